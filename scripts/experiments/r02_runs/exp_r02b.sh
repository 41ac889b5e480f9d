# Round-2 experiment: unified warps + cyclic ring (lib A), cyclic ring only (A0), one-word items (B): parity, then
# per-isoform-count timings over launch shapes and L2 pinning fractions.  Outputs in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02b.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02b.log
P='DICEGP_PLAN_K{C}'
python scripts/class_timing.py 2,3,4,5,6,7,8 1430 "$P=96,4" "$P=128,3" "$P=128,4" "$P=160,3" "$P=192,2" "$P=224,2" "$P=256,2" \
   "DICEGP_L2_PIN=0,0,8,8,8,8,8" "DICEGP_L2_PIN=0,0,5,5,5,5,5" "DICEGP_L2_PIN=8,8,11,11,11,11,11" > gpurun_out/class_A.log 2>&1
DICEGP_LIB=diceseq_b200/libdicegp_sw0.so python scripts/class_timing.py 2,3,4,5,6,7,8 1430 "$P=96,4" "$P=128,4" "$P=224,2" "$P=256,2" \
   "DICEGP_L2_PIN=0,0,8,8,8,8,8" > gpurun_out/class_A0.log 2>&1
DICEGP_LIB=diceseq_b200/libdicegp_iw1.so python scripts/class_timing.py 2,3,4,5,6,7,8 1430 "$P=96,4" "$P=128,4" "$P=128,3" "$P=192,2" "$P=256,2" \
   "DICEGP_L2_PIN=0,0,8,8,8,8,8" > gpurun_out/class_B.log 2>&1
DICEGP_VERBOSE=1 python scripts/quick_timing.py timeseries 10000 > gpurun_out/quick_A.log 2>&1
DICEGP_LIB=diceseq_b200/libdicegp_prof.so python scripts/prof_phases.py big > gpurun_out/phases_r02b.log 2>&1
tail -3 gpurun_out/gputest_r02b.log; grep -v "^\[dicegp\]" gpurun_out/class_A.log | tail -80
