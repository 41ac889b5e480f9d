set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02f.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02f.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_f.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_f.log 2>&1
for C in 4 8; do DICEGP_LIB=diceseq_b200/libdicegp_prof.so python scripts/prof_latency.py $C 148 > gpurun_out/widelat_f_$C.log 2>&1; done
