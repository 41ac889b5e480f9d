# Round-2 profile of the shipped build: per-phase cycle counts (DG_PROFILE build), then `ncu --set full` captures of
# three pass-0 launches of a 2500-gene batch (C = 8, 5, 2) and of the straggler-pass wide kernel.  Outputs in gpurun_out/.
set -x
TAG=${TAG:-r02}
P=diceseq_b200/libdicegp_prof_n4.so
[ -f $P ] && DICEGP_LIB=$P python scripts/prof_phases.py big > gpurun_out/phases_$TAG.log 2>&1
python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/plain2500_$TAG.json 2> gpurun_out/plain2500_$TAG.log || exit 1
for spec in "c8:dicegp_chain_kernel:6" "c5:dicegp_chain_kernel:3" "c2:dicegp_chain_kernel:0" "wide8:dicegp_wide_kernel:0"; do
  IFS=: read name kern skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name \
    python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full_${TAG}_$name.log 2>&1
done
ls -la gpurun_out/ | tail -8
