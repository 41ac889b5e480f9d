# Round-2 profile of the shipped build.  Outputs in gpurun_out/ (the .ncu-rep files stay on the box: gpurun_out/ is
# capped at 64 MiB; the raw-metric and per-source-line CSV exports are what comes back).
#   1. plain bench (must exit 0) + reference arm, 2. ncu launch list of the same command,
#   3. `ncu --set full` captures of one launch each on a 2500-gene batch: pass-0 chain kernel for C = 8 and C = 5,
#      the wide-round kernel (straggler pass, C = 8), the summary kernel (the short-chain class).
set -x
TAG=${TAG:-r02}
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.log || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/plain2500_$TAG.json 2> gpurun_out/plain2500_$TAG.log || exit 1
# 2500 genes -> pass 0 capped at 1024 steps; launches of a batch: 7 pre-run, 1 var, 7 chain (C = 8 first), tile, 7 wide, 4-5 summary
GENES=10000 NAME=${TAG}_chain_c8 KERN=dicegp_chain_kernel SKIP=0 bash scripts/ncu_one.sh > gpurun_out/ncu_one_${TAG}_c8.log 2>&1
NAME=${TAG}_chain_c5 KERN=dicegp_chain_kernel SKIP=3 bash scripts/ncu_one.sh > gpurun_out/ncu_one_${TAG}_c5.log 2>&1
GENES=10000 NAME=${TAG}_wide_c8 KERN=dicegp_wide_kernel SKIP=0 bash scripts/ncu_one.sh > gpurun_out/ncu_one_${TAG}_w8.log 2>&1
NAME=${TAG}_summary KERN=dicegp_summary_kernel SKIP=3 bash scripts/ncu_one.sh > gpurun_out/ncu_one_${TAG}_sum.log 2>&1
rm -f gpurun_out/ncu_sass_*.csv.gz
ls -la gpurun_out/ | tail -30
