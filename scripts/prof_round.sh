# Round profile: plain run first (must exit 0), then the ncu launch list of the same command, then one
# `--set full` capture of the C = 8 streamed pass-0 launch on a 2500-gene batch (FULL=1).  Outputs in gpurun_out/.
set -x
TAG=${TAG:-r01b}
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.log || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
if [ "${FULL:-0}" = "1" ]; then
  python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2500_$TAG.json 2> gpurun_out/plain2500_$TAG.log || exit 1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dicegp_chain_kernel -s 6 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
fi
ls -la gpurun_out/ | tail -6
