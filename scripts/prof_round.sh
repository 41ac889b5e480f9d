set -x
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_r01b.json 2> gpurun_out/plain_r01b.log || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_r01b.log 2>&1
python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2500_r01b.json 2> gpurun_out/plain2500_r01b.log || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dicegp_chain_kernel -s 13 -c 1 -f -o gpurun_out/prof_r01b python bench.py --genes 2500 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r01b.log 2>&1
ls -la gpurun_out/ | tail -8
