# 8-GPU strong-scaling lines (one box): the headline workload and the human-scale sample.  Outputs in gpurun_out/.
set -x
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-weak > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --no-weak --workload human > gpurun_out/bench_r02_human_${N}gpu.json 2> gpurun_out/bench_r02_human_${N}gpu.log
tail -c 600 gpurun_out/bench_r02_${N}gpu.json; tail -c 600 gpurun_out/bench_r02_human_${N}gpu.json
