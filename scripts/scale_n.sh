# Strong-scaling bench line on N GPUs of one box (headline workload).  usage: N=4 bash scripts/scale_n.sh
set -x
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-weak > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.log
tail -c 300 gpurun_out/bench_r02_${N}gpu.json
