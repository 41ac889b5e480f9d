set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02d.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02d.log
DICEGP_VERBOSE=1 python bench.py --no-cpu-baseline > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.log
tail -3 gpurun_out/gputest_r02d.log
