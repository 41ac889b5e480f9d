set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02l.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02l.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_l.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_l.log 2>&1
timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_l.log 2>&1
