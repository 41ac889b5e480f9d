set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02n.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02n.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_n.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_n.log 2>&1
DICEGP_CAP0=4096 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_n_staged.log 2>&1
timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_n.log 2>&1
