set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02v.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02v.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_v.log 2>&1
DICEGP_WIDE_BIG=0 timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_v0.log 2>&1
timeout 300 python scripts/quick_timing.py human 375 > gpurun_out/qhuman375_v.log 2>&1
DICEGP_WIDE_BIG=0 timeout 300 python scripts/quick_timing.py human 375 > gpurun_out/qhuman375_v0.log 2>&1
