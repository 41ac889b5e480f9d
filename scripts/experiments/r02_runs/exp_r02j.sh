set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02j.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02j.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_new.log 2>&1
DICEGP_CAP0=4096 timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_4096.log 2>&1
timeout 300 python scripts/quick_timing.py human 375 > gpurun_out/qhuman375_new.log 2>&1
DICEGP_CAP0=4096 timeout 300 python scripts/quick_timing.py human 375 > gpurun_out/qhuman375_4096.log 2>&1
