set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02w.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02w.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_w.log 2>&1
DICEGP_WIDE_CS16_BYTES=0 timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_w0.log 2>&1
DICEGP_WIDE_CS16_BYTES=16777216 timeout 300 python scripts/quick_timing.py human 3000 > gpurun_out/qhuman_w16.log 2>&1
