set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02e.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02e.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_new.log 2>&1
DICEGP_WIDE_CS=1 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_cs1.log 2>&1
DICEGP_STRAGGLER_STEPS=4096 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_old.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 2500 > gpurun_out/q2500_new.log 2>&1
DICEGP_STRAGGLER_STEPS=4096 timeout 300 python scripts/quick_timing.py timeseries 2500 > gpurun_out/q2500_old.log 2>&1
tail -3 gpurun_out/gputest_r02e.log; grep -h "pass\|rep 1" gpurun_out/q*.log | grep -v "launch"
