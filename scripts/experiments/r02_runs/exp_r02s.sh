set -x
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_s4.log 2>&1
DICEGP_STAGE_MAX=8 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_s8.log 2>&1
DICEGP_STAGE_MAX=20 timeout 300 python scripts/quick_timing.py timeseries 2500 > gpurun_out/q2500_s20.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 2500 > gpurun_out/q2500_s4.log 2>&1
