set -x
export DICEGP_VERBOSE=1
DICEGP_END_MODE=2 DICEGP_WIDE_CS=2 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_cs2end.log 2>&1
DICEGP_END_MODE=2 DICEGP_WIDE_CS=4 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_cs4end.log 2>&1
DICEGP_END_MODE=2 DICEGP_WIDE_CS=2 timeout 300 python scripts/quick_timing.py timeseries 2500 > gpurun_out/q2500_cs2end.log 2>&1
