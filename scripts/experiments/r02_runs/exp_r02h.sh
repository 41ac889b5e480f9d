set -x
export DICEGP_VERBOSE=1
for cap in 1024 2048 3072; do
DICEGP_CAP0=$cap timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_cap$cap.log 2>&1
done
DICEGP_CAP0=2048 timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_cap2048.log 2>&1
DICEGP_CAP0=3072 timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_cap3072.log 2>&1
