set -x
export DICEGP_VERBOSE=1
for G in 1250 2500 10000; do
timeout 300 python scripts/quick_timing.py timeseries $G > gpurun_out/q${G}_desc.log 2>&1
DICEGP_ORDER=1 timeout 300 python scripts/quick_timing.py timeseries $G > gpurun_out/q${G}_asc.log 2>&1
done
