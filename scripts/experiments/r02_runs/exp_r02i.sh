set -x
export DICEGP_VERBOSE=1
for cap in 512 768 1536; do
DICEGP_CAP0=$cap timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_cap$cap.log 2>&1
done
for G in 2500 5000; do for cap in 1024 2048 4096; do
DICEGP_CAP0=$cap timeout 300 python scripts/quick_timing.py timeseries $G > gpurun_out/q${G}_cap$cap.log 2>&1
done; done
DICEGP_CAP0=1024 timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_cap1024.log 2>&1
