set -x
export DICEGP_VERBOSE=1
for cap in 2048 3072 4096; do
DICEGP_CAP0=$cap timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_u$cap.log 2>&1
done
for cap in 1024 2048; do
DICEGP_CAP0=$cap timeout 300 python scripts/quick_timing.py timeseries 5000 > gpurun_out/q5000_u$cap.log 2>&1
done
for cap in 512 1024; do
DICEGP_CAP0=$cap timeout 300 python scripts/quick_timing.py timeseries 2500 > gpurun_out/q2500_u$cap.log 2>&1
done
