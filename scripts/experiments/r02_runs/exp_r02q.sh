set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02q.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02q.log
export DICEGP_VERBOSE=1
for w in 0 8 16; do
DICEGP_LIB=diceseq_b200/libdicegp_sw$w.so timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_sw$w.log 2>&1
done
timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_sw12.log 2>&1
