set -x
export DICEGP_VERBOSE=1
export DICEGP_LIB=diceseq_b200/libdicegp_w512.so
timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_w512.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_w512.log 2>&1
DICEGP_WIDE_NPW=1 timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_w512_p1.log 2>&1
DICEGP_WIDE_NPW=1 timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_w512_p1.log 2>&1
