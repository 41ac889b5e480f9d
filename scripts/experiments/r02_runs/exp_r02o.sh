set -x
export DICEGP_VERBOSE=1
for v in "" _j5 _j6; do
L=diceseq_b200/libdicegp$v.so
DICEGP_LIB=$L timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_nodes$v.log 2>&1
done
