# Round-2 step: new summary kernel + planner: parity, bench, wide-kernel phase profile and shapes.  Outputs in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02c.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02c.log
DICEGP_VERBOSE=1 python bench.py --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.log
P=diceseq_b200/libdicegp_prof.so
for C in 2 4 8; do DICEGP_LIB=$P python scripts/prof_latency.py $C 148 > gpurun_out/widelat_$C.log 2>&1; done
for C in 2 4 8; do DICEGP_WIDE_NCW=8 DICEGP_WIDE_BPS=2 python scripts/prof_latency.py $C 296 > gpurun_out/widelat_bps2_$C.log 2>&1; done
for C in 2 4 8; do python scripts/prof_latency.py $C 296 > gpurun_out/widelat_bps1_$C.log 2>&1; done
tail -3 gpurun_out/gputest_r02c.log; tail -n 3 gpurun_out/widelat_*.log
