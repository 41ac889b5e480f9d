# Round-2 exploration: baseline timings, per-phase cycle counts, forced cluster tiers (development aid).
set -x
P=diceseq_b200/libdicegp_prof_n4.so
DICEGP_VERBOSE=1 python scripts/quick_timing.py timeseries 2500 20000 > gpurun_out/x_base.log 2>&1
DICEGP_LIB=$P python scripts/prof_latency.py 8 148 > gpurun_out/x_lat8.log 2>&1
DICEGP_LIB=$P python scripts/prof_latency.py 4 148 > gpurun_out/x_lat4.log 2>&1
DICEGP_LIB=$P python scripts/prof_latency.py 2 148 > gpurun_out/x_lat2.log 2>&1
DICEGP_LIB=$P python scripts/prof_phases.py big > gpurun_out/x_phases.log 2>&1
DICEGP_LIB=$P DICEGP_WIDE_MIN_BYTES=120000 python scripts/prof_latency.py 8 37 > gpurun_out/x_lat8_cs4.log 2>&1
DICEGP_LIB=$P DICEGP_WIDE_MIN_BYTES=30000 python scripts/prof_latency.py 8 18 > gpurun_out/x_lat8_cs8.log 2>&1
DICEGP_LIB=$P DICEGP_WIDE_MIN_BYTES=120000 python scripts/prof_latency.py 4 74 > gpurun_out/x_lat4_cs2.log 2>&1
DICEGP_VERBOSE=1 DICEGP_WIDE_MIN_BYTES=120000 python scripts/quick_timing.py timeseries 2500 20000 > gpurun_out/x_cluster.log 2>&1
tail -n 8 gpurun_out/x_*.log
