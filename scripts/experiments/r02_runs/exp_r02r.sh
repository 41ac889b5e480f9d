set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02r.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest_r02r.log
export DICEGP_VERBOSE=1
timeout 300 python scripts/quick_timing.py timeseries 10000 > gpurun_out/q10000_r.log 2>&1
timeout 300 python scripts/quick_timing.py timeseries 1250 > gpurun_out/q1250_r.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r.log 2>&1
