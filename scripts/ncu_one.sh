# One `ncu --set full` capture of a chain-kernel launch of a 2500-gene batch, exported to CSV on the box (the .ncu-rep
# stays there: gpurun_out/ is capped at 64 MiB).  usage: NAME=c8 KERN=dicegp_chain_kernel SKIP=6 bash scripts/ncu_one.sh
set -x
NAME=${NAME:-c8}; KERN=${KERN:-dicegp_chain_kernel}; SKIP=${SKIP:-6}; GENES=${GENES:-2500}
REP=/tmp/prof_$NAME
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERN -s $SKIP -c 1 -f -o $REP \
  python bench.py --genes $GENES --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full_$NAME.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$NAME.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$NAME.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_$NAME.csv 60 > gpurun_out/ncu_lines_$NAME.txt 2>&1
ncu -i $REP.ncu-rep --page source --csv --print-source sass > /tmp/sass_$NAME.csv 2>/dev/null
gzip -c /tmp/sass_$NAME.csv > gpurun_out/ncu_sass_$NAME.csv.gz
ls -la gpurun_out/*$NAME* /tmp/*$NAME*
