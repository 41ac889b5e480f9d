# launch shape of the streamed tier for one isoform count: default planner vs forced (threads, chains per SM)
for C in ${CS:-2 3 4}; do
  python scripts/stream_sweep.py $C 1184 2>&1 | tail -1
  for shape in ${SHAPES:-64,8 96,5 128,4 160,3}; do
    thr=${shape%,*}; bps=${shape#*,}
    DICEGP_TIER_THREADS=64,128,256,512,$thr DICEGP_TIER_BPS=8,4,2,1,$bps python scripts/stream_sweep.py $C 1184 2>&1 | tail -1
  done
done
