for C in 8 7 6 5; do
  python scripts/stream_sweep.py $C 1184 2>&1 | tail -1
  DICEGP_TIER_THREADS=64,128,256,512,512 DICEGP_TIER_BPS=8,4,2,1,1 python scripts/stream_sweep.py $C 1184 2>&1 | tail -1
  DICEGP_TIER_THREADS=64,128,256,512,384 DICEGP_TIER_BPS=8,4,2,1,1 python scripts/stream_sweep.py $C 1184 2>&1 | tail -1
done
