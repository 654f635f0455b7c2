mkdir -p gpurun_out
for o in "refill_at=1" "clip_a=1,clip_b=1000" "persist_waves=1" "persist_min_chunk=1024" "rtree=6"; do
  echo "== $o"; ( MA_OPTS2=$o timeout 30 python scripts/dbg_k32.py 1.0 32 1 ) 2>&1 | grep -v '^\[ma\]' | tail -2 | cut -c1-200
done
