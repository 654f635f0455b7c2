mkdir -p gpurun_out
for args in "0.1 32 1" "1.0 16 1" "1.0 32 1" "1.0 32 0"; do
  echo "== $args"; ( timeout 40 python scripts/dbg_k32.py $args ) 2>&1 | grep -v '^\[ma\]' | cut -c1-200
done
