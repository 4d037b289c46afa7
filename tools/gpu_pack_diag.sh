mkdir -p gpurun_out
for pk in 0 1; do
  echo "=== WG_PACK4=$pk timeline"; WG_PACK4=$pk python tools/timeline.py 2>&1 | grep -E "lifetime|tile period|rows with"
  echo "=== WG_PACK4=$pk tile_starts layer 3 (reverse order)"; WG_PACK4=$pk python tools/tile_starts.py 3 2>&1 | tail -3 | cut -c1-150
  echo "=== WG_PACK4=$pk tile_starts layer 2 (forward order)"; WG_PACK4=$pk python tools/tile_starts.py 2 2>&1 | tail -3 | cut -c1-150
done 2>&1 | tee gpurun_out/pack4_diag.txt
