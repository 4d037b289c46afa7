#!/bin/bash
# Diagnostics: tile period of the continuous pair kernel with parts switched off (results invalid; which part paces the tile?)
mkdir -p gpurun_out
for f in ${FLAGS:-0 16 4 8 128 1 20 28 29}; do
  echo "=== WG_DBG_FLAGS=$f"
  WG_DBG_FLAGS=$f timeout 300 python tools/tile_starts.py 3 2>&1 | tail -3 | cut -c1-140
done | tee gpurun_out/flags_${TAG:-v8}.txt
