mkdir -p gpurun_out
python tools/layer_times.py 2>&1 | tail -12 | tee gpurun_out/layer_times_v9.txt
for g in 4 5 6; do echo "== WG_G2AT0=$g"; WG_G2AT0=$g python tools/layer_times.py 2>&1 | grep "layer 0"; done | tee -a gpurun_out/layer_times_v9.txt
echo "== tile period layer 0"; python tools/tile_starts.py 0 2>&1 | tail -3 | cut -c1-150 | tee -a gpurun_out/layer_times_v9.txt
echo "== B1 F200"; python tools/layer_times.py 1 200 2>&1 | tail -12 | tee -a gpurun_out/layer_times_v9.txt
