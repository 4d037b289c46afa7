"""Times the non-headline BASELINE.json configs on one GPU (CUDA events, warm engine) -> JSON lines.

  python tools/measure_configs.py            # configs 1 (B1 x 200 frames), 4 (B1 x 60 s), 5 (forward B24 x 16000 samples)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402


def timed(fn, reps):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  ts.sort()
  return ts[len(ts) // 2], ts[0]


def main():
  hp = HParams()
  dev = torch.device("cuda:0")
  model = WaveGlow(hp).to(dev)
  model.load_state_dict(synthetic_state_dict(hp, 1234))
  model = WaveGlow.remove_weightnorm(model).eval()
  for name, B, F in (("config1_b1_f200", 1, 200), ("config4_b1_60s", 1, 5168), ("config2_b16_f861", 16, 861)):
    mel = synthetic_mel(B, F, 1).to(dev)
    zs = [z.to(dev) for z in synthetic_noise(hp, B, F * 32, 2)]
    med, best = timed(lambda: model.infer_with_noise(mel, zs, 0.666), 10)
    n = B * F * 256
    out = model.infer_with_noise(mel, zs, 0.666)
    print(json.dumps({"config": name, "direction": "infer", "batch": B, "frames": F, "samples": n, "median_ms": med, "best_ms": best,
                      "samples_per_s": n / (med * 1e-3), "x_realtime": n / (med * 1e-3) / 22050,
                      "finite": bool(torch.isfinite(out).all()), "launches": model.engine().launch_count()}), flush=True)
  # config 1 again through the CUDA-graph replay path
  B, F = 1, 200
  mel = synthetic_mel(B, F, 1).to(dev)
  zs = [z.to(dev) for z in synthetic_noise(hp, B, F * 32, 2)]
  model.use_cuda_graphs = True
  med, best = timed(lambda: model.infer_with_noise(mel, zs, 0.666), 10)
  model.use_cuda_graphs = False
  n = B * F * 256
  print(json.dumps({"config": "config1_b1_f200_cuda_graph", "direction": "infer", "batch": B, "frames": F, "samples": n,
                    "median_ms": med, "best_ms": best, "samples_per_s": n / (med * 1e-3), "x_realtime": n / (med * 1e-3) / 22050}), flush=True)
  B, T, F = 24, 16000, 63
  mel = synthetic_mel(B, F, 3).to(dev)
  audio = (torch.rand(B, T, generator=torch.Generator().manual_seed(4)) * 2 - 1).to(dev)
  med, best = timed(lambda: model((mel, audio)), 10)
  z, _, _ = model((mel, audio))
  print(json.dumps({"config": "config5_forward_b24_t16000", "direction": "forward", "batch": B, "samples": B * T, "median_ms": med,
                    "best_ms": best, "samples_per_s": B * T / (med * 1e-3), "finite": bool(torch.isfinite(z).all()),
                    "launches": model.engine().launch_count()}), flush=True)


if __name__ == "__main__":
  main()
