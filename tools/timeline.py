"""Per-phase clock64 timeline of the fused WN layer kernel (diagnostics; run on the GPU box).

  python tools/timeline.py [--batch 16] [--frames 861]

Slots: 0 entry | 1 MMA thread ready | 2 first operands landed | 3 G1a issued | 4 G1b issued | 5 acc a done (epilogue sees it)
| 6 gate a done | 7 acc b done | 8 gate b done | 9 MMA sees acts | 10 G2 issued | 11 D2 done | 12 x' staged | 13 x' stored | 14 exit | 15 first operands of tile 2 (continuous
pair kernel only: there GEMM2 of tile 1 is issued inside tile 2's chunk a, so 9..12 lie after 15)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402

NAMES = ["entry", "mma_ready", "first_full", "g1a_issued", "g1b_issued", "acc_a_done", "gate_a_done", "acc_b_done",
         "gate_b_done", "mma_sees_acts", "g2_issued", "d2_done", "x_staged", "x_stored", "exit"]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=16)
  ap.add_argument("--frames", type=int, default=861)
  a = ap.parse_args()
  hp = HParams()
  dev = torch.device("cuda:0")
  model = WaveGlow(hp).to(dev)
  model.load_state_dict(synthetic_state_dict(hp, 1234))
  model = WaveGlow.remove_weightnorm(model).eval()
  mel = synthetic_mel(a.batch, a.frames, 1).to(dev)
  zs = [z.to(dev) for z in synthetic_noise(hp, a.batch, a.frames * 32, 2)]
  eng = model.engine()
  for _ in range(2):
    model.infer_with_noise(mel, zs, 0.666)
  n_cta = 148
  buf = torch.zeros((n_cta, 16), dtype=torch.int64, device=dev)
  eng.debug_timeline(buf)
  eng.debug_stop_after(hp.n_flows - 1, 3)       # stamps of the 4th layer launch (dilation 8) remain
  model.infer_with_noise(mel, zs, 0.666)
  torch.cuda.synchronize()
  eng.debug_timeline(None)
  eng.debug_stop_after(-1, 0)
  t = buf.cpu().double()
  t = t[(t[:, 0] > 0) & (t[:, 5] > 0)]           # persistent kernel: one row per CTA; phase stamps are of its 2nd tile
  n_tiles = a.batch * 32 * ((a.frames + 127) // 128)
  print(f"rows with stamps: {t.size(0)}; tiles {n_tiles}; per-CTA tiles ~ {n_tiles / max(1, t.size(0)):.2f}")
  rel = t[:, 1:15] - t[:, 0:1]
  med = rel.median(dim=0).values
  print(f"CTAs {n_cta}; median cycles since kernel entry of each CTA:")
  prev = 0.0
  for i, name in enumerate(NAMES[1:]):
    print(f"  {name:14s} {med[i]:10.0f}   (+{med[i] - prev:8.0f})")
    prev = float(med[i])
  dur = (t[:, 14] - t[:, 0])
  print(f"CTA lifetime cycles: median {dur.median():.0f}  p10 {dur.quantile(0.1):.0f}  p90 {dur.quantile(0.9):.0f}")
  span = t[:, 11] - t[:, 2]
  print(f"steady-state tile: first operands of G1a -> D2 done: median {span.median():.0f} cycles")
  if (t[:, 15] > 0).any():
    per = (t[:, 15] - t[:, 2])[t[:, 15] > 0]
    print(f"tile period (first operands of tile 1 -> first operands of tile 2): median {per.median():.0f} cycles")
  g1a = t[:, 5] - t[:, 2]
  g1b = t[:, 7] - t[:, 5]
  print(f"G1a (first operands -> acc a done): median {g1a.median():.0f} cycles = {g1a.median() / 68:.1f} per UTCHMMA (ideal 128)")
  print(f"G1b (acc a done -> acc b done):     median {g1b.median():.0f} cycles = {g1b.median() / 68:.1f} per UTCHMMA")


if __name__ == "__main__":
  main()
