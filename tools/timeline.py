"""Per-phase clock64 timeline of the fused WN layer kernel (diagnostics; run on the GPU box).

  python tools/timeline.py [--batch 16] [--frames 861] [--channels 256|512] [--layer 3]

Stamps are of the SECOND tile of every leader CTA (steady state).  Slots (wg_layer_pair_cont_tc.cuh): 0 kernel entry | 1 MMA warp
ready | 2 first operands of tile 1 landed | 3 last GEMM1 chunk of tile 1 issued | 4 GEMM2 (tile 1) may start: every gate done |
5 last GEMM2 half issued | 6 first accumulator chunk complete (gate warps see it) | 7 first gate done | 8 last chunk complete |
9 last gate done | 10 first x' half complete | 11 last x' half drained and stored | 12 first operands of tile 2 landed |
13 epilogue warps done | 14 exit.  GEMM2 of tile 1 is issued inside tile 2's first chunk, so 4, 5, 10, 11 lie after 12.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402

NAMES = ["entry", "mma_ready", "t1_first_full", "t1_g1_issued", "t1_all_gated", "t1_g2_issued", "t1_acc0_done", "t1_gate0_done",
         "t1_accL_done", "t1_gateL_done", "t1_d2h0_done", "t1_drained", "t2_first_full", "epilogue_done", "exit"]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=16)
  ap.add_argument("--frames", type=int, default=861)
  ap.add_argument("--channels", type=int, default=256)
  ap.add_argument("--layer", type=int, default=3)
  a = ap.parse_args()
  hp = HParams(n_channels=a.channels)
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
  eng.debug_stop_after(hp.n_flows - 1, a.layer)       # stamps of this layer launch of the first flow remain
  model.infer_with_noise(mel, zs, 0.666)
  torch.cuda.synchronize()
  eng.debug_timeline(None)
  eng.debug_stop_after(-1, 0)
  t = buf.cpu().double()
  t = t[(t[:, 0] > 0) & (t[:, 6] > 0)]           # persistent kernel: one row per leader CTA; phase stamps are of its 2nd tile
  print(f"leader CTAs with a second tile: {t.size(0)}")
  rel = t[:, 1:15] - t[:, 0:1]
  med = rel.median(dim=0).values
  print("median cycles since kernel entry (in slot order; see the docstring for which lie in the next tile):")
  for i, name in enumerate(NAMES[1:]):
    print(f"  {i + 1:2d} {name:16s} {med[i]:10.0f}")
  dur = (t[:, 14] - t[:, 0])
  print(f"CTA lifetime cycles: median {dur.median():.0f}  p10 {dur.quantile(0.1):.0f}  p90 {dur.quantile(0.9):.0f}")
  n_mma1 = (5 + 3 * a.channels // 64) * 4 * (a.channels // 128) if a.layer > 0 else (5 * 4 + 3) * (a.channels // 128)
  if (t[:, 12] > 0).any():
    per = (t[:, 12] - t[:, 2])[t[:, 12] > 0]
    print(f"tile period (first operands of tile 1 -> first operands of tile 2): median {per.median():.0f} cycles "
          f"({n_mma1} GEMM1 MMAs of 128 cycles = {n_mma1 * 128})")
  print(f"GEMM1 issue (t1 first operands -> last chunk issued): median {(t[:, 3] - t[:, 2]).median():.0f}")
  print(f"first gate: {(t[:, 7] - t[:, 6]).median():.0f}   last gate: {(t[:, 9] - t[:, 8]).median():.0f}")
  print(f"last gate done -> GEMM2 may start (MMA warp sees it): {(t[:, 4] - t[:, 9]).median():.0f}")
  print(f"GEMM2 issue (all halves): {(t[:, 5] - t[:, 4]).median():.0f}   first x' half complete -> all drained: {(t[:, 11] - t[:, 10]).median():.0f}")


if __name__ == "__main__":
  main()
