"""One small `infer` per channel count and one `forward`, meant to run under `compute-sanitizer --tool memcheck`
(tools/gpu_step.sh <tag> sanitize): ragged lengths, packed tails, chained launches with tile flags, exact carry (512)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402


def main():
  for nc, B, F in ((256, 3, 139), (512, 2, 133)):
    hp = HParams(n_channels=nc)
    m = WaveGlow(hp, precision="fp16").to("cuda:0")
    m.load_state_dict(synthetic_state_dict(hp, 1234))
    m = WaveGlow.remove_weightnorm(m).eval()
    mel = synthetic_mel(B, F, 5).cuda()
    zs = [z.cuda() for z in synthetic_noise(hp, B, F * 32, 6)]
    out = m.infer_with_noise(mel, zs, 0.666)
    torch.cuda.synchronize()
    print(f"{nc} ch infer B{B} x F{F}: finite {bool(torch.isfinite(out).all())} absmax {float(out.abs().max()):.3f}")
    if nc == 256:
      audio = torch.rand((B, 16000), device="cuda") * 2 - 1
      z, log_s, log_det = m((synthetic_mel(B, 63, 7).cuda(), audio))
      torch.cuda.synchronize()
      print(f"{nc} ch forward B{B} x 16000: finite {bool(torch.isfinite(z).all())}")


if __name__ == "__main__":
  main()
