"""Per-launch time of the first flow's kernels (flow step, then WN layers 0..7) from differences of truncated runs
(wg_debug_stop_after), CUDA events, config 2 by default.  Diagnostics; run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 861
NC = int(sys.argv[3]) if len(sys.argv) > 3 else 256
hp = HParams(n_channels=NC)
dev = torch.device("cuda:0")
model = WaveGlow(hp).to(dev)
model.load_state_dict(synthetic_state_dict(hp, 1234))
model = WaveGlow.remove_weightnorm(model).eval()
mel = synthetic_mel(B, F, 1).to(dev)
zs = [z.to(dev) for z in synthetic_noise(hp, B, F * 32, 2)]
eng = model.engine()
k = hp.n_flows - 1


def timed(stop_layer, reps=20):
  eng.debug_stop_after(k, stop_layer)
  for _ in range(3):
    model.infer_with_noise(mel, zs, 0.666)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    model.infer_with_noise(mel, zs, 0.666)
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps * 1e3


prev = None
names = ["noise+melwin+flowstep"] + [f"layer {i}" for i in range(hp.n_layers)] + ["flow step (finish+start)"]
for idx, stop in enumerate(list(range(-1, hp.n_layers)) + [hp.n_layers]):
  t = timed(stop)
  print(f"{names[idx]:28s} cumulative {t:8.1f} us   delta {t - (prev or 0.0):8.1f} us")
  prev = t
eng.debug_stop_after(-1, 0)
