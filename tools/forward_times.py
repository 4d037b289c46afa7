"""Per-call CUDA-event times of the forward direction at BASELINE config 5 (24 x 16000 samples) -- diagnostics."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_state_dict  # noqa: E402

hp = HParams()
dev = torch.device("cuda:0")
model = WaveGlow(hp).to(dev)
model.load_state_dict(synthetic_state_dict(hp, 1234))
model = WaveGlow.remove_weightnorm(model).eval()
B, T, F = 24, 16000, 63
mel = synthetic_mel(B, F, 3).to(dev)
audio = (torch.rand(B, T, generator=torch.Generator().manual_seed(4)) * 2 - 1).to(dev)
eng = model.engine()
if "--after-big" in sys.argv:      # a big infer first (workspace sized for config 2), as tools/measure_configs.py does
  from waveglow_b200.synthetic import synthetic_noise
  m2 = synthetic_mel(16, 861, 1).to(dev)
  z2 = [z.to(dev) for z in synthetic_noise(hp, 16, 861 * 32, 2)]
  for _ in range(3):
    model.infer_with_noise(m2, z2, 0.666)
  torch.cuda.synchronize()
halves = [4, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2]
ts = []
for i in range(24):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  eng.forward(mel, audio, halves, want_log_s=(i % 2 == 0))
  e1.record()
  torch.cuda.synchronize()
  ts.append(e0.elapsed_time(e1))
print("env", {k: v for k, v in os.environ.items() if k.startswith("WG_")}, "launches", eng.launch_count())
print("ms per call:", " ".join(f"{t:.2f}" for t in ts))
ts = []
for i in range(12):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  model((mel, audio))
  e1.record()
  torch.cuda.synchronize()
  ts.append(e0.elapsed_time(e1))
print("through model.forward:", " ".join(f"{t:.2f}" for t in ts))
