"""Start cycle of the first 16 tiles of every CTA pair of one layer launch (WG_DBG_FLAGS=64; diagnostics, run on the GPU box)."""
import os
import sys

import torch

os.environ["WG_DBG_FLAGS"] = str(int(os.environ.get("WG_DBG_FLAGS", "0")) | 64)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.model import WaveGlow  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402

hp = HParams()
dev = torch.device("cuda:0")
model = WaveGlow(hp).to(dev)
model.load_state_dict(synthetic_state_dict(hp, 1234))
model = WaveGlow.remove_weightnorm(model).eval()
B, F = 16, 861
mel = synthetic_mel(B, F, 1).to(dev)
zs = [z.to(dev) for z in synthetic_noise(hp, B, F * 32, 2)]
eng = model.engine()
for _ in range(2):
  model.infer_with_noise(mel, zs, 0.666)
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
buf = torch.zeros((148, 16), dtype=torch.int64, device=dev)
eng.debug_timeline(buf)
eng.debug_stop_after(hp.n_flows - 1, layer)
model.infer_with_noise(mel, zs, 0.666)
torch.cuda.synchronize()
eng.debug_timeline(None)
eng.debug_stop_after(-1, 0)
t = buf.cpu().double()
t = t[t[:, 1] > 0]
d = t[:, 1:] - t[:, :-1]
print(f"layer {layer}: rows {t.size(0)}; median tile period by tile index (cycles):")
print("  " + " ".join(f"{float(x):7.0f}" for x in d.median(dim=0).values))
print("  p90 " + " ".join(f"{float(x):7.0f}" for x in d.quantile(0.9, dim=0)))
