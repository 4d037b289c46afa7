"""Where does Synthesizer construction time go?  Run on the GPU box."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.hparams import HParams
from waveglow_b200.model import WaveGlow
from waveglow_b200.synthetic import synthetic_state_dict
from waveglow_b200.checkpoint import CheckpointWaveglow
from waveglow_b200.synthesizer import Synthesizer

def tick(label, t0):
  torch.cuda.synchronize()
  print(f"{label:40s} {time.perf_counter() - t0:7.3f} s", flush=True)
  return time.perf_counter()

hp = HParams()
sd = synthetic_state_dict(hp, 1234)
torch.zeros(1).cuda()
t = time.perf_counter()
m = WaveGlow(hp); t = tick("WaveGlow(hparams) (random init)", t)
m = m.to("cuda:0"); t = tick(".to(cuda)", t)
m.load_state_dict(sd); t = tick("load_state_dict", t)
m = WaveGlow.remove_weightnorm(m).eval(); t = tick("remove_weightnorm", t)
fw = m.folded_weights(); t = tick("folded_weights()", t)
eng = m.engine(); t = tick("engine(): wg_create + wg_load_weights", t)
mel = torch.zeros(1, 80, 88, device="cuda:0")
m.infer(mel, sigma=0.0); t = tick("first infer (bias mel)", t)
ck = CheckpointWaveglow.from_state_dict(sd, hp)
t = time.perf_counter()
s = Synthesizer(ck, device=torch.device("cuda:0")); t = tick("Synthesizer(checkpoint) total", t)
