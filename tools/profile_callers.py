"""B200 timings of the CALLER rows around the hot path: `Synthesizer.__init__`,
`Synthesizer.infer` (inference / denoising / device->host), and the batched `synthesize` CLI in files per second.

  python tools/profile_callers.py            # JSON lines; run on the GPU box
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveglow_b200.checkpoint import CheckpointWaveglow  # noqa: E402
from waveglow_b200.cli import main as cli_main  # noqa: E402
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.synthesizer import Synthesizer  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_state_dict  # noqa: E402


def main():
  hp = HParams()
  dev = torch.device("cuda:0")
  ck = CheckpointWaveglow.from_state_dict(synthetic_state_dict(hp, 1234), hp, iteration=1)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  synth = Synthesizer(ck, device=dev)
  torch.cuda.synchronize()
  print(json.dumps({"what": "Synthesizer.__init__ (module tree + fold + pack + denoiser bias inference)", "seconds": time.perf_counter() - t0}), flush=True)
  for name, F in (("config1_b1_f200", 200), ("config4_b1_f5168_60s", 5168), ("b1_f861_10s", 861)):
    mel = synthetic_mel(1, F, 11).to(dev)
    for graphs in (False, True):
      synth.model.use_cuda_graphs = graphs
      for _ in range(3):
        synth.infer(mel, sigma=0.666, denoiser_strength=0.0005, seed=1)
      runs = []
      for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = synth.infer(mel, sigma=0.666, denoiser_strength=0.0005, seed=1)
        total = time.perf_counter() - t0
        runs.append((total, r.inference_duration_s, r.denoising_duration_s))
      runs.sort()
      total, inf, den = runs[len(runs) // 2]
      n = F * 256
      print(json.dumps({"what": "Synthesizer.infer", "shape": name, "cuda_graphs": graphs, "samples": n, "total_ms": total * 1e3,
                        "inference_ms": inf * 1e3, "denoising_ms": den * 1e3, "other_ms (seeding, noise draw sync, D2H, numpy)": (total - inf - den) * 1e3,
                        "denoiser_share": den / total, "x_realtime_end_to_end": n / total / 22050}), flush=True)
    synth.model.use_cuda_graphs = False
  # batched CLI: 64 mels x 10 s
  with tempfile.TemporaryDirectory() as tmp:
    ckpt = os.path.join(tmp, "1.pt")
    ck.save(ckpt)
    folder = os.path.join(tmp, "mels")
    os.makedirs(folder)
    for i in range(64):
      np.save(os.path.join(folder, f"u{i:02d}.npy"), synthetic_mel(1, 861, 100 + i)[0].numpy())
    for bs in (1, 16):
      out = os.path.join(tmp, f"wavs{bs}")
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      rc = cli_main(["synthesize", ckpt, folder, "-out", out, "--custom-seed", "3", "--sigma", "0.666", "--batch-size", str(bs)])
      dt = time.perf_counter() - t0
      print(json.dumps({"what": "cli synthesize, 64 mels x 861 frames (10 s each), incl. checkpoint load + engine start", "batch_size": bs,
                        "rc": rc, "seconds": dt, "files_per_s": 64 / dt, "x_realtime": 64 * 861 * 256 / 22050 / dt}), flush=True)


if __name__ == "__main__":
  main()
