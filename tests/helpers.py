"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import functools
import os

import numpy as np
import torch

from oracle import waveglow_oracle as O
from waveglow_b200.hparams import HParams
from waveglow_b200.synthetic import synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity gates (BASELINE.json north_star / SURVEY.md §8d)
FP32_MAX_ABS = 1e-5
FP16_MAX_ABS = 2e-3
FP16_MIN_SNR_DB = 40.0


def load_golden(name):
  d = np.load(os.path.join(GOLDEN, name + ".npz"))
  out = {k: d[k] for k in d.files}
  return out


def golden_noise(g):
  zs = []
  i = 0
  while f"z{i}" in g:
    zs.append(torch.from_numpy(g[f"z{i}"]))
    i += 1
  return zs


@functools.lru_cache(maxsize=4)
def synthetic_weights(n_channels: int, style: str = "parametrized", seed: int = 1234):
  hp = HParams(n_channels=n_channels)
  sd = synthetic_state_dict(hp, seed, style=style)
  return hp, sd, O.fold_weight_norm(sd), O.ModelDims(n_channels=n_channels)


def snr_db(ref: torch.Tensor, out: torch.Tensor) -> float:
  ref = ref.double()
  err = out.double() - ref
  return float(10 * torch.log10((ref ** 2).sum() / (err ** 2).sum().clamp_min(1e-300)))


def max_abs(ref: torch.Tensor, out: torch.Tensor) -> float:
  return float((out.double() - ref.double()).abs().max())
