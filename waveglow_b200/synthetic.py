"""Deterministic synthetic WaveGlow checkpoints, mels and noise.

There is no network for the NVIDIA checkpoints, so benchmarks and parity tests use random-init
weights of the LJS-v3 architecture in the reference's `state_dict` layout
(`parametrizations.weight.original0/1` keys, /root/reference/src/waveglow/model.py:84-113).
All draws come from an explicit CPU `torch.Generator`, so the same (dims, seed) gives the same
tensors in the build container (where the golden fixtures are made with the unmodified reference)
and on the GPU box.

`WN.end` is zero-initialised in the reference (model.py:88-93), which makes the flow degenerate
(s = b = 0, the WN output is discarded).  As SURVEY.md §8c prescribes, the synthetic checkpoint
overwrites `end.weight, end.bias ~ N(0, 0.02)` so parity tests exercise the WN kernels.
"""
import math
from collections import OrderedDict
from typing import Dict, List

import torch

from .hparams import HParams


def _uniform(gen: torch.Generator, shape, bound: float) -> torch.Tensor:
  return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def _conv_init(gen, out_c, in_c, k):
  # torch Conv1d default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)); same for bias
  bound = 1.0 / math.sqrt(in_c * k)
  return _uniform(gen, (out_c, in_c, k), bound), _uniform(gen, (out_c,), bound)


def flow_channels(hp: HParams) -> List[int]:
  """Live audio channels per flow k (model.py:160-175): 8,8,8,8,6,6,6,6,4,4,4,4 for the defaults."""
  n_rem = hp.n_group
  out = []
  for k in range(hp.n_flows):
    if k % hp.n_early_every == 0 and k > 0:
      n_rem -= hp.n_early_size
    out.append(n_rem)
  return out


def synthetic_state_dict(hp: HParams, seed: int = 1234, end_std: float = 0.02,
                         style: str = "parametrized") -> "OrderedDict[str, torch.Tensor]":
  """state_dict with the reference's keys and shapes.

  style: "parametrized" (original0/original1), "legacy" (weight_g/weight_v) or "folded" (.weight).
  """
  gen = torch.Generator(device="cpu")
  gen.manual_seed(seed)
  sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

  def put_wn(prefix, weight, bias):
    sd[prefix + ".bias"] = bias
    if style == "folded":
      sd[prefix + ".weight"] = weight
      return
    g = weight.flatten(1).norm(dim=1).view(-1, 1, 1)
    # perturb g so that the fold is not the identity
    g = g * (1.0 + 0.1 * _uniform(gen, g.shape, 1.0))
    if style == "legacy":
      sd[prefix + ".weight_g"] = g
      sd[prefix + ".weight_v"] = weight
    else:
      sd[prefix + ".parametrizations.weight.original0"] = g
      sd[prefix + ".parametrizations.weight.original1"] = weight

  n_mel = hp.n_mel_channels
  bound = 1.0 / math.sqrt(n_mel * 1024)
  # ConvTranspose1d weight is [in, out, k]; torch's fan_in for it is out*k (dim 1 * receptive field)
  sd["upsample.weight"] = _uniform(gen, (n_mel, n_mel, 1024), bound)
  sd["upsample.bias"] = _uniform(gen, (n_mel,), bound)
  nc = hp.n_channels
  chans = flow_channels(hp)
  for k in range(hp.n_flows):
    c = chans[k]
    h = c // 2
    p = f"WN.{k}"
    for i in range(hp.n_layers):
      w, b = _conv_init(gen, 2 * nc, nc, hp.kernel_size)
      put_wn(f"{p}.in_layers.{i}", w, b)
    for i in range(hp.n_layers):
      rs = 2 * nc if i < hp.n_layers - 1 else nc
      w, b = _conv_init(gen, rs, nc, 1)
      put_wn(f"{p}.res_skip_layers.{i}", w, b)
    w, b = _conv_init(gen, nc, h, 1)
    put_wn(f"{p}.start", w, b)
    sd[f"{p}.end.weight"] = torch.randn((2 * h, nc, 1), generator=gen) * end_std
    sd[f"{p}.end.bias"] = torch.randn((2 * h,), generator=gen) * end_std
    w, b = _conv_init(gen, 2 * nc * hp.n_layers, n_mel * hp.n_group, 1)
    put_wn(f"{p}.cond_layer", w, b)
  for k in range(hp.n_flows):
    c = chans[k]
    q, _ = torch.linalg.qr(torch.randn((c, c), generator=gen))
    if torch.det(q) < 0:
      q[:, 0] = -q[:, 0]
    # a trained 1x1 weight is no longer orthonormal: perturb so that W^-1 != W^T and log|det W| != 0
    q = q @ (torch.eye(c) + 0.03 * torch.randn((c, c), generator=gen))
    sd[f"convinv.{k}.conv.weight"] = q.contiguous().view(c, c, 1)
  return sd


def synthetic_mel(batch: int, frames: int, seed: int, n_mel: int = 80) -> torch.Tensor:
  """log-mel-like input: N(-5, 1.5^2) clamped to [-11.5, 2] (range of log(clamp(x, 1e-5)), taco_stft.py:10-16)."""
  gen = torch.Generator(device="cpu")
  gen.manual_seed(seed)
  mel = torch.randn((batch, n_mel, frames), generator=gen) * 1.5 - 5.0
  return mel.clamp_(-11.5, 2.0)


def synthetic_noise(hp: HParams, batch: int, n_groups: int, seed: int) -> List[torch.Tensor]:
  """Pre-drawn N(0,1) in the order/shape infer consumes it (model.py:239-243, 266-270)."""
  gen = torch.Generator(device="cpu")
  gen.manual_seed(seed)
  chans = flow_channels(hp)
  zs = [torch.randn((batch, chans[-1], n_groups), generator=gen)]
  for k in reversed(range(hp.n_flows)):
    if k % hp.n_early_every == 0 and k > 0:
      zs.append(torch.randn((batch, hp.n_early_size, n_groups), generator=gen))
  return zs
