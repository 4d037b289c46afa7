"""CPU oracle for the WaveGlow inverse/forward flow  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A restatement, in plain torch fp32 CPU ops, of the reference hot path
(`/root/reference/src/waveglow/model.py`):

  * `infer`      follows `WaveGlow.infer`                    model.py:223-274
  * `wn`         follows `WN.forward`                        model.py:115-138
  * gate         follows `fused_add_tanh_sigmoid_multiply`   model.py:13-20
  * 1x1 inverse  follows `Invertible1x1Conv.forward(reverse)` model.py:45-60
  * `forward`    follows `WaveGlow.forward`                  model.py:178-221
  * `fold_weight_norm` follows `WaveGlow.remove_weightnorm`  model.py:276-297
  * `loss`       follows `WaveGlowLoss.forward`              train.py:31-45
  * `draw_noise` follows the noise draws inside infer        model.py:234-244, 261-271

The arithmetic itself lives in torch (third party; reference pins torch>=2.0, this image has 2.11):
every contraction is an `F.conv1d` / `F.conv_transpose1d` call exactly as in the reference so that,
on the same CPU, this restatement is BIT-IDENTICAL to the reference module (see
tests/test_oracle_pinned.py, which runs the reference itself when /root/reference is present, and
tests/golden/*.npz which were produced by the unmodified reference via tools/make_golden.py).

Parity status: PINNED by outputs of the reference itself run in the build container (the
reference's own single known-answer test needs a network download).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product (waveglow_b200/) never does.

`emu` hooks (operand rounding emulation) exist only to derive the tolerance budget of the
bf16/tf32 tensor-core modes on CPU; emu=None is the oracle proper.
"""
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- hparams
class ModelDims:
  """The nine ModelHParams ints (hparams.py:20-31) and the derived per-flow channel counts."""

  def __init__(self, n_mel_channels=80, n_flows=12, n_group=8, n_early_every=4, n_early_size=2,
               n_layers=8, n_channels=256, kernel_size=3, segment_length=16000):
    self.n_mel_channels = n_mel_channels
    self.n_flows = n_flows
    self.n_group = n_group
    self.n_early_every = n_early_every
    self.n_early_size = n_early_size
    self.n_layers = n_layers
    self.n_channels = n_channels
    self.kernel_size = kernel_size
    self.segment_length = segment_length
    # model.py:160-175
    n_half = n_group // 2
    n_rem = n_group
    self.flow_channels: List[int] = []
    for k in range(n_flows):
      if k % n_early_every == 0 and k > 0:
        n_half -= n_early_size // 2
        n_rem -= n_early_size
      self.flow_channels.append(n_rem)
    self.n_remaining_channels = n_rem


# ----------------------------------------------------------------------------- weights
def fold_weight_norm(state_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
  """state_dict in any of the three key styles -> plain `.weight` fp32 (model.py:276-297).

  Accepts `parametrizations.weight.original0/1` (torch>=2.0 weight_norm), legacy
  `weight_g/weight_v` (converter/convert.py:82-84) and already-folded `.weight`.
  w = g * v / ||v||, norm over dims (1,2) per output channel (torch._weight_norm, dim=0).
  """
  out: Dict[str, Tensor] = {}
  pend: Dict[str, Dict[str, Tensor]] = {}
  for key, val in state_dict.items():
    val = val.detach().to(torch.float32).cpu()
    if key.endswith(".parametrizations.weight.original0"):
      pend.setdefault(key[: -len(".parametrizations.weight.original0")], {})["g"] = val
    elif key.endswith(".parametrizations.weight.original1"):
      pend.setdefault(key[: -len(".parametrizations.weight.original1")], {})["v"] = val
    elif key.endswith(".weight_g"):
      pend.setdefault(key[: -len(".weight_g")], {})["g"] = val
    elif key.endswith(".weight_v"):
      pend.setdefault(key[: -len(".weight_v")], {})["v"] = val
    else:
      out[key] = val
  for base, gv in pend.items():
    out[base + ".weight"] = torch._weight_norm(gv["v"], gv["g"], 0)
  return out


# ----------------------------------------------------------------------------- pieces
def upsample_squeeze(w: Dict[str, Tensor], dims: ModelDims, mel: Tensor, n_samples: Optional[int] = None) -> Tensor:
  """ConvTranspose1d(80,80,1024,stride=256) + trim + squeeze to [B, 640, T'] (model.py:225-232 / 186-193).

  n_samples=None: infer-style tail trim (drop the last 768 samples);
  n_samples=T   : forward-style front-aligned trim to the audio length.
  """
  up = F.conv_transpose1d(mel, w["upsample.weight"], w["upsample.bias"], stride=256)
  if n_samples is None:
    up = up[:, :, :-(1024 - 256)]
  else:
    assert up.size(2) >= n_samples
    up = up[:, :, :n_samples]
  g = dims.n_group
  s = up.unfold(2, g, g).permute(0, 2, 1, 3)
  s = s.contiguous().view(s.size(0), s.size(1), -1).permute(0, 2, 1)
  return s


def gate(in_res: Tensor, cond: Tensor, n_channels: int) -> Tensor:
  a = in_res + cond
  return torch.tanh(a[:, :n_channels, :]) * torch.sigmoid(a[:, n_channels:, :])


def wn(w: Dict[str, Tensor], dims: ModelDims, k: int, audio_0: Tensor, spect: Tensor,
       emu: Optional[Callable[[Tensor, str], Tensor]] = None) -> Tensor:
  """WN.forward (model.py:115-138).  emu(x, tag) optionally rounds a GEMM operand."""
  q = (lambda x, tag: x) if emu is None else emu
  p = f"WN.{k}."
  nc = dims.n_channels
  audio = F.conv1d(audio_0, w[p + "start.weight"], w[p + "start.bias"])
  output = torch.zeros_like(audio)
  cond = F.conv1d(q(spect, "mel"), q(w[p + "cond_layer.weight"], "w_cond"), w[p + "cond_layer.bias"])
  for i in range(dims.n_layers):
    d = 2 ** i
    pad = (dims.kernel_size * d - d) // 2
    in_res = F.conv1d(q(audio, "x"), q(w[p + f"in_layers.{i}.weight"], "w_in"), w[p + f"in_layers.{i}.bias"],
                      dilation=d, padding=pad)
    acts = gate(in_res, cond[:, 2 * nc * i: 2 * nc * (i + 1), :], nc)
    rs = F.conv1d(q(acts, "acts"), q(w[p + f"res_skip_layers.{i}.weight"], "w_rs"), w[p + f"res_skip_layers.{i}.bias"])
    if i < dims.n_layers - 1:
      audio = audio + rs[:, :nc, :]
      output = output + rs[:, nc:, :]
    else:
      output = output + rs
  return F.conv1d(q(output, "skip"), q(w[p + "end.weight"], "w_end"), w[p + "end.bias"])


def draw_noise(dims: ModelDims, batch: int, n_groups: int, device="cpu") -> List[Tensor]:
  """The N(0,1) draws of infer in the reference's order, shapes and dtype (model.py:239-243, 266-270).

  Returns [z_init[B, n_remaining, T'], z_k... for k descending with k % n_early_every == 0, k > 0].
  Uses torch's global generator on `device` exactly as the reference does (seed it first).
  """
  zs = [torch.empty(batch, dims.n_remaining_channels, n_groups, dtype=torch.float32, device=device).normal_()]
  for k in reversed(range(dims.n_flows)):
    if k % dims.n_early_every == 0 and k > 0:
      zs.append(torch.empty(batch, dims.n_early_size, n_groups, dtype=torch.float32, device=device).normal_())
  return zs


def w_inverse(w: Dict[str, Tensor], k: int) -> Tensor:
  W = w[f"convinv.{k}.conv.weight"].squeeze(-1)
  return W.float().inverse()[..., None]


def infer(w: Dict[str, Tensor], dims: ModelDims, mel: Tensor, sigma: float, zs: Sequence[Tensor],
          emu: Optional[Callable[[Tensor, str], Tensor]] = None) -> Tensor:
  """WaveGlow.infer with injected pre-drawn noise (model.py:223-274).  w = folded weights."""
  spect = upsample_squeeze(w, dims, mel)
  zi = iter(zs)
  audio = sigma * next(zi)
  assert audio.size(2) == spect.size(2)
  for k in reversed(range(dims.n_flows)):
    n_half = audio.size(1) // 2
    audio_0 = audio[:, :n_half, :]
    audio_1 = audio[:, n_half:, :]
    output = wn(w, dims, k, audio_0, spect, emu)
    s = output[:, n_half:, :]
    b = output[:, :n_half, :]
    audio_1 = (audio_1 - b) / torch.exp(s)
    audio = torch.cat([audio_0, audio_1], 1)
    audio = F.conv1d(audio, w_inverse(w, k), bias=None, stride=1, padding=0)
    if k % dims.n_early_every == 0 and k > 0:
      audio = torch.cat((sigma * next(zi), audio), 1)
  return audio.permute(0, 2, 1).contiguous().view(audio.size(0), -1)


def forward(w: Dict[str, Tensor], dims: ModelDims, mel: Tensor, audio: Tensor) -> Tuple[Tensor, List[Tensor], List[Tensor]]:
  """WaveGlow.forward (model.py:178-221): returns (z[B,8,T'], log_s_list, log_det_W_list)."""
  spect = upsample_squeeze(w, dims, mel, n_samples=audio.size(1))
  g = dims.n_group
  audio = audio.unfold(1, g, g).permute(0, 2, 1)
  output_audio, log_s_list, log_det_W_list = [], [], []
  for k in range(dims.n_flows):
    if k % dims.n_early_every == 0 and k > 0:
      output_audio.append(audio[:, :dims.n_early_size, :])
      audio = audio[:, dims.n_early_size:, :]
    W = w[f"convinv.{k}.conv.weight"]
    log_det_W = audio.size(0) * audio.size(2) * torch.logdet(W.squeeze())
    audio = F.conv1d(audio, W)
    log_det_W_list.append(log_det_W)
    n_half = audio.size(1) // 2
    audio_0 = audio[:, :n_half, :]
    audio_1 = audio[:, n_half:, :]
    output = wn(w, dims, k, audio_0, spect)
    log_s = output[:, n_half:, :]
    b = output[:, :n_half, :]
    audio_1 = torch.exp(log_s) * audio_1 + b
    log_s_list.append(log_s)
    audio = torch.cat([audio_0, audio_1], 1)
  output_audio.append(audio)
  return torch.cat(output_audio, 1), log_s_list, log_det_W_list


def loss(z: Tensor, log_s_list: Sequence[Tensor], log_det_W_list: Sequence[Tensor], sigma: float = 1.0) -> Tensor:
  """WaveGlowLoss.forward (train.py:31-45)."""
  log_s_total = 0
  log_det_W_total = 0
  for i, log_s in enumerate(log_s_list):
    if i == 0:
      log_s_total = torch.sum(log_s)
      log_det_W_total = log_det_W_list[i]
    else:
      log_s_total = log_s_total + torch.sum(log_s)
      log_det_W_total = log_det_W_total + log_det_W_list[i]
  l = torch.sum(z * z) / (2 * sigma * sigma) - log_s_total - log_det_W_total
  return l / (z.size(0) * z.size(1) * z.size(2))
