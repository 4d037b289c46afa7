"""`WaveGlow` -- drop-in for the reference's `waveglow.model.WaveGlow` whose hot path runs on the B200 engine.

The module tree, parameter names and shapes mirror /root/reference/src/waveglow/model.py:142-176 (and
`WN` 68-113, `Invertible1x1Conv` 23-43) exactly, so `load_state_dict`, `state_dict`, `.to(device)`,
`.eval()`, `WaveGlow.remove_weightnorm(model)`, `Denoiser(waveglow=model, ...)` and the reference's
`load_model` work unchanged.  The modules are PARAMETER CONTAINERS only: the arithmetic
of `infer` (model.py:223-274) is done by hand-written sm_100a kernels behind the C ABI
(waveglow_b200/engine.py).  There is no torch/CPU fallback: calling `infer` without a CUDA device or
without the built library raises.
"""
import os
import weakref
from logging import getLogger
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import engine as _engine
from .hparams import HParams
from .synthetic import flow_channels


class Invertible1x1Conv(torch.nn.Module):
  """Holds the C x C weight of model.py:30-43 (random orthonormal init, det = +1)."""

  def __init__(self, c: int):
    super().__init__()
    self.conv = torch.nn.Conv1d(c, c, kernel_size=1, stride=1, padding=0, bias=False)
    W = torch.linalg.qr(torch.empty(c, c).normal_())[0]
    if torch.det(W) < 0:
      W[:, 0] = -1 * W[:, 0]
    self.conv.weight.data = W.view(c, c, 1)


class WN(torch.nn.Module):
  """Parameter container with the layout of model.py:75-113 (weight-normed start / cond / in / res_skip, plain zero-init end)."""

  def __init__(self, n_in_channels: int, n_mel_channels: int, hparams: HParams):
    super().__init__()
    assert hparams.kernel_size % 2 == 1
    assert hparams.n_channels % 2 == 0
    self.n_layers = hparams.n_layers
    self.n_channels = hparams.n_channels
    wn = torch.nn.utils.parametrizations.weight_norm
    self.in_layers = torch.nn.ModuleList()
    self.res_skip_layers = torch.nn.ModuleList()
    self.start = wn(torch.nn.Conv1d(n_in_channels, self.n_channels, 1), name="weight")
    end = torch.nn.Conv1d(self.n_channels, 2 * n_in_channels, 1)
    end.weight.data.zero_()
    end.bias.data.zero_()
    self.end = end
    self.cond_layer = wn(torch.nn.Conv1d(n_mel_channels, 2 * self.n_channels * self.n_layers, 1), name="weight")
    for i in range(self.n_layers):
      dilation = 2 ** i
      padding = int((hparams.kernel_size * dilation - dilation) / 2)
      self.in_layers.append(wn(torch.nn.Conv1d(self.n_channels, 2 * self.n_channels, hparams.kernel_size,
                                               dilation=dilation, padding=padding), name="weight"))
      rs = 2 * self.n_channels if i < self.n_layers - 1 else self.n_channels
      self.res_skip_layers.append(wn(torch.nn.Conv1d(self.n_channels, rs, 1), name="weight"))


def _remove(conv_list: torch.nn.ModuleList) -> torch.nn.ModuleList:
  out = torch.nn.ModuleList()
  for conv in conv_list:
    out.append(torch.nn.utils.parametrize.remove_parametrizations(conv, "weight"))
  return out


class WaveGlow(torch.nn.Module):
  """Same constructor, attributes and call surface as the reference WaveGlow (model.py:141-297).

  Extra, engine-only knobs:
    precision: "fp16" (tcgen05 tensor cores, fp16 operands / fp32 accumulate; default) or "fp32"
               (CUDA-core correctness mode).  Can also be set with `model.set_precision(...)`.
  """

  def __init__(self, hparams: HParams, precision: str = "fp16"):
    super().__init__()
    self.hparams = hparams
    self.upsample = torch.nn.ConvTranspose1d(hparams.n_mel_channels, hparams.n_mel_channels, 1024, stride=256)
    assert hparams.n_group % 2 == 0
    self.n_flows = hparams.n_flows
    self.n_group = hparams.n_group
    self.n_early_every = hparams.n_early_every
    self.n_early_size = hparams.n_early_size
    self.WN = torch.nn.ModuleList()
    self.convinv = torch.nn.ModuleList()
    chans = flow_channels(hparams)
    for k in range(self.n_flows):
      self.convinv.append(Invertible1x1Conv(chans[k]))
      self.WN.append(WN(chans[k] // 2, hparams.n_mel_channels * self.n_group, hparams))
    self.n_remaining_channels = chans[-1]
    self._precision = precision
    self._engine: Optional[_engine.Engine] = None
    self._packed = False
    self.pack_cache = None         # (path, fingerprint), set by Synthesizer(pack_cache=...) right after it built the model
    self.pack_cache_hit = False
    # opt-in: replay each (batch, frames, sigma) shape from a CUDA graph (latency of short single utterances)
    self.use_cuda_graphs = False

  # ------------------------------------------------------------------ engine plumbing
  def set_precision(self, precision: str) -> "WaveGlow":
    assert precision in ("fp16", "fp32")
    if precision != self._precision:
      self._precision = precision
      self._drop_engine()
    return self

  def _drop_engine(self) -> None:
    if self._engine is not None:
      self._engine.close()
    self._engine = None
    self._packed = False

  def invalidate_packed_weights(self) -> None:
    """Call after changing parameters in place; the reference has the same staleness for its cached
    `W_inverse` (model.py:49-58).  load_state_dict / remove_weightnorm / .to() call it for you.
    A packed-weight cache (`pack_cache`) identifies the CHECKPOINT FILE the model was built from, not the live
    parameters, so it is only good for the first pack after construction: it is dropped here."""
    self._packed = False
    self.pack_cache = None

  def load_state_dict(self, *args, **kwargs):
    res = super().load_state_dict(*args, **kwargs)
    self.invalidate_packed_weights()
    return res

  def _apply(self, fn, *args, **kwargs):
    res = super()._apply(fn, *args, **kwargs)
    self._drop_engine()
    return res

  def folded_weights(self) -> Dict[str, torch.Tensor]:
    """name -> folded fp32 weight, keyed like state_dict() after remove_weightnorm (model.py:276-297)."""
    out: Dict[str, torch.Tensor] = {"upsample.weight": self.upsample.weight, "upsample.bias": self.upsample.bias}
    for k in range(self.n_flows):
      wn = self.WN[k]
      p = f"WN.{k}."
      out[p + "start.weight"], out[p + "start.bias"] = wn.start.weight, wn.start.bias
      out[p + "end.weight"], out[p + "end.bias"] = wn.end.weight, wn.end.bias
      out[p + "cond_layer.weight"], out[p + "cond_layer.bias"] = wn.cond_layer.weight, wn.cond_layer.bias
      for i in range(wn.n_layers):
        out[p + f"in_layers.{i}.weight"], out[p + f"in_layers.{i}.bias"] = wn.in_layers[i].weight, wn.in_layers[i].bias
        out[p + f"res_skip_layers.{i}.weight"] = wn.res_skip_layers[i].weight
        out[p + f"res_skip_layers.{i}.bias"] = wn.res_skip_layers[i].bias
      out[f"convinv.{k}.conv.weight"] = self.convinv[k].conv.weight
    return {n: t.detach() for n, t in out.items()}

  def engine(self) -> _engine.Engine:
    device = self.upsample.weight.device
    if device.type != "cuda":
      raise _engine.WgError("WaveGlow.infer runs on the B200 engine only: move the model to a CUDA device "
                            "(there is no CPU fallback)")
    if self._engine is None:
      prec = _engine.PRECISION_FP16 if self._precision == "fp16" else _engine.PRECISION_FP32
      self._engine = _engine.Engine(self.hparams, prec, device)
      self._packed = False
    if not self._packed:
      cache = self.pack_cache
      if cache is not None and os.path.isfile(cache[0]) and self._engine.load_packed(cache[0], cache[1]):
        self.pack_cache_hit = True
      else:
        with torch.no_grad():
          self._engine.load_weights(self.folded_weights())
        self.pack_cache_hit = False
        if cache is not None:
          try:
            self._engine.save_packed(cache[0], cache[1])
          except _engine.WgError as ex:                    # a read-only checkpoint folder must not break synthesis
            getLogger(__name__).warning(f"packed-weight cache not written: {ex}")
      self._packed = True
    return self._engine

  # ------------------------------------------------------------------ reference call surface
  def draw_noise(self, batch: int, n_groups: int, device: torch.device) -> List[torch.Tensor]:
    """N(0,1) from torch's global generator on `device`, in the reference's order, shapes and dtype
    (model.py:239-243 then 266-270 for k = 8, 4) so a seed means the same thing as in the reference."""
    zs = [torch.empty(batch, self.n_remaining_channels, n_groups, dtype=torch.float32, device=device).normal_()]
    for k in reversed(range(self.n_flows)):
      if k % self.n_early_every == 0 and k > 0:
        zs.append(torch.empty(batch, self.n_early_size, n_groups, dtype=torch.float32, device=device).normal_())
    return zs

  def infer_with_noise(self, spect: torch.Tensor, zs: Sequence[torch.Tensor], sigma: float = 1.0) -> torch.Tensor:
    if spect.dim() != 3 or spect.size(1) != self.hparams.n_mel_channels:
      raise ValueError(f"spect must be [B, {self.hparams.n_mel_channels}, frames], got {tuple(spect.shape)}")
    eng = self.engine()
    with torch.no_grad():
      if self.use_cuda_graphs:
        return eng.infer_graphed(spect, zs, sigma)
      return eng.infer(spect, zs, sigma)

  def infer(self, spect: torch.Tensor, sigma: float = 1.0) -> torch.Tensor:
    """audio[B, 256 * frames] = inverse flow of sigma * N(0,1) conditioned on spect[B, 80, frames] (model.py:223-274)."""
    n_groups = spect.size(2) * 256 // self.n_group
    zs = self.draw_noise(spect.size(0), n_groups, spect.device)
    return self.infer_with_noise(spect, zs, sigma)

  def forward(self, forward_input: Tuple[torch.Tensor, torch.Tensor]):
    """(spect[B,80,F], audio[B,T]) -> (z[B,8,T/8], log_s_list, log_det_W_list) like model.py:178-221.

    Inference-only (no autograd graph: the reference's optimiser/backward are out of scope).
    The engine also returns the reductions WaveGlowLoss needs; they are kept in `self.last_forward_stats`
    (`log_s_sum[k]`, `log_det_w[k]`) so `waveglow_b200.model.WaveGlowLoss` does not re-reduce."""
    spect, audio = forward_input
    if spect.dim() != 3 or spect.size(1) != self.hparams.n_mel_channels or audio.dim() != 2:
      raise ValueError("forward_input must be (spect[B, n_mel, frames], audio[B, samples])")
    # model.py:187: assert(spect.size(2) >= audio.size(1)) on the upsampled mel
    assert (spect.size(2) - 1) * 256 + 1024 >= audio.size(1)
    eng = self.engine()
    halves = [c // 2 for c in flow_channels(self.hparams)]
    with torch.no_grad():
      z, log_s_list, sums, dets, zsq = eng.forward(spect, audio, halves, want_log_s=True)
    self.last_forward_stats = ForwardStats(z, log_s_list, sums, dets, zsq)
    # the fused reductions travel WITH the output tuple, so WaveGlowLoss can tell an engine output from foreign tensors
    z._wg_forward_stats = self.last_forward_stats
    return z, log_s_list, [dets[k] for k in range(self.n_flows)]

  @staticmethod
  def remove_weightnorm(model: "WaveGlow") -> "WaveGlow":
    """Folds w = g v / ||v|| for start, in_layers, cond_layer, res_skip_layers (model.py:276-297)."""
    for wnet in model.WN:
      if torch.nn.utils.parametrize.is_parametrized(wnet.start, "weight"):
        wnet.start = torch.nn.utils.parametrize.remove_parametrizations(wnet.start, "weight")
        wnet.in_layers = _remove(wnet.in_layers)
        wnet.cond_layer = torch.nn.utils.parametrize.remove_parametrizations(wnet.cond_layer, "weight")
        wnet.res_skip_layers = _remove(wnet.res_skip_layers)
    model.invalidate_packed_weights()
    return model


class ForwardStats(dict):
  """Reductions the engine fuses into `wg_forward`: `log_s_sum[k]` = sum of flow k's log_s (reduced in
  double by the coupling kernel), `log_det_w[k]` = B * T' * log|det W_k|, `z_sq_sum` = sum z^2 (reduced by the kernel that
  writes z).  Remembers which tensors they describe so a loss computed on other tensors does not pick them up."""

  def __init__(self, z, log_s_list, log_s_sum, log_det_w, z_sq_sum):
    super().__init__(log_s_sum=log_s_sum, log_det_w=log_det_w, z_sq_sum=z_sq_sum)
    # WEAK references: z carries this object as an attribute, so a strong reference back would be a cycle that keeps every
    # forward output alive until Python's cyclic collector runs (measured: 10-120 ms stalls per call from the CUDA caching
    # allocator growing and trimming, profiles/r02/forward_times_cycle.txt)
    self._z = weakref.ref(z)
    self._log_s = [weakref.ref(t) for t in log_s_list] if log_s_list is not None else None

  def describes(self, z, log_s_list) -> bool:
    if z is not self._z() or z._version != 0:
      return False
    if self._log_s is None or len(log_s_list) != len(self._log_s):
      return False
    return all(a is b() and a._version == 0 for a, b in zip(log_s_list, self._log_s))


class WaveGlowLoss(torch.nn.Module):
  """Mirror of train.py:26-45: (sum z^2 / (2 sigma^2) - sum log_s - sum log_det_W) / (B * 8 * T').

  Called on the output of `WaveGlow.forward` of this package it consumes the reductions the engine already made inside
  `wg_forward` (no pass over z or the log_s tensors: three tiny ops on [n_flows] vectors); called on any other tensors
  (or on engine outputs that were modified in place since) it reduces them with torch exactly like the reference."""

  def __init__(self, sigma: float = 1.0):
    super().__init__()
    self.sigma = sigma
    self.used_fused_stats = False

  def forward(self, model_output):
    z, log_s_list, log_det_W_list = model_output
    stats = getattr(z, "_wg_forward_stats", None)
    n = z.size(0) * z.size(1) * z.size(2)
    if isinstance(stats, ForwardStats) and stats.describes(z, log_s_list) and len(log_det_W_list) == stats["log_det_w"].numel():
      self.used_fused_stats = True
      loss = stats["z_sq_sum"][0] / (2 * self.sigma * self.sigma) - stats["log_s_sum"].sum() - stats["log_det_w"].sum()
      return loss / n
    self.used_fused_stats = False
    log_s_total = 0
    log_det_W_total = 0
    for i, log_s in enumerate(log_s_list):
      log_s_total = log_s_total + torch.sum(log_s)
      log_det_W_total = log_det_W_total + log_det_W_list[i]
    loss = torch.sum(z * z) / (2 * self.sigma * self.sigma) - log_s_total - log_det_W_total
    return loss / n


def load_model(hparams: HParams, state_dict: Optional[dict], device: torch.device, precision: str = "fp16") -> WaveGlow:
  """Mirror of train.py:48-55, minus `try_copy_to`'s silent CPU fallback (do not replicate)."""
  model = WaveGlow(hparams, precision=precision).to(device)
  if state_dict is not None:
    model.load_state_dict(state_dict)
  return model
