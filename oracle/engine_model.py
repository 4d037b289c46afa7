"""CPU model of the B200 engine's RESTRUCTURED algorithm  --  TEST INFRASTRUCTURE, NOT PRODUCT.

The CUDA engine does not evaluate WaveGlow.infer (model.py:223-274) op by op.  It uses an
algebraically equivalent form that suits one fused tcgen05 kernel per WN layer:

  * activations are time-major  [B, T', C]  (C contiguous == K-major MMA operand);
  * the conditioning 1x1 conv (model.py:121) is a K-extension of each layer's dilated conv:
        in_i(t) = [W_in_i(tap-1) | W_in_i(tap0) | W_in_i(tap+1) | W_cond_i] . [x(t-d); x(t); x(t+d); S(t)]
    with the two biases folded;
  * (tensor-core mode, `phase_cond=True`) the mel upsampler (ConvTranspose1d, model.py:145-150, 225) is
    composed INTO the conditioning weights: a group-step t = 32 q + p (frame q, phase p) sees only the 4 mel
    frames q..q-3, so  W_cond_i . S(t) = V_{i,p} . [mel(q); mel(q-1); mel(q-2); mel(q-3)] + const  with
    V_{i,p} = W_cond_i . Wup_p  (K = 320 instead of 640; no upsampled tensor is ever formed);
  * the skip path and the `end` 1x1 conv (model.py:130-137) are folded per layer:
        end(sum_i skip_i) = sum_i (W_end W_skip_i) acts_i + (W_end sum_i b_skip_i + b_end)
    so each layer emits only 2h <= 8 "skip-end" channels accumulated in fp32 (`e`);
  * (default 256-channel tensor-core kernel, `first_composed=True`) the `start` 1x1 conv (model.py:117) is composed into
    the first layer: in_0(start(a0))(t) = sum_tau (W_in_0[tau] W_start) a0(t+tau) + (W_in_0[tau] b_start) 1(t+tau), where
    1(t) is 1 inside the sequence and 0 in the zero padding, and the residual carry of that layer is
    x(t) = [W_start | b_start] [a0(t); 1]; the 256-channel x of a flow's first layer is never formed;
  * the residual stream x is carried in the MMA operand dtype (fp16 by default) and updated
    on the tensor core:  x <- [W_res_i | I] . [acts; x] + b_res_i;
  * the audio state is time-major [B, T', 8]; live channels are the LAST C slots, so the two
    early-output concats (model.py:260-271) and the final de-squeeze (model.py:273) are no-ops.

This file restates that pipeline in torch on the CPU with the same rounding points
(`op_dtype`: torch.float16 / torch.bfloat16 / torch.float32 = no rounding) so that
  (1) tests can show the restructuring is exact (op_dtype=float32 vs the oracle, ~1e-6), and
  (2) the tolerance budget of the fp16/bf16 modes is derived on CPU before any GPU run, and
  (3) per-layer intermediates are available to localise a GPU mismatch.
"""
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import waveglow_oracle as O

Tensor = torch.Tensor


def rnd(x: Tensor, dt: torch.dtype) -> Tensor:
  return x if dt == torch.float32 else x.to(dt).to(torch.float32)


def upsample_phase_weights(w: Dict[str, Tensor], dims: O.ModelDims) -> Tensor:
  """Wup_p[p, cj = c*8+j, k = m*80+ci] = Wup[ci, c, 8p + j + 256m]   (float64), p < 32, m < 4."""
  wup = w["upsample.weight"].double()                      # [ci, c, 1024]
  n_mel, g = dims.n_mel_channels, dims.n_group
  phases = 256 // g
  out = torch.zeros(phases, n_mel * g, 4 * n_mel, dtype=torch.float64)
  for p in range(phases):
    for m in range(4):
      blk = wup[:, :, g * p + 256 * m: g * p + 256 * m + g]      # [ci, c, j]
      out[p, :, m * n_mel:(m + 1) * n_mel] = blk.permute(1, 2, 0).reshape(n_mel * g, n_mel)
  return out


def mel_windows(mel: Tensor, n_frames_out: int) -> Tensor:
  """MW[b, q, m*80+ci] = mel[b, ci, q-m] (0 outside [0,F)), q < n_frames_out."""
  B, n_mel, F = mel.shape
  mw = torch.zeros(B, n_frames_out, 4 * n_mel, dtype=mel.dtype)
  for m in range(4):
    lo, hi = m, min(n_frames_out, F + m)
    if hi > lo:
      mw[:, lo:hi, m * n_mel:(m + 1) * n_mel] = mel[:, :, lo - m: hi - m].permute(0, 2, 1)
  return mw


class PackedFlow:
  """Per-flow packed weights in the engine's layout (fp32 master values, rounded on use)."""

  def __init__(self, w: Dict[str, Tensor], dims: O.ModelDims, k: int, wup_p: Optional[Tensor] = None):
    nc, L = dims.n_channels, dims.n_layers
    p = f"WN.{k}."
    c = dims.flow_channels[k]
    h = c // 2
    self.h, self.c = h, c
    self.w_start = w[p + "start.weight"][:, :, 0].double()       # [nc, h]
    self.b_start = w[p + "start.bias"].double()
    w_end = w[p + "end.weight"][:, :, 0].double()                 # [2h, nc]
    w_cond = w[p + "cond_layer.weight"][:, :, 0].double()         # [2 nc L, 640]
    b_cond = w[p + "cond_layer.bias"].double()
    self.v: List[Tensor] = []    # phase-composed conditioning [32, 2nc, 320] (only when wup_p is given)
    self.b1p: List[Tensor] = []  # bias for the phase-composed form (adds W_cond . upsample.bias)
    self.w1: List[Tensor] = []   # [2nc, 3nc + 640]   columns: tap-1 | tap0 | tap+1 | mel
    self.b1: List[Tensor] = []
    self.w_res: List[Optional[Tensor]] = []   # [nc, nc]
    self.b_res: List[Optional[Tensor]] = []
    self.w_se: List[Tensor] = []   # [2h, nc]  skip-end fold
    e_bias = w[p + "end.bias"].double().clone()
    for i in range(L):
      wi = w[p + f"in_layers.{i}.weight"].double()                # [2nc, nc, 3]
      cat = torch.cat([wi[:, :, 0], wi[:, :, 1], wi[:, :, 2], w_cond[2 * nc * i: 2 * nc * (i + 1)]], dim=1)
      self.w1.append(cat.float())
      self.b1.append((w[p + f"in_layers.{i}.bias"].double() + b_cond[2 * nc * i: 2 * nc * (i + 1)]).float())
      if wup_p is not None:
        wc = w_cond[2 * nc * i: 2 * nc * (i + 1)]                       # [2nc, 640]
        self.v.append(torch.einsum("oc,pck->pok", wc, wup_p).float())    # [32, 2nc, 320]
        b_up = w["upsample.bias"].double().repeat_interleave(dims.n_group)
        self.b1p.append((self.b1[-1].double() + wc @ b_up).float())
      wrs = w[p + f"res_skip_layers.{i}.weight"][:, :, 0].double()
      brs = w[p + f"res_skip_layers.{i}.bias"].double()
      if i < L - 1:
        self.w_res.append(wrs[:nc].float())
        self.b_res.append(brs[:nc].float())
        w_skip, b_skip = wrs[nc:], brs[nc:]
      else:
        self.w_res.append(None)
        self.b_res.append(None)
        w_skip, b_skip = wrs, brs
      self.w_se.append((w_end @ w_skip).float())
      e_bias = e_bias + w_end @ b_skip
    self.e_bias = e_bias.float()                                  # [2h]: rows [:h] -> b, [h:] -> s
    self.w_start = self.w_start.float()
    self.b_start = self.b_start.float()
    W = w[f"convinv.{k}.conv.weight"][:, :, 0]
    self.w_inv = W.float().inverse()                              # model.py:51-53


def shift_time(x: Tensor, s: int) -> Tensor:
  """y[:, t] = x[:, t + s] with zeros outside [0, T')  (zero padding of the dilated conv, model.py:100-102)."""
  if s == 0:
    return x
  y = torch.zeros_like(x)
  T = x.size(1)
  if abs(s) >= T:
    return y
  if s > 0:
    y[:, : T - s] = x[:, s:]
  else:
    y[:, -s:] = x[:, : T + s]
  return y


def wn_layers(pf: PackedFlow, dims: O.ModelDims, x: Tensor, S: Tensor, op_dtype: torch.dtype,
              trace: Optional[list] = None, mw: Optional[Tensor] = None, a0: Optional[Tensor] = None,
              exact_carry: bool = False) -> Tensor:
  """x: [B,T',nc] (already in op_dtype precision), S: [B,T',640] (or None with mw = mel windows [B,F',320]).
  a0 [B,T',h] (with mw): the first layer takes the live half of the audio state itself (start conv composed in).
  exact_carry (the engine's 512-channel kernel): the residual stream is carried as a two-term split hi + lo in the operand
  dtype, i.e. exactly to ~2^-22; only the dilated-conv taps see its rounded value.
  Returns e [B,T',2h] fp32 (without e_bias)."""
  nc = dims.n_channels
  store = (lambda v: v) if exact_carry else (lambda v: rnd(v, op_dtype))
  e = torch.zeros(x.size(0), x.size(1), 2 * pf.h, dtype=torch.float32)
  for i in range(dims.n_layers):
    d = 2 ** i
    if a0 is not None and i == 0:
      assert mw is not None and pf.w_res[0] is not None
      # operand rows [a_hi | a_lo | a_hi | 1 | 1] against weight columns [W_hi | W_hi | W_lo | b_hi | b_lo]: two-term splits in
      # the operand dtype, so the start conv enters at nearly fp32 accuracy
      a_hi = rnd(a0, op_dtype)
      a_lo = rnd(a0 - a_hi, op_dtype)
      one = torch.ones_like(a0[..., :1])
      a01 = torch.cat([a_hi, a_lo, a_hi, one, one], dim=2)                                # [B,T',3h+2]

      def split_cols(wb):                                                                 # wb [rows, h+1] float64
        hi = rnd(wb.float(), op_dtype)
        lo = rnd((wb - hi.double()).float(), op_dtype)
        return torch.cat([hi[:, :-1], hi[:, :-1], lo[:, :-1], hi[:, -1:], lo[:, -1:]], dim=1)

      ws1 = torch.cat([pf.w_start.double(), pf.b_start.double()[:, None]], dim=1)        # [nc, h+1]
      pre = pf.b1p[0].clone()
      for tap in range(3):
        wc = pf.w1[0][:, tap * nc:(tap + 1) * nc].double() @ ws1                         # [2nc, h+1]
        pre = pre + shift_time(a01, tap - 1) @ split_cols(wc).t()
      B, T = x.size(0), x.size(1)
      cond = torch.einsum("bqk,pok->bqpo", mw, rnd(pf.v[0], op_dtype))
      pre = pre + cond.reshape(B, -1, 2 * nc)[:, :T]
      acts = rnd(torch.tanh(pre[..., :nc]) * torch.sigmoid(pre[..., nc:]), op_dtype)
      e = e + acts @ rnd(pf.w_se[0], op_dtype).t()
      x = store(a01 @ split_cols(ws1).t() + acts @ rnd(pf.w_res[0], op_dtype).t() + pf.b_res[0])
      if trace is not None:
        trace.append({"layer": 0, "acts": acts.clone(), "x": x.clone(), "e": e.clone()})
      continue
    xt = rnd(x, op_dtype)                                                        # what the taps read (== x without exact_carry)
    if mw is None:
      a = torch.cat([shift_time(xt, -d), xt, shift_time(xt, d), S], dim=2)       # [B,T',3nc+640]
      pre = a @ rnd(pf.w1[i], op_dtype).t() + pf.b1[i]                           # fp32 accumulate
    else:
      a = torch.cat([shift_time(xt, -d), xt, shift_time(xt, d)], dim=2)
      pre = a @ rnd(pf.w1[i][:, : 3 * nc], op_dtype).t() + pf.b1p[i]
      B, T, P = x.size(0), x.size(1), 256 // dims.n_group
      cond = torch.einsum("bqk,pok->bqpo", mw, rnd(pf.v[i], op_dtype))           # [B,F',32,2nc]
      pre = pre + cond.reshape(B, -1, 2 * nc)[:, :T]
    acts = rnd(torch.tanh(pre[..., :nc]) * torch.sigmoid(pre[..., nc:]), op_dtype)
    e = e + acts @ rnd(pf.w_se[i], op_dtype).t()
    if pf.w_res[i] is not None:
      x = store(x + acts @ rnd(pf.w_res[i], op_dtype).t() + pf.b_res[i])
    if trace is not None:
      trace.append({"layer": i, "acts": acts.clone(), "x": x.clone(), "e": e.clone()})
  return e


def upsample_squeeze_tm(w: Dict[str, Tensor], dims: O.ModelDims, mel: Tensor, n_samples: Optional[int] = None) -> Tensor:
  """Time-major squeezed conditioning S[b, t', c*8+j] = up[b, c, 8t'+j]  (model.py:225-232)."""
  return O.upsample_squeeze(w, dims, mel, n_samples).permute(0, 2, 1).contiguous()


def pack_noise(dims: O.ModelDims, zs: Sequence[Tensor], sigma: float) -> Tensor:
  """Initial audio state [B,T',8]: slots = [z_last_early ... z_first_early, z_init] * sigma."""
  parts = [sigma * z.permute(0, 2, 1) for z in zs]      # zs order: init, early(k=8), early(k=4)
  return torch.cat(list(reversed(parts)), dim=2).contiguous()


def infer(w: Dict[str, Tensor], dims: O.ModelDims, mel: Tensor, sigma: float, zs: Sequence[Tensor],
          op_dtype: torch.dtype = torch.float16, trace: Optional[list] = None, phase_cond: bool = False,
          first_composed: bool = False, exact_carry: bool = False) -> Tensor:
  state = pack_noise(dims, zs, sigma)                   # [B,T',8] fp32
  G = dims.n_group
  if phase_cond:
    wup_p = upsample_phase_weights(w, dims)
    S = None
    mw = rnd(mel_windows(mel, (state.size(1) + 31) // 32), op_dtype)
  else:
    wup_p, mw = None, None
    S = rnd(upsample_squeeze_tm(w, dims, mel), op_dtype)
  for k in reversed(range(dims.n_flows)):
    pf = PackedFlow(w, dims, k, wup_p)
    c, h = pf.c, pf.h
    a0 = state[..., G - c: G - c + h]
    a1 = state[..., G - c + h:]
    x = a0 @ pf.w_start.t() + pf.b_start
    if not exact_carry:
      x = rnd(x, op_dtype)
    tr = [] if trace is not None else None
    e = wn_layers(pf, dims, x, S, op_dtype, tr, mw, a0 if first_composed else None, exact_carry) + pf.e_bias
    b, s = e[..., :h], e[..., h:]
    a1 = (a1 - b) * torch.exp(-s)
    live = torch.cat([a0, a1], dim=2) @ pf.w_inv.t()
    state = torch.cat([state[..., : G - c], live], dim=2)
    if trace is not None:
      trace.append({"flow": k, "layers": tr, "e": e.clone(), "state": state.clone()})
  return state.reshape(state.size(0), -1)
