"""Where the tensor mode's error comes from: oracle/engine_model.py's pipeline with ONE rounding point enabled at a time.

  python tools/error_budget.py [channels] [frames]       # CPU only; default 512 channels, 60 frames

Rounding points of the fp16-operand engine: the residual stream x (re-rounded once per layer), the gated activations, the
weights (W_in, V, W_res, W_end W_skip), the mel windows.  Prints rms and max-abs error against the oracle for each point alone,
for all of them, and for an exactly carried residual stream (a hi/lo pair: the taps still see fp16(x)).  DESIGN.md section 2
quotes the 512-channel figures.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import engine_model as E  # noqa: E402
from oracle import waveglow_oracle as O  # noqa: E402
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402

F16, F32 = torch.float16, torch.float32


def run(w, dims, mel, zs, sigma, x_dt, acts_dt, w_dt, mel_dt, exact_carry):
  """E.infer (phase-composed conditioning, explicit start conv) with a rounding dtype per point."""
  state = E.pack_noise(dims, zs, sigma)
  G, nc = dims.n_group, dims.n_channels
  wup_p = E.upsample_phase_weights(w, dims)
  mw = E.rnd(E.mel_windows(mel, (state.size(1) + 31) // 32), mel_dt)
  for k in reversed(range(dims.n_flows)):
    pf = E.PackedFlow(w, dims, k, wup_p)
    c, h = pf.c, pf.h
    a0, a1 = state[..., G - c: G - c + h], state[..., G - c + h:]
    x = E.rnd(a0 @ pf.w_start.t() + pf.b_start, x_dt)
    e = torch.zeros(x.size(0), x.size(1), 2 * h)
    for i in range(dims.n_layers):
      d = 2 ** i
      xr = E.rnd(x, F16) if exact_carry else x
      a = torch.cat([E.shift_time(xr, -d), xr, E.shift_time(xr, d)], dim=2)
      pre = a @ E.rnd(pf.w1[i][:, : 3 * nc], w_dt).t() + pf.b1p[i]
      cond = torch.einsum("bqk,pok->bqpo", mw, E.rnd(pf.v[i], mel_dt))
      pre = pre + cond.reshape(x.size(0), -1, 2 * nc)[:, : x.size(1)]
      acts = E.rnd(torch.tanh(pre[..., :nc]) * torch.sigmoid(pre[..., nc:]), acts_dt)
      e = e + acts @ E.rnd(pf.w_se[i], w_dt).t()
      if pf.w_res[i] is not None:
        x = E.rnd(x + acts @ E.rnd(pf.w_res[i], w_dt).t() + pf.b_res[i], x_dt)
    e = e + pf.e_bias
    a1 = (a1 - e[..., :h]) * torch.exp(-e[..., h:])
    live = torch.cat([a0, a1], dim=2) @ pf.w_inv.t()
    state = torch.cat([state[..., : G - c], live], dim=2)
  return state.reshape(state.size(0), -1)


def main():
  nc = int(sys.argv[1]) if len(sys.argv) > 1 else 512
  frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
  hp = HParams(n_channels=nc)
  w = O.fold_weight_norm(synthetic_state_dict(hp, 1234))
  dims = O.ModelDims(n_channels=nc)
  mel = synthetic_mel(1, frames, 3001)
  zs = synthetic_noise(hp, 1, frames * 32, 3002)
  with torch.no_grad():
    ref = O.infer(w, dims, mel, 0.666, zs)
    cases = {"all fp16 (the engine)": (F16, F16, F16, F16, False), "residual stream x only": (F16, F32, F32, F32, False),
             "gated activations only": (F32, F16, F32, F32, False), "weights only": (F32, F32, F16, F32, False),
             "mel windows + V only": (F32, F32, F32, F16, False), "all but x": (F32, F16, F16, F16, False),
             "exact residual carry, rest fp16": (F32, F16, F16, F16, True), "exact residual carry, nothing else": (F32, F32, F32, F32, True),
             "nothing rounded (restructuring only)": (F32, F32, F32, F32, False)}
    total = None
    for name, (xd, ad, wd, md, carry) in cases.items():
      err = (run(w, dims, mel, zs, 0.666, xd, ad, wd, md, carry) - ref).abs()
      rms = float(err.pow(2).mean().sqrt())
      total = total or rms
      print(f"{nc} ch, {frames} frames  {name:38s} rms {rms:.3e} ({100 * (rms / total) ** 2:5.1f} % of the variance)  max-abs {float(err.max()):.3e}", flush=True)


if __name__ == "__main__":
  main()
