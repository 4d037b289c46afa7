"""Pin the BASELINE.json shapes against the UNMODIFIED reference: decimated golden vectors under tests/golden/.

  python tools/make_golden_large.py [case ...]

The fixtures of tools/make_golden.py are small (<= 40 frames); the headline shapes (configs 2-5: 861 and 5168 frames, 512
channels, the 24 x 16000-sample training batch) produce megabytes of audio per case.  This script runs the reference's own
`WaveGlow.infer` / `WaveGlow.forward` (through tools/ref_loader.py) at those shapes in the build container, asserts
that oracle/waveglow_oracle.py reproduces it BIT FOR BIT at that size too, and commits a decimated view of the result:
every 61st sample of every utterance (61 is prime to the 8-sample group, the 256-sample frame and the 128-frame tile, so all
phases and tile positions are hit) plus the first and last 1024 samples at full resolution (sequence boundaries: zero padding
of the dilated convs, the upsampler's tail trim).  Inputs are NOT stored: mel, noise and weights are regenerated from the
recorded seeds by waveglow_b200/synthetic.py (explicit CPU generators, identical on every machine).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import waveglow_oracle as O  # noqa: E402
from ref_loader import load_reference  # noqa: E402
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict  # noqa: E402

STRIDE = 61
EDGE = 1024
WEIGHT_SEED = 1234
SIGMA = 0.666

INFER_CASES = [
  # name, n_channels, batch, frames, mel seed, noise seed            BASELINE.json
  ("large_infer_c256_b2_f861", 256, 2, 861, 1002, 7),        # configs[1]: two utterances of the headline batch (same seeds as bench.py)
  ("large_infer_c256_b1_f5168", 256, 1, 5168, 4001, 4002),   # configs[3]: 60 s single utterance
  ("large_infer_c512_b1_f140", 512, 1, 140, 3001, 3002),     # configs[2] architecture: one whole 128-frame tile + packed tail
  ("large_infer_c512_b2_f300", 512, 2, 300, 3003, 3004),     # ... multi-tile, two utterances
]
FORWARD_CASES = [
  # name, n_channels, batch, samples, frames, mel seed, audio seed
  ("large_forward_c256_b24_t16000", 256, 24, 16000, 63, 5001, 5002),    # configs[4]
]


def decimate(x: torch.Tensor) -> dict:
  """x [B, n] -> {"strided": x[:, ::61], "head": x[:, :1024], "tail": x[:, -1024:]}"""
  return {"strided": x[:, ::STRIDE].contiguous().numpy(), "head": x[:, :EDGE].contiguous().numpy(),
          "tail": x[:, -EDGE:].contiguous().numpy()}


def main(only):
  WaveGlow, RefHParams = load_reference()
  torch.set_num_threads(os.cpu_count())
  out_dir = os.path.join(ROOT, "tests", "golden")
  models = {}

  def get_model(nc):
    if nc not in models:
      hp = HParams(n_channels=nc)
      sd = synthetic_state_dict(hp, WEIGHT_SEED)
      m = WaveGlow(RefHParams(n_channels=nc))
      res = m.load_state_dict(sd)
      assert not res.missing_keys and not res.unexpected_keys
      m = WaveGlow.remove_weightnorm(m).eval()
      models[nc] = (hp, m, O.fold_weight_norm(sd), O.ModelDims(n_channels=nc))
    return models[nc]

  for name, nc, B, F, mel_seed, noise_seed in INFER_CASES:
    if only and name not in only:
      continue
    hp, m, w, dims = get_model(nc)
    mel = synthetic_mel(B, F, mel_seed)
    zs = synthetic_noise(hp, B, F * 32, noise_seed)
    # the reference draws its noise inside infer from the global generator: feed it the SAME tensors by patching the draw
    t0 = time.perf_counter()
    with torch.no_grad():
      ref = reference_infer_with_noise(m, mel, SIGMA, zs)
    t1 = time.perf_counter()
    with torch.no_grad():
      mine = O.infer(w, dims, mel, SIGMA, zs)
    assert torch.equal(ref, mine), f"{name}: oracle is not bit-identical to the reference at this shape"
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), n_channels=np.int32(nc), batch=np.int32(B), frames=np.int32(F),
                        mel_seed=np.int32(mel_seed), noise_seed=np.int32(noise_seed), weight_seed=np.int32(WEIGHT_SEED),
                        sigma=np.float32(SIGMA), stride=np.int32(STRIDE), edge=np.int32(EDGE), abs_max=np.float32(ref.abs().max()),
                        power=np.float64((ref.double() ** 2).mean()), **decimate(ref))
    print(f"{name}: audio {tuple(ref.shape)} absmax {float(ref.abs().max()):.3f}; reference {t1 - t0:.1f} s; oracle bit-identical", flush=True)

  for name, nc, B, T, F, mel_seed, audio_seed in FORWARD_CASES:
    if only and name not in only:
      continue
    hp, m, w, dims = get_model(nc)
    mel = synthetic_mel(B, F, mel_seed)
    audio = torch.rand((B, T), generator=torch.Generator().manual_seed(audio_seed)) * 2 - 1
    with torch.no_grad():
      z, log_s_list, log_det_W_list = m((mel, audio))
      z2, ls2, ld2 = O.forward(w, dims, mel, audio)
    assert torch.equal(z, z2) and all(torch.equal(a, b) for a, b in zip(log_s_list, ls2)), f"{name}: oracle forward mismatch"
    loss = O.loss(z, log_s_list, log_det_W_list, 1.0)
    zf = z.reshape(B, -1)                               # [B, 8 * T'] channel-major per utterance
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), n_channels=np.int32(nc), batch=np.int32(B), samples=np.int32(T),
                        frames=np.int32(F), mel_seed=np.int32(mel_seed), audio_seed=np.int32(audio_seed),
                        weight_seed=np.int32(WEIGHT_SEED), stride=np.int32(STRIDE), edge=np.int32(EDGE),
                        log_s_sum=np.array([float(t.double().sum()) for t in log_s_list]),
                        log_s_abs_sum=np.array([float(t.double().abs().sum()) for t in log_s_list]),
                        log_det_w=np.array([float(t) for t in log_det_W_list]), loss=np.float64(loss),
                        z_sq_sum=np.float64((z.double() ** 2).sum()), z_abs_max=np.float32(z.abs().max()),
                        log_s0_strided=log_s_list[0].reshape(B, -1)[:, ::STRIDE].contiguous().numpy(),
                        log_s11_strided=log_s_list[11].reshape(B, -1)[:, ::STRIDE].contiguous().numpy(),
                        **decimate(zf))
    print(f"{name}: z {tuple(z.shape)} loss {float(loss):.6f}; oracle bit-identical", flush=True)


def reference_infer_with_noise(model, mel, sigma, zs):
  """`WaveGlow.infer` of the reference (model.py:223-274) with its three noise draws replaced by the given tensors: the
  method draws `torch.FloatTensor(...).normal_()` (model.py:239-243, 266-270); `normal_` is patched to copy from the queue for
  the duration of the call, everything else is the reference's code."""
  queue = list(zs)
  orig = torch.Tensor.normal_

  def fake_normal_(self, *a, **k):
    src = queue.pop(0)
    assert tuple(self.shape) == tuple(src.shape), (tuple(self.shape), tuple(src.shape))
    return self.copy_(src)

  torch.Tensor.normal_ = fake_normal_
  try:
    out = model.infer(mel, sigma=sigma)
  finally:
    torch.Tensor.normal_ = orig
  assert not queue
  return out


if __name__ == "__main__":
  main(set(sys.argv[1:]))
