"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build container.

  python tools/make_golden.py

Each fixture holds: the mel input, the pre-drawn noise tensors (drawn from torch's global CPU generator in
the order the reference draws them), the reference's `WaveGlow.infer` output, and the recipe for the
weights (`synthetic_state_dict(hparams, weight_seed)` -- deterministic, so the 352 MB checkpoint is not
committed).  The script asserts that oracle/waveglow_oracle.py reproduces the reference bit-for-bit before
writing anything, which is what pins the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import waveglow_oracle as O  # noqa: E402
from ref_loader import load_reference, load_reference_module  # noqa: E402
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_state_dict  # noqa: E402

CASES = [
  # name, n_channels, batch, frames, sigma, noise seed, mel seed, weight style
  ("infer_c256_b2_f12", 256, 2, 12, 0.666, 11, 1001, "parametrized"),
  ("infer_c256_b1_f40", 256, 1, 40, 1.0, 12, 1002, "parametrized"),
  ("infer_c256_b3_f7_ragged", 256, 3, 7, 0.666, 13, 1003, "legacy"),
  ("infer_c256_b1_f11_sigma0", 256, 1, 11, 0.0, 14, 1004, "parametrized"),
  ("infer_c512_b1_f10", 512, 1, 10, 0.666, 15, 1005, "parametrized"),
]
FWD_CASES = [
  # name, n_channels, batch, n_samples, frames, seed
  ("forward_c256_b2_t2048", 256, 2, 2048, 8, 21),
  ("forward_c256_b1_t4000", 256, 1, 4000, 16, 22),
]
WEIGHT_SEED = 1234


def main():
  WaveGlow, RefHParams = load_reference()
  out_dir = os.path.join(ROOT, "tests", "golden")
  os.makedirs(out_dir, exist_ok=True)
  models = {}

  def get_model(nc, style):
    key = (nc, style)
    if key not in models:
      hp = HParams(n_channels=nc)
      sd = synthetic_state_dict(hp, WEIGHT_SEED, style=style)
      m = WaveGlow(RefHParams(n_channels=nc))
      missing = m.load_state_dict(sd)
      assert not missing.missing_keys and not missing.unexpected_keys, missing
      m = WaveGlow.remove_weightnorm(m).eval()
      models[key] = (m, O.fold_weight_norm(sd), O.ModelDims(n_channels=nc))
    return models[key]

  for name, nc, B, F, sigma, seed, mel_seed, style in CASES:
    m, w, dims = get_model(nc, style)
    mel = synthetic_mel(B, F, mel_seed)
    torch.manual_seed(seed)
    with torch.no_grad():
      ref = m.infer(mel, sigma=sigma)
    torch.manual_seed(seed)
    zs = O.draw_noise(dims, B, F * 32)
    with torch.no_grad():
      mine = O.infer(w, dims, mel, sigma, zs)
    assert torch.equal(ref, mine), f"{name}: oracle is not bit-identical to the reference"
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), mel=mel.numpy(), audio=ref.numpy(), sigma=np.float32(sigma),
                        n_channels=np.int32(nc), weight_seed=np.int32(WEIGHT_SEED), weight_style=np.array(style),
                        **{f"z{i}": z.numpy() for i, z in enumerate(zs)})
    print(f"{name}: audio {tuple(ref.shape)} absmax {ref.abs().max():.3f}  (oracle bit-identical)")

  for name, nc, B, T, F, seed in FWD_CASES:
    m, w, dims = get_model(nc, "parametrized")
    mel = synthetic_mel(B, F, 2000 + seed)
    g = torch.Generator().manual_seed(seed)
    audio = torch.rand((B, T), generator=g) * 2 - 1
    with torch.no_grad():
      z, log_s_list, log_det_W_list = m((mel, audio))
      z2, ls2, ld2 = O.forward(w, dims, mel, audio)
    assert torch.equal(z, z2) and all(torch.equal(a, b) for a, b in zip(log_s_list, ls2))
    # fp32 logdet goes through LAPACK's LU and depends on the operand's memory layout at the last bits;
    # log_det_W = B * T' * logdet(W) amplifies that by B * T'
    atol = 2e-6 * B * (T // 8)
    assert all(torch.allclose(a, b, rtol=1e-5, atol=atol) for a, b in zip(log_det_W_list, ld2))
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), mel=mel.numpy(), audio_in=audio.numpy(), z=z.numpy(),
                        log_s_sum=np.array([float(t.double().sum()) for t in log_s_list]),
                        log_det_w=np.array([float(t) for t in log_det_W_list]),
                        loss=np.float64(O.loss(z, log_s_list, log_det_W_list, 1.0)),
                        n_channels=np.int32(nc), weight_seed=np.int32(WEIGHT_SEED),
                        **{f"log_s{k}": t.numpy() for k, t in enumerate(log_s_list)})
    print(f"{name}: z {tuple(z.shape)}  (oracle bit-identical)")


def make_denoiser_golden():
  """Runs the reference's own STFT and Denoiser classes (stft.py, denoiser.py) on CPU."""
  from oracle.denoiser_oracle import StftOracle, denoise
  ref_stft = load_reference_module("stft")
  ref_den = load_reference_module("denoiser")
  _, RefHParams = load_reference()
  hp = RefHParams()
  g = torch.Generator().manual_seed(31)
  bias_audio = torch.randn((1, 88 * 256), generator=g) * 0.01
  audio = torch.randn((2, 256 * 24), generator=g) * 0.3

  class FakeWaveglow:          # Denoiser only touches .upsample.weight.{dtype,device} and .infer (denoiser.py:29-47)
    upsample = torch.nn.ConvTranspose1d(1, 1, 4)

    def infer(self, mel, sigma=1.0):
      assert tuple(mel.shape) == (1, 80, 88) and float(mel.abs().sum()) == 0.0 and sigma == 0.0
      return bias_audio.clone()

  dev = torch.device("cpu")
  den = ref_den.Denoiser(waveglow=FakeWaveglow(), hparams=hp, mode="zeros", device=dev)
  stft = ref_stft.STFT(filter_length=hp.filter_length, hop_length=hp.hop_length, win_length=hp.win_length, device=dev)
  mag, phase = stft.transform(audio)
  recon = stft.inverse(mag, phase)
  outs = {f"denoised_{i}": den(audio, strength=s) for i, s in enumerate((0.0005, 0.1, 1.0))}
  so = StftOracle(hp.filter_length, hp.hop_length, hp.win_length)
  m2, p2 = so.transform(audio)
  assert torch.equal(mag, m2) and torch.equal(phase, p2), "denoiser oracle STFT is not bit-identical to the reference"
  for i, s in enumerate((0.0005, 0.1, 1.0)):
    assert torch.equal(outs[f"denoised_{i}"], denoise(so, audio, den.bias_spec, s)), "denoiser oracle mismatch"
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "denoiser_b2_n6144.npz"), audio=audio.numpy(),
                      bias_audio=bias_audio.numpy(), bias_spec=den.bias_spec.numpy(), magnitude=mag.numpy(),
                      recon=recon.numpy(), strengths=np.array([0.0005, 0.1, 1.0], dtype=np.float32),
                      **{k: v.numpy() for k, v in outs.items()})
  print(f"denoiser_b2_n6144: bias_spec {tuple(den.bias_spec.shape)}, denoised {tuple(outs['denoised_0'].shape)} (oracle bit-identical)")


def make_taco_stft_golden():
  """Runs the reference's own TacotronSTFT class (taco_stft.py) on CPU.  Its one third-party ingredient,
  librosa.filters.mel, is absent from this image: oracle/taco_stft_oracle.py::librosa_mel (a restatement of the published
  algorithm) is injected in its place (see that file's header: the filterbank is not pinned against librosa proper)."""
  from oracle import taco_stft_oracle as TO
  load_reference()                                   # registers the stub modules (incl. librosa.filters)
  sys.modules["librosa.filters"].mel = TO.librosa_mel
  sys.modules.pop("waveglow.taco_stft", None)        # re-import so that `from librosa.filters import mel` binds the injection
  ref_taco = load_reference_module("taco_stft")
  hp = ref_taco.TSTFTHParams()
  tstft = ref_taco.TacotronSTFT(hp, torch.device("cpu"))
  g = torch.Generator().manual_seed(41)
  n = 8192
  t = torch.arange(n) / hp.sampling_rate
  # two "utterances": a chirp-like tone mix with noise, and low-level noise (exercises the 1e-5 clamp)
  y0 = 0.4 * torch.sin(2 * np.pi * (220.0 + 900.0 * t) * t) + 0.2 * torch.sin(2 * np.pi * 3100.0 * t) + 0.05 * torch.randn(n, generator=g)
  y1 = 1e-6 * torch.randn(n, generator=g)
  y = torch.stack([y0, y1]).clamp(-1.0, 1.0)
  mel = tstft.mel_spectrogram(y)
  mel_single = tstft.get_mel_tensor(y[0])
  assert torch.equal(mel, TO.mel_spectrogram(y)), "taco_stft oracle is not bit-identical to the reference class"
  assert torch.allclose(mel_single, mel[0], atol=1e-4)      # batch 1 vs 2 may take another conv1d path
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "taco_stft_b2_n8192.npz"), audio=y.numpy(), mel=mel.numpy(),
                      mel_basis=tstft.mel_basis.numpy())
  print(f"taco_stft_b2_n8192: mel {tuple(mel.shape)}, range [{float(mel.min()):.3f}, {float(mel.max()):.3f}] (oracle bit-identical)")


if __name__ == "__main__":
  if "--denoiser-only" not in sys.argv and "--taco-only" not in sys.argv:
    main()
  if "--taco-only" not in sys.argv:
    make_denoiser_golden()
  make_taco_stft_golden()
