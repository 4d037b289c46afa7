"""GPU: the caller rows around the hot path -- Synthesizer / Denoiser / batched CLI -- against the
oracle (reference arithmetic on CPU) on identical mel, weights and noise."""
import numpy as np
import pytest
import torch
from scipy.io.wavfile import read as read_wav

from oracle import waveglow_oracle as O
from oracle.denoiser_oracle import StftOracle, denoise
from tests.helpers import FP16_MAX_ABS, FP32_MAX_ABS, max_abs, synthetic_weights
from waveglow_b200 import audio_utils as au
from waveglow_b200.checkpoint import CheckpointWaveglow
from waveglow_b200.cli import main as cli_main
from waveglow_b200.synthesizer import InferenceResult, Synthesizer, init_global_seeds
from waveglow_b200.synthetic import synthetic_mel

pytestmark = pytest.mark.gpu

_SYNTH = {}


def get_synth(precision):
  if precision not in _SYNTH:
    hp, sd, w, dims = synthetic_weights(256)
    ck = CheckpointWaveglow.from_state_dict(sd, hp, iteration=1)
    _SYNTH[precision] = Synthesizer(ck, device=torch.device("cuda:0"), precision=precision)
  return _SYNTH[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_synthesizer_matches_reference_arithmetic(precision):
  """Synthesizer.infer (synthesizer.py:54-94): wav and wav_denoised vs oracle infer + oracle (conv-basis) denoiser."""
  hp, sd, w, dims = synthetic_weights(256)
  synth = get_synth(precision)
  mel = synthetic_mel(1, 30, 41)
  seed, sigma, strength = 5, 0.9, 0.05
  res = synth.infer(mel.cuda(), sigma=sigma, denoiser_strength=strength, seed=seed)
  assert isinstance(res, InferenceResult) and res.sampling_rate == 22050
  assert res.wav.shape == (30 * 256,) and res.wav.dtype == np.float32 and res.wav_denoised.shape == (30 * 256,)
  assert res.inference_duration_s > 0 and res.denoising_duration_s > 0
  # the same noise the engine drew: re-seed and draw on the device in the reference's order
  init_global_seeds(seed)
  zs = [z.cpu() for z in synth.model.draw_noise(1, 30 * 32, torch.device("cuda:0"))]
  with torch.no_grad():
    ref = O.infer(w, dims, mel, sigma, zs)
    bias_audio = O.infer(w, dims, torch.zeros(1, 80, 88), 0.0, [torch.zeros(1, 4, 88 * 32), torch.zeros(1, 2, 88 * 32), torch.zeros(1, 2, 88 * 32)])
    so = StftOracle(hp.filter_length, hp.hop_length, hp.win_length)
    bias_spec = so.transform(bias_audio)[0][:, :, 0][:, :, None]
    ref_den = denoise(so, ref, bias_spec, strength)
  tol = FP32_MAX_ABS if precision == "fp32" else FP16_MAX_ABS
  e1 = max_abs(ref[0], torch.from_numpy(res.wav))
  e2 = max_abs(ref_den[0, 0], torch.from_numpy(res.wav_denoised))
  print(f"synthesizer {precision}: wav max-abs {e1:.3e}, wav_denoised max-abs {e2:.3e}, was_overamplified {res.was_overamplified}")
  assert e1 <= tol and e2 <= tol
  assert res.was_overamplified == bool(np.abs(res.wav).max() > 1.0)
  # same seed -> same audio (seed semantics of synthesizer.py:56)
  res2 = synth.infer(mel.cuda(), sigma=sigma, denoiser_strength=strength, seed=seed)
  assert np.array_equal(res.wav, res2.wav)


def test_batched_path_equals_one_by_one_loop():
  """infer_batch (on-device normalise + int16) == reference CLI loop: infer -> normalize_wav -> float_to_wav per mel."""
  synth = get_synth("fp16")
  mels = [synthetic_mel(1, 20, 50 + i)[0] for i in range(3)]
  pcms = synth.infer_batch(mels, sigma=0.8, denoiser_strength=0.0005, seed=9)
  for i, m in enumerate(mels):
    r = synth.infer(m[None].cuda(), sigma=0.8, denoiser_strength=0.0005, seed=9)
    expect = au.convert_wav(au.normalize_wav(r.wav_denoised), np.int16)
    diff = np.abs(pcms[i].astype(np.int32) - expect.astype(np.int32)).max()
    assert pcms[i].dtype == np.int16 and pcms[i].shape == expect.shape
    assert diff <= 1, f"utterance {i}: int16 difference {diff}"      # identical up to one rounding tie


def test_cli_synthesize_end_to_end(tmp_path):
  hp, sd, w, dims = synthetic_weights(256)
  ckpt = tmp_path / "1.pt"
  CheckpointWaveglow.from_state_dict(sd, hp, iteration=1).save(ckpt)
  folder = tmp_path / "mels"
  (folder / "sub").mkdir(parents=True)
  lengths = {"a.npy": 12, "b.npy": 12, "sub/c.npy": 17}
  for name, frames in lengths.items():
    np.save(folder / name, synthetic_mel(1, frames, 60 + frames)[0].numpy())
  out = tmp_path / "wavs"
  assert cli_main(["synthesize", str(ckpt), str(folder), "-out", str(out), "--custom-seed", "3", "--sigma", "0.666"]) == 0
  for name, frames in lengths.items():
    sr, pcm = read_wav(out / name.replace(".npy", ".wav"))
    assert sr == 22050 and pcm.dtype == np.int16 and pcm.shape == (frames * 256,)
    assert np.abs(pcm.astype(np.int32)).max() == 32767      # peak-normalised like normalize_wav
  # without --overwrite nothing is re-synthesised; a missing checkpoint is reported, not raised (inference_v2.py:92-96)
  before = (out / "a.wav").stat().st_mtime_ns
  assert cli_main(["synthesize", str(ckpt), str(folder), "-out", str(out), "--custom-seed", "3"]) == 0
  assert (out / "a.wav").stat().st_mtime_ns == before
  assert cli_main(["synthesize", str(tmp_path / "nope.pt"), str(folder)]) == 1


def test_mel_front_end_on_device_matches_reference_golden():
  """taco_stft.TacotronSTFT on cuda:0 (cuFFT + mel matmul + log-clamp) against the reference class's golden mels."""
  from tests.helpers import load_golden
  from waveglow_b200.taco_stft import TacotronSTFT, TSTFTHParams
  g = load_golden("taco_stft_b2_n8192")
  tstft = TacotronSTFT(TSTFTHParams(), torch.device("cuda:0"))
  mel = tstft.mel_spectrogram(torch.from_numpy(g["audio"]).cuda())
  ref = torch.from_numpy(g["mel"])
  assert mel.is_cuda and mel.shape == ref.shape
  err = float((mel.cpu() - ref).abs().max())
  print(f"on-device mel front-end: max-abs (log domain) {err:.3e}")
  assert err <= 2e-3


def test_cli_synthesize_wav_equals_synthesize_of_its_mels(tmp_path):
  """`synthesize-wav` (inference_wav.py:74-130) = mel front-end + `synthesize`: feeding the mels it computes to
  `synthesize` as .npy files gives the same wavs."""
  from scipy.io.wavfile import write as write_wav
  from waveglow_b200.taco_stft import TacotronSTFT, TSTFTHParams
  hp, sd, w, dims = synthetic_weights(256)
  ckpt = tmp_path / "1.pt"
  CheckpointWaveglow.from_state_dict(sd, hp, iteration=1).save(ckpt)
  src, mels = tmp_path / "src", tmp_path / "mels"
  src.mkdir()
  mels.mkdir()
  tstft = TacotronSTFT(TSTFTHParams(), torch.device("cuda:0"))
  g = torch.Generator().manual_seed(77)
  for name, n in (("x.wav", 256 * 11), ("y.wav", 256 * 11), ("z.wav", 256 * 15 + 100)):
    pcm = (torch.randn(n, generator=g).clamp(-3, 3) / 3 * 20000).to(torch.int16).numpy()
    write_wav(str(src / name), 22050, pcm)
    np.save(mels / name.replace(".wav", ".npy"), tstft.get_mel_tensor_from_file(src / name).cpu().numpy())
  out_a, out_b = tmp_path / "a", tmp_path / "b"
  assert cli_main(["synthesize-wav", str(ckpt), str(src), "-out", str(out_a), "--custom-seed", "4", "--sigma", "0.7"]) == 0
  assert cli_main(["synthesize", str(ckpt), str(mels), "-out", str(out_b), "--custom-seed", "4", "--sigma", "0.7"]) == 0
  for name, frames in (("x.wav", 12), ("y.wav", 12), ("z.wav", 16)):
    sr, a = read_wav(out_a / name)
    _, b = read_wav(out_b / name)
    assert sr == 22050 and a.shape == (frames * 256,) and np.array_equal(a, b)


def test_packed_weight_cache_round_trip(tmp_path):
  """the engine's packed arrays written by one Synthesizer are restored by the next one (bit-identical
  audio, no fold / re-pack); other weights, another precision or a damaged file fall back to packing from the state_dict."""
  import time
  hp, sd, w, dims = synthetic_weights(256)
  ck = CheckpointWaveglow.from_state_dict(sd, hp, iteration=1)
  cache = tmp_path / "1.pt.b200pack"
  mel = synthetic_mel(2, 9, 91).cuda()
  dev = torch.device("cuda:0")

  def build(**kw):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = Synthesizer(ck, device=dev, **kw)
    torch.cuda.synchronize()
    return s, time.perf_counter() - t0

  a, t_a = build(precision="fp16", pack_cache=cache, pack_fingerprint=123)
  assert a.model.pack_cache_hit is False and cache.is_file()
  b, t_b = build(precision="fp16", pack_cache=cache, pack_fingerprint=123)
  assert b.model.pack_cache_hit is True
  print(f"Synthesizer start: {t_a:.2f} s packing from the state_dict, {t_b:.2f} s from the cache ({cache.stat().st_size / 1e6:.0f} MB)")
  init_global_seeds(3)
  out_a = a.model.infer(mel, sigma=0.7)
  init_global_seeds(3)
  out_b = b.model.infer(mel, sigma=0.7)
  assert torch.equal(out_a, out_b)
  assert torch.equal(a.denoiser.bias_spec, b.denoiser.bias_spec)
  # forward direction works from the cache too (w_fwd, log-dets)
  audio_in = torch.rand(2, 2048, device=dev) * 2 - 1
  za = a.model((mel[:, :, :8], audio_in))
  zb = b.model((mel[:, :, :8], audio_in))
  assert torch.equal(za[0], zb[0]) and all(torch.equal(x, y) for x, y in zip(za[2], zb[2]))
  # other weights (fingerprint), another precision, a truncated file: rejected, packed as usual, cache rewritten
  c, _ = build(precision="fp16", pack_cache=cache, pack_fingerprint=124)
  assert c.model.pack_cache_hit is False
  d, _ = build(precision="fp32", pack_cache=cache, pack_fingerprint=124)
  assert d.model.pack_cache_hit is False
  e, _ = build(precision="fp32", pack_cache=cache, pack_fingerprint=124)
  assert e.model.pack_cache_hit is True
  init_global_seeds(3)
  out_d = d.model.infer(mel, sigma=0.7)
  init_global_seeds(3)
  assert torch.equal(out_d, e.model.infer(mel, sigma=0.7))
  data = cache.read_bytes()
  cache.write_bytes(data[: len(data) // 2])
  f, _ = build(precision="fp32", pack_cache=cache, pack_fingerprint=124)
  assert f.model.pack_cache_hit is False and cache.stat().st_size == len(data)


def test_fused_denoiser_kernel_matches_reference_golden():
  """Denoiser on cuda:0 = cuFFT STFT -> wg_denoise_spectrum (one in-place kernel: X * max(|X| - s b, 0) / |X|) -> cuFFT inverse,
  against the outputs of the reference's own conv-basis STFT / Denoiser classes (denoiser.py:51-57, stft.py:135-198) for three
  strengths incl. one that clamps most bins to zero; gate 1e-5.  Also the raw kernel against the torch
  formula on a non-contiguous spectrum view."""
  from tests.helpers import load_golden
  from waveglow_b200.denoiser import Denoiser
  from waveglow_b200.engine import denoise_spectrum_
  from waveglow_b200.hparams import HParams
  g = load_golden("denoiser_b2_n6144")
  dev = torch.device("cuda:0")

  class FakeWaveglow:
    upsample = torch.nn.ConvTranspose1d(1, 1, 4).to(dev)

    def infer(self, mel, sigma=1.0):
      assert tuple(mel.shape) == (1, 80, 88) and sigma == 0.0 and mel.is_cuda
      return torch.from_numpy(g["bias_audio"]).to(dev)

  den = Denoiser(FakeWaveglow(), HParams(), mode="zeros", device=dev)
  audio = torch.from_numpy(g["audio"]).to(dev)
  for i, s in enumerate(g["strengths"]):
    out = den(audio, strength=float(s))
    torch.cuda.synchronize()
    ref = torch.from_numpy(g[f"denoised_{i}"])
    err = float((out.cpu() - ref).abs().max())
    print(f"fused denoiser, strength {float(s):g}: max-abs {err:.3e}")
    assert out.is_cuda and out.shape == ref.shape and err <= 1e-5
  gen = torch.Generator().manual_seed(3)
  spec = torch.complex(torch.randn((3, 40, 513), generator=gen), torch.randn((3, 40, 513), generator=gen)).to(dev).transpose(1, 2)
  spec[0, 5, 7] = 0                                          # |X| == 0 stays 0 (atan2(0, 0) = 0 in the reference)
  bias = torch.rand(513, generator=gen).to(dev)
  mag = spec.abs()
  expect = spec * torch.where(mag > 0, torch.clamp(mag - bias[None, :, None] * 0.7, min=0) / mag.clamp_min(1e-30), torch.zeros_like(mag))
  got = denoise_spectrum_(spec.clone(memory_format=torch.preserve_format), bias, 0.7)
  assert not got.is_contiguous() and float((got - expect).abs().max()) <= 1e-6 and complex(got[0, 5, 7]) == 0
