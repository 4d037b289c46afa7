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
