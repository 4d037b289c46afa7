"""CPU: the widened rows that do not need a GPU -- STFT/Denoiser arithmetic against the
golden vectors produced by the reference's own STFT and Denoiser classes, audio post-processing, checkpoint format, CLI parsing."""
import dataclasses
import os
import sys

import numpy as np
import pytest
import torch

from oracle.denoiser_oracle import StftOracle, denoise
from tests.helpers import load_golden
from waveglow_b200 import audio_utils as au
from waveglow_b200.checkpoint import CheckpointWaveglow
from waveglow_b200.cli import build_parser, split_hparams_string
from waveglow_b200.denoiser import STFT, Denoiser
from waveglow_b200.hparams import HParams


def test_denoiser_oracle_matches_reference_golden():
  g = load_golden("denoiser_b2_n6144")
  so = StftOracle(1024, 256, 1024)
  audio = torch.from_numpy(g["audio"])
  mag, _ = so.transform(audio)
  assert float((mag - torch.from_numpy(g["magnitude"])).abs().max()) <= 1e-4
  bias = torch.from_numpy(g["bias_spec"])
  for i, s in enumerate(g["strengths"]):
    out = denoise(so, audio, bias, float(s))
    assert float((out - torch.from_numpy(g[f"denoised_{i}"])).abs().max()) <= 2e-6


def test_fft_stft_and_denoiser_match_reference_golden():
  """The product's FFT-based STFT / Denoiser against outputs of the reference's conv-basis classes (stft.py, denoiser.py)."""
  g = load_golden("denoiser_b2_n6144")
  audio = torch.from_numpy(g["audio"])
  stft = STFT(filter_length=1024, hop_length=256, win_length=1024)
  mag, phase = stft.transform(audio)
  ref_mag = torch.from_numpy(g["magnitude"])
  assert mag.shape == ref_mag.shape
  assert float((mag - ref_mag).abs().max()) <= 1e-4 * max(1.0, float(ref_mag.max()))
  recon = stft.inverse(mag, phase)
  assert recon.shape == torch.from_numpy(g["recon"]).shape
  assert float((recon - torch.from_numpy(g["recon"])).abs().max()) <= 1e-5

  class FakeWaveglow:
    upsample = torch.nn.ConvTranspose1d(1, 1, 4)

    def infer(self, mel, sigma=1.0):
      assert tuple(mel.shape) == (1, 80, 88) and sigma == 0.0
      return torch.from_numpy(g["bias_audio"])

  den = Denoiser(FakeWaveglow(), HParams(), mode="zeros", device=torch.device("cpu"))
  ref_bias = torch.from_numpy(g["bias_spec"])
  assert den.bias_spec.shape == ref_bias.shape
  assert float((den.bias_spec - ref_bias).abs().max()) <= 1e-4 * max(1.0, float(ref_bias.max()))
  for i, s in enumerate(g["strengths"]):
    out = den(audio, strength=float(s))
    ref = torch.from_numpy(g[f"denoised_{i}"])
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-5, f"strength {s}"
  with pytest.raises(Exception):
    Denoiser(FakeWaveglow(), HParams(), mode="bogus", device=torch.device("cpu"))


def test_audio_postprocessing_semantics():
  rng = np.random.default_rng(0)
  wav = (rng.standard_normal(4096) * 0.2).astype(np.float32)
  n = au.normalize_wav(wav)
  assert n.dtype == np.float32 and np.max(np.abs(n)) == 1.0 and not au.is_overamp(n)
  assert au.normalize_wav(np.zeros(16, np.float32)).max() == 0            # silence: no division by zero
  pcm = au.convert_wav(n, np.int16)
  assert pcm.dtype == np.int16 and np.max(np.abs(pcm)) == 32767
  assert np.array_equal(pcm, np.round(n * 32767).astype(np.int16))
  assert au.is_overamp(np.array([0.5, 1.5], np.float32)) and not au.is_overamp(np.array([-1.0, 1.0], np.float32))
  # device path == numpy path, per utterance
  batch = torch.from_numpy(np.stack([wav, wav * 3.0, np.zeros_like(wav)]))
  dev = au.normalize_to_int16_device(batch).numpy()
  for i in range(3):
    assert np.array_equal(dev[i], au.convert_wav(au.normalize_wav(batch[i].numpy()), np.int16))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/waveglow"), reason="reference tree only exists in the build container")
def test_audio_postprocessing_matches_live_reference():
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
  from ref_loader import load_reference_module
  ref = load_reference_module("audio_utils")
  rng = np.random.default_rng(1)
  for scale in (0.01, 0.3, 2.5):
    wav = (rng.standard_normal(2000) * scale).astype(np.float32)
    assert np.array_equal(ref.normalize_wav(wav), au.normalize_wav(wav))
    assert np.array_equal(ref.convert_wav(ref.normalize_wav(wav), np.int16), au.convert_wav(au.normalize_wav(wav), np.int16))
    assert ref.is_overamp(wav) == au.is_overamp(wav)


def test_checkpoint_file_format_round_trip(tmp_path):
  hp = HParams(n_channels=256)
  sd = {"upsample.bias": torch.arange(80, dtype=torch.float32)}
  ck = CheckpointWaveglow.from_state_dict(sd, hp, iteration=7)
  path = tmp_path / "7.pt"
  ck.save(path)
  raw = torch.load(path, weights_only=False)
  assert set(raw.keys()) == {"state_dict", "optimizer", "learning_rate", "iteration", "hparams"}     # checkpoint.py:13-20
  assert raw["hparams"] == dataclasses.asdict(hp)
  # a file as the reference writes it, with an hparam this build does not know: dropped with a warning (checkpoint.py:22-28)
  raw["hparams"]["fp16_run"] = False
  torch.save(raw, path)
  back = CheckpointWaveglow.load(path, torch.device("cpu"))
  assert back.iteration == 7 and back.get_hparams() == hp
  assert torch.equal(back.state_dict["upsample.bias"], sd["upsample.bias"])


def test_cli_arguments_mirror_the_reference():
  p = build_parser()
  ns = p.parse_args(["synthesize", "ckpt.pt", "mels"])
  assert ns.sigma == 1.0 and ns.denoiser_strength == 0.0005 and ns.custom_seed is None and not ns.overwrite
  ns = p.parse_args(["synthesize", "ckpt.pt", "mels", "--sigma", "0.666", "--custom-seed", "3", "-o", "-out", "wavs",
                     "--custom-hparams", "n_channels=512,n_flows=12"])
  assert ns.sigma == 0.666 and ns.custom_seed == 3 and ns.overwrite and str(ns.output_directory) == "wavs"
  assert split_hparams_string(ns.custom_hparams) == {"n_channels": "512", "n_flows": "12"}
  with pytest.raises(SystemExit):
    p.parse_args(["synthesize", "ckpt.pt", "mels", "--sigma", "1.5"])      # validated to [0, 1] like the reference
