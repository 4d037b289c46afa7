"""Mel front-end of `synthesize-wav` on the device.

Mirrors /root/reference/src/waveglow/taco_stft.py:36-125: `TacotronSTFT(hparams, device)` with `mel_spectrogram`,
`get_mel_tensor`, `get_mel_tensor_from_file`, and the `wav_to_float32` loader of audio_utils.py:206-216.  The reference
computes the magnitude STFT as a conv1d with a [1026, 1, 1024] windowed Fourier basis (stft.py:135-159); here it is one
batched cuFFT call (`denoiser.STFT`, pinned against the reference's STFT class), followed by the mel matmul and
`log(clamp(x, 1e-5))` on the device, so a wav goes audio -> mel -> `WaveGlow.infer` without leaving the GPU.

The mel filterbank comes from `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` in the reference (taco_stft.py:67-73,
librosa 0.10.1 in its Pipfile.lock; defaults htk=False, norm="slaney").  librosa is not a dependency of this package:
`mel_filterbank` below restates that published algorithm; tests pin it to an independent implementation
(transformers.audio_utils.mel_filter_bank) and to the golden mels of the reference's own TacotronSTFT class.
"""
from dataclasses import dataclass
from logging import getLogger

import numpy as np
import torch
from scipy.io.wavfile import read

from .denoiser import STFT

FLOAT32_64_MIN_WAV = -1.0      # audio_utils.py:246-247
FLOAT32_64_MAX_WAV = 1.0


@dataclass
class STFTHParams:
  """taco_stft.py:36-42."""
  filter_length: int = 1024
  hop_length: int = 256
  win_length: int = 1024
  window: str = "hann"


@dataclass
class TSTFTHParams(STFTHParams):
  """taco_stft.py:45-50."""
  n_mel_channels: int = 80
  sampling_rate: int = 22050
  mel_fmin: float = 0.0
  mel_fmax: float = 8000.0


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
  """Slaney scale: linear below 1 kHz (200/3 Hz per mel), logarithmic above (27 mels per factor 6.4)."""
  f = np.asarray(f, dtype=np.float64)
  f_sp = 200.0 / 3
  mels = f / f_sp
  min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
  return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
  m = np.asarray(m, dtype=np.float64)
  f_sp = 200.0 / 3
  min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
  return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
  """[n_mels, 1 + n_fft/2] float32 triangular filters, centres equally spaced on the Slaney mel scale, each scaled by
  2 / (f_hi - f_lo) (area normalisation) -- librosa.filters.mel with its defaults."""
  fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
  mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
  fdiff = np.diff(mel_f)
  ramps = np.subtract.outer(mel_f, fftfreqs)
  weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
  for i in range(n_mels):
    lower = -ramps[i] / fdiff[i]
    upper = ramps[i + 2] / fdiff[i + 1]
    weights[i] = np.maximum(0, np.minimum(lower, upper))
  enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
  return (weights * enorm[:, None]).astype(np.float32)


def dynamic_range_compression(x: torch.Tensor, C: float = 1, clip_val: float = 1e-5) -> torch.Tensor:
  """taco_stft.py:10-16."""
  return torch.log(torch.clamp(x, min=clip_val) * C)


def wav_to_float32(path) -> tuple:
  """audio_utils.py:213-216 (+ convert_wav 53-64): PCM -> float32 in [-1, 1) by dividing through -min of the PCM type."""
  sampling_rate, wav = read(str(path))
  if wav.dtype != np.float32:
    if wav.dtype in (np.float32, np.float64):
      wav = (wav / 1.0 * FLOAT32_64_MAX_WAV).astype(np.float32)
    else:
      wav = (wav / (-1.0 * np.iinfo(wav.dtype).min) * FLOAT32_64_MAX_WAV).astype(np.float32)
  return wav, sampling_rate


class TacotronSTFT(torch.nn.Module):
  """Same constructor and methods as the reference's TacotronSTFT (taco_stft.py:53-125)."""

  def __init__(self, hparams: TSTFTHParams, device: torch.device):
    super().__init__()
    self.n_mel_channels = hparams.n_mel_channels
    self.sampling_rate = hparams.sampling_rate
    self.device = torch.device(device)
    self.stft_fn = STFT(device=self.device, filter_length=hparams.filter_length, hop_length=hparams.hop_length,
                        win_length=hparams.win_length, window=hparams.window)
    basis = mel_filterbank(hparams.sampling_rate, hparams.filter_length, hparams.n_mel_channels, hparams.mel_fmin, hparams.mel_fmax)
    self.register_buffer("mel_basis", torch.from_numpy(basis).float().to(self.device))

  def spectral_normalize(self, magnitudes: torch.Tensor) -> torch.Tensor:
    return dynamic_range_compression(magnitudes)

  def spectral_de_normalize(self, magnitudes: torch.Tensor) -> torch.Tensor:
    return torch.exp(magnitudes)

  def mel_spectrogram(self, y: torch.Tensor) -> torch.Tensor:
    """y [B, T] in [-1, 1] -> log-mel [B, n_mel_channels, 1 + T / hop] (taco_stft.py:84-105)."""
    y = y.to(self.device)
    if not (float(y.min()) >= FLOAT32_64_MIN_WAV and float(y.max()) <= FLOAT32_64_MAX_WAV):
      raise AssertionError("input is overamplified")        # taco_stft.py:95-97 asserts the same
    magnitudes, _ = self.stft_fn.transform(y)
    return self.spectral_normalize(torch.matmul(self.mel_basis, magnitudes))

  def get_wav_tensor_from_file(self, wav_path) -> torch.Tensor:
    wav, sampling_rate = wav_to_float32(wav_path)
    if sampling_rate != self.sampling_rate:
      msg = (f"{wav_path}: The sampling rate of the file ({sampling_rate}Hz) doesn't match the target sampling rate "
             f"({self.sampling_rate}Hz)!")
      getLogger(__name__).error(msg)
      raise ValueError(msg)                                  # taco_stft.py:110-114
    return torch.from_numpy(wav)

  def get_mel_tensor_from_file(self, wav_path) -> torch.Tensor:
    return self.get_mel_tensor(self.get_wav_tensor_from_file(wav_path))

  def get_mel_tensor(self, wav_tensor: torch.Tensor) -> torch.Tensor:
    return self.mel_spectrogram(wav_tensor.unsqueeze(0)).squeeze(0)
