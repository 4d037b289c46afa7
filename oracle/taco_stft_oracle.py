"""CPU restatement of the reference's mel front-end  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Follows /root/reference/src/waveglow/taco_stft.py:84-105 (`TacotronSTFT.mel_spectrogram`) and the conv-basis STFT it
calls (stft.py:100-159): reflect padding by filter_length / 2, conv1d with the hann-windowed real / imaginary Fourier
basis at stride hop_length, magnitude, mel matmul, log(clamp(., 1e-5)).

Pinning: tools/make_golden.py runs the reference's own `TacotronSTFT` class (through tools/ref_loader.py) and writes
tests/golden/taco_stft_*.npz; this restatement must reproduce it bit for bit.  ONE ingredient is not the reference's:
`librosa.filters.mel` (librosa 0.10.1 in the reference's Pipfile.lock) is a third-party dependency that is absent from
this image, so `librosa_mel` below restates its published algorithm (Slaney scale, area normalisation) and is injected
into the reference run; it is cross-checked against the independent `transformers.audio_utils.mel_filter_bank`
(tests/test_workflow_cpu.py).  The filterbank itself is therefore "parity unpinned" against librosa proper.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.signal import get_window


def librosa_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kwargs):
  """librosa.filters.mel(htk=False, norm='slaney', dtype=float32), restated (see the header)."""
  if fmax is None:
    fmax = sr / 2.0
  f_sp = 200.0 / 3
  min_log_hz = 1000.0
  min_log_mel = min_log_hz / f_sp
  logstep = np.log(6.4) / 27.0

  def hz_to_mel(f):
    f = np.asanyarray(f, dtype=float)
    m = f / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, m)

  def mel_to_hz(m):
    m = np.asanyarray(m, dtype=float)
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

  weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
  fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
  mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
  fdiff = np.diff(mel_f)
  ramps = np.subtract.outer(mel_f, fftfreqs)
  for i in range(n_mels):
    lower = -ramps[i] / fdiff[i]
    upper = ramps[i + 2] / fdiff[i + 1]
    weights[i] = np.maximum(0, np.minimum(lower, upper))
  enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
  weights *= enorm[:, np.newaxis]
  return weights


def stft_magnitude(y: torch.Tensor, filter_length: int = 1024, hop_length: int = 256) -> torch.Tensor:
  """stft.py:100-159 with win_length == filter_length: y [B, T] -> |STFT| [B, filter_length / 2 + 1, 1 + T / hop]."""
  fourier_basis = np.fft.fft(np.eye(filter_length))
  cutoff = int(filter_length / 2 + 1)
  fourier_basis = np.vstack([np.real(fourier_basis[:cutoff, :]), np.imag(fourier_basis[:cutoff, :])])
  forward_basis = torch.FloatTensor(fourier_basis[:, None, :])
  fft_window = torch.from_numpy(get_window("hann", filter_length, fftbins=True)).float()
  forward_basis = forward_basis * fft_window
  x = y.view(y.size(0), 1, y.size(1))
  x = F.pad(x.unsqueeze(1), (int(filter_length / 2), int(filter_length / 2), 0, 0), mode="reflect").squeeze(1)
  ft = F.conv1d(x, forward_basis, stride=hop_length, padding=0)
  real, imag = ft[:, :cutoff, :], ft[:, cutoff:, :]
  return torch.sqrt(real ** 2 + imag ** 2)


def mel_spectrogram(y: torch.Tensor, sr: int = 22050, filter_length: int = 1024, hop_length: int = 256, n_mels: int = 80,
                    fmin: float = 0.0, fmax: float = 8000.0) -> torch.Tensor:
  """taco_stft.py:84-105."""
  basis = torch.from_numpy(librosa_mel(sr=sr, n_fft=filter_length, n_mels=n_mels, fmin=fmin, fmax=fmax)).float()
  return torch.log(torch.clamp(torch.matmul(basis, stft_magnitude(y, filter_length, hop_length)), min=1e-5))
