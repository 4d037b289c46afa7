"""CPU oracle of the reference's STFT / Denoiser arithmetic  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates /root/reference/src/waveglow/stft.py:98-203 (basis convolution STFT, pseudo-inverse basis inverse,
window sum-square normalisation, stft.py:44-94) and denoiser.py:51-57 in numpy/torch, with the three librosa.util helpers
the reference imports (pad_center, tiny, normalize; librosa is not installed here) restated for the only case the reference
uses (win_length == filter_length, norm=None).  Pinned against the reference's own classes by tools/make_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.signal import get_window


def window_sumsquare(n_frames: int, hop_length: int, win_length: int, n_fft: int) -> np.ndarray:
  n = n_fft + hop_length * (n_frames - 1)
  x = np.zeros(n, dtype=np.float32)
  win_sq = get_window("hann", win_length, fftbins=True) ** 2      # normalize(norm=None) is the identity; pad_center no-op
  for i in range(n_frames):
    s = i * hop_length
    x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
  return x


class StftOracle:
  def __init__(self, filter_length=1024, hop_length=256, win_length=1024):
    assert win_length == filter_length
    self.n, self.hop = filter_length, hop_length
    scale = filter_length / hop_length
    fb = np.fft.fft(np.eye(filter_length))
    cutoff = filter_length // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff]), np.imag(fb[:cutoff])])
    win = torch.from_numpy(get_window("hann", win_length, fftbins=True)).float()
    self.forward_basis = torch.FloatTensor(fb[:, None, :]) * win
    self.inverse_basis = torch.FloatTensor(np.linalg.pinv(scale * fb).T[:, None, :]) * win

  def transform(self, x: torch.Tensor):
    B, n = x.shape
    x = F.pad(x.view(B, 1, 1, n), (self.n // 2, self.n // 2, 0, 0), mode="reflect").squeeze(1)
    ft = F.conv1d(x, self.forward_basis, stride=self.hop)
    c = self.n // 2 + 1
    re, im = ft[:, :c], ft[:, c:]
    return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)

  def inverse(self, mag: torch.Tensor, phase: torch.Tensor):
    rec = torch.cat([mag * torch.cos(phase), mag * torch.sin(phase)], 1)
    inv = F.conv_transpose1d(rec, self.inverse_basis, stride=self.hop)
    ws = window_sumsquare(mag.size(-1), self.hop, self.n, self.n)
    idx = torch.from_numpy(np.where(ws > np.finfo(np.float32).tiny)[0])
    inv[:, :, idx] /= torch.from_numpy(ws)[idx]
    inv *= float(self.n) / self.hop
    return inv[:, :, self.n // 2: -(self.n // 2)]


def denoise(stft: StftOracle, audio: torch.Tensor, bias_spec: torch.Tensor, strength: float) -> torch.Tensor:
  mag, ph = stft.transform(audio.float())
  return stft.inverse(torch.clamp(mag - bias_spec * strength, 0.0), ph)
