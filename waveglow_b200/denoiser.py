"""Denoiser: removes the model's bias spectrum from synthesised audio.

Mirrors /root/reference/src/waveglow/denoiser.py:14-57 and the STFT it uses (stft.py:98-203):
  * bias: `waveglow.infer(zeros[1, 80, 88], sigma=0.0)` once -> magnitude STFT -> first frame (denoiser.py:29-49);
  * per call: STFT -> `clamp(|X| - strength * bias, 0)` with the original phase -> inverse STFT (denoiser.py:51-57).

The reference evaluates the STFT as a conv1d with a [1026, 1, 1024] windowed Fourier basis and the inverse as a
conv_transpose1d with its pseudo-inverse, divided by the window sum-square envelope (stft.py:135-198).  For a hann
window with hop = n_fft / 4 that is exactly the textbook centred STFT / overlap-add inverse, so this module runs it as
batched FFTs on the device (`torch.stft` / `torch.istft`, cuFFT) -- O(N log N) instead of the reference's O(N^2)
basis convolutions, no host-side numpy window loop (stft.py:175-191) per call.  Parity with the reference's own STFT
and Denoiser classes is pinned by tests/golden/denoiser_*.npz (tools/make_golden.py).
"""
import torch

BIAS_MEL_LENGTH = 88   # denoiser.py:11


class STFT(torch.nn.Module):
  """transform(audio[B, n]) -> (magnitude, phase) [B, n_fft/2+1, 1 + n/hop]; inverse(magnitude, phase) -> [B, 1, n']."""

  def __init__(self, device: torch.device = None, filter_length: int = 800, hop_length: int = 200, win_length: int = 800,
               window: str = "hann"):
    super().__init__()
    if window != "hann":
      raise ValueError("only the hann window of the reference's hparams (taco_stft.py:36-42) is supported")
    if win_length != filter_length:
      raise ValueError("win_length must equal filter_length (as in the reference's hparams)")
    self.filter_length = filter_length
    self.hop_length = hop_length
    self.win_length = win_length
    # scipy.signal.get_window('hann', n, fftbins=True) == periodic hann
    self.register_buffer("window", torch.hann_window(win_length, periodic=True, dtype=torch.float32, device=device))

  def _complex(self, audio: torch.Tensor) -> torch.Tensor:
    return torch.stft(audio.float(), n_fft=self.filter_length, hop_length=self.hop_length, win_length=self.win_length,
                      window=self.window, center=True, pad_mode="reflect", normalized=False, onesided=True,
                      return_complex=True)

  def transform(self, input_data: torch.Tensor):
    spec = self._complex(input_data)
    return spec.abs(), torch.atan2(spec.imag, spec.real)        # stft.py:155-159

  def inverse_complex(self, spec: torch.Tensor) -> torch.Tensor:
    out = torch.istft(spec, n_fft=self.filter_length, hop_length=self.hop_length, win_length=self.win_length,
                      window=self.window, center=True, normalized=False, onesided=True)
    return out.unsqueeze(1)                                      # the reference returns [B, 1, n]

  def inverse(self, magnitude: torch.Tensor, phase: torch.Tensor) -> torch.Tensor:
    return self.inverse_complex(torch.complex(magnitude * torch.cos(phase), magnitude * torch.sin(phase)))   # stft.py:164-165

  def forward(self, input_data: torch.Tensor) -> torch.Tensor:
    return self.inverse(*self.transform(input_data))


class Denoiser(torch.nn.Module):
  """Same constructor and call as the reference's Denoiser (denoiser.py:17, 51)."""

  def __init__(self, waveglow, hparams, mode: str = "zeros", device: torch.device = None):
    super().__init__()
    w = waveglow.upsample.weight
    self.device = device if device is not None else w.device
    self.stft = STFT(filter_length=hparams.filter_length, hop_length=hparams.hop_length, win_length=hparams.win_length,
                     device=self.device)
    if mode == "zeros":
      mel_input = torch.zeros((1, hparams.n_mel_channels, BIAS_MEL_LENGTH), dtype=w.dtype, device=w.device)
    elif mode == "normal":
      mel_input = torch.randn((1, hparams.n_mel_channels, BIAS_MEL_LENGTH), dtype=w.dtype, device=w.device)
    else:
      raise Exception(f"Mode {mode} if not supported")
    with torch.no_grad():
      bias_audio = waveglow.infer(mel_input, sigma=0.0).float()
      bias_spec, _ = self.stft.transform(bias_audio.to(self.device))
    self.register_buffer("bias_spec", bias_spec[:, :, 0][:, :, None])

  def forward(self, audio: torch.Tensor, strength: float) -> torch.Tensor:
    """denoiser.py:51-57.  The reference rebuilds the spectrum from clamp(|X| - strength * bias, 0) and the phase of X with
    atan2 / cos / sin; that is the real scaling X * max(|X| - strength * bias, 0) / |X|, done here in place by ONE kernel
    (`wg_denoise_spectrum`) between the two cuFFT transforms: no magnitude / phase / cos / sin tensors, no transcendentals."""
    audio = audio.to(self.device)
    spec = self.stft._complex(audio)
    if spec.is_cuda:
      from .engine import denoise_spectrum_
      return self.stft.inverse_complex(denoise_spectrum_(spec, self.bias_spec.reshape(-1), strength))
    # host tensors (the CPU tests that pin this arithmetic to the reference's golden vectors): the same scaling with torch ops
    mag = spec.abs()
    scale = torch.where(mag > 0, torch.clamp(mag - self.bias_spec * strength, min=0.0) / mag.clamp_min(1e-30), torch.zeros_like(mag))
    return self.stft.inverse_complex(spec * scale)
