"""Post-processing of synthesised audio: peak normalisation and PCM conversion.

Semantics follow /root/reference/src/waveglow/audio_utils.py (`normalize_wav` 67-95, `convert_wav` 53-64,
`float_to_wav` 26-33, `is_overamp` 132-138): a float wav is scaled so that max|x| == 1.0, an int16 wav is
`round(x * 32767)`.  The numpy versions serve the one-utterance `Synthesizer` path; `normalize_to_int16_device` does both
steps on the GPU for the batched CLI so only 2 bytes per sample cross PCIe instead of 4 + a host-side pass.
"""
import numpy as np
import torch
from scipy.io.wavfile import write

FLOAT_MAX_WAV = 1.0
FLOAT_MIN_WAV = -1.0
INT16_MAX = np.iinfo(np.int16).max
INT16_MIN = np.iinfo(np.int16).min


def is_overamp(wav: np.ndarray) -> bool:
  """audio_utils.py:132-138 for float wavs: any sample outside [-1, 1]."""
  if wav.dtype in (np.float32, np.float64):
    return bool(np.min(wav) < FLOAT_MIN_WAV or np.max(wav) > FLOAT_MAX_WAV)
  info = np.iinfo(wav.dtype)
  return bool(np.min(wav) < info.min or np.max(wav) > info.max)


def normalize_wav(wav: np.ndarray) -> np.ndarray:
  """audio_utils.py:67-95 for float input: scale to max|x| == 1 (no-op for silence or already normalised audio)."""
  assert wav.dtype in (np.float32, np.float64)
  max_val = np.max(np.abs(wav))
  if max_val == 0 or max_val == FLOAT_MAX_WAV:
    return wav
  orig = wav.dtype
  out = (wav.astype(np.float32) * FLOAT_MAX_WAV / max_val).astype(orig)
  return out


def convert_wav(wav: np.ndarray, to_dtype=np.int16) -> np.ndarray:
  """audio_utils.py:53-64: float [-1, 1] -> round(x * 32767) int16."""
  if wav.dtype == to_dtype:
    return wav
  out = wav / (-1 * FLOAT_MIN_WAV) * np.iinfo(to_dtype).max
  return np.round(out, 0).astype(to_dtype)


def float_to_wav(wav: np.ndarray, path, dtype=np.int16, sample_rate: int = 22050) -> None:
  """audio_utils.py:26-33."""
  write(filename=str(path), rate=sample_rate, data=convert_wav(wav, dtype))


def normalize_to_int16_device(audio: torch.Tensor) -> torch.Tensor:
  """[B, n] float on the device -> [B, n] int16 on the device: per-utterance peak normalisation + rounding.
  Matches convert_wav(normalize_wav(x)) of the numpy path (same fp32 operations in the same order)."""
  a = audio.float()
  peak = a.abs().amax(dim=1, keepdim=True)
  scale = torch.where((peak == 0) | (peak == FLOAT_MAX_WAV), torch.ones_like(peak), FLOAT_MAX_WAV / peak)
  normed = torch.where((peak == 0) | (peak == FLOAT_MAX_WAV), a, a * FLOAT_MAX_WAV / peak)
  del scale
  return torch.round(normed * float(INT16_MAX)).to(torch.int16)
