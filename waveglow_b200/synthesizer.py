"""`Synthesizer` -- the reference's per-utterance workflow on top of the B200 engine.

Mirrors /root/reference/src/waveglow/synthesizer.py:20-94: same constructor (`checkpoint`, `custom_hparams`, `device`),
same `infer(mel, sigma=, denoiser_strength=, seed=) -> InferenceResult` with the same fields.  Differences, on purpose:
  * no silent CPU fallback (`try_copy_to`, utils.py:308-316): a failing device placement raises;
  * `inference_duration_s` / `denoising_duration_s` are synchronised with the device -- the reference's perf_counter
    brackets only measure launch time on a GPU;
  * `infer_batch` (new): equal-length mels in one engine call with on-device normalisation + int16 quantisation.
"""
import datetime
import os
import random
import time
from dataclasses import dataclass
from logging import getLogger
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .audio_utils import is_overamp, normalize_to_int16_device
from .checkpoint import CheckpointWaveglow
from .denoiser import Denoiser
from .hparams import overwrite_custom_hparams
from .model import WaveGlow, load_model


def checkpoint_fingerprint(path) -> int:
  """Cheap identity of a checkpoint file for the packed-weight cache: size and modification time."""
  st = os.stat(path)
  return (st.st_size * 1000003 + st.st_mtime_ns) & (2 ** 64 - 1)


def init_global_seeds(seed: int) -> None:
  """utils.py:221-231."""
  os.environ["PYTHONHASHSEED"] = str(seed)
  random.seed(seed)
  np.random.seed(seed)
  torch.random.manual_seed(seed)
  torch.manual_seed(seed)
  if torch.cuda.is_available():
    torch.cuda.manual_seed(seed)


@dataclass
class InferenceResult():
  wav: np.ndarray
  wav_denoised: np.ndarray
  sampling_rate: int
  inference_duration_s: float
  denoising_duration_s: float
  was_overamplified: bool
  timepoint: datetime.datetime


class Synthesizer():
  def __init__(self, checkpoint: CheckpointWaveglow, *, custom_hparams: Optional[Dict[str, str]] = None,
               device: torch.device = None, precision: str = "fp16", pack_cache=None, pack_fingerprint: int = 0):
    """pack_cache: optional path of a packed-weight cache.  If it exists and was written for these weights
    (`pack_fingerprint`, e.g. `checkpoint_fingerprint(path)`), this engine configuration and this library version, the
    engine restores its packed arrays from it instead of folding and re-packing the state_dict; otherwise it packs as usual
    and writes the file for the next start."""
    if device is None:
      device = torch.device("cuda:0")
    device = torch.device(device)
    hparams = checkpoint.get_hparams()
    hparams = overwrite_custom_hparams(hparams, custom_hparams)
    model = load_model(hparams, checkpoint.state_dict, device, precision=precision)
    model = WaveGlow.remove_weightnorm(model)
    model = model.eval()
    if pack_cache is not None:
      model.pack_cache = (str(pack_cache), int(pack_fingerprint))
    denoiser = Denoiser(waveglow=model, hparams=hparams, mode="zeros", device=device).to(device)
    self.device = device
    self.hparams = hparams
    self.model = model
    self.denoiser = denoiser

  def _sync(self) -> None:
    if self.device.type == "cuda":
      torch.cuda.synchronize(self.device)

  def infer(self, mel: torch.Tensor, *, sigma: float = 1.0, denoiser_strength: float = 0.0005, seed: int = 0) -> InferenceResult:
    timepoint = datetime.datetime.now()
    init_global_seeds(seed)
    denoising_duration = 0
    mel = mel.to(self.device)
    self._sync()
    start = time.perf_counter()
    with torch.no_grad():
      audio = self.model.infer(mel, sigma=sigma)
      self._sync()
      end = time.perf_counter()
      audio_denoised = audio
      if denoiser_strength > 0:
        start_denoising = time.perf_counter()
        audio_denoised = self.denoiser(audio, strength=denoiser_strength)
        self._sync()
        denoising_duration = time.perf_counter() - start_denoising
    # ONE device -> host copy for both wavs (the reference does two, synthesizer.py:69-70)
    if audio_denoised is audio:
      audio_np: np.ndarray = audio.squeeze().cpu().numpy()
      audio_denoised_np: np.ndarray = audio_np
    else:
      both = torch.stack([audio.squeeze(), audio_denoised.squeeze()]).cpu().numpy()
      audio_np, audio_denoised_np = both[0], both[1]
    was_overamplified = False
    if is_overamp(audio_np):
      was_overamplified = True
      getLogger(__name__).debug("Waveglow output was overamplified.")
    return InferenceResult(sampling_rate=self.hparams.sampling_rate, inference_duration_s=end - start, wav=audio_np,
                           was_overamplified=was_overamplified, wav_denoised=audio_denoised_np,
                           denoising_duration_s=denoising_duration, timepoint=timepoint)

  def infer_batch(self, mels: Sequence[torch.Tensor], *, sigma: float = 1.0, denoiser_strength: float = 0.0005,
                  seed: int = 0) -> List[np.ndarray]:
    """Equal-length mels [80, F] -> list of normalised int16 wavs, bit-compatible with calling `infer` on each mel and
    then `float_to_wav(normalize_wav(result.wav_denoised))` as the reference CLI does (inference_v2.py:112-130): every
    utterance of the reference loop is synthesised from the SAME re-seeded noise, so the batch shares one noise draw."""
    F = mels[0].size(-1)
    assert all(m.dim() == 2 and m.size(-1) == F for m in mels), "infer_batch needs equal-length [n_mel, F] mels"
    B = len(mels)
    mel = torch.stack([m.to(self.device, dtype=torch.float32) for m in mels], 0)
    init_global_seeds(seed)
    with torch.no_grad():
      zs = self.model.draw_noise(1, F * 256 // self.model.n_group, self.device)      # what one reference call draws
      zs = [z.expand(B, -1, -1).contiguous() for z in zs]
      audio = self.model.infer_with_noise(mel, zs, sigma)
      if denoiser_strength > 0:
        audio = self.denoiser(audio, strength=denoiser_strength).squeeze(1)
      pcm = normalize_to_int16_device(audio).cpu().numpy()
    return [pcm[i] for i in range(B)]
