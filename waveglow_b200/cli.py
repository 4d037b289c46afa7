"""`python -m waveglow_b200.cli synthesize CHECKPOINT FOLDER` -- the reference's `waveglow-cli synthesize` on the B200 engine.

Mirrors /root/reference/src/waveglow_cli/inference_v2.py:57-130 (arguments, defaults, seed handling, output naming
`<mel stem>.wav` mirrored under the output directory, `--overwrite`) with one addition the reference only sketched
(`--batch-size`, inference_v2.py:64): mels of equal length are synthesised in one engine call and peak-normalised +
quantised to int16 on the GPU.  Results are identical to the one-by-one loop because the reference
re-seeds the RNG for every utterance (synthesizer.py:56), i.e. equal-length utterances share their noise anyway.
`--include-stats` (MCD / SSIM plots) is reporting-only and not provided.

`python -m waveglow_b200.cli synthesize-wav CHECKPOINT FOLDER` is the reference's `waveglow-cli synthesize-wav`
(waveglow_cli/inference_wav.py:57-130: same arguments, `.wav` inputs): each wav goes through the on-device mel front-end
(`taco_stft.TacotronSTFT`) and then through exactly the same batched path as `synthesize`.
"""
import argparse
import random
import sys
import time
from collections import defaultdict
from logging import getLogger
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch
from scipy.io.wavfile import write as write_wav

from .checkpoint import CheckpointWaveglow
from .synthesizer import Synthesizer, checkpoint_fingerprint

DEFAULT_SIGMA = 1.0                   # waveglow_cli/defaults.py:9
DEFAULT_DENOISER_STRENGTH = 0.0005    # waveglow_cli/defaults.py:10


def _unit_interval(value: str) -> float:
  v = float(value)
  if not 0 <= v <= 1:
    raise argparse.ArgumentTypeError("Value needs to be in interval [0, 1]!")      # argparse_helper.py:125-129
  return v


def split_hparams_string(hparams: Optional[str]) -> Optional[Dict[str, str]]:
  """utils.py:32-38."""
  if hparams is None:
    return None
  return dict([x.split("=") for x in hparams.split(",")])


def add_common(p: argparse.ArgumentParser) -> None:
  """Options shared by `synthesize` (inference_v2.py:57-72) and `synthesize-wav` (inference_wav.py:57-72)."""
  p.add_argument("--sigma", type=_unit_interval, default=DEFAULT_SIGMA, help="sigma used for synthesis")
  p.add_argument("--denoiser-strength", type=_unit_interval, default=DEFAULT_DENOISER_STRENGTH,
                 help="strength of denoising to remove model bias after synthesis")
  p.add_argument("--device", type=str, default="cuda:0", help='CUDA device, e.g. "cuda:0" (there is no CPU fallback)')
  p.add_argument("--custom-hparams", type=str, default=None, metavar="CUSTOM-HYPERPARAMETERS",
                 help="custom hyperparameters comma separated")
  p.add_argument("--custom-seed", type=int, default=None, help="custom seed; if unset a random seed is chosen")
  p.add_argument("--batch-size", type=int, default=64, help="max. utterances of equal length per engine call")
  p.add_argument("--precision", choices=["fp16", "fp32"], default="fp16")
  p.add_argument("--pack-cache", action="store_true",
                 help="keep the engine's packed weights in CHECKPOINT.b200pack and reuse them on the next start")
  p.add_argument("-out", "--output-directory", type=Path, default=None, help="custom output directory if FOLDER should not be used")
  p.add_argument("-o", "--overwrite", action="store_true", help="overwrite already synthesized lines")


def build_parser() -> argparse.ArgumentParser:
  parser = argparse.ArgumentParser(prog="waveglow_b200", description="B200-native WaveGlow synthesis")
  sub = parser.add_subparsers(dest="command", required=True)
  p = sub.add_parser("synthesize", description="Synthesize mel-spectrograms to audio files (.wav).")
  p.add_argument("checkpoint", type=Path, metavar="CHECKPOINT", help="path to checkpoint which should be used for synthesis")
  p.add_argument("folder", type=Path, metavar="FOLDER", help="path to folder which contains the mel-spectrograms (.npy)")
  p.set_defaults(handler=synthesize, source="mel")
  add_common(p)
  p = sub.add_parser("synthesize-wav", description="Synthesize audio files to audio files (.wav).")
  p.add_argument("checkpoint", type=Path, metavar="CHECKPOINT", help="path to checkpoint which should be used for synthesis")
  p.add_argument("folder", type=Path, metavar="FOLDER", help="path to folder which contains the audio files (.wav) that should be synthesized")
  p.set_defaults(handler=synthesize, source="wav")
  add_common(p)
  return parser


def synthesize(ns: argparse.Namespace) -> bool:
  logger = getLogger(__name__)
  if not ns.checkpoint.is_file():
    logger.error("Checkpoint does not exist!")
    return False
  if not ns.folder.is_dir():
    logger.error("Folder does not exist!")
    return False
  out_dir = ns.output_directory if ns.output_directory is not None else ns.folder
  if out_dir.is_file():
    logger.error("Output directory is a file!")
    return False
  seed = ns.custom_seed if ns.custom_seed is not None else random.randint(1, 9999)
  try:
    checkpoint = CheckpointWaveglow.load(ns.checkpoint, torch.device("cpu"))
  except Exception:
    logger.error("Checkpoint couldn't be loaded!")
    return False
  suffix = ".wav" if ns.source == "wav" else ".npy"
  mel_files = sorted(f for f in ns.folder.rglob("*") if f.is_file() and f.suffix.lower() == suffix)
  todo = []
  for f in mel_files:
    wav_path = out_dir / f.relative_to(ns.folder).parent / f"{f.stem}.wav"
    if wav_path.exists() and not ns.overwrite:
      continue
    todo.append((f, wav_path))
  cache = Path(str(ns.checkpoint) + ".b200pack") if ns.pack_cache else None
  synth = Synthesizer(checkpoint, custom_hparams=split_hparams_string(ns.custom_hparams), device=torch.device(ns.device),
                      precision=ns.precision, pack_cache=cache,
                      pack_fingerprint=checkpoint_fingerprint(ns.checkpoint) if ns.pack_cache else 0)
  by_len: Dict[int, List] = defaultdict(list)
  taco_stft = None
  if ns.source == "wav":
    from .taco_stft import TacotronSTFT
    taco_stft = TacotronSTFT(synth.hparams, torch.device(ns.device))      # inference_wav.py:109
  for f, wav_path in todo:
    if taco_stft is not None:
      mel_t = taco_stft.get_mel_tensor_from_file(f)                        # stays on the device (inference_wav.py:114)
      by_len[mel_t.shape[1]].append((mel_t, wav_path))
      continue
    mel = np.load(f)
    if mel.ndim != 2 or mel.shape[0] != synth.hparams.n_mel_channels:
      logger.error(f"{f}: expected a [{synth.hparams.n_mel_channels}, frames] array, got {mel.shape}")
      return False
    by_len[mel.shape[1]].append((torch.from_numpy(mel.astype(np.float32)), wav_path))
  t0 = time.perf_counter()
  n_samples = 0
  for frames, items in sorted(by_len.items()):
    for i in range(0, len(items), ns.batch_size):
      chunk = items[i:i + ns.batch_size]
      pcms = synth.infer_batch([m for m, _ in chunk], sigma=ns.sigma, denoiser_strength=ns.denoiser_strength, seed=seed)
      for pcm, (_, wav_path) in zip(pcms, chunk):
        wav_path.parent.mkdir(parents=True, exist_ok=True)
        write_wav(str(wav_path), synth.hparams.sampling_rate, pcm)
        n_samples += pcm.shape[0]
  dt = time.perf_counter() - t0
  if n_samples:
    logger.info(f"Synthesized {len(todo)} file(s), {n_samples / synth.hparams.sampling_rate:.1f} s of audio in {dt:.2f} s "
                f"({n_samples / synth.hparams.sampling_rate / dt:.0f}x real-time incl. file I/O), seed {seed}.")
  return True


def main(argv=None) -> int:
  import logging
  logging.basicConfig(level=logging.INFO, format="[%(asctime)s] (%(levelname)s) %(message)s")
  ns = build_parser().parse_args(argv)
  ok = ns.handler(ns)
  return 0 if ok else 1


if __name__ == "__main__":
  sys.exit(main())
