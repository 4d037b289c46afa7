"""Checkpoint container with the reference's on-disk format.

A checkpoint file is `torch.save` of a plain dict with the keys `state_dict, optimizer, learning_rate, iteration, hparams`
(/root/reference/src/waveglow/checkpoint.py:13-35); `hparams` is `asdict(HParams)` (model_checkpoint.py:23) and unknown
hparams are dropped with a warning on load (checkpoint.py:22-28).  Files written by the reference load here and vice versa.
"""
from dataclasses import asdict, dataclass
from logging import getLogger
from pathlib import Path
from typing import Optional

import torch

from .hparams import HParams, get_dataclass_from_dict


@dataclass
class CheckpointWaveglow():
  # renaming any of these fields breaks existing checkpoint files (checkpoint.py:14)
  state_dict: dict
  optimizer: dict
  learning_rate: float
  iteration: int
  hparams: dict

  def get_hparams(self) -> HParams:
    res, ignored = get_dataclass_from_dict(self.hparams, HParams)
    if len(ignored) > 0:
      getLogger(__name__).warning(
        f"Ignored these hparams from checkpoint because they did not exist in the current HParams: {ignored}.")
    return res

  def save(self, checkpoint_path: Path) -> None:
    torch.save(asdict(self), checkpoint_path)

  @classmethod
  def load(cls, checkpoint_path: Path, device: torch.device) -> "CheckpointWaveglow":
    checkpoint_path = Path(checkpoint_path)
    assert checkpoint_path.is_file()
    return cls(**torch.load(checkpoint_path, map_location=device, weights_only=False))

  @classmethod
  def from_state_dict(cls, state_dict: dict, hparams: HParams, iteration: int = 0,
                      optimizer: Optional[dict] = None) -> "CheckpointWaveglow":
    return cls(state_dict=dict(state_dict), optimizer=optimizer or {}, learning_rate=hparams.learning_rate,
               iteration=iteration, hparams=asdict(hparams))
