"""Hyper-parameters with the reference's field names (they are serialised into checkpoints).

Mirrors /root/reference/src/waveglow/hparams.py:6-43 and taco_stft.py:36-50: the checkpoint stores
`asdict(HParams)` (model_checkpoint.py:23) and `--custom-hparams "k=v"` overrides by field name
(utils.py:48-59), so names, defaults and the MRO-merged `HParams` are kept verbatim.
"""
import dataclasses
from typing import Dict, Optional, Set, Tuple


# One flat schema instead of the reference's four mixin dataclasses: (field, type, default, where the reference defines it).
# The ORDER is the field order of the reference's MRO-merged `HParams` (hparams.py:42-43), so `asdict` output and positional
# construction agree with checkpoints written by the reference.
_SCHEMA = (
  # OptimizerHParams, hparams.py:35-39
  ("learning_rate", float, 1e-4), ("sigma", float, 1.0), ("batch_size", int, 1),
  # ModelHParams, hparams.py:20-31 (the nine ints wg_config carries; n_layers / n_channels / kernel_size = "WN_config")
  ("segment_length", int, 16000), ("n_mel_channels", int, 80), ("n_flows", int, 12), ("n_group", int, 8),
  ("n_early_every", int, 4), ("n_early_size", int, 2), ("n_layers", int, 8), ("n_channels", int, 256), ("kernel_size", int, 3),
  # STFTHParams / TSTFTHParams, taco_stft.py:36-50
  ("filter_length", int, 1024), ("hop_length", int, 256), ("win_length", int, 1024), ("window", str, "hann"),
  ("sampling_rate", int, 22050), ("mel_fmin", float, 0.0), ("mel_fmax", float, 8000.0),
  # ExperimentHParams, hparams.py:7-16 (training bookkeeping; carried so that checkpoints round-trip)
  ("epochs", int, 100000), ("iters_per_checkpoint", int, 2000), ("epochs_per_checkpoint", int, 1), ("seed", int, 1234),
  ("cache_wavs", bool, False), ("cudnn_enabled", bool, True), ("cudnn_benchmark", bool, False),
)

HParams = dataclasses.make_dataclass(
  "HParams", [(name, typ, dataclasses.field(default=default)) for name, typ, default in _SCHEMA], module=__name__)
HParams.__doc__ = "All hyper-parameters of the reference (`HParams`, hparams.py:42-43) as one dataclass; see _SCHEMA."


def get_dataclass_from_dict(params: Dict, dc=HParams) -> Tuple[HParams, Set[str]]:
  """Unknown keys are dropped and reported (checkpoint.py:22-28, utils.py:93-97)."""
  names = {f.name for f in dataclasses.fields(dc)}
  res = {k: v for k, v in params.items() if k in names}
  ignored = {k for k in params.keys() if k not in names}
  return dc(**res), ignored


def overwrite_custom_hparams(hparams: HParams, custom: Optional[Dict[str, str]]) -> HParams:
  """`--custom-hparams` semantics (utils.py:48-59, 62-80): unknown key raises, value cast to the default's type."""
  if custom is None:
    return hparams
  available = dataclasses.asdict(hparams)
  typed = {}
  for key, val in custom.items():
    if key not in available:
      raise Exception(f"unknown hparam '{key}'")
    typed[key] = None if val == "" else type(available[key])(val)
  return dataclasses.replace(hparams, **typed)
