"""Import the UNMODIFIED reference `waveglow.model` from /root/reference through a stub loader.

Only usable in the build container (the GPU box has no /root/reference).  Used by
`tools/make_golden.py` to generate the committed fixtures under tests/golden/ and by the
container-only oracle-pinning test.  Recipe follows SURVEY.md §8c: the reference's `model.py`
pulls librosa/matplotlib/... through `hparams -> taco_stft` and `utils`, none of which touch the
arithmetic, so empty stub modules are registered for the missing names and the package shell is
registered with an explicit __path__ to bypass `waveglow/__init__.py`.
"""
import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"

_MISSING = [
  "librosa", "librosa.util", "librosa.filters", "librosa.core", "matplotlib", "matplotlib.pylab",
  "matplotlib.pyplot", "matplotlib.figure", "matplotlib.ticker", "fastdtw", "fastdtw.fastdtw", "imageio", "skimage",
  "skimage.metrics", "mel_cepstral_distance", "ordered_set", "wget", "gdown", "plotly",
  "plotly.offline", "plotly.graph_objs", "pandas_ods_reader",
]


class _Anything(types.ModuleType):
  """Module stub whose every attribute is a harmless callable/class placeholder."""

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    obj = type(name, (), {"__init__": lambda self, *a, **k: None})
    setattr(self, name, obj)
    return obj


def available() -> bool:
  return os.path.isdir(os.path.join(REF_SRC, "waveglow"))


def _install_librosa_util_shims():
  """librosa is not installed; the reference's STFT needs three helpers from librosa.util (stft.py:36).  For the only
  case the reference uses (win_length == filter_length, norm=None) they are trivial."""
  import numpy as np
  util = sys.modules["librosa.util"]

  def pad_center(data, size, axis=-1, **kwargs):
    n = data.shape[axis]
    assert size >= n
    lpad = (size - n) // 2
    widths = [(0, 0)] * data.ndim
    widths[axis] = (lpad, size - n - lpad)
    return np.pad(data, widths, mode="constant")

  util.pad_center = pad_center
  util.tiny = lambda x: np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny
  util.normalize = lambda S, norm=None, **kw: S


def load_reference_module(name: str):
  """Imports `waveglow.<name>` of the reference (after load_reference() has prepared the stubs)."""
  load_reference()
  return importlib.import_module("waveglow." + name)


def load_reference():
  """Returns (WaveGlow, HParams) classes of the reference, unmodified."""
  if not available():
    raise RuntimeError("reference tree not present (this only works in the build container)")
  for name in _MISSING:
    if name in sys.modules:
      continue
    try:
      importlib.import_module(name)
    except Exception:
      stub = _Anything(name)
      stub.__path__ = []  # let "from pkg.sub import x" resolve through sys.modules
      sys.modules[name] = stub
  if "librosa.util" in sys.modules and isinstance(sys.modules["librosa.util"], _Anything):
    _install_librosa_util_shims()      # must precede the first import of waveglow.stft (binds the names at import time)
  if "waveglow" not in sys.modules or not hasattr(sys.modules["waveglow"], "__path__"):
    shell = types.ModuleType("waveglow")
    shell.__path__ = [os.path.join(REF_SRC, "waveglow")]
    sys.modules["waveglow"] = shell
  model = importlib.import_module("waveglow.model")
  hparams = importlib.import_module("waveglow.hparams")
  return model.WaveGlow, hparams.HParams
