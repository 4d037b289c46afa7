"""waveglow_b200 -- B200-native engine behind the reference's `WaveGlow.infer` call surface.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + C ABI), `engine.py` (ctypes binding),
`model.py` (the reference's module tree as a parameter container whose `infer` calls the engine),
`hparams.py` (field names serialised into checkpoints) and `synthetic.py` (deterministic benchmark inputs).
"""
from .hparams import HParams  # noqa: F401
from .model import WaveGlow, load_model  # noqa: F401

__all__ = ["HParams", "WaveGlow", "load_model"]
