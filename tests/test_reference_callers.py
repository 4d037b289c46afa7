"""Drop-in proof with the reference's OWN callers (the `synthesize` path "must keep working unmodified
on top of the new engine"): the unmodified `waveglow.synthesizer.Synthesizer`, `waveglow.denoiser.Denoiser`,
`waveglow.train.load_model` and `waveglow.model_checkpoint.CheckpointWaveglow` of /root/reference run with ONE change -- the
class `load_model` instantiates (`waveglow.train.WaveGlow`) is this package's `waveglow_b200.model.WaveGlow`, which is exactly the
one-line swap INTEGRATION.md describes.

Build-container only: it needs /root/reference (absent on the GPU box) and there is no GPU here, so the C-ABI engine behind
`WaveGlow.engine()` is replaced, FOR THIS TEST, by a stand-in that evaluates the same call with the CPU oracle.  What is proven
is the whole host-side contract the reference's callers rely on -- constructor signature, `load_state_dict` of a reference
checkpoint, `model.remove_weightnorm(model)` called through an instance, `.eval()`, `.upsample.weight.dtype/.device`,
`infer(mel, sigma=...)` incl. `sigma=0.0` on a zero mel (Denoiser), the noise draw order under `init_global_seeds(seed)`, the
shape / dtype / device of the result -- by requiring BIT-IDENTICAL `InferenceResult.wav` and `.wav_denoised` to the all-reference
run.  The kernels themselves are proven against the same oracle by the `-m gpu` tests.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_loader  # noqa: E402

from oracle import waveglow_oracle as O  # noqa: E402
from waveglow_b200 import model as b200_model  # noqa: E402
from waveglow_b200.hparams import HParams  # noqa: E402
from waveglow_b200.synthetic import synthetic_mel, synthetic_state_dict  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree (build container only)")


class OracleEngine:
  """Stand-in for waveglow_b200.engine.Engine where no GPU exists: same `infer(mel, zs, sigma)` contract, evaluated by the
  CPU oracle from the folded weights the model would hand to `wg_load_weights`."""

  def __init__(self, model):
    self.w = {k: v.detach().float().cpu() for k, v in model.folded_weights().items()}
    self.dims = O.ModelDims(n_channels=model.hparams.n_channels)
    self.calls = []

  def infer(self, mel, zs, sigma):
    self.calls.append((tuple(mel.shape), float(sigma)))
    with torch.no_grad():
      return O.infer(self.w, self.dims, mel.float().cpu(), sigma, [z.cpu() for z in zs])


@pytest.fixture
def reference(monkeypatch):
  ref_loader.load_reference()
  mods = {name: ref_loader.load_reference_module(name) for name in ("synthesizer", "train", "model", "model_checkpoint", "hparams", "denoiser")}

  def engine(self):
    if getattr(self, "_oracle_engine", None) is None or not self._packed:
      self._oracle_engine = OracleEngine(self)
      self._packed = True
    return self._oracle_engine

  monkeypatch.setattr(b200_model.WaveGlow, "engine", engine)
  return mods


def test_reference_synthesizer_and_denoiser_run_unmodified_on_the_new_model(reference, monkeypatch):
  ref = reference
  hp = HParams()
  sd = synthetic_state_dict(hp, 1234)
  RefHParams = ref["hparams"].HParams
  from dataclasses import asdict
  ck = ref["model_checkpoint"].CheckpointWaveglow(state_dict=sd, optimizer={}, learning_rate=1e-4, iteration=1, hparams=asdict(RefHParams()))
  cpu = torch.device("cpu")
  mel = synthetic_mel(1, 24, 77)

  # all-reference run
  assert ref["train"].WaveGlow is ref["model"].WaveGlow
  synth_ref = ref["synthesizer"].Synthesizer(ck, device=cpu)
  res_ref = synth_ref.infer(mel, sigma=0.8, denoiser_strength=0.05, seed=11)

  # the same, unmodified callers; only the class load_model builds is swapped
  monkeypatch.setattr(ref["train"], "WaveGlow", b200_model.WaveGlow)
  synth_new = ref["synthesizer"].Synthesizer(ck, custom_hparams={"n_channels": "256"}, device=cpu)
  assert isinstance(synth_new.model, b200_model.WaveGlow) and isinstance(synth_new.denoiser, ref["denoiser"].Denoiser)
  assert not synth_new.model.training
  eng = synth_new.model.engine()
  assert eng.calls == [((1, 80, 88), 0.0)]                       # Denoiser.__init__: zero mel, sigma 0 (denoiser.py:29-47)
  assert torch.equal(synth_new.denoiser.bias_spec, synth_ref.denoiser.bias_spec)
  res_new = synth_new.infer(mel, sigma=0.8, denoiser_strength=0.05, seed=11)
  assert eng.calls[-1] == ((1, 80, 24), 0.8)
  assert res_new.wav.dtype == np.float32 and res_new.wav.shape == (24 * 256,)
  assert np.array_equal(res_new.wav, res_ref.wav)                 # same seed -> same noise order -> same audio, bit for bit
  assert np.array_equal(res_new.wav_denoised, res_ref.wav_denoised)
  assert res_new.sampling_rate == res_ref.sampling_rate == 22050 and res_new.was_overamplified == res_ref.was_overamplified
  # another seed gives other audio through both
  res2 = synth_new.infer(mel, sigma=0.8, denoiser_strength=0.0, seed=12)
  assert not np.array_equal(res2.wav, res_new.wav) and np.array_equal(res2.wav, res2.wav_denoised)


def test_reference_model_and_new_model_exchange_state_dicts(reference):
  """Checkpoints written from either class load into the other (parametrized and folded layouts), key for key."""
  ref = reference
  hp = HParams()
  sd = synthetic_state_dict(hp, 99)
  m_ref = ref["model"].WaveGlow(ref["hparams"].HParams())
  m_new = b200_model.WaveGlow(hp)
  assert list(m_ref.state_dict().keys()) == list(m_new.state_dict().keys())
  assert not m_ref.load_state_dict(sd).missing_keys and not m_new.load_state_dict(sd).missing_keys
  res = m_new.load_state_dict(m_ref.state_dict())
  assert not res.missing_keys and not res.unexpected_keys
  f_ref = ref["model"].WaveGlow.remove_weightnorm(m_ref).state_dict()
  f_new = b200_model.WaveGlow.remove_weightnorm(m_new).state_dict()
  assert list(f_ref.keys()) == list(f_new.keys()) and all(torch.equal(f_ref[k], f_new[k]) for k in f_ref)
