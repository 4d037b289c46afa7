"""CPU: the C ABI loads and exports what include/waveglow_b200.h declares; host-side mirror of the reference interface."""
import ctypes
import dataclasses
import os
import re

import pytest
import torch

import __graft_entry__ as graft
from tests.helpers import synthetic_weights
from waveglow_b200 import engine as wg_engine
from waveglow_b200.hparams import HParams, get_dataclass_from_dict, overwrite_custom_hparams
from waveglow_b200.model import WaveGlow
from waveglow_b200.synthetic import flow_channels, synthetic_noise, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
  graft.build()
  return wg_engine.load_library()


def test_library_exports_every_declared_symbol(lib):
  header = open(os.path.join(ROOT, "include", "waveglow_b200.h")).read()
  declared = set(re.findall(r"\b(wg_[a-z_0-9]+)\s*\(", header))
  declared -= {"wg_status", "wg_precision"}
  assert {"wg_create", "wg_load_weights", "wg_infer", "wg_forward", "wg_workspace_bytes", "wg_destroy"} <= declared
  for sym in sorted(declared):
    assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
  assert set(wg_engine.EXPORTED_SYMBOLS) <= declared
  assert b"sm_100a" in lib.wg_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(lib):
  """No CPU fallback: without a device wg_create returns WG_ERR_CUDA and says why."""
  hp = HParams()
  cfg = wg_engine._Config(hp.n_mel_channels, hp.n_flows, hp.n_group, hp.n_early_every, hp.n_early_size, hp.n_layers,
                          hp.n_channels, hp.kernel_size, hp.segment_length, wg_engine.PRECISION_FP16, 0)
  h = ctypes.c_void_p()
  rc = lib.wg_create(ctypes.byref(cfg), ctypes.byref(h))
  assert rc == -3
  assert b"no CPU fallback" in lib.wg_last_error(h)
  lib.wg_destroy(h)


def test_create_rejects_unsupported_hparams(lib):
  hp = HParams()
  cfg = wg_engine._Config(hp.n_mel_channels, hp.n_flows, 4, hp.n_early_every, hp.n_early_size, hp.n_layers,
                          hp.n_channels, hp.kernel_size, hp.segment_length, wg_engine.PRECISION_FP16, 0)
  h = ctypes.c_void_p()
  rc = lib.wg_create(ctypes.byref(cfg), ctypes.byref(h))
  assert rc == -2 and b"n_group=8" in lib.wg_last_error(h)
  lib.wg_destroy(h)


def test_model_infer_on_cpu_raises():
  model = WaveGlow(HParams())
  with pytest.raises(wg_engine.WgError, match="no CPU fallback"):
    model.infer(torch.zeros(1, 80, 4), sigma=1.0)


def test_hparams_field_names_and_overrides():
  hp = HParams()
  d = dataclasses.asdict(hp)
  for name, val in dict(n_mel_channels=80, n_flows=12, n_group=8, n_early_every=4, n_early_size=2, n_layers=8,
                        n_channels=256, kernel_size=3, segment_length=16000, sampling_rate=22050, sigma=1.0).items():
    assert d[name] == val
  hp2 = overwrite_custom_hparams(hp, {"n_channels": "512"})
  assert hp2.n_channels == 512 and isinstance(hp2.n_channels, int)
  with pytest.raises(Exception):
    overwrite_custom_hparams(hp, {"nope": "1"})
  hp3, ignored = get_dataclass_from_dict({**d, "legacy_field": 3})
  assert hp3 == hp and ignored == {"legacy_field"}


@pytest.mark.parametrize("style", ["parametrized", "legacy"])
def test_state_dict_layout_and_three_key_styles(style):
  """The module tree ingests the reference's checkpoint layouts and folds like remove_weightnorm."""
  hp, sd, w, dims = synthetic_weights(256, style)
  model = WaveGlow(hp)
  if style == "parametrized":
    assert set(model.state_dict().keys()) == set(sd.keys()) and len(sd) == 686
  res = model.load_state_dict(sd)
  assert not res.missing_keys and not res.unexpected_keys
  model = WaveGlow.remove_weightnorm(model)
  folded_sd = model.state_dict()
  assert len(folded_sd) == 470
  for k, v in model.folded_weights().items():
    assert torch.equal(v, w[k]), k
  # already-folded checkpoints load too
  model2 = WaveGlow.remove_weightnorm(WaveGlow(hp))
  res = model2.load_state_dict(folded_sd)
  assert not res.missing_keys and not res.unexpected_keys


def test_flow_channels_and_noise_shapes():
  hp = HParams()
  assert flow_channels(hp) == [8, 8, 8, 8, 6, 6, 6, 6, 4, 4, 4, 4]
  zs = synthetic_noise(hp, 3, 64, 1)
  assert [tuple(z.shape) for z in zs] == [(3, 4, 64), (3, 2, 64), (3, 2, 64)]
  model = WaveGlow.__new__(WaveGlow)
  torch.nn.Module.__init__(model)
  model.n_remaining_channels, model.n_flows, model.n_early_every, model.n_early_size = 4, 12, 4, 2
  torch.manual_seed(5)
  a = model.draw_noise(2, 32, torch.device("cpu"))
  from oracle import waveglow_oracle as O
  torch.manual_seed(5)
  b = O.draw_noise(O.ModelDims(), 2, 32)
  assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_synthetic_checkpoint_is_deterministic_and_non_degenerate():
  hp = HParams()
  a = synthetic_state_dict(hp, 7)
  b = synthetic_state_dict(hp, 7)
  assert all(torch.equal(a[k], b[k]) for k in a)
  assert float(a["WN.0.end.weight"].abs().max()) > 0   # the reference zero-inits `end`: would hide the WN kernels


def test_bench_reference_arm_prints_the_contract_line():
  """`bench.py --impl reference` (the one place besides tests / smoke that may execute oracle/) runs on the host cores and
  prints ONE JSON line with the contract's keys; no GPU, no CUDA library involved."""
  import json
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                     capture_output=True, text=True, timeout=900, cwd=root)
  assert r.returncode == 0, r.stderr[-500:]
  lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
  assert d["value"] > 0 and d["steps"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
  # the unmodified reference where /root/reference exists (build container), the oracle port elsewhere (GPU box)
  assert d["cpu_baseline"]["kind"] == ("reference" if os.path.isdir("/root/reference/src/waveglow") else "port")
  assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  # the line says what it timed: one utterance of the headline length (861 frames), not a shorter stand-in
  assert "1 utterance x 861 frames (220416 samples)" in d["cpu_baseline"]["sample"]
  assert abs(d["value"] * d["ms_per_step"] * 1e-3 - 220416) < 1.0
  assert "configs[1]" in d["config"]["workload"] and "ONE utterance" in d["config"]["reference_arm_sample"]
  assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert "workload" in d["config"] and "model" not in d["config"]


def test_roofline_traffic_entries_name_a_committed_capture():
  """bench.py copies `roofline.traffic` from profiles/traffic.json (an ncu capture cannot run inside the timed program): every
  entry must carry a positive byte count and point at a summary file that is committed, and the headline workload must have one."""
  import json
  import re
  with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
    tr = json.load(f)
  entries = {k: v for k, v in tr.items() if not k.startswith("_")}
  assert "c256_b16_f861_fp16" in entries
  for key, ent in entries.items():
    assert re.fullmatch(r"c(256|512)_b\d+_f\d+_(fp16|fp32)", key), key
    assert ent["bytes_per_launch"] > 0
    m = re.match(r"(profiles/\S+)", ent["source"])
    assert m and os.path.isfile(os.path.join(ROOT, m.group(1))), ent["source"]
