"""CPU: the oracle against the golden vectors produced by the unmodified reference, and (in the build
container only) against the reference itself run live."""
import os
import sys

import pytest
import torch

from oracle import engine_model as E
from oracle import waveglow_oracle as O
from tests.helpers import FP16_MAX_ABS, FP16_MIN_SNR_DB, FP32_MAX_ABS, golden_noise, load_golden, max_abs, snr_db, synthetic_weights

INFER_CASES = ["infer_c256_b2_f12", "infer_c256_b1_f40", "infer_c256_b3_f7_ragged", "infer_c256_b1_f11_sigma0",
               "infer_c512_b1_f10"]
FWD_CASES = ["forward_c256_b2_t2048", "forward_c256_b1_t4000"]

# The fixtures were made on the build container's CPU; another CPU may pick other oneDNN kernels, so
# "bit-identical" is relaxed to 20 ulp-ish of the audio scale when the machine differs.
CROSS_MACHINE_ATOL = 2e-6


@pytest.mark.parametrize("name", INFER_CASES)
def test_oracle_infer_matches_reference_golden(name):
  g = load_golden(name)
  hp, sd, w, dims = synthetic_weights(int(g["n_channels"]), str(g["weight_style"]))
  with torch.no_grad():
    out = O.infer(w, dims, torch.from_numpy(g["mel"]), float(g["sigma"]), golden_noise(g))
  ref = torch.from_numpy(g["audio"])
  assert out.shape == ref.shape
  assert max_abs(ref, out) <= CROSS_MACHINE_ATOL * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("name", FWD_CASES)
def test_oracle_forward_matches_reference_golden(name):
  g = load_golden(name)
  hp, sd, w, dims = synthetic_weights(int(g["n_channels"]))
  with torch.no_grad():
    z, log_s_list, log_det_w = O.forward(w, dims, torch.from_numpy(g["mel"]), torch.from_numpy(g["audio_in"]))
  assert max_abs(torch.from_numpy(g["z"]), z) <= 1e-5
  for k, ls in enumerate(log_s_list):
    assert max_abs(torch.from_numpy(g[f"log_s{k}"]), ls) <= 1e-5
  n = z.size(0) * z.size(2)
  for k, ld in enumerate(log_det_w):
    assert abs(float(ld) - float(g["log_det_w"][k])) <= 2e-6 * n + 1e-5 * abs(float(g["log_det_w"][k]))
  assert abs(float(O.loss(z, log_s_list, log_det_w)) - float(g["loss"])) <= 1e-5


def test_forward_then_infer_round_trip():
  """Size-independent property: infer(forward(audio)) == audio when the z of forward is fed back (sigma = 1)."""
  hp, sd, w, dims = synthetic_weights(256)
  g = load_golden("forward_c256_b2_t2048")
  mel = torch.from_numpy(g["mel"])
  audio = torch.from_numpy(g["audio_in"])
  with torch.no_grad():
    z, _, _ = O.forward(w, dims, mel, audio)
    # forward emits cat([early(k=4), early(k=8), final]); infer wants [final, early(k=8), early(k=4)]
    zs = [z[:, 4:], z[:, 2:4], z[:, :2]]
    mel_pad = torch.cat([mel, torch.zeros(mel.size(0), mel.size(1), 0)], 2)
    # forward trims the upsampled mel at the FRONT to len(audio); infer trims the tail: same first 256*F samples
    back = O.infer(w, dims, mel_pad, 1.0, zs)
  assert max_abs(audio, back[:, : audio.size(1)]) <= 2e-4


@pytest.mark.parametrize("name", ["infer_c256_b2_f12", "infer_c256_b3_f7_ragged", "infer_c512_b1_f10"])
def test_engine_model_fp32_is_exact_restructuring(name):
  """The engine's algebraic restructuring (K-extended conditioning, skip->end fold, time-major state) is exact."""
  g = load_golden(name)
  hp, sd, w, dims = synthetic_weights(int(g["n_channels"]), str(g["weight_style"]))
  with torch.no_grad():
    out = E.infer(w, dims, torch.from_numpy(g["mel"]), float(g["sigma"]), golden_noise(g), torch.float32)
  assert max_abs(torch.from_numpy(g["audio"]), out) <= FP32_MAX_ABS


@pytest.mark.parametrize("name", ["infer_c256_b2_f12", "infer_c256_b3_f7_ragged"])
def test_engine_model_composed_first_layer_is_exact_and_in_budget(name):
  """start conv composed into each flow's first layer (and the upsampler into the conditioning): exact in fp32, and
  inside the tensor-mode gate with fp16 operands."""
  g = load_golden(name)
  hp, sd, w, dims = synthetic_weights(256, str(g["weight_style"]))
  ref = torch.from_numpy(g["audio"])
  mel, sigma, zs = torch.from_numpy(g["mel"]), float(g["sigma"]), golden_noise(g)
  with torch.no_grad():
    out32 = E.infer(w, dims, mel, sigma, zs, torch.float32, phase_cond=True, first_composed=True)
    out16 = E.infer(w, dims, mel, sigma, zs, torch.float16, phase_cond=True, first_composed=True)
  assert max_abs(ref, out32) <= FP32_MAX_ABS
  assert max_abs(ref, out16) <= FP16_MAX_ABS and snr_db(ref, out16) >= FP16_MIN_SNR_DB


@pytest.mark.parametrize("name", ["infer_c256_b2_f12", "infer_c256_b1_f40"])
def test_engine_model_fp16_budget(name):
  """fp16 operands / fp32 accumulate / fp16 residual stream stay inside the BF16-TF32-mode gate; bf16 does not."""
  g = load_golden(name)
  hp, sd, w, dims = synthetic_weights(256)
  ref = torch.from_numpy(g["audio"])
  with torch.no_grad():
    out = E.infer(w, dims, torch.from_numpy(g["mel"]), float(g["sigma"]), golden_noise(g), torch.float16)
  assert max_abs(ref, out) <= FP16_MAX_ABS and snr_db(ref, out) >= FP16_MIN_SNR_DB
  if name == "infer_c256_b1_f40":
    with torch.no_grad():
      out_bf = E.infer(w, dims, torch.from_numpy(g["mel"]), float(g["sigma"]), golden_noise(g), torch.bfloat16)
    # documents WHY the engine's operand type is fp16 and not bf16
    assert max_abs(ref, out_bf) > FP16_MAX_ABS


def test_fold_weight_norm_three_key_styles_agree():
  _, _, w_a, _ = synthetic_weights(256, "parametrized")
  _, _, w_b, _ = synthetic_weights(256, "legacy")
  assert set(w_a.keys()) == set(w_b.keys()) and len(w_a) == 470
  for k in w_a:
    assert torch.equal(w_a[k], w_b[k]), k
  w_c = O.fold_weight_norm(w_a)
  for k in w_a:
    assert torch.equal(w_a[k], w_c[k]), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/waveglow"), reason="reference tree only exists in the build container")
def test_oracle_bit_identical_to_live_reference():
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
  from ref_loader import load_reference
  from waveglow_b200.synthetic import synthetic_mel
  WaveGlow, RefHParams = load_reference()
  hp, sd, w, dims = synthetic_weights(256)
  m = WaveGlow(RefHParams())
  m.load_state_dict(sd)
  m = WaveGlow.remove_weightnorm(m).eval()
  mel = synthetic_mel(2, 9, 77)
  torch.manual_seed(3)
  with torch.no_grad():
    ref = m.infer(mel, sigma=0.9)
  torch.manual_seed(3)
  zs = O.draw_noise(dims, 2, 9 * 32)
  with torch.no_grad():
    out = O.infer(w, dims, mel, 0.9, zs)
  assert torch.equal(ref, out)
