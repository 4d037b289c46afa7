"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI.

Stage-by-stage checks compare the engine's intermediates with oracle/engine_model.py (same restructured
algorithm, same rounding points); end-to-end checks compare with the golden vectors produced by the
unmodified reference and, at larger sizes, use size-independent properties."""
import pytest
import torch

from oracle import engine_model as E
from oracle import waveglow_oracle as O
from tests.helpers import (FP16_MAX_ABS, FP16_MIN_SNR_DB, FP32_MAX_ABS, golden_noise, load_golden, max_abs, snr_db,
                           synthetic_weights)
from waveglow_b200.model import WaveGlow
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise

pytestmark = pytest.mark.gpu

_MODELS = {}


def get_model(n_channels: int, precision: str, style: str = "parametrized") -> WaveGlow:
  """precision: "fp32", "fp16" (the tensor-core path: fused layer on CTA pairs, continuous-issue schedule; 256 and 512
  channels run the same templated kernel), or the fp16 engine with one of its switches set: "fp16-seg32" / "fp16-seg64"
  (WG_SEG) force the segment size of the tiles; "fp16-nopack" (WG_PACK4=0) switches the packed ragged tails off and
  "fp16-nofirst" (WG_FIRST=0) the composition of the start conv into each flow's first layer, "fp16-nochain" (WG_CHAIN=0) the
  chaining of a flow's generic layers into one launch with tile-level dependencies (one launch per layer instead),
  "fp16-nocarry" (WG_EXACT_CARRY=0, 512 channels) the exact fp16 hi + lo residual carry (x' rounded to fp16 per layer)."""
  import os
  key = (n_channels, precision, style)
  if key not in _MODELS:
    hp, sd, w, dims = synthetic_weights(n_channels, style)
    m = WaveGlow(hp, precision="fp16" if precision.startswith("fp16") else precision).to("cuda:0")
    m.load_state_dict(sd)
    m = WaveGlow.remove_weightnorm(m).eval()
    if precision.startswith("fp16-seg"):
      os.environ["WG_SEG"] = precision[len("fp16-seg"):]
    if precision == "fp16-nopack":
      os.environ["WG_PACK4"] = "0"
    if precision == "fp16-nofirst":
      os.environ["WG_FIRST"] = "0"
    if precision == "fp16-nochain":
      os.environ["WG_CHAIN"] = "0"
    if precision == "fp16-nocarry":
      os.environ["WG_EXACT_CARRY"] = "0"
    try:
      m.engine()          # the engine reads its switches when it is created
    finally:
      os.environ.pop("WG_SEG", None)
      os.environ.pop("WG_PACK4", None)
      os.environ.pop("WG_FIRST", None)
      os.environ.pop("WG_CHAIN", None)
      os.environ.pop("WG_EXACT_CARRY", None)
    _MODELS[key] = m
  return _MODELS[key]


def run(model, mel, zs, sigma):
  out = model.infer_with_noise(mel.cuda(), [z.cuda() for z in zs], sigma)
  torch.cuda.synchronize()
  return out.cpu()


def test_tcgen05_descriptor_gemm():
  """TMA(SWIZZLE_128B) -> tcgen05.mma(kind::f16) -> tcgen05.ld in isolation: D = A . B^T."""
  eng = get_model(256, "fp16").engine()
  g = torch.Generator().manual_seed(0)
  for dt in (torch.float16, torch.bfloat16):
    a = (torch.randn((256, 320), generator=g)).to(dt)
    b = (torch.randn((512, 320), generator=g)).to(dt)
    d = eng.debug_gemm(a.cuda(), b.cuda()).cpu()
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    err = float((d.double() - ref).abs().max())
    print(f"debug gemm {dt}: max err {err:.3e} (ref absmax {float(ref.abs().max()):.2f})")
    assert err <= 2e-3, f"tcgen05 GEMM mismatch for {dt}: {err}"


def test_tcgen05_cta_pair_gemm():
  """cta_group::2 mechanisms in isolation: cluster launch, paired TMEM allocation, TMA loads completing on the leader
  CTA's barrier, multicast tcgen05.commit, M = 256 across the pair with each CTA staging half of B."""
  eng = get_model(256, "fp16").engine()
  g = torch.Generator().manual_seed(1)
  a = torch.randn((512, 384), generator=g).to(torch.float16)
  b = torch.randn((512, 384), generator=g).to(torch.float16)
  ref = a.double() @ b.double().t()
  for ks in (64, 32):      # operand stages of 64 channels (SWIZZLE_128B) and of 32 channels (SWIZZLE_64B)
    d = eng.debug_gemm2(a.cuda(), b.cuda(), ks=ks).cpu()
    torch.cuda.synchronize()
    err = float((d.double() - ref).abs().max())
    print(f"debug gemm2 (cta_group::2, {ks}-channel stages): max err {err:.3e} (ref absmax {float(ref.abs().max()):.2f})")
    assert err <= 2e-3


@pytest.mark.parametrize("precision,nc", [("fp32", 256), ("fp16", 256), ("fp16-nofirst", 256), ("fp16", 512), ("fp16-nofirst", 512),
                                          ("fp16-nocarry", 512)])
def test_stagewise_first_flow(precision, nc):
  """cond (upsample+squeeze), start conv, then every WN layer of the first flow, then the coupling step."""
  op_dtype = torch.float32 if precision == "fp32" else torch.float16
  tol = 2e-5 if precision == "fp32" else 6e-3      # intermediates are O(1..10); fp16 has 2^-11 relative steps
  hp, sd, w, dims = synthetic_weights(nc)
  model = get_model(nc, precision)
  eng = model.engine()
  B, F = 2, 7            # T' = 224: one full and one ragged 128-tile, dilation 128 reaches across both
  T = F * 32
  mel = synthetic_mel(B, F, 31)
  zs = synthetic_noise(hp, B, T, 32)
  sigma = 0.8
  trace = []
  phase_cond = precision != "fp32"     # tensor-core mode composes the upsampler into the conditioning weights
  first_composed = precision in ("fp16", "fp16-nocarry")  # ... and the start conv into each flow's first layer
  # the 512-channel kernel carries the residual stream as fp16 hi + fp16 lo (unless switched off)
  exact_carry = phase_cond and nc == 512 and precision != "fp16-nocarry"
  with torch.no_grad():
    ref_audio = E.infer(w, dims, mel, sigma, zs, op_dtype, trace, phase_cond=phase_cond, first_composed=first_composed,
                        exact_carry=exact_carry)
  k = dims.n_flows - 1
  flow_tr = trace[0]
  assert flow_tr["flow"] == k
  S_ref = E.rnd(E.upsample_squeeze_tm(w, dims, mel), op_dtype)
  pf = E.PackedFlow(w, dims, k)
  state0 = E.pack_noise(dims, zs, sigma)
  x0_ref = state0[..., 4:6] @ pf.w_start.t() + pf.b_start
  if not exact_carry:
    x0_ref = E.rnd(x0_ref, op_dtype)
  failures = []

  def x_view(buf, with_lo=True):
    """engine x buffer -> time-major [B, T, nc] (tensor-core mode keeps it phase-major [B][32][T/32][nc]; with the exact carry
    the stream is the sum of the hi buffer and its rounding-residual buffer, except for the input of the last layer, which has
    no residual output: nothing would read its lo part, so it is not written)"""
    raw = eng.debug_view(buf, B, F).float().cpu()
    if exact_carry and with_lo:
      raw = raw + eng.debug_view(buf.replace("x", "xlo"), B, F).float().cpu()
    if precision != "fp32":
      return raw.view(B, 32, T // 32, nc).permute(0, 2, 1, 3).reshape(B, T, nc)
    return raw.view(B, T, nc)

  def check(name, got, ref, t=tol):
    err = float((got.double() - ref.double()).abs().max())
    scale = float(ref.abs().max())
    print(f"[{precision} {nc}] {name}: max err {err:.3e} (ref absmax {scale:.3f})")
    if not err <= t * max(1.0, scale):
      failures.append(f"{name}: {err:.3e}")

  try:
    eng.debug_stop_after(k, -1)
    run(model, mel, zs, sigma)
    if phase_cond:
      check("mel windows", eng.debug_view("cond", B, F).float().view(B, F, 320).cpu(), E.rnd(E.mel_windows(mel, F), op_dtype))
    else:
      check("cond", eng.debug_view("cond", B, F).float().view(B, T, 640).cpu(), S_ref)
    if first_composed:
      a0 = eng.debug_view("a0", B, F).float().cpu().view(B, 32, T // 32, 16).permute(0, 2, 1, 3).reshape(B, T, 16)
      a0_ref = torch.zeros(B, T, 16)
      hi = E.rnd(state0[..., 4:6], op_dtype)
      a0_ref[..., 0:2] = hi
      a0_ref[..., 4:6] = E.rnd(state0[..., 4:6] - hi, op_dtype)
      a0_ref[..., 8:10] = hi
      a0_ref[..., 12:14] = 1.0
      check("[a_hi | a_lo | a_hi | 1 | 1] rows after start", a0, a0_ref)
    else:
      check("x after start", x_view("x0"), x0_ref)
    for i in range(dims.n_layers):
      eng.debug_stop_after(k, i)
      run(model, mel, zs, sigma)
      lt = flow_tr["layers"][i]
      check(f"e after layer {i}", eng.debug_view("e", B, F).view(B, T, 8).cpu()[..., : 2 * pf.h], lt["e"])
      if i < dims.n_layers - 1:
        check(f"x after layer {i}", x_view(f"x{i + 1}", with_lo=i + 1 < dims.n_layers - 1), lt["x"])      # the buffer layer i + 1 reads
    eng.debug_stop_after(k, dims.n_layers)
    state = run(model, mel, zs, sigma).view(B, T, 8)
    check("state after first flow", state, flow_tr["state"], t=tol)
  finally:
    eng.debug_stop_after(-1, 0)
  out = run(model, mel, zs, sigma)
  check("final audio vs engine model", out, ref_audio, t=(2e-5 if precision == "fp32" else 2e-3))
  assert not failures, "; ".join(failures)


@pytest.mark.parametrize("name", ["infer_c256_b2_f12", "infer_c256_b1_f40", "infer_c256_b3_f7_ragged",
                                  "infer_c256_b1_f11_sigma0", "infer_c512_b1_f10"])
def test_fp32_mode_matches_reference_golden(name):
  g = load_golden(name)
  model = get_model(int(g["n_channels"]), "fp32", str(g["weight_style"]))
  out = run(model, torch.from_numpy(g["mel"]), golden_noise(g), float(g["sigma"]))
  ref = torch.from_numpy(g["audio"])
  err = max_abs(ref, out)
  print(f"fp32 {name}: max-abs {err:.3e}")
  assert out.shape == ref.shape and err <= FP32_MAX_ABS


@pytest.mark.parametrize("name", ["infer_c256_b2_f12", "infer_c256_b1_f40", "infer_c256_b3_f7_ragged",
                                  "infer_c256_b1_f11_sigma0"])
def test_fp16_tensor_core_mode_matches_reference_golden(name):
  g = load_golden(name)
  model = get_model(256, "fp16", str(g["weight_style"]))
  out = run(model, torch.from_numpy(g["mel"]), golden_noise(g), float(g["sigma"]))
  ref = torch.from_numpy(g["audio"])
  err, snr = max_abs(ref, out), snr_db(ref, out)
  print(f"fp16 {name}: max-abs {err:.3e}  snr {snr:.1f} dB")
  assert out.shape == ref.shape
  assert err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB


@pytest.mark.parametrize("mode", ["fp16", "fp16-seg32", "fp16-nofirst"])
def test_512_channel_tensor_core_path_matches_reference_golden(mode):
  """The 512-channel instantiation of the fused layer kernel (BASELINE config 3 architecture): 4 accumulator chunks, two x'
  halves, 128 KB of gated activations in shared memory."""
  g = load_golden("infer_c512_b1_f10")
  model = get_model(512, mode, str(g["weight_style"]))
  out = run(model, torch.from_numpy(g["mel"]), golden_noise(g), float(g["sigma"]))
  ref = torch.from_numpy(g["audio"])
  err, snr = max_abs(ref, out), snr_db(ref, out)
  print(f"512 channels {mode}: max-abs {err:.3e}  snr {snr:.1f} dB")
  assert out.shape == ref.shape and err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB


@pytest.mark.parametrize("nc,B,F", [(256, 1, 200), (256, 6, 140), (256, 3, 861), (256, 9, 40), (256, 1, 5), (512, 1, 200), (512, 5, 300)])
def test_chained_layers_equal_one_launch_per_layer(nc, B, F):
  """The generic layers of a flow run as ONE launch whose tiles wait only for the tiles of the previous layer that produced
  their tap windows (completion counters in global memory, polled by the TMA warp).  Per tile the arithmetic is the one of
  the per-layer launches (WG_CHAIN=0), so the audio must be BIT-IDENTICAL: any dependency missed (a tile reading x or e rows
  before they were written) shows up here.  Shapes: config 1 (one short utterance: 64 pair-tiles on 74 CTA pairs, consecutive
  items of a pair depend on each other), packed tails with missing utterances, the headline length, no whole block at all, a
  single 5-frame tile, and 512 channels (two x' halves).  Run three times: the schedule is timing dependent."""
  hp, sd, w, dims = synthetic_weights(nc)
  mel = synthetic_mel(B, F, 800 + F)
  zs = synthetic_noise(hp, B, F * 32, 801 + F)
  ref = run(get_model(nc, "fp16-nochain"), mel, zs, 0.666)
  assert torch.isfinite(ref).all()
  for rep in range(3):
    out = run(get_model(nc, "fp16"), mel, zs, 0.666)
    assert torch.equal(out, ref), f"repetition {rep}: max-abs {max_abs(ref, out):.3e}"


@pytest.mark.parametrize("n_held", [1, 60, 131])
def test_chained_launch_completes_when_part_of_the_gpu_is_taken(n_held):
  """The chained launch spins on flags raised by other CTAs of the same grid.  It must not assume that the whole grid is
  resident: here another stream holds `n_held` SMs (a CTA with 200 KB of shared memory each) for ~60 ms while `infer` runs, so
  only some CTA pairs can be scheduled -- with 131 held at most 8 pairs, and with 1 held a pair is split.  Work items are
  claimed in order by whichever pairs run, so the call completes (late) with bit-identical audio.  A static assignment of
  items to pairs would wait forever for tiles owned by pairs that cannot start (the watchdog would trap after seconds)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  eng = model.engine()
  mel = synthetic_mel(2, 300, 901)
  zs = synthetic_noise(hp, 2, 300 * 32, 902)
  ref = run(model, mel, zs, 0.666)
  side = torch.cuda.Stream()
  mel_d, zs_d = mel.cuda(), [z.cuda() for z in zs]
  torch.cuda.synchronize()
  eng.debug_occupy(n_held, 120_000_000, side)        # ~60 ms at 2 GHz
  out = model.infer_with_noise(mel_d, zs_d, 0.666)   # enqueued while the other stream holds its SMs
  torch.cuda.synchronize()
  assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("mode", ["fp16-seg32", "fp16-seg64"])
@pytest.mark.parametrize("name", ["infer_c256_b2_f12", "infer_c256_b1_f40", "infer_c256_b3_f7_ragged"])
def test_segment_tiles_match_reference_golden(name, mode):
  """Tiles built from 32- or 64-frame segments (several TMA boxes per operand load, tiles spanning utterances), forced with
  WG_SEG on shapes that would otherwise use one 128-frame box: same golden gates, and bit-identical to the default choice
  (the segment size only changes which rows share a tile)."""
  g = load_golden(name)
  model = get_model(256, mode, str(g["weight_style"]))
  mel, zs, sigma = torch.from_numpy(g["mel"]), golden_noise(g), float(g["sigma"])
  out = run(model, mel, zs, sigma)
  ref = torch.from_numpy(g["audio"])
  err, snr = max_abs(ref, out), snr_db(ref, out)
  print(f"{mode} {name}: max-abs {err:.3e}  snr {snr:.1f} dB")
  assert out.shape == ref.shape and err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB
  assert torch.equal(out, run(get_model(256, "fp16", str(g["weight_style"])), mel, zs, sigma))


@pytest.mark.parametrize("B,F", [(6, 140), (4, 861), (9, 40)])
def test_packed_ragged_tails_vs_oracle_and_unpacked(B, F):
  """The frames past the last whole 128-frame block are tiled as 32 frames x 4 consecutive utterances (one 4-D TMA box):
  B = 6 leaves a group with two missing utterances, 861 frames is the headline length (6 blocks + 93 frames -> 3 packed
  tiles per 4 utterances), 40 frames has no whole block at all.  Same gates against the oracle, and bit-identical to the
  engine with the packing switched off (only the assignment of rows to tiles changes)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  mel = synthetic_mel(B, F, 511 + F)
  zs = synthetic_noise(hp, B, F * 32, 512 + F)
  out = run(model, mel, zs, 0.666)
  assert torch.equal(out, run(get_model(256, "fp16-nopack"), mel, zs, 0.666))
  if F <= 200:
    with torch.no_grad():
      ref = O.infer(w, dims, mel, 0.666, zs)
    err, snr = max_abs(ref, out), snr_db(ref, out)
    print(f"fp16 B{B} F{F} (packed tails): max-abs {err:.3e}  snr {snr:.1f} dB")
    assert err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB
  alone = run(model, mel[B - 1:B], [z[B - 1:B] for z in zs], 0.666)
  assert torch.equal(out[B - 1:B], alone)


def test_packed_ragged_tails_512_channels():
  """Packed tails on the 512-channel instantiation: bit-identical to the same kernel without the packing."""
  hp, sd, w, dims = synthetic_weights(512)
  B, F = 6, 140
  mel = synthetic_mel(B, F, 611)
  zs = synthetic_noise(hp, B, F * 32, 612)
  out = run(get_model(512, "fp16"), mel, zs, 0.666)
  assert torch.isfinite(out).all()
  assert torch.equal(out, run(get_model(512, "fp16-nopack"), mel, zs, 0.666))


def test_first_layer_composition_vs_explicit_start_conv():
  """WG_FIRST=0 keeps the explicit start conv (the flow step writes the 256-channel x): both forms stay inside the gates
  against the reference golden and differ from each other only by operand rounding."""
  g = load_golden("infer_c256_b1_f40")
  mel, zs, sigma = torch.from_numpy(g["mel"]), golden_noise(g), float(g["sigma"])
  ref = torch.from_numpy(g["audio"])
  a = run(get_model(256, "fp16"), mel, zs, sigma)
  b = run(get_model(256, "fp16-nofirst"), mel, zs, sigma)
  for name, out in (("composed", a), ("explicit", b)):
    err, snr = max_abs(ref, out), snr_db(ref, out)
    print(f"first layer {name}: max-abs {err:.3e}  snr {snr:.1f} dB")
    assert err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB
  assert max_abs(a, b) <= 2e-3      # each is within 2e-3 of the reference; measured 1.3e-3 between them


def test_segment_choice_short_utterances_vs_oracle():
  """B=8 x 72 frames, no whole 128-frame block: the engine packs 32 frames x 4 utterances per tile (6 tile entries instead
  of 8 ragged ones; before the packed tails existed it picked 32-frame segments here, which WG_SEG still forces)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  B, F = 8, 72
  mel = synthetic_mel(B, F, 411)
  zs = synthetic_noise(hp, B, F * 32, 412)
  with torch.no_grad():
    ref = O.infer(w, dims, mel, 0.666, zs)
  out = run(model, mel, zs, 0.666)
  err, snr = max_abs(ref, out), snr_db(ref, out)
  print(f"fp16 B8 F72 (32-frame segments): max-abs {err:.3e}  snr {snr:.1f} dB")
  assert err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB
  alone = run(model, mel[5:6], [z[5:6] for z in zs], 0.666)
  assert torch.equal(out[5:6], alone)


@pytest.mark.parametrize("B,F", [(1, 1), (5, 3), (2, 129), (1, 257)])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_edge_shapes_vs_oracle(B, F, precision):
  """Single frame (shorter than the upsampler's 4-frame support), odd batch, and lengths that cross the 128-frame tile
  boundary by one frame (the second / third tile of a phase holds a single valid row)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, precision)
  mel = synthetic_mel(B, F, 300 + F)
  zs = synthetic_noise(hp, B, F * 32, 301 + F)
  with torch.no_grad():
    ref = O.infer(w, dims, mel, 0.75, zs)
  out = run(model, mel, zs, 0.75)
  err = max_abs(ref, out)
  print(f"edge {precision} B={B} F={F}: max-abs {err:.3e}")
  assert out.shape == (B, F * 256) and torch.isfinite(out).all()
  assert err <= (FP32_MAX_ABS if precision == "fp32" else FP16_MAX_ABS)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("nc,n_flows,n_early_every,n_early_size,n_layers,B,F", [
  (256, 6, 2, 2, 4, 2, 45),      # early outputs at k = 2, 4; 4 layers per WN (chained launch of 3)
  (256, 4, 4, 2, 10, 1, 70),     # no early output at all (8 channels drawn at once); 10 layers: dilation 512, one launch per layer
  (512, 3, 2, 4, 3, 2, 33),      # one early output of 4 channels: flows run on 8, 8, 4 channels
  (256, 2, 4, 2, 1, 3, 20),      # a single layer per WN: no residual output anywhere, explicit start conv
  (256, 5, 1, 0, 2, 2, 50),      # n_early_size 0: the "early output" every flow is empty
])
def test_non_default_hparams_vs_oracle(nc, n_flows, n_early_every, n_early_size, n_layers, B, F, precision):
  """Architectures other than the published one (the reference builds them from the same nine hparams, model.py:142-176): the
  engine derives its launch plan from the configuration, nothing is specialised to 12 flows x 8 layers."""
  from waveglow_b200.hparams import HParams
  from waveglow_b200.synthetic import synthetic_state_dict
  fields = dict(n_channels=nc, n_flows=n_flows, n_early_every=n_early_every, n_early_size=n_early_size, n_layers=n_layers)
  hp = HParams(**fields)
  sd = synthetic_state_dict(hp, 77)
  w, dims = O.fold_weight_norm(sd), O.ModelDims(**fields)
  model = WaveGlow(hp, precision=precision).to("cuda:0")
  model.load_state_dict(sd)
  model = WaveGlow.remove_weightnorm(model).eval()
  mel = synthetic_mel(B, F, 41)
  zs = synthetic_noise(hp, B, F * 32, 42)
  assert [z.shape[1] for z in zs] == [dims.n_remaining_channels] + [n_early_size] * (len(zs) - 1)
  with torch.no_grad():
    ref = O.infer(w, dims, mel, 0.7, zs)
  out = run(model, mel, zs, 0.7)
  err = max_abs(ref, out)
  print(f"hparams {fields} {precision}: max-abs {err:.3e} (ref absmax {float(ref.abs().max()):.2f})")
  assert out.shape == ref.shape == (B, F * 256) and torch.isfinite(out).all()
  assert err <= (FP32_MAX_ABS if precision == "fp32" else FP16_MAX_ABS)
  # and the training direction of the same architecture
  audio = torch.rand((B, F * 256), generator=torch.Generator().manual_seed(43)) * 2 - 1
  with torch.no_grad():
    z_ref, ls_ref, ld_ref = O.forward(w, dims, mel, audio)
  z, log_s_list, log_det_list = model((mel.cuda(), audio.cuda()))
  torch.cuda.synchronize()
  tol = FP32_MAX_ABS if precision == "fp32" else FP16_MAX_ABS
  assert z.shape == z_ref.shape and max_abs(z_ref, z.cpu()) <= tol * max(1.0, float(z_ref.abs().max()))
  assert len(log_s_list) == n_flows and all(max_abs(a, b.cpu()) <= tol for a, b in zip(ls_ref, log_s_list))


def test_fp16_vs_oracle_medium_size():
  """B=4 x 2.3 s: against the oracle itself (CPU, a few seconds)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  B, F = 4, 200
  mel = synthetic_mel(B, F, 1001)
  zs = synthetic_noise(hp, B, F * 32, 9)
  with torch.no_grad():
    ref = O.infer(w, dims, mel, 0.666, zs)
  out = run(model, mel, zs, 0.666)
  err, snr = max_abs(ref, out), snr_db(ref, out)
  print(f"fp16 B4 F200: max-abs {err:.3e}  snr {snr:.1f} dB")
  assert err <= FP16_MAX_ABS and snr >= FP16_MIN_SNR_DB


def test_seed_semantics_and_determinism():
  """model.infer draws its noise from torch's global CUDA generator in the reference's order; same seed, same audio."""
  model = get_model(256, "fp16")
  mel = synthetic_mel(1, 16, 5).cuda()
  torch.manual_seed(123)
  a = model.infer(mel, sigma=0.7)
  torch.manual_seed(123)
  zs = model.draw_noise(1, 16 * 32, mel.device)
  b = model.infer_with_noise(mel, zs, 0.7)
  torch.manual_seed(123)
  c = model.infer(mel, sigma=0.7)
  assert torch.equal(a, b) and torch.equal(a, c)
  assert a.shape == (1, 16 * 256) and a.dtype == torch.float32 and a.device == mel.device
  assert torch.isfinite(a).all()


def test_batch_independence_and_locality_full_length():
  """Size-independent properties at the benchmark's utterance length (10 s, F=861):
  (1) utterance b of a batch equals the same utterance synthesised alone (no cross-batch interaction);
  (2) a WN sees +-255 group-steps per flow => changing the mel after frame f0 cannot change audio more than
      12 flows * 255 group-steps (+ the 4-frame upsampler support) before it."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  B, F = 3, 861
  mel = synthetic_mel(B, F, 1002)
  zs = synthetic_noise(hp, B, F * 32, 10)
  full = run(model, mel, zs, 0.666)
  assert torch.isfinite(full).all()
  alone = run(model, mel[1:2], [z[1:2] for z in zs], 0.666)
  assert torch.equal(full[1:2], alone)
  f0 = 600
  mel2 = mel.clone()
  mel2[:, :, f0:] = synthetic_mel(B, F - f0, 77)
  full2 = run(model, mel2, zs, 0.666)
  reach = (12 * 255 + 3 * 32) * 8
  cut = f0 * 256 - reach
  assert torch.equal(full[:, :cut], full2[:, :cut])
  assert not torch.equal(full[:, f0 * 256:], full2[:, f0 * 256:])


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name", ["forward_c256_b2_t2048", "forward_c256_b1_t4000"])
def test_forward_direction_matches_reference_golden(name, precision):
  """WaveGlow.forward (model.py:178-221) + WaveGlowLoss (train.py:31-45): z, log_s, sums, log-dets, loss.
  t4000 has T' = 500 (not a multiple of 32 nor 128): exercises the padded phase view of the tensor-core path."""
  from waveglow_b200.model import WaveGlowLoss
  g = load_golden(name)
  model = get_model(256, precision)
  z, log_s_list, log_det_list = model((torch.from_numpy(g["mel"]).cuda(), torch.from_numpy(g["audio_in"]).cuda()))
  torch.cuda.synchronize()
  z_ref = torch.from_numpy(g["z"])
  tol = 1e-5 if precision == "fp32" else 2e-3
  ez = max_abs(z_ref, z.cpu())
  print(f"forward {precision} {name}: z max-abs {ez:.3e} (ref absmax {float(z_ref.abs().max()):.2f})")
  assert z.shape == z_ref.shape and ez <= tol
  # sums: fp32 mode 1e-6 of sum|log_s|; fp16 mode 5e-4 (per-element gate is 2e-3 max-abs; observed 1.2e-4)
  rel = 1e-6 if precision == "fp32" else 5e-4
  n = z.size(0) * z.size(2)
  for k, ls in enumerate(log_s_list):
    ls_ref = torch.from_numpy(g[f"log_s{k}"])
    assert ls.shape == ls_ref.shape and max_abs(ls_ref, ls.cpu()) <= tol
    s_ref = float(g["log_s_sum"][k])
    s_got = float(model.last_forward_stats["log_s_sum"][k])
    # tolerance relative to the scale of the summands (the sum itself may cancel to ~0)
    scale = float(ls_ref.double().abs().sum())
    assert abs(s_got - s_ref) <= rel * scale + 1e-6, (k, s_got, s_ref)
    # the reference's own fp32 LU log-det carries ~2e-6 * B * T' of noise (tools/make_golden.py)
    assert abs(float(log_det_list[k]) - float(g["log_det_w"][k])) <= 2e-6 * n + 1e-5 * abs(float(g["log_det_w"][k]))
  loss = float(WaveGlowLoss(1.0)((z, log_s_list, log_det_list)))
  print(f"forward {precision} {name}: loss {loss:.6f} ref {float(g['loss']):.6f}")
  assert abs(loss - float(g["loss"])) <= (1e-5 if precision == "fp32" else 5e-4) * max(1.0, abs(float(g["loss"])))


def test_forward_then_infer_round_trip_on_gpu():
  """Size-independent property: infer(forward(audio) as noise, sigma=1) reproduces the audio (both fused paths)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  B, F = 2, 64
  mel = synthetic_mel(B, F, 90).cuda()
  audio = (torch.rand(B, F * 256, generator=torch.Generator().manual_seed(3)) * 2 - 1).cuda()
  z, _, _ = model((mel, audio))
  back = model.infer_with_noise(mel, [z[:, 4:].contiguous(), z[:, 2:4].contiguous(), z[:, :2].contiguous()], 1.0)
  torch.cuda.synchronize()
  err = max_abs(audio.cpu(), back.cpu())
  print(f"forward->infer round trip: max-abs {err:.3e}")
  assert err <= 4e-3


def test_512_channel_properties():
  """512-channel tensor-core path at a moderate size: finite, deterministic, no cross-utterance interaction, and
  infer(forward(audio)) == audio."""
  hp, sd, w, dims = synthetic_weights(512)
  model = get_model(512, "fp16")
  B, F = 3, 150
  mel = synthetic_mel(B, F, 71)
  zs = synthetic_noise(hp, B, F * 32, 72)
  full = run(model, mel, zs, 0.666)
  assert torch.isfinite(full).all() and torch.equal(full, run(model, mel, zs, 0.666))
  alone = run(model, mel[2:3], [z[2:3] for z in zs], 0.666)
  assert torch.equal(full[2:3], alone)
  audio = (torch.rand(B, F * 256, generator=torch.Generator().manual_seed(5)) * 2 - 1).cuda()
  z, _, _ = model((mel.cuda(), audio))
  back = model.infer_with_noise(mel.cuda(), [z[:, 4:].contiguous(), z[:, 2:4].contiguous(), z[:, :2].contiguous()], 1.0)
  torch.cuda.synchronize()
  err = max_abs(audio.cpu(), back.cpu())
  print(f"512-channel forward->infer round trip: max-abs {err:.3e}")
  assert err <= 4e-3


def test_cuda_graph_replay_is_bit_identical():
  """Opt-in CUDA-graph replay of the whole infer call (launch-bound single short utterances)."""
  hp, sd, w, dims = synthetic_weights(256)
  model = get_model(256, "fp16")
  mel = synthetic_mel(1, 40, 81)
  outs = []
  for seed in (1, 2):
    zs = synthetic_noise(hp, 1, 40 * 32, seed)
    plain = run(model, mel, zs, 0.7)
    model.use_cuda_graphs = True
    try:
      graphed = run(model, mel, zs, 0.7)      # first call captures, second replays
      graphed2 = run(model, mel, zs, 0.7)
    finally:
      model.use_cuda_graphs = False
    assert torch.equal(plain, graphed) and torch.equal(plain, graphed2)
    outs.append(plain)
  assert not torch.equal(outs[0], outs[1])


def test_errors_are_loud():
  from waveglow_b200.engine import WgError
  model = get_model(256, "fp16")
  mel = synthetic_mel(1, 4, 5).cuda()
  with pytest.raises(WgError, match="noise tensors"):
    model.infer_with_noise(mel, [torch.zeros(1, 4, 128).cuda()], 1.0)
  with pytest.raises(ValueError):
    model.infer_with_noise(torch.zeros(1, 3, 4).cuda(), [], 1.0)
  # raw pointers cross the ABI: a noise tensor of the wrong channel count, length, batch or device must not reach the kernels
  good = [torch.zeros(1, 4, 128), torch.zeros(1, 2, 128), torch.zeros(1, 2, 128)]
  for i, bad in ((0, torch.zeros(1, 2, 128)), (1, torch.zeros(1, 2, 96)), (2, torch.zeros(2, 2, 128))):
    zs = [z.cuda() for z in good]
    zs[i] = bad.cuda()
    with pytest.raises(WgError, match=f"noise tensor {i} has shape"):
      model.infer_with_noise(mel, zs, 1.0)
  with pytest.raises(WgError, match="is on cpu"):
    model.infer_with_noise(mel, [good[0].cuda(), good[1], good[2].cuda()], 1.0)
  # empty inputs: the reference raises as well (reshape of 0 elements / zero-length ConvTranspose1d input, model.py:225-232)
  for shape in ((0, 80, 4), (1, 80, 0)):
    with pytest.raises((WgError, ValueError, RuntimeError)):
      model.infer(torch.zeros(shape).cuda(), sigma=0.5)
  # ... and the engine still works afterwards
  assert torch.isfinite(model.infer(mel, sigma=0.5)).all()


def test_reloaded_weights_invalidate_cuda_graphs_and_pack_cache(tmp_path):
  """a captured CUDA graph bakes in the pointers of the packed weights, and the packed-weight cache identifies
  the checkpoint FILE, not the live parameters.  After load_state_dict both must be dropped: the next call runs the NEW
  weights (eagerly and through a fresh graph)."""
  from waveglow_b200.checkpoint import CheckpointWaveglow
  from waveglow_b200.synthesizer import Synthesizer
  from waveglow_b200.synthetic import synthetic_state_dict
  hp, sd, w, dims = synthetic_weights(256)
  sd2 = synthetic_state_dict(hp, 4321)
  mel = synthetic_mel(1, 24, 95)
  zs = synthetic_noise(hp, 1, 24 * 32, 96)
  cache = tmp_path / "ck.b200pack"
  synth = Synthesizer(CheckpointWaveglow.from_state_dict(sd, hp, iteration=1), device=torch.device("cuda:0"), pack_cache=cache,
                      pack_fingerprint=77)
  model = synth.model
  assert cache.is_file() and model.pack_cache is not None
  model.use_cuda_graphs = True
  old = run(model, mel, zs, 0.7)
  assert torch.equal(old, run(model, mel, zs, 0.7))             # replayed
  assert torch.equal(old, run(get_model(256, "fp16"), mel, zs, 0.7))
  # new weights into the same module (keys of the folded layout: remove_weightnorm already ran)
  fresh = WaveGlow(hp).to("cuda:0")
  fresh.load_state_dict(sd2)
  fresh = WaveGlow.remove_weightnorm(fresh).eval()
  expect = run(fresh, mel, zs, 0.7)
  model.load_state_dict(fresh.state_dict())
  assert model.pack_cache is None                               # the cache describes the old checkpoint file
  new = run(model, mel, zs, 0.7)                                # graph path: must re-capture against the re-packed weights
  assert not torch.equal(new, old) and torch.equal(new, expect)
  assert torch.equal(new, run(model, mel, zs, 0.7))
  model.use_cuda_graphs = False
  assert torch.equal(new, run(model, mel, zs, 0.7))
  assert model.pack_cache_hit is False
