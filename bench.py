#!/usr/bin/env python
"""bench.py -- audio samples/s (and x real-time) of WaveGlow `infer` on N B200s, next to the CPU reference path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--frames F] [--precision fp16|fp32]
                  [--channels 256|512] [--scaling weak|strong --global-batch G]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (`WaveGlow.infer`, reference model.py:223-274) over one batch of synthetic mels.
Default workload at every N (weak scaling, utterance-sharded, no data-path collective): BASELINE.json configs[1] per GPU --
WaveGlow-256 (LJS-v3 arch, random init with end ~ N(0, 0.02)), batch 16 x 10 s (861 frames), sigma 0.666.
`--scaling strong --channels 512 --global-batch 32` is BASELINE.json configs[2]: a fixed global batch split 32/16/8/4 per GPU.
Prints ONE JSON line on rank 0 (contract in the task statement): `value` = samples/s with inputs resident in HBM,
`e2e` = the same through the public `model.infer` call with pinned-host mel in and pinned-host audio out,
`roofline` = the fused WN-layer tcgen05 kernel against the measured dense-bf16 peak, `parity` = utterance 0 of the timed
batch against the CPU oracle, `cpu_baseline` = the reference's CPU path timed on this box's host cores on a bounded sample.
`--impl reference` times that CPU path alone: the UNMODIFIED reference (`waveglow.model.WaveGlow.infer` through
tools/ref_loader.py) where /root/reference exists, else the oracle port of the same torch ops.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

SAMPLING_RATE = 22050
WEIGHT_SEED = 1234
MEL_SEED = 1002
NOISE_SEED = 7
SIGMA = 0.666


def algorithmic_flops_per_sample(nc: int, n_flows: int = 12, n_layers: int = 8, n_group: int = 8):
  """SURVEY.md §8d counting rule: 2 FLOP per MAC, dilated conv dense K = 3 nc, all per group-step, summed over flows."""
  chans = [8, 8, 8, 8, 6, 6, 6, 6, 4, 4, 4, 4][:n_flows]
  per_step = 0
  layer_part = 0
  for c in chans:
    h = c // 2
    start = h * nc
    cond = 640 * 2 * nc * n_layers
    inl = n_layers * 3 * nc * 2 * nc
    rs = (n_layers - 1) * nc * 2 * nc + nc * nc
    end = nc * 2 * h
    inv = c * c
    per_step += 2 * (start + cond + inl + rs + end + inv)
    layer_part += 2 * (cond + inl + rs)
  return per_step / n_group, layer_part / n_group


class ClockSampler:
  """Samples SM clock, throttle reasons and board power of one GPU through NVML while the timed region runs, and reads the
  board's energy counter on both sides of it (the 1000 W cap sets the clock in this workload: J per step is what ranks
  kernel variants)."""

  def __init__(self, index: int):
    self.samples, self.power_w, self.reasons, self.max_mhz = [], [], set(), None
    self.energy_mj = [None, None]
    self._stop = threading.Event()
    self._thread = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception as ex:  # NVML missing: report that instead of inventing numbers
      self.nv = None
      self.err = str(ex)

  def _energy(self):
    try:
      return self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h)      # millijoules since driver load
    except Exception:
      return None

  def _run(self):
    nv = self.nv
    names = {
      "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
      "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
      "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
      "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
    }
    while not self._stop.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
          self.power_w.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
        except Exception:
          pass
        try:
          mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
          mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in names.items():
          if mask & bit:
            self.reasons.add(name)
      except Exception:
        pass
      self._stop.wait(0.05)

  def __enter__(self):
    if self.nv is not None:
      self.energy_mj[0] = self._energy()
      self._thread = threading.Thread(target=self._run, daemon=True)
      self._thread.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    if self._thread is not None:
      self._thread.join()
      self.energy_mj[1] = self._energy()

  def summary(self):
    if self.nv is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
    s = sorted(self.samples)
    out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
           "samples": len(s)}
    if self.power_w:
      p = sorted(self.power_w)
      out["power_w_median"] = p[len(p) // 2]
    return out

  def energy_joules(self):
    a, b = self.energy_mj
    return None if a is None or b is None else (b - a) / 1e3


def reference_available() -> bool:
  return os.path.isdir("/root/reference/src/waveglow")


def cpu_reference_run(channels: int, frames: int, steps: int, warmup: int):
  """Times the reference's CPU `infer` on ONE utterance of `frames` mel frames, all host cores (the reference's
  `set_torch_thread_to_max`, utils.py:27-29).  kind "reference": the unmodified `waveglow.model.WaveGlow.infer`
  (model.py:223-274) imported from /root/reference through tools/ref_loader.py -- only where that tree exists (the build
  container); kind "port": oracle/waveglow_oracle.py, the same torch CPU ops, bit-identical to it (tests/test_oracle.py)."""
  import multiprocessing
  import torch
  from waveglow_b200.hparams import HParams
  from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict
  cores = multiprocessing.cpu_count()
  torch.set_num_threads(cores)
  hp = HParams(n_channels=channels)
  sd = synthetic_state_dict(hp, WEIGHT_SEED)
  mel = synthetic_mel(1, frames, MEL_SEED)
  if reference_available():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from ref_loader import load_reference
    RefWaveGlow, RefHParams = load_reference()
    model = RefWaveGlow(RefHParams(n_channels=channels))
    model.load_state_dict(sd)
    model = RefWaveGlow.remove_weightnorm(model).eval()
    kind, what = "reference", "the UNMODIFIED reference waveglow.model.WaveGlow.infer (model.py:223-274) via tools/ref_loader.py"

    def run():
      torch.manual_seed(NOISE_SEED)
      return model.infer(mel, sigma=SIGMA)
  else:
    from oracle import waveglow_oracle as O
    w = O.fold_weight_norm(sd)
    dims = O.ModelDims(n_channels=channels)
    zs = synthetic_noise(hp, 1, frames * 32, NOISE_SEED)
    kind, what = "port", "oracle/waveglow_oracle.py, the reference's own torch CPU ops (model.py:223-274; /root/reference is not on this box)"

    def run():
      return O.infer(w, dims, mel, SIGMA, zs)
  times = []
  with torch.no_grad():
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      run()
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
  return {"samples_per_step": frames * 256, "mean_s": sum(times) / len(times), "best_s": min(times), "cores": cores,
          "kind": kind, "what": what}


def cpu_model_name():
  try:
    for line in open("/proc/cpuinfo"):
      if line.startswith("model name"):
        return line.split(":", 1)[1].strip()
  except Exception:
    pass
  return "unknown"


def oracle_parity(channels: int, mel0, zs0, out0):
  """Utterance 0 of the timed batch against the CPU oracle (fp32 reference arithmetic) on the same mel, weights and noise."""
  import torch
  from oracle import waveglow_oracle as O
  from waveglow_b200.hparams import HParams
  from waveglow_b200.synthetic import synthetic_state_dict
  hp = HParams(n_channels=channels)
  w = O.fold_weight_norm(synthetic_state_dict(hp, WEIGHT_SEED))
  dims = O.ModelDims(n_channels=channels)
  t0 = time.perf_counter()
  with torch.no_grad():
    ref = O.infer(w, dims, mel0, SIGMA, zs0)
  dt = time.perf_counter() - t0
  err = (out0.double() - ref.double())
  snr = float(10 * torch.log10((ref.double() ** 2).sum() / (err ** 2).sum().clamp_min(1e-300)))
  return {"max_abs": float(err.abs().max()), "snr_db": snr, "shape": list(ref.shape), "ref_abs_max": float(ref.abs().max()),
          "vs": "oracle/waveglow_oracle.py (fp32 CPU, bit-identical to the reference's infer) on utterance 0 of the timed batch",
          "oracle_seconds": dt}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--batch", type=int, default=16, help="utterances per GPU (weak scaling)")
  ap.add_argument("--frames", type=int, default=861, help="mel frames per utterance (861 = 10 s)")
  ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
  ap.add_argument("--channels", type=int, default=256, choices=[256, 512], help="WN channels (512 = BASELINE config 3 architecture)")
  ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                  help="weak: --batch utterances per GPU; strong: --global-batch utterances split over the GPUs (SURVEY §8e)")
  ap.add_argument("--global-batch", type=int, default=32, help="utterances in total with --scaling strong (config 3: 32)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-parity", action="store_true")
  args = ap.parse_args()

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  strong = args.scaling == "strong"
  global_batch = args.global_batch if strong else args.batch * world
  if not strong and args.channels == 256 and args.batch == 16 and args.frames == 861:
    which = "BASELINE.json configs[1]"
  elif strong and args.channels == 512 and args.frames == 861 and args.global_batch == 32:
    which = "BASELINE.json configs[2]"
  elif args.channels == 512 and args.frames == 861:
    which = "BASELINE.json configs[2] architecture and utterance length"
  else:
    which = "non-headline shape"
  per_gpu_max = -(-global_batch // world)
  x_bytes = per_gpu_max * ((args.frames + 127) // 128 * 128) * 32 * args.channels * 2
  secs = args.frames * 256 / SAMPLING_RATE
  if strong:
    split = f"global batch {global_batch} x {args.frames} frames ({secs:.1f} s) split over the GPUs ({per_gpu_max} per GPU at N={args.gpus})"
  else:
    split = f"batch {args.batch} x {args.frames} frames ({secs:.1f} s) per GPU"
  config = {"workload": f"WaveGlow-{args.channels} infer, {split}, sigma {SIGMA} ({which})",
            "n_channels": args.channels, "batch_per_gpu": per_gpu_max, "frames": args.frames, "sigma": SIGMA,
            "sharding": "by utterance, one process per GPU, no data-path collective; the NCCL all_gather of each step's audio "
                        "runs asynchronously under the next step's kernels",
            "reference_arm_sample": f"the CPU arm (--impl reference, cpu_baseline) times ONE utterance of this workload per step "
                                    f"(1 x {args.frames} frames = {args.frames * 256} samples), all host cores; samples/s is per sample, so the "
                                    "ratio compares equal utterance lengths",
            "l2": f"working set per step (residual streams 2 x {x_bytes / 1e6:.0f} MB + packed weights > 1 GB) "
                  + (">> 126 MB L2, no flush needed" if 2 * x_bytes > 252e6 else "comparable to the 126 MB L2: small-shape run, not a headline number")}
  metric = f"audio samples/sec, WaveGlow-{args.channels} infer"
  scaling = "strong" if strong else "weak"

  if args.impl == "reference":
    if rank != 0:
      return
    r = cpu_reference_run(args.channels, args.frames, max(1, args.steps), max(0, args.warmup))
    val = r["samples_per_step"] / r["mean_s"]
    line = {"impl": "reference", "metric": metric, "value": val, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["mean_s"] * 1e3,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "rtf_x_realtime": val / SAMPLING_RATE, "config": config,
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "cpu": cpu_model_name(),
                             "sample": f"1 utterance x {args.frames} frames ({r['samples_per_step']} samples) of the workload per step, "
                                       f"mean of {max(1, args.steps)} steps after {max(0, args.warmup)} warm-up; {r['what']}; all host threads"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return

  import torch
  import torch.distributed as dist
  from waveglow_b200.hparams import HParams
  from waveglow_b200.model import WaveGlow
  from waveglow_b200.sharding import AudioGather, shard_inputs, shard_range
  from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict
  import __graft_entry__ as graft

  if not torch.cuda.is_available():
    raise SystemExit("bench.py --impl b200 needs a GPU (no CPU fallback)")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  if rank == 0:
    graft.build()                 # no-op when the in-tree library matches the sources (content hash)
  if world > 1:
    dist.barrier()                # nobody loads the library before rank 0 is done with it

  hp = HParams(n_channels=args.channels)
  model = WaveGlow(hp, precision=args.precision).to(dev)
  model.load_state_dict(synthetic_state_dict(hp, WEIGHT_SEED))
  model = WaveGlow.remove_weightnorm(model).eval()
  F = args.frames
  T = F * 256
  lo, hi = shard_range(global_batch, rank, world)
  B = hi - lo
  # noise is drawn for the global batch with one seed and sliced per rank, so results do not depend on N
  mel_g, zs_g = shard_inputs(synthetic_mel(global_batch, F, MEL_SEED), synthetic_noise(hp, global_batch, F * 32, NOISE_SEED), rank, world)
  mel_host = mel_g.contiguous().pin_memory()
  zs = [z.contiguous().to(dev) for z in zs_g]
  mel = mel_host.to(dev)
  audio_host = torch.empty((B, T), dtype=torch.float32).pin_memory()
  eng = model.engine()
  pending = [None]

  def step_device():
    out = model.infer_with_noise(mel, zs, SIGMA)
    if world > 1:
      # the only collective: final gather of the audio over NCCL, asynchronous -- it overlaps the NEXT step's kernels
      # and is joined one step later (and before the timed region closes)
      if pending[0] is not None:
        pending[0].wait()
      pending[0] = AudioGather(out, global_batch)
    return out

  def drain():
    if pending[0] is not None:
      pending[0].wait()
      pending[0] = None

  def step_e2e():
    m = mel_host.to(dev, non_blocking=True)
    out = model.infer(m, sigma=SIGMA)        # public API: draws its noise on the device like the reference
    audio_host.copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(max(args.warmup, 3)):
    step_device()
  drain()
  barrier()
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
  with ClockSampler(local_rank) as clocks:
    barrier()
    evs[0].record()
    for i in range(args.steps):
      step_device()
      evs[i + 1].record()
    drain()
    end_ev = torch.cuda.Event(enable_timing=True)
    end_ev.record()
    barrier()
  elapsed_ms = evs[0].elapsed_time(end_ev)
  last_step_ms = evs[-2].elapsed_time(evs[-1])      # the step whose layer launches the engine's own events cover
  layer_ms, layer_launches = eng.profile_layers()   # events recorded around the layer launches of the last timed step
  launches_per_step = eng.launch_count()
  energy_j = clocks.energy_joules()
  out_check = step_device()
  drain()
  torch.cuda.synchronize()
  finite = bool(torch.isfinite(out_check).all())
  parity = None
  if rank == 0 and not args.no_parity:
    parity = oracle_parity(args.channels, mel_g[0:1].contiguous(), [z[0:1].contiguous() for z in zs_g], out_check[0:1].cpu())
    parity["gate"] = {"max_abs": 2e-3, "snr_db": 40.0} if args.precision == "fp16" else {"max_abs": 1e-5}
    parity["pass"] = bool(parity["max_abs"] <= parity["gate"]["max_abs"] and parity["snr_db"] >= parity["gate"].get("snr_db", -1e9))
  if world > 1:
    dist.barrier()

  for _ in range(2):
    step_e2e()
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step_e2e()
  barrier()
  e2e_s = time.perf_counter() - t0

  t_dev = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
  elapsed_ms, e2e_ms = float(t_dev[0]), float(t_dev[1])
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  total_samples = global_batch * T * args.steps
  value = total_samples / (elapsed_ms * 1e-3)
  e2e_value = total_samples / (e2e_ms * 1e-3)
  flops_sample, layer_flops_sample = algorithmic_flops_per_sample(args.channels)
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
  peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step; fp16 and bf16 share the tcgen05 kind::f16 rate)" \
      if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
  # tensor FLOPs the layer kernels actually issue per group-step, averaged over the 8 layers of a flow:
  # GEMM1 K = 3 nc + 320 (upsampler composed into the conditioning), K = 320 + 3 x 16 in a flow's first layer (start conv
  # composed in); GEMM2 N = nc residual + identity carry (one K = 16 step in the first layer, none in the last);
  # skip->end fold (8 outputs per group-step on warp-level mma.sync)
  nc = args.channels
  se = 2 * nc * 8
  g1_full, g1_first = 2 * (3 * nc + 320) * 2 * nc, 2 * (320 + 3 * 16) * 2 * nc
  g2_mid, g2_first = 2 * nc * nc + 2 * nc * 32, 2 * nc * nc + 2 * 16 * nc
  exec_per_step_layer = ((g1_first + g2_first + se) + 6 * (g1_full + g2_mid + se) + (g1_full + se)) / 8
  kernel = "wn_layer_fp32_kernel" if args.precision == "fp32" else f"wn_layer_pair_cont_kernel<{args.channels}, 64, first, mode>"
  # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over a flow's launches: taken from the committed
  # capture of this workload (profiles/traffic.json names the .csv it came from), null when this workload was not captured
  traffic, traffic_src = None, None
  try:
    tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    ent = tr.get(f"c{nc}_b{B}_f{F}_{args.precision}")
    if ent:
      traffic, traffic_src = ent["bytes_per_launch"], ent["source"]
  except Exception:
    pass
  roofline = None
  if layer_launches:
    per_launch_flops = layer_flops_sample * B * T / layer_launches
    per_launch_ms = layer_ms / layer_launches
    achieved = per_launch_flops / (per_launch_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": kernel,
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)", "traffic_source": traffic_src,
                "peak_source": peak_src, "launch_ms": per_launch_ms, "launches_per_step": layer_launches,
                "launch_unit": "one WN layer (the unit of launch_ms / launches_per_step / algorithmic_flop_per_launch); the generic "
                               "layers of a flow run as one chained launch of this kernel, so a step has fewer launches than layers "
                               "(gpu_launches counts real launches)",
                "algorithmic_flop_per_launch": per_launch_flops,
                "executed_tensor_tflops": exec_per_step_layer * (B * T / 8) / (per_launch_ms * 1e-3) / 1e12,
                "executed_over_algorithmic": exec_per_step_layer * (B * T / 8) / per_launch_flops,
                "executed_frac_of_peak": exec_per_step_layer * (B * T / 8) / (per_launch_ms * 1e-3) / 1e12 / peak_tf,
                "layer_kernels_share_of_step": layer_ms / last_step_ms}
  line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world,
          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
          "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
          "dtype": "fp16 operands, fp32 accumulate (tcgen05 kind::f16)" if args.precision == "fp16" else "fp32",
          "data": "synthetic", "rtf_x_realtime": value / SAMPLING_RATE, "rtf_x_realtime_per_gpu": value / SAMPLING_RATE / world,
          "algorithmic_tflops": value * flops_sample / 1e12, "config": config, "clocks": clocks.summary(),
          "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": mel_host.numel() * 4,
                  "d2h_bytes_per_step": audio_host.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                  "api": "WaveGlow.infer(spect, sigma) with pinned-host mel in, pinned-host audio out, noise drawn on device"},
          "gpu_launches": launches_per_step * args.steps, "output_finite": finite, "parity": parity, "roofline": roofline}
  if energy_j is not None:
    line["energy"] = {"j_per_step": energy_j / args.steps, "avg_w": energy_j / (elapsed_ms * 1e-3),
                      "source": "NVML total-energy counter of rank 0's GPU around the timed region"}
  if not args.no_cpu_baseline and world == 1:
    r = cpu_reference_run(args.channels, F, 2, 1)
    v = r["samples_per_step"] / r["best_s"]
    line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "cpu": cpu_model_name(),
                            "rtf_x_realtime": v / SAMPLING_RATE,
                            "sample": f"1 utterance x {F} frames ({r['samples_per_step']} samples) of the workload, 1 warm-up + best of 2; "
                                      f"{r['what']}; all host threads"}
  print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
