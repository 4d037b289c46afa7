#!/usr/bin/env python
"""bench.py -- audio samples/s (and x real-time) of WaveGlow-256 `infer` on N B200s, next to the CPU reference path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--frames F] [--precision fp16|fp32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (`WaveGlow.infer`, reference model.py:223-274) over one batch of synthetic
mels.  Workload at every N (weak scaling, utterance-sharded, no data-path collective): BASELINE.json configs[1] per
GPU -- WaveGlow-256 (LJS-v3 arch, random init with end ~ N(0, 0.02)), batch 16 x 10 s (861 frames), sigma 0.666.
Prints ONE JSON line on rank 0 (contract in the task statement): `value` = samples/s with inputs resident in HBM,
`e2e` = the same through the public `model.infer` call with pinned-host mel in and pinned-host audio out,
`roofline` = the fused WN-layer tcgen05 kernel against the measured dense-bf16 peak, `cpu_baseline` = the oracle
port of the reference timed on this box's host cores on a bounded sample.
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
SIGMA = 0.666
CPU_SAMPLE_FRAMES = 200      # configs[0]-sized sample: 1 utterance x 200 frames = 51 200 samples


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
  """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

  def __init__(self, index: int):
    self.samples, self.reasons, self.max_mhz = [], set(), None
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
      self._thread = threading.Thread(target=self._run, daemon=True)
      self._thread.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    if self._thread is not None:
      self._thread.join()

  def summary(self):
    if self.nv is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
            "samples": len(s)}


def cpu_reference_run(frames: int, steps: int, warmup: int):
  """Times the oracle port of the reference's CPU `infer` (same torch ops as model.py) on all host cores."""
  import multiprocessing
  import torch
  from oracle import waveglow_oracle as O
  from waveglow_b200.hparams import HParams
  from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict
  cores = multiprocessing.cpu_count()
  torch.set_num_threads(cores)       # set_torch_thread_to_max(), reference utils.py:27-29
  hp = HParams()
  w = O.fold_weight_norm(synthetic_state_dict(hp, WEIGHT_SEED))
  dims = O.ModelDims()
  mel = synthetic_mel(1, frames, 1002)
  zs = synthetic_noise(hp, 1, frames * 32, 7)
  times = []
  with torch.no_grad():
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      O.infer(w, dims, mel, SIGMA, zs)
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
  n = frames * 256
  mean = sum(times) / len(times)
  return {"samples_per_step": n, "mean_s": mean, "best_s": min(times), "cores": cores}


def cpu_model_name():
  try:
    for line in open("/proc/cpuinfo"):
      if line.startswith("model name"):
        return line.split(":", 1)[1].strip()
  except Exception:
    pass
  return "unknown"


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--batch", type=int, default=16, help="utterances per GPU")
  ap.add_argument("--frames", type=int, default=861, help="mel frames per utterance (861 = 10 s)")
  ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
  ap.add_argument("--channels", type=int, default=256, choices=[256, 512], help="WN channels (512 = BASELINE config 3 architecture)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if args.channels == 256 and args.batch == 16 and args.frames == 861:
    which = "BASELINE.json configs[1]"
  elif args.channels == 512 and args.frames == 861:
    which = "BASELINE.json configs[2] architecture and utterance length"
  else:
    which = "non-headline shape"
  x_bytes = args.batch * ((args.frames + 127) // 128 * 128) * 32 * args.channels * 2
  config = {"workload": f"WaveGlow-{args.channels} infer, batch {args.batch} x {args.frames} frames ({args.frames * 256 / SAMPLING_RATE:.1f} s) per GPU, "
                        f"sigma {SIGMA} ({which})",
            "n_channels": args.channels, "batch_per_gpu": args.batch, "frames": args.frames, "sigma": SIGMA,
            "sharding": "by utterance, one process per GPU, no data-path collective; NCCL all_gather of the audio at the end of each step",
            "l2": f"working set per step (residual streams 2 x {x_bytes / 1e6:.0f} MB + packed weights > 1 GB) "
                  + (">> 126 MB L2, no flush needed" if 2 * x_bytes > 252e6 else "comparable to the 126 MB L2: small-shape run, not a headline number")}

  if args.impl == "reference":
    if rank != 0:
      return
    r = cpu_reference_run(CPU_SAMPLE_FRAMES, max(1, args.steps), max(1, args.warmup))
    val = r["samples_per_step"] / r["mean_s"]
    line = {"impl": "reference", "metric": "audio samples/sec, WaveGlow-256 infer", "value": val, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["mean_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "rtf_x_realtime": val / SAMPLING_RATE, "config": config,
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": r["cores"], "kind": "port", "cpu": cpu_model_name(),
                             "sample": f"1 utterance x {CPU_SAMPLE_FRAMES} frames ({r['samples_per_step']} samples) of the same synthetic workload per step; "
                                       "oracle/waveglow_oracle.py = the reference's torch CPU ops (model.py:223-274), all host threads"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return

  import torch
  import torch.distributed as dist
  from waveglow_b200.hparams import HParams
  from waveglow_b200.model import WaveGlow
  from waveglow_b200.sharding import gather_audio, shard_inputs
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
  B, F = args.batch, args.frames
  T = F * 256
  # noise is drawn for the global batch with one seed and sliced per rank, so results do not depend on N
  mel_g, zs_g = shard_inputs(synthetic_mel(B * world, F, 1002), synthetic_noise(hp, B * world, F * 32, 7), rank, world)
  mel_host = mel_g.contiguous().pin_memory()
  zs = [z.contiguous().to(dev) for z in zs_g]
  mel = mel_host.to(dev)
  audio_host = torch.empty((B, T), dtype=torch.float32).pin_memory()
  eng = model.engine()

  def step_device():
    out = model.infer_with_noise(mel, zs, SIGMA)
    if world > 1:
      gather_audio(out, B * world)         # the only collective: final gather of the audio over NCCL
    return out

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
  barrier()
  layer_ms, layer_launches = 0.0, 0
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
  with ClockSampler(local_rank) as clocks:
    barrier()
    evs[0].record()
    for i in range(args.steps):
      step_device()
      evs[i + 1].record()
    barrier()
  elapsed_ms = evs[0].elapsed_time(evs[-1])
  last_step_ms = evs[-2].elapsed_time(evs[-1])      # the step whose layer launches the engine's own events cover
  ms, n = eng.profile_layers()        # events recorded around the layer launches of the last timed step
  layer_ms, layer_launches = ms, n
  launches_per_step = eng.launch_count()
  out_check = step_device()
  finite = bool(torch.isfinite(out_check).all())

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

  total_samples = world * B * T * args.steps
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
  # skip->end fold (8 outputs on mma.sync in the 256-channel kernel, 16 tcgen05 columns in the wide path)
  nc = args.channels
  se = 2 * nc * (8 if nc == 256 else 16)
  g1_full, g1_first = 2 * (3 * nc + 320) * 2 * nc, 2 * (320 + 3 * 16) * 2 * nc
  g2_mid, g2_first = 2 * nc * nc + 2 * nc * 32, 2 * nc * nc + 2 * 16 * nc
  exec_per_step_layer = ((g1_first + g2_first + se) + 6 * (g1_full + g2_mid + se) + (g1_full + se)) / 8
  roofline = None
  if layer_launches:
    per_launch_flops = layer_flops_sample * B * T / layer_launches
    per_launch_ms = layer_ms / layer_launches
    achieved = per_launch_flops / (per_launch_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": ("wn_layer_fp32_kernel" if args.precision == "fp32" else
                           "wn_layer_pair_cont_kernel" if args.channels == 256 else "wn_gate_wide_kernel + wn_res_wide_kernel"),
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel on this workload, averaged over the 8
                # layers of a flow, from profiles/r01/ncu_full_layer_v12_first_and_generic_summary.csv (ncu --set full: generic
                # layer 271 + 214 MB, first layer 34 + 195 MB); only valid for the default workload
                "traffic": 453e6 if (args.batch == 16 and args.frames == 861 and args.precision == "fp16" and args.channels == 256) else None,
                "traffic_unit": "bytes per launch (ncu)",
                "peak_source": peak_src, "launch_ms": per_launch_ms, "launches_per_step": layer_launches,
                "algorithmic_flop_per_launch": per_launch_flops,
                "executed_tensor_tflops": exec_per_step_layer * (B * T / 8) / (per_launch_ms * 1e-3) / 1e12,
                "executed_over_algorithmic": exec_per_step_layer * (B * T / 8) / per_launch_flops,
                "layer_kernels_share_of_step": layer_ms / last_step_ms}
  line = {"metric": f"audio samples/sec, WaveGlow-{args.channels} infer", "value": value, "unit": "samples/s", "n_gpus": world,
          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "fp16 operands, fp32 accumulate (tcgen05 kind::f16)" if args.precision == "fp16" else "fp32",
          "data": "synthetic", "rtf_x_realtime": value / SAMPLING_RATE, "rtf_x_realtime_per_gpu": value / SAMPLING_RATE / world,
          "algorithmic_tflops": value * flops_sample / 1e12, "config": config, "clocks": clocks.summary(),
          "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": mel_host.numel() * 4,
                  "d2h_bytes_per_step": audio_host.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                  "api": "WaveGlow.infer(spect, sigma) with pinned-host mel in, pinned-host audio out, noise drawn on device"},
          "gpu_launches": launches_per_step * args.steps, "output_finite": finite, "roofline": roofline}
  if not args.no_cpu_baseline and world == 1:
    r = cpu_reference_run(CPU_SAMPLE_FRAMES, 3, 1)
    v = r["samples_per_step"] / r["best_s"]
    line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": r["cores"], "kind": "port", "cpu": cpu_model_name(),
                            "rtf_x_realtime": v / SAMPLING_RATE,
                            "sample": f"1 utterance x {CPU_SAMPLE_FRAMES} frames ({r['samples_per_step']} samples), 1 warm-up + best of 3; "
                                      "oracle port of the reference's torch CPU ops (model.py:223-274), all host threads"}
  print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
