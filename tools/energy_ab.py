"""Energy per step of kernel variants on ONE box: the layer kernel runs at the 1000 W power cap, so the
SM clock -- and with it the time of a step -- follows the energy a step needs.  Each variant (a set of the engine's
environment switches, read once in wg_create) runs in its own process: 3 warm-up steps, then steps for >= `--seconds` with the
board's NVML energy counter read on both sides of the synchronised region; reported: ms per step (CUDA events), J per step,
average W, median SM clock.  Variants with WG_DBG_FLAGS switch PARTS of the kernel off (their audio is invalid): they attribute
energy, they are not candidates.

  python tools/energy_ab.py                     # all variants, 256 channels 16 x 861 and 512 channels 8 x 861
  python tools/energy_ab.py --child 256 16      # one measurement in this process with the current environment
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS_256 = [
  ("production", {}),
  ("one launch per layer (WG_CHAIN=0)", {"WG_CHAIN": "0"}),
  ("no programmatic dependent launch (WG_PDL=0)", {"WG_PDL": "0"}),
  ("no packed ragged tails (WG_PACK4=0)", {"WG_PACK4": "0"}),
  ("explicit start conv (WG_FIRST=0)", {"WG_FIRST": "0"}),
  ("smallest grid with the same tile rounds (WG_TRIM=1)", {"WG_TRIM": "1"}),
  ("[invalid] operand loads off (flags 1, per-layer launches)", {"WG_DBG_FLAGS": "1", "WG_CHAIN": "0"}),
  ("[invalid] gate math off (flags 4)", {"WG_DBG_FLAGS": "4"}),
  ("[invalid] x' stores off (flags 8)", {"WG_DBG_FLAGS": "8"}),
  ("[invalid] x' stores as full 128-byte lines, rows scrambled (flags 8192)", {"WG_DBG_FLAGS": "8192"}),
  ("[invalid] skip-end GEMM off (flags 16)", {"WG_DBG_FLAGS": "16"}),
  ("[invalid] identity carry off (flags 128)", {"WG_DBG_FLAGS": "128"}),
  ("[invalid] all of the above off (flags 157, per-layer launches)", {"WG_DBG_FLAGS": "157", "WG_CHAIN": "0"}),
  ("[diagnostic] chained, no dependency polls (flags 256)", {"WG_DBG_FLAGS": "256"}),
  ("[diagnostic] chained, no polls, no completion counts (flags 768)", {"WG_DBG_FLAGS": "768"}),
  ("[diagnostic] chained, static items, no polls / counts (flags 1792)", {"WG_DBG_FLAGS": "1792"}),
  ("one launch per layer (WG_CHAIN=0, again)", {"WG_CHAIN": "0"}),
  ("production (again)", {}),
]
VARIANTS_512 = [
  ("production: exact hi + lo residual carry", {}),
  ("residual stream rounded per layer (WG_EXACT_CARRY=0)", {"WG_EXACT_CARRY": "0"}),
  ("one launch per layer (WG_CHAIN=0)", {"WG_CHAIN": "0"}),
  ("production (again)", {}),
]


def child(channels, batch, seconds):
  import pynvml
  import torch
  from waveglow_b200.hparams import HParams
  from waveglow_b200.model import WaveGlow
  from waveglow_b200.synthetic import synthetic_mel, synthetic_noise, synthetic_state_dict
  pynvml.nvmlInit()
  h = pynvml.nvmlDeviceGetHandleByIndex(0)
  hp = HParams(n_channels=channels)
  m = WaveGlow(hp, precision="fp16").to("cuda:0")
  m.load_state_dict(synthetic_state_dict(hp, 1234))
  m = WaveGlow.remove_weightnorm(m).eval()
  F = 861
  mel = synthetic_mel(batch, F, 1002).cuda()
  zs = [z.cuda() for z in synthetic_noise(hp, batch, F * 32, 7)]
  for _ in range(3):
    m.infer_with_noise(mel, zs, 0.666)
  torch.cuda.synchronize()
  clocks, stop = [], threading.Event()

  def sample():
    while not stop.is_set():
      clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
      stop.wait(0.05)

  th = threading.Thread(target=sample, daemon=True)
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  th.start()
  e0, t0, steps = pynvml.nvmlDeviceGetTotalEnergyConsumption(h), time.perf_counter(), 0
  ev0.record()
  while time.perf_counter() - t0 < seconds:
    for _ in range(4):
      m.infer_with_noise(mel, zs, 0.666)
    steps += 4
    torch.cuda.synchronize()
  ev1.record()
  torch.cuda.synchronize()
  e1, t1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h), time.perf_counter()
  stop.set()
  th.join()
  print(json.dumps({"ms_per_step": ev0.elapsed_time(ev1) / steps, "j_per_step": (e1 - e0) / 1e3 / steps, "avg_w": (e1 - e0) / 1e3 / (t1 - t0),
                    "steps": steps, "sm_mhz_median": statistics.median(clocks) if clocks else None}))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--child", nargs=2, type=int, default=None)
  ap.add_argument("--seconds", type=float, default=5.0)
  ap.add_argument("--only", type=int, default=0, help="256 or 512: only that channel count")
  ap.add_argument("--match", default="", help="only variants whose name contains one of these comma-separated substrings")
  args = ap.parse_args()
  if args.child:
    return child(args.child[0], args.child[1], args.seconds)
  for channels, batch, variants in ((256, 16, VARIANTS_256), (512, 8, VARIANTS_512)):
    if args.only and args.only != channels:
      continue
    print(f"# {channels} channels, {batch} x 861 frames per step; >= {args.seconds:.0f} s of back-to-back steps per variant, NVML energy counter")
    print(f"{'variant':58s} {'ms/step':>8s} {'J/step':>8s} {'avg W':>7s} {'SM MHz':>7s}")
    for name, env in variants:
      if args.match and not any(m in name for m in args.match.split(",")):
        continue
      e = dict(os.environ)
      for k in ("WG_CHAIN", "WG_PDL", "WG_PACK4", "WG_FIRST", "WG_TRIM", "WG_DBG_FLAGS", "WG_EXACT_CARRY"):
        e.pop(k, None)
      e.update(env)
      r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(channels), str(batch), "--seconds", str(args.seconds)],
                         env=e, capture_output=True, text=True, timeout=300)
      line = [l for l in r.stdout.splitlines() if l.startswith("{")]
      if r.returncode != 0 or not line:
        print(f"{name:58s} failed: {r.stderr.strip().splitlines()[-1][:120] if r.stderr.strip() else r.returncode}")
        continue
      d = json.loads(line[-1])
      print(f"{name:58s} {d['ms_per_step']:8.2f} {d['j_per_step']:8.2f} {d['avg_w']:7.0f} {d['sm_mhz_median'] or 0:7.0f}", flush=True)
    print()


if __name__ == "__main__":
  main()
