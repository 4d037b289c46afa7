"""stdin: bench.py's JSON line -> one short line (A/B scripts)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d.get("roofline") or {}
e = d.get("energy") or {}
print("ms_per_step %.2f  x_realtime %.0f  e2e_ms %.2f  launch_ms %.4f  exec_frac %.3f  J/step %.1f  clocks %s" % (
  d["ms_per_step"], d["rtf_x_realtime"], d["e2e"]["ms_per_step"], r.get("launch_ms", 0), r.get("executed_frac_of_peak", 0),
  e.get("j_per_step", 0), d["clocks"]))
