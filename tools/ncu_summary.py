"""Condense ncu output into the small text/CSV summaries kept under profiles/ (the .ncu-rep files stay in gpurun_out/).

  python tools/ncu_summary.py full  gpurun_out/prof_layer.ncu-rep  > profiles/rNN/ncu_full_<what>_summary.csv
  python tools/ncu_summary.py list  gpurun_out/launches.csv "<command that was profiled>"  > profiles/rNN/ncu_launches_<what>_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
  "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
  "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
  "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
  "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
  "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
  "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
  "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__block_size",
  "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
  "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
  "smsp__inst_executed.sum",
]


def full(rep):
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  head, units, data = rows[0], rows[1], rows[2:]
  col = {n: i for i, n in enumerate(head)}
  w = csv.writer(sys.stdout, quoting=csv.QUOTE_NONNUMERIC)
  sys.stdout.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))) + "\n")
  w.writerow(["Kernel Name", ""] + [d[col["Kernel Name"]][:60] for d in data])
  for m in METRICS:
    if m in col:
      w.writerow([m, units[col[m]]] + [d[col[m]] for d in data])


def launch_list(path, what):
  """Aggregates per kernel name: all captured launches, and the LAST `infer` call of the run (from its pack_noise_kernel launch
  to the end of the capture: the timed / end-to-end step of the profiled bench command)."""
  lines = [l for l in open(path) if not l.startswith("==")]
  launches = []
  for row in csv.DictReader(lines):
    try:
      v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
      continue
    if row.get("Metric Unit", "") in ("ns", "nsecond"):
      v /= 1e3
    elif row.get("Metric Unit", "") in ("ms", "msecond"):
      v *= 1e3
    launches.append((row["Kernel Name"].split("(")[0][:70], v))

  def table(rows, title):
    agg = collections.OrderedDict()
    for name, v in rows:
      agg.setdefault(name, []).append(v)
    total = sum(v for _, v in rows)
    print(f"{title}: {len(rows)} launches, total {total / 1e3:.2f} ms")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
      print(f"{name:72s} launches {len(v):4d}  total {sum(v) / 1e3:9.3f} ms  share {100 * sum(v) / total:5.1f}%  avg {sum(v) / len(v):8.1f} us")

  print(f"ncu launch list (gpu__time_duration.sum, cold-cache, serialised) of: {what}")
  starts = [i for i, (name, _) in enumerate(launches) if "pack_noise_kernel" in name]
  if starts:
    table(launches[starts[-1]:], "last infer call of the run (one step)")
    print()
  table(launches, "whole capture")


if __name__ == "__main__":
  if sys.argv[1] == "full":
    full(sys.argv[2])
  else:
    launch_list(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
