#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: duration, DRAM bytes, tensor-pipe and DRAM utilisation, achieved
occupancy -- the numbers quoted in profiles/ and the `traffic` field of bench.py's roofline (profiles/r02_traffic.json).

    python tools/ncu_kernel_summary.py gpurun_out/a_prof_conv.ncu-rep [--json profiles/r02_traffic.json --key c4/bf16 --pairs 704]

Reads the report with `ncu -i <rep> --page raw --csv` (run here, no GPU needed)."""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

METRICS = {
  "gpu__time_duration.sum": "ns",
  "dram__bytes_read.sum": "B",
  "dram__bytes_write.sum": "B",
  "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "%",
  "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed": "%",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed": "%",
  "sm__warps_active.avg.pct_of_peak_sustained_active": "%",
  "lts__t_bytes.sum": "B",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3,
         "ms": 1e6, "msecond": 1e6, "second": 1e9, "s": 1e9, "%": 1.0, "": 1.0}


def rows(rep):
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
  start = out.find('"ID"')
  r = list(csv.reader(io.StringIO(out[start:])))
  header, units, data = r[0], r[1], r[2:]
  return header, units, data


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("rep")
  ap.add_argument("--json")
  ap.add_argument("--key")
  ap.add_argument("--pairs", type=int, default=0)
  ap.add_argument("--kernel", default="", help="regex: the kernel whose DRAM bytes go into --json (default: longest total)")
  a = ap.parse_args()
  header, units, data = rows(a.rep)
  col = {h: i for i, h in enumerate(header)}
  name_i = col["Kernel Name"]
  agg = {}
  for d in data:
    name = re.sub(r"\(.*$", "", d[name_i])[:90]
    e = agg.setdefault(name, {"n": 0})
    e["n"] += 1
    for m in METRICS:
      if m in col and d[col[m]] not in ("", "n/a", "no data"):
        try:
          v = float(d[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        except ValueError:
          continue
        e[m] = e.get(m, 0.0) + v
  print("| kernel | launches | avg us | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | DRAM % | warps active % |")
  print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
  order = sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0.0))
  for name, e in order:
    n = e["n"]
    t = e.get("gpu__time_duration.sum", 0.0) / n
    rd, wr = e.get("dram__bytes_read.sum", 0.0) / n, e.get("dram__bytes_write.sum", 0.0) / n
    print("| `%s` | %d | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f |" % (
      name, n, t / 1e3, rd / 1e6, wr / 1e6, (rd + wr) / max(t, 1.0),
      e.get("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 0.0) / n,
      e.get("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) / n,
      e.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) / n))
  if a.json and a.key:
    pick = None
    for name, e in order:
      if (not a.kernel) or re.search(a.kernel, name):
        pick = (name, e)
        break
    if pick:
      name, e = pick
      n = e["n"]
      try:
        cur = json.load(open(a.json))
      except Exception:
        cur = {}
      cur[a.key] = {"kernel": name, "launches_captured": n, "pairs_per_gpu": a.pairs,
                    "dram_bytes_per_launch": (e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)) / n,
                    "dram_read_bytes_per_launch": e.get("dram__bytes_read.sum", 0.0) / n,
                    "dram_write_bytes_per_launch": e.get("dram__bytes_write.sum", 0.0) / n,
                    "avg_us_under_ncu": e.get("gpu__time_duration.sum", 0.0) / n / 1e3,
                    "source": "ncu --set full --clock-control none, %s (tools/ncu_kernel_summary.py)" % a.rep.split("/")[-1]}
      json.dump(cur, open(a.json, "w"), indent=1)
      print("\nwrote %s[%s] = %s" % (a.json, a.key, name), file=sys.stderr)


if __name__ == "__main__":
  main()
