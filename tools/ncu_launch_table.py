#!/usr/bin/env python
"""Turns an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`)
into the per-kernel markdown table kept under profiles/ (launches, total ms, share of the GPU time of the command).

    python tools/ncu_launch_table.py gpurun_out/launches.csv "title" "command" > profiles/rNN_launch_shares.md

Per-launch times under ncu are cold-cache and serialised: the SHARES are what is compared with the live CUDA-event
breakdown of bench.py (roofline.by_kind / other_kernels_ms_per_step), not the absolute times."""
import csv
import io
import re
import sys


def read_rows(path):
  with open(path, newline="") as f:
    text = f.read()
  # ncu prefixes the CSV with "==PROF==" lines and whatever the application printed
  start = text.find('"ID"')
  if start < 0:
    raise SystemExit("no ncu CSV header in %s" % path)
  return list(csv.DictReader(io.StringIO(text[start:])))


def short(name):
  name = re.sub(r"\(.*$", "", name)  # drop the argument list
  return name[:96]


def main():
  if len(sys.argv) < 2:
    raise SystemExit(__doc__)
  rows = read_rows(sys.argv[1])
  title = sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"
  cmd = sys.argv[3] if len(sys.argv) > 3 else ""
  agg = {}
  for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
      continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
  total = sum(a[1] for a in agg.values())
  print("# %s\n" % title)
  if cmd:
    print("Command: `%s`\n" % cmd)
  print("%d launches, %.3f ms of GPU time in total (cold-cache, serialised: compare shares, not absolutes).\n" %
        (sum(a[0] for a in agg.values()), total))
  print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
  for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ms, 100.0 * ms / total if total else 0.0))
  ours = sum(ms for k, (n, ms) in agg.items() if "iic::" in k)
  print("\nKernels of this library (`iic::`): %.1f%% of the GPU time; the rest is torch plumbing (copies, fills, "
        "gradient accumulation adds)." % (100.0 * ours / total if total else 0.0))


if __name__ == "__main__":
  main()
