#!/usr/bin/env python
"""Where does a kernel stall?  Summarises the source page of an ncu report (captured with --set full --import-source on;
`ncu -i X.ncu-rep --page source --csv`): per captured launch the warp-stall sample totals by reason, and the SASS
instructions that collect the most samples (with their dominant reasons).

    python tools/ncu_source_hot.py gpurun_out/prof.ncu-rep [kernel-name-substring] [top=12]
"""
import csv
import io
import subprocess
import sys


def launches(rep):
  out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
  cur = None
  for row in csv.reader(io.StringIO(out)):
    if not row:
      continue
    if row[0] == "Kernel Name":
      cur = {"name": row[1], "hdr": None, "rows": []}
      yield cur
    elif cur is not None and cur["hdr"] is None:
      cur["hdr"] = row
    elif cur is not None:
      cur["rows"].append(row)


def main():
  rep = sys.argv[1]
  sub = sys.argv[2] if len(sys.argv) > 2 else ""
  top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
  for i, k in enumerate(list(launches(rep))):
    if sub not in k["name"] or k["hdr"] is None:
      continue
    h = k["hdr"]
    col = {n: j for j, n in enumerate(h)}
    reasons = [n for n in h if n.startswith("stall_") and "(Not Issued)" not in n]
    samp = col["Warp Stall Sampling (All Samples)"]

    def num(r, j):
      try:
        return float(r[j])
      except Exception:
        return 0.0

    total = sum(num(r, samp) for r in k["rows"])
    print("== launch %d: %s  (%d SASS instructions, %.0f stall samples)" % (i, k["name"][:70], len(k["rows"]), total))
    by = {n: sum(num(r, col[n]) for r in k["rows"]) for n in reasons}
    print("   by reason: " + ", ".join("%s %.1f%%" % (n[6:], 100 * v / max(total, 1)) for n, v in sorted(by.items(), key=lambda kv: -kv[1])[:8]))
    hot = sorted(k["rows"], key=lambda r: -num(r, samp))[:top]
    for r in hot:
      rs = sorted(((n[6:], num(r, col[n])) for n in reasons), key=lambda kv: -kv[1])[:2]
      print("   %5.1f%%  %-58s %s" % (100 * num(r, samp) / max(total, 1), r[col["Source"]].strip()[:58],
                                      ", ".join("%s %.0f" % x for x in rs if x[1] > 0)))


if __name__ == "__main__":
  main()
