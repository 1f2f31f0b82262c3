#!/usr/bin/env python
"""Proves that an edit left already-validated kernels untouched: compiles csrc/<file>.cu at a git revision (default HEAD)
and in the working tree, and compares the SASS of every kernel the old object contains, function by function.

    python tools/sass_diff.py elementwise.cu [rev]

Kernels that exist only in the new object (added variants) are listed; any difference in an old kernel is an error.
Used when a new, not-yet-validated variant (a template instantiation behind a switch) is added to a file whose existing
kernels have already passed the GPU suite."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "iic_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def functions(obj):
  out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
  d, cur = {}, None
  for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
      cur = m.group(1)
      d[cur] = []
      continue
    if cur is not None:
      l = re.sub(r"/\*[0-9a-f]{4}\*/", "", line)           # instruction address
      l = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip()    # encoding (contains line-table free bits)
      if l:
        d[cur].append(l)
  return d


def main():
  src = sys.argv[1]
  rev = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
  with tempfile.TemporaryDirectory() as tmp:
    for tag in ("old", "new"):
      d = os.path.join(tmp, tag, "iic_b200", "csrc")
      os.makedirs(d)
      os.makedirs(os.path.join(tmp, tag, "include"))
      for f in os.listdir(CSRC):
        if f.endswith(".cuh") or f == src:
          if tag == "old":
            text = subprocess.run(["git", "-C", ROOT, "show", "%s:iic_b200/csrc/%s" % (rev, f)], capture_output=True, text=True).stdout
          else:
            text = open(os.path.join(CSRC, f)).read()
          open(os.path.join(d, f), "w").write(text)
      hdr = (subprocess.run(["git", "-C", ROOT, "show", "%s:include/iic_b200.h" % rev], capture_output=True, text=True).stdout
             if tag == "old" else open(os.path.join(ROOT, "include", "iic_b200.h")).read())
      open(os.path.join(tmp, tag, "include", "iic_b200.h"), "w").write(hdr)
      subprocess.run(["/usr/local/cuda/bin/nvcc"] + FLAGS + ["-c", os.path.join(d, src), "-o", os.path.join(tmp, tag + ".o")], check=True)
    a, b = functions(os.path.join(tmp, "old.o")), functions(os.path.join(tmp, "new.o"))
  bad = 0
  for k in sorted(a):
    kb = k if k in b else None
    if kb is None:  # a template parameter was appended: accept the unique new function that starts with the same stem
      stem = re.sub(r"E+v.*$", "", k)
      cands = [x for x in b if x not in a and x.startswith(stem)]
      kb = next((x for x in cands if a[k] == b[x]), None)
    if kb is None or a[k] != b[kb]:
      bad += 1
      print("DIFFERENT:", k)
  new_only = [k for k in b if k not in a]
  print("%s @ %s: %d kernels compared, %d different, %d only in the working tree" % (src, rev, len(a), bad, len(new_only)))
  for k in new_only:
    print("  new:", k[:110])
  sys.exit(1 if bad else 0)


if __name__ == "__main__":
  main()
