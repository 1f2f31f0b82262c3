#!/usr/bin/env python
"""End-to-end precision of each mode of the CUDA path against the fp32 / fp64 CPU oracle on the WELL-CONDITIONED
fixture of tests/test_gpu_precision.py (informative clustering: centred head bias, head gain, two correlated views,
batch >= 32 pairs), plus a short Adam trajectory per mode.  Prints one JSON object; run on the GPU box:
    python tools/precision_probe.py [--sz 32] [--pairs 64] [--steps 40] [--modes bf16,fp32]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import precision_fixture as fx  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--sz", type=int, default=32)
  ap.add_argument("--pairs", type=int, default=64)
  ap.add_argument("--steps", type=int, default=40)
  ap.add_argument("--modes", default="bf16,fp32")
  ap.add_argument("--head", default="B")
  ap.add_argument("--fp64", action="store_true")
  a = ap.parse_args()
  torch.set_num_threads(len(os.sched_getaffinity(0)))
  f = fx.Fixture(a.sz, a.pairs, a.head)
  t0 = time.time()
  ref = f.oracle(torch.float32)
  out = {"fixture": f.describe(), "oracle_fp32_s": time.time() - t0, "oracle_loss": ref["loss"]}
  if a.fp64:
    r64 = f.oracle(torch.float64)
    out["oracle_fp32_vs_fp64"] = fx.compare(ref, r64)
    ref = r64
  for mode in a.modes.split(","):
    got = f.cuda(mode)
    out[mode] = fx.compare(got, ref)
  if a.steps > 0:
    traj = {m: f.cuda_trajectory(m, a.steps) for m in a.modes.split(",")}
    out["trajectory"] = {m: [round(v, 6) for v in t] for m, t in traj.items()}
    if "fp32" in traj:
      base = np.array(traj["fp32"])
      for m, t in traj.items():
        if m != "fp32":
          d = np.abs(np.array(t) - base) / np.abs(base)
          out["trajectory_rel_dev_vs_fp32_" + m] = {"max": float(d.max()), "mean": float(d.mean()), "last": float(d[-1])}
  print(json.dumps(out, indent=1))


if __name__ == "__main__":
  main()
