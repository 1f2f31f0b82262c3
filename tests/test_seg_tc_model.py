"""CPU: the work decomposition, operand addressing, ring-slot barrier protocol and partial / reduce layouts of the
tensor-core segmentation kernels (csrc/seg_joint_tc.cu), exercised through their Python models (tools/models/) against
the oracle.  (The UMMA descriptor semantics the kernels rely on are checked on hardware by tools/umma_sw64_probe.cu.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "models"))

from oracle import seg_losses as oseg  # noqa: E402


def test_seg_joint_tc_model_matches_oracle():
  import seg_joint_tc_model as M
  rng = np.random.default_rng(0)
  for (n, h, T) in [(2, 12, 3), (1, 14, 6)]:
    x1m, x2m = rng.random((n, h, h, 16)), rng.random((n, h, h, 16))
    got = M.sim_joint(x1m, x2m, T)
    A = oseg.seg_joint_displacements(torch.from_numpy(x1m).permute(0, 3, 1, 2), torch.from_numpy(x2m).permute(0, 3, 1, 2), T)
    want = A.permute(2, 3, 0, 1).reshape(-1, 16, 16).numpy()
    assert np.abs(got - want).max() < 1e-10 * np.abs(want).max()


def test_seg_corr_tc_model_matches_direct_sum():
  import seg_corr_tc_model as M
  rng = np.random.default_rng(1)
  for (n, h, T, sgn) in [(2, 12, 3, 1), (1, 9, 3, -1)]:
    V = 2 * T + 1
    inp = rng.random((n, h, h, 16))
    H = rng.standard_normal((V * V, 16, 16))
    H = (H + H.transpose(0, 2, 1)) / 2
    a, b = M.sim(inp, H, T, sgn, 0.37), M.ref(inp, H, T, sgn, 0.37)
    assert np.abs(a - b).max() < 1e-10 * np.abs(b).max()
