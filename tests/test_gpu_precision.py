"""End-to-end precision of every shipped mode on the WELL-CONDITIONED fixture (tests/precision_fixture.py), against
(a) the golden produced by the unmodified reference network (tests/golden/precision.npz, make_golden.py gen_precision)
and (b) the fp32 CPU oracle run live (full per-parameter gradients).  These are the stated tolerances of DESIGN.md S4.

Why the bf16 numbers are what they are: the forward error of bf16 storage through the 34-layer trunk is ~6 % of the
trunk feature (measured on the CPU with the oracle rounding at the product's storage points); a ReLU network's gradient
is discontinuous in its pre-activations, so the gradient error scales like sqrt(forward error) (sign flips), not
linearly -- fp32 itself is 3e-3 away from fp64 on this fixture for a 1e-6 forward error.  bf16 mode is therefore held to
(1) forward quantities, (2) agreement with the reference algorithm evaluated WITH bf16 storage (the oracle's
bf16_rounding emulation): it must not be further from fp32 than that emulation is, and (3) a training trajectory that
tracks the fp32 one.  The mode that meets an fp32 tolerance on the tensor cores is tf32x3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import precision_fixture as fx  # noqa: E402
from tests.conftest import load_golden  # noqa: E402

SZ, PAIRS, HEAD = 32, 64, "B"


@pytest.fixture(scope="module")
def setup():
  g = load_golden("precision.npz").sub("wc/%d_%d_%s" % (SZ, PAIRS, HEAD))
  f = fx.Fixture(SZ, PAIRS, HEAD, trunk_mean=torch.from_numpy(g["trunk_mean"]))
  return g, f, f.oracle(torch.float32)


def _vs_golden(got, g, out_atol, loss_rtol, norm_rtol, grad_rel):
  assert abs(got["loss"] - float(g["loss"])) <= loss_rtol * abs(float(g["loss"]))
  assert np.abs(got["out"].numpy() - g["ref_out"]).max() <= out_atol
  for pn, norm in zip(g["grad_names"], g["grad_norms"]):
    pn = str(pn)
    if pn not in got["grads"]:
      assert norm == 0.0, pn
      continue
    mine = float(got["grads"][pn].norm())
    assert abs(mine - norm) <= norm_rtol * norm + 1e-9, (pn, mine, norm)
  for key in g.sub("grad"):
    want = torch.from_numpy(g["grad/" + key]).double()
    rel = float((got["grads"][key] - want).norm() / want.norm())
    assert rel <= grad_rel, (key, rel)


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_fp32_grade_modes_meet_the_fp32_tolerance(mode, setup):
  """fp32 SIMT and 3xTF32 tensor-core convolutions: loss 1e-4 relative, outputs 5e-4, every parameter gradient within
  4e-2 relative L2 of the reference's (2.5e-2 in total; the fp32 reference itself is 3e-3..6e-3 from fp64 here),
  cosine >= 0.9995.  Measured on a B200: fp32 3e-6 / 5.5e-5 / 8.1e-3 / 6.3e-3 / 0.99997, tf32x3 1e-5 / 1.8e-4 / 1.8e-2 /
  1.55e-2 / 0.99984 (the tensor core's fp32 accumulation is not round-to-nearest)."""
  g, f, ref = setup
  got = f.cuda(mode)
  _vs_golden(got, g, out_atol=5e-4, loss_rtol=1e-4, norm_rtol=2e-2, grad_rel=4e-2)
  c = fx.compare(got, ref)
  assert c["loss_rel"] < 1e-4 and c["out_max_abs"] < 5e-4, c
  assert c["grad_rel_l2_total"] < 2.5e-2 and c["grad_rel_l2_max"] < 4e-2 and c["grad_cos_min"] > 0.9995, c


def test_tf32_mode_tolerance(setup):
  """Single-pass kind::tf32 (10-bit mantissa operands, fp32 storage): forward within 1 %, gradients within 25 %
  (measured: loss 2.5e-4, outputs 2.7e-2, gradients 0.148, minimum cosine 0.983)."""
  g, f, ref = setup
  c = fx.compare(f.cuda("tf32"), ref)
  assert c["loss_rel"] < 1e-2 and c["out_max_abs"] < 5e-2, c
  assert c["grad_rel_l2_total"] < 0.25 and c["grad_cos_median"] > 0.97, c


def test_bf16_mode_tolerance_and_agreement_with_bf16_storage_oracle(setup):
  g, f, ref = setup
  got = f.cuda("bf16")
  c = fx.compare(got, ref)
  emu = f.oracle(torch.float32, rounding=True)  # the reference algorithm with bf16 rounding at the product's storage points
  e = fx.compare(emu, ref)
  print("bf16 cuda vs fp32:", c)
  print("bf16 emu  vs fp32:", e)
  # (1) forward quantities
  assert c["loss_rel"] < 6e-2 and c["out_max_abs"] < 0.35, c
  assert abs(got["loss"] - float(g["loss"])) < 6e-2 * abs(float(g["loss"]))
  # (2) no further from fp32 than bf16 storage itself makes the reference algorithm
  assert c["grad_rel_l2_total"] < 1.3 * e["grad_rel_l2_total"] + 0.05, (c["grad_rel_l2_total"], e["grad_rel_l2_total"])
  assert c["out_max_abs"] < 1.5 * e["out_max_abs"] + 0.02
  assert c["grad_cos_median"] > min(0.80, e["grad_cos_median"] - 0.05), (c["grad_cos_median"], e["grad_cos_median"])
  # (3) still a descent direction for every parameter
  assert c["grad_cos_min"] > 0.5, c


def test_training_trajectories_track_fp32(setup):
  """40 Adam steps (lr 1e-4) on the fixture batch: every mode's loss curve stays within 3 % (bf16; measured 2.3 % at the
  first step -- its forward loss error -- and 0.2 % from step 3 on) / 0.5 % (tf32x3; measured 0.12 %) of the fp32-SIMT
  curve, and the loss goes down (measured -0.51 -> -2.27 in all modes)."""
  g, f, ref = setup
  base = np.array(f.cuda_trajectory("fp32", 40))
  assert abs(base[0] - float(g["loss"])) < 1e-4 * abs(float(g["loss"])) and base[-1] < base[0] - 0.05
  for mode, tol in (("tf32x3", 5e-3), ("bf16", 3e-2)):
    t = np.array(f.cuda_trajectory(mode, 40))
    dev = np.abs(t - base) / np.abs(base)
    assert dev.max() < tol, (mode, float(dev.max()), t[-1], base[-1])
