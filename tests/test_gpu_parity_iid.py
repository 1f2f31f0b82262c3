"""GPU parity: fused IID_loss kernel vs the oracle and the reference-generated goldens.

Tolerances (fp32 path; stated per SURVEY.md S8c): loss within 2e-6 abs + 2e-6 rel of the fp32
reference value; gradients within 1e-5 relative to the largest gradient entry of the fp64 oracle."""
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import iid_losses as oracle_iid  # noqa: E402
from oracle import weights  # noqa: E402

LOSS_ATOL, LOSS_RTOL, GRAD_RTOL, GRAD_ATOL = 2e-6, 2e-6, 1e-5, 2e-7


def _api():
  from iic_b200.utils.cluster import IID_losses
  return IID_losses


def _close_loss(a, b):
  return abs(a - b) <= LOSS_ATOL + LOSS_RTOL * abs(b)


def _check_case(z, zt, lamb, ref_loss=None, ref_loss1=None):
  api = _api()
  zc = torch.from_numpy(np.asarray(z)).cuda().requires_grad_(True)
  ztc = torch.from_numpy(np.asarray(zt)).cuda().requires_grad_(True)
  loss, loss1 = api.IID_loss(zc, ztc, lamb=lamb)
  cf = oracle_iid.iid_loss_closed_form(z, zt, lamb=lamb)
  assert _close_loss(loss.item(), cf["loss"]), (loss.item(), cf["loss"])
  assert _close_loss(loss1.item(), cf["loss_no_lamb"]), (loss1.item(), cf["loss_no_lamb"])
  if ref_loss is not None:
    assert _close_loss(loss.item(), float(ref_loss)) and _close_loss(loss1.item(), float(ref_loss1))
  loss.backward()
  for got, want in [(zc.grad, cf["dz"]), (ztc.grad, cf["dzt"])]:
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got.cpu().numpy().astype(np.float64) - want).max()
    assert err <= GRAD_RTOL * scale + GRAD_ATOL, (err, scale)
  return loss, loss1


def test_goldens(golden_iid):
  for name in golden_iid.names("iid/"):
    c = golden_iid.sub("iid/" + name)
    _check_case(c["z"], c["zt"], float(c["lamb"]), c["ref_loss"], c["ref_loss_no_lamb"])
    # gradient also against the stored fp64 autograd of the restated reference
    zc = torch.from_numpy(c["z"]).cuda().requires_grad_(True)
    ztc = torch.from_numpy(c["zt"]).cuda().requires_grad_(True)
    _api().IID_loss(zc, ztc, lamb=float(c["lamb"]))[0].backward()
    scale = np.abs(c["orc_dz_f64"]).max()
    assert np.abs(zc.grad.cpu().numpy() - c["orc_dz_f64"]).max() <= GRAD_RTOL * scale
    joint = _api().compute_joint(zc.detach(), ztc.detach())
    np.testing.assert_allclose(joint.cpu().numpy(), c["ref_joint"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("n,k,lamb", [(1, 2, 1.0), (7, 3, 1.5), (33, 10, 1.0), (704, 10, 1.0), (704, 70, 1.0),
                                      (1408, 70, 1.2), (4096, 10, 1.0), (2800, 140, 1.0), (1000, 24, 1.0),
                                      (660, 50, 1.0)])
def test_random_shapes(n, k, lamb):
  l = weights.normal("gpu.iid.l.%d.%d" % (n, k), (n, k))
  lt = 2.0 * l + 0.7 * weights.normal("gpu.iid.lt.%d.%d" % (n, k), (n, k))
  z, zt = torch.softmax(2.0 * l, 1).numpy(), torch.softmax(lt, 1).numpy()
  _check_case(z, zt, lamb)


def test_no_grad_and_loss_no_lamb_backward():
  api = _api()
  l = weights.normal("gpu.iid.ng", (200, 10))
  z = torch.softmax(l, 1).cuda()
  zt = torch.softmax(l + 0.5 * weights.normal("gpu.iid.ng2", (200, 10)), 1).cuda()
  with torch.no_grad():  # cluster_eval.py:281-288 calls the loss under no_grad
    a, b = api.IID_loss(z, zt, lamb=1.3)
  assert not a.requires_grad
  zr, ztr = z.clone().requires_grad_(True), zt.clone().requires_grad_(True)
  a2, b2 = api.IID_loss(zr, ztr, lamb=1.3)
  assert a2.item() == a.item() and b2.item() == b.item()
  b2.backward()  # gradient of loss_no_lamb == gradient of the loss at lamb = 1
  cf = oracle_iid.iid_loss_closed_form(z.cpu().numpy(), zt.cpu().numpy(), lamb=1.0)
  assert np.abs(zr.grad.cpu().numpy() - cf["dz"]).max() <= GRAD_RTOL * np.abs(cf["dz"]).max()
  with pytest.raises(AssertionError):
    api.IID_loss(z, zt[:, :5])
  with pytest.raises(RuntimeError):
    api.IID_loss(z.cpu(), zt.cpu())


def test_subheads_one_launch_and_upstream_scale():
  api = _api()
  from iic_b200 import kernels
  S, n, k = 5, 704, 10
  l = weights.normal("gpu.iid.sh", (S, n, k))
  z = torch.softmax(2 * l, 2).cuda().requires_grad_(True)
  zt = torch.softmax(2 * l + weights.normal("gpu.iid.sh2", (S, n, k)), 2).cuda().requires_grad_(True)
  kernels.launch_count(reset=True)
  losses, losses1 = api.IID_loss_subheads(z, zt, lamb=1.0)
  assert kernels.launch_count() == 1
  (losses.mean() * 3.0).backward()
  for s in range(S):
    cf = oracle_iid.iid_loss_closed_form(z[s].detach().cpu().numpy(), zt[s].detach().cpu().numpy())
    assert _close_loss(losses[s].item(), cf["loss"])
    want = cf["dz"] * 3.0 / S
    assert np.abs(z.grad[s].cpu().numpy() - want).max() <= GRAD_RTOL * np.abs(want).max()


def test_phases_equal_fused_and_shard():
  """PARTIAL -> (sum of per-shard joints) -> FINISH reproduces the fused single-device result:
  the multi-GPU algorithm of SURVEY.md S8e, emulated on one device."""
  from iic_b200 import _lib, kernels
  S, n, k = 2, 704, 70
  l = weights.normal("gpu.iid.ph", (S, n, k))
  z = torch.softmax(2 * l, 2).cuda()
  zt = torch.softmax(2 * l + weights.normal("gpu.iid.ph2", (S, n, k)), 2).cuda()
  loss_f, dz_f, dzt_f, _ = kernels.iid_loss(z, zt, 1.0, sys.float_info.epsilon, True)
  joint = torch.zeros(S, k, k, device="cuda")
  shards = [(0, 352), (352, 704)]
  for lo, hi in shards:
    j = torch.empty(S, k, k, device="cuda")
    kernels.iid_loss(z[:, lo:hi].contiguous(), zt[:, lo:hi].contiguous(), 1.0, sys.float_info.epsilon, False,
                     phase=_lib.PHASE_PARTIAL, joint_ws=j)
    joint += j
  for lo, hi in shards:
    loss_s, dz_s, dzt_s, _ = kernels.iid_loss(z[:, lo:hi].contiguous(), zt[:, lo:hi].contiguous(), 1.0,
                                               sys.float_info.epsilon, True, phase=_lib.PHASE_FINISH, joint_ws=joint)
    assert torch.allclose(loss_s, loss_f, rtol=1e-5, atol=1e-6)
    assert torch.allclose(dz_s, dz_f[:, lo:hi], rtol=1e-4, atol=1e-8)
    assert torch.allclose(dzt_s, dzt_f[:, lo:hi], rtol=1e-4, atol=1e-8)


def test_full_size_properties():
  """BASELINE sizes (n=704*8, k=70, S=5): size-independent properties instead of an oracle run --
  symmetry in the arguments, loss == loss_no_lamb iff lamb == 1, one-hot identical views -> -log k."""
  api = _api()
  S, n, k = 5, 5632, 70
  l = weights.normal("gpu.iid.full", (S, n, k))
  z = torch.softmax(3 * l, 2).cuda()
  zt = torch.softmax(3 * l + weights.normal("gpu.iid.full2", (S, n, k)), 2).cuda()
  a, a1 = api.IID_loss_subheads(z, zt, lamb=1.0)
  b, _ = api.IID_loss_subheads(zt, z, lamb=1.0)
  assert torch.equal(a, a1)
  assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
  c, c1 = api.IID_loss_subheads(z, zt, lamb=1.5)
  assert torch.allclose(c1, a, rtol=1e-6, atol=1e-7) and (c < a).all()
  oh = torch.eye(k, device="cuda").repeat(n // k + 1, 1)[:n - n % k]
  l1, _ = api.IID_loss(oh, oh.clone())
  assert abs(l1.item() + np.log(k)) < 1e-5
