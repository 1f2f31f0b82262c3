"""Drop-in for xu-ji/IIC ``code/utils/cluster/IID_losses.py`` (IID_loss :6-33,
compute_joint :36-47) backed by the fused sm_100a kernel (csrc/iid_loss.cu).

Same signatures, same return values: ``IID_loss(x_out, x_tf_out, lamb=1.0,
EPS=sys.float_info.epsilon) -> (loss, loss_no_lamb)``, both differentiable
w.r.t. the two softmax inputs.  One kernel launch computes the joint, the MI
and the analytic gradient; under ``torch.no_grad()`` / non-grad inputs
(cluster_eval.py:281-288) the gradient sweep is skipped.

With ``iic_b200.distributed.enable()`` the batch is the concatenation of every
rank's rows: the un-normalised [k,k] joint is all-reduced between the kernel's
PARTIAL and FINISH phases, so every rank gets the global loss and the gradient
for its own rows.
"""
import sys

import torch

from ... import _lib, distributed, kernels


class _IIDLoss(torch.autograd.Function):
  @staticmethod
  def forward(ctx, z, zt, lamb, eps, want_grad):
    # z, zt: [S, n, k] fp32 contiguous
    if distributed.active():
      S, n, k = z.shape
      joint = torch.empty((S, k, k), device=z.device, dtype=torch.float32)
      kernels.iid_loss(z, zt, lamb, eps, False, phase=_lib.PHASE_PARTIAL, joint_ws=joint)
      distributed.allreduce_sum_(joint)
      loss, dz, dzt, _ = kernels.iid_loss(z, zt, lamb, eps, want_grad, phase=_lib.PHASE_FINISH, joint_ws=joint)
      ctx.joint = joint if want_grad else None
    else:
      loss, dz, dzt, _ = kernels.iid_loss(z, zt, lamb, eps, want_grad)
      ctx.joint = None
    ctx.set_materialize_grads(False)
    ctx.lamb, ctx.eps = lamb, eps
    if want_grad:
      ctx.save_for_backward(z, zt, dz, dzt)
    return loss[:, 0], loss[:, 1]

  @staticmethod
  def backward(ctx, g_loss, g_nolamb):
    z, zt, dz, dzt = ctx.saved_tensors
    gz = gzt = None
    if g_loss is not None:
      s = g_loss.reshape(-1, 1, 1)
      gz, gzt = dz * s, dzt * s
    if g_nolamb is not None:
      # rarely used (the reference only logs loss_no_lamb): re-evaluate the gradient with lamb = 1
      if ctx.joint is not None:
        _, dz1, dzt1, _ = kernels.iid_loss(z, zt, 1.0, ctx.eps, True, phase=_lib.PHASE_FINISH, joint_ws=ctx.joint)
      else:
        _, dz1, dzt1, _ = kernels.iid_loss(z, zt, 1.0, ctx.eps, True)
      s1 = g_nolamb.reshape(-1, 1, 1)
      gz = dz1 * s1 if gz is None else gz + dz1 * s1
      gzt = dzt1 * s1 if gzt is None else gzt + dzt1 * s1
    return gz, gzt, None, None, None


def _want_grad(a, b):
  # the gradient sweep is skipped for no-grad callers (cluster_eval.py:281-288)
  return torch.is_grad_enabled() and (a.requires_grad or b.requires_grad)


def _prep(x):
  assert x.dim() == 2, "IID_loss expects (bn, k) softmax outputs"
  if not x.is_cuda:
    raise RuntimeError("iic_b200.IID_loss: CUDA tensors only (the CPU implementation lives in oracle/ as a test "
                       "checker and is not a fallback)")
  return x.float().contiguous().unsqueeze(0)


def IID_loss(x_out, x_tf_out, lamb=1.0, EPS=sys.float_info.epsilon):
  # has had softmax applied
  bn, k = x_out.size()
  assert (x_tf_out.size(0) == bn and x_tf_out.size(1) == k)
  loss, loss_no_lamb = _IIDLoss.apply(_prep(x_out), _prep(x_tf_out), float(lamb), float(EPS),
                                      _want_grad(x_out, x_tf_out))
  return loss[0], loss_no_lamb[0]


def IID_loss_subheads(x_outs, x_tf_outs, lamb=1.0, EPS=sys.float_info.epsilon):
  """All S sub-heads in ONE launch.  x_outs / x_tf_outs: [S, bn, k] tensors (what the
  iic_b200 nets return before unbinding) or lists of S (bn, k) tensors.
  Returns (loss[S], loss_no_lamb[S]); ``loss.mean()`` is the reference's
  avg_loss_batch (cluster_sobel_twohead.py:323-336)."""
  if isinstance(x_outs, (list, tuple)):
    x_outs = torch.stack(list(x_outs))
  if isinstance(x_tf_outs, (list, tuple)):
    x_tf_outs = torch.stack(list(x_tf_outs))
  assert x_outs.dim() == 3 and x_outs.shape == x_tf_outs.shape
  return _IIDLoss.apply(x_outs.float().contiguous(), x_tf_outs.float().contiguous(), float(lamb), float(EPS),
                        _want_grad(x_outs, x_tf_outs))


def compute_joint(x_out, x_tf_out):
  """(k, k) symmetrised, normalised joint (reference :36-47).  Forward only: gradients flow
  through IID_loss (which fuses this computation), not through this helper."""
  bn, k = x_out.size()
  assert (x_tf_out.size(0) == bn and x_tf_out.size(1) == k)
  z, zt = _prep(x_out.detach()), _prep(x_tf_out.detach())
  if distributed.active():
    joint = torch.empty((1, k, k), device=z.device, dtype=torch.float32)
    kernels.iid_loss(z, zt, 1.0, sys.float_info.epsilon, False, phase=_lib.PHASE_PARTIAL, joint_ws=joint)
    distributed.allreduce_sum_(joint)
    _, _, _, p = kernels.iid_loss(z, zt, 1.0, sys.float_info.epsilon, False, phase=_lib.PHASE_FINISH, joint_ws=joint,
                                  want_joint=True)
  else:
    _, _, _, p = kernels.iid_loss(z, zt, 1.0, sys.float_info.epsilon, False, want_joint=True)
  return p[0]
