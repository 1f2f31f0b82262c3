"""Drop-in for xu-ji/IIC ``code/utils/segmentation/IID_losses.py``:
``IID_segmentation_loss`` (:14-83) and ``IID_segmentation_loss_uncollapsed`` (:86-159), same
signatures and return values, backed by csrc/seg_loss.cu + the MI kernel of csrc/iid_loss.cu.

  x2 is brought into x1's frame (affine resample), both are masked, the joint over the
  (2T+1)^2 displacement window is accumulated, then
    uncollapsed: one MI per displacement (each normalised, attached), averaged;
    collapsed:   displacements summed first (via the box-filter identity),
                 normaliser detached (:60), one MI.
  Gradients w.r.t. both inputs are computed analytically in the same call.

With ``iic_b200.distributed.enable()`` the un-normalised joint ([(2T+1)^2, k, k] or [k, k]) is SUM
all-reduced before the MI, so every rank sees the global loss and gets the gradient of its own images."""
from sys import float_info

import torch

from ... import distributed, kernels

EPS = float_info.epsilon
RENDER = False


def _seg_forward(x1, x2, theta, mask, lamb, T, collapsed, want_grad, shift):
  """-> (out [2] = (loss, loss_no_lamb), dx1, dx2) with dx* = d loss / d x* (None unless want_grad)."""
  n, k, h, w = x1.shape
  x1m, x2m = kernels.seg_prepare(x1, x2, theta, mask, shift)
  if collapsed:
    b1 = kernels.box_filter(x1m, k, T)
    joint = kernels.seg_joint(b1, x2m, k, 0)  # [1, k, k]
  else:
    joint = kernels.seg_joint(x1m, x2m, k, T)  # [(2T+1)^2, k, k]
  if distributed.active():
    distributed.allreduce_sum_(joint)
  loss, H = kernels.joint_mi(joint, lamb, EPS, collapsed, want_grad)
  V2 = 1 if collapsed else (2 * T + 1) ** 2
  out = loss.sum(dim=0) / float(V2)  # (:150-157): mean over displacements
  dx1 = dx2 = None
  if want_grad:
    if collapsed:
      d_b1 = kernels.seg_corr_bwd(x2m, H, k, 0, 1, 1.0)
      dx1m = kernels.box_filter(d_b1, k, T)  # the zero-padded box filter is self-adjoint
      dx2m = kernels.seg_corr_bwd(b1, H, k, 0, -1, 1.0)
    else:
      dx1m = kernels.seg_corr_bwd(x2m, H, k, T, 1, 1.0 / V2)
      dx2m = kernels.seg_corr_bwd(x1m, H, k, T, -1, 1.0 / V2)
    dx1, dx2 = kernels.seg_unprepare(dx1m, dx2m, theta, mask, k, shift)
  return out, dx1, dx2


class _SegLoss(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x1, x2, theta, mask, lamb, T, collapsed, want_grad, shift):
    out, dx1, dx2 = _seg_forward(x1, x2, theta, mask, lamb, T, collapsed, want_grad, shift)
    if want_grad:
      ctx.save_for_backward(dx1, dx2, x1, x2, theta, mask)
      ctx.args = (T, collapsed, shift)
    ctx.set_materialize_grads(False)
    return out[0], out[1]

  @staticmethod
  def backward(ctx, g_loss, g_nolamb):
    dx1, dx2, x1, x2, theta, mask = ctx.saved_tensors
    g1 = g2 = None
    if g_loss is not None:
      g1, g2 = dx1 * g_loss, dx2 * g_loss
    if g_nolamb is not None:
      # rarely used (the reference only logs loss_no_lamb, :78-81): re-evaluate the analytic gradient with lamb = 1
      T, collapsed, shift = ctx.args
      _, e1, e2 = _seg_forward(x1, x2, theta, mask, 1.0, T, collapsed, True, shift)
      g1 = e1 * g_nolamb if g1 is None else g1 + e1 * g_nolamb
      g2 = e2 * g_nolamb if g2 is None else g2 + e2 * g_nolamb
    return g1, g2, None, None, None, None, None, None, None


def random_translation_multiple_draw(half_side_min, half_side_max):
  """The random part of the reference's ``random_translation_multiple`` (code/utils/segmentation/transforms.py:146-166,
  called at IID_losses.py:29-32 / :101-104): one (x, y) displacement for the whole batch, drawn from numpy's GLOBAL
  generator with the reference's own calls in the reference's order (so ``np.random.seed`` reproduces its runs).  The
  shift itself -- x2_inv read at (x + tx, y + ty), zero outside the frame, which is what the pad-and-crop at :150-163
  does -- is applied inside the resampling kernel (iic_seg_prepare_shift)."""
  import numpy as np
  t = np.random.randint(half_side_min, half_side_max + 1, size=(2,))
  polarities = np.random.choice([-1, 1], size=(2,), replace=True)
  t *= polarities
  return int(t[0]), int(t[1])


def _seg_loss(collapsed, x1_outs, x2_outs, all_affine2_to_1, all_mask_img1, lamb, half_T_side_dense,
              half_T_side_sparse_min, half_T_side_sparse_max):
  assert (x1_outs.requires_grad)
  assert (x2_outs.requires_grad)
  assert (not all_affine2_to_1.requires_grad)
  assert (not all_mask_img1.requires_grad)
  assert (x1_outs.shape == x2_outs.shape)
  shift = (0, 0)
  if (half_T_side_sparse_min != 0) or (half_T_side_sparse_max != 0):
    shift = random_translation_multiple_draw(half_T_side_sparse_min, half_T_side_sparse_max)
  if not x1_outs.is_cuda:
    raise RuntimeError("iic_b200 segmentation losses: CUDA tensors only (no CPU fallback)")
  bn, k, h, w = x1_outs.shape
  want_grad = torch.is_grad_enabled()
  return _SegLoss.apply(x1_outs.float().contiguous(), x2_outs.float().contiguous(),
                        all_affine2_to_1.float().contiguous(), all_mask_img1.reshape(bn, h, w).float().contiguous(),
                        float(lamb), int(half_T_side_dense), collapsed, want_grad, shift)


def IID_segmentation_loss(x1_outs, x2_outs, all_affine2_to_1=None, all_mask_img1=None, lamb=1.0,
                          half_T_side_dense=None, half_T_side_sparse_min=None, half_T_side_sparse_max=None):
  return _seg_loss(True, x1_outs, x2_outs, all_affine2_to_1, all_mask_img1, lamb, half_T_side_dense,
                   half_T_side_sparse_min, half_T_side_sparse_max)


def IID_segmentation_loss_uncollapsed(x1_outs, x2_outs, all_affine2_to_1=None, all_mask_img1=None, lamb=1.0,
                                      half_T_side_dense=None, half_T_side_sparse_min=None,
                                      half_T_side_sparse_max=None):
  return _seg_loss(False, x1_outs, x2_outs, all_affine2_to_1, all_mask_img1, lamb, half_T_side_dense,
                   half_T_side_sparse_min, half_T_side_sparse_max)
