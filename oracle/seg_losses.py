"""Oracle (CPU, test infrastructure only) for the segmentation IIC objective.

Restates xu-ji/IIC:
  * ``perform_affine_tf``  -> code/utils/segmentation/transforms.py:131-143
  * ``IID_segmentation_loss``             -> code/utils/segmentation/IID_losses.py:14-83
  * ``IID_segmentation_loss_uncollapsed`` -> code/utils/segmentation/IID_losses.py:86-159
and adds the displacement-explicit evaluation (``seg_joint_displacements``) and
the box-filter identity for the collapsed form, both used to
cross-check the CUDA kernels at sizes the conv formulation cannot reach.

Version drift: torch 0.4.1's ``affine_grid``/``grid_sample`` behaved like
``align_corners=True``; for square maps with translation-free affines (all the
reference generates) both conventions agree to ~4e-6.  The
oracle pins ``align_corners=True`` -- the reference's era semantics.

Sparse random displacement (``half_T_side_sparse_*``, unused by the published
commands: examples/commands.txt:74-103 set min=max=0) is restated as
``random_translation_multiple`` (code/utils/segmentation/transforms.py:146-166)
and pinned against the unmodified reference function under ``np.random.seed``.

Parity pinning: tests/golden/seg_loss_*.npz, generated from the reference.
"""
from sys import float_info

import torch
import torch.nn.functional as F

EPS = float_info.epsilon


def perform_affine_tf(data, tf_matrices, align_corners=True):
  n, k, h, w = data.shape
  assert tf_matrices.shape == (n, 2, 3)
  grid = F.affine_grid(tf_matrices, list(data.shape), align_corners=align_corners)
  return F.grid_sample(data, grid, mode="bilinear", padding_mode="zeros",
                       align_corners=align_corners)


def random_translation_multiple(data, half_side_min, half_side_max):
  """code/utils/segmentation/transforms.py:146-166: zero-pad by half_side_max, draw one (x, y) displacement with
  magnitude in [min, max] and random sign from numpy's global generator, crop back to (h, w)."""
  import numpy as np
  n, c, h, w = data.shape
  data = F.pad(data, (half_side_max, half_side_max, half_side_max, half_side_max), "constant", 0)
  t = np.random.randint(half_side_min, half_side_max + 1, size=(2,))
  polarities = np.random.choice([-1, 1], size=(2,), replace=True)
  t *= polarities
  t += half_side_max
  return data[:, :, t[1]:(t[1] + h), t[0]:(t[0] + w)]


def _masked_pair(x1, x2, affine2_to_1, mask, align_corners, sparse=(0, 0)):
  assert x1.shape == x2.shape
  n, k, h, w = x1.shape
  x2_inv = perform_affine_tf(x2, affine2_to_1, align_corners)  # reference :27 / :99
  if sparse[0] or sparse[1]:  # reference :29-32 / :101-104
    x2_inv = random_translation_multiple(x2_inv, sparse[0], sparse[1])
  m = mask.view(n, 1, h, w)
  return x1 * m, x2_inv * m  # reference :42-45 / :114-117


def seg_joint_displacements(x1m, x2m, T):
  """A[c, c', u, v] = sum_{n,y,x} x1m[n,c,y+u-T,x+v-T] * x2m[n,c',y,x].

  This is what ``F.conv2d(x1^T, weight=x2^T, padding=T)`` (reference :53 /
  :125) evaluates; index convention verified in SURVEY.md S8 a11.
  """
  a = x1m.permute(1, 0, 2, 3).contiguous()
  b = x2m.permute(1, 0, 2, 3).contiguous()
  return F.conv2d(a, weight=b, padding=(T, T))  # (k, k, 2T+1, 2T+1)


def _mi_terms(P, p_row, p_col, lamb):
  P = torch.where(P < EPS, torch.full_like(P, EPS), P)
  p_row = torch.where(p_row < EPS, torch.full_like(p_row, EPS), p_row)
  p_col = torch.where(p_col < EPS, torch.full_like(p_col, EPS), p_col)
  return -P * (torch.log(P) - lamb * torch.log(p_row) - lamb * torch.log(p_col))


def IID_segmentation_loss(x1_outs, x2_outs, all_affine2_to_1=None, all_mask_img1=None,
                          lamb=1.0, half_T_side_dense=None,
                          half_T_side_sparse_min=None, half_T_side_sparse_max=None,
                          align_corners=True):
  """Collapsed form; reference :14-83.  The normaliser is detached (:60)."""
  x1m, x2m = _masked_pair(x1_outs, x2_outs, all_affine2_to_1, all_mask_img1, align_corners,
                          (half_T_side_sparse_min or 0, half_T_side_sparse_max or 0))
  P = seg_joint_displacements(x1m, x2m, half_T_side_dense).sum(dim=(2, 3))  # :53-55
  P = P / float(P.detach().sum())  # :60-61 (python float => no gradient through the norm)
  P = (P + P.t()) / 2.  # :64
  p_row = P.sum(dim=1, keepdim=True)  # :67  (k,1)
  p_col = P.sum(dim=0, keepdim=True)  # :68  (1,k)
  loss = _mi_terms(P, p_row, p_col, lamb).sum()
  loss_no_lamb = _mi_terms(P, p_row, p_col, 1.0).sum()
  return loss, loss_no_lamb


def IID_segmentation_loss_uncollapsed(x1_outs, x2_outs, all_affine2_to_1=None,
                                      all_mask_img1=None, lamb=1.0, half_T_side_dense=None,
                                      half_T_side_sparse_min=None,
                                      half_T_side_sparse_max=None, align_corners=True):
  """One MI per displacement, averaged; reference :86-159."""
  k = x1_outs.shape[1]
  x1m, x2m = _masked_pair(x1_outs, x2_outs, all_affine2_to_1, all_mask_img1, align_corners,
                          (half_T_side_sparse_min or 0, half_T_side_sparse_max or 0))
  A = seg_joint_displacements(x1m, x2m, half_T_side_dense)
  side = 2 * half_T_side_dense + 1
  P = A.permute(2, 3, 0, 1)  # (T,T,k,k) :133
  P = P / P.sum(dim=(2, 3), keepdim=True)  # :134-135, attached to the graph
  P = (P + P.transpose(2, 3)) / 2.  # :138
  # reference :141-142 -- note the (swapped) naming there: "p_i_mat" is the sum
  # over dim 2; the MI expression is symmetric in the two marginals.
  p_a = P.sum(dim=2, keepdim=True).expand(-1, -1, k, -1)
  p_b = P.sum(dim=3, keepdim=True).expand(-1, -1, -1, k)
  loss = _mi_terms(P, p_a, p_b, lamb).sum() / (side * side)
  loss_no_lamb = _mi_terms(P, p_a, p_b, 1.0).sum() / (side * side)
  return loss, loss_no_lamb


def collapsed_joint_box_filter(x1m, x2m, T):
  """sum_{u,v} A[c,c',u,v] = sum_{n,p} Box_{(2T+1)^2}(x1m)[n,c,p] * x2m[n,c',p].

  Zero-padded box sum.  O(k) per pixel instead of O(k (2T+1)^2).
  """
  k = x1m.shape[1]
  side = 2 * T + 1
  box = F.conv2d(x1m, torch.ones(k, 1, side, side, dtype=x1m.dtype), padding=T, groups=k)
  return torch.einsum("nchw,ndhw->cd", box, x2m)
