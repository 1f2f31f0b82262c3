"""Drop-in for the GPU part of xu-ji/IIC ``code/utils/segmentation/transforms.py``:
``perform_affine_tf`` (:131-143) = affine_grid + bilinear grid_sample with zeros padding.

torch 0.4.1 semantics (``align_corners=True``) are used; for the square maps and translation-free
affines the reference generates, this agrees with today's default to ~4e-6.
``random_translation_multiple`` (:146-166) is unused by every published command (min=max=0) and
not provided."""
import torch

from ... import kernels


class _AffineTf(torch.autograd.Function):
  @staticmethod
  def forward(ctx, data, theta):
    n, k, h, w = data.shape
    ones = torch.ones((n, h, w), device=data.device, dtype=torch.float32)
    _, x2m = kernels.seg_prepare(data, data, theta, ones)
    ctx.save_for_backward(theta, ones)
    ctx.k = k
    return x2m[..., :k].permute(0, 3, 1, 2).contiguous()

  @staticmethod
  def backward(ctx, g):
    theta, ones = ctx.saved_tensors
    n, k, h, w = g.shape
    kp = kernels.seg_kp(k)
    gm = torch.zeros((n, h, w, kp), device=g.device, dtype=torch.float32)
    gm[..., :k] = g.permute(0, 2, 3, 1)
    _, dx2 = kernels.seg_unprepare(torch.zeros_like(gm), gm, theta, ones, k)
    return dx2, None


def perform_affine_tf(data, tf_matrices):
  # expects 4D tensor, we preserve gradients if there are any
  n_i, k, h, w = data.shape
  n_i2, r, c = tf_matrices.shape
  assert (n_i == n_i2)
  assert (r == 2 and c == 3)
  if not data.is_cuda:
    raise RuntimeError("iic_b200.perform_affine_tf: CUDA tensors only (no CPU fallback)")
  return _AffineTf.apply(data.float().contiguous(), tf_matrices.detach().float().contiguous())
