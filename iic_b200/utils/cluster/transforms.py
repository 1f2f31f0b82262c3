"""Drop-in for the GPU part of xu-ji/IIC ``code/utils/cluster/transforms.py``:
``sobel_process`` (:47-96).  The reference builds two nn.Conv2d modules and
uploads their weights on every call; here it is one stencil kernel."""
import torch

from ... import kernels


def sobel_process(imgs, include_rgb, using_IR=False):
  bn, c, h, w = imgs.size()
  if not imgs.is_cuda:
    raise RuntimeError("iic_b200.sobel_process: CUDA tensors only (no CPU fallback)")
  # channel-count asserts of the reference (:52,:56,:60,:64) are enforced by the C-ABI (-> AssertionError)
  out = kernels.sobel(imgs.detach().float().contiguous(), include_rgb, using_IR)
  return out
