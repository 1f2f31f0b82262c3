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


def rgb_sobel_process(imgs_rgb):
  """RGB batch (n,3,h,w), uint8 or fp32 in [0,1] -> (n,2,h,w): the grey conversion of the reference's dataloader
  (``custom_greyscale_to_tensor``, :12-16) and ``sobel_process(grey, include_rgb=False)`` in one kernel
  (SURVEY.md S8f row 2)."""
  assert imgs_rgb.dim() == 4 and imgs_rgb.size(1) == 3
  if not imgs_rgb.is_cuda:
    raise RuntimeError("iic_b200.rgb_sobel_process: CUDA tensors only (no CPU fallback)")
  x = imgs_rgb.detach()
  if x.dtype != torch.uint8:
    x = x.float()
  return kernels.grey_sobel(x.contiguous())
