"""Oracle (CPU, test infrastructure only) for ``sobel_process``.

Restates xu-ji/IIC ``code/utils/cluster/transforms.py:47-96``.  The reference
cannot run on CPU (hard ``.cuda()`` at :72,:78), so parity for this function is
"restated, unpinned by execution"; it is pinned instead by hand-computable
known answers (tests/test_oracle_golden.py: a horizontal ramp has dx == -2*slope*... etc).

Filters (cross-correlation, zero padding 1, no bias):
  dx: [[1,0,-1],[2,0,-2],[1,0,-1]]   (:69)
  dy: [[1,2,1],[0,0,0],[-1,-2,-1]]   (:75)
Channel conventions (:50-66, :84-94):
  not IR, no rgb : (n,1,h,w) grey            -> (n,2,h,w) [dx,dy]
  not IR, rgb    : (n,4,h,w) rgb+grey        -> (n,5,h,w) [rgb,dx,dy]
  IR, no rgb     : (n,2,h,w) grey+ir         -> (n,3,h,w) [dx,dy,ir]
  IR, rgb        : (n,5,h,w) rgb+grey+ir     -> (n,6,h,w) [rgb,dx,dy,ir]
"""
import torch
import torch.nn.functional as F

SOBEL_DX = [[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]]
SOBEL_DY = [[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]]


def sobel_process(imgs, include_rgb, using_IR=False):
  n, c, h, w = imgs.shape
  rgb = ir = None
  if not using_IR:
    if not include_rgb:
      assert c == 1
      grey = imgs
    else:
      assert c == 4
      rgb, grey = imgs[:, :3], imgs[:, 3:4]
  else:
    if not include_rgb:
      assert c == 2
      grey, ir = imgs[:, 0:1], imgs[:, 1:2]
    else:
      assert c == 5
      rgb, grey, ir = imgs[:, :3], imgs[:, 3:4], imgs[:, 4:5]
  wx = torch.tensor(SOBEL_DX, dtype=imgs.dtype).view(1, 1, 3, 3)
  wy = torch.tensor(SOBEL_DY, dtype=imgs.dtype).view(1, 1, 3, 3)
  dx = F.conv2d(grey, wx, padding=1)
  dy = F.conv2d(grey, wy, padding=1)
  parts = ([rgb] if rgb is not None else []) + [dx, dy] + ([ir] if ir is not None else [])
  return torch.cat(parts, dim=1).detach()


def grey_from_rgb(imgs):
  """``custom_greyscale_to_tensor(include_rgb=False)`` of the reference (code/utils/cluster/transforms.py:12-16) on a
  batch: uint8 RGB -> PIL "L" (ImagingConvert rgb2l: (19595 R + 38470 G + 7471 B + 0x8000) >> 16) -> to_tensor (/255);
  fp32 RGB in [0,1] -> the same weights without the uint8 rounding (0.299, 0.587, 0.114)."""
  assert imgs.dim() == 4 and imgs.shape[1] == 3
  if imgs.dtype == torch.uint8:
    v = imgs.to(torch.int64)
    l = (19595 * v[:, 0:1] + 38470 * v[:, 1:2] + 7471 * v[:, 2:3] + 0x8000) >> 16
    return l.float() / 255.0
  return 0.299 * imgs[:, 0:1] + 0.587 * imgs[:, 1:2] + 0.114 * imgs[:, 2:3]
