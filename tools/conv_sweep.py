"""Per-layer timing of the tcgen05 conv kernel on the distinct geometries of ClusterNet5g @ 96x96.
usage: python tools/conv_sweep.py [n_images] [bf16|tf32|tf32x3]   (default 1408 = 704 pairs x 2 views, bf16)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from iic_b200 import kernels as K
from iic_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
MODE = sys.argv[2] if len(sys.argv) > 2 else "bf16"
BF16 = {"bf16": _lib.BF16, "tf32": _lib.TF32, "tf32x3": _lib.TF32X3}[MODE]  # compute mode handed to the conv entry points
STORE = _lib.BF16 if MODE == "bf16" else _lib.F32
TDT = torch.bfloat16 if MODE == "bf16" else torch.float32
LAYERS = [  # name, h, cin, cout, k, stride, pad, count in the net
  ("l1 3x3 64->64 @49", 49, 64, 64, 3, 1, 1, 6),
  ("l2 3x3s2 64->128 @49", 49, 64, 128, 3, 2, 1, 1),
  ("l2 1x1s2 64->128 @49", 49, 64, 128, 1, 2, 0, 1),
  ("l2 3x3 128->128 @25", 25, 128, 128, 3, 1, 1, 7),
  ("l3 3x3s2 128->256 @25", 25, 128, 256, 3, 2, 1, 1),
  ("l3 1x1s2 128->256 @25", 25, 128, 256, 1, 2, 0, 1),
  ("l3 3x3 256->256 @13", 13, 256, 256, 3, 1, 1, 11),
  ("l4 3x3s2 256->512 @13", 13, 256, 512, 3, 2, 1, 1),
  ("l4 1x1s2 256->512 @13", 13, 256, 512, 1, 2, 0, 1),
  ("l4 3x3 512->512 @7", 7, 512, 512, 3, 1, 1, 5),
]


def timeit(fn, reps=5):
  fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


rows, tot = [], {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
flops_tot = 0.0
for name, h, cin, cout, k, s, p, cnt in LAYERS:
  g = K.conv_geom(N, h, h, cin, cout, k, k, s, p, 1)
  x = torch.randn(N, h, h, cin, device="cuda").to(TDT)
  w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
  dy = torch.randn(N, g.oh, g.ow, cout, device="cuda").to(TDT)
  wp, wt = K.pack_weight(w, K.weight_dtype(STORE, BF16), 0), K.pack_weight(w, K.weight_dtype(STORE, BF16), 1)
  gw = torch.zeros_like(w)
  fl = 2.0 * N * g.oh * g.ow * cout * k * k * cin
  t = {"fprop": timeit(lambda: K.conv_fprop(x, wp, g, BF16)),
       "dgrad": timeit(lambda: K.conv_dgrad(dy, wt, g, BF16)),
       "wgrad": timeit(lambda: K.conv_wgrad(x, dy, g, BF16, gw, False))}
  rows.append((name, cnt, fl, t))
  for kk in t:
    tot[kk] += t[kk] * cnt
  flops_tot += fl * cnt
  print("%-26s x%-2d %7.1f GFLOP  fprop %7.3f ms %6.0f TF | dgrad %7.3f ms %6.0f TF | wgrad %7.3f ms %6.0f TF" % (
    name, cnt, fl / 1e9, t["fprop"], fl / t["fprop"] / 1e9, t["dgrad"], fl / t["dgrad"] / 1e9, t["wgrad"], fl / t["wgrad"] / 1e9))
  del x, dy
print("mode", MODE)
print("TOTAL per view-batch of %d images: fprop %.1f ms, dgrad %.1f ms, wgrad %.1f ms; %.1f TFLOP each => %.0f / %.0f / %.0f TFLOP/s" % (
  N, tot["fprop"], tot["dgrad"], tot["wgrad"], flops_tot / 1e12, flops_tot / tot["fprop"] / 1e9, flops_tot / tot["dgrad"] / 1e9,
  flops_tot / tot["wgrad"] / 1e9))
