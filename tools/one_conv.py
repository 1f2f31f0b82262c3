"""Run one conv geometry (for ncu): python tools/one_conv.py <kind> <h> <cin> <cout> <k> <stride> <pad> [n=1408] [reps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iic_b200 import kernels as K
from iic_b200._lib import BF16
kind, h, cin, cout, k, s, p = sys.argv[1], *[int(a) for a in sys.argv[2:8]]
N = int(sys.argv[8]) if len(sys.argv) > 8 else 1408
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 3
g = K.conv_geom(N, h, h, cin, cout, k, k, s, p, 1)
x = torch.randn(N, h, h, cin, device="cuda").bfloat16()
w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
dy = torch.randn(N, g.oh, g.ow, cout, device="cuda").bfloat16()
wp, wt = K.pack_weight(w, BF16, 0), K.pack_weight(w, BF16, 1)
gw = torch.zeros_like(w)
for _ in range(reps):
  if kind == "fprop": K.conv_fprop(x, wp, g, BF16)
  elif kind == "dgrad": K.conv_dgrad(dy, wt, g, BF16)
  else: K.conv_wgrad(x, dy, g, BF16, gw, False)
torch.cuda.synchronize()
