"""Achieved HBM bandwidth of the BatchNorm / elementwise passes at ClusterNet5g shapes (1408 images)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iic_b200 import kernels as K

N = int(sys.argv[1]) if len(sys.argv) > 1 else 704  # images per view (BN statistics are per view)
SHAPES = [("stem 96x96x64", 96, 64), ("l1 49x49x64", 49, 64), ("l2 25x25x128", 25, 128), ("l3 13x13x256", 13, 256), ("l4 7x7x512", 7, 512)]


def timeit(fn, reps=5):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


for name, h, C in SHAPES:
  y = torch.randn(N, h, h, C, device="cuda").bfloat16()
  g = torch.randn(N, h, h, C, device="cuda").bfloat16()
  act = torch.relu(torch.randn(N, h, h, C, device="cuda")).bfloat16()
  gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
  dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
  ss, mi = K.bn_stats(y, gamma, beta, 1e-5, 0.1, None, None, False)
  nb = y.numel() * 2
  dy, go, out = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
  t_stats = timeit(lambda: K.bn_stats(y, gamma, beta, 1e-5, 0.1, None, None, False))
  t_apply = timeit(lambda: K.bn_apply(y, ss, True, res=act, out=out))
  t_b1 = timeit(lambda: K.bn_bwd(g, None, y, mi, gamma, dg, db, False, False, dy=dy, mask_ss=ss))
  t_b2 = timeit(lambda: K.bn_bwd(g, act, y, mi, gamma, dg, db, False, True, dy=dy, g_out=go))
  mi2 = [mi, mi.clone()]
  ss2 = [ss, ss.clone()]
  y2, g2, act2 = torch.cat([y, y]), torch.cat([g, g]), torch.cat([act, act])
  t_f1 = timeit(lambda: K.bn_bwd_fused(g2, None, y2, mi2, gamma, dg, db, False, False, mask_sss=ss2)) / 2
  t_f2 = timeit(lambda: K.bn_bwd_fused(g2, act2, y2, mi2, gamma, dg, db, False, True)) / 2
  print("%-16s fused (2 views in one launch, per view): bwd(bn1) %.3f ms %5.0f GB/s | bwd(bn2+gout) %.3f ms %5.0f GB/s" % (
    name, t_f1, 5 * nb / t_f1 / 1e6, t_f2, 8 * nb / t_f2 / 1e6))
  del y2, g2, act2
  print("%-16s %6.1f MB/tensor | stats %.3f ms %5.0f GB/s | apply+res %.3f ms %5.0f GB/s | bwd(bn1) %.3f ms %5.0f GB/s | bwd(bn2+gout) %.3f ms %5.0f GB/s" % (
    name, nb / 1e6, t_stats, nb / t_stats / 1e6, t_apply, 3 * nb / t_apply / 1e6, t_b1, 5 * nb / t_b1 / 1e6, t_b2, 8 * nb / t_b2 / 1e6))
