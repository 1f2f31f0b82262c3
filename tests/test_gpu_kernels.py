"""GPU unit tests of the individual sm_100a kernels against plain PyTorch fp32 ops.

(Parity with the reference proper is in test_gpu_parity_*.py, against the oracle and the
golden vectors; these tests localise a failure to one kernel.)"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import transforms as otf  # noqa: E402
from oracle import weights  # noqa: E402


def _K():
  from iic_b200 import kernels
  return kernels


def _dev():
  return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _no_tf32():
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  yield


def to_nhwc(x, dt):
  return x.permute(0, 2, 3, 1).contiguous().to(dt)


def from_nhwc(x):
  return x.float().permute(0, 3, 1, 2).contiguous()


CONV_CASES = [
  # n, h, cin, cout, k, stride, pad, dil
  (3, 13, 64, 64, 3, 1, 1, 1),
  (2, 49, 64, 64, 3, 1, 1, 1),
  (3, 25, 64, 128, 3, 2, 1, 1),
  (3, 25, 64, 128, 1, 2, 0, 1),
  (5, 13, 128, 256, 3, 2, 1, 1),
  (4, 7, 256, 512, 3, 1, 1, 1),
  (9, 7, 512, 512, 3, 1, 1, 1),
  (2, 12, 64, 128, 5, 1, 2, 1),
  (2, 16, 256, 512, 3, 1, 1, 2),
  (1, 5, 128, 128, 3, 1, 1, 1),
  (32, 49, 64, 64, 3, 1, 1, 1),  # > 2 tiles per SM: resident-weights / two-tile variant of the TMA kernel
]


def _conv_inputs(case, seed=0):
  n, h, cin, cout, k, s, p, d = case
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(n, cin, h, h, generator=g)
  w = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))
  oh = (h + 2 * p - d * (k - 1) - 1) // s + 1
  dy = torch.randn(n, cout, oh, oh, generator=g)
  add = torch.randn(n, cin, h, h, generator=g)
  return x.cuda(), w.cuda(), dy.cuda(), add.cuda()


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_fprop_dgrad_wgrad(case, mode):
  K = _K()
  from iic_b200._lib import BF16, F32
  n, h, cin, cout, k, s, p, d = case
  dt, tdt = (F32, torch.float32) if mode == "fp32" else (BF16, torch.bfloat16)
  x, w, dy, add = _conv_inputs(case)
  if mode == "bf16":  # compare against the same rounded operands
    x, w, dy, add = [t.bfloat16().float() for t in (x, w, dy, add)]
  g = K.conv_geom(n, h, h, cin, cout, k, k, s, p, d)
  # exact reference: CPU fp64 (cuDNN may pick FFT / Winograd algorithms with ~1e-2 error)
  xr = x.double().cpu().requires_grad_(True)
  wr = w.double().cpu().requires_grad_(True)
  yr = F.conv2d(xr, wr, None, s, p, d)
  yr.backward(dy.double().cpu())
  yr = yr.detach().float().cuda()
  xr_grad, wr_grad = xr.grad.float().cuda(), wr.grad.float().cuda()
  tol = dict(rtol=2e-4, atol=2e-4) if mode == "fp32" else dict(rtol=2e-2, atol=2e-2)

  xh = to_nhwc(x, tdt)
  y = K.conv_fprop(xh, K.pack_weight(w, dt, 0), g, dt)
  torch.cuda.synchronize()
  err = (from_nhwc(y) - yr).abs().max().item()
  assert torch.allclose(from_nhwc(y), yr, **tol), "fprop max err %g" % err

  dyh = to_nhwc(dy, tdt)
  dx = K.conv_dgrad(dyh, K.pack_weight(w, dt, 1), g, dt)
  torch.cuda.synchronize()
  err = (from_nhwc(dx) - xr_grad).abs().max().item()
  assert torch.allclose(from_nhwc(dx), xr_grad, **tol), "dgrad max err %g" % err
  dx2 = K.conv_dgrad(dyh, K.pack_weight(w, dt, 1), g, dt, addend=to_nhwc(add, tdt))
  assert torch.allclose(from_nhwc(dx2), xr_grad + add, **tol), "dgrad+addend"

  gw = torch.zeros_like(w)
  K.conv_wgrad(xh, dyh, g, dt, gw, False)
  torch.cuda.synchronize()
  scale = wr_grad.abs().max().item()
  err = (gw - wr_grad).abs().max().item()
  wtol = 2e-4 if mode == "fp32" else 1e-2
  assert err <= wtol * scale + 1e-4, "wgrad max err %g (scale %g)" % (err, scale)
  K.conv_wgrad(xh, dyh, g, dt, gw, True)  # accumulate
  assert (gw - 2 * wr_grad).abs().max().item() <= 2 * wtol * scale + 2e-4


def test_tc_conv_exact_small_integers():
  """bf16 tensor-core path must be EXACT on small-integer operands (fp32 accumulation of
  exactly representable products): catches any swizzle / descriptor / gather indexing slip."""
  K = _K()
  from iic_b200._lib import BF16
  n, h, cin, cout, k = 2, 9, 128, 128, 3
  g = torch.Generator().manual_seed(3)
  x = torch.randint(-1, 2, (n, cin, h, h), generator=g).float()
  w = torch.randint(-1, 2, (cout, cin, k, k), generator=g).float()
  dy = torch.randint(-1, 2, (n, cout, h, h), generator=g).float()
  geo = K.conv_geom(n, h, h, cin, cout, k, k, 1, 1, 1)
  xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
  ref = F.conv2d(xr, wr, None, 1, 1)  # CPU fp64: exact integers
  ref.backward(dy.double())
  assert ref.abs().max() <= 256 and xr.grad.abs().max() <= 256  # exactly representable in bf16
  xh, dyh = to_nhwc(x.cuda(), torch.bfloat16), to_nhwc(dy.cuda(), torch.bfloat16)
  y = K.conv_fprop(xh, K.pack_weight(w.cuda(), BF16, 0), geo, BF16)
  assert torch.equal(from_nhwc(y).cpu(), ref.detach().float())
  dx = K.conv_dgrad(dyh, K.pack_weight(w.cuda(), BF16, 1), geo, BF16)
  assert torch.equal(from_nhwc(dx).cpu(), xr.grad.float())
  gw = torch.zeros_like(w).cuda()
  K.conv_wgrad(xh, dyh, geo, BF16, gw, False)
  assert torch.equal(gw.cpu(), wr.grad.float())  # fp32 output of exact integer sums


@pytest.mark.parametrize("case", [(6, 25, 128, 128, 3, 1, 1), (5, 13, 256, 256, 3, 1, 1), (3, 7, 512, 512, 3, 1, 1),
                                  (4, 13, 256, 512, 1, 2, 0), (3, 12, 128, 128, 5, 1, 2), (40, 9, 128, 128, 3, 1, 1),
                                  (3, 25, 64, 128, 3, 2, 1)])
def test_wgrad_multi_tile_work_items_exact(case):
  """Option wgrad_mt: 2 or 3 (tap, cin) tiles per dy k-block, single TMEM accumulator buffer.  Exact on small integers
  (fp32 sums of exactly representable products), with and without accumulation, for 3 x 128-row tiles (N = 128), 2 tiles
  (N = 256, one and two N tiles), a single 2-tile item (1x1), 5x5 taps, many split-K items, and a shape that stays on the
  one-tile path."""
  K = _K()
  from iic_b200._lib import BF16
  n, h, cin, cout, k, s, p = case
  g = torch.Generator().manual_seed(11)
  oh = (h + 2 * p - k) // s + 1
  x = torch.randint(-1, 2, (n, cin, h, h), generator=g).float()
  dy = torch.randint(-1, 2, (n, cout, oh, oh), generator=g).float()
  wr = torch.zeros(cout, cin, k, k, dtype=torch.float64, requires_grad=True)
  F.conv2d(x.double(), wr, None, s, p).backward(dy.double())
  geo = K.conv_geom(n, h, h, cin, cout, k, k, s, p, 1)
  xh, dyh = to_nhwc(x.cuda(), torch.bfloat16), to_nhwc(dy.cuda(), torch.bfloat16)
  base = torch.randint(-3, 4, (cout, cin, k, k), generator=g).float().cuda()
  for fused in (True, False):
    old = K.WGRAD_FUSED_UNPACK["on"]
    K.WGRAD_FUSED_UNPACK["on"] = fused
    try:
      with K.options(wgrad_mt=1):
        gw = torch.zeros_like(base)
        K.conv_wgrad(xh, dyh, geo, BF16, gw, False)
        acc = base.clone()
        K.conv_wgrad(xh, dyh, geo, BF16, acc, True)
    finally:
      K.WGRAD_FUSED_UNPACK["on"] = old
    assert torch.equal(gw.cpu(), wr.grad.float()), "max |err| %g" % (gw.cpu() - wr.grad.float()).abs().max().item()
    assert torch.equal((acc - base).cpu(), wr.grad.float())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(5, 13, 13, 64), (2, 25, 25, 128), (3, 7, 7, 512), (64, 3, 3, 256)])
def test_bn_forward_backward(mode, shape):
  K = _K()
  tdt = torch.float32 if mode == "fp32" else torch.bfloat16
  n, h, w, C = shape
  g = torch.Generator().manual_seed(1)
  y = (torch.randn(n, C, h, w, generator=g) * 2 + 0.5).cuda()
  res = torch.randn(n, C, h, w, generator=g).cuda()
  gamma = (torch.rand(C, generator=g) + 0.5).cuda()
  beta = (torch.randn(C, generator=g) * 0.1).cuda()
  dout = torch.randn(n, C, h, w, generator=g).cuda()
  if mode == "bf16":
    y, res, dout = [t.bfloat16().float() for t in (y, res, dout)]
  rm, rv = torch.zeros(C).cuda(), torch.ones(C).cuda()
  yh = to_nhwc(y, tdt)
  ss, mi = K.bn_stats(yh, gamma, beta, 1e-5, 0.1, rm, rv, False)
  # reference
  yr = y.clone().requires_grad_(True)
  gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
  rr = res.clone().requires_grad_(True)
  rm_r, rv_r = torch.zeros(C).cuda(), torch.ones(C).cuda()
  o = F.relu(F.batch_norm(yr, rm_r, rv_r, gr, br, True, 0.1, 1e-5) + rr)
  o.backward(dout)
  assert torch.allclose(rm, rm_r, atol=1e-5) and torch.allclose(rv, rv_r, rtol=1e-4, atol=1e-5)
  out = K.bn_apply(yh, ss, relu=True, res=to_nhwc(res, tdt))
  tol = dict(rtol=1e-4, atol=1e-4) if mode == "fp32" else dict(rtol=2e-2, atol=2e-2)
  assert torch.allclose(from_nhwc(out), o.detach(), **tol)
  dg, db = torch.empty(C).cuda(), torch.empty(C).cuda()
  dy, gres = K.bn_bwd(to_nhwc(dout, tdt), out, yh, mi, gamma, dg, db, False, True)
  if mode == "fp32":
    assert torch.allclose(from_nhwc(gres), rr.grad, **tol)
    assert torch.allclose(from_nhwc(dy), yr.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(dg, gr.grad, rtol=1e-3, atol=1e-3) and torch.allclose(db, br.grad, rtol=1e-3, atol=1e-3)
  else:
    # the ReLU mask is taken from the bf16-rounded output: compare away from the kink
    far = (o.detach().abs() > 0.05) | (o.detach() == 0)
    assert ((from_nhwc(dy) - yr.grad).abs()[far] < 0.05 * yr.grad.abs().max()).float().mean() > 0.995
    assert torch.allclose(dg, gr.grad, rtol=5e-2, atol=0.5) and torch.allclose(db, br.grad, rtol=5e-2, atol=0.5)
  # BN directly followed by ReLU: the mask is recomputed from y instead of reading the activation
  o3 = F.relu(F.batch_norm(yr2 := y.clone().requires_grad_(True), None, None, gamma, beta, True, 0.1, 1e-5))
  o3.backward(dout)
  dy3, _ = K.bn_bwd(to_nhwc(dout, tdt), None, yh, mi, gamma, dg, db, False, False, mask_ss=ss)
  if mode == "fp32":
    assert torch.allclose(from_nhwc(dy3), yr2.grad, rtol=1e-3, atol=1e-4)
  else:
    assert ((from_nhwc(dy3) - yr2.grad).abs() < 0.05 * yr2.grad.abs().max()).float().mean() > 0.99
  # eval mode uses the running statistics
  ss_e, _ = K.bn_stats(yh, gamma, beta, 1e-5, 0.1, rm, rv, True)
  oe = F.batch_norm(y, rm_r, rv_r, gamma, beta, False, 0.1, 1e-5)
  assert torch.allclose(from_nhwc(K.bn_apply(yh, ss_e, relu=False)), oe, **tol)
  # downsample-style residual: out = relu(bn(y) + bn_r(res))
  ss_r, _ = K.bn_stats(to_nhwc(res, tdt), gamma, beta, 1e-5, 0.1, None, None, False)
  o2 = F.relu(F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5) +
              F.batch_norm(res, None, None, gamma, beta, True, 0.1, 1e-5))
  out2 = K.bn_apply(yh, ss, relu=True, res=to_nhwc(res, tdt), rss=ss_r)
  assert torch.allclose(from_nhwc(out2), o2, **tol)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("hw,pad", [(96, 1), (24, 0), (13, 0), (32, 1)])
def test_bn_relu_maxpool(mode, hw, pad):
  K = _K()
  tdt = torch.float32 if mode == "fp32" else torch.bfloat16
  n, C = 2, 64
  g = torch.Generator().manual_seed(2)
  y = torch.randn(n, C, hw, hw, generator=g).cuda()
  if mode == "bf16":
    y = y.bfloat16().float()
  gamma = (torch.rand(C, generator=g) + 0.5).cuda()
  beta = (torch.randn(C, generator=g) * 0.1).cuda()
  yh = to_nhwc(y, tdt)
  ss, mi = K.bn_stats(yh, gamma, beta, 1e-5, 0.1, None, None, False)
  yr = y.clone().requires_grad_(True)
  bno = F.batch_norm(yr, None, None, gamma, beta, True, 0.1, 1e-5)
  bno.retain_grad()  # the kernel returns the gradient w.r.t. the BN output (pool routing x ReLU mask)
  a = F.relu(bno)
  pr = F.max_pool2d(a, 2, 2, pad)
  out = K.bn_relu_maxpool(yh, ss, pad)
  tol = dict(rtol=1e-4, atol=1e-4) if mode == "fp32" else dict(rtol=2e-2, atol=2e-2)
  assert out.shape[1] == pr.shape[2]
  assert torch.allclose(from_nhwc(out), pr.detach(), **tol)
  dp = torch.randn(pr.shape, generator=torch.Generator().manual_seed(5)).cuda()
  if mode == "bf16":
    dp = dp.bfloat16().float()
  pr.backward(dp)
  gmask = K.bn_relu_maxpool_bwd(yh, ss, to_nhwc(dp, tdt), pad)
  if mode == "fp32":
    assert torch.allclose(from_nhwc(gmask), bno.grad, rtol=1e-5, atol=1e-6)
  else:
    agree = (from_nhwc(gmask) - bno.grad).abs() < 1e-2
    assert agree.float().mean() > 0.995  # bf16 rounding can flip an arg-max between near-equal values


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_avgpool_and_layout(mode):
  K = _K()
  from iic_b200._lib import BF16, F32
  dt, tdt = (F32, torch.float32) if mode == "fp32" else (BF16, torch.bfloat16)
  x = torch.randn(5, 512, 7, 7, generator=torch.Generator().manual_seed(0)).cuda()
  xh = K.nchw_to_nhwc(x, dt)
  assert torch.equal(xh, to_nhwc(x, tdt))
  assert torch.equal(K.nhwc_to_nchw(xh), xh.float().permute(0, 3, 1, 2))
  f = K.avgpool(xh)
  assert torch.allclose(f, xh.float().mean(dim=(1, 2)), rtol=1e-5, atol=1e-6)
  df = torch.randn(5, 512).cuda()
  dx = K.avgpool_bwd(df, tuple(xh.shape), dt)
  assert torch.allclose(dx.float(), (df / 49.)[:, None, None, :].expand(5, 7, 7, 512), rtol=1e-2, atol=1e-6)
  assert torch.equal(K.cast(x, dt), x.to(tdt))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,k,pad,hw", [(2, 3, 1, 32), (1, 5, 2, 24), (5, 3, 1, 20)])
def test_stem(mode, cin, k, pad, hw):
  K = _K()
  from iic_b200._lib import BF16, F32
  dt, tdt = (F32, torch.float32) if mode == "fp32" else (BF16, torch.bfloat16)
  n, cout = 3, 64
  g = torch.Generator().manual_seed(4)
  x = torch.randn(n, cin, hw, hw, generator=g).cuda()
  w = (torch.randn(cout, cin, k, k, generator=g) * 0.3).cuda()
  dy = torch.randn(n, cout, hw, hw, generator=g).cuda()
  if mode == "bf16":
    dy = dy.bfloat16().float()
  geo = K.conv_geom(n, hw, hw, cin, cout, k, k, 1, pad, 1)
  y = K.stem_fprop(x, w, geo, dt)
  wr = w.clone().requires_grad_(True)
  yr = F.conv2d(x, wr, None, 1, pad)
  tol = dict(rtol=1e-4, atol=1e-4) if mode == "fp32" else dict(rtol=2e-2, atol=2e-2)
  assert torch.allclose(from_nhwc(y), yr.detach(), **tol)
  yr.backward(dy)
  gw = torch.zeros_like(w)
  # bf16 mode: the tcgen05 stem wgrad (kernels.STEM_WGRAD_TC) rounds the input patches to bf16 as well (2^-9 per element)
  wtol = 4e-3 if (mode == "bf16" and K.STEM_WGRAD_TC["on"]) else 1e-3
  K.stem_wgrad(x, to_nhwc(dy, tdt), geo, dt, gw, False)
  assert torch.allclose(gw, wr.grad, rtol=1e-3, atol=wtol * wr.grad.abs().max().item())
  K.stem_wgrad(x, to_nhwc(dy, tdt), geo, dt, gw, True)
  assert torch.allclose(gw, 2 * wr.grad, rtol=1e-3, atol=2 * wtol * wr.grad.abs().max().item())


@pytest.mark.parametrize("n,F_,S,k", [(37, 512, 5, 10), (64, 512, 5, 70), (33, 4608, 5, 50), (5, 512, 1, 3),
                                      (700, 4608, 5, 10), (1408, 512, 5, 10)])  # (the last two: split-K logits / dw GEMMs)
def test_heads(n, F_, S, k):
  K = _K()
  g = torch.Generator().manual_seed(6)
  feat = torch.randn(n, F_, generator=g).cuda()
  w = (torch.randn(S * k, F_, generator=g) * 0.05).cuda()
  b = (torch.randn(S * k, generator=g) * 0.1).cuda()
  dz = torch.randn(S, n, k, generator=g).cuda()
  z = K.heads_fwd(feat, w, b, S, k)
  fr, wr, br = [t.clone().requires_grad_(True) for t in (feat, w, b)]
  zr = torch.softmax((fr @ wr.t() + br).view(n, S, k), dim=2).permute(1, 0, 2)
  assert torch.allclose(z, zr.detach(), rtol=1e-4, atol=1e-6)
  zr.backward(dz)
  dw, db, dfeat = K.heads_bwd(feat, w, z, dz, S, k, True)
  # (fp32 sums of n terms in a different order than torch: absolute tolerance 1e-5 of the largest entry per 100 rows)
  atol = lambda t: max(1e-5, 1e-5 * t.abs().max().item() * max(1.0, n / 100.0))  # noqa: E731
  assert torch.allclose(dw, wr.grad, rtol=1e-3, atol=atol(wr.grad))
  assert torch.allclose(db, br.grad, rtol=1e-3, atol=atol(br.grad))
  assert torch.allclose(dfeat, fr.grad, rtol=1e-3, atol=atol(fr.grad))


def test_sobel_matches_oracle():
  from iic_b200.utils.cluster.transforms import sobel_process
  for (c, rgb, ir) in [(1, False, False), (4, True, False), (2, False, True), (5, True, True)]:
    x = weights.uniform("sobel.gpu.%d" % c, (3, c, 33, 47))
    o = sobel_process(x.cuda(), rgb, ir)
    r = otf.sobel_process(x, rgb, ir)
    assert o.shape == r.shape
    assert torch.allclose(o.cpu(), r, rtol=0, atol=2e-6)
  with pytest.raises(AssertionError):
    sobel_process(torch.zeros(1, 3, 8, 8).cuda(), False)


def test_adam_matches_torch():
  K = _K()
  g = torch.Generator().manual_seed(7)
  shapes = [(64, 2, 3, 3), (64,), (512, 512, 3, 3), (10, 512), (3,)] * 13  # > 48 tensors: several launches
  ps = [torch.randn(s, generator=g).cuda() for s in shapes]
  ref = [p.clone().requires_grad_(True) for p in ps]
  opt = torch.optim.Adam(ref, lr=1e-3)
  ms = [torch.zeros_like(p) for p in ps]
  vs = [torch.zeros_like(p) for p in ps]
  for step in range(1, 4):
    gs = [torch.randn(s, generator=g).cuda() for s in shapes]
    for r, gg in zip(ref, gs):
      r.grad = gg.clone()
    opt.step()
    K.adam_step(ps, gs, ms, vs, 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
  for p, r in zip(ps, ref):
    assert torch.allclose(p, r.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("h,cin,cout,k,s,p", [(49, 64, 64, 3, 1, 1), (49, 64, 128, 3, 2, 1), (25, 128, 128, 3, 1, 1),
                                               (13, 256, 256, 3, 1, 1), (13, 256, 512, 1, 2, 0), (7, 512, 512, 3, 1, 1)])
def test_full_size_conv_adjoint_identities(h, cin, cout, k, s, p):
  """BASELINE size (88 pairs per GPU x 2 views = 176 images, ClusterNet5g layer shapes): the three
  tensor-core kernels must be mutually adjoint, <conv(x,w), dy> == <x, dgrad(dy,w)> == <w, wgrad(x,dy)>,
  a size-independent property that needs no oracle run."""
  K = _K()
  from iic_b200._lib import BF16
  n = 176
  g = torch.Generator().manual_seed(11)
  x = torch.randn(n, h, h, cin, generator=g).cuda().bfloat16()
  w = (torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))).cuda().bfloat16().float()
  geo = K.conv_geom(n, h, h, cin, cout, k, k, s, p, 1)
  dy = torch.randn(n, geo.oh, geo.ow, cout, generator=g).cuda().bfloat16()
  y = K.conv_fprop(x, K.pack_weight(w, BF16, 0), geo, BF16)
  dx = K.conv_dgrad(dy, K.pack_weight(w, BF16, 1), geo, BF16)
  dw = torch.zeros_like(w)
  K.conv_wgrad(x, dy, geo, BF16, dw, False)
  a = (y.double() * dy.double()).sum().item()
  b = (x.double() * dx.double()).sum().item()
  c = (w.double() * dw.double()).sum().item()
  scale = (y.double().norm() * dy.double().norm()).item()
  assert abs(a - c) <= 2e-3 * scale and abs(b - c) <= 2e-3 * scale, (a, b, c, scale)


@pytest.mark.parametrize("n,h,cin,cout,k,s,p", [(6, 13, 64, 64, 3, 1, 1), (4, 25, 64, 128, 3, 2, 1), (10, 7, 256, 512, 3, 1, 1),
                                                 (2, 49, 64, 64, 3, 1, 1), (32, 49, 64, 64, 3, 1, 1)])
def test_conv_epilogue_bn_statistics(n, h, cin, cout, k, s, p):
  """BN statistics accumulated in the tcgen05 conv epilogue (two views in one launch, statistics kept
  per view) == statistics of a separate pass over the stored output."""
  K = _K()
  from iic_b200._lib import BF16
  g = torch.Generator().manual_seed(21)
  x = torch.randn(n, h, h, cin, generator=g).cuda().bfloat16()
  x[n // 2:] = x[n // 2:] * 2.0 + 0.5  # make the two views' statistics clearly different
  w = (torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))).cuda()
  gamma = (torch.rand(cout, generator=g) + 0.5).cuda()
  beta = (torch.randn(cout, generator=g) * 0.1).cuda()
  geo = K.conv_geom(n, h, h, cin, cout, k, k, s, p, 1)
  wp = K.pack_weight(w, BF16, 0)
  y, partial, nblk = K.conv_fprop_stats(x, wp, geo, BF16, 2)
  y_ref = K.conv_fprop(x, wp, geo, BF16)
  assert torch.equal(y, y_ref)
  M = (n // 2) * geo.oh * geo.ow
  for v in range(2):
    rm, rv = torch.zeros(cout).cuda(), torch.ones(cout).cuda()
    rm2, rv2 = torch.zeros(cout).cuda(), torch.ones(cout).cuda()
    ss, mi = K.bn_stats_from_partials(partial, nblk, 2, v, M, gamma, beta, 1e-5, 0.1, rm, rv)
    ss2, mi2 = K.bn_stats(y[v * (n // 2):(v + 1) * (n // 2)], gamma, beta, 1e-5, 0.1, rm2, rv2, False)
    assert torch.allclose(mi[:cout], mi2[:cout], rtol=0, atol=3e-3 * mi2[:cout].abs().max().item() + 1e-3)  # mean
    assert torch.allclose(mi[cout:], mi2[cout:], rtol=5e-3, atol=0)  # invstd
    assert torch.allclose(ss, ss2, rtol=1e-2, atol=5e-3)
    assert torch.allclose(rv, rv2, rtol=1e-2, atol=1e-4)
  # one view: second slot of every partial row is zero, first holds the whole batch
  y1, partial1, nblk1 = K.conv_fprop_stats(x, wp, geo, BF16, 1)
  assert partial1.shape[1] == 2 and float(partial1[:, 1].abs().max()) == 0.0
  ss, mi = K.bn_stats_from_partials(partial1, nblk1, 2, 0, 2 * M, gamma, beta, 1e-5, 0.1, None, None)
  ss2, mi2 = K.bn_stats(y, gamma, beta, 1e-5, 0.1, None, None, False)
  assert torch.allclose(ss, ss2, rtol=1e-2, atol=5e-3)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(6, 13, 13, 64), (4, 7, 7, 512), (2, 5, 5, 2048), (64, 25, 25, 128)])
@pytest.mark.parametrize("views", [1, 2])
def test_bn_multi_view_kernels_match_per_view_chain(mode, shape, views):
  """One-launch-for-all-views kernels (bn_apply_views, the cooperative bn_bwd_fused, bn_stats_from_partials_views'
  consumer layout) against the per-view launch chain they replace: same arithmetic per element, so the
  tolerance only covers the different partial-sum grouping."""
  K = _K()
  tdt = torch.float32 if mode == "fp32" else torch.bfloat16
  n, h, w, C = shape
  g = torch.Generator().manual_seed(31)
  y = (torch.randn(n, h, w, C, generator=g) * 2 + 0.5).cuda().to(tdt)
  y[n // 2:] = (y[n // 2:].float() * 0.5 - 1.0).to(tdt)
  res = torch.randn(n, h, w, C, generator=g).cuda().to(tdt)
  dout = torch.randn(n, h, w, C, generator=g).cuda().to(tdt)
  gamma = (torch.rand(C, generator=g) + 0.5).cuda()
  beta = (torch.randn(C, generator=g) * 0.1).cuda()
  nv = n // views
  ss = torch.empty(views, 2 * C).cuda()
  mi = torch.empty(views, 2 * C).cuda()
  for v in range(views):
    K.bn_stats(y[v * nv:(v + 1) * nv], gamma, beta, 1e-5, 0.1, None, None, False, ss=ss[v], mi=mi[v])
  # forward
  out = K.bn_apply_views(y, ss, True, views, res=res, rss=ss)
  out_ref = torch.empty_like(y)
  for v in range(views):
    sl = slice(v * nv, (v + 1) * nv)
    K.bn_apply(y[sl], ss[v], True, res=res[sl], rss=ss[v], out=out_ref[sl])
  assert torch.equal(out, out_ref)
  # backward, both mask flavours
  for act, mss in ((out, None), (None, [ss[v] for v in range(views)]), (None, None)):
    dg, db = torch.full((C,), 3.0).cuda(), torch.full((C,), -2.0).cuda()
    dy, go = K.bn_bwd_fused(dout, act, y, [mi[v] for v in range(views)], gamma, dg, db, True, act is not None, mask_sss=mss)
    dg_r, db_r = torch.full((C,), 3.0).cuda(), torch.full((C,), -2.0).cuda()
    dy_r, go_r = torch.empty_like(y), torch.empty_like(y)
    for v in range(views):
      sl = slice(v * nv, (v + 1) * nv)
      K.bn_bwd(dout[sl], None if act is None else act[sl], y[sl], mi[v], gamma, dg_r, db_r, True, act is not None,
               dy=dy_r[sl], g_out=go_r[sl], mask_ss=None if mss is None else mss[v])
    scale = dy_r.float().abs().max().item()
    assert (dy.float() - dy_r.float()).abs().max().item() <= (1e-5 if mode == "fp32" else 1e-2) * scale
    assert torch.allclose(dg, dg_r, rtol=1e-4, atol=1e-3 * dg_r.abs().max().item())
    assert torch.allclose(db, db_r, rtol=1e-4, atol=1e-3 * db_r.abs().max().item())
    if act is not None:
      assert torch.equal(go, go_r)
    else:
      assert go is None


def test_bn_stats_from_partials_all_views_in_one_launch():
  K = _K()
  from iic_b200._lib import BF16
  g = torch.Generator().manual_seed(22)
  n, h, cin, cout = 6, 13, 64, 128
  x = torch.randn(n, h, h, cin, generator=g).cuda().bfloat16()
  x[n // 2:] = x[n // 2:] * 2.0 + 0.5
  w = (torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))).cuda()
  gamma = (torch.rand(cout, generator=g) + 0.5).cuda()
  beta = (torch.randn(cout, generator=g) * 0.1).cuda()
  geo = K.conv_geom(n, h, h, cin, cout, 3, 3, 1, 1, 1)
  for views in (1, 2):
    y, partial, nblk = K.conv_fprop_stats(x, K.pack_weight(w, BF16, 0), geo, BF16, views)
    M = (n // views) * geo.oh * geo.ow
    rm, rv = torch.zeros(cout).cuda(), torch.ones(cout).cuda()
    rm2, rv2 = torch.zeros(cout).cuda(), torch.ones(cout).cuda()
    ss, mi = K.bn_stats_from_partials_views(partial, nblk, 2, views, M, gamma, beta, 1e-5, 0.1, rm, rv)
    for v in range(views):  # same fold order => identical to the per-view call; running stats chained view by view
      ss1, mi1 = K.bn_stats_from_partials(partial, nblk, 2, v, M, gamma, beta, 1e-5, 0.1, rm2, rv2)
      assert torch.equal(ss[v], ss1) and torch.equal(mi[v], mi1)
    assert torch.equal(rm, rm2) and torch.equal(rv, rv2)


@pytest.fixture(params=[1, 2], ids=["quad-consecutive", "quad-interleaved"])
def quad_stem(request):
  """The fused-statistics entry point belongs to the quad stem kernels; both thread layouts are checked whatever the
  process default (IIC_STEM_QUAD) is."""
  K = _K()
  with K.options(stem_quad=request.param):
    yield


@pytest.mark.parametrize("cin,k,pad,hw", [(2, 3, 1, 24), (1, 5, 2, 24), (5, 3, 1, 16), (2, 3, 1, 96)])
@pytest.mark.parametrize("views", [1, 2])
def test_stem_fprop_with_fused_bn_statistics(cin, k, pad, hw, views, quad_stem):
  """Quad stem kernel: same output as the plain entry point, statistics == a separate pass (fp32 accumulators vs
  bf16-rounded storage: loose tolerance), per view."""
  K = _K()
  from iic_b200._lib import BF16
  g = torch.Generator().manual_seed(41)
  n = 6
  x = torch.randn(n, cin, hw, hw, generator=g).cuda()
  x[n // 2:] = x[n // 2:] * 1.5 + 0.3
  w = (torch.randn(64, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))).cuda()
  gamma = (torch.rand(64, generator=g) + 0.5).cuda()
  beta = (torch.randn(64, generator=g) * 0.1).cuda()
  geo = K.conv_geom(n, hw, hw, cin, 64, k, k, 1, pad, 1)
  y_ref = F.conv2d(x.double().cpu(), w.double().cpu(), None, 1, pad).permute(0, 2, 3, 1).float().cuda()
  fused = K.stem_fprop_stats(x, w, geo, BF16, views)
  assert fused is not None
  y, partial, nblk = fused
  assert torch.equal(y, K.stem_fprop(x, w, geo, BF16))
  assert torch.allclose(y.float(), y_ref, rtol=1e-2, atol=1e-2)
  M = (n // views) * hw * hw
  ss, mi = K.bn_stats_from_partials_views(partial, nblk, 2, views, M, gamma, beta, 1e-5, 0.1, None, None)
  for v in range(views):
    sl = slice(v * (n // views), (v + 1) * (n // views))
    mean = y_ref[sl].mean(dim=(0, 1, 2))
    var = y_ref[sl].var(dim=(0, 1, 2), unbiased=False)
    assert torch.allclose(mi[v, :64], mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(mi[v, 64:], 1.0 / torch.sqrt(var + 1e-5), rtol=1e-4)
  # fp32 storage takes the same kernel
  from iic_b200._lib import F32
  y32 = K.stem_fprop(x, w, geo, F32)
  assert torch.allclose(y32, y_ref, rtol=1e-4, atol=1e-5)


def test_batched_weight_packing_matches_single_calls():
  K = _K()
  from iic_b200._lib import BF16, F32
  g = torch.Generator().manual_seed(42)
  ws = [torch.randn(co, ci, k, k, generator=g).cuda() for co, ci, k in ((64, 64, 3), (128, 64, 1), (128, 64, 3), (512, 256, 3))]
  from iic_b200._lib import TF32X3
  for dt in (BF16, F32, TF32X3):
    plan = K.PackPlan(ws, (0, 1), dt)
    out = plan.run()
    for i, w in enumerate(ws):
      for kind in (0, 1):
        assert torch.equal(out[(i, kind)], K.pack_weight(w, dt, kind))
  # 3xTF32 weights: [raw plane | lo plane], lo = w - (w with the 13 low mantissa bits cleared), exact in fp32
  w = ws[2]
  both, raw = K.pack_weight(w, TF32X3, 0), K.pack_weight(w, F32, 0)
  assert both.shape == (2,) + tuple(raw.shape) and torch.equal(both[0], raw)
  hi = (raw.view(torch.int32) & -8192).view(torch.float32)
  assert torch.equal(both[1], raw - hi) and float(both[1].abs().max()) > 0


HALO_GEOMS = [
  # n, h: Wp = h + 2, R = 256 // Wp output rows per work item; covers R > H (one item per image, box taller than
  # the image), ragged last item (H % R != 0), R == 1 (Wp = 162) and a row too wide for two pipeline stages (h = 200: im2col fallback)
  (3, 5), (5, 13), (2, 30), (3, 49), (2, 96), (1, 126), (1, 160), (1, 200),
]


@pytest.mark.parametrize("n,h", HALO_GEOMS)
@pytest.mark.parametrize("variant", ["tma-store", "direct-store", "im2col-wgrad", "tma-addend"])
def test_halo_kernels_forced_exact_small_integers(n, h, variant):
  """The halo kernels (3x3 / stride 1 / pad 1 / 64 -> 64: fprop, dgrad, wgrad) forced on every geometry they accept
  (option conv_halo = 2; by default they only run where they pay): EXACT on small-integer operands against CPU fp64,
  and bit-identical to the im2col kernels (conv_halo = 0), including the dgrad addend and the fused BN statistics."""
  K = _K()
  from iic_b200._lib import BF16
  halo_wgrad = 0 if variant == "im2col-wgrad" else 1
  halo_store = 0 if variant == "direct-store" else 1
  if halo_wgrad == 0 and h not in (13, 49):
    pytest.skip("im2col wgrad beside the halo fprop/dgrad: two geometries are enough")
  g = torch.Generator().manual_seed(100 + h)
  x = torch.randint(-1, 2, (n, 64, h, h), generator=g).float()
  w = torch.randint(-1, 2, (64, 64, 3, 3), generator=g).float()
  dy = torch.randint(-1, 2, (n, 64, h, h), generator=g).float()
  add = torch.randint(-2, 3, (n, 64, h, h), generator=g).float()
  geo = K.conv_geom(n, h, h, 64, 64, 3, 3, 1, 1, 1)
  xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
  ref = F.conv2d(xr, wr, None, 1, 1)
  ref.backward(dy.double())
  assert ref.abs().max() <= 256 and xr.grad.abs().max() <= 250
  xh, dyh, addh = [to_nhwc(t.cuda(), torch.bfloat16) for t in (x, dy, add)]
  wp0, wp1 = K.pack_weight(w.cuda(), BF16, 0), K.pack_weight(w.cuda(), BF16, 1)

  def run():
    y = K.conv_fprop(xh, wp0, geo, BF16)
    dx = K.conv_dgrad(dyh, wp1, geo, BF16)
    dx2 = K.conv_dgrad(dyh, wp1, geo, BF16, addend=addh)
    gw = torch.zeros_like(w).cuda()
    K.conv_wgrad(xh, dyh, geo, BF16, gw, False)
    st = K.conv_fprop_stats(xh, wp0, geo, BF16, 2) if n % 2 == 0 else K.conv_fprop_stats(xh, wp0, geo, BF16, 1)
    torch.cuda.synchronize()
    return y, dx, dx2, gw, st

  # "tma-addend": the dgrad addend arrives by one TMA load in the staging buffer (option halo_addend_tma)
  with K.options(conv_halo=2, conv_halo_wgrad=halo_wgrad, conv_halo_store=halo_store,
                 halo_addend_tma=1 if variant == "tma-addend" else 0):
    y, dx, dx2, gw, st = run()
  with K.options(conv_halo=0):
    y0, dx0, dx20, gw0, st0 = run()
  assert torch.equal(from_nhwc(y).cpu(), ref.detach().float()), "halo fprop"
  assert torch.equal(from_nhwc(dx).cpu(), xr.grad.float()), "halo dgrad"
  assert torch.equal(from_nhwc(dx2).cpu(), (xr.grad + add.double()).float()), "halo dgrad + addend"
  assert torch.equal(gw.cpu(), wr.grad.float()), "halo wgrad"
  for a, b in ((y, y0), (dx, dx0), (dx2, dx20), (gw, gw0), (st[0], st0[0])):
    assert torch.equal(a, b)
  # statistics: integer sums, exact in fp32 whatever the partition into per-CTA partial rows
  views = 2 if n % 2 == 0 else 1
  tot = st[1][:st[2]].double().sum(0).cpu()  # [2 slots][{sum, sumsq}][64]
  yr = ref.detach()
  for v in range(views):
    sl = yr[v * (n // views):(v + 1) * (n // views)]
    assert torch.equal(tot[v, 0], sl.sum(dim=(0, 2, 3))) and torch.equal(tot[v, 1], (sl * sl).sum(dim=(0, 2, 3)))
  if views == 1:
    assert float(tot[1].abs().max()) == 0.0


def test_runtime_options_roundtrip():
  K = _K()
  for name in ("conv_halo", "conv_halo_wgrad", "stem_quad", "dgrad_prefetch", "tc2_mt2", "conv_halo_store", "stem_bwd_v2", "bn_bwd_ctas", "conv_halo_stats", "tf32x3_raw_hi", "wgrad_mt", "halo_addend_tma", "dgrad_s2_mt"):
    v = K.get_option(name)
    with K.options(**{name: 0}):
      assert K.get_option(name) == 0
    assert K.get_option(name) == v
  with pytest.raises(AssertionError):
    K.get_option("no_such_option")


@pytest.mark.parametrize("cin,k,pad,hw", [(2, 3, 1, 24), (1, 5, 2, 24), (5, 3, 1, 16)])
def test_stem_quad_kernel_equals_one_pixel_kernel(cin, k, pad, hw):
  """The two stem conv kernels accumulate in the same (ci, a, b) order: bit-identical fp32 results."""
  K = _K()
  from iic_b200._lib import F32
  g = torch.Generator().manual_seed(43)
  x = torch.randn(4, cin, hw, hw, generator=g).cuda()
  w = (torch.randn(64, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))).cuda()
  geo = K.conv_geom(4, hw, hw, cin, 64, k, k, 1, pad, 1)
  ys = []
  for mode in (0, 1, 2):
    with K.options(stem_quad=mode):
      ys.append(K.stem_fprop(x, w, geo, F32))
  assert torch.allclose(ys[1], ys[0], rtol=1e-6, atol=1e-6)
  assert torch.equal(ys[2], ys[1])  # the two quad layouts differ only in which thread owns which channel


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,hw,pool_pad,views,n", [(2, 32, 1, 2, 6), (2, 96, 1, 2, 4), (1, 24, 0, 1, 3), (2, 20, 0, 2, 2),
                                                     (2, 18, 1, 1, 5)])
def test_stem_backward_fused_matches_chain_and_autograd(mode, cin, hw, pool_pad, views, n):
  """iic_stem_bwd_fused (conv3x3 -> BN -> ReLU -> MaxPool backward in two passes over (y, dpool)) against the
  three-kernel chain it replaces (max-pool backward, BatchNorm backward, stem wgrad) and, in fp32 mode, against
  torch autograd in double precision -- per-view BatchNorm statistics, both pool paddings, accumulate flags."""
  K = _K()
  from iic_b200._lib import BF16, F32
  dt, tdt = (F32, torch.float32) if mode == "fp32" else (BF16, torch.bfloat16)
  g = torch.Generator().manual_seed(77)
  x = torch.randn(n, cin, hw, hw, generator=g).cuda()
  if views == 2:
    x[n // 2:] = x[n // 2:] * 1.7 + 0.4
  w = (torch.randn(64, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))).cuda()
  gamma = (torch.rand(64, generator=g) + 0.5).cuda()
  beta = (torch.randn(64, generator=g) * 0.2).cuda()
  geo = K.conv_geom(n, hw, hw, cin, 64, 3, 3, 1, 1, 1)
  y = K.stem_fprop(x, w, geo, dt)
  nv = n // views
  ss = torch.empty(views, 128, device="cuda")
  mi = torch.empty(views, 128, device="cuda")
  for v in range(views):
    K.bn_stats(y[v * nv:(v + 1) * nv], gamma, beta, 1e-5, 0.1, None, None, False, ss=ss[v], mi=mi[v])
  oh = (hw + 2 * pool_pad - 2) // 2 + 1
  dpool = torch.randn(n, oh, oh, 64, generator=g).cuda().to(tdt)
  # the chain
  gmask = torch.empty_like(y)
  for v in range(views):
    K.bn_relu_maxpool_bwd(y[v * nv:(v + 1) * nv], ss[v], dpool[v * nv:(v + 1) * nv], pool_pad, out=gmask[v * nv:(v + 1) * nv])
  dg1, db1 = torch.zeros(64).cuda(), torch.zeros(64).cuda()
  dy, _ = K.bn_bwd_fused(gmask, None, y, [mi[v] for v in range(views)], gamma, dg1, db1, False, False)
  gw1 = torch.zeros_like(w)
  K.stem_wgrad(x, dy, geo, dt, gw1, False)
  # fused
  assert K.stem_bwd_fused_workspace(geo, pool_pad, views, dt) > 0
  dg2, db2, gw2 = torch.full((64,), 7.0).cuda(), torch.full((64,), 7.0).cuda(), torch.full_like(w, 7.0)
  assert K.stem_bwd_fused(x, y, dpool, ss, mi, gamma, dg2, db2, False, gw2, False, geo, pool_pad, dt)
  torch.cuda.synchronize()
  sg, sb, sw = dg1.abs().max().item(), db1.abs().max().item(), gw1.abs().max().item()
  assert (dg2 - dg1).abs().max().item() <= 2e-4 * sg + 1e-5
  assert (db2 - db1).abs().max().item() <= 2e-4 * sb + 1e-5
  wtol = 2e-4 if mode == "fp32" else 1e-2  # the chain rounds dy to bf16 before the wgrad, the fused pass keeps fp32
  assert (gw2 - gw1).abs().max().item() <= wtol * sw + 1e-5
  # accumulate flags
  assert K.stem_bwd_fused(x, y, dpool, ss, mi, gamma, dg2, db2, True, gw2, True, geo, pool_pad, dt)
  assert (dg2 - 2 * dg1).abs().max().item() <= 4e-4 * sg + 2e-5
  assert (gw2 - 2 * gw1).abs().max().item() <= 2 * wtol * sw + 2e-5
  if mode == "fp32":
    xr, wr = x.double().cpu(), w.double().cpu().requires_grad_(True)
    gr, br = gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, 1)
    outs = [F.max_pool2d(torch.relu(F.batch_norm(yr[v * nv:(v + 1) * nv], None, None, gr, br, True, 0.1, 1e-5)), 2, 2, pool_pad)
            for v in range(views)]
    torch.cat(outs).backward(dpool.double().cpu().permute(0, 3, 1, 2))
    K.stem_bwd_fused(x, y, dpool, ss, mi, gamma, dg2, db2, False, gw2, False, geo, pool_pad, dt)
    assert torch.allclose(gw2.cpu().double(), wr.grad, rtol=0, atol=1e-3 * wr.grad.abs().max().item())
    assert torch.allclose(dg2.cpu().double(), gr.grad, rtol=0, atol=1e-3 * gr.grad.abs().max().item())
    assert torch.allclose(db2.cpu().double(), br.grad, rtol=0, atol=1e-3 * br.grad.abs().max().item())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,hw,pool_pad,views,n", [(2, 32, 1, 2, 6), (2, 96, 1, 2, 4), (1, 24, 0, 1, 3), (2, 20, 0, 2, 2),
                                                     (2, 18, 1, 1, 5)])
def test_stem_backward_dy_matches_chain(mode, cin, hw, pool_pad, views, n):
  """iic_stem_bwd_dy (max-pool routing + ReLU + BatchNorm backward in two passes over (y, dpool), dy written) against the
  two-kernel chain it replaces: dy, dgamma, dbeta, accumulate flag."""
  K = _K()
  from iic_b200._lib import BF16, F32
  dt, tdt = (F32, torch.float32) if mode == "fp32" else (BF16, torch.bfloat16)
  g = torch.Generator().manual_seed(78)
  x = torch.randn(n, cin, hw, hw, generator=g).cuda()
  if views == 2:
    x[n // 2:] = x[n // 2:] * 1.7 + 0.4
  w = (torch.randn(64, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))).cuda()
  gamma = (torch.rand(64, generator=g) + 0.5).cuda()
  beta = (torch.randn(64, generator=g) * 0.2).cuda()
  geo = K.conv_geom(n, hw, hw, cin, 64, 3, 3, 1, 1, 1)
  old = K.STEM_FPROP_TC["on"]
  K.STEM_FPROP_TC["on"] = False
  try:
    y = K.stem_fprop(x, w, geo, dt)
  finally:
    K.STEM_FPROP_TC["on"] = old
  nv = n // views
  ss = torch.empty(views, 128, device="cuda")
  mi = torch.empty(views, 128, device="cuda")
  for v in range(views):
    K.bn_stats(y[v * nv:(v + 1) * nv], gamma, beta, 1e-5, 0.1, None, None, False, ss=ss[v], mi=mi[v])
  oh = (hw + 2 * pool_pad - 2) // 2 + 1
  dpool = torch.randn(n, oh, oh, 64, generator=g).cuda().to(tdt)
  gmask = torch.empty_like(y)
  for v in range(views):
    K.bn_relu_maxpool_bwd(y[v * nv:(v + 1) * nv], ss[v], dpool[v * nv:(v + 1) * nv], pool_pad, out=gmask[v * nv:(v + 1) * nv])
  dg1, db1 = torch.zeros(64).cuda(), torch.zeros(64).cuda()
  dy1, _ = K.bn_bwd_fused(gmask, None, y, [mi[v] for v in range(views)], gamma, dg1, db1, False, False)
  dg2, db2 = torch.full((64,), 7.0).cuda(), torch.full((64,), 7.0).cuda()
  dy2 = K.stem_bwd_dy(y, dpool, ss, mi, gamma, dg2, db2, False, geo, pool_pad, dt)
  assert dy2 is not None
  torch.cuda.synchronize()
  sg, sb, sd = dg1.abs().max().item(), db1.abs().max().item(), dy1.float().abs().max().item()
  assert (dg2 - dg1).abs().max().item() <= 2e-4 * sg + 1e-5
  assert (db2 - db1).abs().max().item() <= 2e-4 * sb + 1e-5
  dtol = 2e-5 if mode == "fp32" else 1e-2  # (bf16: one rounding of dy; the coefficients differ in the last fp32 bits)
  assert (dy2.float() - dy1.float()).abs().max().item() <= dtol * sd
  dy3 = K.stem_bwd_dy(y, dpool, ss, mi, gamma, dg2, db2, True, geo, pool_pad, dt)
  assert (dg2 - 2 * dg1).abs().max().item() <= 4e-4 * sg + 2e-5 and torch.equal(dy3, dy2)


def test_stem_backward_fused_reports_unsupported_geometries():
  K = _K()
  from iic_b200._lib import BF16
  assert K.stem_bwd_fused_workspace(K.conv_geom(2, 24, 24, 1, 64, 5, 5, 1, 2, 1), 0, 1, BF16) == 0  # 5x5 (ClusterNet6c)
  assert K.stem_bwd_fused_workspace(K.conv_geom(2, 32, 32, 5, 64, 3, 3, 1, 1, 1), 1, 1, BF16) == 0  # cin 5 (net10a)
  assert K.stem_bwd_fused_workspace(K.conv_geom(2, 33, 33, 2, 64, 3, 3, 1, 1, 1), 0, 1, BF16) == 0  # a row no window covers
  assert K.stem_bwd_fused_workspace(K.conv_geom(3, 32, 32, 2, 64, 3, 3, 1, 1, 1), 1, 2, BF16) == 0  # 3 images, 2 views
  assert K.stem_bwd_fused_workspace(K.conv_geom(4, 32, 32, 2, 64, 3, 3, 1, 1, 1), 1, 2, BF16) > 0


@pytest.mark.parametrize("n,h,cin,cout,k,s,p", [(2, 9, 128, 128, 3, 1, 1), (5, 13, 128, 128, 3, 1, 1), (4, 25, 64, 128, 3, 2, 1),
                                                 (3, 25, 64, 128, 1, 2, 0), (2, 30, 128, 128, 3, 1, 1)])
def test_tc2_two_tile_work_items_forced_exact_small_integers(n, h, cin, cout, k, s, p):
  """N = 128 tiles of the TMA kernel with two 128-row accumulator tiles per weight k-block (option tc2_mt2 = 2 forces the
  variant at test sizes; by default it needs two work items per SM): exact on small integers, bit-identical to the
  one-tile variant, fused BN statistics and dgrad addend included."""
  K = _K()
  from iic_b200._lib import BF16
  g = torch.Generator().manual_seed(200 + h + cin)
  x = torch.randint(-1, 2, (n, cin, h, h), generator=g).float()
  w = torch.randint(-1, 2, (cout, cin, k, k), generator=g).float()
  geo = K.conv_geom(n, h, h, cin, cout, k, k, s, p, 1)
  dy = torch.randint(-1, 2, (n, cout, geo.oh, geo.ow), generator=g).float()
  add = torch.randint(-2, 3, (n, cin, h, h), generator=g).float()
  xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
  ref = F.conv2d(xr, wr, None, s, p)
  ref.backward(dy.double())
  assert ref.abs().max() <= 256 and xr.grad.abs().max() <= 250
  xh, dyh, addh = [to_nhwc(t.cuda(), torch.bfloat16) for t in (x, dy, add)]
  wp0, wp1 = K.pack_weight(w.cuda(), BF16, 0), K.pack_weight(w.cuda(), BF16, 1)

  def run():
    y = K.conv_fprop(xh, wp0, geo, BF16)
    st = K.conv_fprop_stats(xh, wp0, geo, BF16, 2 if n % 2 == 0 else 1)
    dx = K.conv_dgrad(dyh, wp1, geo, BF16)
    dx2 = K.conv_dgrad(dyh, wp1, geo, BF16, addend=addh)
    torch.cuda.synchronize()
    return y, st, dx, dx2

  with K.options(tc2_mt2=2):
    y, st, dx, dx2 = run()
  with K.options(tc2_mt2=0):
    y0, st0, dx0, dx20 = run()
  assert torch.equal(from_nhwc(y).cpu(), ref.detach().float()), "fprop"
  assert torch.equal(from_nhwc(dx).cpu(), xr.grad.float()), "dgrad"
  assert torch.equal(from_nhwc(dx2).cpu(), (xr.grad + add.double()).float()), "dgrad + addend"
  for a, b in ((y, y0), (st[0], st0[0]), (dx, dx0), (dx2, dx20)):
    assert torch.equal(a, b)
  views = 2 if n % 2 == 0 else 1
  tot = st[1][:st[2]].double().sum(0).cpu()
  yr = ref.detach()
  for v in range(views):
    sl = yr[v * (n // views):(v + 1) * (n // views)]
    assert torch.equal(tot[v, 0], sl.sum(dim=(0, 2, 3))) and torch.equal(tot[v, 1], (sl * sl).sum(dim=(0, 2, 3)))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("shape,views", [((6, 13, 13, 64), 2), ((4, 7, 7, 512), 1), ((2, 5, 5, 2048), 2), ((64, 25, 25, 128), 2)])
def test_bn_relu_bitmask_variant_equals_activation_mask(mode, shape, views):
  """1-bit ReLU mask (iic_bn_apply_views_mask / iic_bn_bwd_fused_bits, engine switch bn_bitmask): same output as the plain
  apply, bits == (out > 0), and a backward bit-identical to the one that re-reads the activation."""
  K = _K()
  tdt = torch.float32 if mode == "fp32" else torch.bfloat16
  n, h, w, C = shape
  g = torch.Generator().manual_seed(55)
  y = torch.randn(shape, generator=g).cuda().to(tdt)
  res = torch.randn(shape, generator=g).cuda().to(tdt)
  gin = torch.randn(shape, generator=g).cuda().to(tdt)
  gamma = (torch.rand(C, generator=g) + 0.5).cuda()
  beta = (torch.randn(C, generator=g) * 0.3).cuda()
  nv = n // views
  ss = torch.empty(views, 2 * C, device="cuda")
  mi = torch.empty(views, 2 * C, device="cuda")
  for v in range(views):
    K.bn_stats(y[v * nv:(v + 1) * nv], gamma, beta, 1e-5, 0.1, None, None, False, ss=ss[v], mi=mi[v])
  out_ref = K.bn_apply_views(y, ss, True, views, res=res)
  out, mbits = K.bn_apply_views_mask(y, ss, views, res=res)
  assert torch.equal(out, out_ref)
  M = n * h * w
  want = (out.view(M, C // 8, 8) > 0).to(torch.int32)
  weights_ = (2 ** torch.arange(8, device="cuda", dtype=torch.int32)).view(1, 1, 8)
  assert torch.equal(mbits.to(torch.int32), (want * weights_).sum(-1))
  mis = [mi[v] for v in range(views)]
  dg1, db1 = torch.zeros(C).cuda(), torch.zeros(C).cuda()
  dg2, db2 = torch.full((C,), 3.0).cuda(), torch.full((C,), 3.0).cuda()
  dy1, go1 = K.bn_bwd_fused(gin, out, y, mis, gamma, dg1, db1, False, True)
  dy2, go2 = K.bn_bwd_fused_bits(gin, mbits, y, mis, gamma, dg2, db2, False, True)
  assert torch.equal(dy1, dy2) and torch.equal(go1, go2) and torch.equal(dg1, dg2) and torch.equal(db1, db2)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,hw,pool_pad,views,n", [(2, 32, 1, 2, 6), (2, 96, 1, 2, 4), (1, 24, 0, 1, 3), (2, 6, 0, 1, 3),
                                                     (2, 18, 1, 1, 5)])
def test_stem_backward_fused_v2(mode, cin, hw, pool_pad, views, n):
  """Second version of the fused stem backward's wgrad pass (option stem_bwd_v2: half-warp cooperative fetch of the
  input neighbourhood + shuffles); (2, 6, 0, 1, 3) has an odd number of windows, i.e. an idle half-warp at the tail."""
  K = _K()
  with K.options(stem_bwd_v2=1):
    test_stem_backward_fused_matches_chain_and_autograd(mode, cin, hw, pool_pad, views, n)


@pytest.mark.parametrize("n,h", [(2, 13), (4, 49), (2, 96), (6, 5), (3, 30)])
@pytest.mark.parametrize("variant", ["tma-store", "direct-store"])
def test_halo_fprop_with_per_lane_running_statistics(n, h, variant):
  """conv_halo_stats = 1: the halo fprop keeps per-lane running BatchNorm sums over a CTA's work items (one warp
  reduction per CTA and view); statistics and output must equal the per-chunk variant exactly on small integers."""
  K = _K()
  with K.options(conv_halo_stats=1):
    test_halo_kernels_forced_exact_small_integers(n, h, variant)


@pytest.mark.parametrize("case", [(3, 13, 64, 64, 3, 1, 1, 1), (32, 49, 64, 64, 3, 1, 1, 1), (3, 25, 64, 128, 3, 2, 1, 1),
                                  (3, 25, 64, 128, 1, 2, 0, 1), (4, 7, 256, 512, 3, 1, 1, 1), (2, 12, 64, 128, 5, 1, 2, 1)])
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_wgrad_written_in_torch_layout_equals_wgrad_plus_unpack(case, mode):
  """iic_conv_wgrad_oihw (the split-K fold writes [cout][cin][kh][kw] itself on the bf16 path) == iic_conv_wgrad +
  iic_unpack_wgrad, bit for bit (same partial sums, same fold order), with and without accumulation."""
  K = _K()
  from iic_b200._lib import BF16, F32
  n, h, cin, cout, k, s, p, d = case
  dt, tdt = (F32, torch.float32) if mode == "fp32" else (BF16, torch.bfloat16)
  x, w, dy, _ = _conv_inputs(case)
  g = K.conv_geom(n, h, h, cin, cout, k, k, s, p, d)
  xh, dyh = to_nhwc(x, tdt), to_nhwc(dy, tdt)
  base = torch.randn_like(w)
  outs = {}
  for fused in (False, True):
    old = K.WGRAD_FUSED_UNPACK["on"]
    K.WGRAD_FUSED_UNPACK["on"] = fused
    try:
      a = torch.zeros_like(w)
      K.conv_wgrad(xh, dyh, g, dt, a, False)
      b = base.clone()
      K.conv_wgrad(xh, dyh, g, dt, b, True)
      outs[fused] = (a, b)
    finally:
      K.WGRAD_FUSED_UNPACK["on"] = old
  assert torch.equal(outs[False][0], outs[True][0])
  assert torch.equal(outs[False][1], outs[True][1])


@pytest.mark.parametrize("cin,k,pad,hw,n", [(2, 3, 1, 32, 3), (2, 3, 1, 96, 4), (1, 5, 2, 24, 5), (3, 3, 1, 20, 2), (2, 3, 1, 7, 1),
                                            (5, 3, 1, 20, 2), (4, 3, 1, 128, 2), (2, 5, 2, 24, 3)])  # (the last three: 33..64 taps)
def test_stem_wgrad_on_tensor_cores_matches_the_simt_kernel(cin, k, pad, hw, n):
  """Stem wgrad on tcgen05 (stem_tc.cu: patches gathered into shared memory, dy by TMA; option STEM_WGRAD_TC) against
  torch autograd on the same bf16-rounded operands and against the SIMT Gram-product kernel."""
  K = _K()
  from iic_b200._lib import BF16
  cout = 64
  g = torch.Generator().manual_seed(9)
  x = torch.randn(n, cin, hw, hw, generator=g).cuda()
  dy = torch.randn(n, cout, hw, hw, generator=g).cuda().bfloat16().float()
  geo = K.conv_geom(n, hw, hw, cin, cout, k, k, 1, pad, 1)
  wr = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
  F.conv2d(x.bfloat16().float(), wr, None, 1, pad).backward(dy)  # (the patches are rounded to bf16 by the gather)
  dyh = to_nhwc(dy, torch.bfloat16)
  base = torch.randn(cout, cin, k, k, generator=g).cuda()
  a = torch.zeros_like(base)
  assert K.stem_wgrad_tc(x, dyh, geo, a, False)
  scale = wr.grad.abs().max().item()
  assert (a - wr.grad).abs().max().item() <= 2e-3 * scale
  b = base.clone()
  assert K.stem_wgrad_tc(x, dyh, geo, b, True)
  assert (b - base - wr.grad).abs().max().item() <= 2e-3 * scale
  c = torch.zeros_like(base)
  old = K.STEM_WGRAD_TC["on"]
  K.STEM_WGRAD_TC["on"] = False
  try:
    K.stem_wgrad(x, dyh, geo, BF16, c, False)
  finally:
    K.STEM_WGRAD_TC["on"] = old
  assert (a - c).abs().max().item() <= 1e-2 * scale  # (the SIMT kernel reads x in fp32)


@pytest.mark.parametrize("cin,k,pad,hw,n,views", [(2, 3, 1, 32, 4, 2), (2, 3, 1, 96, 2, 2), (1, 5, 2, 24, 5, 1), (3, 3, 1, 20, 3, 1),
                                                  (2, 3, 1, 16, 6, 2), (2, 3, 1, 7, 1, 1)])
def test_stem_fprop_on_tensor_cores(cin, k, pad, hw, n, views):
  """Stem conv + BN statistics on tcgen05 (stem_tc.cu: patches gathered into shared memory, resident weights, TMA-stored
  bf16 output) against torch on the same bf16-rounded operands, and its statistics partials against the sums of the
  fp32 results."""
  K = _K()
  g = torch.Generator().manual_seed(21)
  x = torch.randn(n, cin, hw, hw, generator=g).cuda()
  w = (torch.randn(64, cin, k, k, generator=g) * 0.3).cuda()
  geo = K.conv_geom(n, hw, hw, cin, 64, k, k, 1, pad, 1)
  r = K.stem_fprop_stats_tc(x, w, geo, views)
  if views == 2 and (n // 2 * hw * hw) % 128 != 0:
    assert r is None
    return
  assert r is not None
  y, partial, nblk = r
  yr = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), None, 1, pad)  # exact products of the rounded operands
  got = from_nhwc(y).double()
  assert (got - yr).abs().max().item() <= 1e-2 * yr.abs().max().item()  # bf16 rounding of the stored result
  tot = partial[:nblk].double().sum(0)  # [2 views][{sum, sumsq}][64]
  nv = n // views
  for v in range(views):
    sl = yr[v * nv:(v + 1) * nv]
    s1, s2 = sl.sum(dim=(0, 2, 3)), (sl * sl).sum(dim=(0, 2, 3))
    assert (tot[v, 0] - s1).abs().max().item() <= 1e-4 * s2.sqrt().max().item() * (sl[:, 0].numel() ** 0.5)
    assert torch.allclose(tot[v, 1], s2, rtol=1e-4, atol=0)
  if views == 1:
    assert float(tot[1].abs().max()) == 0.0


@pytest.mark.parametrize("n,h,cin,cout,k,p", [(4, 25, 64, 128, 3, 1), (3, 49, 64, 128, 3, 1), (3, 25, 128, 256, 3, 1), (2, 24, 64, 128, 5, 2),
                                              (3, 25, 64, 128, 1, 0), (2, 13, 256, 512, 3, 1)])
def test_stride2_dgrad_resident_and_two_tile_variants_exact(n, h, cin, cout, k, p):
  """Option dgrad_s2_mt = 2: every parity class of a stride-2 dgrad through the resident-weights (cin 64) or two-tile
  (cin 128, with tc2_mt2 = 2) kernels, scattering epilogue with and without addend: exact on small integers and
  bit-identical to the one-tile launches."""
  K = _K()
  from iic_b200._lib import BF16
  g = torch.Generator().manual_seed(300 + h + cin)
  w = torch.randint(-1, 2, (cout, cin, k, k), generator=g).float()
  geo = K.conv_geom(n, h, h, cin, cout, k, k, 2, p, 1)
  dy = torch.randint(-1, 2, (n, cout, geo.oh, geo.ow), generator=g).float()
  add = torch.randint(-2, 3, (n, cin, h, h), generator=g).float()
  xr = torch.zeros(n, cin, h, h, dtype=torch.float64, requires_grad=True)
  F.conv2d(xr, w.double(), None, 2, p).backward(dy.double())
  assert xr.grad.abs().max() <= 250
  dyh, addh = to_nhwc(dy.cuda(), torch.bfloat16), to_nhwc(add.cuda(), torch.bfloat16)
  wp1 = K.pack_weight(w.cuda(), BF16, 1)
  outs = []
  for mode in (0, 2):
    with K.options(dgrad_s2_mt=mode, tc2_mt2=2 if mode else 1):
      dx = K.conv_dgrad(dyh, wp1, geo, BF16)
      dx2 = K.conv_dgrad(dyh, wp1, geo, BF16, addend=addh)
      torch.cuda.synchronize()
    outs.append((dx, dx2))
  for dx, dx2 in outs:
    assert torch.equal(from_nhwc(dx).cpu(), xr.grad.float())
    assert torch.equal(from_nhwc(dx2).cpu(), (xr.grad + add.double()).float())
