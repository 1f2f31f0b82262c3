"""tcgen05 kind::tf32 convolutions on fp32 activations (csrc/conv_tf32.cu): IIC_TF32 and the error-compensated
IIC_TF32X3, through the C-ABI, against CPU fp64 convolutions.

  * exact on small-integer operands (every product and partial sum is representable): any slip in a UMMA descriptor,
    the swizzle, the im2col gather, the MN-major wgrad layout, the parity-class scatter or the hi/lo split shows up
    as a wrong integer;
  * on random data: IIC_TF32 within 2e-3 of the operand scale (10-bit mantissas), IIC_TF32X3 within 5e-5 (measured 1.3e-5) -- the fp32
    SIMT kernel's own tolerance class -- which is the "stated fp32 tolerance" of the tensor-core path."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
  # n, h, cin, cout, k, stride, pad, dil
  (3, 13, 64, 64, 3, 1, 1, 1),
  (2, 49, 64, 64, 3, 1, 1, 1),
  (3, 25, 64, 128, 3, 2, 1, 1),
  (3, 25, 64, 128, 1, 2, 0, 1),
  (5, 13, 128, 256, 3, 2, 1, 1),
  (4, 7, 256, 512, 3, 1, 1, 1),
  (2, 12, 64, 128, 5, 1, 2, 1),
  (2, 16, 256, 512, 3, 1, 1, 2),
  (1, 5, 128, 128, 3, 1, 1, 1),
  (20, 24, 64, 192, 3, 1, 1, 1),  # N = 192 -> 64-wide tiles, several tiles per CTA
]


def _modes():
  from iic_b200._lib import TF32, TF32X3
  return {"tf32": TF32, "tf32x3": TF32X3}


def to_nhwc(x):
  return x.permute(0, 2, 3, 1).contiguous()


def from_nhwc(x):
  return x.permute(0, 3, 1, 2).contiguous()


def _run(case, mode, x, w, dy, add):
  from iic_b200 import kernels as K
  from iic_b200._lib import F32
  n, h, cin, cout, k, s, p, d = case
  dt = _modes()[mode]
  g = K.conv_geom(n, h, h, cin, cout, k, k, s, p, d)
  xh, dyh = to_nhwc(x), to_nhwc(dy)
  y = from_nhwc(K.conv_fprop(xh, K.pack_weight(w, K.weight_dtype(F32, dt), 0), g, dt))
  wt = K.pack_weight(w, K.weight_dtype(F32, dt), 1)  # (3xTF32: [raw plane | lo plane])
  dx = from_nhwc(K.conv_dgrad(dyh, wt, g, dt))
  dx2 = from_nhwc(K.conv_dgrad(dyh, wt, g, dt, addend=to_nhwc(add)))
  gw = torch.zeros_like(w)
  K.conv_wgrad(xh, dyh, g, dt, gw, False)
  gw2 = gw.clone()
  K.conv_wgrad(xh, dyh, g, dt, gw2, True)
  torch.cuda.synchronize()
  return y, dx, dx2, gw, gw2


def _ref(case, x, w, dy):
  n, h, cin, cout, k, s, p, d = case
  xr, wr = x.double().cpu().requires_grad_(True), w.double().cpu().requires_grad_(True)
  yr = F.conv2d(xr, wr, None, s, p, d)
  yr.backward(dy.double().cpu())
  return yr.detach(), xr.grad, wr.grad


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["tf32", "tf32x3"])
def test_tf32_conv_exact_small_integers(case, mode):
  n, h, cin, cout, k, s, p, d = case
  g = torch.Generator().manual_seed(3)
  oh = (h + 2 * p - d * (k - 1) - 1) // s + 1
  x = torch.randint(-1, 2, (n, cin, h, h), generator=g).float().cuda()
  w = torch.randint(-1, 2, (cout, cin, k, k), generator=g).float().cuda()
  dy = torch.randint(-1, 2, (n, cout, oh, oh), generator=g).float().cuda()
  add = torch.randint(-3, 4, (n, cin, h, h), generator=g).float().cuda()
  yr, dxr, dwr = _ref(case, x, w, dy)
  y, dx, dx2, gw, gw2 = _run(case, mode, x, w, dy, add)
  for name, got, want in (("fprop", y, yr), ("dgrad", dx, dxr), ("dgrad+addend", dx2, dxr + add.double().cpu()),
                          ("wgrad", gw, dwr), ("wgrad accumulate", gw2, 2 * dwr)):
    bad = (got.double().cpu() != want)
    assert not bad.any(), "%s: %d of %d wrong, max |err| %g" % (
      name, int(bad.sum()), bad.numel(), (got.double().cpu() - want).abs().max().item())


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["tf32", "tf32x3"])
def test_tf32_conv_random_data_tolerance(case, mode):
  n, h, cin, cout, k, s, p, d = case
  g = torch.Generator().manual_seed(0)
  oh = (h + 2 * p - d * (k - 1) - 1) // s + 1
  x = torch.randn(n, cin, h, h, generator=g).cuda()
  w = (torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))).cuda()
  dy = torch.randn(n, cout, oh, oh, generator=g).cuda()
  add = torch.randn(n, cin, h, h, generator=g).cuda()
  yr, dxr, dwr = _ref(case, x, w, dy)
  y, dx, dx2, gw, gw2 = _run(case, mode, x, w, dy, add)
  tol = 2e-3 if mode == "tf32" else 5e-5  # (measured 1.3e-5: fp32 accumulation order in the tensor core)
  for name, got, want in (("fprop", y, yr), ("dgrad", dx, dxr), ("dgrad+addend", dx2, dxr + add.double().cpu()),
                          ("wgrad", gw, dwr), ("wgrad accumulate", gw2, 2 * dwr)):
    err = (got.double().cpu() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= tol * scale, "%s: max |err| %g, scale %g (rel %g)" % (name, err, scale, err / scale)


def test_tf32x3_recovers_bits_that_tf32_drops():
  """x = 1 + 2^-14 is not a tf32 number: a single kind::tf32 pass loses the 2^-14; the hi/lo split keeps it."""
  from iic_b200 import kernels as K
  from iic_b200._lib import F32, TF32, TF32X3
  n, h, c = 2, 8, 64
  x = torch.full((n, h, h, c), 1.0 + 2.0 ** -14).cuda()
  w = torch.zeros(c, c, 1, 1)
  w[torch.arange(c), torch.arange(c), 0, 0] = 1.0 + 2.0 ** -13  # identity 1x1 with its own low bits
  w = w.cuda()
  g = K.conv_geom(n, h, h, c, c, 1, 1, 1, 0, 1)
  want = (1.0 + 2.0 ** -14) * (1.0 + 2.0 ** -13)
  y3 = K.conv_fprop(x, K.pack_weight(w, TF32X3, 0), g, TF32X3)
  y1 = K.conv_fprop(x, K.pack_weight(w, F32, 0), g, TF32)
  assert (y3.double() - want).abs().max().item() < 2e-7
  assert (y1.double() - want).abs().max().item() > 5e-5  # plain tf32 sees 1.0 * 1.0


@pytest.mark.parametrize("raw_hi", [0, 1])
@pytest.mark.parametrize("case", CASES[:5])
def test_tf32x3_raw_hi_operand(case, raw_hi):
  """Option tf32x3_raw_hi (default 1): the splitter writes only lo and the tensor core truncates the raw fp32 stage
  itself; 0 = hi written explicitly.  Same tolerance for both -- if the hardware ROUNDED the low 13 bits instead, hi + lo would double count up to
  2^-11 of every operand (5e-4 relative) and this fails."""
  from iic_b200 import kernels as K
  n, h, cin, cout, k, s, p, d = case
  g = torch.Generator().manual_seed(1)
  oh = (h + 2 * p - d * (k - 1) - 1) // s + 1
  x = torch.randn(n, cin, h, h, generator=g).cuda()
  w = (torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))).cuda()
  dy = torch.randn(n, cout, oh, oh, generator=g).cuda()
  add = torch.randn(n, cin, h, h, generator=g).cuda()
  yr, dxr, dwr = _ref(case, x, w, dy)
  with K.options(tf32x3_raw_hi=raw_hi):
    y, dx, dx2, gw, gw2 = _run(case, "tf32x3", x, w, dy, add)
  for name, got, want in (("fprop", y, yr), ("dgrad", dx, dxr), ("dgrad+addend", dx2, dxr + add.double().cpu()),
                          ("wgrad", gw, dwr)):
    err = (got.double().cpu() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 5e-5 * scale, "%s: max |err| %g, scale %g (rel %g)" % (name, err, scale, err / scale)
