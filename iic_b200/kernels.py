"""Thin typed wrappers: torch CUDA tensors -> raw pointers -> libiic_b200.so.

Each function enqueues its kernel(s) on ``torch.cuda.current_stream()`` and
returns immediately.  Inputs must be contiguous CUDA tensors; nothing here
touches the CPU path (there is none).
"""
import ctypes

import torch

from . import _lib
from ._lib import BF16, F32, ConvGeom, check

_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}


def torch_dtype(dt):
  return _TORCH_DT[dt]


def iic_dtype(t):
  if t.dtype == torch.float32:
    return F32
  if t.dtype == torch.bfloat16:
    return BF16
  raise AssertionError("unsupported dtype %s" % t.dtype)


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
  if t is None:
    return None
  assert t.is_cuda, "iic_b200 runs on CUDA tensors only (no CPU fallback)"
  assert t.is_contiguous(), "iic_b200 kernels need contiguous tensors"
  return ctypes.c_void_p(t.data_ptr())


def conv_geom(n, h, w, cin, cout, kh, kw, stride, pad, dil):
  oh = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
  ow = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1
  return ConvGeom(n, h, w, cin, oh, ow, cout, kh, kw, stride, pad, dil)


# ---- optional per-launch timing of the convolution kernels (bench.py roofline leg) ---------------
_conv_timing = {"on": False, "records": []}


def conv_timing(enable):
  _conv_timing["on"] = bool(enable)
  _conv_timing["records"] = []


def conv_timing_summary():
  """-> dict kind -> (launches, algorithmic FLOPs, device ms); call after torch.cuda.synchronize()."""
  out = {}
  for kind, flops, e0, e1 in _conv_timing["records"]:
    n, f, ms = out.get(kind, (0, 0.0, 0.0))
    out[kind] = (n + 1, f + flops, ms + e0.elapsed_time(e1))
  return out


class _timed(object):
  def __init__(self, kind, g):
    self.kind, self.g = kind, g

  def __enter__(self):
    if _conv_timing["on"]:
      self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      self.e0.record()
    return self

  def __exit__(self, *exc):
    if _conv_timing["on"]:
      self.e1.record()
      g = self.g
      flops = 2.0 * g.n * g.oh * g.ow * g.cout * g.kh * g.kw * g.cin  # 2*MAC, identical for fprop/dgrad/wgrad
      _conv_timing["records"].append((self.kind, flops, self.e0, self.e1))
    return False


def launch_count(reset=False):
  return int(_lib.lib().iic_launch_count(1 if reset else 0))


# ---- losses ---------------------------------------------------------------------------------
def iid_loss(z, zt, lamb, eps, want_grad, phase=_lib.PHASE_FUSED, joint_ws=None, want_joint=False):
  """z, zt: [S, n, k] fp32.  Returns (loss[S,2] | None, dz | None, dzt | None, joint_out | None)."""
  S, n, k = z.shape
  assert zt.shape == z.shape and z.dtype == torch.float32 and zt.dtype == torch.float32
  loss = None if phase == _lib.PHASE_PARTIAL else torch.empty((S, 2), device=z.device, dtype=torch.float32)
  dz = torch.empty_like(z) if (want_grad and phase != _lib.PHASE_PARTIAL) else None
  dzt = torch.empty_like(zt) if (want_grad and phase != _lib.PHASE_PARTIAL) else None
  jout = torch.empty((S, k, k), device=z.device, dtype=torch.float32) if want_joint else None
  check(_lib.lib().iic_iid_loss(_p(z), _p(zt), S, n, k, float(lamb), float(eps), _p(loss), _p(dz), _p(dzt),
                                _p(joint_ws), _p(jout), phase, _stream()), "iic_iid_loss")
  return loss, dz, dzt, jout


def sobel(imgs, include_rgb, using_ir):
  n, c, h, w = imgs.shape
  cout = (3 if include_rgb else 0) + 2 + (1 if using_ir else 0)
  out = torch.empty((n, cout, h, w), device=imgs.device, dtype=torch.float32)
  check(_lib.lib().iic_sobel(_p(imgs), _p(out), n, c, h, w, int(bool(include_rgb)), int(bool(using_ir)), _stream()),
        "iic_sobel")
  return out


# ---- layout ---------------------------------------------------------------------------------
def nchw_to_nhwc(x, dt):
  n, c, h, w = x.shape
  out = torch.empty((n, h, w, c), device=x.device, dtype=_TORCH_DT[dt])
  check(_lib.lib().iic_nchw_to_nhwc(_p(x), _p(out), dt, n, c, h, w, _stream()), "iic_nchw_to_nhwc")
  return out


def nhwc_to_nchw(x):
  n, h, w, c = x.shape
  out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
  check(_lib.lib().iic_nhwc_to_nchw(_p(x), iic_dtype(x), _p(out), n, c, h, w, _stream()), "iic_nhwc_to_nchw")
  return out


def cast(x, dt):
  out = torch.empty(x.shape, device=x.device, dtype=_TORCH_DT[dt])
  check(_lib.lib().iic_cast(_p(x), iic_dtype(x), _p(out), dt, x.numel(), _stream()), "iic_cast")
  return out


# ---- convolution ----------------------------------------------------------------------------
def pack_weight(w, dt, kind):
  cout, cin, kh, kw = w.shape
  shape = (cout, kh, kw, cin) if kind == 0 else (cin, kh, kw, cout)
  out = torch.empty(shape, device=w.device, dtype=_TORCH_DT[dt])
  check(_lib.lib().iic_pack_weight(_p(w), _p(out), dt, kind, cout, cin, kh, kw, _stream()), "iic_pack_weight")
  return out


def conv_fprop(x, wp, g, dt):
  y = torch.empty((g.n, g.oh, g.ow, g.cout), device=x.device, dtype=_TORCH_DT[dt])
  with _timed("fprop", g):
    check(_lib.lib().iic_conv_fprop(_p(x), _p(wp), _p(y), ctypes.byref(g), dt, _stream()), "iic_conv_fprop")
  return y


def conv_dgrad(dy, wpt, g, dt, addend=None):
  dx = torch.empty((g.n, g.h, g.w, g.cin), device=dy.device, dtype=_TORCH_DT[dt])
  with _timed("dgrad", g):
    check(_lib.lib().iic_conv_dgrad(_p(dy), _p(wpt), _p(addend), _p(dx), ctypes.byref(g), dt, _stream()),
          "iic_conv_dgrad")
  return dx


def conv_wgrad(x, dy, g, dt, grad_out, accumulate):
  """Accumulates (or writes) the torch-layout [cout][cin][kh][kw] fp32 gradient into grad_out."""
  nbytes = int(_lib.lib().iic_conv_wgrad_workspace(ctypes.byref(g), dt))
  ws = torch.empty((max(nbytes, 4) + 3) // 4, device=x.device, dtype=torch.float32)
  dwp = torch.empty((g.cout, g.kh, g.kw, g.cin), device=x.device, dtype=torch.float32)
  with _timed("wgrad", g):
    check(_lib.lib().iic_conv_wgrad(_p(x), _p(dy), _p(dwp), _p(ws), ctypes.byref(g), dt, _stream()), "iic_conv_wgrad")
  check(_lib.lib().iic_unpack_wgrad(_p(dwp), _p(grad_out), int(bool(accumulate)), g.cout, g.cin, g.kh, g.kw, _stream()),
        "iic_unpack_wgrad")
  return grad_out


def stem_fprop(x_nchw, w, g, dt):
  y = torch.empty((g.n, g.oh, g.ow, g.cout), device=x_nchw.device, dtype=_TORCH_DT[dt])
  check(_lib.lib().iic_stem_fprop(_p(x_nchw), _p(w), _p(y), ctypes.byref(g), dt, _stream()), "iic_stem_fprop")
  return y


def stem_wgrad(x_nchw, dy, g, dt, grad_out, accumulate):
  ws = torch.empty(2 * 1024 * 1024, device=dy.device, dtype=torch.float32)  # 8 MB of per-block partials
  check(_lib.lib().iic_stem_wgrad(_p(x_nchw), _p(dy), _p(grad_out), int(bool(accumulate)), _p(ws), ws.numel() * 4,
                                  ctypes.byref(g), dt, _stream()), "iic_stem_wgrad")
  return grad_out


# ---- batch norm / pooling ----------------------------------------------------------------------
def bn_stats(y, gamma, beta, eps, momentum, running_mean, running_var, use_running):
  C = y.shape[-1]
  M = y.numel() // C
  dev = y.device
  ws = torch.empty(2 * C, device=dev, dtype=torch.float64)
  ss = torch.empty(2 * C, device=dev, dtype=torch.float32)
  mi = torch.empty(2 * C, device=dev, dtype=torch.float32)
  check(_lib.lib().iic_bn_stats(_p(y), iic_dtype(y), M, C, _p(gamma), _p(beta), float(eps), float(momentum),
                                _p(running_mean), _p(running_var), int(bool(use_running)), _p(ws), _p(ss), _p(mi),
                                _stream()), "iic_bn_stats")
  return ss, mi


def bn_apply(y, ss, relu, res=None, rss=None):
  C = y.shape[-1]
  M = y.numel() // C
  out = torch.empty_like(y)
  check(_lib.lib().iic_bn_apply(_p(y), _p(ss), _p(res), _p(rss), _p(out), iic_dtype(y), M, C, int(bool(relu)),
                                _stream()), "iic_bn_apply")
  return out


def bn_relu_maxpool(y, ss, pad):
  n, h, w, C = y.shape
  oh, ow = (h + 2 * pad - 2) // 2 + 1, (w + 2 * pad - 2) // 2 + 1
  out = torch.empty((n, oh, ow, C), device=y.device, dtype=y.dtype)
  check(_lib.lib().iic_bn_relu_maxpool(_p(y), _p(ss), _p(out), iic_dtype(y), n, h, w, C, pad, oh, ow, _stream()),
        "iic_bn_relu_maxpool")
  return out


def bn_relu_maxpool_bwd(y, ss, dpool, pad):
  n, h, w, C = y.shape
  _, oh, ow, _ = dpool.shape
  g = torch.empty_like(y)
  check(_lib.lib().iic_bn_relu_maxpool_bwd(_p(y), _p(ss), _p(dpool), _p(g), iic_dtype(y), n, h, w, C, pad, oh, ow,
                                           _stream()), "iic_bn_relu_maxpool_bwd")
  return g


def bn_bwd(g_in, act, y, mi, gamma, dgamma, dbeta, accumulate, want_g_out):
  """Returns (dy, g_masked | None); writes/accumulates dgamma, dbeta."""
  C = y.shape[-1]
  M = y.numel() // C
  dt = iic_dtype(y)
  sums = torch.empty(2 * C, device=y.device, dtype=torch.float64)
  check(_lib.lib().iic_bn_bwd_reduce(_p(g_in), _p(act), _p(y), _p(mi), dt, M, C, _p(sums), _stream()),
        "iic_bn_bwd_reduce")
  dy = torch.empty_like(y)
  g_out = torch.empty_like(y) if want_g_out else None
  check(_lib.lib().iic_bn_bwd_apply(_p(g_in), _p(act), _p(y), _p(mi), _p(gamma), _p(sums), _p(dy), _p(g_out),
                                    _p(dgamma), _p(dbeta), int(bool(accumulate)), dt, M, C, _stream()),
        "iic_bn_bwd_apply")
  return dy, g_out


def avgpool(x):
  n, h, w, C = x.shape
  feat = torch.empty((n, C), device=x.device, dtype=torch.float32)
  check(_lib.lib().iic_avgpool(_p(x), iic_dtype(x), _p(feat), n, h * w, C, _stream()), "iic_avgpool")
  return feat


def avgpool_bwd(dfeat, shape, dt):
  n, h, w, C = shape
  dx = torch.empty(shape, device=dfeat.device, dtype=_TORCH_DT[dt])
  check(_lib.lib().iic_avgpool_bwd(_p(dfeat), _p(dx), dt, n, h * w, C, _stream()), "iic_avgpool_bwd")
  return dx


# ---- heads ----------------------------------------------------------------------------------
def heads_fwd(feat, w, b, S, k):
  n, F = feat.shape
  logits = torch.empty((n, S * k), device=feat.device, dtype=torch.float32)
  z = torch.empty((S, n, k), device=feat.device, dtype=torch.float32)
  check(_lib.lib().iic_heads_fwd(_p(feat), _p(w), _p(b), _p(logits), _p(z), n, F, S, k, _stream()), "iic_heads_fwd")
  return z


def heads_bwd(feat, w, z, dz, S, k, want_dfeat):
  n, F = feat.shape
  dlog = torch.empty((n, S * k), device=feat.device, dtype=torch.float32)
  dw = torch.empty_like(w)
  db = torch.empty(S * k, device=feat.device, dtype=torch.float32)
  dfeat = torch.empty_like(feat) if want_dfeat else None
  check(_lib.lib().iic_heads_bwd(_p(feat), _p(w), _p(z), _p(dz), _p(dlog), _p(dw), _p(db), _p(dfeat), n, F, S, k,
                                 _stream()), "iic_heads_bwd")
  return dw, db, dfeat


# ---- optimiser ------------------------------------------------------------------------------
def adam_step(params, grads, exp_avgs, exp_avg_sqs, lr, beta1, beta2, eps, weight_decay, step):
  T = len(params)
  ptrs = (ctypes.c_void_p * (4 * T))()
  sizes = (ctypes.c_longlong * T)()
  for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avgs, exp_avg_sqs)):
    for t in (p, g, m, v):
      assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    ptrs[4 * i], ptrs[4 * i + 1], ptrs[4 * i + 2], ptrs[4 * i + 3] = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
    sizes[i] = p.numel()
  check(_lib.lib().iic_adam_step(ptrs, sizes, T, float(lr), float(beta1), float(beta2), float(eps),
                                 float(weight_decay), int(step), _stream()), "iic_adam_step")
