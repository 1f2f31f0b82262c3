"""Deterministic, platform-independent tensors addressed by name.

Golden fixtures cannot ship 21M-parameter state dicts, so both the golden
generator (which drives the real reference modules) and the tests (which drive
the oracle and the CUDA path) rebuild identical parameters from a name and a
shape.  numpy's PCG64 stream is stable for a given numpy version, unlike
``torch.randn`` whose CPU kernels may vectorise differently per ISA.
"""
import zlib

import numpy as np
import torch


def _rng(name, salt=0):
  return np.random.Generator(np.random.PCG64(zlib.crc32(name.encode()) + 7919 * salt))


def normal(name, shape, std=1.0, mean=0.0, salt=0, dtype=np.float32):
  a = _rng(name, salt).standard_normal(size=tuple(shape)) * std + mean
  return torch.from_numpy(a.astype(dtype))


def uniform(name, shape, lo=0.0, hi=1.0, salt=0, dtype=np.float32):
  a = _rng(name, salt).random(size=tuple(shape)) * (hi - lo) + lo
  return torch.from_numpy(a.astype(dtype))


def fill_state_dict(module, salt=0, head_gain=50.0):
  """Overwrite every parameter of ``module`` with named deterministic values.

  conv weights: N(0, sqrt(2/fan_out)) (the reference's kaiming fan_out init,
  code/archs/cluster/residual.py:75-78); BN gamma ~ U[0.5,1.5], beta ~
  N(0,0.1) (non-trivial so dgamma/dbeta are exercised); Linear / 1x1-head
  weights N(0, 0.01*head_gain) -- SURVEY S7 step 0: the reference's N(0,0.01)
  head init gives a uniform softmax and an uninformative loss, so parity runs
  scale it.  Buffers (running stats) are left at their defaults.
  """
  with torch.no_grad():
    for name, p in module.named_parameters():
      if p.dim() == 4 and "head" not in name:
        fan_out = p.shape[0] * p.shape[2] * p.shape[3]
        v = normal(name, p.shape, std=(2.0 / fan_out) ** 0.5, salt=salt)
      elif p.dim() == 4:  # segmentation 1x1 head conv
        v = normal(name, p.shape, std=0.01 * head_gain, salt=salt)
      elif p.dim() == 2:
        v = normal(name, p.shape, std=0.01 * head_gain / 5.0, salt=salt)
      elif name.endswith("weight"):
        v = uniform(name, p.shape, 0.5, 1.5, salt=salt)
      else:
        v = normal(name, p.shape, std=0.1, salt=salt)
      p.copy_(v.to(p.dtype))
  return module
