"""Oracle (CPU, test infrastructure only) for the clustering IIC objective.

Restates xu-ji/IIC ``code/utils/cluster/IID_losses.py``:
  * ``compute_joint``  -> reference :36-47
  * ``IID_loss``       -> reference :6-33
plus an independent closed-form fp64 numpy evaluation (loss and analytic
gradient) that the CUDA kernel's math mirrors.

The reference's forward runs unmodified on torch 2.x, but its backward does not
(``:17-19`` write in place into ``expand``-ed views).  The torch restatement
below materialises the two marginals with ``.clone()`` after the expand, which
is bit-identical in the forward pass (checked by tests/test_oracle_golden.py
against the unmodified reference) and differentiable.

Parity pinning: the reference has no golden vectors; this oracle is pinned
against tests/golden/iid_loss_*.npz generated from the reference itself.
"""
import sys

import numpy as np
import torch

EPS_DEFAULT = sys.float_info.epsilon


def compute_joint(x_out, x_tf_out):
  """P = normalise(symmetrise(sum_n z_n z'_n^T)); reference :36-47."""
  bn, k = x_out.shape
  assert x_tf_out.shape[0] == bn and x_tf_out.shape[1] == k
  outer = x_out[:, :, None] * x_tf_out[:, None, :]  # (bn, k, k), reference :42
  joint = outer.sum(dim=0)  # :43
  joint = (joint + joint.t()) / 2.  # :44
  joint = joint / joint.sum()  # :45
  return joint


def IID_loss(x_out, x_tf_out, lamb=1.0, EPS=EPS_DEFAULT):
  """(loss, loss_no_lamb); reference :6-33.  Inputs are softmax outputs."""
  k = x_out.shape[1]
  P = compute_joint(x_out, x_tf_out)
  assert P.shape == (k, k)
  # marginals are taken BEFORE the clamp (reference :12-14), then all three are
  # clamped independently (:17-19)
  p_row = P.sum(dim=1).view(k, 1).expand(k, k).clone()
  p_col = P.sum(dim=0).view(1, k).expand(k, k).clone()
  P = torch.where(P < EPS, torch.full_like(P, EPS), P)
  p_col = torch.where(p_col < EPS, torch.full_like(p_col, EPS), p_col)
  p_row = torch.where(p_row < EPS, torch.full_like(p_row, EPS), p_row)
  loss = (-P * (torch.log(P) - lamb * torch.log(p_col) - lamb * torch.log(p_row))).sum()
  loss_no_lamb = (-P * (torch.log(P) - torch.log(p_col) - torch.log(p_row))).sum()
  return loss, loss_no_lamb


def iid_loss_closed_form(z, zt, lamb=1.0, eps=EPS_DEFAULT, dtype=np.float64):
  """Closed-form loss + analytic gradients in numpy (no autograd).

  Returns dict(loss, loss_no_lamb, dz, dzt, joint) -- SURVEY.md S8 a8:
    A = Z^T Z'; s = sum A; P = (A + A^T) / (2 s)
    G = dloss/dP (with the reference's clamp semantics: an overwritten entry
        passes no gradient), dB = (G - <G,P>)/s, H = (dB + dB^T)/2,
    dZ = Z' H^T, dZ' = Z H.
  """
  z = np.asarray(z, dtype=dtype)
  zt = np.asarray(zt, dtype=dtype)
  eps = dtype(eps)
  A = z.T @ zt
  s = A.sum()
  P = (A + A.T) / 2. / s
  pi = P.sum(axis=1)
  pj = P.sum(axis=0)
  mP = P >= eps
  mi = pi >= eps
  mj = pj >= eps
  Pc = np.where(mP, P, eps)
  pic = np.where(mi, pi, eps)
  pjc = np.where(mj, pj, eps)
  lP, li, lj = np.log(Pc), np.log(pic), np.log(pjc)

  def one(lam):
    loss = -(Pc * (lP - lam * lj[None, :] - lam * li[:, None])).sum()
    G = mP * (-(lP - lam * lj[None, :] - lam * li[:, None]) - 1.0)
    G = G + lam * (mi * (Pc.sum(axis=1) / pic))[:, None]
    G = G + lam * (mj * (Pc.sum(axis=0) / pjc))[None, :]
    dB = (G - (G * P).sum()) / s
    H = (dB + dB.T) / 2.
    return loss, zt @ H.T, z @ H

  loss, dz, dzt = one(dtype(lamb))
  loss1, _, _ = one(dtype(1.0))
  return dict(loss=loss, loss_no_lamb=loss1, dz=dz, dzt=dzt, joint=P)
