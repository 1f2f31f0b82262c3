"""Multi-GPU plumbing for the sharded hot path.

One process per GPU (torchrun); rank r owns a contiguous slice of the image
pairs -- the same chunking ``torch.nn.DataParallel`` used in the reference
(code/scripts/cluster/cluster_sobel_twohead.py:180-181), so BatchNorm statistics
stay per-rank exactly as they were per-replica.  Two exchanges exist:
  1. the [S,k,k] un-normalised joint is SUM all-reduced before normalisation
     (inside IID_loss, between the PARTIAL and FINISH kernel phases);
  2. weight gradients are SUM all-reduced (the loss is already globally
     normalised, so it is a sum, not a mean) -- ``allreduce_gradients``.
torch.distributed (NCCL over NVLink/NVSwitch) carries both.
"""
import torch
import torch.distributed as dist

_group = {"enabled": False, "pg": None}


def enable(process_group=None):
  """Turn on the cross-rank joint all-reduce inside the loss functions."""
  assert dist.is_available() and dist.is_initialized(), "init torch.distributed first"
  _group["enabled"] = True
  _group["pg"] = process_group


def disable():
  _group["enabled"] = False
  _group["pg"] = None


def active():
  return _group["enabled"] and dist.get_world_size(_group["pg"]) > 1


def group():
  return _group["pg"]


_checked_shapes = set()


def allreduce_sum_(t):
  """In-place SUM over the ranks.  The first time a shape is seen, the ranks compare it (one small all-gather): a rank
  whose (S, k) differs -- a different config, a ragged last batch routed to another head -- would otherwise hang or
  silently corrupt the NCCL collective."""
  key = (tuple(t.shape), t.dtype)
  if key not in _checked_shapes:
    mine = torch.tensor([t.dim()] + list(t.shape) + [0] * (7 - t.dim()), dtype=torch.int64, device=t.device)
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size(_group["pg"]))]
    dist.all_gather(every, mine, group=_group["pg"])
    shapes = [tuple(int(v) for v in e[1:1 + int(e[0])]) for e in every]
    if any(sh != shapes[0] for sh in shapes):
      raise RuntimeError("iic_b200.distributed: ranks disagree on the shape of a summed tensor: %s" % (shapes,))
    _checked_shapes.add(key)
  dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group["pg"])
  return t


def allreduce_gradients(params, bucket_bytes=64 << 20):
  """SUM all-reduce of .grad over all ranks, coalesced into flat buckets."""
  if not active():
    return
  grads = [p.grad for p in params if p.grad is not None]
  bucket, size = [], 0

  def flush():
    if not bucket:
      return
    flat = torch.cat([g.reshape(-1) for g in bucket])
    allreduce_sum_(flat)
    off = 0
    for g in bucket:
      n = g.numel()
      g.copy_(flat[off:off + n].view_as(g))
      off += n

  for g in grads:
    bucket.append(g)
    size += g.numel() * g.element_size()
    if size >= bucket_bytes:
      flush()
      bucket, size = [], 0
  flush()


def emulate_sharded_backward(net, xs, xts, head="B", lamb=1.0, sobel=False):
  """One-device emulation of the W-rank sharded step (SURVEY.md S8e "Equivalence check"): the chunks ``xs[r]`` /
  ``xts[r]`` are pushed through ``net`` one after the other (so BatchNorm statistics are per chunk, as they are per
  rank), the PARTIAL joints are summed, every chunk is FINISHed against the global joint and backpropagated -- the
  parameter gradients accumulate to what the SUM all-reduce produces.  Returns the (global) mean loss.
  Used by tests/test_gpu_multi.py and ``bench.py --verify``."""
  import sys

  from . import _lib, kernels
  from .step import _to_net_input
  assert not active(), "emulation runs on one device without a process group"
  zs, zts = [], []
  for x, xt in zip(xs, xts):
    x, xt = _to_net_input(x, sobel, False), _to_net_input(xt, sobel, False)
    if hasattr(net, "forward_stacked_pair"):
      z, zt = net.forward_stacked_pair(x, xt, head=head)
    else:
      z, zt = net.forward_stacked(x, head=head), net.forward_stacked(xt, head=head)
    zs.append(z)
    zts.append(zt)
  S, _, k = zs[0].shape
  eps = sys.float_info.epsilon
  joint = torch.zeros(S, k, k, device=zs[0].device)
  for z, zt in zip(zs, zts):
    j = torch.empty_like(joint)
    kernels.iid_loss(z.detach().contiguous(), zt.detach().contiguous(), lamb, eps, False, phase=_lib.PHASE_PARTIAL,
                     joint_ws=j)
    joint += j
  loss = None
  for z, zt in zip(zs, zts):
    loss, dz, dzt, _ = kernels.iid_loss(z.detach().contiguous(), zt.detach().contiguous(), lamb, eps, True,
                                        phase=_lib.PHASE_FINISH, joint_ws=joint)
    torch.autograd.backward([z, zt], [dz / S, dzt / S])
  return loss[:, 0].mean()
