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


def allreduce_sum_(t):
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
