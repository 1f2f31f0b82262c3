"""2-GPU (NCCL) test of the sharded hot path: per-rank BN, [S,k,k] joint all-reduce inside IID_loss,
SUM all-reduce of the weight gradients.  The 2-rank result must equal the one-device emulation of the
same algorithm (chunked forwards + summed PARTIAL joints), which test_gpu_parity_nets.py pins to the
oracle with per-chunk BatchNorm.  Skipped unless >= 2 GPUs are visible (gpurun --gpus 2)."""
import os
import socket
import sys
import tempfile
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import weights  # noqa: E402

CFG = dict(in_channels=2, input_sz=32, num_sub_heads=3, output_k_A=12, output_k_B=6, batchnorm_track=True)


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _build(dev):
  import iic_b200.archs as archs
  net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **CFG))
  weights.fill_state_dict(net, salt=9)
  return net.to(dev).train()


def _data():
  x = weights.normal("multi.x", (12, 2, 32, 32))
  return x, x + 0.3 * weights.normal("multi.xt", (12, 2, 32, 32))


def _worker(rank, world, port, outdir, use_arena):
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  from iic_b200 import distributed as iicd
  from iic_b200.utils.cluster.IID_losses import IID_loss
  iicd.enable()
  net = _build(dev)
  x, xt = _data()
  per = x.shape[0] // world
  xs, xts = x[rank * per:(rank + 1) * per].to(dev), xt[rank * per:(rank + 1) * per].to(dev)
  if use_arena:
    # the path bench.py times: flat in-place gradients, buckets all-reduced on a side stream during the backward
    from iic_b200.arena import GradArena
    from iic_b200.step import iic_cluster_step
    arena = GradArena(net, bucket_bytes=4 << 20)
    loss, _ = iic_cluster_step(net, None, xs, xts, head="A", lamb=1.2, sobel=False, arena=arena)
    assert len(arena.reduce_log) == len(arena.buckets) > 3 and arena.reduce_log[-1] == len(arena.buckets) - 1
  else:
    o, ot = net(xs, head="A"), net(xts, head="A")
    loss = sum(IID_loss(a, b, lamb=1.2)[0] for a, b in zip(o, ot)) / len(o)
    loss.backward()
    iicd.allreduce_gradients(net.parameters())
  torch.cuda.synchronize()
  torch.save({"loss": loss.item(), "grads": {n: p.grad.cpu() for n, p in net.named_parameters() if p.grad is not None}},
             os.path.join(outdir, "rank%d.pt" % rank))  # (the idle head's arena gradients are exact zeros)
  dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("use_arena", [False, True], ids=["cat-allreduce", "arena-overlapped"])
def test_two_ranks_match_single_device_emulation(use_arena):
  import torch.multiprocessing as mp
  from iic_b200 import _lib, kernels
  world = 2
  with tempfile.TemporaryDirectory() as d:
    mp.spawn(_worker, args=(world, _free_port(), d, use_arena), nprocs=world, join=True)
    res = [torch.load(os.path.join(d, "rank%d.pt" % r)) for r in range(world)]
  # all ranks hold the same loss and, after the all-reduce, the same gradients
  assert abs(res[0]["loss"] - res[1]["loss"]) < 1e-6
  for n in res[0]["grads"]:
    assert torch.allclose(res[0]["grads"][n], res[1]["grads"][n], rtol=1e-5, atol=1e-8), n
  # one-device emulation
  dev = torch.device("cuda", 0)
  net = _build(dev)
  x, xt = _data()
  per = x.shape[0] // world
  zs = [net.forward_stacked(x[r * per:(r + 1) * per].to(dev), head="A") for r in range(world)]
  zts = [net.forward_stacked(xt[r * per:(r + 1) * per].to(dev), head="A") for r in range(world)]
  S, k = 3, 12
  joint = torch.zeros(S, k, k, device=dev)
  for z, zt in zip(zs, zts):
    j = torch.empty_like(joint)
    kernels.iid_loss(z.detach().contiguous(), zt.detach().contiguous(), 1.2, sys.float_info.epsilon, False,
                     phase=_lib.PHASE_PARTIAL, joint_ws=j)
    joint += j
  for z, zt in zip(zs, zts):
    loss, dz, dzt, _ = kernels.iid_loss(z.detach().contiguous(), zt.detach().contiguous(), 1.2, sys.float_info.epsilon,
                                        True, phase=_lib.PHASE_FINISH, joint_ws=joint)
    torch.autograd.backward([z, zt], [dz / S, dzt / S])
  assert abs(loss[:, 0].mean().item() - res[0]["loss"]) < 2e-6
  for n, p in net.named_parameters():
    if p.grad is None:
      assert n not in res[0]["grads"] or float(res[0]["grads"][n].abs().max()) == 0.0, n
      continue
    want, got = p.grad.cpu(), res[0]["grads"][n]
    assert ((got - want).norm() / (want.norm() + 1e-30)).item() < 2e-3, n
