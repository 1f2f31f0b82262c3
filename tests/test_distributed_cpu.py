"""CPU, world_size=2 over gloo: the host-side logic of the N>1 path.

  * iic_b200.distributed: enable/active, SUM all-reduce of the joint, bucketed gradient all-reduce;
  * the sharding algebra the CUDA phases implement (PARTIAL joint per rank -> SUM all-reduce ->
    FINISH with local rows) reproduces the single-process loss and per-row gradients -- evaluated
    here with the oracle's closed form as the per-rank math (the CUDA kernels themselves are
    covered by tests/test_gpu_parity_iid.py::test_phases_equal_fused_and_shard)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from iic_b200 import distributed as iicd
    from oracle import iid_losses as oi
    from oracle import weights
    assert not iicd.active()
    iicd.enable()
    assert iicd.active()
    # ---- gradient all-reduce: SUM (not mean), bucket boundaries must not matter ----
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in [(3, 5), (7,), (2, 2, 2), (11,)]]
    for i, p in enumerate(ps):
      p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    ps.append(torch.nn.Parameter(torch.zeros(4)))  # no grad: skipped
    iicd.allreduce_gradients(ps, bucket_bytes=64)
    for i, p in enumerate(ps[:-1]):
      assert torch.all(p.grad == 3.0 * (i + 1)), (i, p.grad)
    assert ps[-1].grad is None
    # ---- sharded joint: rank r owns rows [r*n/W, (r+1)*n/W) (DataParallel's contiguous scatter) ----
    n, k, lamb = 64, 6, 1.3
    l = weights.normal("dist.l", (n, k))
    z = torch.softmax(2 * l, 1).double().numpy()
    zt = torch.softmax(2 * l + weights.normal("dist.lt", (n, k)), 1).double().numpy()
    full = oi.iid_loss_closed_form(z, zt, lamb=lamb)
    lo, hi = rank * n // world, (rank + 1) * n // world
    joint = torch.from_numpy(z[lo:hi].T @ zt[lo:hi])  # PARTIAL
    iicd.allreduce_sum_(joint)                        # exchange
    A = joint.numpy()
    # FINISH: every rank evaluates the MI from the global joint and the gradient of ITS rows
    s = A.sum()
    P = (A + A.T) / 2 / s
    eps = np.finfo(np.float64).eps
    pi, pj = P.sum(1), P.sum(0)
    Pc = np.maximum(P, eps)
    loss = -(Pc * (np.log(Pc) - lamb * np.log(np.maximum(pj, eps))[None, :] - lamb * np.log(np.maximum(pi, eps))[:, None])).sum()
    G = (P >= eps) * (-(np.log(Pc) - lamb * np.log(pj)[None, :] - lamb * np.log(pi)[:, None]) - 1.0)
    G = G + lamb * (Pc.sum(1) / pi)[:, None] + lamb * (Pc.sum(0) / pj)[None, :]
    H = (G - (G * P).sum()) / s
    H = (H + H.T) / 2
    dz_local = zt[lo:hi] @ H.T
    assert abs(loss - full["loss"]) < 1e-12
    assert np.abs(dz_local - full["dz"][lo:hi]).max() < 1e-12
    # ---- ranks that disagree on the shape of a summed tensor get an error instead of a hang ----
    bad = torch.zeros(1, 3 + rank, 3 + rank)
    try:
      iicd.allreduce_sum_(bad)
      raise AssertionError("shape disagreement went unnoticed")
    except RuntimeError as e:
      assert "disagree" in str(e), e
    iicd.disable()
    assert not iicd.active()
    q.put((rank, "ok"))
  except Exception as e:  # pragma: no cover
    q.put((rank, "FAIL %r" % (e,)))
  finally:
    dist.destroy_process_group()


def test_world_size_2_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, "ok"), (1, "ok")], res


class _TinyTwoHead(torch.nn.Module):
  """trunk.* / head_A.* / head_B.* parameter names like the real networks (what GradArena keys its layout on)."""

  def __init__(self):
    super().__init__()
    self.trunk = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                     torch.nn.Linear(16, 12))
    self.head_A = torch.nn.Linear(12, 5)
    self.head_B = torch.nn.Linear(12, 3)

  def forward(self, x, head):
    return (self.head_A if head == "A" else self.head_B)(self.trunk(x))


def _arena_worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from iic_b200 import distributed as iicd
    from iic_b200.arena import GradArena
    iicd.enable()
    torch.manual_seed(0)
    net = _TinyTwoHead()
    ref = _TinyTwoHead()
    ref.load_state_dict(net.state_dict())
    arena = GradArena(net, bucket_bytes=600)
    # layout: trunk parameters in the order the backward produces them (last layer first), the head groups at the end
    assert arena.names[0] == "trunk.4.bias" and arena.names[5] == "trunk.0.weight", arena.names
    assert [n.split(".")[0] for n in arena.names[6:]] == ["head_B", "head_B", "head_A", "head_A"], arena.names
    assert len(arena.buckets) >= 3 and arena.buckets[-1] == (6, 10), arena.buckets
    torch.manual_seed(100)
    xs = torch.randn(2 * world, 8)
    for step, head in enumerate(["B", "B", "A"]):
      arena.begin_step(overlap=True)
      assert all(float(p.grad.abs().sum()) == 0 for p in net.parameters())
      net(xs[2 * rank:2 * rank + 2], head).pow(2).sum().backward()
      # the trunk buckets were reduced DURING the backward, in production order; the head bucket waits for flush()
      assert arena.reduce_log == list(range(len(arena.buckets) - 1)), arena.reduce_log
      arena.flush()
      assert arena.reduce_log == list(range(len(arena.buckets)))
      ref.zero_grad()
      ref(xs, head).pow(2).sum().backward()  # all ranks' rows on one process: the SUM of the per-rank gradients
      for (n, p), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
        want = torch.zeros_like(p) if r.grad is None else r.grad
        assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-6), (step, n)
        assert p.grad.data_ptr() == arena.flat.data_ptr() + 4 * arena.slices[arena.names.index(n)][0]
      live_a = arena.is_live(net.head_A.weight)
      assert live_a == (step == 2) and arena.is_live(net.head_B.weight) and arena.is_live(net.trunk[0].weight)
    # a dropped view (zero_grad(set_to_none=True)) is restored by begin_step
    net.zero_grad(set_to_none=True)
    arena.begin_step(overlap=False)
    assert arena.holds(net.trunk[0].weight)
    net(xs[2 * rank:2 * rank + 2], "A").sum().backward()
    assert arena.reduce_log == []  # overlap off: nothing before flush
    arena.flush()
    assert arena.reduce_log == list(range(len(arena.buckets)))
    iicd.disable()
    q.put((rank, "ok"))
  except Exception as e:  # pragma: no cover
    import traceback
    q.put((rank, "FAIL %r %s" % (e, traceback.format_exc())))
  finally:
    dist.destroy_process_group()


def test_grad_arena_world_size_2_gloo():
  """GradArena (iic_b200/arena.py): layout, liveness, in-place bucketed SUM all-reduce launched during the backward."""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_arena_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, "ok"), (1, "ok")], res
