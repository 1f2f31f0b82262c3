"""One training step as a CUDA graph.

A step of the hot path is ~310 kernel launches of this library plus a few dozen torch plumbing ops, enqueued from Python
through ctypes: ~4 ms of host time.  At the benchmark's weak-scaling batch (704 pairs per GPU, 43 ms of GPU work) that
hides behind the GPU; at the strong-scaling point of BASELINE.json (704 pairs over 8 GPUs = 88 per GPU, ~6 ms) and on
the small networks (ClusterNet6c 24x24) the step is launch bound.  ``GraphedStep`` captures ONE step -- zero_grad
(arena memset), sobel, trunk, heads, loss, backward, bucketed all-reduces on the side stream, fused Adam -- and replays
it with one ``cudaGraphLaunch``.  What makes the step capturable:

  * all gradients live in a ``GradArena`` (static addresses, no autograd allocation of .grad);
  * ``FusedAdam.graph_safe``: step counts in device memory, bias corrections computed by the kernel;
  * inputs are copied into static device buffers before each replay; the loss scalars are static outputs;
  * a graph is specific to (head, batch shape): capture one per head, as the reference alternates heads per epoch.
"""
import torch

from .step import iic_cluster_step, iic_seg_step


class GraphedStep(object):
  def __init__(self, net, optimiser, arena, example_batch, kind="cluster", warmup=3, **step_kwargs):
    assert arena is not None, "GraphedStep needs a GradArena (static gradient addresses)"
    self.net, self.opt, self.arena, self.kw = net, optimiser, arena, dict(step_kwargs)
    self.fn = iic_cluster_step if kind == "cluster" else iic_seg_step
    dev = next(net.parameters()).device
    self.static = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example_batch]
    for dst, src in zip(self.static, example_batch):
      dst.copy_(src)
    optimiser.graph_safe = True
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for _ in range(warmup):  # lazily created state (Adam moments, pack plans, scratch, NCCL communicators) must exist
        self.fn(net, optimiser, *self.static, arena=arena, **self.kw)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    from . import kernels
    self.graph = torch.cuda.CUDAGraph()
    c0 = kernels.launch_count()
    with torch.cuda.graph(self.graph):
      self.out = self.fn(net, optimiser, *self.static, arena=arena, **self.kw)
    # kernels of this library inside the graph (the host-side launch counter does not see replays)
    self.launches_per_replay = kernels.launch_count() - c0
    self.warmup_steps = warmup  # these were real optimiser steps

  def __call__(self, *batch):
    """batch: host (pinned) or device tensors with the shapes of the example batch.  Returns the (static) loss tensors."""
    for dst, src in zip(self.static, batch):
      dst.copy_(src, non_blocking=True)
    self.graph.replay()
    self.opt.note_replay()
    return self.out
