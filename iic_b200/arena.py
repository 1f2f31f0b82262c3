"""Gradient arena: every ``.grad`` of a network is a view of ONE flat fp32 buffer.

Why ("bucketed and overlapped with backward"): the reference's ``DataParallel`` reduces the
replicas' gradients tensor by tensor after the backward; round 1 of this repo did ``torch.cat`` -> NCCL all-reduce ->
~110 ``copy_`` kernels on the compute stream after ``backward()``.  With the arena

  * ``zero_grad`` is one memset, and the trunk's backward (``_engine.TrunkFunction``) accumulates its weight gradients
    straight into the arena views (no autograd ``AccumulateGrad`` copy per parameter);
  * the parameters are laid out in the order the backward PRODUCES their gradients (layer4 first, stem last, the two
    head groups at the end), so a bucket is a contiguous slice that can be SUM all-reduced in place -- no cat, no copy --
    on a side stream as soon as its last gradient has been enqueued, while the backward of the earlier layers is still
    running on the compute stream; ``wait()`` makes the compute stream wait for the reductions before the optimiser.

Reference semantics kept: torch 0.4.1's ``zero_grad()`` zero-fills gradients that exist and leaves never-computed ones
``None`` (so Adam skips a head until it has been trained once, then keeps decaying its moments while the other head
trains: cluster_sobel_twohead.py:287,355).  Here a parameter is "live" once a gradient has been produced for it;
``FusedAdam`` steps live parameters only.
"""
import torch

from . import distributed


class GradArena(object):
  def __init__(self, net, bucket_bytes=16 << 20, align=32):
    named = list(net.named_parameters())
    heads = [(n, p) for n, p in named if not n.startswith("trunk.")]
    trunk = [(n, p) for n, p in named if n.startswith("trunk.")]
    order = list(reversed(trunk)) + list(reversed(heads))  # production order of the backward, head groups last
    assert order and all(p.dtype == torch.float32 for _, p in order), "GradArena: fp32 parameters only"
    dev = order[0][1].device
    # (a CPU network is accepted so that the bucket / liveness logic can be tested over gloo; reductions then run
    # synchronously)
    self.names, self.params, self.slices = [], [], []
    off = 0
    for n, p in order:
      self.names.append(n)
      self.params.append(p)
      self.slices.append((off, off + p.numel()))
      off += (p.numel() + align - 1) // align * align
    self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
    self._index = {}
    for i, (p, (a, b)) in enumerate(zip(self.params, self.slices)):
      p.grad = self.flat[a:b].view_as(p)
      self._index[id(p)] = i
      p._iic_arena = self  # _engine.GradSink looks this up
      p.register_post_accumulate_grad_hook(self._autograd_hook)
    # buckets: contiguous runs of parameters; the head groups get a bucket of their own (the head that is not being
    # trained never produces a gradient, its bucket is reduced by flush())
    self.buckets = []
    ntrunk = len(trunk)
    start, size = 0, 0
    for i in range(len(order)):
      size += (self.slices[i][1] - self.slices[i][0]) * 4
      last_of_group = (i == ntrunk - 1) or (i == len(order) - 1)
      if size >= bucket_bytes or last_of_group:
        self.buckets.append((start, i + 1))
        start, size = i + 1, 0
    self._bucket_of = {}
    for b, (lo, hi) in enumerate(self.buckets):
      for i in range(lo, hi):
        self._bucket_of[i] = b
    self.live = [False] * len(order)
    self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
    self._overlap = False
    self.profile = False  # record CUDA events around every bucket reduction (timeline(), bench.py --overlap-report)
    self._events = None
    self._reset_step()

  # ---- per-step protocol ------------------------------------------------------------------------
  def _reset_step(self):
    self._pending = [hi - lo for lo, hi in self.buckets]
    self._marked = [False] * len(self.params)
    self._launched = [False] * len(self.buckets)
    self._producers = set()  # streams other than the compute stream that wrote gradients in this step
    self.reduce_log = []  # bucket ids in launch order (tests)

  def begin_step(self, overlap=True):
    """zero_grad (one memset) and re-arm the buckets.  ``overlap``: reduce each bucket as soon as it is complete;
    only valid when every parameter receives its gradient from ONE backward node per step (pair-batched trunk)."""
    for p, (a, b) in zip(self.params, self.slices):
      if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * a:
        p.grad = self.flat[a:b].view_as(p)  # someone replaced / dropped the view (zero_grad(set_to_none=True), ...)
    self.flat.zero_()
    self._overlap = bool(overlap) and distributed.active()
    self._reset_step()
    if self.profile and self._side is not None:
      self._events = {"t0": torch.cuda.Event(enable_timing=True), "buckets": [], "bwd_end": None, "all_end": None}
      self._events["t0"].record()
    else:
      self._events = None

  def holds(self, p):
    i = self._index.get(id(p))
    return i is not None and p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * self.slices[i][0]

  def is_live(self, p):
    i = self._index.get(id(p))
    return True if i is None else self.live[i]

  def _autograd_hook(self, p):
    self.mark(p)

  def mark(self, p, producer=None):
    """The gradient of ``p`` for this step has been enqueued on the current stream (or on the stream ``producer``)."""
    i = self._index.get(id(p))
    if i is None:
      return
    if producer is not None:
      self._producers.add(producer)
    self.live[i] = True
    if self._marked[i]:
      return
    self._marked[i] = True
    b = self._bucket_of[i]
    self._pending[b] -= 1
    if self._pending[b] == 0 and self._overlap:
      self._reduce(b)

  def _reduce(self, b):
    if self._launched[b]:
      return
    self._launched[b] = True
    lo, hi = self.buckets[b]
    a, e = self.slices[lo][0], self.slices[hi - 1][1]
    self.reduce_log.append(b)
    if self._side is None:
      distributed.allreduce_sum_(self.flat[a:e])
      return
    cur = torch.cuda.current_stream(self.flat.device)
    self._side.wait_stream(cur)
    for ps in self._producers:
      self._side.wait_stream(ps)
    with torch.cuda.stream(self._side):
      if self._events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
      distributed.allreduce_sum_(self.flat[a:e])
      if self._events is not None:
        e1.record()
        self._events["buckets"].append((b, (e - a) * 4, e0, e1))

  def flush(self):
    """After backward(): reduce whatever has not been reduced yet (same order on every rank) and make the compute
    stream wait for all reductions."""
    if not distributed.active():
      return
    if self._events is not None:  # the backward has been enqueued up to here on the compute stream
      self._events["bwd_end"] = torch.cuda.Event(enable_timing=True)
      self._events["bwd_end"].record()
    for b in range(len(self.buckets)):
      self._reduce(b)
    if self._side is not None:
      torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
    if self._events is not None:
      self._events["all_end"] = torch.cuda.Event(enable_timing=True)
      self._events["all_end"].record()

  def timeline(self):
    """After a profiled step (``profile = True``) and a device synchronise: milliseconds since begin_step of every bucket
    reduction on the side stream, of the end of the backward on the compute stream, and of the point where the compute
    stream has the reduced gradients.  ``exposed_ms`` = all-reduce time the backward did not hide."""
    ev = self._events
    if ev is None or ev["bwd_end"] is None:
      return None
    t0 = ev["t0"]
    rows = [{"bucket": b, "bytes": nbytes, "start_ms": t0.elapsed_time(e0), "end_ms": t0.elapsed_time(e1)}
            for b, nbytes, e0, e1 in ev["buckets"]]
    bwd_end, all_end = t0.elapsed_time(ev["bwd_end"]), t0.elapsed_time(ev["all_end"])
    return {"buckets": rows, "backward_end_ms": bwd_end, "gradients_ready_ms": all_end, "exposed_ms": max(0.0, all_end - bwd_end),
            "allreduce_busy_ms": sum(r["end_ms"] - r["start_ms"] for r in rows)}

  def grad_checksum(self):
    """fp64 sum and sum of squares of the whole arena (bench.py --verify)."""
    f = self.flat.double()
    return float(f.sum()), float((f * f).sum())
