"""Fused multi-tensor Adam: drop-in for ``torch.optim.Adam(params, lr=...)`` as built by the
reference (code/utils/cluster/general.py:8-9, stepped at cluster_sobel_twohead.py:355).

One kernel launch per 48 parameter tensors instead of ~5 per tensor.  ``state_dict()`` uses
torch's layout (``step``, ``exp_avg``, ``exp_avg_sq``) so optimiser checkpoints interchange.

``zero_grad_like_reference(net)`` (a function, below) reproduces torch 0.4.1's ``zero_grad()``
(zero-fill instead of None): the head that is not being trained still receives Adam moment
decay and parameter updates, as in the reference.  ``iic_b200.step`` defaults to
that behaviour (``set_to_none=False``).

Optimiser checkpoints written by the reference (torch 0.4.1) keep ``state['step']`` as a Python
int and carry an ``amsgrad`` key in the param groups; both are accepted (``__setstate__``)."""
import torch

from . import kernels


class FusedAdam(torch.optim.Optimizer):
  def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
    # With a GradArena every parameter has a (zero-filled) .grad from the start; `grad_filter(p) -> bool` then plays
    # the part of torch's `p.grad is None` test (a parameter that never received a gradient is not stepped).
    self.grad_filter = None

  def __setstate__(self, state):
    # what torch.optim.Adam.__setstate__ does: legacy checkpoints (torch < 1.12, incl. the reference's 0.4.1) store
    # `step` as a Python int
    super().__setstate__(state)
    for group in self.param_groups:
      group.setdefault("amsgrad", False)
      assert not group["amsgrad"], "FusedAdam: amsgrad is not implemented (the reference never enables it)"
    for st in self.state.values():
      if len(st) != 0 and not torch.is_tensor(st["step"]):
        st["step"] = torch.tensor(float(st["step"]))

  @torch.no_grad()
  def step(self, closure=None):
    """``self.graph_safe = True``: the step count of every launch group also lives in a device scalar that is incremented
    on the stream, and the kernel computes the bias corrections from it -- the enqueued work is then identical from step
    to step and can be captured into a CUDA graph (iic_b200/graph.py).  While a capture is in progress the host-side
    ``state['step']`` counters are NOT advanced (nothing executes); ``note_replay()`` advances them per replay."""
    assert closure is None
    graph = getattr(self, "graph_safe", False)
    capturing = graph and torch.cuda.is_current_stream_capturing()
    self._last_states = []
    for group in self.param_groups:
      buckets = {}  # step value -> (params, grads, exp_avgs, exp_avg_sqs); parameters that joined later step separately
      for p in group["params"]:
        if p.grad is None or (self.grad_filter is not None and not self.grad_filter(p)):
          continue
        st = self.state[p]
        if len(st) == 0:
          st["step"] = torch.tensor(0.0)
          st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
          st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if not torch.is_tensor(st["step"]):  # state assigned by hand after load_state_dict
          st["step"] = torch.tensor(float(st["step"]))
        s = int(st["step"].item()) + 1
        if not capturing:
          st["step"] += 1
        self._last_states.append(st)
        b = buckets.setdefault(s, ([], [], [], []))
        b[0].append(p)
        b[1].append(p.grad.contiguous())
        b[2].append(st["exp_avg"])
        b[3].append(st["exp_avg_sq"])
      for s, (ps, gs, ms, vs) in buckets.items():
        step_dev = None
        if graph:
          key = tuple(id(p) for p in ps)
          store = self.__dict__.setdefault("_dev_steps", {})
          step_dev = store.get(key)
          if step_dev is None:
            step_dev = store[key] = torch.full((1,), float(s - 1), device=ps[0].device, dtype=torch.float32)
          step_dev.add_(1.0)  # on the stream: executed eagerly, recorded under capture
        kernels.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                          group["weight_decay"], s, step_dev=step_dev)

  def note_replay(self):
    """Host-side bookkeeping for one replay of a captured step: advance the ``state['step']`` counters of the
    parameters the captured step updates (the device-side counters advance inside the graph)."""
    for st in self._last_states:
      st["step"] += 1


def zero_grad_like_reference(net):
  """torch 0.4.1 ``Module.zero_grad()`` (the call at cluster_sobel_twohead.py:287): gradients that exist are detached
  and zero-filled, gradients that were never computed stay ``None``.  Same as ``zero_grad(set_to_none=False)`` today."""
  for p in net.parameters():
    if p.grad is not None:
      p.grad.detach_()
      p.grad.zero_()
