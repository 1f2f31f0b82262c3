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
    assert closure is None
    for group in self.param_groups:
      ps, gs, ms, vs = [], [], [], []
      step = None
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
        st["step"] += 1
        s = int(st["step"].item())
        if step is None:
          step = s
        if s != step:  # parameters that joined later: separate launch group
          kernels.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                            group["weight_decay"], step)
          ps, gs, ms, vs, step = [], [], [], [], s
        ps.append(p)
        gs.append(p.grad.contiguous())
        ms.append(st["exp_avg"])
        vs.append(st["exp_avg_sq"])
      if ps:
        kernels.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                          group["weight_decay"], step)


def zero_grad_like_reference(net):
  """torch 0.4.1 ``Module.zero_grad()`` (the call at cluster_sobel_twohead.py:287): gradients that exist are detached
  and zero-filled, gradients that were never computed stay ``None``.  Same as ``zero_grad(set_to_none=False)`` today."""
  for p in net.parameters():
    if p.grad is not None:
      p.grad.detach_()
      p.grad.zero_()
