"""The WELL-CONDITIONED end-to-end fixture behind the stated tolerances of the tensor-core modes
(tests/test_gpu_precision.py, tools/precision_probe.py, __graft_entry__.smoke()).

Round 1's net goldens are a randomly-initialised ClusterNet5g at batch 3-6 whose clustering has collapsed (every image
lands in the same cluster, MI ~ 1e-3): there the loss gradient is the difference of nearly equal terms and any rounding
is amplified without bound, so they cannot carry a tolerance for a reduced-precision mode.  This fixture removes the
two causes without shipping trained weights (21 M parameters cannot travel in a git fixture):

  * batch >= 32 pairs per view, so BatchNorm statistics are estimates, not noise;
  * the sub-head biases are centred on the mean trunk feature (b = -W f_mean, f_mean from the fp32 oracle, stored in
    the golden file) and the head weights have gain 200 (N(0, 0.4)): the random-feature clustering is then spread
    over all k clusters and consistent between the views (x_tf = x + 0.03 noise), MI ~ 0.5 nats -- the regime of a
    network that has started to train (reference logs: loss -0.5 ... -2.2).

Everything is rebuilt from names (oracle/weights.py), so the unmodified reference (tests/golden/make_golden.py
gen_precision), the CPU oracle and the CUDA path all see bit-identical parameters and inputs."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from oracle import iid_losses as oracle_iid  # noqa: E402
from oracle import nets as oracle_nets  # noqa: E402
from oracle import weights  # noqa: E402

HEAD_GAIN = 200.0
NOISE = 0.03


def config(sz):
  return dict(in_channels=2, input_sz=sz, num_sub_heads=5, output_k_A=70, output_k_B=10, batchnorm_track=True)


def inputs(sz, pairs):
  x = weights.normal("wc.x", (pairs, 2, sz, sz))
  return x, x + NOISE * weights.normal("wc.xt", tuple(x.shape))


def prepare(net, sz, pairs, head, trunk_mean=None):
  """Named deterministic parameters + centred head biases.  ``net``: any module with the reference's parameter names
  (reference module, oracle, or the iic_b200 network BEFORE .cuda()).  ``trunk_mean`` [512]: mean trunk feature of
  view x in train mode; computed with ``net`` itself (CPU, fp32) when None.  Returns trunk_mean."""
  weights.fill_state_dict(net, head_gain=HEAD_GAIN)
  if trunk_mean is None:
    x, _ = inputs(sz, pairs)
    was = net.training
    net.train()
    with torch.no_grad():
      trunk_mean = net(x, trunk_features=True).mean(0)
    net.train(was)
    for m in net.modules():  # undo the running-statistics update of the probe forward
      if isinstance(m, torch.nn.BatchNorm2d) and m.track_running_stats:
        m.reset_running_stats()
  with torch.no_grad():
    for h in (net.head_A if head == "A" else net.head_B).heads:
      h[0].bias.copy_(-(h[0].weight @ trunk_mean.to(h[0].weight.dtype)))
  return trunk_mean


def result(loss, outs, net):
  return {"loss": float(loss), "out": torch.stack(list(outs)).detach().double().cpu(),
          "grads": {n: p.grad.detach().double().cpu() for n, p in net.named_parameters() if p.grad is not None}}


def compare(got, ref):
  """Error summary of `got` against `ref` (both from result())."""
  rel, cos = {}, {}
  num = den = 0.0
  for n, g in ref["grads"].items():
    a = got["grads"][n]
    d = float((a - g).norm())
    rel[n] = d / (float(g.norm()) + 1e-300)
    cos[n] = float((a.flatten() @ g.flatten()) / (a.norm() * g.norm() + 1e-300))
    num += d * d
    den += float(g.norm()) ** 2
  worst = sorted(rel, key=rel.get)[-3:]
  return {"loss_rel": abs(got["loss"] - ref["loss"]) / abs(ref["loss"]), "loss": got["loss"],
          "out_max_abs": float((got["out"] - ref["out"]).abs().max()),
          "grad_rel_l2_total": (num / den) ** 0.5, "grad_rel_l2_median": float(np.median(list(rel.values()))),
          "grad_rel_l2_max": max(rel.values()), "grad_cos_min": min(cos.values()),
          "grad_cos_median": float(np.median(list(cos.values()))),
          "worst": {n: round(rel[n], 4) for n in worst}}


class Fixture(object):
  def __init__(self, sz=32, pairs=64, head="B", trunk_mean=None):
    self.sz, self.pairs, self.head = sz, pairs, head
    self.x, self.xt = inputs(sz, pairs)
    self.trunk_mean = trunk_mean

  def describe(self):
    return {"net": "ClusterNet5gTwoHead", "input_sz": self.sz, "pairs": self.pairs, "head": self.head,
            "head_gain": HEAD_GAIN, "view_noise": NOISE}

  def _oracle_net(self, dtype=torch.float32):
    net = oracle_nets.ClusterNet5gTwoHead(Namespace(**config(self.sz)))
    self.trunk_mean = prepare(net, self.sz, self.pairs, self.head, self.trunk_mean)
    return net.to(dtype).train()

  def oracle(self, dtype=torch.float32, rounding=False):
    net = self._oracle_net(dtype)
    with oracle_nets.bf16_rounding(rounding):
      o, ot = net(self.x.to(dtype), head=self.head), net(self.xt.to(dtype), head=self.head)
      loss = sum(oracle_iid.IID_loss(a, b)[0] for a, b in zip(o, ot)) / len(o)
      loss.backward()
    return result(loss.item(), o, net)

  def cuda_net(self, precision):
    import iic_b200.archs as archs
    if self.trunk_mean is None:
      self._oracle_net()
    net = archs.ClusterNet5gTwoHead(Namespace(precision=precision, **config(self.sz)))
    prepare(net, self.sz, self.pairs, self.head, self.trunk_mean)
    return net.cuda().train()

  def cuda(self, precision):
    from iic_b200.utils.cluster.IID_losses import IID_loss
    net = self.cuda_net(precision)
    o, ot = net(self.x.cuda(), head=self.head), net(self.xt.cuda(), head=self.head)
    loss = sum(IID_loss(a, b)[0] for a, b in zip(o, ot)) / len(o)
    loss.backward()
    torch.cuda.synchronize()
    return result(loss.item(), o, net)

  def cuda_trajectory(self, precision, steps, lr=1e-4):
    """`steps` Adam steps of iic_cluster_step on the fixture batch (sobel off: the fixture is the network input)."""
    from iic_b200.optim import FusedAdam
    from iic_b200.step import iic_cluster_step
    net = self.cuda_net(precision)
    opt = FusedAdam(net.parameters(), lr=lr)
    x, xt = self.x.cuda(), self.xt.cuda()
    losses = [iic_cluster_step(net, opt, x, xt, head=self.head, sobel=False)[0] for _ in range(steps)]
    return [float(l) for l in losses]
