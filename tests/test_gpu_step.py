"""The functions bench.py times, checked end to end (VERDICT r1 "test what you benchmark"):
``iic_b200.step.iic_cluster_step`` / ``iic_seg_step`` (zero_grad -> sobel -> net x2 -> loss -> backward -> Adam) and
``FusedAdam`` as a class, against the oracle networks stepped by ``torch.optim.Adam`` on the CPU
(code/scripts/cluster/cluster_sobel_twohead.py:286-355, code/scripts/segmentation/segmentation_twohead.py:262-361).

fp32 mode (reference precision; the engine and the glue are shared with the tensor-core modes).  Tolerances: the loss
of the first step within 2e-5 abs, of later steps within 5e-4; after the steps, parameter *updates* within 2 % relative L2 (Adam's first steps are
+-lr * sign(g): elements whose gradient is ~0 may flip), Adam first moments 2e-3, second moments 4e-3, BatchNorm
running statistics 1e-4."""
import copy
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu


from oracle import iid_losses as oracle_iid  # noqa: E402
from oracle import nets as oracle_nets  # noqa: E402
from oracle import seg_losses as oracle_seg  # noqa: E402
from oracle import transforms as oracle_tf  # noqa: E402
from oracle import weights  # noqa: E402

CFG = dict(in_channels=2, input_sz=32, num_sub_heads=3, output_k_A=14, output_k_B=6, batchnorm_track=True)
LR = 1e-3


def _rel(a, b):
  return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _oracle_cluster_step(ora, opt, g, gt, head, lamb):
  opt.zero_grad(set_to_none=False)  # torch 0.4.1 semantics
  x, xt = oracle_tf.sobel_process(g, False), oracle_tf.sobel_process(gt, False)
  o, ot = ora(x, head=head), ora(xt, head=head)
  loss = sum(oracle_iid.IID_loss(a, b, lamb=lamb)[0] for a, b in zip(o, ot)) / len(o)
  loss.backward()
  opt.step()
  return loss.item()


def _sync_oracle_to(net, opt, ora, oopt):
  """Make the oracle (network, BatchNorm buffers, Adam moments and step counts) identical to the CUDA side, so that the
  next step is again a ONE-step comparison: Adam's +-lr updates flip sign where a gradient is ~0, and a 34-layer network at
  batch 10 amplifies that into a 1 % loss difference two steps later (measured) -- a property of the optimiser, not of the
  code under test."""
  ora.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
  names, onames = dict(net.named_parameters()), dict(ora.named_parameters())
  for k, p in names.items():
    st = opt.state.get(p, {})
    if len(st) == 0:
      assert len(oopt.state.get(onames[k], {})) == 0, k
      continue
    ost = oopt.state[onames[k]]
    ost["exp_avg"].copy_(st["exp_avg"].cpu())
    ost["exp_avg_sq"].copy_(st["exp_avg_sq"].cpu())
    assert int(ost["step"]) == int(st["step"]), (k, ost["step"], st["step"])


@pytest.mark.parametrize("pair_batched,use_arena", [(True, False), (False, False), (True, True), (False, True)])
def test_cluster_step_matches_oracle_with_torch_adam(pair_batched, use_arena):
  """Every step is compared as ONE step from identical states (the oracle is re-synchronised after each): the loss to
  2e-5, Adam's first moments to 2e-2 and second moments to 4e-2 (linear / quadratic in the gradient), the parameter
  update element-wise (an Adam update is ~lr * sign(g): the elements whose gradient is below the fp32 noise flip, each
  flip costs 2 lr -- at most 3 % of a tensor's elements, or 2, may disagree by more than lr / 2), running statistics 1e-4; the idle head follows torch 0.4.1's zero_grad semantics."""
  import iic_b200.archs as archs
  from iic_b200.arena import GradArena
  from iic_b200.optim import FusedAdam
  from iic_b200.step import iic_cluster_step
  net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **CFG))
  weights.fill_state_dict(net, salt=5)
  ora = oracle_nets.ClusterNet5gTwoHead(Namespace(**CFG))
  ora.load_state_dict(net.state_dict())
  net.cuda().train()
  ora.train()
  opt = FusedAdam(net.parameters(), lr=LR)
  oopt = torch.optim.Adam(ora.parameters(), lr=LR)
  arena = GradArena(net) if use_arena else None
  names, onames = dict(net.named_parameters()), dict(ora.named_parameters())
  heads = ["B", "B", "A", "B"]  # head A joins at step 3; head B keeps being updated (moment decay) at step 3
  for i, head in enumerate(heads):
    g = weights.uniform("step.g%d" % i, (10, 1, 32, 32))
    gt = (g + 0.05 * weights.normal("step.gt%d" % i, (10, 1, 32, 32))).clamp(0, 1)
    before = {k: p.detach().cpu().clone() for k, p in names.items()}
    loss, loss_nl = iic_cluster_step(net, opt, g.cuda(), gt.cuda(), head=head, lamb=1.2, pair_batched=pair_batched,
                                     arena=arena)
    want = _oracle_cluster_step(ora, oopt, g, gt, head, 1.2)
    assert abs(loss.item() - want) < 2e-5, (i, loss.item(), want)
    sd, osd = net.state_dict(), ora.state_dict()
    for k in osd:
      if k.endswith("num_batches_tracked"):
        assert int(sd[k]) == int(osd[k]), k
      elif "running" in k:
        assert torch.allclose(sd[k].cpu(), osd[k], rtol=1e-4, atol=1e-5), k
    for k, p in names.items():
      ost = oopt.state.get(onames[k], {})
      st = opt.state.get(p, {})
      assert (len(st) == 0) == (len(ost) == 0), (i, k)  # the same parameters are live on both sides
      if len(st) == 0:
        assert torch.equal(p.detach().cpu(), before[k]), k
        continue
      assert int(st["step"]) == int(ost["step"]), (i, k, st["step"], ost["step"])
      assert _rel(st["exp_avg"].cpu(), ost["exp_avg"]) < 2e-2, (i, k)
      assert _rel(st["exp_avg_sq"].cpu(), ost["exp_avg_sq"]) < 4e-2, (i, k)
      upd, wantu = p.detach().cpu() - before[k], onames[k].detach() - before[k]
      assert wantu.abs().max() > 0, (i, k)
      flipped = ((upd - wantu).abs() > 0.5 * LR).float().sum().item()  # elements whose ~lr-sized update disagrees
      assert flipped <= max(2.0, 0.03 * upd.numel()), (i, k, flipped, upd.numel())
    _sync_oracle_to(net, opt, ora, oopt)
  assert int(opt.state[names["head_A.heads.0.0.weight"]]["step"]) == 2  # joined at step 3, decayed at step 4
  assert int(opt.state[names["head_B.heads.0.0.weight"]]["step"]) == 4


def test_set_to_none_skips_the_idle_head():
  """set_to_none=True (modern torch default) is an explicit opt-in: the head that is not trained is not stepped."""
  import iic_b200.archs as archs
  from iic_b200.optim import FusedAdam
  from iic_b200.step import iic_cluster_step
  net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **CFG))
  weights.fill_state_dict(net, salt=6)
  net.cuda().train()
  opt = FusedAdam(net.parameters(), lr=LR)
  g = weights.uniform("step.n", (6, 1, 32, 32)).cuda()
  iic_cluster_step(net, opt, g, g.flip(3), head="A", set_to_none=True)
  wb = net.head_B.heads[0][0].weight.detach().clone()
  wa = net.head_A.heads[0][0].weight.detach().clone()
  iic_cluster_step(net, opt, g, g.flip(3), head="A", set_to_none=True)
  assert torch.equal(wb, net.head_B.heads[0][0].weight) and not torch.equal(wa, net.head_A.heads[0][0].weight)
  assert net.head_B.heads[0][0].weight.grad is None


def test_fused_adam_class_vs_torch_and_legacy_state():
  """FusedAdam.step through the class: param groups, late joiners, weight decay; state dicts interchange with
  torch.optim.Adam in both directions, including the reference's torch-0.4.1 layout (int `step`, `amsgrad` key)."""
  from iic_b200.optim import FusedAdam
  torch.manual_seed(0)
  shapes = [(64, 3, 3, 3), (64,), (10, 512), (1,), (130, 7)]
  ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
  qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
  opt = FusedAdam([{"params": ps[:3]}, {"params": ps[3:], "lr": 3e-3, "weight_decay": 0.01}], lr=1e-3, betas=(0.8, 0.95))
  ref = torch.optim.Adam([{"params": qs[:3]}, {"params": qs[3:], "lr": 3e-3, "weight_decay": 0.01}], lr=1e-3, betas=(0.8, 0.95))
  for it in range(4):
    for i, (p, q) in enumerate(zip(ps, qs)):
      if i == 2 and it < 2:
        p.grad = q.grad = None  # joins at the third step
        continue
      gr = torch.randn_like(p)
      p.grad, q.grad = gr.clone(), gr.clone()
    opt.step()
    ref.step()
  for p, q in zip(ps, qs):
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
  # torch -> FusedAdam, with the state rewritten the way torch 0.4.1 stored it
  sd = ref.state_dict()
  for st in sd["state"].values():
    st["step"] = int(st["step"])
  for gr in sd["param_groups"]:
    gr["amsgrad"] = False
  opt2 = FusedAdam([{"params": ps[:3]}, {"params": ps[3:]}])
  opt2.load_state_dict(sd)
  ref2 = torch.optim.Adam([{"params": qs[:3]}, {"params": qs[3:]}])
  ref2.load_state_dict(copy.deepcopy(ref.state_dict()))
  for p, q in zip(ps, qs):
    gr = torch.randn_like(p)
    p.grad, q.grad = gr.clone(), gr.clone()
  opt2.step()
  ref2.step()
  for p, q in zip(ps, qs):
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
  # FusedAdam -> torch
  ref3 = torch.optim.Adam([{"params": qs[:3]}, {"params": qs[3:]}])
  ref3.load_state_dict(opt2.state_dict())
  assert int(ref3.state[qs[0]]["step"]) == 5


def test_backward_guards():
  """Second backward through the same trunk forward, and an in-place weight update between forward and backward,
  raise (stock autograd would: freed buffers / version counter) instead of silently using stale data."""
  import iic_b200.archs as archs
  from iic_b200.utils.cluster.IID_losses import IID_loss_subheads
  net = archs.ClusterNet5gTwoHead(Namespace(precision="bf16", **CFG))
  weights.fill_state_dict(net, salt=7)
  net.cuda().train()
  x = weights.normal("guard.x", (4, 2, 32, 32)).cuda()
  z, zt = net.forward_stacked_pair(x, x.flip(3))
  loss = IID_loss_subheads(z, zt)[0].mean()
  loss.backward(retain_graph=True)
  with pytest.raises(RuntimeError, match="already been differentiated"):
    loss.backward()
  z, zt = net.forward_stacked_pair(x, x.flip(3))
  loss = IID_loss_subheads(z, zt)[0].mean()
  with torch.no_grad():
    net.trunk.layer1[0].conv1.weight.mul_(1.01)
  with pytest.raises(RuntimeError, match="modified in place"):
    loss.backward()


def test_seg_step_matches_oracle_with_torch_adam():
  """iic_seg_step (segmentation_twohead.py:262-361) against the oracle stepped by torch.optim.Adam, one step at a time
  from re-synchronised states (see _sync_oracle_to): loss 5e-5, Adam moments 2e-2 / 4e-2, updates element-wise."""
  import iic_b200.archs as archs
  from iic_b200.optim import FusedAdam
  from iic_b200.step import iic_seg_step
  cfg = dict(in_channels=5, input_sz=32, num_sub_heads=1, output_k_A=15, output_k_B=3, batchnorm_track=True)
  net = archs.SegmentationNet10aTwoHead(Namespace(precision="fp32", **cfg))
  weights.fill_state_dict(net, head_gain=20.0)
  ora = oracle_nets.SegmentationNet10aTwoHead(Namespace(**cfg))
  ora.load_state_dict(net.state_dict())
  net.cuda().train()
  ora.train()
  opt, oopt = FusedAdam(net.parameters(), lr=LR), torch.optim.Adam(ora.parameters(), lr=LR)
  names, onames = dict(net.named_parameters()), dict(ora.named_parameters())
  n = 2
  theta = torch.zeros(n, 2, 3)
  theta[:, 0, 0] = 1.
  theta[:, 1, 1] = 1.
  theta[1, 0, 0] = -1.
  for i, (head, lamb, unc) in enumerate([("A", 1.0, True), ("B", 1.5, True), ("A", 1.0, False)]):
    img = weights.uniform("seg.step%d" % i, (n, 4, 32, 32))
    img_tf = (img + 0.05 * weights.normal("seg.stept%d" % i, (n, 4, 32, 32))).clamp(0, 1)
    mask = (weights.uniform("seg.stepm%d" % i, (n, 32, 32)) < 0.7).float()
    before = {k: p.detach().cpu().clone() for k, p in names.items()}
    loss, _ = iic_seg_step(net, opt, img.cuda(), img_tf.cuda(), theta.cuda(), mask.cuda(), head=head, lamb=lamb,
                           half_T_side_dense=3, uncollapsed=unc)
    oopt.zero_grad(set_to_none=False)
    x, xt = oracle_tf.sobel_process(img, True), oracle_tf.sobel_process(img_tf, True)
    o, ot = ora(x, head=head), ora(xt, head=head)
    fn = oracle_seg.IID_segmentation_loss_uncollapsed if unc else oracle_seg.IID_segmentation_loss
    ol, _ = fn(o[0], ot[0], all_affine2_to_1=theta, all_mask_img1=mask, lamb=lamb, half_T_side_dense=3,
               half_T_side_sparse_min=0, half_T_side_sparse_max=0)
    ol.backward()
    oopt.step()
    assert abs(loss.item() - ol.item()) < 5e-5 * max(1.0, abs(ol.item())), (i, loss.item(), ol.item())
    for k, p in names.items():
      ost, st = oopt.state.get(onames[k], {}), opt.state.get(p, {})
      assert (len(st) == 0) == (len(ost) == 0), (i, k)
      if len(st) == 0:
        continue
      assert _rel(st["exp_avg"].cpu(), ost["exp_avg"]) < 2e-2, (i, k, _rel(st["exp_avg"].cpu(), ost["exp_avg"]))
      assert _rel(st["exp_avg_sq"].cpu(), ost["exp_avg_sq"]) < 4e-2, (i, k)
      upd, wantu = p.detach().cpu() - before[k], onames[k].detach() - before[k]
      flipped = ((upd - wantu).abs() > 0.5 * LR).float().sum().item()
      assert flipped <= max(2.0, 0.03 * upd.numel()), (i, k, flipped, upd.numel())
    _sync_oracle_to(net, opt, ora, oopt)


def test_graphed_step_equals_eager_steps():
  """iic_b200.graph.GraphedStep (one cudaGraphLaunch per step) against the eager step: same losses and parameters over
  several steps (fp32 mode, identical kernels -> tight), Adam step counters advance, the idle head is untouched."""
  import iic_b200.archs as archs
  from iic_b200.arena import GradArena
  from iic_b200.graph import GraphedStep
  from iic_b200.optim import FusedAdam
  from iic_b200.step import iic_cluster_step
  nets, opts, arenas = [], [], []
  for _ in range(2):
    net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **CFG))
    weights.fill_state_dict(net, salt=31)
    net.cuda().train()
    nets.append(net)
    opts.append(FusedAdam(net.parameters(), lr=LR))
    arenas.append(GradArena(net))
  batches = []
  for i in range(6):
    g = weights.uniform("graph.g%d" % i, (8, 1, 32, 32))
    batches.append((g.cuda(), (g + 0.05 * weights.normal("graph.t%d" % i, (8, 1, 32, 32))).clamp(0, 1).cuda()))
  # eager net: 3 warm-up steps on batch 0 (what GraphedStep does while it builds), then batches 1..5
  eager_losses = []
  for _ in range(3):
    iic_cluster_step(nets[0], opts[0], *batches[0], head="B", arena=arenas[0])
  for b in batches[1:]:
    eager_losses.append(iic_cluster_step(nets[0], opts[0], *b, head="B", arena=arenas[0])[0].item())
  gs = GraphedStep(nets[1], opts[1], arenas[1], batches[0], head="B", warmup=3)
  graph_losses = [gs(*b)[0].item() for b in batches[1:]]
  for a, b in zip(eager_losses, graph_losses):
    assert abs(a - b) < 1e-6, (eager_losses, graph_losses)
  for (k, p), (_, q) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), k
  pa, pb = dict(nets[1].named_parameters()), opts[1]
  assert int(pb.state[pa["trunk.conv1.weight"]]["step"]) == 8
  assert len(pb.state[pa["head_A.heads.0.0.weight"]]) == 0  # never trained: no state, untouched


def test_wgrad_on_second_stream_gives_identical_gradients():
  """OPTIONS["wgrad_stream"] (weight-gradient convolutions on a second stream, overlapping the BatchNorm backward of
  the next stage; library option bn_bwd_ctas = 1) only changes the schedule: losses and every gradient are bit-identical
  to the single-stream step, eager and with the arena, over two steps (stream-ordering bugs show up as garbage)."""
  import iic_b200.archs as archs
  from iic_b200 import kernels as K
  from iic_b200.archs import _engine as E
  from iic_b200.arena import GradArena
  from iic_b200.step import iic_cluster_step
  res = {}
  for mode in ("single", "second-stream"):
    for use_arena in (False, True):
      net = archs.ClusterNet5gTwoHead(Namespace(precision="bf16", **CFG))
      weights.fill_state_dict(net, salt=41)
      net.cuda().train()
      arena = GradArena(net) if use_arena else None
      old = E.OPTIONS["wgrad_stream"]
      E.OPTIONS["wgrad_stream"] = mode != "single"
      try:
        with K.options(bn_bwd_ctas=1):  # (both: the grid of the BatchNorm backward fixes its summation order)
          out = []
          for i in range(2):
            g = weights.uniform("ws.g%d" % i, (24, 1, 32, 32)).cuda()
            loss, _ = iic_cluster_step(net, None, g, g.flip(3), head="B", arena=arena)
            torch.cuda.synchronize()
            out.append((loss.item(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
      finally:
        E.OPTIONS["wgrad_stream"] = old
      res[(mode, use_arena)] = out
  for use_arena in (False, True):
    a, b = res[("single", use_arena)], res[("second-stream", use_arena)]
    for (la, ga), (lb, gb) in zip(a, b):
      assert la == lb
      assert ga.keys() == gb.keys()
      for k in ga:
        assert torch.equal(ga[k], gb[k]), k
