"""GPU parity of the full networks (trunk + sub-heads + IID loss + backward) against the
golden vectors produced by the reference modules, and against the CPU oracle.

Stated tolerances, and where they come from (measured on the CPU oracle, see DESIGN.md "Parity"):
the golden configurations are a randomly-initialised 34-layer BN ResNet at batch 3-6 -- a chaotic
map: the fp32 oracle already differs from the fp64 oracle by up to 1.3 % relative L2 per parameter
gradient (ReLU sign flips), and two bf16-storage emulations that differ only in accumulation
precision differ by ~45 % (median) in the gradients while their outputs agree to ~3e-2.
  precision="fp32" (fp32 SIMT convolutions; same engine/orchestration as bf16): softmax outputs
      within 2e-4 abs, loss within 2e-5 abs, gradient norms within 1e-2 relative, stored full
      gradients within 5e-2 relative L2 (4x the oracle's own fp32-vs-fp64 noise).
  precision="bf16" (tcgen05 bf16 MMA, fp32 TMEM accumulate, bf16 activations): tight parity is
      established per residual block against the oracle with bf16 rounding at the same storage
      points (test_bf16_block_matches_rounding_oracle, 3e-2 relative L2); end-to-end the outputs
      must stay within 0.1 abs and the shallow 6c net within 15 % gradient-norm / cos > 0.98."""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402
from oracle import iid_losses as oracle_iid  # noqa: E402
from oracle import nets as oracle_nets  # noqa: E402
from oracle import weights  # noqa: E402


def _build(name, precision):
  import iic_b200.archs as archs
  ctor, cfg, batch, head, lamb = make_golden.NET_SPECS[name]
  net = archs.__dict__[ctor](Namespace(precision=precision, **cfg))
  weights.fill_state_dict(net)
  net.cuda().train()
  x, xt = make_golden.net_input("net." + name, cfg, batch)
  return net, x.cuda(), xt.cuda(), head, lamb


def _step(net, x, xt, head, lamb):
  from iic_b200.utils.cluster.IID_losses import IID_loss
  kw = {} if head is None else {"head": head}
  o, ot = net(x, **kw), net(xt, **kw)
  loss = sum(IID_loss(a, b, lamb=lamb)[0] for a, b in zip(o, ot)) / len(o)
  loss.backward()
  return o, ot, loss


NAMES = ["5g2h_32_A", "5g2h_32_B", "5g2h_96_B", "5g_64", "6c2h_24_A", "6c_24"]


@pytest.mark.parametrize("name", NAMES)
def test_fp32_mode_matches_reference_goldens(name, golden_nets):
  c = golden_nets.sub("net/" + name)
  net, x, xt, head, lamb = _build(name, "fp32")
  o, ot, loss = _step(net, x, xt, head, lamb)
  got = torch.stack(o).detach().cpu().numpy()
  np.testing.assert_allclose(got, c["ref_out"], rtol=0, atol=2e-4)
  np.testing.assert_allclose(torch.stack(ot).detach().cpu().numpy(), c["ref_out_tf"], rtol=0, atol=2e-4)
  assert abs(loss.item() - float(c["loss"])) < 2e-5
  params = dict(net.named_parameters())
  assert list(params) == [str(s) for s in c["grad_names"]]
  for pn, norm in zip(c["grad_names"], c["grad_norms"]):
    g = params[str(pn)].grad
    mine = 0.0 if g is None else float(g.double().norm())
    assert abs(mine - norm) <= 1e-2 * norm + 1e-7, (str(pn), mine, norm)
  for key in c.sub("grad"):
    want = c["grad/" + key].astype(np.float64)
    got_g = params[key].grad.cpu().numpy().astype(np.float64)
    rel = np.linalg.norm(got_g - want) / (np.linalg.norm(want) + 1e-30)
    assert rel <= 5e-2, (key, rel)
  sd = net.state_dict()
  for key in c.sub("buf"):
    np.testing.assert_allclose(sd[key].cpu().numpy(), c["buf/" + key], rtol=1e-4, atol=1e-5)
  # eval-mode forward (running statistics) of the trunk feature
  net.eval()
  with torch.no_grad():
    kw = {} if head is None else {"head": head}
    f = net(x, trunk_features=True, **kw)
  want = c["ref_trunk_eval"]
  np.testing.assert_allclose(f.cpu().numpy()[:, :want.shape[1]], want, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("name", ["5g2h_32_A", "5g2h_96_B", "6c2h_24_A"])
def test_bf16_mode_end_to_end_vs_goldens(name, golden_nets):
  c = golden_nets.sub("net/" + name)
  net, x, xt, head, lamb = _build(name, "bf16")
  o, ot, loss = _step(net, x, xt, head, lamb)
  got = torch.stack(o).detach().cpu().numpy()
  assert np.isfinite(got).all() and np.abs(got.sum(-1) - 1).max() < 1e-4
  assert np.abs(got - c["ref_out"]).max() < 0.1, np.abs(got - c["ref_out"]).max()
  params = dict(net.named_parameters())
  assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in params.items() if "head_" + ("B" if head == "A" else "A") not in n)
  if not name.startswith("6c"):
    return  # deep random ResNet at batch <= 6: chaotic under bf16 rounding (see module docstring)
  bad = []
  for pn, norm in zip(c["grad_names"], c["grad_norms"]):
    g = params[str(pn)].grad
    mine = 0.0 if g is None else float(g.double().norm())
    if abs(mine - norm) > 0.15 * norm + 1e-6:
      bad.append((str(pn), mine, norm))
  assert len(bad) <= len(params) // 20, bad[:10]
  # direction of the full gradient: cosine similarity with the fp32 reference gradient
  for key in c.sub("grad"):
    want = c["grad/" + key].ravel().astype(np.float64)
    gotg = params[key].grad.cpu().numpy().ravel().astype(np.float64)
    cos = float(want @ gotg / (np.linalg.norm(want) * np.linalg.norm(gotg) + 1e-30))
    assert cos > 0.98, (key, cos)


def test_oracle_live_batch_and_kwargs():
  """Against the CPU oracle run live (not only fixtures): other batch size, eval mode, the
  forward kwargs of SURVEY.md S3.4."""
  import iic_b200.archs as archs
  cfg = dict(in_channels=2, input_sz=32, num_sub_heads=3, output_k_A=20, output_k_B=5, batchnorm_track=True)
  net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **cfg))
  weights.fill_state_dict(net, salt=3)
  ora = oracle_nets.ClusterNet5gTwoHead(Namespace(**cfg))
  ora.load_state_dict(net.state_dict())
  net.cuda().train()
  ora.train()
  x = weights.normal("live.x", (9, 2, 32, 32))
  for head in ("A", "B"):
    a = net(x.cuda(), head=head)
    b = ora(x, head=head)
    assert isinstance(a, list) and len(a) == 3
    for u, v in zip(a, b):
      assert torch.allclose(u.cpu(), v, rtol=0, atol=2e-4)
  f = net(x.cuda(), trunk_features=True)
  assert f.shape == (9, 512) and torch.allclose(f.cpu(), ora(x, trunk_features=True), rtol=1e-3, atol=1e-3)
  dup = net(x.cuda(), kmeans_use_features=True)
  ora(x, kmeans_use_features=True)  # keep the two nets' running statistics in step
  assert len(dup) == 3 and dup[0].shape == (9, 512)
  net.eval(), ora.eval()
  with torch.no_grad():
    pen = net(x.cuda(), trunk_features=True, penultimate_features=True)
    assert torch.allclose(pen.cpu(), ora(x, trunk_features=True, penultimate_features=True), rtol=1e-3, atol=1e-3)
  with pytest.raises(AssertionError):
    net(x.cuda(), head="C")
  with pytest.raises(RuntimeError):
    net(x)  # CPU tensors: no fallback


def test_shard_equivalence_per_rank_batchnorm():
  """W-GPU result == single-device oracle with BN statistics per contiguous chunk.
  Emulated on one GPU: two chunked forwards + summed partial joints (what the NCCL path does)."""
  import iic_b200.archs as archs
  from iic_b200 import _lib, kernels
  cfg = dict(in_channels=2, input_sz=32, num_sub_heads=2, output_k_A=12, output_k_B=4, batchnorm_track=False)
  net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **cfg))
  weights.fill_state_dict(net, salt=5)
  ora = oracle_nets.ClusterNet5gTwoHead(Namespace(**cfg))
  ora.load_state_dict(net.state_dict())
  net.cuda().train(), ora.train()
  x = weights.normal("shard.x", (8, 2, 32, 32))
  xt = x + 0.3 * weights.normal("shard.xt", (8, 2, 32, 32))
  # oracle: chunked forward (per-chunk BN), loss on the concatenated batch
  oz = [torch.cat([ora(x[i:i + 4], head="A")[s] for i in (0, 4)]) for s in range(2)]
  ozt = [torch.cat([ora(xt[i:i + 4], head="A")[s] for i in (0, 4)]) for s in range(2)]
  oloss = sum(oracle_iid.IID_loss(a, b)[0] for a, b in zip(oz, ozt)) / 2
  oloss.backward()
  # ours: per-shard forward, PARTIAL joints summed, FINISH per shard
  zs, zts = [], []
  for i in (0, 4):
    zs.append(net.forward_stacked(x[i:i + 4].cuda(), head="A"))
    zts.append(net.forward_stacked(xt[i:i + 4].cuda(), head="A"))
  joint = torch.zeros(2, 12, 12, device="cuda")
  for z, zt in zip(zs, zts):
    j = torch.empty_like(joint)
    kernels.iid_loss(z.detach().contiguous(), zt.detach().contiguous(), 1.0, sys.float_info.epsilon, False,
                     phase=_lib.PHASE_PARTIAL, joint_ws=j)
    joint += j
  for z, zt in zip(zs, zts):
    loss, dz, dzt, _ = kernels.iid_loss(z.detach().contiguous(), zt.detach().contiguous(), 1.0,
                                        sys.float_info.epsilon, True, phase=_lib.PHASE_FINISH, joint_ws=joint)
    assert abs(loss[:, 0].mean().item() - oloss.item()) < 2e-5
    torch.autograd.backward([z, zt], [dz / 2, dzt / 2])
  op = dict(ora.named_parameters())
  for pn, p in net.named_parameters():
    if p.grad is None:
      continue
    want = op[pn].grad
    assert ((p.grad.cpu() - want).norm() / want.norm()).item() <= 5e-2, pn


@pytest.fixture(params=[1, 2], ids=["halo-auto", "halo-forced"])
def halo_mode(request):
  """conv_halo = 2 pushes the 64-channel 3x3 convs through the halo kernels (fprop with fused BN statistics, dgrad
  with the residual addend, wgrad) at test sizes; by default they only take over at production sizes."""
  from iic_b200 import kernels
  with kernels.options(conv_halo=request.param):
    yield request.param


@pytest.mark.parametrize("cin,cout,stride,hw,n", [(64, 64, 1, 17, 5), (64, 128, 2, 17, 6), (128, 256, 2, 9, 8),
                                                  (256, 512, 2, 13, 4), (512, 512, 1, 7, 9), (64, 64, 1, 49, 2)])
def test_bf16_block_matches_rounding_oracle(cin, cout, stride, hw, n, halo_mode):
  """One BasicBlock (2 tcgen05 convs, 2-3 BNs, residual) forward + backward in bf16 mode against the
  oracle block with bf16 rounding emulated at the product's storage points (oracle/nets.py q/qw).
  A single block is not chaotic, so this is a tight check of every bf16 kernel in composition."""
  import torch.nn as nn
  from iic_b200.archs import _engine as E
  from iic_b200.archs.cluster.residual import BasicBlock
  from iic_b200._lib import BF16
  if halo_mode == 2 and cin != 64:
    pytest.skip("no 64 -> 64 conv in this block")
  ds = None
  if stride != 1 or cin != cout:
    ds = nn.Sequential(E.ConvParams(cin, cout, 1, stride, 0), E.BNParams(cout, False))
  blk = BasicBlock(cin, cout, stride, ds, track_running_stats=False)
  weights.fill_state_dict(blk, salt=11)
  oblk = oracle_nets._Block(cin, cout, stride, False)
  oblk.load_state_dict(blk.state_dict())
  blk.cuda()
  tag = "blk.%d.%d.%d" % (cin, cout, stride)
  x = torch.relu(weights.normal(tag + ".x", (n, cin, hw, hw))).bfloat16().float()
  oh = (hw + 2 - 3) // stride + 1
  dout = weights.normal(tag + ".d", (n, cout, oh, oh)).bfloat16().float()
  ctx = E._Ctx(BF16, True, True)
  out = E.block_forward(ctx, blk, x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16())
  sink = E.GradSink()
  dx = E.block_backward(ctx, sink, ctx.saved[-1], dout.permute(0, 2, 3, 1).contiguous().cuda().bfloat16())
  xo = x.clone().requires_grad_(True)
  with oracle_nets.bf16_rounding():
    oo = oblk(oracle_nets.q(xo))
    oo.backward(dout)

  def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

  assert rel(out.float().permute(0, 3, 1, 2).cpu(), oo.detach()) < 1e-2
  assert rel(dx.float().permute(0, 3, 1, 2).cpu(), xo.grad) < 3e-2
  for (pn, p), (_, po) in zip(blk.named_parameters(), oblk.named_parameters()):
    assert rel(sink.get(p).cpu(), po.grad) < 5e-2, pn  # (BN statistics come from the fp32 accumulators, the oracle's from the rounded y)


@pytest.mark.parametrize("arch,cfg", [
  ("ClusterNet5gTwoHead", dict(in_channels=2, input_sz=32, num_sub_heads=3, output_k_A=20, output_k_B=5, batchnorm_track=True)),
  ("ClusterNet6cTwoHead", dict(in_channels=1, input_sz=24, num_sub_heads=2, output_k_A=12, output_k_B=4, batchnorm_track=True)),
])
def test_pair_batched_pass_equals_two_forward_calls(arch, cfg):
  """forward_stacked_pair pushes both views through the trunk in one pass with per-view BatchNorm
  statistics: it must reproduce two separate forward calls (outputs, gradients, running stats)."""
  import iic_b200.archs as archs
  from iic_b200.utils.cluster.IID_losses import IID_loss_subheads
  nets_ = []
  for _ in range(2):
    net = archs.__dict__[arch](Namespace(precision="fp32", **cfg))
    weights.fill_state_dict(net, salt=4)
    nets_.append(net.cuda().train())
  x = weights.normal("pair.x", (10, cfg["in_channels"], cfg["input_sz"], cfg["input_sz"])).cuda()
  xt = x + 0.3 * weights.normal("pair.xt", tuple(x.shape)).cuda()
  za, zta = nets_[0].forward_stacked(x, head="A"), nets_[0].forward_stacked(xt, head="A")
  zb, ztb = nets_[1].forward_stacked_pair(x, xt, head="A")
  assert torch.allclose(za, zb, rtol=0, atol=2e-6) and torch.allclose(zta, ztb, rtol=0, atol=2e-6)
  IID_loss_subheads(za, zta)[0].mean().backward()
  IID_loss_subheads(zb, ztb)[0].mean().backward()
  for (n1, p1), (_, p2) in zip(nets_[0].named_parameters(), nets_[1].named_parameters()):
    if p1.grad is None:
      assert p2.grad is None
      continue
    assert ((p1.grad - p2.grad).norm() / (p1.grad.norm() + 1e-30)).item() < 2e-3, n1
  sd1, sd2 = nets_[0].state_dict(), nets_[1].state_dict()
  for key in sd1:
    if "running" in key or "num_batches" in key:
      assert torch.allclose(sd1[key].float(), sd2[key].float(), rtol=1e-5, atol=1e-6), key


@pytest.mark.parametrize("cin,cout,stride,hw,n", [(64, 64, 1, 17, 5), (64, 128, 2, 17, 6), (256, 512, 2, 13, 4)])
def test_bf16_block_with_relu_bitmask(cin, cout, stride, hw, n):
  """Engine switch bn_bitmask: the residual block forward / backward must give what it gives with the switch off."""
  import torch.nn as nn
  from iic_b200.archs import _engine as E
  from iic_b200.archs.cluster.residual import BasicBlock
  from iic_b200._lib import BF16
  ds = None
  if stride != 1 or cin != cout:
    ds = nn.Sequential(E.ConvParams(cin, cout, 1, stride, 0), E.BNParams(cout, False))
  blk = BasicBlock(cin, cout, stride, ds, track_running_stats=False)
  weights.fill_state_dict(blk, salt=13)
  blk.cuda()
  x = torch.relu(weights.normal("bits.x", (2 * n, cin, hw, hw))).permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
  oh = (hw + 2 - 3) // stride + 1
  dout = weights.normal("bits.d", (2 * n, oh, oh, cout)).cuda().bfloat16()
  results = []
  for flag in (False, True):
    prev = E.OPTIONS["bn_bitmask"]
    E.OPTIONS["bn_bitmask"] = flag
    try:
      ctx = E._Ctx(BF16, True, True, groups=2)
      out = E.block_forward(ctx, blk, x)
      assert (len(ctx.mbits) == 1) == flag
      sink = E.GradSink()
      dx = E.block_backward(ctx, sink, ctx.saved[-1], dout)
      results.append((out, dx, [sink.get(p).clone() for p in blk.parameters()]))
    finally:
      E.OPTIONS["bn_bitmask"] = prev
  (o0, d0, g0), (o1, d1, g1) = results
  assert torch.equal(o0, o1) and torch.equal(d0, d1)
  for a, b in zip(g0, g1):
    assert torch.equal(a, b)


@pytest.mark.parametrize("halo,mt2,addtma", [(2, 2, 1), (0, 0, 0), (2, 1, 0)])
@pytest.mark.parametrize("cin,cout,stride,hw,n", [(64, 64, 1, 17, 5), (128, 128, 1, 13, 4), (256, 256, 1, 7, 3), (64, 128, 2, 17, 6),
                                                  (128, 128, 1, 30, 2)])
def test_bf16_block_with_masked_addend(cin, cout, stride, hw, n, halo, mt2, addtma):
  """Engine switch masked_addend: the masked residual gradient is never written; conv1's dgrad epilogue (halo kernel with
  the TMA-loaded or the per-lane addend, im2col kernel with one or two tiles per work item) or the downsample BatchNorm
  backward apply the mask bits to d_out.  Bit-identical to the materialised copy."""
  import torch.nn as nn
  from iic_b200.archs import _engine as E
  from iic_b200.archs.cluster.residual import BasicBlock
  from iic_b200 import kernels as K
  from iic_b200._lib import BF16
  ds = None
  if stride != 1 or cin != cout:
    ds = nn.Sequential(E.ConvParams(cin, cout, 1, stride, 0), E.BNParams(cout, False))
  blk = BasicBlock(cin, cout, stride, ds, track_running_stats=False)
  weights.fill_state_dict(blk, salt=17)
  blk.cuda()
  x = torch.relu(weights.normal("madd.x", (2 * n, cin, hw, hw))).permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
  oh = (hw + 2 - 3) // stride + 1
  dout = weights.normal("madd.d", (2 * n, oh, oh, cout)).cuda().bfloat16()
  results = []
  for flag in (False, True):
    prev = (E.OPTIONS["bn_bitmask"], E.OPTIONS["masked_addend"])
    E.OPTIONS["bn_bitmask"], E.OPTIONS["masked_addend"] = True, flag
    try:
      with K.options(conv_halo=halo, tc2_mt2=mt2, halo_addend_tma=addtma):
        ctx = E._Ctx(BF16, True, True, groups=2)
        out = E.block_forward(ctx, blk, x)
        sink = E.GradSink()
        dx = E.block_backward(ctx, sink, ctx.saved[-1], dout)
        torch.cuda.synchronize()
      results.append((out, dx, [sink.get(p).clone() for p in blk.parameters()]))
    finally:
      E.OPTIONS["bn_bitmask"], E.OPTIONS["masked_addend"] = prev
  (o0, d0, g0), (o1, d1, g1) = results
  assert torch.equal(o0, o1)
  assert torch.equal(d0, d1), "max |diff| %g" % (d0.float() - d1.float()).abs().max().item()
  for a, b in zip(g0, g1):
    assert torch.equal(a, b)
