"""SURVEY S8f row 2 on the GPU: dataloader tail fused with sobel (RGB -> grey -> sobel), slab assembly, and the
repeat handling of view 1 (cluster_sobel_twohead.py:290-318: the same tf1 batch in every slab)."""
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import transforms as otf  # noqa: E402
from oracle import weights  # noqa: E402


def test_grey_sobel_matches_oracle_fp32_and_uint8():
  from iic_b200.utils.cluster.transforms import rgb_sobel_process
  x = weights.uniform("batch.rgb", (5, 3, 33, 47))
  got = rgb_sobel_process(x.cuda())
  want = otf.sobel_process(otf.grey_from_rgb(x), False)
  assert got.shape == want.shape == (5, 2, 33, 47)
  assert torch.allclose(got.cpu(), want, rtol=0, atol=3e-6)
  u = (weights.uniform("batch.u8", (4, 3, 96, 96)) * 256).clamp(0, 255).to(torch.uint8)
  got = rgb_sobel_process(u.cuda())
  want = otf.sobel_process(otf.grey_from_rgb(u), False)  # PIL's integer "L" formula, then /255
  assert torch.allclose(got.cpu(), want, rtol=0, atol=3e-6)
  with pytest.raises(AssertionError):
    rgb_sobel_process(torch.zeros(1, 4, 8, 8).cuda())


def test_assemble_slabs_is_the_reference_loop():
  from iic_b200.step import assemble_slabs
  n, R = 6, 3
  a = weights.uniform("slab.a", (n, 1, 16, 16))
  tfs = [weights.uniform("slab.t%d" % d, (n, 1, 16, 16)) for d in range(R)]
  # cluster_sobel_twohead.py:290-318
  all_imgs, all_tf = torch.zeros(R * n, 1, 16, 16), torch.zeros(R * n, 1, 16, 16)
  for d in range(R):
    all_imgs[d * n:(d + 1) * n] = a
    all_tf[d * n:(d + 1) * n] = tfs[d]
  for src in ("host", "device"):
    g1, g2 = assemble_slabs(a.pin_memory() if src == "host" else a.cuda(),
                            [t.pin_memory() if src == "host" else t.cuda() for t in tfs])
    assert g1.is_cuda and torch.equal(g1.cpu(), all_imgs) and torch.equal(g2.cpu(), all_tf)


@pytest.mark.parametrize("head", ["A", "B"])
def test_repeat_handling_equals_slab_assembled_step(head):
  """Forwarding the unique tf1 images once and repeating their softmax rows gives the loss and the parameter
  gradients of the reference's slab-assembled batch (fp32 mode; only running_var's unbiased factor differs)."""
  import iic_b200.archs as archs
  from iic_b200.step import assemble_slabs, iic_cluster_step
  cfg = dict(in_channels=2, input_sz=32, num_sub_heads=3, output_k_A=14, output_k_B=6, batchnorm_track=True)
  n, R = 4, 3
  a = weights.uniform("rep.a", (n, 1, 32, 32))
  tfs = [(a + 0.05 * weights.normal("rep.t%d" % d, (n, 1, 32, 32))).clamp(0, 1) for d in range(R)]
  nets = []
  for _ in range(2):
    net = archs.ClusterNet5gTwoHead(Namespace(precision="fp32", **cfg))
    weights.fill_state_dict(net, salt=11)
    nets.append(net.cuda().train())
  full, dedup = nets
  all_imgs, all_tf = assemble_slabs(a.cuda(), [t.cuda() for t in tfs])
  l_full, _ = iic_cluster_step(full, None, all_imgs, all_tf, head=head, lamb=1.1, pair_batched=False)
  l_dd, _ = iic_cluster_step(dedup, None, a.cuda(), all_tf, head=head, lamb=1.1, repeats=R)
  assert abs(l_full.item() - l_dd.item()) < 2e-6
  for (k, p), (_, q) in zip(full.named_parameters(), dedup.named_parameters()):
    if p.grad is None:
      assert q.grad is None, k
      continue
    rel = ((p.grad - q.grad).norm() / (p.grad.norm() + 1e-30)).item()
    assert rel < 1e-2, (k, rel)  # fp32 re-association between the two schedules (measured 3.4e-3 on the stem)
  sf, sd = full.state_dict(), dedup.state_dict()
  for k in sf:
    if k.endswith("running_mean"):
      assert torch.allclose(sf[k], sd[k], rtol=1e-4, atol=1e-6), k
