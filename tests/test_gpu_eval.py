"""GPU tests of the evaluation path (csrc/eval.cu, iic_b200/utils/cluster/{eval_metrics,cluster_eval}.py,
iic_b200/utils/segmentation/segmentation_eval.py) against torch / numpy and the oracle's restatement
of the reference loops.  Index and integer work: everything must be exact.

Validated on a B200 in round 2 (gpurun_out/a_tests.log, profiles/r02_session_a.md)."""
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import eval_metrics as oem

pytestmark = pytest.mark.gpu


def _K():
  from iic_b200 import kernels
  return kernels


@pytest.mark.parametrize("S,n,k", [(5, 704, 10), (5, 333, 70), (1, 17, 3), (2, 64, 255), (3, 40, 33)])
def test_argmax_rows(S, n, k):
  K = _K()
  g = torch.Generator().manual_seed(k)
  z = torch.softmax(torch.randn(S, n, k, generator=g), dim=2).cuda()
  assert torch.equal(K.argmax_rows(z).long(), torch.argmax(z, dim=2))
  # ties -> lowest index; NaN counts as the maximum
  z2 = torch.zeros(4, k).cuda()
  z2[1, k - 1] = 1.0
  z2[2, k // 2] = z2[2, k - 1] = 0.5
  z2[3, k - 1] = float("nan")
  assert K.argmax_rows(z2).tolist() == [0, k - 1, k // 2, k - 1]


@pytest.mark.parametrize("n,k,h", [(3, 15, 32), (2, 3, 128), (1, 255, 9)])
def test_argmax_channels(n, k, h):
  K = _K()
  g = torch.Generator().manual_seed(n + k)
  x = torch.randn(n, k, h, h, generator=g).cuda()
  x[0, :, 0, 0] = 0.25  # a full tie
  want = torch.argmax(x, dim=1)
  got = K.argmax_channels(x).long()
  assert got[0, 0, 0].item() == 0
  want[0, 0, 0] = 0
  assert torch.equal(got, want)


@pytest.mark.parametrize("S,n,pk,tk", [(5, 7040, 10, 10), (5, 5000, 70, 10), (1, 1 << 18, 15, 15), (2, 3000, 255, 40), (1, 5, 3, 3)])
def test_confusion_counts(S, n, pk, tk):
  K = _K()
  rng = np.random.RandomState(pk + n % 97)
  p = rng.randint(-1, pk + 1, (S, n)).astype(np.int32)  # includes out-of-range labels: ignored
  t = rng.randint(0, tk, n).astype(np.int32)
  m = (rng.rand(n) < 0.7).astype(np.uint8)

  def ref(mask):
    out = np.zeros((S, pk, tk), dtype=np.int64)
    for s in range(S):
      ok = (p[s] >= 0) & (p[s] < pk) & (mask != 0)
      np.add.at(out[s], (p[s][ok], t[ok]), 1)
    return out

  pc, tc, mc = torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(m).cuda()
  c = K.confusion_counts(pc, tc, pk, tk)
  assert np.array_equal(c.cpu().numpy(), ref(np.ones(n)))
  cm = K.confusion_counts(pc, tc, pk, tk, mask=mc)
  assert np.array_equal(cm.cpu().numpy(), ref(m))
  K.confusion_counts(pc, tc, pk, tk, mask=mc, counts=c)  # accumulate
  assert np.array_equal(c.cpu().numpy(), ref(np.ones(n)) + ref(m))


@pytest.mark.parametrize("n,k,seed", [(700, 10, 0), (3000, 10, 1), (90, 3, 2)])
def test_match_functions_equal_the_reference_loops(n, k, seed):
  from iic_b200.utils.cluster import eval_metrics as em
  rng = np.random.RandomState(seed)
  t = rng.randint(0, k, n)
  perm = rng.permutation(k)
  p = np.where(rng.rand(n) < 0.7, perm[t], rng.randint(0, k, n))
  pc, tc = torch.from_numpy(p).int().cuda(), torch.from_numpy(t).int().cuda()
  assert em._original_match(pc, tc, k, k) == oem.original_match(p, t, k, k)
  match = em._hungarian_match(pc, tc, k, k)
  omatch, ocost = oem.hungarian_match(p, t, k, k)
  votes = em.confusion(pc, tc, k, k)
  assert sum(n - votes[a, b] for a, b in match) == ocost
  re = torch.from_numpy(oem.reorder(p, match)).int().cuda()
  assert em._acc(re, tc, k) == oem.acc(oem.reorder(p, match), t, k)
  with pytest.raises(AssertionError):
    em._acc(pc.cpu(), tc.cpu(), k)  # CPU tensors: the reference asserts is_cuda, and there is no fallback


class _FakeNet(torch.nn.Module):
  """S linear sub-heads on the flattened image (test scaffolding: only the evaluation code is under test)."""

  def __init__(self, S, fin, k, seed):
    super().__init__()
    g = torch.Generator().manual_seed(seed)
    self.w = torch.randn(S, fin, k, generator=g).cuda() * 3.0

  def forward(self, x, head="B"):
    f = x.reshape(x.shape[0], -1)
    return [torch.softmax(f @ self.w[i], dim=1) for i in range(self.w.shape[0])]


@pytest.mark.parametrize("mode", ["IID", "IID+"])
def test_cluster_subheads_eval_matches_reference_flow(mode):
  from iic_b200.utils.cluster.cluster_eval import cluster_subheads_eval
  S, k, fin = 3, 10, 16
  cfg = Namespace(output_k=k, gt_k=k, num_sub_heads=S, eval_mode="hung", mode=mode, batch_sz=50, include_rgb=False,
                  mapping_assignment_partitions=["train"], mapping_test_partitions=["train"])
  g = torch.Generator().manual_seed(9)
  protos = torch.randn(k, 1, 4, 4, generator=g)

  def loader(nb, seed):
    gg = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(nb):
      t = torch.randint(0, k, (50,), generator=gg)
      out.append((protos[t] + 0.3 * torch.randn(50, 1, 4, 4, generator=gg), t))
    return out

  net = _FakeNet(S, fin, k, 1)
  assign, test = loader(4, 100), loader(3, 200)
  stats = cluster_subheads_eval(cfg, net, assign, test, sobel=False)

  def flat(ld):
    preds = [[] for _ in range(S)]
    for x, _ in ld:
      outs = net(x.cuda())
      for i in range(S):
        preds[i].append(torch.argmax(outs[i], dim=1).cpu().numpy())
    return [np.concatenate(p) for p in preds], np.concatenate([t.numpy() for _, t in ld])

  pa, ta = flat(assign)
  matches = [oem.hungarian_match(pa[i], ta, k, k)[0] for i in range(S)]
  train = [oem.acc(oem.reorder(pa[i], matches[i]), ta, k) for i in range(S)]
  assert np.allclose(stats["train_accs"], train, rtol=0, atol=1e-7)
  if mode == "IID":
    want = train
  else:
    pt, tt = flat(test)
    # the product's matches (equal-cost ties may be resolved differently from scipy's): relabel with them
    want = [oem.acc(oem.reorder(pt[i], stats["best_train_sub_head_match"] if i == stats["best_train_sub_head"] else matches[i]), tt, k)
            for i in range(S)]
  assert np.allclose(stats["test_accs"], want, rtol=0, atol=1e-7)
  assert stats["best_train_sub_head"] == int(np.argmax(np.array(train, dtype=np.float32)))
  assert stats["best"] == stats["test_accs"][stats["best_train_sub_head"]]
