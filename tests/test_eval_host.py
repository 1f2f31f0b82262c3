"""Host arithmetic of the evaluation drop-in (iic_b200/utils/cluster/eval_metrics.py) against the
oracle's loop-for-loop restatement of the reference, on CPU: matches and accuracies computed from a votes table must
equal what the reference gets from its k x k masked reductions and its relabelling loop."""
import numpy as np
import pytest

from oracle import eval_metrics as oem


def _votes(p, t, pk, tk):
  v = np.zeros((pk, tk), dtype=np.int64)
  np.add.at(v, (p, t), 1)
  return v


def _em():
  import importlib.util
  import os
  import sys
  import types
  # import the host module without loading the CUDA library: give it a stub `kernels` (the functions under test take
  # a ready votes table)
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  name = "_iic_eval_metrics_host_only"
  if name in sys.modules:
    return sys.modules[name]
  src = open(os.path.join(root, "iic_b200", "utils", "cluster", "eval_metrics.py")).read()
  src = src.replace("from ... import kernels", "kernels = None")
  mod = types.ModuleType(name)
  exec(compile(src, "eval_metrics.py", "exec"), mod.__dict__)
  sys.modules[name] = mod
  return mod


@pytest.mark.parametrize("n,pk,tk,seed", [(500, 10, 10, 0), (2000, 70, 10, 1), (300, 5, 3, 2), (50, 10, 10, 3)])
def test_original_match_and_accuracy_from_votes(n, pk, tk, seed):
  em = _em()
  rng = np.random.RandomState(seed)
  t = rng.randint(0, tk, n)
  p = (t * 3 + rng.randint(0, pk, n) * (rng.rand(n) < 0.4)) % pk  # correlated with the labels, with noise
  votes = _votes(p, t, pk, tk)
  match = em.match_from_votes_original(votes)
  assert match == oem.original_match(p, t, pk, tk)
  assert em.acc_from_votes(votes, match, n) == oem.acc(oem.reorder(p, match), t, max(pk, tk))


@pytest.mark.parametrize("n,k,seed", [(700, 10, 0), (3000, 10, 1), (90, 3, 2), (40, 10, 4)])
def test_hungarian_match_and_accuracy_from_votes(n, k, seed):
  em = _em()
  rng = np.random.RandomState(seed)
  t = rng.randint(0, k, n)
  perm = rng.permutation(k)
  p = np.where(rng.rand(n) < 0.7, perm[t], rng.randint(0, k, n))
  votes = _votes(p, t, k, k)
  assert np.array_equal(votes, oem.num_correct_table(p, t, k).astype(np.int64))
  match = em.match_from_votes_hungarian(votes, n)
  omatch, ocost = oem.hungarian_match(p, t, k, k)
  cost = sum(n - votes[a, b] for a, b in match)
  assert cost == ocost  # same optimum (the assignment itself may differ among ties)
  assert sorted(a for a, _ in match) == list(range(k)) and sorted(b for _, b in match) == list(range(k))
  assert em.acc_from_votes(votes, match, n) == oem.acc(oem.reorder(p, match), t, k)
  assert em.acc_from_votes(votes, omatch, n) == oem.acc(oem.reorder(p, omatch), t, k)


def test_perfect_clustering_is_accuracy_one():
  em = _em()
  t = np.repeat(np.arange(10), 70)
  p = (t + 3) % 10
  votes = _votes(p, t, 10, 10)
  m = em.match_from_votes_hungarian(votes, t.size)
  assert em.acc_from_votes(votes, m, t.size) == 1.0 and dict(m) == {(c + 3) % 10: c for c in range(10)}
