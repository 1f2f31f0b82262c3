"""Drop-in for xu-ji/IIC ``code/utils/cluster/eval_metrics.py`` (_original_match :9-26, _hungarian_match :29-53,
_acc :56-70).

The reference fills its ``num_correct[c1, c2]`` table with one device reduction and one host synchronisation per pair
(k x k of them: 4 900 for the over-clustering head with k = 70).  Here ONE histogram launch (csrc/eval.cu,
``iic_confusion_counts``) produces the whole table; everything after it is host arithmetic on a k x k matrix, as in the
reference ("num_correct is small", :46).  Same signatures, same return values: lists of ``(out_c, gt_c)`` tuples / a
Python float.  ``sklearn.utils.linear_assignment_`` (removed from scikit-learn) is replaced by
``scipy.optimize.linear_sum_assignment``: same optimal cost; among equal-cost assignments the chosen one may differ.
"""
import numpy as np
import torch

from ... import kernels


def _as_labels(t):
  assert isinstance(t, torch.Tensor) and t.is_cuda, "flat predictions / targets must be CUDA tensors (eval_metrics.py:12-14)"
  return t.reshape(-1).to(torch.int32)


def confusion(flat_preds, flat_targets, preds_k, targets_k):
  """votes[out_c, gt_c] = #samples predicted out_c whose label is gt_c, as a host int64 array (one launch, one copy)."""
  p, t = _as_labels(flat_preds), _as_labels(flat_targets)
  assert p.shape == t.shape
  return kernels.confusion_counts(p, t, preds_k, targets_k)[0].cpu().numpy()


def match_from_votes_original(votes):
  """Many-to-one: every output channel goes to the ground-truth class it co-occurs with most (first one on ties,
  the strict '>' of eval_metrics.py:21)."""
  return [(int(out_c), int(np.argmax(votes[out_c]))) for out_c in range(votes.shape[0])]


def match_from_votes_hungarian(votes, num_samples):
  from scipy.optimize import linear_sum_assignment
  assert votes.shape[0] == votes.shape[1], "one to one (eval_metrics.py:36)"
  rows, cols = linear_sum_assignment(num_samples - votes)
  return [(int(r), int(c)) for r, c in zip(rows, cols)]


def _original_match(flat_preds, flat_targets, preds_k, targets_k):
  return match_from_votes_original(confusion(flat_preds, flat_targets, preds_k, targets_k))


def _hungarian_match(flat_preds, flat_targets, preds_k, targets_k):
  assert preds_k == targets_k
  return match_from_votes_hungarian(confusion(flat_preds, flat_targets, preds_k, targets_k), int(flat_targets.numel()))


def acc_from_votes(votes, match, num_samples):
  """Accuracy after relabelling predictions through `match` (what the reference gets by rewriting flat_preds,
  cluster_eval.py:213-226, and comparing): the relabelled prediction is right exactly on the matched cells."""
  return int(sum(votes[out_c, gt_c] for out_c, gt_c in match)) / float(num_samples)


def _acc(preds, targets, num_k, verbose=0):
  assert isinstance(preds, torch.Tensor) and isinstance(targets, torch.Tensor) and preds.is_cuda and targets.is_cuda
  assert preds.shape == targets.shape
  votes = confusion(preds, targets, num_k, num_k)
  n = int(preds.numel())
  assert int(votes.sum()) == n, "a label is outside [0, num_k) (eval_metrics.py:65)"
  return int(np.trace(votes)) / float(n)


def _nmi(preds, targets):
  from sklearn import metrics
  return metrics.normalized_mutual_info_score(targets, preds)


def _ari(preds, targets):
  from sklearn import metrics
  return metrics.adjusted_rand_score(targets, preds)
