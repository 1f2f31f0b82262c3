"""CPU restatement (numpy) of the reference's evaluation arithmetic -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, loop for loop,
  code/utils/cluster/eval_metrics.py:9-26    _original_match
  code/utils/cluster/eval_metrics.py:29-53   _hungarian_match
  code/utils/cluster/eval_metrics.py:56-70   _acc
  code/utils/cluster/cluster_eval.py:209-226 relabelling of the predictions through a match
Parity unpinned by execution: the reference module cannot be imported here (it needs
``sklearn.utils.linear_assignment_``, removed from scikit-learn, and asserts CUDA tensors), and the reference has no
tests; the assignment problem is solved with scipy's ``linear_sum_assignment`` (same optimal cost).
"""
import numpy as np


def original_match(flat_preds, flat_targets, preds_k, targets_k):
  out_to_gts, out_to_gts_scores = {}, {}
  for out_c in range(preds_k):
    for gt_c in range(targets_k):
      tp_score = int(((flat_preds == out_c) * (flat_targets == gt_c)).sum())  # :20
      if (out_c not in out_to_gts) or (tp_score > out_to_gts_scores[out_c]):   # :21
        out_to_gts[out_c] = gt_c
        out_to_gts_scores[out_c] = tp_score
  return sorted(out_to_gts.items())


def num_correct_table(flat_preds, flat_targets, num_k):
  num_correct = np.zeros((num_k, num_k))
  for c1 in range(num_k):
    for c2 in range(num_k):
      num_correct[c1, c2] = int(((flat_preds == c1) * (flat_targets == c2)).sum())  # :43
  return num_correct


def hungarian_match(flat_preds, flat_targets, preds_k, targets_k):
  from scipy.optimize import linear_sum_assignment
  assert preds_k == targets_k  # :36
  num_samples = flat_targets.shape[0]
  cost = num_samples - num_correct_table(flat_preds, flat_targets, preds_k)  # :47
  rows, cols = linear_sum_assignment(cost)
  return [(int(r), int(c)) for r, c in zip(rows, cols)], float(cost[rows, cols].sum())


def acc(preds, targets, num_k):
  assert preds.shape == targets.shape
  assert preds.max() < num_k and targets.max() < num_k  # :65
  return int((preds == targets).sum()) / float(preds.shape[0])  # :67


def reorder(flat_preds, match):
  """cluster_eval.py:213-216: predictions relabelled through (pred_i, target_i) pairs; unmatched values become 0."""
  reordered = np.zeros(flat_preds.shape[0], dtype=flat_preds.dtype)
  for pred_i, target_i in match:
    reordered[flat_preds == pred_i] = target_i
  return reordered
