"""Drop-in for xu-ji/IIC ``code/utils/cluster/cluster_eval.py``: same function names, arguments and
result dictionaries; the device work goes through csrc/eval.cu (arg-max of all sub-heads in one launch, one
confusion-count launch per data set instead of k x k reductions per sub-head) and the fused IID loss kernel.

  _clustering_get_data         :15-75     net outputs -> flat predictions per sub-head + targets
  cluster_subheads_eval        :78-149    matches from the assignment set, accuracies on the test set
  _get_assignment_data_matches :152-233
  get_subhead_using_loss       :236-317   sub-head with the lowest IID loss, no labels
  cluster_eval                 :320-369   eval-mode wrapper that updates config.epoch_*
"""
import sys

import numpy as np
import torch

from ... import kernels
from . import eval_metrics as em
from .IID_losses import IID_loss_subheads
from .transforms import sobel_process


def _stack(outs):
  return outs if isinstance(outs, torch.Tensor) else torch.stack(list(outs))


def _clustering_get_data(config, net, dataloader, sobel=False, using_IR=False, get_soft=False, verbose=0):
  """Flat int32 predictions (one tensor per sub-head) and targets of a whole data loader, on the GPU."""
  assert not using_IR  # IR is a segmentation input (cluster_eval.py:21)
  preds, soft, targets = [], [], []
  for batch in dataloader:
    imgs = batch[0].cuda()
    if sobel:
      imgs = sobel_process(imgs, config.include_rgb, using_IR=using_IR)
    with torch.no_grad():
      z = _stack(net(imgs))  # [S, n, k]
    assert z.dim() == 3 and z.shape[2] == config.output_k
    preds.append(kernels.argmax_rows(z))
    targets.append(batch[1].cuda().reshape(-1).to(torch.int32))
    if get_soft:
      soft.append(z)
  flat = torch.cat(preds, dim=1)
  flat_predss_all = [flat[i] for i in range(config.num_sub_heads)]
  flat_targets_all = torch.cat(targets)
  if not get_soft:
    return flat_predss_all, flat_targets_all
  s = torch.cat(soft, dim=1)
  return flat_predss_all, flat_targets_all, [s[i] for i in range(config.num_sub_heads)]


def _votes(flat_predss_all, flat_targets_all, preds_k, targets_k):
  """[S, preds_k, targets_k] co-occurrence counts of every sub-head: one launch, one device-to-host copy."""
  p = torch.stack([x.reshape(-1).to(torch.int32) for x in flat_predss_all])
  return kernels.confusion_counts(p, flat_targets_all.reshape(-1).to(torch.int32), preds_k, targets_k).cpu().numpy()


def _match(votes, mode, num_samples):
  if mode == "hung":
    return em.match_from_votes_hungarian(votes, num_samples)
  if mode == "orig":
    return em.match_from_votes_original(votes)
  assert False, "config.eval_mode must be 'hung' or 'orig' (cluster_eval.py:195-205)"


def _get_assignment_data_matches(net, mapping_assignment_dataloader, config, sobel=False, using_IR=False, get_data_fn=None,
                                 just_matches=False, verbose=0):
  """Best match per sub-head on the assignment set, and its accuracy there."""
  flat_predss_all, flat_targets_all = get_data_fn(config, net, mapping_assignment_dataloader, sobel=sobel,
                                                  using_IR=using_IR, verbose=verbose)
  assert flat_predss_all[0].shape == flat_targets_all.shape
  n = int(flat_targets_all.shape[0])
  votes = _votes(flat_predss_all, flat_targets_all, config.output_k, config.gt_k)
  all_matches = [_match(votes[i], config.eval_mode, n) for i in range(config.num_sub_heads)]
  if just_matches:
    return all_matches
  all_accs = np.zeros(config.num_sub_heads, dtype=np.float32)
  for i, match in enumerate(all_matches):
    assert len({out_c for out_c, _ in match}) == config.output_k  # each output_k must get mapped (:219)
    all_accs[i] = em.acc_from_votes(votes[i], match, n)
  return all_matches, all_accs


def cluster_subheads_eval(config, net, mapping_assignment_dataloader, mapping_test_dataloader, sobel, using_IR=False,
                          get_data_fn=_clustering_get_data, use_sub_head=None, verbose=0):
  """Accuracy of every sub-head (matches made on the assignment data, measured on the test data); `best` is the
  sub-head that is best on the assignment data, or `use_sub_head`."""
  all_matches, train_accs = _get_assignment_data_matches(net, mapping_assignment_dataloader, config, sobel=sobel,
                                                         using_IR=using_IR, get_data_fn=get_data_fn, verbose=verbose)
  best_sub_head = int(np.argmax(train_accs))
  if config.num_sub_heads > 1 and use_sub_head is not None:
    best_sub_head = use_sub_head
  if config.mode == "IID":
    assert config.mapping_assignment_partitions == config.mapping_test_partitions
    test_accs = train_accs
  elif config.mode == "IID+":
    flat_predss_all, flat_targets_all = get_data_fn(config, net, mapping_test_dataloader, sobel=sobel, using_IR=using_IR,
                                                    verbose=verbose)
    n = int(flat_targets_all.shape[0])
    votes = _votes(flat_predss_all, flat_targets_all, config.output_k, config.gt_k)
    test_accs = np.array([em.acc_from_votes(votes[i], all_matches[i], n) for i in range(config.num_sub_heads)],
                         dtype=np.float32)
  else:
    assert False
  return {"test_accs": list(test_accs), "avg": np.mean(test_accs), "std": np.std(test_accs),
          "best": test_accs[best_sub_head], "worst": test_accs.min(), "best_train_sub_head": best_sub_head,
          "best_train_sub_head_match": all_matches[best_sub_head], "train_accs": list(train_accs)}


def get_subhead_using_loss(config, dataloaders_head_B, net, sobel, lamb, compare=False):
  """Index of the sub-head of head B with the lowest summed IID loss over the data (no labels used)."""
  net.eval()
  module = net.module if hasattr(net, "module") else net
  loss_per_sub_head = np.zeros(config.num_sub_heads)
  for b_i, tup in enumerate(zip(*dataloaders_head_B)):
    module.zero_grad()
    plain = tup[0][0]
    # the reference pairs every transformed loader with the SAME plain batch (cluster_eval.py:262-274)
    all_imgs = torch.cat([plain.cuda() for _ in range(config.num_dataloaders)])
    all_imgs_tf = torch.cat([tup[1 + d][0].cuda() for d in range(config.num_dataloaders)])
    assert all_imgs.shape == all_imgs_tf.shape
    if sobel:
      all_imgs = sobel_process(all_imgs, config.include_rgb)
      all_imgs_tf = sobel_process(all_imgs_tf, config.include_rgb)
    with torch.no_grad():
      z, zt = _stack(net(all_imgs, head="B")), _stack(net(all_imgs_tf, head="B"))
      loss, _ = IID_loss_subheads(z, zt, lamb=lamb)  # all sub-heads in one launch
    loss_per_sub_head += loss.detach().cpu().numpy().astype(np.float64)
    if b_i % 100 == 0:
      print("at batch %d" % b_i)
      sys.stdout.flush()
  best_sub_head_loss = int(np.argmin(loss_per_sub_head))
  if compare:
    print(loss_per_sub_head)
    print("best sub_head by loss: %d" % best_sub_head_loss)
    best_epoch = int(np.argmax(np.array(config.epoch_acc)))
    stats = config.epoch_stats[best_epoch]
    by_eval = stats["best_train_sub_head"] if "best_train_sub_head" in stats else stats["best_head"]
    accs = stats["test_accs"] if "test_accs" in stats else stats["all"]
    print("best sub_head by eval: %d" % by_eval)
    print("... loss select acc: %f, eval select acc: %f" % (accs[best_sub_head_loss], accs[by_eval]))
  net.train()
  return best_sub_head_loss


def cluster_eval(config, net, mapping_assignment_dataloader, mapping_test_dataloader, sobel, use_sub_head=None,
                 print_stats=False):
  if config.double_eval:  # evaluation in train mode as well: BatchNorm then uses batch statistics (:323-341)
    stats2 = cluster_subheads_eval(config, net, mapping_assignment_dataloader=mapping_assignment_dataloader,
                                   mapping_test_dataloader=mapping_test_dataloader, sobel=sobel, use_sub_head=use_sub_head)
    if print_stats:
      print("double eval stats:")
      print(stats2)
    else:
      config.double_eval_stats.append(stats2)
      config.double_eval_acc.append(stats2["best"])
      config.double_eval_avg_subhead_acc.append(stats2["avg"])
  net.eval()
  stats = cluster_subheads_eval(config, net, mapping_assignment_dataloader=mapping_assignment_dataloader,
                                mapping_test_dataloader=mapping_test_dataloader, sobel=sobel, use_sub_head=use_sub_head)
  net.train()
  if print_stats:
    print("eval stats:")
    print(stats)
    return None
  acc = stats["best"]
  is_best = len(config.epoch_acc) > 0 and acc > max(config.epoch_acc)
  config.epoch_stats.append(stats)
  config.epoch_acc.append(acc)
  config.epoch_avg_subhead_acc.append(stats["avg"])
  return is_best
