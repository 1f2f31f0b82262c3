"""Drop-in for xu-ji/IIC ``code/utils/segmentation/segmentation_eval.py`` (segmentation_eval :12-40,
_segmentation_get_data :43-140): per-pixel arg-max over the output channels on the device
(csrc/eval.cu), masked flat predictions / targets handed to the shared sub-head evaluation of cluster_eval.py."""
import torch

from ... import kernels
from ..cluster.cluster_eval import cluster_subheads_eval
from ..cluster.transforms import sobel_process


def _segmentation_get_data(config, net, dataloader, sobel=False, using_IR=False, verbose=0):
  """Flat uint8 predictions per sub-head and targets, restricted to the pixels whose mask is set."""
  assert config.output_k <= 255
  preds, targets, masks = [], [], []
  for imgs, flat_targets, mask in dataloader:
    imgs = imgs.cuda()
    if sobel:
      imgs = sobel_process(imgs, config.include_rgb, using_IR=using_IR)
    with torch.no_grad():
      x_outs = net(imgs)
    assert x_outs[0].shape[1] == config.output_k
    assert x_outs[0].shape[2] == config.input_sz and x_outs[0].shape[3] == config.input_sz
    preds.append(torch.stack([kernels.argmax_channels(x.float()).reshape(-1) for x in x_outs]))
    targets.append(flat_targets.cuda().reshape(-1))
    masks.append(mask.cuda().reshape(-1))
  keep = torch.cat(masks).bool()
  flat = torch.cat(preds, dim=1)
  flat_predss_all = [flat[i][keep].to(torch.uint8) for i in range(config.num_sub_heads)]
  flat_targets_all = torch.cat(targets)[keep].to(torch.uint8)
  assert flat_predss_all[0].dim() == 1 and flat_predss_all[0].shape == flat_targets_all.shape
  return flat_predss_all, flat_targets_all


def segmentation_eval(config, net, mapping_assignment_dataloader, mapping_test_dataloader, sobel, using_IR=False, verbose=0,
                      return_only=False):
  torch.cuda.empty_cache()
  net.eval()
  stats = cluster_subheads_eval(config, net, mapping_assignment_dataloader=mapping_assignment_dataloader,
                                mapping_test_dataloader=mapping_test_dataloader, sobel=sobel, using_IR=using_IR,
                                get_data_fn=_segmentation_get_data, verbose=verbose)
  net.train()
  torch.cuda.empty_cache()
  if return_only:
    return stats
  acc = stats["best"]
  is_best = len(config.epoch_acc) > 0 and acc > max(config.epoch_acc)
  config.epoch_stats.append(stats)
  config.epoch_acc.append(acc)
  config.epoch_avg_subhead_acc.append(stats["avg"])
  return is_best
