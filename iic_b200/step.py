"""The per-batch training step of the reference's clustering scripts
(code/scripts/cluster/cluster_sobel_twohead.py:286-355), written against the drop-in API.

    zero_grad -> sobel_process(x), sobel_process(x_tf) -> net(x, head), net(x_tf, head)
    -> IID_loss per sub-head, mean -> backward -> (all-reduce of gradients) -> optimiser.step()

Inputs are the grey (and optionally rgb) image batches the dataloaders hand to the script
(host or device tensors; host tensors are copied on the current stream)."""
import torch

from . import distributed
from .utils.cluster.IID_losses import IID_loss_subheads
from .utils.cluster.transforms import sobel_process


def iic_cluster_step(net, optimiser, imgs, imgs_tf, head="B", lamb=1.0, include_rgb=False, sobel=True,
                     set_to_none=True, pair_batched=True):
  """Returns (avg_loss, avg_loss_no_lamb) as 0-dim device tensors (no host sync)."""
  if optimiser is not None:
    optimiser.zero_grad(set_to_none=set_to_none)
  else:
    net.zero_grad(set_to_none=set_to_none)
  dev = next(net.parameters()).device
  if not imgs.is_cuda:
    imgs = imgs.to(dev, non_blocking=True)
    imgs_tf = imgs_tf.to(dev, non_blocking=True)
  if sobel:
    imgs = sobel_process(imgs, include_rgb)
    imgs_tf = sobel_process(imgs_tf, include_rgb)
  if hasattr(net, "forward_stacked_pair") and pair_batched:
    x_outs, x_tf_outs = net.forward_stacked_pair(imgs, imgs_tf, head=head)
  elif hasattr(net, "forward_stacked"):
    x_outs = net.forward_stacked(imgs, head=head)
    x_tf_outs = net.forward_stacked(imgs_tf, head=head)
  else:
    x_outs, x_tf_outs = net(imgs), net(imgs_tf)
  loss, loss_no_lamb = IID_loss_subheads(x_outs, x_tf_outs, lamb=lamb)
  avg_loss, avg_loss_no_lamb = loss.mean(), loss_no_lamb.mean()
  avg_loss.backward()
  distributed.allreduce_gradients(net.parameters())
  if optimiser is not None:
    optimiser.step()
  return avg_loss.detach(), avg_loss_no_lamb.detach()
