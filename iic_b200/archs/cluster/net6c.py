"""ClusterNet6c -- drop-in for code/archs/cluster/net6c.py:10-88 (4 x conv5x5-BN-ReLU with 3
max-pools, flatten, num_sub_heads x (Linear + Softmax))."""
import torch.nn as nn

from .. import _engine as E
from ... import kernels as K
from .vgg import VGGNet, VGGTrunk

__all__ = ["ClusterNet6c"]


def run_vgg_features(trunk, ctx, xin):
  """Executes the conv-BN-ReLU(-MaxPool) units of a VGG ``features`` list; first unit is the stem."""
  a = None
  for i, u in enumerate(trunk._plan):
    conv, bn = trunk.features[u["conv"]], trunk.features[u["bn"]]
    pool_pad = 0 if u["pool"] else None
    if i == 0:
      a = E.stem_forward(ctx, conv, bn, xin, pool_pad)
    else:
      a = E.convbn_forward(ctx, conv, bn, a, pool_pad)
  return a


class ClusterNet6cTrunk(VGGTrunk):
  def __init__(self, config):
    super().__init__()
    self.batchnorm_track = config.batchnorm_track
    self.precision = getattr(config, "precision", "bf16")
    self.conv_size = 5
    self.pad = 2
    self.cfg = ClusterNet6c.cfg
    self.in_channels = config.in_channels if hasattr(config, 'in_channels') else 3
    self.features = self._make_layers()

  def forward(self, x, groups=1):
    def run(ctx, xin):
      a = run_vgg_features(self, ctx, xin)
      n, h, w, c = a.shape
      dt = ctx.dt
      # flatten in the reference's NCHW order (net6c.py:26-27)
      flat = K.nhwc_to_nchw(a).reshape(n, -1)

      def finisher(dfeat):
        return K.nchw_to_nhwc(dfeat.float().reshape(n, c, h, w).contiguous(), dt)

      return flat, finisher

    return E.run_trunk(self, run, x, groups)


def _feat_size(config, cfg):
  if config.input_sz == 24:
    features_sp_size = 3
  elif config.input_sz == 64:
    features_sp_size = 8
  return cfg[-1][0] * features_sp_size * features_sp_size


class ClusterNet6cHead(E.SubHeads):
  def __init__(self, config):
    super().__init__(_feat_size(config, ClusterNet6c.cfg), config.output_k, config.num_sub_heads)
    self.batchnorm_track = config.batchnorm_track
    self.cfg = ClusterNet6c.cfg


class ClusterNet6c(VGGNet):
  cfg = [(64, 1), ('M', None), (128, 1), ('M', None), (256, 1), ('M', None), (512, 1)]

  def __init__(self, config):
    super().__init__()
    self.batchnorm_track = config.batchnorm_track
    self.trunk = ClusterNet6cTrunk(config)
    self.head = ClusterNet6cHead(config)
    self._initialize_weights()

  def forward(self, x, kmeans_use_features=False, trunk_features=False, penultimate_features=False):
    if penultimate_features:
      print("Not needed/implemented for this arch")
      exit(1)
    x = self.trunk(x)
    if trunk_features:  # for semisup
      return x
    return self.head(x, kmeans_use_features=kmeans_use_features)  # returns list
