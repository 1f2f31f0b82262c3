"""ResNet plumbing of the reference (code/archs/cluster/residual.py:10-85) as parameter
containers + plan builders; the math runs in the sm_100a kernels via ``_engine``."""
import torch.nn as nn

from .._engine import BNParams, ConvParams, initialize_weights


class BasicBlock(nn.Module):
  """conv3x3-BN-ReLU-conv3x3-BN (+1x1/s conv-BN downsample of the input), add, ReLU
  (reference residual.py:10-43).  Executed by ``_engine.block_forward``."""
  expansion = 1

  def __init__(self, inplanes, planes, stride=1, downsample=None, track_running_stats=None):
    super().__init__()
    assert (track_running_stats is not None)
    self.conv1 = ConvParams(inplanes, planes, 3, stride, 1)
    self.bn1 = BNParams(planes, track_running_stats)
    self.conv2 = ConvParams(planes, planes, 3, 1, 1)
    self.bn2 = BNParams(planes, track_running_stats)
    self.downsample = downsample
    self.stride = stride


class ResNetTrunk(nn.Module):
  def _make_layer(self, block, planes, blocks, stride=1):
    downsample = None
    if stride != 1 or self.inplanes != planes * block.expansion:
      downsample = nn.Sequential(ConvParams(self.inplanes, planes * block.expansion, 1, stride, 0),
                                 BNParams(planes * block.expansion, self.batchnorm_track))
    layers = [block(self.inplanes, planes, stride, downsample, track_running_stats=self.batchnorm_track)]
    self.inplanes = planes * block.expansion
    for _ in range(1, blocks):
      layers.append(block(self.inplanes, planes, track_running_stats=self.batchnorm_track))
    return nn.Sequential(*layers)


class ResNet(nn.Module):
  def _initialize_weights(self):
    initialize_weights(self, "fan_out")  # residual.py:75-85
