"""VGG plumbing of the reference (code/archs/cluster/vgg.py:4-54): the ``features`` Sequential
keeps the reference's indices (conv, BN, ReLU, pool entries) so state_dict keys match."""
import torch.nn as nn

from .._engine import BNParams, ConvParams, Identity, initialize_weights


class VGGTrunk(nn.Module):
  def _make_layers(self, batch_norm=True):
    assert batch_norm, "the IIC configurations always use batch norm"
    layers, plan = [], []
    in_channels = self.in_channels
    for tup in self.cfg:
      assert (len(tup) == 2)
      out, dilation = tup
      if out == 'M':
        layers += [Identity()]
        plan[-1]["pool"] = True  # fused into the preceding conv-BN-ReLU unit
      elif out == 'A':
        raise NotImplementedError("'A' (avg-pool) entries are not used by any IIC configuration")
      else:
        conv = ConvParams(in_channels, out, self.conv_size, 1, self.pad, dilation)
        bn = BNParams(out, self.batchnorm_track)
        plan.append({"conv": len(layers), "bn": len(layers) + 1, "pool": False})
        layers += [conv, bn, Identity()]
        in_channels = out
    self._plan = plan
    return nn.Sequential(*layers)


class VGGNet(nn.Module):
  def _initialize_weights(self, mode='fan_in'):
    initialize_weights(self, mode)  # vgg.py:42-54
