"""SegmentationNet10a -- drop-in for code/archs/segmentation/net10a.py:13-80 (VGG-style trunk
[(64,1),(128,1),M,(256,1),(256,1),(512,2),(512,2)] 3x3 pad 1 -- the two dilated layers shrink the map --
and num_sub_heads x (conv1x1 padding=1 -> Softmax2d) -> bilinear resize to input_sz)."""
import torch
import torch.nn as nn

from .. import _engine as E
from ... import kernels as K
from ..cluster.net6c import run_vgg_features
from ..cluster.vgg import VGGNet, VGGTrunk

__all__ = ["SegmentationNet10a"]


class SegmentationNet10aTrunk(VGGTrunk):
  def __init__(self, config, cfg):
    super().__init__()
    self.batchnorm_track = config.batchnorm_track
    self.precision = getattr(config, "precision", "bf16")
    assert (config.input_sz % 2 == 0)
    self.conv_size = 3
    self.pad = 1
    self.cfg = cfg
    self.in_channels = config.in_channels if hasattr(config, 'in_channels') else 3
    self.features = self._make_layers()

  def forward(self, x):
    def run(ctx, xin):
      a = run_vgg_features(self, ctx, xin)  # NHWC feature map, do not flatten
      return a, (lambda d: d.contiguous())

    return E.run_trunk(self, run, x)


class _SegHeadFn(torch.autograd.Function):
  @staticmethod
  def forward(ctx, feat, w, input_sz):
    out, zlow = K.seg_head_fwd(feat.contiguous(), w.contiguous(), input_sz, input_sz)
    ctx.save_for_backward(feat, w, zlow)
    return out

  @staticmethod
  def backward(ctx, dout):
    feat, w, zlow = ctx.saved_tensors
    dfeat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
    dw = K.seg_head_bwd(feat.contiguous(), w.contiguous(), zlow, dout.contiguous().float(), dfeat, False)
    return dfeat, dw, None


class SegmentationNet10aHead(nn.Module):
  def __init__(self, config, output_k, cfg):
    super().__init__()
    self.batchnorm_track = config.batchnorm_track
    self.cfg = cfg
    num_features = self.cfg[-1][0]
    self.num_sub_heads = config.num_sub_heads
    # keys heads.{i}.0.weight [k, 512, 1, 1] as in the reference's nn.Sequential(Conv2d, Softmax2d)
    self.heads = nn.ModuleList([nn.Sequential(E.ConvParams(num_features, output_k, 1, 1, 1), E.Identity())
                                for _ in range(self.num_sub_heads)])
    self.input_sz = config.input_sz

  def forward(self, x):
    results = []
    for i in range(self.num_sub_heads):
      w = self.heads[i][0].weight
      results.append(_SegHeadFn.apply(x, w.view(w.shape[0], w.shape[1]), self.input_sz))
    return results


class SegmentationNet10a(VGGNet):
  cfg = [(64, 1), (128, 1), ('M', None), (256, 1), (256, 1), (512, 2), (512, 2)]  # 30x30 recep field

  def __init__(self, config):
    super().__init__()
    self.batchnorm_track = config.batchnorm_track
    self.trunk = SegmentationNet10aTrunk(config, cfg=SegmentationNet10a.cfg)
    self.head = SegmentationNet10aHead(config, output_k=config.output_k, cfg=SegmentationNet10a.cfg)
    self._initialize_weights()

  def forward(self, x):
    x = self.trunk(x)
    x = self.head(x)
    return x
