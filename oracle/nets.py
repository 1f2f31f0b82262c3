"""Oracle (CPU, test infrastructure only): the reference networks restated as
plain ``torch.nn`` modules with the reference's state_dict key names.

Restates xu-ji/IIC:
  * BasicBlock / _make_layer / init  -> code/archs/cluster/residual.py:10-85
  * ClusterNet5g(TwoHead)            -> code/archs/cluster/net5g.py:10-103, net5g_two_head.py:11-81
  * VGG trunk builder / init         -> code/archs/cluster/vgg.py:4-54
  * ClusterNet6c(TwoHead)            -> code/archs/cluster/net6c.py:10-88, net6c_two_head.py:9-98
  * SegmentationNet10a(TwoHead)      -> code/archs/segmentation/net10a.py:13-80, net10a_twohead.py:8-31

Because the key names and shapes match, ``oracle_net.load_state_dict(
reference_net.state_dict())`` works; tests/golden/make_golden.py uses exactly
that (with named deterministic weights, oracle/weights.py) to pin these
restatements against the real reference modules' outputs and gradients.
"""
import contextlib

import torch
import torch.nn as nn
import torch.nn.functional as F

# ---- optional emulation of the product's bf16 storage points ---------------------------------
# The sm_100a path in bf16 mode (iic_b200, precision="bf16") keeps every NHWC activation and
# every activation-gradient in bf16 and feeds bf16 operands to the tensor cores (fp32 accumulate).
# `bf16_rounding()` makes this oracle round at exactly those points (forward value AND the
# gradient flowing back through the same point), so the bf16 path can be checked against "the
# reference algorithm with bf16 storage" instead of only loosely against the fp32 reference.
_ROUND = {"on": False}


class _RoundBF16(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x):
    return x.bfloat16().to(x.dtype)

  @staticmethod
  def backward(ctx, g):
    return g.bfloat16().to(g.dtype)


def q(x):
  """storage point of an activation (and of its gradient)"""
  return _RoundBF16.apply(x) if _ROUND["on"] else x


def qw(w):
  """tensor-core weight operand (value only; the gradient stays fp32)"""
  return (w + (w.detach().bfloat16().to(w.dtype) - w.detach())) if _ROUND["on"] else w


@contextlib.contextmanager
def bf16_rounding(on=True):
  old = _ROUND["on"]
  _ROUND["on"] = on
  try:
    yield
  finally:
    _ROUND["on"] = old


def _conv(m, x, stem=False):
  w = m.weight if stem else qw(m.weight)
  return q(F.conv2d(x, w, None, m.stride, m.padding, m.dilation))


def _bn(c, track):
  return nn.BatchNorm2d(c, track_running_stats=track)


class _Block(nn.Module):
  """conv3x3-BN-ReLU-conv3x3-BN (+ 1x1/s downsample of the input) -add-ReLU."""

  def __init__(self, cin, cout, stride, track):
    super().__init__()
    self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
    self.bn1 = _bn(cout, track)
    self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
    self.bn2 = _bn(cout, track)
    self.downsample = None
    if stride != 1 or cin != cout:
      self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                      _bn(cout, track))

  def forward(self, x):
    y = q(F.relu(self.bn1(_conv(self.conv1, x))))
    y = self.bn2(_conv(self.conv2, y))
    r = x if self.downsample is None else self.downsample[1](_conv(self.downsample[0], x))
    return q(F.relu(y + r))


def _init_resnet(net):
  for m in net.modules():
    if isinstance(m, nn.Conv2d):
      nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    elif isinstance(m, nn.BatchNorm2d):
      m.weight.data.fill_(1)
      m.bias.data.zero_()
    elif isinstance(m, nn.Linear):
      m.weight.data.normal_(0, 0.01)
      m.bias.data.zero_()


def _init_vgg(net):
  for m in net.modules():
    if isinstance(m, nn.Conv2d):
      nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="relu")
    elif isinstance(m, nn.BatchNorm2d):
      m.weight.data.fill_(1)
      m.bias.data.zero_()
    elif isinstance(m, nn.Linear):
      m.weight.data.normal_(0, 0.01)
      m.bias.data.zero_()


class Trunk5g(nn.Module):
  def __init__(self, config):
    super().__init__()
    t = config.batchnorm_track
    self.conv1 = nn.Conv2d(config.in_channels, 64, 3, 1, 1, bias=False)
    self.bn1 = _bn(64, t)
    self.maxpool = nn.MaxPool2d(2, 2, 1)
    cin = 64
    for li, (planes, nblk, stride) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)], 1):
      blocks = []
      for b in range(nblk):
        blocks.append(_Block(cin, planes, stride if b == 0 else 1, t))
        cin = planes
      setattr(self, "layer%d" % li, nn.Sequential(*blocks))
    self.avgpool = nn.AvgPool2d({96: 7, 64: 5, 32: 3}[config.input_sz], stride=1)

  def forward(self, x, penultimate_features=False):
    x = q(self.maxpool(F.relu(self.bn1(_conv(self.conv1, x, stem=True)))))
    x = self.layer3(self.layer2(self.layer1(x)))
    if not penultimate_features:
      x = self.avgpool(self.layer4(x))
    return x.reshape(x.size(0), -1)


class _SubHeads(nn.Module):
  """num_sub_heads x (Linear -> Softmax(dim=1)); keys ``heads.{i}.0.{weight,bias}``."""

  def __init__(self, nfeat, k, num_sub_heads):
    super().__init__()
    self.heads = nn.ModuleList([nn.Sequential(nn.Linear(nfeat, k), nn.Softmax(dim=1))
                                for _ in range(num_sub_heads)])

  def forward(self, x, kmeans_use_features=False):
    return [x if kmeans_use_features else h(x) for h in self.heads]


class ClusterNet5g(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.trunk = Trunk5g(config)
    self.head = _SubHeads(512, config.output_k, config.num_sub_heads)
    _init_resnet(self)

  def forward(self, x, kmeans_use_features=False, trunk_features=False, penultimate_features=False):
    x = self.trunk(x, penultimate_features=penultimate_features)
    return x if trunk_features else self.head(x, kmeans_use_features)


class ClusterNet5gTwoHead(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.trunk = Trunk5g(config)
    self.head_A = _SubHeads(512, config.output_k_A, config.num_sub_heads)
    self.head_B = _SubHeads(512, config.output_k_B, config.num_sub_heads)
    _init_resnet(self)

  def forward(self, x, head="B", kmeans_use_features=False, trunk_features=False,
              penultimate_features=False):
    x = self.trunk(x, penultimate_features=penultimate_features)
    if trunk_features:
      return x
    assert head in ("A", "B")
    return (self.head_A if head == "A" else self.head_B)(x, kmeans_use_features)


def _vgg_features(cfg, cin, ksz, pad, track):
  layers = []
  for out, dil in cfg:
    if out == "M":
      layers.append(nn.MaxPool2d(2, 2))
    elif out == "A":
      layers.append(nn.AvgPool2d(2, 2))
    else:
      layers += [nn.Conv2d(cin, out, ksz, 1, pad, dilation=dil, bias=False), _bn(out, track),
                 nn.ReLU(inplace=True)]
      cin = out
  return nn.Sequential(*layers)


def _run_vgg(features, x):
  """features = [conv, bn, relu, (pool)]*: the product stores conv outputs and the
  (pooled) activations, so those are the rounding points."""
  first, pending = True, False
  for m in features:
    if isinstance(m, nn.Conv2d):
      if pending:
        x = q(x)
      x = _conv(m, x, stem=first)
      first, pending = False, False
    else:
      x = m(x)
      pending = True
  return q(x) if pending else x


CFG_6C = [(64, 1), ("M", None), (128, 1), ("M", None), (256, 1), ("M", None), (512, 1)]
CFG_10A = [(64, 1), (128, 1), ("M", None), (256, 1), (256, 1), (512, 2), (512, 2)]


class Trunk6c(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.features = _vgg_features(CFG_6C, config.in_channels, 5, 2, config.batchnorm_track)

  def forward(self, x):
    x = _run_vgg(self.features, x)
    return x.reshape(x.size(0), -1)


def _feat6c(config):
  return 512 * {24: 3, 64: 8}[config.input_sz] ** 2


class ClusterNet6c(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.trunk = Trunk6c(config)
    self.head = _SubHeads(_feat6c(config), config.output_k, config.num_sub_heads)
    _init_vgg(self)

  def forward(self, x, kmeans_use_features=False, trunk_features=False, penultimate_features=False):
    assert not penultimate_features
    x = self.trunk(x)
    return x if trunk_features else self.head(x, kmeans_use_features)


class ClusterNet6cTwoHead(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.trunk = Trunk6c(config)
    self.head_A = _SubHeads(_feat6c(config), config.output_k_A, config.num_sub_heads)
    self.head_B = _SubHeads(_feat6c(config), config.output_k_B, config.num_sub_heads)
    _init_vgg(self)

  def forward(self, x, head="B", kmeans_use_features=False, trunk_features=False,
              penultimate_features=False):
    assert not penultimate_features
    x = self.trunk(x)
    if trunk_features:
      return x
    assert head in ("A", "B")
    return (self.head_A if head == "A" else self.head_B)(x, kmeans_use_features)


class Trunk10a(nn.Module):
  def __init__(self, config):
    super().__init__()
    cin = config.in_channels if hasattr(config, "in_channels") else 3
    self.features = _vgg_features(CFG_10A, cin, 3, 1, config.batchnorm_track)

  def forward(self, x):
    return _run_vgg(self.features, x)


class _SegHeads(nn.Module):
  """num_sub_heads x (conv1x1 **padding=1**, no bias -> Softmax2d) then bilinear
  resize to input_sz (align_corners=False); keys ``heads.{i}.0.weight``."""

  def __init__(self, k, config):
    super().__init__()
    self.heads = nn.ModuleList([nn.Sequential(nn.Conv2d(512, k, 1, 1, padding=1, bias=False),
                                              nn.Softmax2d()) for _ in range(config.num_sub_heads)])
    self.input_sz = config.input_sz

  def forward(self, x):
    return [F.interpolate(h(x), size=self.input_sz, mode="bilinear", align_corners=False)
            for h in self.heads]


class SegmentationNet10a(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.trunk = Trunk10a(config)
    self.head = _SegHeads(config.output_k, config)
    _init_vgg(self)

  def forward(self, x):
    return self.head(self.trunk(x))


class SegmentationNet10aTwoHead(nn.Module):
  def __init__(self, config):
    super().__init__()
    self.trunk = Trunk10a(config)
    self.head_A = _SegHeads(config.output_k_A, config)
    self.head_B = _SegHeads(config.output_k_B, config)
    _init_vgg(self)

  def forward(self, x, head="B"):
    x = self.trunk(x)
    assert head in ("A", "B")
    return (self.head_A if head == "A" else self.head_B)(x)
