"""a checkpoint written by the reference's own modules (state_dict of the unmodified reference net,
optimiser state of torch.optim.Adam) loads into the iic_b200 network / FusedAdam with the reference's own calls, and
back.  CPU only (modules hold parameters; no kernel runs)."""
import collections
import os
from argparse import Namespace

import pytest
import torch

from oracle import refshim

CFGS = {
  "ClusterNet5gTwoHead": dict(in_channels=2, input_sz=32, num_sub_heads=3, output_k_A=20, output_k_B=5, batchnorm_track=True),
  "ClusterNet6cTwoHead": dict(in_channels=1, input_sz=24, num_sub_heads=2, output_k_A=12, output_k_B=4, batchnorm_track=False),
  "SegmentationNet10aTwoHead": dict(in_channels=5, input_sz=32, num_sub_heads=1, output_k_A=6, output_k_B=3, batchnorm_track=True),
}


@pytest.mark.skipif(not refshim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", sorted(CFGS))
def test_reference_checkpoint_round_trip(name, tmp_path):
  import iic_b200.archs as archs
  from iic_b200.utils.checkpoint import load_into, load_reference_state_dict
  ref = refshim.load()
  torch.manual_seed(3)
  rnet = getattr(ref, name)(Namespace(**CFGS[name]))
  for b in rnet.buffers():  # make the BatchNorm buffers distinguishable from a fresh net's
    if b.dtype.is_floating_point:
      b.add_(torch.rand_like(b))
  path = os.path.join(str(tmp_path), "latest_net.pytorch")
  torch.save(rnet.state_dict(), path)  # cluster_sobel_twohead.py:429
  net = archs.__dict__[name](Namespace(**CFGS[name]))
  net.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage))  # :176-178, verbatim
  want = rnet.state_dict()
  got = net.state_dict()
  assert list(got.keys()) == list(want.keys())
  for k in want:
    assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k
  # saved from a DataParallel wrapper: keys carry a "module." prefix
  path2 = os.path.join(str(tmp_path), "dp.pytorch")
  torch.save(collections.OrderedDict(("module." + k, v) for k, v in want.items()), path2)
  net2 = archs.__dict__[name](Namespace(**CFGS[name]))
  load_into(net2, path2)
  assert all(torch.equal(net2.state_dict()[k], want[k]) for k in want)
  assert list(load_reference_state_dict(path2).keys()) == list(want.keys())
  # and back: the reference module accepts what iic_b200 saves
  torch.save(net.state_dict(), path)
  rnet2 = getattr(ref, name)(Namespace(**CFGS[name]))
  rnet2.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage))
  assert all(torch.equal(rnet2.state_dict()[k], want[k]) for k in want)


def test_adam_state_interchanges_with_torch():
  from iic_b200.optim import FusedAdam
  ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
  ref = torch.optim.Adam(ps, lr=1e-4)
  for p in ps:
    p.grad = torch.randn_like(p)
  ref.step()
  sd = ref.state_dict()
  ours = FusedAdam(ps, lr=1e-4)
  ours.load_state_dict(sd)  # cluster_sobel_twohead.py:187
  out = ours.state_dict()
  assert out["param_groups"][0]["lr"] == 1e-4 and out["param_groups"][0]["betas"] == (0.9, 0.999)
  for i in range(2):
    for key in ("exp_avg", "exp_avg_sq"):
      assert torch.equal(out["state"][i][key], sd["state"][i][key])
    assert float(out["state"][i]["step"]) == 1.0
  back = torch.optim.Adam(ps, lr=1e-4)
  back.load_state_dict(out)


def test_adam_state_written_by_torch_0_4_1_loads():
  """the reference's optimiser checkpoints (torch 0.4.1) keep `step` as a Python int and an `amsgrad`
  key in the param groups; FusedAdam.load_state_dict must accept them (resume path, cluster_sobel_twohead.py:187)."""
  from iic_b200.optim import FusedAdam
  ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
  legacy = {"state": {i: {"step": 7, "exp_avg": torch.randn_like(p), "exp_avg_sq": torch.rand_like(p)}
                      for i, p in enumerate(ps)},
            "param_groups": [{"lr": 1e-4, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                              "params": [0, 1]}]}
  ours = FusedAdam(ps, lr=1e-3)
  ours.load_state_dict(legacy)
  for p in ps:
    st = ours.state[p]
    assert torch.is_tensor(st["step"]) and float(st["step"]) == 7.0
  assert ours.param_groups[0]["lr"] == 1e-4
  # a pickled optimiser (torch.save(optimiser)) goes through __setstate__ as well
  import pickle
  again = pickle.loads(pickle.dumps(ours))
  assert all(torch.is_tensor(again.state[p]["step"]) for p in again.param_groups[0]["params"])
