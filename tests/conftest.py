import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden(dict):
  def sub(self, prefix):
    prefix = prefix.rstrip("/") + "/"
    return Golden({k[len(prefix):]: v for k, v in self.items() if k.startswith(prefix)})

  def names(self, prefix=""):
    pl = len(prefix)
    return sorted({k[pl:].split("/")[0] for k in self if k.startswith(prefix)})


def load_golden(fname):
  with np.load(os.path.join(GOLDEN, fname), allow_pickle=False) as z:
    return Golden({k.replace("|", "/"): z[k] for k in z.files})


@pytest.fixture(scope="session")
def golden_iid():
  return load_golden("iid_loss.npz")


@pytest.fixture(scope="session")
def golden_seg():
  return load_golden("seg_loss.npz")


@pytest.fixture(scope="session")
def golden_nets():
  return load_golden("nets.npz")
