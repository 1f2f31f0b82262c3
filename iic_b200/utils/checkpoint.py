"""Checkpoint wire format of xu-ji/IIC.

The reference saves ``net.module.state_dict()`` / ``optimiser.state_dict()`` with ``torch.save``
(code/scripts/cluster/cluster_sobel_twohead.py:425-438) and restores them with
``net.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage))`` (:176-178) before ``net.cuda()``.
The iic_b200 networks are ``nn.Module``s holding parameters under the reference's names, shapes and (torch) layouts, and
``FusedAdam`` keeps torch's Adam state layout, so those two calls work unchanged.  This module only deals with what differs
between the files found in the wild: pickles written by Python 2 (the published models) and state dicts saved from a
``DataParallel`` wrapper (keys prefixed with ``module.``)."""
import torch


def _load(path, weights_only):
  try:
    return torch.load(path, map_location="cpu", weights_only=weights_only)
  except UnicodeDecodeError:  # written by Python 2
    return torch.load(path, map_location="cpu", weights_only=weights_only, encoding="latin1")


def load_reference_state_dict(path, trusted=False):
  """torch.load for a reference checkpoint: CPU tensors, Python-2 pickles accepted, ``module.`` prefixes removed.
  A state dict is tensors in an (Ordered)dict, so the restricted unpickler (``weights_only=True``) is enough; the full
  pickle machinery -- which executes code from the file -- is only used when the caller says the file is ``trusted``."""
  try:
    sd = _load(path, True)
  except Exception:  # noqa: BLE001 -- pickle.UnpicklingError and friends; the message of the retry is the useful one
    if not trusted:
      raise
    sd = _load(path, False)
  if not isinstance(sd, dict):
    raise AssertionError("%s does not hold a state dict" % path)
  if sd and all(isinstance(k, str) and k.startswith("module.") for k in sd):
    sd = type(sd)((k[len("module."):], v) for k, v in sd.items())
  return sd


def load_into(net, path, strict=True):
  """``net.load_state_dict`` of a reference checkpoint; raises on missing / unexpected keys or shape mismatches (strict),
  exactly like the reference's own call."""
  return net.load_state_dict(load_reference_state_dict(path), strict=strict)
