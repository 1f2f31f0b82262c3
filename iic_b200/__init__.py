"""iic_b200: B200-native (sm_100a) implementation of the IIC training hot path.

Drop-in mirror of the reference's Python interface for that path:

    from iic_b200.utils.cluster.IID_losses import IID_loss, compute_joint
    from iic_b200.utils.cluster.transforms import sobel_process
    from iic_b200.utils.segmentation.IID_losses import IID_segmentation_loss, IID_segmentation_loss_uncollapsed
    import iic_b200.archs as archs;  net = archs.__dict__[config.arch](config)

All device work is done by hand-written CUDA kernels in libiic_b200.so (C-ABI in
include/iic_b200.h) called through ctypes.  PyTorch supplies device memory,
streams, autograd bookkeeping and torch.distributed -- plumbing only.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
