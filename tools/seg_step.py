"""C5-shape segmentation step (SegmentationNet10aTwoHead 128x128x5, k_A=15, T=10, uncollapsed loss):
python tools/seg_step.py [pairs=15] [head=A] -> ms per step with a loss/net breakdown."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
import iic_b200.archs as archs
from iic_b200.utils.segmentation.IID_losses import IID_segmentation_loss_uncollapsed, IID_segmentation_loss

n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
head = sys.argv[2] if len(sys.argv) > 2 else "A"
cfg = Namespace(in_channels=5, input_sz=128, num_sub_heads=1, output_k_A=15, output_k_B=3, batchnorm_track=True, precision="bf16")
torch.manual_seed(0)
net = archs.SegmentationNet10aTwoHead(cfg).cuda().train()
x1, x2 = torch.rand(n, 5, 128, 128, device="cuda"), torch.rand(n, 5, 128, 128, device="cuda")
theta = torch.tensor([[1., 0, 0], [0, 1, 0]], device="cuda").repeat(n, 1, 1)
theta[::2, 0, 0] = -1
mask = (torch.rand(n, 128, 128, device="cuda") < 0.7).float()


def step(loss_fn, timing=None):
  net.zero_grad(set_to_none=True)
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
  ev[0].record()
  o, ot = net(x1, head=head)[0], net(x2, head=head)[0]
  ev[1].record()
  l, _ = loss_fn(o, ot, all_affine2_to_1=theta, all_mask_img1=mask, lamb=1.0, half_T_side_dense=10,
                 half_T_side_sparse_min=0, half_T_side_sparse_max=0)
  ev[2].record()
  l.backward()
  ev[3].record()
  if timing is not None:
    torch.cuda.synchronize()
    timing.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
  return l


out = {}
for name, fn in [("uncollapsed", IID_segmentation_loss_uncollapsed), ("collapsed", IID_segmentation_loss)]:
  for _ in range(2):
    step(fn)
  t = []
  for _ in range(5):
    step(fn, t)
  f, l, b = [sum(x[i] for x in t) / len(t) for i in range(3)]
  k = 15 if head == "A" else 3
  flop_loss = 2.0 * k * k * 441 * n * 128 * 128 * (3 if name == "uncollapsed" else 0)
  out[name] = dict(pairs=n, head=head, ms_forward_nets=f, ms_loss_fwd_and_grad=l, ms_backward_nets=b, ms_step=f + l + b,
                   pairs_per_s=n / (f + l + b) * 1e3, loss_tflops=(flop_loss / (l * 1e-3) / 1e12) if flop_loss else None)
print(json.dumps(out))
