#!/usr/bin/env python
"""Benchmark of the IIC training hot path (BASELINE.json metric): img-pairs/s through
ClusterNet5gTwoHead + IID_loss (k=10 head B, 5 sub-heads) on synthetic 96x96 batches.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference ...                     # the reference algorithm on the host CPUs

A step = one pass of the reference's per-batch loop (cluster_sobel_twohead.py:286-355):
zero_grad, sobel x2, net(x), net(x_tf), IID loss over 5 sub-heads, backward, gradient all-reduce
(N>1), Adam step.  `value` times it with the grey image batches already resident in HBM;
`e2e` times the same call with PINNED HOST batches (H2D inside the timed region, loss read back).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONV_FLOP_PER_PAIR_96 = 36342890496.0  # 2*MAC, fprop+dgrad+wgrad of the 36 convs, two views
METRIC = "img-pairs/s through ClusterNet5g+IID_loss at 1/2/4/8 B200 vs CPU ref"


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--pairs-per-gpu", type=int, default=704)
  ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
  ap.add_argument("--head", default="B")
  ap.add_argument("--cpu-pairs", type=int, default=32, help="bounded CPU sample: image pairs per reference step")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-roofline", action="store_true")
  return ap.parse_args()


def net_config(precision=None):
  from argparse import Namespace
  cfg = dict(in_channels=2, input_sz=96, num_sub_heads=5, output_k_A=70, output_k_B=10, batchnorm_track=True)
  if precision is not None:
    cfg["precision"] = precision
  return Namespace(**cfg)


def peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      p = json.load(f)
    return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, sustained bf16)")
  return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                    "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], None, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[0]))
        mx = float(r[1])
        for nme, v in zip(names, r[3:7]):
          if v.lower().startswith("active"):
            reasons.add(nme)
      except Exception:
        pass
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port of xu-ji/IIC) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_step_fn(pairs, head, threads=None):
  import torch

  from oracle import iid_losses as oracle_iid
  from oracle import nets as oracle_nets
  from oracle import transforms as oracle_tf
  if threads:
    torch.set_num_threads(threads)
  torch.manual_seed(0)
  net = oracle_nets.ClusterNet5gTwoHead(net_config())
  net.train()
  opt = torch.optim.Adam(net.parameters(), lr=1e-4)
  grey = torch.rand(pairs, 1, 96, 96)
  grey_tf = torch.rand(pairs, 1, 96, 96)

  def step():
    opt.zero_grad()
    x, xt = oracle_tf.sobel_process(grey, False), oracle_tf.sobel_process(grey_tf, False)
    o, ot = net(x, head=head), net(xt, head=head)
    loss = sum(oracle_iid.IID_loss(a, b)[0] for a, b in zip(o, ot)) / len(o)
    loss.backward()
    opt.step()
    return float(loss)

  return step


def time_cpu_reference(pairs, head, steps, warmup, budget_s=25.0):
  import torch
  step = cpu_reference_step_fn(pairs, head)
  t0 = time.time()
  for _ in range(max(1, warmup)):
    step()
  per = (time.time() - t0) / max(1, warmup)
  n = max(1, min(steps, int(budget_s / max(per, 1e-3))))
  t0 = time.time()
  for _ in range(n):
    step()
  dt = (time.time() - t0) / n
  return dict(value=pairs / dt, unit="img-pairs/s", cores=torch.get_num_threads(), kind="port",
              sample="%d steps of %d img-pairs (96x96), oracle port of the reference (torch CPU fp32), "
                     "sobel+fwd x2+IID_loss x5+bwd+Adam" % (n, pairs)), dt, n


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  base, dt, n = time_cpu_reference(args.cpu_pairs, args.head, args.steps, min(args.warmup, 1), budget_s=120.0)
  line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "img-pairs/s", "n_gpus": args.gpus,
          "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": workload_config(args, args.cpu_pairs, 1), "cpu_baseline": base,
          "e2e": {"value": base["value"], "unit": "img-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  print(json.dumps(line))


def workload_config(args, pairs_per_gpu, n):
  return {"workload": "STL10-shape IIC step: ClusterNet5gTwoHead(96x96, sobel 2ch, 5 sub-heads, k_A=70/k_B=10) head %s, "
                      "IID_loss lamb=1, Adam" % args.head,
          "pairs_per_gpu": pairs_per_gpu, "global_batch": pairs_per_gpu * n, "input": "96x96x1 grey pair -> sobel",
          "parallelism": "dp%d (pairs sharded, per-rank BN, allreduce of [5,k,k] joint + weight grads)" % n,
          "l2": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist

  import iic_b200.archs as archs
  from iic_b200 import distributed as iicd
  from iic_b200 import kernels
  from iic_b200.optim import FusedAdam
  from iic_b200.step import iic_cluster_step

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    iicd.enable()
  assert world == args.gpus, "--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world)

  B = args.pairs_per_gpu
  torch.manual_seed(0)
  net = archs.ClusterNet5gTwoHead(net_config(args.precision)).to(dev)
  net.train()
  opt = FusedAdam(net.parameters(), lr=1e-4)
  g = torch.Generator().manual_seed(1000 + rank)
  host = [torch.rand(B, 1, 96, 96, generator=g).pin_memory() for _ in range(2)]
  resident = [h.to(dev) for h in host]

  def step(src):
    return iic_cluster_step(net, opt, src[0], src[1], head=args.head, lamb=1.0)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(src, steps, read_back):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    last = None
    for _ in range(steps):
      last = step(src)
      if read_back:
        last = (last[0].item(), last[1].item())  # device -> host read of the step's result
    e1.record()
    barrier()
    wall = time.time() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() / 1e3, wall, last

  for _ in range(args.warmup):
    step(resident)
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  kernels.launch_count(reset=True)
  sec, wall, last = timed(resident, args.steps, False)
  launches = kernels.launch_count()
  clocks = sampler.stop() if rank == 0 else None
  value = B * world * args.steps / sec

  # end to end: pinned host batches in, loss scalars out, every step
  step(host)
  sec_e, wall_e, last_e = timed(host, args.steps, True)
  e2e = B * world * args.steps / sec_e

  roof = None
  if not args.no_roofline:
    kernels.conv_timing(True)
    for _ in range(2):
      step(resident)
    torch.cuda.synchronize()
    summ = kernels.conv_timing_summary()
    kernels.conv_timing(False)
    pk = peaks()
    other = {k: {"launches": v[0] // 2, "ms_per_step": v[2] / 2} for k, v in summ.items() if v[1] == 0.0}
    summ = {k: v for k, v in summ.items() if v[1] > 0.0}
    flops = sum(v[1] for v in summ.values())
    ms = sum(v[2] for v in summ.values())
    nl = sum(v[0] for v in summ.values())
    ach = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM fprop/dgrad/wgrad)",
            "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
            # DRAM bytes per launch of the family's most frequent member from the committed `ncu --set full` capture
            # (not measured in this run): conv_halo_kernel fprop, 352 images of 49x49x64 -> 108.5 MB read + 65.4 MB
            # written against 216.4 MB algorithmic (input once + output once; part of the output was still in L2)
            "traffic": 173.9e6, "traffic_unit": "B per launch (ncu, profiles/r01_ncu_halo.md: conv_halo_kernel fprop at 352 images; "
                                                "algorithmic 216.4e6 B)",
            "peak_source": pk["src"], "launches_timed": nl,
            "flop_per_launch_avg": flops / max(nl, 1), "ms_per_launch_avg": ms / max(nl, 1),
            "by_kind": {k: {"launches": v[0], "tflops": v[1] / (v[2] * 1e-3) / 1e12 if v[2] > 0 else 0.0,
                            "ms_per_step": v[2] / 2} for k, v in summ.items()},
            "conv_share_of_step": (ms / 2) / (sec / args.steps * 1e3), "other_kernels_ms_per_step": other,
            "whole_step_frac": value / world * CONV_FLOP_PER_PAIR_96 / 1e12 / pk["tflops"]}

  cpu = None
  if rank == 0 and not args.no_cpu_baseline and world == 1:
    cpu, _, _ = time_cpu_reference(args.cpu_pairs, args.head, 3, 1, budget_s=20.0)

  if rank == 0:
    from iic_b200.archs import _engine
    variants = {k: kernels.get_option(k) for k in ("conv_halo", "conv_halo_wgrad", "conv_halo_store", "stem_quad", "dgrad_prefetch", "tc2_mt2", "tc_cpasync")}
    variants.update({k: int(v) for k, v in _engine.OPTIONS.items()})
    line = {"metric": METRIC, "value": value, "unit": "img-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, B, world), "clocks": clocks,
            "e2e": {"value": e2e, "unit": "img-pairs/s", "h2d_bytes_per_step": 2 * B * 96 * 96 * 4 * world,
                    "d2h_bytes_per_step": 8 * world, "ms_per_step": sec_e / args.steps * 1e3},
            "gpu_launches": launches, "loss": float(last[0]), "wall_s": wall, "kernel_variants": variants}
    if roof is not None:
      line["roofline"] = roof
    if cpu is not None:
      line["cpu_baseline"] = cpu
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
  if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
  args = parse()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
