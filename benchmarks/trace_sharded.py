#!/usr/bin/env python
"""torchrun --nproc-per-node N benchmarks/trace_sharded.py : device timeline (torch profiler, CUPTI)
of one sharded renumber step on rank 0 -- kernel / memcpy names with start offsets and durations."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
  rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  from fastremap_b200 import api, sharded
  from fastremap_b200.synth import synth_device
  S = 1024
  n = S ** 3
  vol = synth_device((S, S, S), (46 * world, 46, 46), np.uint64, seed=0, zero_frac=0.1, z0=rank * S, nz_total=S * world).reshape(-1).view(torch.uint8)
  work = torch.empty_like(vol)

  def step():
    if world > 1:
      return sharded.renumber_sharded(work, n, np.uint64, write_wide=True, offset=rank * n, total=world * n)
    return api.renumber_device(work, n, np.uint64, write_wide=True)

  for _ in range(3):
    work.copy_(vol)
    step()
  torch.cuda.synchronize()
  from torch.profiler import profile, ProfilerActivity
  with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    work.copy_(vol)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
  if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = None
    for e in evs:
      if "k_table_init" in e.name and t0 is None:
        t0 = e.time_range.start
      if t0 is not None:
        print("%9.1f us  +%8.1f us  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:70]))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
