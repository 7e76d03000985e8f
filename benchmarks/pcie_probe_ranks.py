#!/usr/bin/env python
"""
PCIe / host-memory floor of the sharded end-to-end path: every rank of one box moves what a sharded
renumber(in_place=True) of its 1024^3 uint64 slab must move -- 8 GiB up, then 8 + 4 GiB down, pinned host memory --
all ranks at the same time, with and without NUMA binding of the host buffers.

  torchrun --nproc-per-node N benchmarks/pcie_probe_ranks.py            (N = 1 works too)

Prints one JSON line per probe (rank 0): the slowest rank's time and the aggregate rate.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
  from fastremap_b200 import device as D
  rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  dev = torch.device("cuda", local)
  up, down = 8 << 30, 12 << 30
  dev_buf = torch.empty(down, dtype=torch.uint8, device=dev)

  def sync():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def measure(name, fn, moved, note):
    best = None
    for _ in range(3):
      sync()
      t0 = time.perf_counter()
      fn()
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
      best = dt if best is None else min(best, dt)
    if rank == 0:
      print(json.dumps({"probe": name, "ranks": world, "seconds_max_over_ranks": best, "GB/s_per_rank": moved / best / 1e9,
                        "GB/s_all_ranks": world * moved / best / 1e9, "note": note}), flush=True)
    return best

  for bind in (False, True):
    numa = D.bind_host_to_gpu(local) if bind else None
    host_up = torch.empty(up, dtype=torch.uint8, pin_memory=True)
    host_down = torch.empty(down, dtype=torch.uint8, pin_memory=True)
    host_up.fill_(1)
    host_down.fill_(2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    note = "host buffers bound to the GPU's NUMA node: %s" % (numa,) if bind else "host buffers wherever the allocator put them"

    def h2d():
      dev_buf[:up].copy_(host_up, non_blocking=True)

    def d2h():
      host_down.copy_(dev_buf, non_blocking=True)

    def seq():
      dev_buf[:up].copy_(host_up, non_blocking=True)
      host_down.copy_(dev_buf, non_blocking=True)

    def duplex():
      with torch.cuda.stream(s1):
        dev_buf[:up].copy_(host_up, non_blocking=True)
      with torch.cuda.stream(s2):
        host_down[:up].copy_(dev_buf[:up], non_blocking=True)

    t_up = measure("h2d 8 GiB", h2d, up, note)
    t_dn = measure("d2h 12 GiB", d2h, down, note)
    t_seq = measure("h2d 8 GiB then d2h 12 GiB (what a sharded renumber must do: no id is final before the merge)", seq, up + down, note)
    measure("h2d 8 GiB and d2h 8 GiB at once", duplex, 2 * up, note)
    if rank == 0:
      print(json.dumps({"probe": "sharded renumber e2e floor", "ranks": world, "Gvox/s_all_ranks": world * (1 << 30) / t_seq / 1e9,
                        "bound": bind}), flush=True)
    del host_up, host_down
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
