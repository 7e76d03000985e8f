#!/usr/bin/env python
"""
torchrun --nproc-per-node N benchmarks/check_sharded.py
Parity of the multi-GPU renumber (NCCL key-table merge) against the oracle's renumber of the whole
volume: every rank renumbers its z-slab with fastremap_b200.sharded.renumber_sharded; rank 0
gathers the slabs and compares ids + mapping bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  from fastremap_b200 import api, sharded
  from fastremap_b200 import device as D
  from fastremap_b200.synth import synth_device, synth_numpy
  ok = True
  for dtype, shape, grid, pz, start in ((np.uint64, (96, 128, 160), (14, 9, 11), True, 1),
                                        (np.uint32, (64, 100, 72), (9, 10, 8), False, 5),
                                        (np.uint16, (48, 64, 64), (6, 7, 8), True, 1)):
    nz_total = shape[0]
    bounds = [nz_total * r // world for r in range(world + 1)]
    z0, z1 = bounds[rank], bounds[rank + 1]
    slab = synth_device((z1 - z0, shape[1], shape[2]), grid, dtype, seed=7, zero_frac=0.15, noise=0.05, z0=z0, nz_total=nz_total)
    buf = slab.reshape(-1).view(torch.uint8).clone()
    n_local = (z1 - z0) * shape[1] * shape[2]
    res = sharded.renumber_sharded(buf, n_local, dtype, start=start, preserve_zero=pz, write_wide=True)
    out = D.download(res.out, n_local, res.out_dtype)
    mapping = api._mapping_to_dict(res, dtype, pz)
    gathered = [None] * world
    dist.all_gather_object(gathered, (out, str(res.out_dtype)))
    if rank == 0:
      import oracle as orc
      orc.build()
      full = synth_numpy(shape, grid, dtype, seed=7, zero_frac=0.15, noise=0.05)
      e_out, e_map = orc.renumber(np.copy(full), start=start, preserve_zero=pz)
      got = np.concatenate([g[0] for g in gathered]).reshape(shape)
      same = got.dtype == e_out.dtype and np.array_equal(got, e_out) and mapping == e_map
      print("sharded renumber %s world=%d labels=%d -> %s" % (np.dtype(dtype).name, world, len(e_map), "OK" if same else "MISMATCH"), flush=True)
      ok = ok and same
  flag = torch.tensor([1 if ok else 0], device="cuda")
  dist.broadcast(flag, 0)
  dist.destroy_process_group()
  sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
  main()
