#!/usr/bin/env python
"""renumber end to end from an ORDINARY (pageable) numpy array against a page-locked one."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import fastremap_b200 as fr
from fastremap_b200.synth import synth_device

shape = (512, 1024, 1024)
n = int(np.prod(shape))
vol = synth_device(shape, (23, 46, 46), np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
pinned = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True)
pinned.copy_(vol)
src = pinned.numpy().view(np.uint64).reshape(shape)
pageable = np.empty(shape, dtype=np.uint64)
pinned_work = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True).numpy().view(np.uint64).reshape(shape)
del vol
for name, work in (("pinned", pinned_work), ("pageable", pageable)):
  for in_place in (True, False):
    best = None
    for i in range(3):
      work[...] = src
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      out, mapping = fr.renumber(work, in_place=in_place)
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      if i:
        best = dt if best is None else min(best, dt)
      del out
    print(json.dumps({"host_memory": name, "in_place": in_place, "seconds": best, "Gvox/s": n / best / 1e9}), flush=True)
