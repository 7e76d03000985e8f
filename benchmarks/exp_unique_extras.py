#!/usr/bin/env python
"""unique with return_index / return_inverse / return_counts on a device-resident 1024^3 volume (public API, CUDA tensor in)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import fastremap_b200 as fr
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
S = 1024
D.dev()
for dtype in (np.uint32, np.uint64):
  vol = synth_device((S, S, S), (46, 46, 46), dtype, seed=3, zero_frac=0.1)
  for kw in ({"return_counts": True}, {"return_index": True}, {"return_inverse": True},
             {"return_index": True, "return_inverse": True, "return_counts": True}):
    best = None
    for _ in range(3):
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      out = fr.unique(vol, **kw)
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      best = dt if best is None else min(best, dt)
      del out
    print(json.dumps({"dtype": np.dtype(dtype).name, "kwargs": sorted(kw), "ms": round(best * 1e3, 3)}), flush=True)
  del vol
