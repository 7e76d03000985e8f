#!/usr/bin/env python
"""Device-resident renumber of 1024^3 volumes at every unsigned width: ms per call and the stage split."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
S = 1024
n = S ** 3
D.dev()
only = set(sys.argv[1:])  # e.g. "uint32": that dtype only
for dtype, grid in ((np.uint8, 6), (np.uint16, 36), (np.uint32, 46), (np.uint64, 46)):
  if only and np.dtype(dtype).name not in only:
    continue
  vol = synth_device((S, S, S), (grid,) * 3, dtype, seed=3, zero_frac=0.1).reshape(-1).view(torch.uint8)
  work = torch.empty_like(vol)
  best = None
  for _ in range(4):
    work.copy_(vol)
    timers = {}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = api.renumber_device(work, n, dtype, write_wide=True, timers=timers, row_bytes=S * np.dtype(dtype).itemsize)
    e1.record()
    torch.cuda.synchronize()
    w = np.dtype(dtype).itemsize
    wo = np.dtype(res.out_dtype).itemsize
    alg = n * (w + w + w + (wo if wo < w else 0))
    row = {"dtype": np.dtype(dtype).name, "labels": res.n_labels, "out": str(res.out_dtype), "total_ms": round(e0.elapsed_time(e1), 3),
           "alg_GB/s": round(alg / e0.elapsed_time(e1) / 1e6, 1)}
    row.update({k: round(a.elapsed_time(b), 3) for k, (a, b) in timers.items()})
    best = row if best is None or row["total_ms"] < best["total_ms"] else best
  print(json.dumps(best), flush=True)
  del vol, work, res
