#!/usr/bin/env python
"""Device-resident unique(return_counts=True) of 1024^3 volumes at every unsigned width."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
from bench_ops import peak
S = 1024
n = S ** 3
D.dev()
only = set(sys.argv[1:])
for dtype, grid in ((np.uint8, 6), (np.uint16, 36), (np.uint32, 46), (np.uint64, 46)):
  if only and np.dtype(dtype).name not in only:
    continue
  w = np.dtype(dtype).itemsize
  vol = synth_device((S, S, S), (grid,) * 3, dtype, seed=3, zero_frac=0.1).reshape(-1).view(torch.uint8)
  best = None
  for _ in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    uniq, cts, nu = api.unique_counts_device(vol, n, dtype, row_bytes=S * w, plane_bytes=S * S * w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
  print(json.dumps({"dtype": np.dtype(dtype).name, "n_unique": int(nu), "ms": round(best, 3), "alg_GB/s": round(n * w / best / 1e6, 1),
                    "frac_of_measured_hbm_peak": round(n * w / best / 1e6 / peak(), 3)}), flush=True)
  del vol
