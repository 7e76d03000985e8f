#!/usr/bin/env python
"""Plain vs batched K3 lookup (label -> position table, out of place) at growing label counts, 1024^3."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
S = 1024
n = S ** 3
D.dev()
POLICY = _lib.MISS_ERROR if 'error' in sys.argv[1:] else _lib.MISS_KEEP


def ev():
  e = torch.cuda.Event(enable_timing=True)
  e.record()
  return e


for dtype in (np.uint64,) if 'u64' in sys.argv[1:] else (np.uint64, np.uint32):
  w = np.dtype(dtype).itemsize
  for g in (32, 46, 66, 84, 108, 165):
    vol = synth_device((S, S, S), (g, g, g), dtype, seed=3, zero_frac=0.1).reshape(-1).view(torch.uint8)
    uniq, cts, nu = api.unique_counts_device(vol, n, dtype, row_bytes=S * w)
    t = api.table_from_arrays(uniq, None, nu, dtype, op=_lib.OP_SET, positions=True, resolve=False)
    out = D.alloc(n * w)
    res = {}
    for name, rows in (("plain", 0), ("batched2", 2)):
      best = None
      for _ in range(3):
        torch.cuda.synchronize()
        a = ev()
        api._relabel(vol, n, np.dtype(dtype), t, POLICY, dst_wide=out, batch_rows=rows)
        b = ev()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
      res[name] = round(best, 3)
    print(json.dumps({"dtype": np.dtype(dtype).name, "labels": int(nu), "table_MiB": round(t.cap * 16 / 2**20, 1), **res}), flush=True)
    del vol, out, t, uniq, cts
