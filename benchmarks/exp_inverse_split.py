#!/usr/bin/env python
"""Stage split of unique(return_inverse=True) on a device-resident 1024^3 volume."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
S = 1024
n = S ** 3
D.dev()


def ev():
  e = torch.cuda.Event(enable_timing=True)
  e.record()
  return e


for dtype in (np.uint32, np.uint64):
  w = np.dtype(dtype).itemsize
  vol = synth_device((S, S, S), (46, 46, 46), dtype, seed=3, zero_frac=0.1).reshape(-1).view(torch.uint8)
  for rep in range(3):
    torch.cuda.synchronize()
    e0 = ev()
    uniq, cts, nu = api.unique_counts_device(vol, n, dtype, row_bytes=S * w)
    e1 = ev()
    t = api.table_from_arrays(uniq, None, nu, dtype, op=_lib.OP_SET, positions=True, resolve=False)
    e2 = ev()
    out = D.alloc(n * w)
    api._relabel(vol, n, np.dtype(dtype), t, _lib.MISS_ERROR, dst_wide=out, batch_rows=api._batch_rows(nu))
    e3 = ev()
    api._relabel(vol, n, np.dtype(dtype), t, _lib.MISS_ERROR, dst_wide=out, batch_rows=0)
    e4 = ev()
    wide = api._cast(out, np.dtype("u%d" % w), np.uint64, n) if w < 8 else out
    e5 = ev()
    torch.cuda.synchronize()
  print(json.dumps({"dtype": np.dtype(dtype).name, "nu": int(nu), "counts_ms": round(e0.elapsed_time(e1), 3), "table_ms": round(e1.elapsed_time(e2), 3),
                    "relabel_batched_ms": round(e2.elapsed_time(e3), 3), "relabel_plain_ms": round(e3.elapsed_time(e4), 3),
                    "cast_ms": round(e4.elapsed_time(e5), 3), "batch_rows": api._batch_rows(nu)}), flush=True)
  del vol, out, wide
