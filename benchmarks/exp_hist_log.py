#!/usr/bin/env python
"""Distribution of the per-warp eviction-log lengths of k_hist on the unique configs (debugging aid)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api  # noqa: E402
from fastremap_b200 import device as D  # noqa: E402
from fastremap_b200.synth import synth_device  # noqa: E402

S = 1024
n = S ** 3
D.dev()
L = _lib.load()
vol = synth_device((S, S, S), (100,) * 3, np.uint32, seed=1, zero_frac=0.1).reshape(-1).view(torch.uint8)
for plane in (0, S * S * 4):
  table = api.LabelTable(np.uint32, api._default_entries(n, None, 512), 0)
  ws_bytes = int(L.frb_hist_ws_bytes(n))
  ws = D.alloc(ws_bytes)
  _lib.check(L.frb_hist_hash(D.ptr(vol), n, D.code(np.uint32), D.ptr(table.slots), table.cap, D.ptr(table.meta), S * 4, plane,
                             D.ptr(ws), ws_bytes, D.stream()))
  torch.cuda.synchronize()
  off = int(L.frb_hist_ws_flags_offset())
  counts = ws[:off].view(torch.int32).cpu().numpy()
  live = counts[counts > 0]
  flag = int(ws[off:off + 4].view(torch.int32).item())
  # how many records carry label 0
  print(json.dumps({"plane_bytes": plane, "regions": int(counts.size), "live": int(live.size), "records": int(live.sum()),
                    "min": int(live.min()), "median": float(np.median(live)), "max": int(live.max()), "direct_flag": flag,
                    "pct": [int(x) for x in np.percentile(live, [1, 10, 25, 50, 75, 90, 99])],
                    "labels": int(table.read_meta()[_lib.META_COUNT])}), flush=True)

# which labels dominate the records of the plane walk (same-address atomics serialise when the logs are folded)
cap = (ws_bytes - off - 256) // 16 // counts.size
rec = ws[off + 256: off + 256 + counts.size * cap * 16].view(torch.int64).reshape(counts.size, cap, 2)
valid = torch.arange(cap, device=rec.device)[None, :] < torch.from_numpy(counts.astype(np.int64)).to(rec.device)[:, None]
keys = rec[:, :, 0][valid]
u, c = torch.unique(keys, return_counts=True)
top = torch.topk(c, 8)
print(json.dumps({"records": int(keys.numel()), "distinct": int(u.numel()), "top_keys": u[top.indices].tolist(), "top_counts": top.values.tolist()}))
