#!/usr/bin/env python
"""k_hist_apply_log alone, on the logs of one counting pass, into tables of different sizes (debugging aid)."""
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
c = D.code(np.uint32)
vol = synth_device((S, S, S), (100,) * 3, np.uint32, seed=1, zero_frac=0.1).reshape(-1).view(torch.uint8)
table = api.LabelTable(np.uint32, 1 << 22, 0)
ws_bytes = int(L.frb_hist_ws_bytes(n))
ws = D.alloc(ws_bytes)
_lib.check(L.frb_hist_hash(D.ptr(vol), n, c, D.ptr(table.slots), table.cap, D.ptr(table.meta), S * 4, S * S * 4, D.ptr(ws), ws_bytes, D.stream()))
torch.cuda.synchronize()
flush = torch.empty(1 << 28, dtype=torch.uint8, device=vol.device)
for entries in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
  ms = []
  for rep in range(4):
    t = api.LabelTable(np.uint32, entries, 0)
    flush.fill_(rep)  # push the fresh table out of L2
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.frb_hist_apply_log(D.ptr(ws), ws_bytes, n, c, D.ptr(t.slots), t.cap, D.ptr(t.meta), D.stream()))
    e1.record()
    torch.cuda.synchronize()
    if rep:
      ms.append(e0.elapsed_time(e1))
    meta = t.read_meta()
  print(json.dumps({"table_slots": t.cap, "table_MiB": t.cap * 16 >> 20, "labels": int(meta[_lib.META_COUNT]), "overflow": int(meta[_lib.META_OVERFLOW]),
                    "apply_ms": float(np.mean(ms))}), flush=True)
