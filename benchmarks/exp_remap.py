#!/usr/bin/env python
"""Experiment: remap / mask_except K3 time as a function of the table's load factor."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device

def timed(fn, setup, reps=3):
  ms = []
  for i in range(2 + reps):
    setup(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    if i >= 2: ms.append(e0.elapsed_time(e1))
  return min(ms)

S = 1024; n = S ** 3
L = _lib.load()
for name, g, frac_keep, policy in (("remap", 108, None, _lib.MISS_KEEP), ("mask_except", 215, 2, _lib.MISS_FILL)):
  vol = synth_device((S, S, S), (g, g, g), np.uint64, seed=2 if name == "remap" else 4, zero_frac=0.1 if name == "remap" else 0.0).reshape(-1).view(torch.uint8)
  uniq, _, nu = api.unique_counts_device(vol, n, np.uint64)
  if name == "remap":
    nk = min(1_000_000, nu); keys = uniq[: nk * 8]
  else:
    k = uniq[: nu * 8].view(torch.int64)[::2].contiguous(); nk = k.numel(); keys = k.view(torch.uint8)
  vals = torch.arange(1, nk + 1, dtype=torch.int64, device=vol.device).view(torch.uint8)
  work = torch.empty_like(vol)
  for mult in (4,):
    t = api.LabelTable(np.uint64, nk * mult, 0, slots_per_entry=2)
    _lib.check(L.frb_table_insert(D.ptr(keys), 8, D.ptr(vals), 8, 0, nk, _lib.OP_SET, 0, D.ptr(t.slots), t.cap, D.ptr(t.meta), D.stream()))
    for batch in (0, 2, 4):
      ms = timed(lambda: api._relabel(work, n, np.uint64, t, policy, dst_wide=work, batch_rows=batch), lambda: work.copy_(vol))
      print(json.dumps({"op": name, "keys": nk, "labels": nu, "cap": t.cap, "load": nk / t.cap, "batch_rows": batch, "ms": ms}), flush=True)
  del vol, work, uniq
