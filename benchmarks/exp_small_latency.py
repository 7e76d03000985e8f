#!/usr/bin/env python
"""Wall-clock latency of the public API on small host arrays (the fixed costs: launches, host round trips, dict building)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import fastremap_b200 as fr
from fastremap_b200.synth import synth_numpy
import bench
ref, kind = bench.cpu_impl()   # the compiled reference (oracle/_ref) on this box's host, else the oracle port
for side in (32, 64, 128, 256):
  g = max(side // 22, 1)
  a = synth_numpy((side, side, side), (g, g, g), np.uint64, seed=1, zero_frac=0.1)
  table = {int(k): int(k) + 1 for k in np.unique(a)[:100]}
  ops = {"renumber": lambda x: fr.renumber(x, in_place=False), "unique_counts": lambda x: fr.unique(x, return_counts=True),
         "remap_100": lambda x: fr.remap(x, table, preserve_missing_labels=True), "mask_except_3": lambda x: fr.mask_except(x, [1, 2, 3]),
         "minmax": fr.minmax, "asfortranarray": lambda x: fr.asfortranarray(np.copy(x))}
  row = {"side": side, "MiB": round(a.nbytes / 2**20, 2)}
  for name, fn in ops.items():
    best = None
    for i in range(6):
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      fn(a)
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      if i >= 2:
        best = dt if best is None else min(best, dt)
    row[name + "_ms"] = round(best * 1e3, 3)
  ref_ops = {"renumber": lambda x: ref.renumber(x, in_place=False), "unique_counts": lambda x: ref.unique(x, return_counts=True),
             "remap_100": lambda x: ref.remap(x, table, preserve_missing_labels=True), "mask_except_3": lambda x: ref.mask_except(x, [1, 2, 3]),
             "minmax": ref.minmax}
  for name, fn in ref_ops.items():
    best = None
    for i in range(4):
      t0 = time.perf_counter()
      fn(a)
      dt = time.perf_counter() - t0
      if i >= 1:
        best = dt if best is None else min(best, dt)
    row[name + "_%s_cpu_ms" % kind] = round(best * 1e3, 3)
  print(json.dumps(row), flush=True)
