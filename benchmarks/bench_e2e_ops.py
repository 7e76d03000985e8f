#!/usr/bin/env python
"""End-to-end (host array in, host array out) timing of the lookup family at 1024^3 uint64 with pinned
buffers: the chunk pipeline against the plain upload / kernel / download sequence."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import fastremap_b200 as fr
from fastremap_b200 import api
from fastremap_b200.synth import synth_device

S = 1024
n = S ** 3
vol = synth_device((S, S, S), (46, 46, 46), np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
host_src = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True)
host_work = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True)
host_src.copy_(vol)
labels = np.unique(host_src.numpy().view(np.uint64)[: 1 << 24])
keep = [int(x) for x in labels[::2]]
table = {int(k): i + 1 for i, k in enumerate(labels[::2])}
work = host_work.numpy().view(np.uint64).reshape(S, S, S)
stream_min = api.STREAM_MIN_BYTES
for name, fn in (("mask_except(in_place=True), %d kept labels" % len(keep), lambda: fr.mask_except(work, keep, in_place=True)),
                 ("remap(preserve_missing_labels=True, in_place=True), %d entries" % len(table), lambda: fr.remap(work, table, preserve_missing_labels=True, in_place=True))):
  for mode, smin in (("pipeline", stream_min), ("sequential", 1 << 62)):
    api.STREAM_MIN_BYTES = smin
    best = None
    for i in range(3):
      host_work.copy_(host_src)
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      fn()
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      if i:
        best = dt if best is None else min(best, dt)
    print(json.dumps({"op": name, "mode": mode, "seconds": best, "Gvox/s": n / best / 1e9}), flush=True)
