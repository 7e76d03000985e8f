#!/usr/bin/env python
"""PCIe calibration for the e2e number: pinned H2D / D2H / both at once, and the streamed renumber
at several chunk sizes.  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
  import fastremap_b200 as fr
  from fastremap_b200 import api
  from fastremap_b200.synth import synth_device
  nbytes = 8 << 30
  host_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
  host_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
  dev_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  dev_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

  def timed(fn, reps=3):
    best = None
    for _ in range(reps):
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      fn()
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      best = dt if best is None else min(best, dt)
    return best

  def h2d():
    dev_a.copy_(host_a, non_blocking=True)

  def d2h():
    host_b.copy_(dev_b, non_blocking=True)

  def both():
    with torch.cuda.stream(s1):
      dev_a.copy_(host_a, non_blocking=True)
    with torch.cuda.stream(s2):
      host_b.copy_(dev_b, non_blocking=True)

  def chunked(chunk):
    def f():
      with torch.cuda.stream(s1):
        for a in range(0, nbytes, chunk):
          dev_a[a:a + chunk].copy_(host_a[a:a + chunk], non_blocking=True)
      with torch.cuda.stream(s2):
        for a in range(0, nbytes, chunk):
          host_b[a:a + chunk].copy_(dev_b[a:a + chunk], non_blocking=True)
    return f

  for name, fn, moved in (("h2d", h2d, nbytes), ("d2h", d2h, nbytes), ("both", both, 2 * nbytes),
                          ("both_chunked_64MiB", chunked(64 << 20), 2 * nbytes), ("both_chunked_16MiB", chunked(16 << 20), 2 * nbytes)):
    t = timed(fn)
    print(json.dumps({"probe": name, "seconds": t, "GB/s": moved / t / 1e9}), flush=True)

  side = 1024
  vol = synth_device((side, side, side), (46, 46, 46), np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
  host_a.copy_(vol)
  del dev_a, dev_b
  torch.cuda.synchronize()
  work = host_b.numpy().view(np.uint64).reshape(side, side, side)
  for chunk_mib in (16, 32, 64, 128, 256):
    api.STREAM_CHUNK_BYTES = chunk_mib << 20
    best = None
    for i in range(3):
      host_b.copy_(host_a)
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      out, mapping = fr.renumber(work, in_place=True)
      dt = time.perf_counter() - t0
      if i:
        best = dt if best is None else min(best, dt)
      del out, mapping
    print(json.dumps({"probe": "renumber_e2e", "chunk_MiB": chunk_mib, "seconds": best, "Gvox/s": side ** 3 / best / 1e9}), flush=True)


if __name__ == "__main__":
  main()
