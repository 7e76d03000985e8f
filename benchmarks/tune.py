#!/usr/bin/env python
"""Sweep the tile-run knob (frb_set_tile_run) over the streaming kernels on the headline volumes."""
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


def timed(fn, reps=4, setup=None):
  ms = []
  for i in range(1 + reps):
    if setup:
      setup()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    if i >= 1:
      ms.append(e0.elapsed_time(e1))
  return float(np.mean(ms))


def main():
  S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  runs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 16, 64, 256, 1 << 20]
  n = S ** 3
  D.dev()
  L = _lib.load()
  vol = synth_device((S, S, S), (46 * S // 1024,) * 3, np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
  work = torch.empty_like(vol)
  res = api.renumber_device(vol.clone(), n, np.uint64)
  dense32 = res.out.clone()
  del res
  sparse32 = synth_device((S, S, S), (100 * S // 1024,) * 3, np.uint32, seed=1, zero_frac=0.1).reshape(-1).view(torch.uint8)
  v64 = vol.view(torch.int64)
  print(json.dumps({"torch_sum_int64_ms": timed(lambda: v64.sum()), "torch_max_int64_ms": timed(lambda: v64.max()),
                    "torch_copy_ms": timed(lambda: work.copy_(vol)),
                    "torch_eq_count_ms": timed(lambda: (v64 == 5).sum())}), flush=True)
  for run in runs:
    L.frb_set_tile_run(run)
    row = {"tile_run": run}
    probe_out = torch.zeros(2, dtype=torch.int64, device=vol.device)
    nar = torch.empty(n * 4, dtype=torch.uint8, device=vol.device)
    for mode in (0, 1, 2, 4, 6):
      row["probe%d_ms" % mode] = timed(lambda: _lib.check(L.frb_stream_probe(D.ptr(vol), n * 8, D.ptr(work), D.ptr(nar), mode, D.ptr(probe_out), D.stream())))
    del nar
    row["minmax_ms"] = timed(lambda: api._scan(vol, n, np.uint64))
    timers = {}

    def ren():
      api.renumber_device(work, n, np.uint64, write_wide=True, timers=timers)
    tot = timed(ren, setup=lambda: work.copy_(vol))
    row["renumber_ms"] = tot
    row["k1_ms"] = timers["k1"][0].elapsed_time(timers["k1"][1])
    row["k3_ms"] = timers["k3"][0].elapsed_time(timers["k3"][1])
    row["unique_dense32_hash_ms"] = timed(lambda: api.unique_counts_device(dense32, n, np.uint32, strategy="hash"))
    row["unique_sparse32_hash_ms"] = timed(lambda: api.unique_counts_device(sparse32, n, np.uint32, strategy="hash"))
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
  main()
