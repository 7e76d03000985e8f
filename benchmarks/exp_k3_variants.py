#!/usr/bin/env python
"""K3 on 64-bit labels with big tables: k_relabel_batched (0) vs k_relabel_runs (1) vs the sampler's choice (-1), across run
lengths (1024^3 uint64, cells^3 cells => runs of 1024 / cells voxels), remap through the first min(10^6, all) labels and
mask_except keeping every second label."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))
from fastremap_b200 import _lib, api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
from bench_ops import timed, digest, flat_u8, peak

S = 1024
n = S ** 3
L = _lib.load()
D.dev()
cells = [int(c) for c in sys.argv[1:]] or [46, 72, 108, 150, 215]
for g in cells:
  vol = flat_u8(synth_device((S, S, S), (g, g, g), np.uint64, seed=2, zero_frac=0.1))
  uniq, _, nu = api.unique_counts_device(vol, n, np.uint64)
  work = torch.empty_like(vol)
  for what in ("remap", "mask_except"):
    if what == "remap":
      Lk = min(1_000_000, nu)
      keys = uniq[: Lk * 8]
      vals = flat_u8(torch.arange(1, Lk + 1, dtype=torch.int64, device=vol.device))
      table = api.table_from_arrays(keys, vals, Lk, np.uint64)
      policy, kw = _lib.MISS_KEEP, {}
    else:
      keep = uniq[: nu * 8].view(torch.int64)[::2].contiguous()
      Lk = keep.numel()
      kb = flat_u8(keep)
      table = api.table_from_arrays(kb, kb, Lk, np.uint64)
      policy, kw = _lib.MISS_FILL, {"fill_bits": 0}
    row = {"cells": g, "run_voxels": round(S / g, 2), "labels_in_volume": nu, "op": what, "table_labels": Lk, "table_MiB": table.cap * 16 >> 20}
    for variant in (0, 1, -1):
      L.frb_set_relabel_variant(variant)
      m, mn, _ = timed(lambda: api._relabel(work, n, np.uint64, table, policy, dst_wide=work, batch_rows=api._batch_rows(Lk), **kw), 4,
                       setup=lambda: work.copy_(vol))
      name = {0: "batched", 1: "runs", -1: "auto"}[variant]
      row[name + "_ms"] = round(m, 3)
      row[name + "_digest"] = digest(work)
    L.frb_set_relabel_variant(-1)
    row["auto_frac_of_peak"] = round(16 * n / (row["auto_ms"] / 1e3) / 1e9 / peak(), 3)
    assert row["batched_digest"] == row["runs_digest"] == row["auto_digest"], row
    for k in ("batched_digest", "runs_digest", "auto_digest"):
      del row[k]
    print(json.dumps(row), flush=True)
    del table
  del vol, work, uniq
