#!/usr/bin/env python
"""Driver for ncu captures of the non-headline configs: runs each selected op twice on its 1024^3 synthetic volume.

  python benchmarks/prof_ops.py [side] [ops]      ops: comma list of unique_sparse,mask_except,component_map,remap,unique_dense
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api  # noqa: E402
from fastremap_b200 import device as D  # noqa: E402
from fastremap_b200.synth import synth_device  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ops = set((sys.argv[2] if len(sys.argv) > 2 else "unique_sparse,mask_except,component_map").split(","))
n = S ** 3
sc = S / 1024.0
D.dev()


def flat_u8(t):
  return t.reshape(-1).view(torch.uint8)


if "unique_sparse" in ops:
  vol = flat_u8(synth_device((S, S, S), (max(int(100 * sc), 1),) * 3, np.uint32, seed=1, zero_frac=0.1))
  for _ in range(2):
    out = api.unique_counts_device(vol, n, np.uint32, row_bytes=S * 4, plane_bytes=S * S * 4)
  print("unique_sparse", out[2])
  del vol

if "unique_dense" in ops:
  vol = synth_device((S, S, S), (max(int(46 * sc), 1),) * 3, np.uint64, seed=0, zero_frac=0.1)
  res = api.renumber_device(flat_u8(vol).clone(), n, np.uint64)
  dense = res.out
  del vol
  for _ in range(2):
    out = api.unique_counts_device(dense, n, np.uint32, row_bytes=S * 4, plane_bytes=S * S * 4)
  print("unique_dense", out[2])
  del dense, res

if "remap" in ops:
  vol = flat_u8(synth_device((S, S, S), (max(int(108 * sc), 1),) * 3, np.uint64, seed=2, zero_frac=0.1))
  uniq, _, nu = api.unique_counts_device(vol, n, np.uint64)
  L = min(1_000_000, nu)
  keys = uniq[: L * 8]
  vals = flat_u8(torch.arange(1, L + 1, dtype=torch.int64, device=vol.device))
  table = api.table_from_arrays(keys, vals, L, np.uint64)
  work = torch.empty_like(vol)
  for _ in range(2):
    work.copy_(vol)
    api._relabel(work, n, np.uint64, table, _lib.MISS_KEEP, dst_wide=work, batch_rows=api._batch_rows(L), plane_bytes=S * S * 8)
  print("remap", L)
  del vol, work, table

if "mask_except" in ops or "component_map" in ops:
  g = max(int(215 * sc), 1)
  cc = flat_u8(synth_device((S, S, S), (g, g, g), np.uint64, seed=4, zero_frac=0.0))
  if "mask_except" in ops:
    uniq, _, nu = api.unique_counts_device(cc, n, np.uint64)
    keep = uniq[: nu * 8].view(torch.int64)[::2].contiguous()
    kb = flat_u8(keep)
    table = api.table_from_arrays(kb, kb, keep.numel(), np.uint64)
    work = torch.empty_like(cc)
    for _ in range(2):
      work.copy_(cc)
      api._relabel(work, n, np.uint64, table, _lib.MISS_FILL, fill_bits=0, dst_wide=work, batch_rows=api._batch_rows(keep.numel()), plane_bytes=S * S * 8)
    print("mask_except", keep.numel())
    del work, table, uniq, keep, kb
  if "component_map" in ops:
    par = flat_u8(synth_device((S, S, S), (g, g, g), np.uint64, seed=5, zero_frac=0.0, cell_div=4))
    for _ in range(2):
      out = api.component_map_device(cc, n, np.uint64, par, np.uint64, max_labels=g ** 3, row_bytes=S * 8, plane_bytes=S * S * 8)
    print("component_map", out[2])
torch.cuda.synchronize()
print("ok")
