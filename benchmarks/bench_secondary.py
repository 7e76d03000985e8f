#!/usr/bin/env python
"""
Device-resident timing of the secondary kernels of the hot path (SURVEY 8 rows a3, a12 fast path, a13, f3): refit's
cast, the one-entry replace, remap_from_array, indices, point_cloud's compaction, foreground / minmax, plus the
helpers swap_halves and pack_pairs.  One JSON line per kernel with the fraction of the measured HBM peak on
algorithmic bytes (every input byte read once, every output byte written once).

  python benchmarks/bench_secondary.py [--side 1024] [--reps 5]
"""
import argparse
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
from benchmarks.bench_ops import flat_u8, peak, timed  # noqa: E402


def report(name, n, bytes_total, mean_ms, min_ms, **extra):
  gbs = bytes_total / (mean_ms / 1e3) / 1e9
  line = {"op": name, "voxels": n, "ms": mean_ms, "ms_min": min_ms, "alg_bytes": bytes_total, "alg_GB/s": gbs,
          "frac_of_measured_hbm_peak": gbs / peak()}
  line.update(extra)
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--side", type=int, default=1024)
  ap.add_argument("--reps", type=int, default=5)
  args = ap.parse_args()
  S = args.side
  n = S ** 3
  g = max(46 * S // 1024, 1)
  D.dev()
  L = _lib.load()
  st = D.stream

  vol64 = flat_u8(synth_device((S, S, S), (g, g, g), np.uint64, seed=0, zero_frac=0.1))
  res = api.renumber_device(vol64.clone(), n, np.uint64)
  ids32 = res.out[: n * 4]            # dense uint32 ids 0..~88k
  nl = int(res.n_labels)
  del res

  # ---- refit's astype (fastremap.pyx:301)
  for src, sdt, ddt in ((vol64, np.uint64, np.uint32), (ids32, np.uint32, np.uint64), (ids32, np.uint32, np.uint16), (ids32, np.uint32, np.uint8)):
    ws, wd = np.dtype(sdt).itemsize, np.dtype(ddt).itemsize
    dst = D.alloc(n * wd)
    m, mn, _ = timed(lambda: _lib.check(L.frb_cast(D.ptr(src), D.code(sdt), D.ptr(dst), D.code(ddt), n, st())), args.reps)
    report("cast %s -> %s (refit astype)" % (np.dtype(sdt).name, np.dtype(ddt).name), n, n * (ws + wd), m, mn)
    del dst

  # ---- remap's one-entry fast path (fastremap.pyx:721-729), in place
  for buf, dt in ((vol64, np.uint64), (ids32, np.uint32)):
    w = np.dtype(dt).itemsize
    work = buf.clone()
    m, mn, _ = timed(lambda: _lib.check(L.frb_replace_one(D.ptr(work), D.ptr(work), n, D.code(dt), 5, 7, st())), args.reps)
    report("replace_one %s in place" % np.dtype(dt).name, n, 2 * n * w, m, mn)
    del work

  # ---- remap_from_array (fastremap.pyx:771-791), in place, table = a permutation of the ids
  for dt in (np.uint32, np.uint64):
    w = np.dtype(dt).itemsize
    src = ids32 if dt == np.uint32 else api._cast(ids32, np.uint32, np.uint64, n)
    vals = torch.randperm(nl + 1, device=src.device).to(torch.int64)
    vals = flat_u8(vals.to(torch.int32) if dt == np.uint32 else vals)
    work = src.clone()
    m, mn, _ = timed(lambda: _lib.check(L.frb_remap_from_array(D.ptr(work), D.ptr(work), n, D.ptr(vals), nl + 1, D.code(dt), st())), args.reps,
                     setup=lambda: work.copy_(src[: work.numel()]))
    report("remap_from_array %s in place, %d-entry value array" % (np.dtype(dt).name, nl + 1), n, 2 * n * w, m, mn)
    del work

  # ---- indices(arr, value) (fastremap.pyx:104-119): ordered positions of one label
  meta = D.alloc_meta()
  ws_bytes = int(L.frb_indices_ws_bytes(n))
  ws = D.alloc(ws_bytes)
  out = D.alloc(n // 4 * 8)

  def indices_u32():
    _lib.check(L.frb_indices_count(D.ptr(ids32), n, D.code(np.uint32), 0, D.ptr(meta), D.ptr(ws), ws_bytes, st()))
    cnt = int(D.read_u64(meta)[_lib.META_LIST_COUNT])
    _lib.check(L.frb_indices_emit(D.ptr(ids32), n, D.code(np.uint32), 0, D.ptr(out), cnt, D.ptr(ws), st()))
    return cnt
  m, mn, cnt = timed(indices_u32, args.reps)
  report("indices uint32 (count + emit, %d hits)" % cnt, n, 2 * n * 4 + cnt * 8, m, mn)

  # ---- point_cloud's ordered compaction of the foreground (fastremap.pyx:1444-1540)
  ws_bytes2 = int(L.frb_nonzero_ws_bytes(n))
  ws2 = D.alloc(ws_bytes2)

  def nonzero_u32():
    _lib.check(L.frb_nonzero_count(D.ptr(ids32), n, D.code(np.uint32), D.ptr(meta), D.ptr(ws2), ws_bytes2, st()))
    return int(D.read_u64(meta)[_lib.META_LIST_COUNT])
  m, mn, nfg = timed(nonzero_u32, args.reps)
  report("point_cloud foreground count uint32 (%d voxels)" % nfg, n, n * 4, m, mn)
  del out
  lab, idx = D.alloc(nfg * 8), D.alloc(nfg * 8)
  m, mn, _ = timed(lambda: _lib.check(L.frb_nonzero_emit(D.ptr(ids32), n, D.code(np.uint32), D.ptr(lab), D.ptr(idx), nfg, D.ptr(ws2), st())), args.reps)
  report("point_cloud foreground emit uint32 -> (label, index) uint64 pairs", n, n * 4 + nfg * 16, m, mn)
  del lab, idx

  # ---- foreground / minmax / pixel_pairs (one fused pass, K6)
  for buf, dt in ((vol64, np.uint64), (ids32, np.uint32)):
    m, mn, _ = timed(lambda: api._scan(buf, n, dt), args.reps)
    report("minmax + pixel_pairs + foreground %s (K6)" % np.dtype(dt).name, n, n * np.dtype(dt).itemsize, m, mn)

  # ---- helpers: edge-list column swap, (parent, component) packing
  dst = D.alloc(n * 4)
  m, mn, _ = timed(lambda: _lib.check(L.frb_swap_halves(D.ptr(ids32), D.ptr(dst), n // 2, 8, st())), args.reps)
  report("swap_halves of (N, 2) uint32 edge rows", n // 2, 2 * n * 4, m, mn)
  del dst
  packed = D.alloc(n * 8)
  m, mn, _ = timed(lambda: _lib.check(L.frb_pack_pairs(D.ptr(ids32), 4, D.ptr(ids32), 4, D.ptr(packed), n, st())), args.reps)
  report("pack_pairs (uint32, uint32) -> uint64", n, n * 16, m, mn)


if __name__ == "__main__":
  main()
