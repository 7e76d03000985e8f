#!/usr/bin/env python
"""
Device-resident timing of the other BASELINE.json configs (unique / remap / mask_except /
component_map) on one B200.  Prints one JSON line per op: ms, Gvox/s, algorithmic GB/s and the
fraction of the measured HBM peak.  Inputs are generated on the device; the timed region is the
kernel sequence of the op (table builds timed separately), CUDA events, 2 warm-ups + `reps` runs.

  python benchmarks/bench_ops.py [--side 1024] [--reps 5] [--ops unique_dense,unique_sparse,remap,mask_except,component_map]
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


def peak():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, reps, setup=None):
  ms = []
  for i in range(2 + reps):
    if setup:
      setup()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
      ms.append(e0.elapsed_time(e1))
  return float(np.mean(ms)), float(np.min(ms)), out


def report(name, n, alg_bytes_per_voxel, mean_ms, min_ms, **extra):
  gbs = alg_bytes_per_voxel * n / (mean_ms / 1e3) / 1e9
  line = {"op": name, "voxels": n, "ms": mean_ms, "ms_min": min_ms, "Gvox/s": n / (mean_ms / 1e3) / 1e9,
          "alg_bytes_per_voxel": alg_bytes_per_voxel, "alg_GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak()}
  line.update(extra)
  print(json.dumps(line), flush=True)


def flat_u8(t):
  return t.reshape(-1).view(torch.uint8)


def digest(buf, nbytes=None):
  """order-independent 64-bit digest of a device buffer of 8-byte words (A/B runs of kernel variants must agree)"""
  v = (buf if nbytes is None else buf[:nbytes]).view(torch.int64)
  return "%016x" % ((int(v.sum().item()) ^ (int((v * 0x1E3779B97F4A7C15 - 0x61C8864680B583EB >> 7).sum().item()) << 1)) & 0xFFFFFFFFFFFFFFFF)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--side", type=int, default=1024)
  ap.add_argument("--reps", type=int, default=5)
  ap.add_argument("--ops", default="unique_dense,unique_sparse,remap,mask_except,component_map,minmax")
  ap.add_argument("--no-plane", action="store_true", help="withhold the plane_bytes hint: linear walk everywhere (A/B runs)")
  ap.add_argument("--tile-run", type=int, default=0, help="frb_set_tile_run: consecutive tiles per CTA run of the linear walk (A/B runs)")
  args = ap.parse_args()
  S, ops = args.side, set(args.ops.split(","))
  PL = 0 if args.no_plane else 1
  n = S ** 3
  sc = S / 1024.0
  D.dev()
  if args.tile_run:
    _lib.load().frb_set_tile_run(args.tile_run)

  if "unique_dense" in ops or "minmax" in ops:
    vol = synth_device((S, S, S), (max(int(46 * sc), 1),) * 3, np.uint64, seed=0, zero_frac=0.1)
    res = api.renumber_device(flat_u8(vol).clone(), n, np.uint64)
    dense = res.out  # uint32 ids 0..~88k
    assert res.out_dtype == np.dtype("uint32"), res.out_dtype
    del res
    if "minmax" in ops:
      m, mn, _ = timed(lambda: api._scan(flat_u8(vol), n, np.uint64), args.reps)
      report("minmax+pixel_pairs u64 (K6)", n, 8, m, mn)
    del vol
    if "unique_dense" in ops:
      m, mn, out = timed(lambda: api.unique_counts_device(dense, n, np.uint32, row_bytes=S * 4, plane_bytes=PL * S * S * 4), args.reps)
      report("unique(return_counts) u32 dense ids, incl. minmax pre-pass (config 2a)", n, 4, m, mn, n_unique=out[2])
      m, mn, out = timed(lambda: api.unique_counts_device(dense, n, np.uint32, strategy="hash", row_bytes=S * 4), args.reps)
      report("unique(return_counts) u32 dense ids, forced hash strategy", n, 4, m, mn, n_unique=out[2])
    del dense

  if "unique_sparse" in ops:
    vol = flat_u8(synth_device((S, S, S), (max(int(100 * sc), 1),) * 3, np.uint32, seed=1, zero_frac=0.1))
    m, mn, out = timed(lambda: api.unique_counts_device(vol, n, np.uint32, row_bytes=S * 4, plane_bytes=PL * S * S * 4), args.reps)
    report("unique(return_counts) u32 sparse ids, incl. minmax pre-pass (config 2b)", n, 4, m, mn, n_unique=out[2])
    m, mn, out = timed(lambda: api.unique_counts_device(vol, n, np.uint32, row_bytes=S * 4), args.reps)
    report("unique(return_counts) u32 sparse ids, linear walk (no plane hint)", n, 4, m, mn, n_unique=out[2])
    if S <= 512:
      m, mn, out = timed(lambda: api.unique_counts_device(vol, n, np.uint32, strategy="sort"), max(args.reps // 2, 1))
      report("unique(return_counts) u32 sparse ids, forced sort strategy", n, 4, m, mn, n_unique=out[2])
    del vol

  if "remap" in ops:
    vol = flat_u8(synth_device((S, S, S), (max(int(108 * sc), 1),) * 3, np.uint64, seed=2, zero_frac=0.1))
    uniq, _, nu = api.unique_counts_device(vol, n, np.uint64)
    L = min(1_000_000, nu)
    keys = uniq[: L * 8]
    vals = flat_u8(torch.arange(1, L + 1, dtype=torch.int64, device=vol.device))
    m, mn, table = timed(lambda: api.table_from_arrays(keys, vals, L, np.uint64), args.reps)
    print(json.dumps({"op": "remap table build from %d-entry key/value arrays" % L, "ms": m, "ms_min": mn}), flush=True)
    work = torch.empty_like(vol)
    m, mn, _ = timed(lambda: api._relabel(work, n, np.uint64, table, _lib.MISS_KEEP, dst_wide=work, batch_rows=api._batch_rows(L), plane_bytes=PL * S * S * 8), args.reps,
                     setup=lambda: work.copy_(vol))
    report("remap u64 in place, %d-entry table, preserve_missing_labels (config 3)" % L, n, 16, m, mn, labels_in_volume=nu, digest=digest(work))
    del vol, work, table

  if "mask_except" in ops or "component_map" in ops:
    g = max(int(215 * sc), 1)
    cc = flat_u8(synth_device((S, S, S), (g, g, g), np.uint64, seed=4, zero_frac=0.0))
    if "mask_except" in ops:
      uniq, _, nu = api.unique_counts_device(cc, n, np.uint64)
      keep = uniq[: nu * 8].view(torch.int64)[::2].contiguous()
      kb = flat_u8(keep)
      m, mn, table = timed(lambda: api.table_from_arrays(kb, kb, keep.numel(), np.uint64), args.reps)
      print(json.dumps({"op": "mask_except table build from %d labels" % keep.numel(), "ms": m, "ms_min": mn}), flush=True)
      work = torch.empty_like(cc)
      m, mn, _ = timed(lambda: api._relabel(work, n, np.uint64, table, _lib.MISS_FILL, fill_bits=0, dst_wide=work, batch_rows=api._batch_rows(keep.numel()), plane_bytes=PL * S * S * 8), args.reps,
                       setup=lambda: work.copy_(cc))
      report("mask_except u64 in place, %d kept of %d labels (config 5)" % (keep.numel(), nu), n, 16, m, mn, digest=digest(work))
      del work, table, uniq, keep, kb
    if "component_map" in ops:
      par = flat_u8(synth_device((S, S, S), (g, g, g), np.uint64, seed=5, zero_frac=0.0, cell_div=4))
      m, mn, out = timed(lambda: api.component_map_device(cc, n, np.uint64, par, np.uint64, max_labels=g ** 3, row_bytes=S * 8, plane_bytes=PL * S * S * 8), args.reps)
      report("component_map (u64, u64) device arrays (config 5)", n, 16, m, mn, n_components=out[2],
             digest=digest(out[0], out[2] * 8) + digest(out[1], out[2] * 8))
      del par
    del cc


if __name__ == "__main__":
  main()
