"""
BASELINE.json configs 1, 2, 3 and 5 at their stated sizes on one B200 (config 4 is the 8-GPU run of bench.py).

Per config: device-resident timing with CUDA events (2 warm-ups, `reps` timed runs, inputs >> L2 and restored or
re-read between runs), the roofline fraction on SURVEY 8(d)'s algorithmic bytes per voxel, the compiled reference
(`oracle/_ref`, or the oracle port when it is absent) timed on the box's host beside it on the stated volume / slab,
and bit-exact parity of the GPU result against that very reference run.  Where the generator gives a closed form
(component_map, mask_except, remap) the full-size GPU result is checked against it as well.

Used by bench.py (the `configs` list of the N=1 line) and runnable alone:
  python benchmarks/configs.py [--only 1,2a,2b,3,5a,5b] [--reps 5] [--slab-planes 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import fastremap_b200 as fr  # noqa: E402
from fastremap_b200 import _lib, api  # noqa: E402
from fastremap_b200 import device as D  # noqa: E402
from fastremap_b200.synth import synth_cell_labels, synth_device  # noqa: E402

NOMINAL_HBM_GBS = 8000.0  # the "~8 TB/s" north_star quotes
BASELINE_CONFIGS = {
  "1": "fastremap.renumber(in_place=True) on a synthetic 512^3 uint64 connectomics-like segmentation (~100k labels)",
  "2": "fastremap.unique(return_counts=True) on 1024^3 uint32 synthetic segmentation, 1 B200",
  "3": "fastremap.remap with a 1M-entry dict, preserve_missing_labels=True, on 1024^3 uint64",
  "5": "mask_except / component_map on paired 1024^3 uint64 volumes with ~10M distinct labels (hash-table stress)",
}


def flat_u8(t):
  return t.reshape(-1).view(torch.uint8)


def timed(fn, reps, setup=None, warm=3):
  """(median, min, last result) of `reps` timed runs after `warm` untimed ones, CUDA events.  The median, because a single run that
  meets a cudaMalloc of the caching allocator (first use of a size) is 2x the others and would be a fifth of a 5-run mean."""
  ms, out = [], None
  for i in range(warm + reps):
    if setup:
      setup()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    if i >= warm:
      ms.append(e0.elapsed_time(e1))
  return float(np.median(ms)), float(np.min(ms)), out


def roofline(n, bytes_per_voxel, ms, peak, kernel):
  gbs = bytes_per_voxel * n / (ms / 1e3) / 1e9
  return {"bound": "hbm", "kernel": kernel, "alg_bytes_per_voxel": bytes_per_voxel, "achieved": gbs, "peak": peak, "unit": "GB/s",
          "frac": gbs / peak, "frac_of_nominal_8TBs": gbs / NOMINAL_HBM_GBS, "traffic": None}


def to_host(t):
  """CUDA tensor -> numpy array (plain pageable memory: the reference mutates / owns it)."""
  return t.cpu().numpy()


def cpu_record(kind, what, voxels, secs):
  return {"value": voxels / secs / 1e9, "unit": "Gvox/s", "cores": 1, "kind": kind, "sample": what, "seconds": secs,
          "host_cores_available": os.cpu_count()}


def record(cid, workload, n, ms, ms_min, roof, cpu, parity, how, **extra):
  line = {"id": cid, "config": BASELINE_CONFIGS[cid[0]], "workload": workload, "voxels": n, "ms": ms, "ms_stat": "median of the timed runs (3 warm-ups)", "ms_min": ms_min,
          "Gvox/s": n / (ms / 1e3) / 1e9, "roofline": roof, "cpu_baseline": cpu, "parity": bool(parity), "parity_how": how}
  line.update(extra)
  return line


# ---------------------------------------------------------------------------------------------- config 1
def config1(ref, kind, peak, reps, **_):
  S, g = 512, 46
  n = S ** 3
  vol = synth_device((S, S, S), (g, g, g), np.uint64, seed=0, zero_frac=0.1)
  pristine = flat_u8(vol)
  work = torch.empty_like(pristine)
  ms, mn, res = timed(lambda: api.renumber_device(work, n, np.uint64, write_wide=True, row_bytes=S * 8), reps,
                      setup=lambda: work.copy_(pristine))
  labels = int(res.n_labels)
  out_dtype = str(res.out_dtype)
  del res, work
  host = to_host(vol)
  del vol, pristine
  ref_in = np.copy(host)
  t0 = time.perf_counter()
  ref_out, ref_map = ref.renumber(ref_in, in_place=True)
  secs = time.perf_counter() - t0
  g_in = np.copy(host)
  g_out, g_map = fr.renumber(g_in, in_place=True)
  parity = (g_out.dtype == ref_out.dtype and np.array_equal(g_out, ref_out) and np.array_equal(g_in, ref_in) and g_map == ref_map)
  return record(
    "1", "renumber(in_place=True) of synth(512^3 uint64, 46^3 cells, seed 0, 10%% zeros): %d labels, refit to %s" % (labels, out_dtype),
    n, ms, mn, roofline(n, 20, ms, peak, "K1 k_renumber_build + K2 rank + K3 k_relabel_auto, whole op"),
    cpu_record(kind, "the whole 512^3 volume, renumber(in_place=True), input copy excluded, 1 thread", n, secs), parity,
    "fastremap_b200.renumber(numpy, in_place=True) vs the reference on the whole volume: narrowed array, ids left in the "
    "caller's buffer and the mapping dict, all bit-exact", labels=labels)


# ---------------------------------------------------------------------------------------------- config 2
def _unique_config(cid, vol_t, what, ref, kind, peak, reps):
  n = vol_t.numel()
  ms, mn, out = timed(lambda: fr.unique(vol_t, return_counts=True), reps)
  g_u, g_c = to_host(out[0]), to_host(out[1])
  del out
  host = to_host(vol_t)
  t0 = time.perf_counter()
  r_u, r_c = ref.unique(host, return_counts=True)
  secs = time.perf_counter() - t0
  parity = (g_u.dtype == r_u.dtype and np.array_equal(g_u, r_u) and np.array_equal(g_c.astype(np.uint64), np.asarray(r_c).astype(np.uint64))
            and int(g_c.sum()) == n)
  return record(
    cid, "unique(return_counts=True) of %s: %d distinct labels" % (what, g_u.size), n, ms, mn,
    roofline(n, 4, ms, peak, "k_hist<4, HashSink> + table export + cooperative sort, through fastremap_b200.unique(cuda tensor)"),
    cpu_record(kind, "the whole 1024^3 uint32 volume, unique(return_counts=True), 1 thread", n, secs), parity,
    "fastremap_b200.unique(cuda tensor, return_counts=True) vs the reference on the whole volume: labels and counts bit-exact",
    n_unique=int(g_u.size))


def config2a(ref, kind, peak, reps, **_):
  S, g = 1024, 46
  n = S ** 3
  vol = synth_device((S, S, S), (g, g, g), np.uint64, seed=0, zero_frac=0.1)
  res = api.renumber_device(flat_u8(vol), n, np.uint64)
  assert res.out_dtype == np.dtype("uint32"), res.out_dtype
  dense = res.out[: n * 4].view(torch.uint32).reshape(S, S, S)
  del vol
  return _unique_config("2a", dense, "the dense ids renumber gives synth(1024^3, 46^3 cells) as uint32", ref, kind, peak, reps)


def config2b(ref, kind, peak, reps, **_):
  S, g = 1024, 100
  vol = synth_device((S, S, S), (g, g, g), np.uint32, seed=1, zero_frac=0.1)
  return _unique_config("2b", vol, "synth(1024^3 uint32, 100^3 cells, seed 1): sparse ids uniform in [1, 2^32)", ref, kind, peak, reps)


# ---------------------------------------------------------------------------------------------- config 3
def _sorted_counts(buf, n, dtype):
  _, cts, nu = api.unique_counts_device(buf, n, dtype)
  return torch.sort(cts[: nu * 8].view(torch.int64)).values, nu


def config3(ref, kind, peak, reps, slab_planes=256, **_):
  S, g = 1024, 108
  n = S ** 3
  vol_t = synth_device((S, S, S), (g, g, g), np.uint64, seed=2, zero_frac=0.1)
  vol = flat_u8(vol_t)
  uniq, _, nu = api.unique_counts_device(vol, n, np.uint64)
  L = min(1_000_000, nu)
  keys_dev = uniq[: L * 8]
  vals_dev = flat_u8(torch.arange(1, L + 1, dtype=torch.int64, device=vol.device))
  tb_ms, _, table = timed(lambda: api.table_from_arrays(keys_dev, vals_dev, L, np.uint64), reps)
  work = torch.empty_like(vol)
  ms, mn, _ = timed(lambda: api._relabel(work, n, np.uint64, table, _lib.MISS_KEEP, dst_wide=work, batch_rows=api._batch_rows(L), plane_bytes=S * S * 8), reps,
                    setup=lambda: work.copy_(vol))
  del table
  # the public call on the whole volume (dict -> arrays -> upload -> table -> K3), in place on a device copy
  keys_np = D.download(keys_dev, L, np.uint64)
  tbl = dict(zip(keys_np.tolist(), range(1, L + 1)))
  work.copy_(vol)
  work_t = work.view(torch.uint64).reshape(S, S, S)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  out_t = fr.remap(work_t, tbl, preserve_missing_labels=True, in_place=True)
  torch.cuda.synchronize()
  api_ms = (time.perf_counter() - t0) * 1e3
  assert out_t.data_ptr() == work_t.data_ptr()
  # full size: the mapping is injective (no preserved 63-bit label lies in 1..L), so the multiset of counts survives
  c_before, nu_b = _sorted_counts(vol, n, np.uint64)
  c_after, nu_a = _sorted_counts(work, n, np.uint64)
  full_ok = bool(nu_a == nu_b and torch.equal(c_before, c_after))
  # slab: first planes of the GPU's full-volume result vs the reference on the same planes with the same dict
  g_slab = to_host(out_t[:slab_planes])
  host = to_host(vol_t[:slab_planes])
  t0 = time.perf_counter()
  r_slab = ref.remap(host, tbl, preserve_missing_labels=True, in_place=True)
  secs = time.perf_counter() - t0
  parity = g_slab.dtype == r_slab.dtype and np.array_equal(g_slab, r_slab) and full_ok
  nslab = slab_planes * S * S
  return record(
    "3", "remap(in place, preserve_missing_labels=True) of synth(1024^3 uint64, 108^3 cells, seed 2; %d labels) through the dict "
    "{first %d sorted labels -> 1..%d}" % (nu, L, L), n, ms, mn,
    roofline(n, 16, ms, peak, "k_relabel_batched<8,2> (K3, 16-byte slots, home-pair probe)"),
    cpu_record(kind, "the first %d z-planes (%d voxels) with the full %d-entry dict, remap(in_place=True), 1 thread" % (slab_planes, nslab, L),
               nslab, secs), parity,
    "first %d planes of fastremap_b200.remap(cuda tensor, dict, in_place=True) run on the WHOLE volume vs the reference on those planes, "
    "bit-exact; whole volume: label count and sorted count multiset unchanged by the (injective) mapping" % slab_planes,
    table_build_ms=tb_ms, api_ms_with_dict_conversion=api_ms, dict_entries=L, labels_in_volume=int(nu), full_size_check=full_ok)


# ---------------------------------------------------------------------------------------------- config 5
def _cc_volume(S, g):
  return synth_device((S, S, S), (g, g, g), np.uint64, seed=4, zero_frac=0.0)


def config5a(ref, kind, peak, reps, slab_planes=256, **_):
  S, g = 1024, 215
  n = S ** 3
  cc_t = _cc_volume(S, g)
  cc = flat_u8(cc_t)
  uniq, cts, nu = api.unique_counts_device(cc, n, np.uint64)
  keep = uniq[: nu * 8].view(torch.int64)[::2].contiguous()
  kept_voxels = int(cts[: nu * 8].view(torch.int64)[::2].sum().item())
  kb = flat_u8(keep)
  nk = keep.numel()
  tb_ms, _, table = timed(lambda: api.table_from_arrays(kb, kb, nk, np.uint64), reps)
  work = torch.empty_like(cc)
  ms, mn, _ = timed(lambda: api._relabel(work, n, np.uint64, table, _lib.MISS_FILL, fill_bits=0, dst_wide=work, batch_rows=api._batch_rows(nk), plane_bytes=S * S * 8),
                    reps, setup=lambda: work.copy_(cc))
  del table
  keep_list = D.download(kb, nk, np.uint64).tolist()
  work.copy_(cc)
  work_t = work.view(torch.uint64).reshape(S, S, S)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  out_t = fr.mask_except(work_t, keep_list, in_place=True)
  torch.cuda.synchronize()
  api_ms = (time.perf_counter() - t0) * 1e3
  assert out_t.data_ptr() == work_t.data_ptr()
  # full size: the voxels that survive are exactly the voxels of the kept labels (no label is 0 here)
  nonzero = api._scan(work, n, np.uint64)[3]
  full_ok = bool(nonzero == kept_voxels)
  g_slab = to_host(out_t[:slab_planes])
  host = to_host(cc_t[:slab_planes])
  t0 = time.perf_counter()
  r_slab = ref.mask_except(host, keep_list, in_place=True)
  secs = time.perf_counter() - t0
  parity = g_slab.dtype == r_slab.dtype and np.array_equal(g_slab, r_slab) and full_ok
  nslab = slab_planes * S * S
  return record(
    "5a", "mask_except(in place) of synth(1024^3 uint64, 215^3 cells, seed 4, no zeros; %d labels), keeping every second sorted label (%d)"
    % (nu, nk), n, ms, mn, roofline(n, 16, ms, peak, "k_relabel_batched<8,2> (K3, fill policy, table resident in HBM)"),
    cpu_record(kind, "the first %d z-planes (%d voxels) with the full %d-label keep list, mask_except(in_place=True), 1 thread"
               % (slab_planes, nslab, nk), nslab, secs), parity,
    "first %d planes of fastremap_b200.mask_except(cuda tensor, list, in_place=True) run on the WHOLE volume vs the reference on those "
    "planes, bit-exact; whole volume: surviving voxels == summed counts of the kept labels" % slab_planes,
    table_build_ms=tb_ms, api_ms_with_list_conversion=api_ms, kept_labels=nk, labels_in_volume=int(nu), full_size_check=full_ok)


def config5b(ref, kind, peak, reps, slab_planes=256, **_):
  S, g = 1024, 215
  n = S ** 3
  cc_t = _cc_volume(S, g)
  par_t = synth_device((S, S, S), (g, g, g), np.uint64, seed=5, zero_frac=0.0, cell_div=4)
  cc, par = flat_u8(cc_t), flat_u8(par_t)
  ms, mn, out = timed(lambda: api.component_map_device(cc, n, np.uint64, par, np.uint64, max_labels=g ** 3, row_bytes=S * 8, plane_bytes=S * S * 8), reps)
  keys, pars, nl = out
  # full size, closed form: parent is a function of the cell, so the map must be {label(cell): parent label(cell)}
  k_np, p_np = D.download(keys, nl, np.uint64), D.download(pars, nl, np.uint64)
  order = np.argsort(k_np, kind="stable")
  cells_c = synth_cell_labels((g, g, g), np.uint64, seed=4, zero_frac=0.0).reshape(-1)
  cells_p = synth_cell_labels((g, g, g), np.uint64, seed=5, zero_frac=0.0, cell_div=4).reshape(-1)
  eo = np.argsort(cells_c, kind="stable")
  e_k, e_p = cells_c[eo], cells_p[eo]
  if e_k.size > 1 and (e_k[1:] == e_k[:-1]).any():  # two cells with one label: keep the last cell in memory order (largest index)
    last = np.append(e_k[1:] != e_k[:-1], True)
    e_k, e_p = e_k[last], e_p[last]
  full_ok = bool(np.array_equal(k_np[order], e_k) and np.array_equal(p_np[order], e_p))
  del out, keys, pars
  # slab: the public call on the first planes vs the reference's dict
  g_map = fr.component_map(cc_t[:slab_planes], par_t[:slab_planes])
  hc, hp = to_host(cc_t[:slab_planes]), to_host(par_t[:slab_planes])
  t0 = time.perf_counter()
  r_map = ref.component_map(hc, hp)
  secs = time.perf_counter() - t0
  parity = g_map == r_map and full_ok
  nslab = slab_planes * S * S
  return record(
    "5b", "component_map of (synth cc: 1024^3 uint64, 215^3 cells, seed 4; parent: seed 5, 4 cells per parent): %d components" % nl,
    n, ms, mn, roofline(n, 16, ms, peak, "K7 k_component_build (atomicMax of run starts) + table scan / gather"),
    cpu_record(kind, "the first %d z-planes (%d voxels, %d components) of both volumes, component_map, 1 thread" % (slab_planes, nslab, len(r_map)),
               nslab, secs), parity,
    "fastremap_b200.component_map(cuda tensors) on the first %d planes vs the reference's dict, equal; whole volume: the device arrays equal "
    "the closed form {label(cell): parent(cell)} of the generator for all %d components" % (slab_planes, nl),
    n_components=int(nl), full_size_check=full_ok)


CONFIGS = [("1", config1), ("2a", config2a), ("2b", config2b), ("3", config3), ("5a", config5a), ("5b", config5b)]


def run_configs(ref, kind, peak, reps=5, only=None, slab_planes=256, log=None):
  out = []
  for cid, fn in CONFIGS:
    if only and cid not in only:
      continue
    t0 = time.perf_counter()
    try:
      rec = fn(ref, kind, peak, reps, slab_planes=slab_planes)
    except Exception as e:  # a failing config must show up in the line, not hide the others
      rec = {"id": cid, "config": BASELINE_CONFIGS[cid[0]], "error": repr(e), "parity": False}
    rec["wall_s"] = time.perf_counter() - t0
    torch.cuda.empty_cache()
    if log:
      log(rec)
    out.append(rec)
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--only", default="")
  ap.add_argument("--reps", type=int, default=5)
  ap.add_argument("--slab-planes", type=int, default=256)
  args = ap.parse_args()
  sys.path.insert(0, ROOT)
  import bench
  D.dev()
  ref, kind = bench.cpu_impl()
  peak, _ = bench.load_peaks()
  only = set(args.only.split(",")) if args.only else None
  run_configs(ref, kind, peak, args.reps, only, args.slab_planes, log=lambda r: print(json.dumps(r), flush=True))


if __name__ == "__main__":
  main()
