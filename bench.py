#!/usr/bin/env python
"""
bench.py -- headline benchmark: renumber of a 1024^3 uint64 synthetic segmentation (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N ...            the unmodified reference on host cores

One "step" = one renumber (in_place=True semantics: ids written back at the old width + narrowed
copy + label mapping) of one 1024^3 uint64 volume per GPU.  N>1 (torchrun, one rank per GPU) is
WEAK scaling by default: every rank owns one 1024^3 z-slab of an (N*1024, 1024, 1024) volume and the
per-shard key tables are merged with one NCCL all-gather (fastremap_b200/sharded.py); at N=8 that is
BASELINE config 4 (2048^3 voxels).  `--scaling strong` splits ONE 1024^3 volume N ways instead.
Both arms print the same `config` dict (workload_config).

value   : Gvox/s with the input resident in HBM, CUDA-event time per step summed over K steps
          (the un-timed restore of the input between steps rewrites 8 GiB, which also flushes L2)
e2e     : Gvox/s through the public API fastremap_b200.renumber(numpy array, in_place=True) with
          pinned host buffers: H2D of the volume, kernels, D2H of both result arrays, dict build
roofline: the dominant kernel (K3 k_relabel_auto: 8 B read + 8 B wide write + 4 B narrow write per
          voxel = 20 algorithmic bytes/voxel), timed with CUDA events on the launching stream
cpu_baseline: oracle/_ref (the compiled reference) renumber(in_place=True) on the WHOLE 1024^3 volume, 1 thread,
          and the GPU result must equal it bit for bit (array, caller's buffer, dict) or no number is printed.
configs : (N=1) BASELINE.json configs 1, 2, 3, 5 at size, each with ms, roofline, cpu_baseline and parity
          (benchmarks/configs.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIDE = 1024
GRID = 46                 # cells per 1024 voxels per axis  (~88 k labels per 1024^3, ~91 % equal neighbours)
DTYPE = np.uint64
SEED = 0
ZERO_FRAC = 0.1
LABELS_PER_SLAB = 87554   # distinct non-zero labels of one 1024^3 slab of this generator (asserted at run time)
ALG_BYTES_PER_VOXEL = 20  # 8 read + 8 in-place write + 4 narrow write
NOMINAL_HBM_GBS = 8000.0  # the "~8 TB/s" north_star quotes


def workload_config(world, scaling):
  """The `config` dict BOTH arms print (identical for a given N and scaling mode)."""
  weak = scaling == "weak"
  planes = SIDE * world if weak else SIDE
  return {
    "workload": "renumber(in_place=True) of a (%d, 1024, 1024) uint64 synthetic segmentation (%d GiB), %d x 46 x 46 cells of splitmix64 ids, "
                "10%% zeros (%d labels per 1024^3, ~91%% equal neighbours), refit to uint32" % (planes, planes * 8 // 1024, GRID * planes // SIDE, LABELS_PER_SLAB),
    "volume": [planes, SIDE, SIDE], "dtype": "uint64", "cells": [GRID * planes // SIDE, GRID, GRID], "seed": SEED, "zero_frac": ZERO_FRAC,
    "scaling": scaling,
    "parallelism": "single GPU" if world == 1 else "z-slabs of %d planes on %d GPUs, NCCL all-gather key-table merge" % (planes // world, world),
    "l2": "per-GPU input %.0f GiB >> 126 MB L2; the un-timed restore between steps rewrites it" % (planes // world * 8 / 1024.0),
  }


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
  """SM clock / throttle reasons sampled through NVML every few ms DURING the timed region."""

  def __init__(self, index=0):
    self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None

  def start(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      # CUDA_VISIBLE_DEVICES may remap indices; fall back to index 0 if out of range
      try:
        self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
      except Exception:
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
      self.nv = pynvml
      self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception as e:  # pragma: no cover
      self.err = repr(e)
      return
    self.thread = threading.Thread(target=self._run, daemon=True)
    self.thread.start()

  def _run(self):
    nv = self.nv
    while not self.stop_flag:
      try:
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
          reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
          reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.rows.append((sm, int(reasons)))
      except Exception as e:  # pragma: no cover
        self.err = repr(e)
        return
      time.sleep(0.002)

  def stop(self):
    if self.thread is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
    self.stop_flag = True
    self.thread.join(timeout=2)
    nv = self.nv
    names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
    reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
    sm = [r[0] for r in self.rows]
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm), "reasons": reasons,
            "samples": len(sm)}


def import_reference():
  ref_dir = os.path.join(ROOT, "oracle", "_ref")
  if not os.path.isdir(os.path.join(ref_dir, "fastremap")):
    return None
  import importlib.util
  spec = importlib.util.spec_from_file_location("fastremap", os.path.join(ref_dir, "fastremap", "__init__.py"),
                                                submodule_search_locations=[os.path.join(ref_dir, "fastremap")])
  mod = importlib.util.module_from_spec(spec)
  sys.modules["fastremap"] = mod
  try:
    spec.loader.exec_module(mod)
  except Exception:
    sys.modules.pop("fastremap", None)
    return None
  return mod


def cpu_impl():
  """(module, kind): the compiled reference if present, else the oracle port."""
  ref = import_reference()
  if ref is not None:
    return ref, "reference"
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  orc.build()
  return orc, "port"


def host_volume(z0=0, planes=SIDE, nz_total=SIDE, gz=GRID):
  """Planes [z0, z0+planes) of the workload volume, generated on the host by the generator's numpy twin."""
  from concurrent.futures import ThreadPoolExecutor
  from fastremap_b200.synth import synth_numpy
  out = np.empty((planes, SIDE, SIDE), dtype=DTYPE)
  step = 16

  def fill(a):
    b = min(a + step, planes)
    out[a:b] = synth_numpy((b - a, SIDE, SIDE), (gz, GRID, GRID), DTYPE, seed=SEED, zero_frac=ZERO_FRAC, z0=z0 + a, nz_total=nz_total)
  with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
    list(ex.map(fill, range(0, planes, step)))
  return out


def parallel_copy(dst, src):
  """dst[...] = src split over a few threads (numpy releases the GIL in copyto): the un-timed restore of the CPU legs."""
  from concurrent.futures import ThreadPoolExecutor
  k = min(16, os.cpu_count() or 1)
  cuts = np.linspace(0, src.shape[0], k + 1).astype(int)
  with ThreadPoolExecutor(max_workers=k) as ex:
    list(ex.map(lambda i: np.copyto(dst[cuts[i]:cuts[i + 1]], src[cuts[i]:cuts[i + 1]]), range(k)))


def sample_desc(world, scaling):
  if world == 1 or scaling == "strong":
    return "renumber(in_place=True) on the WHOLE 1024^3 uint64 workload volume (2^30 voxels, 8 GiB), input copy excluded, 1 thread (the reference is single-threaded)"
  return ("renumber(in_place=True) on the first 1024 z-planes (rank 0's slab: 2^30 voxels, 8 GiB) of the (%d, 1024, 1024) workload volume, "
          "input copy excluded, 1 thread (the reference is single-threaded)" % (SIDE * world))


# ------------------------------------------------------------------------------------ reference arm


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if rank != 0:
    return
  impl, kind = cpu_impl()
  weak = args.scaling == "weak"
  # the whole volume at N=1 (and for a strong-scaling split); rank 0's 1024^3 slab of the N x larger volume otherwise
  sample = host_volume(0, SIDE, SIDE * world if weak else SIDE, GRID * world if weak else GRID)
  work = np.empty_like(sample)
  nvox = sample.size
  for _ in range(args.warmup):
    parallel_copy(work, sample)
    impl.renumber(work, in_place=True)
  total = 0.0
  for _ in range(args.steps):
    parallel_copy(work, sample)     # un-timed restore
    t0 = time.perf_counter()
    out, mapping = impl.renumber(work, in_place=True)
    total += time.perf_counter() - t0
  value = nvox * args.steps / total / 1e9
  line = {
    "impl": "reference", "metric": "renumber throughput, 1024^3 uint64 synthetic segmentation", "value": value,
    "unit": "Gvox/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
    "dtype": "u64", "data": "synthetic",
    "config": workload_config(world, args.scaling),
    "observed": {"labels": len(mapping) - 1, "out_dtype": str(out.dtype), "voxels_per_step": int(nvox)},
    "cpu_baseline": {"value": value, "unit": "Gvox/s", "cores": 1, "kind": kind, "sample": sample_desc(world, args.scaling),
                     "host_cores_available": os.cpu_count()},
    "e2e": {"value": value, "unit": "Gvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm


def multi_gpu_parity(args, world, rank, planes, n, nz_total, gz, pristine, work):
  """Parity of the sharded path at volume scale (every N > 1 run, outside the timed region):
    digest   the global mapping (labels in id order, ids) hashes to the same value on every rank;
    prefix   ids depend only on the part of the volume in front of a voxel, so rank 0's first 256 z-planes must equal
             what the compiled reference makes of those planes alone -- narrowed ids, ids left at the old width and
             the mapping of the labels met there, bit for bit;
    single   (N = 2) every rank's slab result equals the corresponding half of ONE GPU's renumber of the whole volume.
  Returns a dict for the bench line; a failed check makes bench.py refuse to print a number."""
  import torch
  import torch.distributed as dist
  from fastremap_b200 import api, sharded
  from fastremap_b200 import device as D
  from fastremap_b200.synth import synth_device
  dev = D.dev()
  work.copy_(pristine)
  res = sharded.renumber_sharded(work, n, DTYPE, write_wide=True, offset=rank * n, total=world * n, row_bytes=SIDE * 8)
  torch.cuda.synchronize()
  nl = int(res.n_labels)
  keys = res.keys[: nl * 8].view(torch.int64)
  ids = res.ids[: nl * 8].view(torch.int64)
  pos = torch.arange(1, nl + 1, dtype=torch.int64, device=dev)
  digest = int(((keys * -7046029254386353131) ^ (ids * 0x2545F4914F6CDD1D) ^ pos).sum().item())  # int64 arithmetic wraps: fine for a digest
  mine = torch.tensor([digest, nl], dtype=torch.int64, device=dev)
  allv = torch.zeros(world * 2, dtype=torch.int64, device=dev)
  dist.all_gather_into_tensor(allv, mine)
  rows = allv.cpu().view(world, 2).tolist()
  out = {"labels": nl, "mapping_digest_equal_on_all_ranks": all(r == rows[0] for r in rows)}

  wo = np.dtype(res.out_dtype).itemsize
  narrow_t = res.out[: n * wo].view(D.torch_dtype(res.out_dtype))
  if rank == 0:
    pp = min(256, planes)
    impl, kind = cpu_impl()
    ref_in = host_volume(0, pp, nz_total, gz)
    ref_out, ref_map = impl.renumber(ref_in, in_place=True)
    m = pp * SIDE * SIDE
    g_narrow = narrow_t[:m].cpu().numpy().reshape(pp, SIDE, SIDE)
    g_wide = work[: m * 8].view(torch.int64).cpu().numpy().view(np.uint64).reshape(pp, SIDE, SIDE)
    k = len(ref_map) - 1   # labels met in those planes (0 -> 0 is implicit)
    g_keys = keys[:k].cpu().numpy().view(np.uint64).tolist()
    g_ids = ids[:k].cpu().numpy().tolist()
    out["rank0_first_%d_planes_vs_%s" % (pp, kind)] = bool(
      np.array_equal(g_narrow.astype(np.uint64), ref_out.astype(np.uint64)) and np.array_equal(g_wide, ref_in)
      and dict(zip(g_keys, g_ids)) == {a: b for a, b in ref_map.items() if a != 0})
    del ref_in, ref_out, ref_map, g_narrow, g_wide

  if world == 2:
    ok = torch.ones(1, dtype=torch.int64, device=dev)
    if rank == 0:
      whole = synth_device((planes * 2, SIDE, SIDE), (gz, GRID, GRID), DTYPE, seed=SEED, zero_frac=ZERO_FRAC, z0=0,
                           nz_total=nz_total).reshape(-1).view(torch.uint8)
      single = api.renumber_device(whole, 2 * n, DTYPE, write_wide=False, row_bytes=SIDE * 8)
      s_narrow = single.out[: 2 * n * wo].view(D.torch_dtype(res.out_dtype))
      same = bool(single.out_dtype == res.out_dtype and single.n_labels == nl and torch.equal(s_narrow[:n], narrow_t))
      other = torch.empty(n, dtype=D.torch_dtype(res.out_dtype), device=dev)
      dist.recv(other.view(torch.uint8), src=1)
      same = same and torch.equal(s_narrow[n:], other)
      same = same and torch.equal(single.keys[: nl * 8], res.keys[: nl * 8]) and torch.equal(single.ids[: nl * 8], res.ids[: nl * 8])
      out["n2_slabs_vs_single_gpu_renumber_of_whole_volume"] = same
      del whole, single, other
    else:
      dist.send(narrow_t.view(torch.uint8).contiguous(), dst=0)
  bad = [k_ for k_, v in out.items() if v is False]
  flag = torch.tensor([len(bad)], dtype=torch.int64, device=dev)
  dist.all_reduce(flag)
  if int(flag.item()):
    raise SystemExit("bench.py: multi-GPU parity failed on rank %d: %s" % (rank, bad))
  del res
  return out



def run_ours(args):
  import torch
  import torch.distributed as dist
  import fastremap_b200 as fr
  from fastremap_b200 import _lib, api, sharded
  from fastremap_b200 import device as D
  from fastremap_b200.synth import synth_device

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local_rank)
  numa = D.bind_host_to_gpu(local_rank) if world > 1 else None  # host buffers next to this rank's GPU
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  dev = D.dev()
  weak = args.scaling == "weak"
  planes = SIDE if weak else SIDE // world           # z-planes of this rank's slab
  if SIDE % world:
    raise SystemExit("bench.py: --scaling strong needs a GPU count that divides 1024")
  n = planes * SIDE * SIDE
  nz_total = SIDE * world if weak else SIDE
  gz = GRID * world if weak else GRID
  peak, peak_src = load_peaks()

  pristine = synth_device((planes, SIDE, SIDE), (gz, GRID, GRID), DTYPE, seed=SEED, zero_frac=ZERO_FRAC,
                          z0=rank * planes, nz_total=nz_total).reshape(-1).view(torch.uint8)
  work = torch.empty_like(pristine)
  torch.cuda.synchronize()

  timers = {}

  def step_device():
    if world > 1:
      return sharded.renumber_sharded(work, n, DTYPE, write_wide=True, offset=rank * n, total=world * n, timers=timers, row_bytes=SIDE * 8)
    return api.renumber_device(work, n, DTYPE, write_wide=True, timers=timers, row_bytes=SIDE * 8)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    work.copy_(pristine)
    res = step_device()
  torch.cuda.synchronize()
  out_dtype = str(res.out_dtype)
  n_labels = int(res.n_labels)
  del res
  if world == 1 and n_labels != LABELS_PER_SLAB:
    raise SystemExit("bench.py: the workload volume has %d labels, the config says %d" % (n_labels, LABELS_PER_SLAB))

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  launches0 = _lib.launch_count()
  step_ms, k3_ms, k1_ms, k2_ms, merge_ms = [], [], [], [], []
  stage_ms = {}
  barrier()
  for _ in range(args.steps):
    work.copy_(pristine)          # un-timed restore (also rewrites 8 GiB: L2 holds nothing of the input)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timers.clear()                # stage events of THIS step only
    e0.record()
    res = step_device()
    e1.record()
    torch.cuda.synchronize()
    step_ms.append(e0.elapsed_time(e1))
    if timers:
      k1_ms.append(timers["k1"][0].elapsed_time(timers["k1"][1]))
      k3_ms.append(timers["k3"][0].elapsed_time(timers["k3"][1]))
      k2_ms.append(timers["k2"][0].elapsed_time(timers["k2"][1]))
      if "merge" in timers:
        merge_ms.append(timers["merge"][0].elapsed_time(timers["merge"][1]))
      for name, (a, b) in timers.items():
        stage_ms.setdefault(name, []).append(a.elapsed_time(b))
      stage_ms.setdefault("start_to_k1", []).append(e0.elapsed_time(timers["k1"][0]))
      stage_ms.setdefault("k2_to_k3", []).append(timers["k2"][1].elapsed_time(timers["k3"][0]))
      stage_ms.setdefault("k3_to_end", []).append(timers["k3"][1].elapsed_time(e1))
      if "merge" in timers:
        stage_ms.setdefault("k1_to_merge", []).append(timers["k1"][1].elapsed_time(timers["merge"][0]))
        stage_ms.setdefault("merge_to_k2", []).append(timers["merge"][1].elapsed_time(timers["k2"][0]))
    del res
  barrier()
  launches = _lib.launch_count() - launches0
  clocks = sampler.stop() if rank == 0 else None

  total_ms = float(sum(step_ms))
  per_rank = None
  if world > 1 and args.no_e2e:
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"rank": rank, "ms_per_step": total_ms / args.steps,
                                      "stages_ms": {k: round(float(np.mean(v)), 4) for k, v in stage_ms.items()}})
  if world > 1:
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
  value = n * world * args.steps / (total_ms / 1e3) / 1e9

  if args.no_e2e:
    if rank == 0:
      print(json.dumps({"profiling_run": True, "value": value, "ms_per_step": total_ms / args.steps,
                        "k1_ms": float(np.mean(k1_ms)) if k1_ms else None, "k3_ms": float(np.mean(k3_ms)) if k3_ms else None,
                        "k2_ms": float(np.mean(k2_ms)) if k2_ms else None,
                        "merge_ms": float(np.mean(merge_ms)) if merge_ms else None,
                        "stages_ms": {k: round(float(np.mean(v)), 4) for k, v in stage_ms.items()},
                        "per_rank": per_rank,
                        "gpu_launches": int(launches), "labels": n_labels}))
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- end to end through the public API with pinned host buffers (every rank: its own slab)
  e2e_steps = max(1, min(args.steps, 3))
  host_src = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True)
  host_work = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True)
  host_src.copy_(pristine)
  torch.cuda.synchronize()
  work_np = host_work.numpy().view(DTYPE).reshape(planes, SIDE, SIDE)
  e2e_s, d2h = 0.0, 0
  for i in range(1 + e2e_steps):
    host_work.copy_(host_src)
    barrier()
    t0 = time.perf_counter()
    if world > 1:
      out, mapping = sharded.renumber_host_sharded(work_np, offset=rank * n, total=world * n)
    else:
      out, mapping = fr.renumber(work_np, in_place=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
      t = torch.tensor([dt], dtype=torch.float64, device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      dt = float(t.item())
    if i > 0:
      e2e_s += dt
    d2h = n * 8 + out.nbytes + 2 * 8 * len(mapping)
    del out, mapping
  e2e_value = n * world * e2e_steps / e2e_s / 1e9

  # ---- the same call on an ORDINARY (pageable) numpy array: what a drop-in user passes (N=1 only; informational)
  e2e_pageable = None
  if world == 1:
    del host_work, work_np
    plain = np.empty((SIDE, SIDE, SIDE), dtype=DTYPE)
    src_np = host_src.numpy().view(DTYPE).reshape(SIDE, SIDE, SIDE)
    best = None
    for i in range(2):
      plain[...] = src_np
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      out, mapping = fr.renumber(plain, in_place=True)
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      best = dt if best is None else min(best, dt)
      del out, mapping
    e2e_pageable = {"value": n / best / 1e9, "unit": "Gvox/s",
                    "note": "same call on a pageable numpy array (staging ring + memcpy threads), best of 2; not the headline"}
    del plain, src_np
  del host_src

  # ---- parity of the sharded path at volume scale (N > 1), see multi_gpu_parity
  parity_n = multi_gpu_parity(args, world, rank, planes, n, nz_total, gz, pristine, work) if world > 1 else None

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- dominant-kernel roofline (CUDA events around K3 on the launching stream, rank 0)
  roofline = None
  if k3_ms:
    k3 = float(np.mean(k3_ms))
    achieved = ALG_BYTES_PER_VOXEL * n / (k3 / 1e3) / 1e9
    whole = ALG_BYTES_PER_VOXEL * n * world / (total_ms / args.steps / 1e3) / 1e9 / world   # per GPU
    roofline = {"bound": "hbm", "kernel": "k_relabel_auto<8> (K3: lookup + wide store + fused narrow store)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS, "traffic": None, "traffic_capture": None,
                "peak_source": peak_src, "ms_per_launch": k3,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_VOXEL * n,
                "k1_build_ms": float(np.mean(k1_ms)), "k2_rank_ms": float(np.mean(k2_ms)),
                "table_merge_ms": float(np.mean(merge_ms)) if merge_ms else None,
                "whole_op": {"algorithmic_GBps": whole, "frac": whole / peak, "frac_of_nominal_8TBs": whole / NOMINAL_HBM_GBS}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and planes == SIDE:
      try:
        with open(prof) as f:
          tr = json.load(f)
        roofline["traffic"] = tr.get("k_relabel_dram_bytes_per_launch")
        roofline["traffic_capture"] = tr.get("capture")   # the ncu --set full capture the figure was read from (not this run)
      except Exception:
        pass

  # ---- CPU baseline beside it (rank 0, N=1 only): the compiled reference on the WHOLE volume + bit-exact parity
  cpu_baseline = None
  configs = None
  if world == 1 and not args.no_cpu:
    del pristine, work
    torch.cuda.empty_cache()
    impl, kind = cpu_impl()
    sample = host_volume()
    g_in = np.empty_like(sample)
    parallel_copy(g_in, sample)
    g_out, g_map = fr.renumber(g_in, in_place=True)
    ref_in = sample                 # the reference renumbers the generated volume itself, in place (nothing else needs it)
    t0 = time.perf_counter()
    ref_out, ref_map = impl.renumber(ref_in, in_place=True)
    secs = time.perf_counter() - t0
    parity = bool(g_out.dtype == ref_out.dtype and np.array_equal(g_out, ref_out) and np.array_equal(g_in, ref_in) and g_map == ref_map)
    cpu_baseline = {"value": sample.size / secs / 1e9, "unit": "Gvox/s", "cores": 1, "kind": kind, "sample": sample_desc(1, "weak"),
                    "host_cores_available": os.cpu_count(), "seconds": secs, "parity_vs_gpu": parity,
                    "parity_how": "fastremap_b200.renumber(numpy, in_place=True) on the whole 1024^3 volume vs the reference: narrowed array, "
                                  "ids left in the caller's buffer and the mapping dict, bit-exact"}
    if not parity:
      raise SystemExit("bench.py: GPU result differs from the CPU %s on the 1024^3 volume -- refusing to report a number" % kind)
    del sample, g_in, g_out, g_map, ref_in, ref_out, ref_map
    if not args.no_configs:
      from benchmarks import configs as cfg
      configs = cfg.run_configs(impl, kind, peak, reps=5)

  line = {
    "metric": "renumber throughput, 1024^3 uint64 synthetic segmentation", "value": value, "unit": "Gvox/s",
    "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
    "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
    "config": workload_config(world, args.scaling),
    "observed": {"labels": n_labels, "out_dtype": out_dtype, "host_binding": numa,
                 "hbm_fraction_of_peak_whole_op": ALG_BYTES_PER_VOXEL * n * world * args.steps / (total_ms / 1e3) / 1e9 / (peak * world)},
    "e2e": {"value": e2e_value, "unit": "Gvox/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": int(d2h),
            "steps": e2e_steps,
            "api": "fastremap_b200.renumber(pinned numpy array, in_place=True)" if world == 1 else
                   "fastremap_b200.sharded.renumber_host_sharded(pinned numpy slab) per rank (chunked: upload under K1, download under K3)"},
    "e2e_pageable": e2e_pageable,
    "gpu_launches": int(launches), "clocks": clocks,
  }
  if roofline:
    line["roofline"] = roofline
  if cpu_baseline:
    line["cpu_baseline"] = cpu_baseline
  if configs is not None:
    line["configs"] = configs
  if parity_n is not None:
    line["parity"] = parity_n
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
  ap.add_argument("--no-e2e", action="store_true", help="device-resident loop only (profiling runs; prints a short line)")
  ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 1/2/3/5 block of the N=1 line")
  ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                  help="N>1: weak = one 1024^3 slab per GPU (default; N=8 is BASELINE config 4), strong = ONE 1024^3 volume split N ways")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
