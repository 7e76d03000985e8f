"""
Multi-GPU renumber: one process per GPU, the volume sharded along contiguous ranges of the flat
memory-order array (z-slabs for a C-order (z, y, x) volume).

Rank r owns flat range [offset_r, offset_r + n_r); every voxel of rank r precedes every voxel of
rank r+1 in memory order, so the global first-occurrence index of a label is the minimum over
ranks of (offset_r + local first index).  The only exchange step is the merge of the per-shard
key tables:

  K1 local build -> export (key, first index) -> all-gather (variable length) ->
  every rank inserts all pairs into a fresh table with atomicMin -> K2 rank (identical on all
  ranks) -> K3 relabel of the own slab.

`allgather_varlen` is device-agnostic so the exchange logic is testable with the gloo backend.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def allgather_varlen(t: torch.Tensor, group=None):
  """All-gather of 1-D tensors whose lengths differ per rank (NCCL has no allgatherv): exchange the
  counts, pad to the maximum, gather, slice.  Returns the list of per-rank tensors."""
  world = dist.get_world_size(group)
  count = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
  counts = torch.zeros(world, dtype=torch.int64, device=t.device)
  dist.all_gather_into_tensor(counts, count, group=group)
  counts_h = counts.cpu().tolist()
  m = max(max(counts_h), 1)
  padded = torch.zeros(m, dtype=t.dtype, device=t.device)
  padded[: t.numel()] = t
  out = torch.empty(world * m, dtype=t.dtype, device=t.device)
  dist.all_gather_into_tensor(out, padded, group=group)
  return [out[r * m: r * m + counts_h[r]] for r in range(world)]


def shard_offsets(n_local: int, group=None):
  """(flat offset of this rank, total voxels) from every rank's local voxel count."""
  world, rank = dist.get_world_size(group), dist.get_rank(group)
  dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
  counts = torch.zeros(world, dtype=torch.int64, device=dev)
  dist.all_gather_into_tensor(counts, torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
  c = counts.cpu().tolist()
  return sum(c[:rank]), sum(c)


def allgather_pairs(keys: torch.Tensor, vals: torch.Tensor, group=None):
  """Variable-length all-gather of parallel (key, value) int64 lists with ONE count exchange and ONE
  payload collective: every rank sends [keys | vals] padded to the longest list."""
  world = dist.get_world_size(group)
  n = keys.numel()
  count = torch.tensor([n], dtype=torch.int64, device=keys.device)
  counts = torch.zeros(world, dtype=torch.int64, device=keys.device)
  dist.all_gather_into_tensor(counts, count, group=group)
  counts_h = counts.cpu().tolist()
  m = max(max(counts_h), 1)
  packed = torch.zeros(2 * m, dtype=torch.int64, device=keys.device)
  packed[:n] = keys
  packed[m: m + n] = vals
  out = torch.empty(world * 2 * m, dtype=torch.int64, device=keys.device)
  dist.all_gather_into_tensor(out, packed, group=group)
  ks = [out[r * 2 * m: r * 2 * m + counts_h[r]] for r in range(world)]
  vs = [out[r * 2 * m + m: r * 2 * m + m + counts_h[r]] for r in range(world)]
  return ks, vs


def renumber_sharded(src_buf, n_local, dtype, start=1, preserve_zero=True, write_wide=False, max_labels=None,
                     group=None, offset=None, total=None):
  """Renumber this rank's slab consistently with the whole sharded volume.  Returns RenumberResult
  (keys / ids hold the GLOBAL mapping, identical on every rank).  offset / total (flat offset of the
  slab, total voxels) are exchanged once here unless the caller already knows them."""
  from . import _lib
  from . import api
  from . import device as D
  if offset is None or total is None:
    offset, total = shard_offsets(n_local, group)

  def merge(local: "api.LabelTable"):
    L = _lib.load()
    nl = int(local.read_meta()[_lib.META_COUNT])
    k64, v64 = D.alloc(nl * 8), D.alloc(nl * 8)
    _lib.check(L.frb_table_export(D.ptr(local.slots), local.cap, D.ptr(local.meta), D.ptr(k64), D.ptr(v64), nl, D.stream()),
               "frb_table_export")
    keys, vals = allgather_pairs(k64[: nl * 8].view(torch.int64), v64[: nl * 8].view(torch.int64), group)
    merged = api.LabelTable(dtype, sum(k.numel() for k in keys), _lib.EMPTY)

    def build(t):
      for k, v in zip(keys, vals):
        if k.numel():
          _lib.check(L.frb_table_insert(D.ptr(k), 8, D.ptr(v), 8, 0, k.numel(), _lib.OP_MIN, t.direct, D.ptr(t.slots),
                                        t.cap, D.ptr(t.meta), D.stream()), "frb_table_insert")

    api._run_build(build, merged)
    return merged

  return api.renumber_device(src_buf, n_local, dtype, start=start, preserve_zero=preserve_zero, write_wide=write_wide,
                             max_labels=max_labels, flat_offset=offset, total_n=total, merge=merge)
