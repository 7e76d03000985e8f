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


MERGED_SLOTS_PER_ENTRY = 4


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
                     group=None, offset=None, total=None, timers=None, row_bytes=0):
  """Renumber this rank's slab consistently with the whole sharded volume.  Returns RenumberResult
  (keys / ids hold the GLOBAL mapping, identical on every rank).  offset / total (flat offset of the
  slab, total voxels) are exchanged once here unless the caller already knows them.

  Two host round trips per call: one for the per-rank label counts (sizes of the exchange), one for
  the merged label count (refit dtype) -- everything else is queued on the stream."""
  from . import _lib
  from . import api
  from . import device as D
  L = _lib.load()
  dtype = np.dtype(dtype)
  w = dtype.itemsize
  c = D.code(dtype)
  world = dist.get_world_size(group)
  if offset is None or total is None:
    offset, total = shard_offsets(n_local, group)

  api._ev(timers, "init", 0)
  # volumes of one job look alike: the previous call's exchange sizes AND the local table size it ended up with are
  # remembered per (group, world, dtype, slab size, hint), so a table that had to grow once starts grown next time
  key = (id(group), world, dtype.str, int(n_local), max_labels)
  guess = _exchange_guess.get(key)
  # the local table is only built and exported (K3 looks labels up in the rank's compact table): small, so that its
  # init and export scans are cheap and it stays in L2 while K1 streams the slab
  local = api.LabelTable(dtype, api._default_entries(n_local, max_labels, voxels_per_label=1024), _lib.EMPTY)
  if guess is not None and not local.direct and guess[2] > local.cap:
    local.cap = guess[2]
    local.slots = D.alloc(local.cap * 16)
    local.clear()
  api._ev(timers, "init", 1)
  limits = api._narrow_limits(dtype, start, int(total))
  if guess is not None and limits is not None:
    res = _renumber_sharded_one_shot(src_buf, n_local, dtype, c, start, preserve_zero, write_wide, group, world, offset, total, timers,
                                     row_bytes, local, limits, guess, key)
    if res is not None:
      return res
    # the guess was too small (or a table overflowed): nothing was written, take the two-phase path

  def build(t):
    api._ev(timers, "k1", 0)
    _lib.check(L.frb_renumber_build(D.ptr(src_buf), n_local, c, offset, 1 if preserve_zero else 0, D.ptr(t.slots),
                                    t.cap, D.ptr(t.meta), row_bytes, D.stream()), "frb_renumber_build")
    api._ev(timers, "k1", 1)

  merged, counts, own_keys = _two_phase_merge(build, local, dtype, group, world, timers)
  _exchange_guess[key] = _guess_from(counts, local.cap)
  return _rank_and_relabel(src_buf, n_local, dtype, start, preserve_zero, write_wide, merged, own_keys, counts[dist.get_rank(group)],
                           total, limits, timers)


def _two_phase_merge(build, local, dtype, group, world, timers=None):
  """Export this rank's (label, first index) pairs, all-gather them (sizes first: one host round trip) and insert
  everybody's pairs into a fresh table with atomicMin.  `build(table)` runs K1 over the slab; it is repeated on a
  grown table when any rank's table overflowed (all ranks take the same decision).
  Returns (merged table, per-rank label counts, this rank's exported keys as a uint64 device buffer)."""
  from . import _lib
  from . import api
  from . import device as D
  L = _lib.load()
  while True:
    cap_pairs = local.entries_max()
    k64, v64 = D.alloc(cap_pairs * 8), D.alloc(cap_pairs * 8)
    build(local)
    api._ev(timers, "merge", 0)
    api._ev(timers, "export", 0)
    _lib.check(L.frb_table_export(D.ptr(local.slots), local.cap, D.ptr(local.meta), D.ptr(k64), D.ptr(v64), cap_pairs,
                                  D.stream()), "frb_table_export")
    api._ev(timers, "export", 1)
    # every rank learns every rank's (count, overflow flag, exported pairs) in one small collective
    api._ev(timers, "xchg_meta", 0)
    metas = torch.empty(world * _lib.META_WORDS, dtype=torch.int64, device=local.meta.device)
    dist.all_gather_into_tensor(metas, local.meta, group=group)
    metas_h = metas.cpu().view(world, _lib.META_WORDS).tolist()            # host round trip 1
    api._ev(timers, "xchg_meta", 1)
    counts = [int(r[_lib.META_COUNT]) for r in metas_h]
    overfull = [bool(r[_lib.META_OVERFLOW]) or (not local.direct and 2 * cnt > local.cap) for r, cnt in zip(metas_h, counts)]
    if not any(overfull):
      break
    local.grow(max(counts))  # all ranks take the same decision: tables stay the same size everywhere

  api._ev(timers, "xchg_pairs", 0)
  m = max(max(counts), 1)
  packed = torch.empty(2 * m * 8, dtype=torch.uint8, device=k64.device)
  packed[: m * 8] = k64[: m * 8]
  packed[m * 8:] = v64[: m * 8]
  gathered = torch.empty(world * 2 * m * 8, dtype=torch.uint8, device=k64.device)
  dist.all_gather_into_tensor(gathered, packed, group=group)
  api._ev(timers, "xchg_pairs", 1)

  api._ev(timers, "insert", 0)
  # the merged table is inserted into, ranked (one scan) and projected from -- K3 never sees it: 4 slots per label
  # keep its init and K2's scan short (it used to be sized like a per-voxel lookup table, 8+ slots per label)
  merged = api.LabelTable(dtype, sum(counts), _lib.EMPTY, slots_per_entry=MERGED_SLOTS_PER_ENTRY)
  # the local COUNT word is what every rank exported (export never drops pairs: cap_pairs >= count)
  _lib.check(L.frb_table_insert_gathered(D.ptr(gathered), m, world, D.ptr(metas), _lib.OP_MIN, merged.direct, D.ptr(merged.slots),
                                         merged.cap, D.ptr(merged.meta), D.stream()), "frb_table_insert_gathered")
  api._ev(timers, "insert", 1)
  api._ev(timers, "merge", 1)
  return merged, counts, k64


# K3 looks every voxel up in the rank's own table, and a sparser table means fewer lookups that go beyond the home pair
# (single GPU: 3.52 ms with the 128 MiB table that n / 256 entries give it, 3.80 ms with a 32 MiB one at 88 k labels):
# the own table gets up to 64 slots per label while it stays at 2^23 slots (128 MiB), never fewer than the lookup default.
OWN_TABLE_SLOTS = 1 << 23


def _own_slots_per_entry(own_capacity: int) -> int:
  from . import api
  return max(api.LOOKUP_SLOTS_PER_ENTRY, min(64, OWN_TABLE_SLOTS // max(int(own_capacity), 1)))


def _tables_on_side_stream(dtype, merged_entries, own_capacity):
  """The merged table and the rank's own lookup table, allocated and initialised on the side stream -- under K1 and the
  exchange instead of between them and K3.  Returns (merged, own or None, event to wait for before their first use)."""
  from . import _lib
  from . import api
  from . import device as D
  main = torch.cuda.current_stream()
  side = D.side_stream()
  with torch.cuda.stream(side):
    merged = api.LabelTable(dtype, merged_entries, _lib.EMPTY, slots_per_entry=MERGED_SLOTS_PER_ENTRY)
    own = None if merged.direct else api.LabelTable(dtype, own_capacity, _lib.EMPTY, slots_per_entry=_own_slots_per_entry(own_capacity))
    ready = torch.cuda.Event()
    ready.record(side)
  for t in (merged, own):
    if t is not None:  # allocated on the side stream, used (and outlived) by kernels of the main stream
      t.slots.record_stream(main)
      t.meta.record_stream(main)
  return merged, own, ready


def _own_table(merged, own_keys, own_count_dev, own_capacity, dtype, own=None):
  """A lookup table for K3 that holds only THIS slab's labels with the global ranks K2 has just written into the
  merged table.  At 8 GPUs the merged table is 8x the size of a single GPU's and later ranks' labels sit deeper in
  its probe chains (K3 3.53 -> 3.69 ms); looked up through its own compact table every rank's K3 costs what a
  single GPU's does.  Direct tables (8 / 16-bit labels) are used as they are."""
  from . import _lib
  from . import api
  from . import device as D
  if merged.direct:
    return merged
  if own is None:
    own = api.LabelTable(dtype, own_capacity, _lib.EMPTY, slots_per_entry=_own_slots_per_entry(own_capacity))
  _lib.check(_lib.load().frb_table_project(D.ptr(own_keys), own_count_dev, own_capacity, D.ptr(merged.slots), merged.cap, D.ptr(merged.meta),
                                           D.ptr(own.slots), own.cap, D.ptr(own.meta), D.stream()), "frb_table_project")
  return own


def _rank_and_relabel(src_buf, n_local, dtype, start, preserve_zero, write_wide, merged, own_keys, own_count, total, limits, timers):
  """K2 on the merged table (identical on every rank), then K3 of the own slab through the rank's compact table."""
  import ctypes
  from . import _lib
  from . import api
  rb = api._RankBuffers(merged, dtype)
  api._ev(timers, "k2", 0)
  api._rank(merged, rb, dtype, start, 0, max(int(total), 1))
  api._ev(timers, "k2", 1)
  api._ev(timers, "project", 0)
  lookup = _own_table(merged, own_keys, None, max(int(own_count), 1), dtype)
  api._ev(timers, "project", 1)
  if limits is not None:  # K3 queued behind K2: the refit width is picked on the device
    ranked = torch.cuda.Event()
    ranked.record()
    api._ev(timers, "k3", 0)
    narrow = api._relabel_auto(src_buf, n_local, dtype, lookup, preserve_zero, start, write_wide, limits)
    api._ev(timers, "k3", 1)
    meta = merged.read_meta_after(ranked)                                  # host round trip 2 (K3 keeps running)
  else:
    meta = merged.read_meta()
  if meta[_lib.META_OVERFLOW]:
    raise api.FastremapB200Error("renumber_sharded: merged label table overflowed")
  if limits is not None:
    return api._auto_result(src_buf, narrow, dtype, start, merged, rb, meta)
  return api._renumber_finish(src_buf, n_local, dtype, start, preserve_zero, write_wide, lookup, rb, meta, timers)


def renumber_host_sharded(arr, start=1, preserve_zero=True, group=None, offset=None, total=None, max_labels=None):
  """renumber(in_place=True) of this rank's HOST slab (a C- or F-contiguous, writeable numpy array) of a sharded
  volume: the caller's array receives the ids at the old width, the narrowed copy is returned with the GLOBAL
  mapping -- what fastremap.renumber(whole_volume, in_place=True) leaves in these planes (fastremap.pyx:121-161).

  A chunk pipeline around the one exchange step: K1 consumes 128 MiB chunks as they arrive from the host (upload of
  chunk c+1 under K1 of chunk c); after the merge and K2, K3 relabels chunk by chunk and both result streams (ids
  at the old width, narrowed copy) go back while the next chunk is relabelled; the mapping dict is built meanwhile.
  No id is final before the merge, so the two PCIe directions cannot overlap: the floor is upload + download time."""
  from . import _lib
  from . import api
  from . import device as D
  L = _lib.load()
  D.dev()
  dtype = arr.dtype
  w = dtype.itemsize
  c = D.code(dtype)
  n = arr.size
  nbytes = n * w
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  if offset is None or total is None:
    offset, total = shard_offsets(n, group)
  order = "F" if arr.flags["F_CONTIGUOUS"] and not arr.flags["C_CONTIGUOUS"] else "C"
  host_u8 = D.as_torch_bytes(D.flat_view(arr))
  hint = D.row_bytes(arr.shape, order, dtype)
  chunk = max(api.STREAM_CHUNK_BYTES // 16 * 16, 16)
  bounds = [(a, min(a + chunk, nbytes)) for a in range(0, nbytes, chunk)]
  cur = torch.cuda.current_stream()
  s_in, s_out, s_out2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
  dev_buf = D.alloc(nbytes)
  for s_ in (s_in, s_out, s_out2):
    s_.wait_stream(cur)
  up = D.ChunkUploader(host_u8, dev_buf, bounds, s_in)
  down_w = D.ChunkDownloader(host_u8, s_out)
  down_n = None
  try:
    local = api.LabelTable(dtype, api._default_entries(n, max_labels, voxels_per_label=1024), _lib.EMPTY)

    def build(t):
      for i, (a, b) in enumerate(bounds):
        cur.wait_event(up.event(i))
        _lib.check(L.frb_renumber_build(D.ptr(dev_buf[a:b]), (b - a) // w, c, offset + a // w, 1 if preserve_zero else 0, D.ptr(t.slots),
                                        t.cap, D.ptr(t.meta), hint, D.stream()), "frb_renumber_build")

    merged, counts, own_keys = _two_phase_merge(build, local, dtype, group, world)
    rb = api._RankBuffers(merged, dtype)
    api._rank(merged, rb, dtype, start, 0, max(int(total), 1))
    lookup = _own_table(merged, own_keys, None, max(int(counts[rank]), 1), dtype)
    meta = merged.read_meta()
    if meta[_lib.META_OVERFLOW]:
      raise api.FastremapB200Error("renumber_host_sharded: merged label table overflowed")
    nl = int(meta[_lib.META_ASSIGNED])
    out_dtype = api._out_dtype(dtype, start, nl)
    wo = out_dtype.itemsize
    narrowed = wo < w
    out = arr
    if out_dtype != dtype:
      out = D.host_empty(arr.shape, out_dtype, order)
      down_n = D.ChunkDownloader(D.as_torch_bytes(D.flat_view(out)), s_out2)
    narrow = D.alloc(n * wo) if narrowed else None
    add = int(np.int64(start))
    for a, b in bounds:
      ne = (b - a) // w
      e0 = a // w
      view = dev_buf[a:b]
      nview = narrow[e0 * wo:(e0 + ne) * wo] if narrowed else None
      api._relabel(view, ne, dtype, lookup, _lib.MISS_ERROR, 0, preserve_zero, view, nview, wo if narrowed else 0, add=add)
      if out_dtype != dtype and not narrowed:  # widened ids (start so large that the next id needs a wider dtype)
        nview = api._cast(view, dtype, out_dtype, ne)
      ev = torch.cuda.Event()
      ev.record(cur)
      down_w.submit(a, b, view, ev)
      if down_n is not None:
        down_n.submit(e0 * wo, (e0 + ne) * wo, nview[: ne * wo], ev)
    res = api.RenumberResult()
    res.keys, res.ids, res.n_labels = rb.keys, rb.ids, nl
    mapping = api._mapping_to_dict(res, dtype, preserve_zero)   # host work under the downloads
    down_w.finish()
    if down_n is not None:
      down_n.finish()
    cur.synchronize()
    return out, mapping
  finally:
    up.abort()
    down_w.abort()
    if down_n is not None:
      down_n.abort()


# One-shot exchange.  The two-phase path above needs the per-rank label counts on the HOST between its
# two collectives (they size the padded all-gather and the merged table).  Volumes of one job look
# alike, so the counts of the previous call (+25 %) are used as the sizes of the next one: K1, export,
# ONE all-gather of [meta | keys | first indices], insert, K2 and K3 are queued without a host round
# trip, and the insert kernel itself checks the guess (frb_table_insert_packed raises OVERFLOW, which
# parks K2 and K3) -- the host learns about a miss from the meta block it reads anyway and falls back.
_exchange_guess = {}
one_shot_stats = {"hit": 0, "miss": 0}


def _guess_from(counts, local_cap):
  m = (int(max(counts) * 1.25) + 1024) // 1024 * 1024
  return m, int(sum(counts) * 1.25) + 1024, int(local_cap)


def _renumber_sharded_one_shot(src_buf, n_local, dtype, c, start, preserve_zero, write_wide, group, world, offset, total, timers,
                               row_bytes, local, limits, guess, key):
  from . import _lib
  from . import api
  from . import device as D
  L = _lib.load()
  m, total_labels = guess[0], guess[1]
  MW = _lib.META_WORDS
  block = (MW + 2 * m) * 8
  packed = D.alloc(block)
  api._ev(timers, "k1", 0)
  _lib.check(L.frb_renumber_build(D.ptr(src_buf), n_local, c, offset, 1 if preserve_zero else 0, D.ptr(local.slots),
                                  local.cap, D.ptr(local.meta), row_bytes, D.stream()), "frb_renumber_build")
  api._ev(timers, "k1", 1)
  merged, own, tables_ready = _tables_on_side_stream(dtype, total_labels, m)   # host work and init kernels under K1
  api._ev(timers, "merge", 0)
  api._ev(timers, "export", 0)
  keys_out, vals_out = packed[MW * 8:(MW + m) * 8], packed[(MW + m) * 8:block]
  _lib.check(L.frb_table_export(D.ptr(local.slots), local.cap, D.ptr(local.meta), D.ptr(keys_out), D.ptr(vals_out), m, D.stream()),
             "frb_table_export")
  packed[: MW * 8].view(torch.int64).copy_(local.meta)
  api._ev(timers, "export", 1)
  api._ev(timers, "xchg_pairs", 0)
  gathered = torch.empty(world * block, dtype=torch.uint8, device=packed.device)
  dist.all_gather_into_tensor(gathered, packed[:block], group=group)
  api._ev(timers, "xchg_pairs", 1)
  api._ev(timers, "insert", 0)
  torch.cuda.current_stream().wait_event(tables_ready)
  _lib.check(L.frb_table_insert_packed(D.ptr(gathered), m, world, 0 if local.direct else local.cap, _lib.OP_MIN, merged.direct,
                                       D.ptr(merged.slots), merged.cap, D.ptr(merged.meta), D.stream()), "frb_table_insert_packed")
  api._ev(timers, "insert", 1)
  api._ev(timers, "merge", 1)
  rb = api._RankBuffers(merged, dtype)
  api._ev(timers, "k2", 0)
  api._rank(merged, rb, dtype, start, 0, max(int(total), 1))
  api._ev(timers, "k2", 1)
  api._ev(timers, "project", 0)
  import ctypes
  own_count_dev = ctypes.c_void_p(packed.data_ptr() + 8 * _lib.META_COUNT)   # this rank's label count, still on the device
  lookup = _own_table(merged, keys_out, own_count_dev, m, dtype, own)
  api._ev(timers, "project", 1)
  ranked = torch.cuda.Event()
  ranked.record()
  api._ev(timers, "k3", 0)
  narrow = api._relabel_auto(src_buf, n_local, dtype, lookup, preserve_zero, start, write_wide, limits)
  api._ev(timers, "k3", 1)
  # host round trip (K3 keeps running): the merged meta and every rank's count, through the side stream
  side = D.side_stream()
  side.wait_event(ranked)
  with torch.cuda.stream(side):
    counts_h = gathered.view(torch.int64).view(world, MW + 2 * m)[:, _lib.META_COUNT].to("cpu")
    meta = merged.meta.to("cpu").numpy().view(np.uint64)
  if meta[_lib.META_OVERFLOW]:
    one_shot_stats["miss"] += 1
    _exchange_guess.pop(key, None)
    torch.cuda.current_stream().synchronize()  # the parked K3 must be out of the way before src is walked again
    return None
  one_shot_stats["hit"] += 1
  _exchange_guess[key] = _guess_from([int(x) for x in counts_h.tolist()], local.cap)
  return api._auto_result(src_buf, narrow, dtype, start, merged, rb, meta)


# ------------------------------------------------------------------------------------ other ops
# unique merges per-shard (label, count) lists by SUM, component_map per-shard
# (label, parent) lists by "last shard wins" (= MAX of the run-start index, because every voxel of
# rank r precedes every voxel of rank r+1), the remap family needs no data exchange (every rank
# builds the same table; only the first-missing-index status is reduced), minmax / pixel_pairs
# reduce a few scalars.


def _export_pairs(table):
  """All (key bits, value) pairs of a label table as two int64 device tensors."""
  from . import _lib
  from . import device as D
  L = _lib.load()
  nl = int(table.read_meta()[_lib.META_COUNT])
  k64, v64 = D.alloc(nl * 8), D.alloc(nl * 8)
  _lib.check(L.frb_table_export(D.ptr(table.slots), table.cap, D.ptr(table.meta), D.ptr(k64), D.ptr(v64), nl, D.stream()),
             "frb_table_export")
  return k64[: nl * 8].view(torch.int64), v64[: nl * 8].view(torch.int64)


def unique_sharded(buf, n_local, dtype, max_labels=None, group=None, row_bytes=0, plane_bytes=0):
  """unique(return_counts=True) of the whole sharded volume, replicated on every rank.
  Returns (uniq_buf, counts_buf, n_unique) like api.unique_counts_device."""
  from . import _lib
  from . import api
  from . import device as D
  L = _lib.load()
  dtype = np.dtype(dtype)
  local = api.LabelTable(dtype, api._default_entries(n_local, max_labels), 0)
  ws_bytes = int(L.frb_hist_ws_bytes(n_local))
  ws = D.alloc(ws_bytes)

  def build_local(t):
    _lib.check(L.frb_hist_hash(D.ptr(buf), n_local, D.code(dtype), D.ptr(t.slots), t.cap, D.ptr(t.meta), row_bytes, plane_bytes,
                               D.ptr(ws), ws_bytes, D.stream()), "frb_hist_hash")

  api._run_build(build_local, local)
  keys, counts = allgather_pairs(*_export_pairs(local), group)
  merged = api.LabelTable(dtype, sum(k.numel() for k in keys), 0)

  def build(t):
    for k, v in zip(keys, counts):
      if k.numel():
        _lib.check(L.frb_table_insert(D.ptr(k), 8, D.ptr(v), 8, 0, k.numel(), _lib.OP_ADD, t.direct, D.ptr(t.slots),
                                      t.cap, D.ptr(t.meta), D.stream()), "frb_table_insert")

  meta = api._run_build(build, merged)
  nu = int(meta[_lib.META_COUNT])
  k64, v64 = _export_pairs(merged)
  ws_bytes = int(L.frb_sort_pairs_ws_bytes(nu))
  ws = D.alloc(ws_bytes)
  uniq, cts = D.alloc(nu * dtype.itemsize), D.alloc(nu * 8)
  _lib.check(L.frb_sort_table_pairs(D.ptr(k64), D.ptr(v64), nu, None, D.code(dtype), D.ptr(uniq), D.ptr(cts), D.ptr(ws), ws_bytes,
                                    D.stream()), "frb_sort_table_pairs")
  return uniq, cts, nu


def component_map_sharded(cc_buf, n_local, cc_dtype, parent_buf, parent_dtype, max_labels=None, group=None, row_bytes=0):
  """component_map of the whole sharded pair of volumes, replicated on every rank.
  Returns (keys_buf, parents_buf, n_labels), unordered, like api.component_map_device."""
  from . import _lib
  from . import api
  from . import device as D
  L = _lib.load()
  cc_dtype, parent_dtype = np.dtype(cc_dtype), np.dtype(parent_dtype)
  # a run that crosses a slab boundary starts in the EARLIER slab: every rank learns the label its predecessor ended with
  wc = cc_dtype.itemsize
  world, rank = dist.get_world_size(group), dist.get_rank(group)
  last = torch.zeros(2, dtype=torch.int64, device=cc_buf.device)  # [has voxels, last label bits]
  if n_local:
    last[0] = 1
    last[1:2].view(torch.uint8)[:wc] = cc_buf[(n_local - 1) * wc: n_local * wc]
  lasts = torch.zeros(2 * world, dtype=torch.int64, device=cc_buf.device)
  dist.all_gather_into_tensor(lasts, last, group=group)
  lasts_h = lasts.cpu().view(world, 2).tolist()
  prev = None
  for r in range(rank - 1, -1, -1):  # nearest non-empty slab in front of this one
    if lasts_h[r][0]:
      prev = lasts_h[r][1] & 0xFFFFFFFFFFFFFFFF
      break
  keys_l, pars_l, nl = api.component_map_device(cc_buf, n_local, cc_dtype, parent_buf, parent_dtype, max_labels=max_labels,
                                                row_bytes=row_bytes, prev_label_bits=prev)
  # widen the raw bit patterns to 64 bits for the exchange
  ucc, upar = np.dtype("u%d" % cc_dtype.itemsize), np.dtype("u%d" % parent_dtype.itemsize)
  k64 = api._cast(keys_l, ucc, np.uint64, nl)[: nl * 8].view(torch.int64)
  p64 = api._cast(pars_l, upar, np.uint64, nl)[: nl * 8].view(torch.int64)
  keys, pars = allgather_pairs(k64, p64, group)
  allk, allp = torch.cat(keys), torch.cat(pars)
  total = allk.numel()
  if total == 0:
    return D.alloc(0), D.alloc(0), 0
  # rank-major concatenation: the LAST occurrence of a label comes from the highest rank that saw
  # it, i.e. from the largest run-start index -- positions + MAX, then resolve (like remap_from_array_kv)
  t = api.table_from_arrays(allk, allp, total, np.uint64, op=_lib.OP_MAX, positions=True)
  mk, mp = _export_pairs(t)
  n = mk.numel()
  keys_out = api._cast(mk.view(torch.uint8), np.uint64, ucc, n)
  pars_out = api._cast(mp.view(torch.uint8), np.uint64, upar, n)
  return keys_out, pars_out, n


def relabel_sharded(buf, n_local, dtype, keys_dev, vals_dev, n_keys, policy, fill_bits=0, offset=0, message="{} was not in the remap table.",
                    group=None):
  """remap / mask / mask_except / remap_from_array_kv of this rank's slab, in place.  Every rank builds
  the same table from the same (keys, values); no data is exchanged.  With the error policy the
  first missing flat index is reduced (min) over the ranks and every rank raises the same KeyError."""
  from . import _lib
  from . import api
  from . import device as D
  dtype = np.dtype(dtype)
  t = api.table_from_arrays(keys_dev, vals_dev, n_keys, dtype, op=_lib.OP_SET) if n_keys else api.LabelTable(dtype, 1, 0)
  api._relabel(buf, n_local, dtype, t, policy, fill_bits=fill_bits, dst_wide=buf, batch_rows=api._batch_rows(n_keys))
  if policy != _lib.MISS_ERROR:
    return
  idx = int(t.read_meta()[_lib.META_MISSING_INDEX])
  label = 0
  if idx != _lib.EMPTY:
    w = dtype.itemsize
    label = int(D.download(buf[idx * w:(idx + 1) * w], 1, np.dtype("u%d" % w))[0])
    idx += offset
  first = allreduce_first_missing(idx if idx != _lib.EMPTY else -1, label, group)
  if first is not None:
    raise KeyError(message.format(api._to_py(first, dtype)))


def allreduce_first_missing(global_index: int, label_bits: int, group=None):
  """Min-reduce of (first missing flat index, its label) over the ranks; global_index < 0 = none.
  Returns the label bits of the globally first missing voxel, or None."""
  world = dist.get_world_size(group)
  dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
  big = (1 << 62)
  mine = torch.tensor([global_index if global_index >= 0 else big, label_bits & 0x7FFFFFFFFFFFFFFF, (label_bits >> 63) & 1],
                      dtype=torch.int64, device=dev)
  allv = torch.zeros(world * 3, dtype=torch.int64, device=dev)
  dist.all_gather_into_tensor(allv, mine, group=group)
  rows = allv.cpu().view(world, 3).tolist()
  best = min(rows, key=lambda r: r[0])
  if best[0] >= big:
    return None
  return best[1] | (best[2] << 63)


def scan_sharded(buf, n_local, dtype, group=None):
  """(min, max, pixel_pairs, nonzero) of the whole sharded volume: local fused scan + a reduction of
  the scalars; a pair that straddles a slab boundary is added from the exchanged edge voxels."""
  from . import api
  from . import device as D
  dtype = np.dtype(dtype)
  w = dtype.itemsize
  dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
  if n_local:
    mn, mx, pairs, nz = api._scan(buf, n_local, dtype)
    first = int(D.download(buf[:w], 1, np.dtype("u%d" % w))[0])
    last = int(D.download(buf[(n_local - 1) * w: n_local * w], 1, np.dtype("u%d" % w))[0])
  else:
    mn = mx = pairs = nz = first = last = 0
  return reduce_scan(mn, mx, pairs, nz, first, last, n_local, np.issubdtype(dtype, np.signedinteger), dev, group)


def reduce_scan(mn, mx, pairs, nz, first_bits, last_bits, n_local, signed, dev, group=None):
  """Reduction step of scan_sharded (device-agnostic: testable with gloo).  Returns
  (min, max, pixel_pairs, nonzero) of the whole volume, (None, None, 0, 0) if it is empty."""
  world = dist.get_world_size(group)

  def split(v):  # two non-negative int64 halves of a 64-bit pattern (or of a signed value)
    v &= 0xFFFFFFFFFFFFFFFF
    return [v & 0xFFFFFFFF, v >> 32]

  def join(lo, hi):
    v = lo | (hi << 32)
    return v - (1 << 64) if signed and v >> 63 else v

  mine = torch.tensor([n_local, pairs, nz] + split(mn) + split(mx) + split(first_bits) + split(last_bits),
                      dtype=torch.int64, device=dev)
  allv = torch.zeros(world * mine.numel(), dtype=torch.int64, device=dev)
  dist.all_gather_into_tensor(allv, mine, group=group)
  live = [r for r in allv.cpu().view(world, mine.numel()).tolist() if r[0] > 0]
  if not live:
    return None, None, 0, 0
  gmn = min(join(r[3], r[4]) for r in live)
  gmx = max(join(r[5], r[6]) for r in live)
  straddling = sum(1 for a, b in zip(live[:-1], live[1:]) if (a[9], a[10]) == (b[7], b[8]))
  return gmn, gmx, sum(r[1] for r in live) + straddling, sum(r[2] for r in live)
