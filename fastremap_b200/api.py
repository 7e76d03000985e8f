"""
Host-side mirror of the reference's Python-level wrappers (fastremap/fastremap.pyx, "L2").

Same names, argument names, defaults, return types and exceptions as the reference's public
surface (fastremap/__init__.py:1-24, fastremap/fastremap.pyi), so the reference's own tests read
unchanged against this module.  Every array loop is a CUDA kernel behind the C ABI of
include/fastremap_b200.h; nothing here falls back to the CPU.

Inputs may be numpy arrays (uploaded, results downloaded) or CUDA torch tensors (device-resident:
results are CUDA tensors; with in_place=True a contiguous tensor is rewritten where it lies).

Keyword-only extras that do not change any default: `max_labels` (table sizing hint).
"""
from __future__ import annotations

import ctypes
import functools

import numpy as np
import torch

from . import _lib
from . import device as D
from ._lib import FastremapB200Error
from .dtypes import fit_dtype

_u64 = np.uint64


def _L():
  return _lib.load()


def _pow2_at_least(v: int) -> int:
  v = max(int(v), 1)
  return 1 << (v - 1).bit_length()


def _to_py(bits: int, dtype) -> int:
  """Raw 64-bit pattern at dtype width -> Python int in that dtype."""
  dtype = np.dtype(dtype)
  bits &= (1 << (8 * dtype.itemsize)) - 1
  if np.issubdtype(dtype, np.signedinteger) and bits >> (8 * dtype.itemsize - 1):
    bits -= 1 << (8 * dtype.itemsize)
  return bits


def _to_bits(value, dtype) -> int:
  """Python int -> raw pattern at dtype width; OverflowError when not representable (the
  reference gets that from Cython's integer conversion)."""
  dtype = np.dtype(dtype)
  info = np.iinfo(dtype)
  value = int(value)
  if value < info.min or value > info.max:
    raise OverflowError("value too large to convert to %s" % dtype.name)
  return value & ((1 << (8 * dtype.itemsize)) - 1)


def _require_int(dtype, what):
  if not D.is_int_dtype(dtype):
    raise TypeError("fastremap_b200.%s supports integer dtypes (uint8-uint64, int8-int64); got %s. "
                    "There is no CPU fallback for other dtypes." % (what, np.dtype(dtype)))


# ------------------------------------------------------------------------------------- label table


class LabelTable:
  """Device label table: `cap` 16-byte slots + meta block (see include/fastremap_b200.h)."""

  def __init__(self, dtype, min_entries: int, init_val: int, slots_per_entry: int = 2):
    """slots_per_entry: 2 for tables that are BUILT from a volume (load factor <= 1/2: their cost is
    the init / export scans), LOOKUP_SLOTS_PER_ENTRY for tables that are only looked up per voxel."""
    dtype = np.dtype(dtype)
    self.dtype = dtype
    direct_cap = int(_L().frb_table_direct_cap(D.code(dtype)))
    self.direct = 1 if direct_cap else 0
    self.cap = direct_cap if direct_cap else max(_pow2_at_least(slots_per_entry * max(min_entries, 1)), 4096)
    self.init_val = init_val & _lib.EMPTY
    self.slots = D.alloc(self.cap * 16)
    self.meta = D.alloc_meta()
    self.clear()

  def clear(self):
    _lib.check(_L().frb_table_init(D.ptr(self.slots), self.cap, D.ptr(self.meta), self.init_val, D.stream()),
               "frb_table_init")

  def grow(self, at_least_entries: int = 0):
    if self.direct:
      raise FastremapB200Error("direct label table cannot overflow")
    self.cap = max(self.cap * 4, _pow2_at_least(4 * max(at_least_entries, 1)))
    self.slots = D.alloc(self.cap * 16)
    self.clear()

  def read_meta(self) -> np.ndarray:
    return D.read_u64(self.meta)

  def read_meta_after(self, event) -> np.ndarray:
    """The meta block as it is once `event` (recorded on the launching stream) has happened, read
    through a side stream: the host does not wait for kernels queued behind the event (renumber's K3
    leaves every word the host needs untouched)."""
    side = D.side_stream()
    side.wait_event(event)
    with torch.cuda.stream(side):
      host = self.meta.to("cpu", non_blocking=False)
    return host.numpy().view(np.uint64)

  def entries_max(self) -> int:
    """Most labels this table is allowed to hold before it is grown (load factor 0.5; +1 for the
    out-of-band all-ones key)."""
    return self.cap + 1 if self.direct else self.cap // 2 + 2

  def overfull(self, meta) -> bool:
    if self.direct:
      return False
    return bool(meta[_lib.META_OVERFLOW]) or int(meta[_lib.META_COUNT]) * 2 > self.cap


def _default_entries(n: int, max_labels, voxels_per_label: int = 256) -> int:
  """Initial sizing of a hash table for an n-voxel volume when the label count is unknown:
  n/256 labels (4.2 M at 1024^3; a 128 MiB table), at least 32 Ki; grown x4 on overflow.  renumber keeps this
  generous default because K3 looks every voxel up in the same table and a sparser table means shorter probe
  chains (measured: K3 3.52 ms with the 128 MiB table, 3.80 ms with a 32 MiB one at 88 k labels).  The counting
  pass of unique only ever ADDS to its table, from a log, and prefers one that stays resident in the 126 MB L2:
  it passes voxels_per_label=512 (a 64 MiB table at 1024^3: half the init and export scans; at 32 MiB a volume with 900 k labels loads it to 0.43 and folding the log takes 0.39 instead of 0.05 ms)."""
  if max_labels is not None:
    return int(max_labels) + 1
  return int(min(max(n // voxels_per_label, 1 << 15), 1 << 27))


def _run_build(fn, table: LabelTable):
  """Run a table-building launch; on overflow grow the table and repeat.  Returns meta."""
  while True:
    fn(table)
    meta = table.read_meta()
    if not table.overfull(meta):
      return meta
    table.grow(int(meta[_lib.META_COUNT]))


# A per-voxel lookup costs the length of its dependent probe chain (the warp waits for its slowest
# lane), not the size of the table: measured on B200, remap of 1024^3 uint64 through 10^6 labels
# takes 6.8 ms at load 0.48, 4.2 ms at 0.24, 3.8 ms at 0.12.  So lookup tables get 8+ slots per label.
LOOKUP_SLOTS_PER_ENTRY = 8
# Sparser still pays: a warp walks the slow path (probes beyond the home pair) as soon as ONE of its 64
# lookups needs it, which at load 1/8 (both slots of a pair taken: 1/64) is most warps.  Measured at 1024^3
# uint64: remap through 10^6 labels 3.60 / 3.41 / 3.07 ms and mask_except with 5 M kept labels 4.52 / 4.18 /
# 4.00 ms at 8 / 16 / 32 slots per label (benchmarks/exp_slots_per_entry.py).  So: up to 32, as long as
# the table stays under 2 GiB.
LOOKUP_SLOTS_PER_ENTRY_MAX = 32
LOOKUP_TABLE_MAX_BYTES = 2 << 30


def _lookup_slots_per_entry(n_labels: int) -> int:
  spe = LOOKUP_SLOTS_PER_ENTRY_MAX
  while spe > LOOKUP_SLOTS_PER_ENTRY and n_labels * spe * 16 > LOOKUP_TABLE_MAX_BYTES:
    spe //= 2
  return spe
# ... and once the table no longer fits L1, K3 issues the table loads of two rows of a warp-tile
# before it consumes either (mask_except with 5 M labels: 5.0 -> 4.5 ms).
BATCH_LOOKUP_MIN_LABELS = 1 << 14


def _batch_rows(n_labels: int) -> int:
  return 2 if n_labels >= BATCH_LOOKUP_MIN_LABELS else 0


def table_from_arrays(keys_dev, vals_dev, L: int, dtype, op=_lib.OP_SET, positions=False, resolve=True) -> LabelTable:
  """Build a lookup table from device arrays of L keys / values of `dtype` (K: frb_table_insert).
  positions: the value of a key is its position in the list (combined with `op`), which `resolve`
  then replaces by vals[position]."""
  dtype = np.dtype(dtype)
  w = dtype.itemsize
  init = 0
  table = LabelTable(dtype, L, init, slots_per_entry=_lookup_slots_per_entry(L))

  def build(t):
    _lib.check(_L().frb_table_insert(D.ptr(keys_dev), w, None if positions else D.ptr(vals_dev), w, 0, L, op,
                                     t.direct, D.ptr(t.slots), t.cap, D.ptr(t.meta), D.stream()), "frb_table_insert")

  _run_build(build, table)
  if positions and resolve:
    _lib.check(_L().frb_table_resolve(D.ptr(vals_dev), w, D.ptr(table.slots), table.cap, D.ptr(table.meta), D.stream()),
               "frb_table_resolve")
  return table


def _relabel(src_buf, n, dtype, table: LabelTable, policy, fill_bits=0, zero_fixed=False, dst_wide=None, dst_narrow=None,
             narrow_width=0, add=0, guard_overflow=False, batch_rows=0, plane_bytes=0):
  _lib.check(_L().frb_relabel(D.ptr(src_buf), D.ptr(dst_wide), D.ptr(dst_narrow), narrow_width, n, D.code(dtype),
                              D.ptr(table.slots), table.cap, table.direct, D.ptr(table.meta), policy, fill_bits,
                              1 if zero_fixed else 0, add & _lib.EMPTY, 1 if guard_overflow else 0, batch_rows,
                              plane_bytes, D.stream()), "frb_relabel")


def _cast(src_buf, src_dtype, dst_dtype, n):
  dst = D.alloc(n * np.dtype(dst_dtype).itemsize)
  _lib.check(_L().frb_cast(D.ptr(src_buf), D.code(src_dtype), D.ptr(dst), D.code(dst_dtype), n, D.stream()), "frb_cast")
  return dst


# ------------------------------------------------------------------------------------- scans


def _scan(vol_buf, n, dtype):
  """(min, max, pixel_pairs, nonzero) of a flat device buffer in one fused pass."""
  out = torch.zeros(4, dtype=torch.int64, device=D.dev())
  _lib.check(_L().frb_minmax_pairs(D.ptr(vol_buf), n, D.code(dtype), D.ptr(out), D.stream()), "frb_minmax_pairs")
  r = D.read_u64(out)
  return _to_py(int(r[0]), dtype), _to_py(int(r[1]), dtype), int(r[2]), int(r[3])


def minmax(arr):
  """(min(arr), max(arr)) in a single pass; (None, None) if empty (reference: fastremap.pyx:70-93)."""
  if not isinstance(arr, torch.Tensor):
    arr = np.asarray(arr)
    if arr.size == 0:
      return None, None
  elif arr.numel() == 0:
    return None, None
  v = D.ingest(arr, alias_ok=True)
  if v.dtype == np.dtype("bool"):
    v.dtype = np.dtype("u1")
  _require_int(v.dtype, "minmax")
  mn, mx, _, _ = _scan(v.buf, v.n, v.dtype)
  return mn, mx


def pixel_pairs(labels):
  """Number of equal adjacent memory locations (reference: fastremap.pyx:829-853)."""
  size = labels.numel() if isinstance(labels, torch.Tensor) else np.asarray(labels).size
  if size == 0:
    return 0
  v = D.ingest(labels, alias_ok=True)
  _require_int(v.dtype, "pixel_pairs")
  return _scan(v.buf, v.n, v.dtype)[2]


def foreground(arr):
  """Number of non-zero voxels (reference: fastremap.pyx:1407-1421); integer dtypes only."""
  is_t = isinstance(arr, torch.Tensor)
  if not is_t:
    arr = np.asarray(arr)
  adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  if not D.is_int_dtype(adtype):
    raise TypeError("No matching signature found")
  if (arr.numel() if is_t else arr.size) == 0:
    return 0
  v = D.ingest(arr, alias_ok=True)
  return _scan(v.buf, v.n, v.dtype)[3]


def indices(arr, value):
  """Ascending positions i with arr[i] == value, as uint64 (reference: fastremap.pyx:104-119).
  1-D integer (or bool) arrays; a value the dtype cannot hold raises OverflowError like the reference."""
  is_t = isinstance(arr, torch.Tensor)
  if not is_t:
    arr = np.asarray(arr)
  nd = arr.dim() if is_t else arr.ndim
  if nd != 1:
    raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % nd)
  adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  if adtype == np.dtype("bool"):
    arr = arr.view(torch.uint8) if is_t else arr.view(np.uint8)
    adtype = np.dtype("u1")
  _require_int(adtype, "indices")
  bits = _to_bits(value, adtype)
  n = arr.numel() if is_t else arr.size

  def wrap(a_np):
    return torch.from_numpy(a_np.view(np.int64)).to(arr.device).view(torch.uint64) if is_t else a_np

  if n == 0:
    return wrap(np.zeros(0, dtype=np.uint64))
  v = D.ingest(arr, alias_ok=True)
  L = _L()
  meta = D.alloc_meta()
  ws_bytes = int(L.frb_indices_ws_bytes(n))
  ws = D.alloc(ws_bytes)
  _lib.check(L.frb_indices_count(D.ptr(v.buf), n, D.code(adtype), bits, D.ptr(meta), D.ptr(ws), ws_bytes, D.stream()),
             "frb_indices_count")
  cnt = int(D.read_u64(meta)[_lib.META_LIST_COUNT])
  out = D.alloc(cnt * 8)
  _lib.check(L.frb_indices_emit(D.ptr(v.buf), n, D.code(adtype), bits, D.ptr(out), cnt, D.ptr(ws), D.stream()), "frb_indices_emit")
  if is_t:
    return out[: cnt * 8].view(torch.uint64)
  return D.download(out, cnt, np.uint64)


def refit(arr, value=None, increase_only=False, exotics=False):
  """Resize to the smallest dtype of the same kind that fits `value` (reference: fastremap.pyx:259-301)."""
  is_t = isinstance(arr, torch.Tensor)
  adtype = D.numpy_dtype(arr.dtype) if is_t else np.asarray(arr).dtype
  if value is None:
    min_value, max_value = minmax(arr)
    if min_value is None or max_value is None:
      min_value = 0
      max_value = 0
    value = max_value if abs(max_value) > abs(min_value) else min_value
  dtype = np.dtype(fit_dtype(adtype, value, exotics=exotics))
  if increase_only and dtype.itemsize <= adtype.itemsize:
    return arr
  elif dtype == adtype:
    return arr
  _require_int(adtype, "refit")
  v = D.ingest(arr, alias_ok=True)
  out = _cast(v.buf, v.dtype, dtype, v.n) if v.n else D.alloc(0)
  return D.emit(out, dtype, v.shape, v.order, v.is_torch)


# ------------------------------------------------------------------------------------- renumber


class RenumberResult:
  """Device-side result of the renumber kernels (all buffers are flat uint8 CUDA tensors)."""
  __slots__ = ("out", "out_dtype", "keys", "ids", "n_labels", "table", "wide", "keys_host", "ids_host")

  def __init__(self):
    self.keys_host = self.ids_host = None  # set by the small-array path: the mapping came back with the counters


def _ev(timers, name, which):
  if timers is not None:
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    timers.setdefault(name, [None, None])[which] = e


class _RankBuffers:
  """Mapping arrays (label / id, indexed by rank) and the K2 workspace, sized for a table."""

  def __init__(self, table: LabelTable, dtype):
    self.capacity = table.entries_max()
    w = np.dtype(dtype).itemsize
    self.keys = D.alloc(self.capacity * w)
    self.ids = D.alloc(self.capacity * w)
    self.ws_bytes = int(_L().frb_renumber_rank_ws_bytes(self.capacity))
    self.ws = D.alloc(self.ws_bytes)


def _rank(table: LabelTable, rb: _RankBuffers, dtype, start, index_base, index_count):
  _lib.check(_L().frb_renumber_rank(D.ptr(table.slots), table.cap, D.ptr(table.meta), D.code(dtype), int(np.int64(start)),
                                    index_base, index_count, D.ptr(rb.keys), D.ptr(rb.ids), rb.capacity, D.ptr(rb.ws),
                                    rb.ws_bytes, D.stream()), "frb_renumber_rank")


def _out_dtype(dtype, start, nl):
  """fastremap.pyx:253-255: fit to the NEXT id (typed, so it wraps in the array dtype) unless |start| is larger."""
  factor = _to_py(int(np.int64(start)) + nl, dtype)
  if abs(start) > abs(factor):
    factor = start
  return np.dtype(fit_dtype(dtype, factor))


def _narrow_limits(dtype, start, n_max):
  if not isinstance(start, (int, np.integer)):
    return None
  return _narrow_limits_cached(np.dtype(dtype).str, int(start), int(n_max))


@functools.lru_cache(maxsize=256)
def _narrow_limits_cached(dtype, start, n_max):
  """(lim1, lim2, lim4) for frb_relabel_auto: the refit width is w' iff n_labels < lim[w'] (and not
  < the next smaller limit).  None when the width is not a monotone function of the label count on
  [0, n_max] (negative start, counters that wrap in the array dtype: fastremap.pyx:225, :253-255)."""
  dtype = np.dtype(dtype)
  info = np.iinfo(dtype)
  if start < 0 or start + n_max > int(info.max):
    return None
  lims = []
  for wo in (1, 2, 4):
    if _out_dtype(dtype, start, 0).itemsize > wo:
      lims.append(0)
      continue
    lo, hi = 0, int(n_max)  # largest nl in [0, n_max] whose refit width is <= wo
    while lo < hi:
      mid = (lo + hi + 1) // 2
      if _out_dtype(dtype, start, mid).itemsize <= wo:
        lo = mid
      else:
        hi = mid - 1
    lims.append(lo + 1)
  return tuple(lims)


def _relabel_auto(src_buf, n, dtype, table: LabelTable, preserve_zero, start, write_wide, limits):
  """K3 with the refit width picked on the device (queued behind K2 without a host round trip).
  Returns the narrow buffer (n * w/2 bytes; which part is meaningful is known once meta is read)."""
  w = np.dtype(dtype).itemsize
  narrow = D.alloc(n * max(w // 2, 1))
  _lib.check(_L().frb_relabel_auto(D.ptr(src_buf), D.ptr(src_buf), D.ptr(narrow), n, D.code(dtype), D.ptr(table.slots), table.cap,
                                   table.direct, D.ptr(table.meta), 1 if preserve_zero else 0, int(np.int64(start)) & _lib.EMPTY,
                                   1 if write_wide else 0, limits[0], limits[1], limits[2], D.stream()), "frb_relabel_auto")
  return narrow


def _auto_result(src_buf, narrow, dtype, start, table, rb, meta) -> RenumberResult:
  nl = int(meta[_lib.META_ASSIGNED])
  res = RenumberResult()
  res.keys, res.ids, res.n_labels, res.table = rb.keys, rb.ids, nl, table
  res.out_dtype = _out_dtype(dtype, start, nl)
  narrowed = res.out_dtype.itemsize < np.dtype(dtype).itemsize
  res.out = narrow if narrowed else src_buf
  res.wide = src_buf  # holds ids when write_wide or when nothing narrower fitted; callers use it only then
  return res


def renumber_device(src_buf, n, dtype, start=1, preserve_zero=True, write_wide=False, max_labels=None,
                    flat_offset=0, total_n=None, timers=None, row_bytes=0) -> RenumberResult:
  """
  K1 -> K2 -> K3 on a flat device buffer.
    write_wide: also overwrite src_buf with the new ids at the input width (in_place=True).
    timers: optional dict; receives CUDA event pairs "k1" / "k2" / "k3" recorded on the launching stream.
    row_bytes: length of the volume's fastest axis in bytes (0 = unknown): a hint that aims K1's
    "same as the voxels directly above" filter.
  The three kernels are queued back to back: the label count stays on the device (K2 reads it there,
  K3 picks the refit width from it), and the host reads the table's meta block once, afterwards.
  """
  dtype = np.dtype(dtype)
  c = D.code(dtype)
  L = _L()
  table = LabelTable(dtype, _default_entries(n, max_labels), _lib.EMPTY)
  index_count = max(int(total_n if total_n is not None else n + flat_offset), 1)

  def build(t):
    _ev(timers, "k1", 0)
    _lib.check(L.frb_renumber_build(D.ptr(src_buf), n, c, flat_offset, 1 if preserve_zero else 0, D.ptr(t.slots), t.cap,
                                    D.ptr(t.meta), row_bytes, D.stream()), "frb_renumber_build")
    _ev(timers, "k1", 1)

  while True:
    build(table)
    rb = _RankBuffers(table, dtype)   # allocated while K1 runs (host work in front of the first launch is on the critical path)
    _ev(timers, "k2", 0)
    _rank(table, rb, dtype, start, 0, index_count)
    _ev(timers, "k2", 1)
    limits = _narrow_limits(dtype, start, min(n, rb.capacity))  # host work while K1 runs (cached across calls)
    if limits is not None:
      ranked = torch.cuda.Event()
      ranked.record()
      _ev(timers, "k3", 0)
      narrow = _relabel_auto(src_buf, n, dtype, table, preserve_zero, start, write_wide, limits)
      _ev(timers, "k3", 1)
      meta = table.read_meta_after(ranked)  # K3 keeps running while the host learns the label count
    else:
      meta = table.read_meta()
    if not meta[_lib.META_OVERFLOW]:
      break
    table.grow(int(meta[_lib.META_COUNT]))  # K3 is guarded by the flag: the input is intact, start over
  if limits is not None:
    return _auto_result(src_buf, narrow, dtype, start, table, rb, meta)
  return _renumber_finish(src_buf, n, dtype, start, preserve_zero, write_wide, table, rb, meta, timers)


def _renumber_finish(src_buf, n, dtype, start, preserve_zero, write_wide, table, rb, meta, timers=None) -> RenumberResult:
  """K3 of renumber once the table holds ranks and the host knows their number: pick the refit dtype,
  relabel (+ fused narrow store).  The general path (any `start`, wrapping counters)."""
  w = dtype.itemsize
  nl = int(meta[_lib.META_ASSIGNED])
  out_dtype = _out_dtype(dtype, start, nl)
  wo = out_dtype.itemsize
  res = RenumberResult()
  res.keys, res.ids, res.n_labels, res.table, res.out_dtype = rb.keys, rb.ids, nl, table, out_dtype
  narrow = D.alloc(n * wo) if wo < w else None
  need_wide = write_wide or wo >= w
  wide_dst = src_buf if need_wide else None
  _ev(timers, "k3", 0)
  _relabel(src_buf, n, dtype, table, _lib.MISS_ERROR, 0, preserve_zero, wide_dst, narrow, wo if narrow is not None else 0,
           add=int(np.int64(start)))
  _ev(timers, "k3", 1)
  res.wide = src_buf if need_wide else None
  if wo < w:
    res.out = narrow
  elif wo == w:
    res.out = src_buf
  else:
    res.out = _cast(src_buf, dtype, out_dtype, n)
  return res


# ------------------------------------------------------------------------------------- small arrays
# Up to this many voxels renumber / unique run as ONE cooperative kernel (csrc/frb_small.cu): a dozen launches and two
# host round trips are what a 32^3 call used to cost (0.4-0.5 ms; the reference's serial loop: 0.1 ms).
SMALL_MAX_VOXELS = 1 << 21


def _small_buffers(n, dtype):
  """(workspace, head) for the small kernels: head = [meta | keys | ids] comes back to the host with one copy."""
  L = _L()
  c = D.code(dtype)
  w = np.dtype(dtype).itemsize
  cap_list = int(L.frb_small_list_capacity(n, c))
  ws_bytes = int(L.frb_small_ws_bytes(n, c))
  ws = D.alloc(ws_bytes)
  lb = (cap_list * max(w, 8) + 255) // 256 * 256
  head = D.alloc(128 + 2 * lb)
  return ws, ws_bytes, head, lb, cap_list


def _renumber_small(src_buf, n, dtype, start, preserve_zero, write_wide):
  """renumber of a small flat device buffer with one kernel launch.  Returns a RenumberResult (mapping already on the
  host), or None when the array holds more labels than the small table takes (src_buf is untouched then) or the
  refit width is not a monotone function of the label count (negative / wrapping `start`)."""
  dtype = np.dtype(dtype)
  limits = _narrow_limits(dtype, start, n)
  if limits is None:
    return None
  L = _L()
  w = dtype.itemsize
  ws, ws_bytes, head, lb, cap_list = _small_buffers(n, dtype)
  narrow = D.alloc(n * w)
  wide = src_buf if write_wide else D.alloc(n * w)  # the relabel phase reads a[i] and writes out[i]: aliasing is safe
  meta_ptr = ctypes.c_void_p(head.data_ptr())
  keys_ptr = ctypes.c_void_p(head.data_ptr() + 128)
  ids_ptr = ctypes.c_void_p(head.data_ptr() + 128 + lb)
  _lib.check(L.frb_renumber_small(D.ptr(src_buf), n, D.code(dtype), int(np.int64(start)), 1 if preserve_zero else 0, D.ptr(wide), D.ptr(narrow),
                                  limits[0], limits[1], limits[2], keys_ptr, ids_ptr, meta_ptr, D.ptr(ws), ws_bytes, D.stream()),
             "frb_renumber_small")
  host = head.cpu().numpy()                       # the one host round trip: counters + mapping
  meta = host[:128].view(np.uint64)
  if meta[_lib.META_OVERFLOW]:
    return None
  nl = int(meta[_lib.META_ASSIGNED])
  res = RenumberResult()
  res.n_labels, res.table = nl, None
  res.keys, res.ids = head[128:128 + lb], head[128 + lb:128 + 2 * lb]
  res.keys_host = host[128:128 + nl * w].view(dtype)
  res.ids_host = host[128 + lb:128 + lb + nl * w].view(dtype)
  res.out_dtype = _out_dtype(dtype, start, nl)
  res.wide = wide
  if res.out_dtype.itemsize < w:
    res.out = narrow
  elif res.out_dtype.itemsize == w:
    res.out = wide
  else:
    res.out = _cast(wide, dtype, res.out_dtype, n)
  return res


def _unique_small(buf, n, dtype):
  """(uniq_buf, counts_buf, n_unique) of a small flat device buffer with one kernel launch, or None (too many labels)."""
  dtype = np.dtype(dtype)
  L = _L()
  w = dtype.itemsize
  ws, ws_bytes, head, lb, cap_list = _small_buffers(n, dtype)
  meta_ptr = ctypes.c_void_p(head.data_ptr())
  uniq_ptr = ctypes.c_void_p(head.data_ptr() + 128)
  cts_ptr = ctypes.c_void_p(head.data_ptr() + 128 + lb)
  _lib.check(L.frb_unique_small(D.ptr(buf), n, D.code(dtype), uniq_ptr, cts_ptr, meta_ptr, D.ptr(ws), ws_bytes, D.stream()), "frb_unique_small")
  meta = D.read_u64(head[:128].view(torch.int64))
  if meta[_lib.META_OVERFLOW]:
    return None
  return head[128:128 + lb], head[128 + lb:128 + 2 * lb], int(meta[_lib.META_LIST_COUNT])


# Host arrays at least this large are renumbered by the chunk pipeline below.
STREAM_MIN_BYTES = 256 << 20
STREAM_CHUNK_BYTES = 128 << 20
STREAM_LOOKAHEAD = 2


def _grow_rehash(table: LabelTable, dtype) -> LabelTable:
  """A bigger table holding the same (label, value) pairs -- ranks and pending first indices alike."""
  L = _L()
  meta = table.read_meta()
  count = int(meta[_lib.META_COUNT])
  k64, v64 = D.alloc(count * 8), D.alloc(count * 8)
  _lib.check(L.frb_table_export(D.ptr(table.slots), table.cap, D.ptr(table.meta), D.ptr(k64), D.ptr(v64), count, D.stream()),
             "frb_table_export")
  new = LabelTable(dtype, 1, _lib.EMPTY)
  new.cap = max(table.cap * 4, _pow2_at_least(4 * max(count, 1)))
  new.slots = D.alloc(new.cap * 16)
  new.clear()
  _lib.check(L.frb_table_insert(D.ptr(k64), 8, D.ptr(v64), 8, 0, count, _lib.OP_SET, 0, D.ptr(new.slots), new.cap,
                                D.ptr(new.meta), D.stream()), "frb_table_insert")
  new.meta[_lib.META_ASSIGNED] = table.meta[_lib.META_ASSIGNED]
  return new


def _renumber_streamed(arr: np.ndarray, start, preserve_zero, in_place_eff, max_labels):
  """
  renumber of a (large) host array as a chunk pipeline.  The id of a label depends only on the part
  of the array in front of its first occurrence, so chunk c can be built, ranked and relabelled
  (K1, K2, K3) as soon as it has arrived, while chunk c+1 is still on its way to the device and the
  ids of chunk c-1 are on their way back: H2D, kernels and D2H run on three streams and the PCIe
  link is busy in both directions.  Only the refit dtype needs the final label count, so the
  narrowed copy is cast and downloaded at the end.
  Returns (result array, mapping) like renumber().
  """
  D.dev()
  dtype = arr.dtype
  n = arr.size
  nbytes = n * dtype.itemsize
  order = "F" if arr.flags["F_CONTIGUOUS"] else "C"
  host_u8 = D.as_torch_bytes(D.flat_view(arr))
  hint = D.row_bytes(arr.shape, order, dtype)
  chunk = max(STREAM_CHUNK_BYTES // 16 * 16, 16)
  bounds = [(a, min(a + chunk, nbytes)) for a in range(0, nbytes, chunk)]
  C = len(bounds)

  cur = torch.cuda.current_stream()
  s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
  dev_buf = D.alloc(nbytes)
  table = LabelTable(dtype, _default_entries(n, max_labels), _lib.EMPTY)
  rb = _RankBuffers(table, dtype)
  flags = torch.zeros((C, _lib.META_WORDS), dtype=torch.int64, pin_memory=True)
  s_in.wait_stream(cur)
  s_out.wait_stream(cur)

  ev_meta, ev_k3 = [None] * C, [None] * C
  up = D.ChunkUploader(host_u8, dev_buf, bounds, s_in)      # pageable arrays go through a staging ring
  down = D.ChunkDownloader(host_u8, s_out) if in_place_eff else None
  try:
    return _renumber_streamed_run(arr, start, preserve_zero, in_place_eff, dtype, order, hint, bounds, dev_buf, table, rb, flags,
                                  cur, s_out, up, down, ev_meta, ev_k3)
  finally:
    up.abort()
    if down is not None:
      down.abort()


def _renumber_streamed_run(arr, start, preserve_zero, in_place_eff, dtype, order, hint, bounds, dev_buf, table, rb, flags,
                           cur, s_out, up, down, ev_meta, ev_k3):
  L = _L()
  w = dtype.itemsize
  c = D.code(dtype)
  n = arr.size
  C = len(bounds)
  add = int(np.int64(start))

  def enqueue(i):
    a, b = bounds[i]
    ne = (b - a) // w
    cur.wait_event(up.event(i))
    view = dev_buf[a:b]
    _lib.check(L.frb_renumber_build(D.ptr(view), ne, c, a // w, 1 if preserve_zero else 0, D.ptr(table.slots), table.cap,
                                    D.ptr(table.meta), hint, D.stream()), "frb_renumber_build")
    _rank(table, rb, dtype, start, a // w, ne)
    flags[i].copy_(table.meta, non_blocking=True)
    ev_meta[i] = torch.cuda.Event()
    ev_meta[i].record(cur)
    _relabel(view, ne, dtype, table, _lib.MISS_ERROR, 0, preserve_zero, view, None, 0, add=add, guard_overflow=True)
    ev_k3[i] = torch.cuda.Event()
    ev_k3[i].record(cur)

  issued = done = 0
  while done < C:
    if issued < C and issued - done <= STREAM_LOOKAHEAD:
      enqueue(issued)
      issued += 1
      continue
    ev_meta[done].synchronize()
    if int(flags[done, _lib.META_OVERFLOW]) != 0:
      # table (or mapping arrays) too small: chunks >= done are untouched on the device (K2 / K3 are
      # guarded by the overflow flag) -- move everything into a bigger table and redo from `done`
      cur.synchronize()
      table = _grow_rehash(table, dtype)
      old = rb
      rb = _RankBuffers(table, dtype)
      rb.keys[: old.keys.numel()].copy_(old.keys)
      rb.ids[: old.ids.numel()].copy_(old.ids)
      issued = done
      continue
    if in_place_eff:
      a, b = bounds[done]
      down.submit(a, b, dev_buf[a:b], ev_k3[done])
    done += 1

  cur.synchronize()
  meta = table.read_meta()
  nl = int(meta[_lib.META_ASSIGNED])
  res = RenumberResult()
  res.keys, res.ids, res.n_labels, res.table = rb.keys, rb.ids, nl, table
  res.out_dtype = out_dtype = _out_dtype(dtype, start, nl)
  mapping = _mapping_to_dict(res, dtype, preserve_zero)
  if out_dtype == dtype and in_place_eff:
    down.finish()
    return arr, mapping
  out_buf = dev_buf if out_dtype == dtype else _cast(dev_buf, dtype, out_dtype, n)
  out = D.host_empty(arr.shape, out_dtype, order)  # page-locked
  out_u8 = D.as_torch_bytes(D.flat_view(out))
  s_side = torch.cuda.Stream() if down is not None and not down.direct else s_out  # do not queue behind the staged pieces
  s_side.wait_stream(cur)
  with torch.cuda.stream(s_side):
    out_u8.copy_(out_buf[: out_u8.numel()], non_blocking=True)
  if down is not None:
    down.finish()
  s_side.synchronize()
  return out, mapping


def _mapping_to_dict(res: RenumberResult, dtype, preserve_zero) -> dict:
  if res.keys_host is not None:
    keys, ids = res.keys_host.tolist(), res.ids_host.tolist()
  else:
    keys = D.download(res.keys, res.n_labels, dtype).tolist()
    ids = D.download(res.ids, res.n_labels, dtype).tolist()
  mapping = {0: 0} if preserve_zero else {}
  mapping.update(zip(keys, ids))
  return mapping


def renumber(arr, start=1, preserve_zero=True, in_place=False, *, max_labels=None):
  """
  Renumber the labels of `arr` in order of first appearance in memory, starting from `start`
  (0 stays 0 when preserve_zero), and refit to the narrowest dtype that holds the next id.
  Returns (renumbered array, {old: new}).  Reference: fastremap.pyx:121-161 and :196-257.
  """
  is_t = isinstance(arr, torch.Tensor)
  if not is_t:
    arr = np.asarray(arr)
  size = arr.numel() if is_t else arr.size
  if size == 0:
    return arr, {}
  adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  if adtype == np.dtype("bool") and preserve_zero:
    return arr, {0: 0, 1: start}
  elif adtype == np.dtype("bool"):
    arr = arr.view(torch.uint8) if is_t else arr.view(np.uint8)
    adtype = np.dtype("u1")
  _require_int(adtype, "renumber")

  if is_t:
    v = D.ingest(arr, alias_ok=bool(in_place))
    in_place_eff = v.aliases_input
  else:
    in_place_eff = bool(in_place) and (arr.flags["F_CONTIGUOUS"] or arr.flags["C_CONTIGUOUS"]) and arr.flags.writeable
    if arr.nbytes >= STREAM_MIN_BYTES:
      if not (arr.flags["F_CONTIGUOUS"] or arr.flags["C_CONTIGUOUS"]):
        arr = np.copy(arr, order="C")
      return _renumber_streamed(arr, start, preserve_zero, in_place_eff, max_labels)
    v = D.ingest(arr)
  res = _renumber_small(v.buf, v.n, v.dtype, start, preserve_zero, in_place_eff) if v.n <= SMALL_MAX_VOXELS else None
  if res is None:
    res = renumber_device(v.buf, v.n, v.dtype, start=start, preserve_zero=preserve_zero, write_wide=in_place_eff,
                          max_labels=max_labels, row_bytes=v.row_bytes)
  mapping = _mapping_to_dict(res, v.dtype, preserve_zero)

  if is_t:
    if res.out_dtype == v.dtype:
      out = arr if in_place_eff else D.emit(res.out, v.dtype, v.shape, v.order, True)
    else:
      out = D.emit(res.out, res.out_dtype, v.shape, v.order, True)
    return out, mapping

  if in_place_eff:
    # the caller's buffer receives the ids at the old width (what the reference leaves behind)
    D.download_into(res.wide, v.nbytes, D.flat_view(v.host))
    if res.out_dtype == v.dtype:
      return v.host, mapping
  return D.emit(res.out, res.out_dtype, v.shape, v.order, False), mapping


# ------------------------------------------------------------------------------------- remap family


def _check_missing(table: LabelTable, out_buf, dtype, message: str):
  meta = table.read_meta()
  idx = int(meta[_lib.META_MISSING_INDEX])
  if idx != _lib.EMPTY:
    w = np.dtype(dtype).itemsize
    label = D.download(out_buf[idx * w:(idx + 1) * w], 1, dtype).tolist()[0]  # a missing label is left as it was
    raise KeyError(message.format(label))


def _dict_to_arrays(table: dict, dtype):
  """dict -> typed key / value arrays; OverflowError if a key or value does not fit the dtype."""
  dtype = np.dtype(dtype)
  n = len(table)
  try:
    keys = np.fromiter(table.keys(), dtype=dtype, count=n)
    vals = np.fromiter(table.values(), dtype=dtype, count=n)
  except (OverflowError, ValueError, TypeError):
    keys = np.array([_to_bits(k, dtype) for k in table.keys()], dtype=np.uint64).astype("u%d" % dtype.itemsize).view(dtype)
    vals = np.array([_to_bits(x, dtype) for x in table.values()], dtype=np.uint64).astype("u%d" % dtype.itemsize).view(dtype)
  return keys, vals


def _labels_to_array(labels, dtype):
  """list of Python ints -> typed array; OverflowError if one does not fit the dtype (what Cython's
  integer conversion raises in the reference, fastremap.pyx:505-506)."""
  dtype = np.dtype(dtype)
  try:
    return np.fromiter(labels, dtype=dtype, count=len(labels))  # numpy >= 2 refuses out-of-range Python ints
  except (OverflowError, ValueError, TypeError):
    return np.array([_to_bits(x, dtype) for x in labels], dtype=np.uint64).astype("u%d" % dtype.itemsize).view(dtype)


def remap(arr, table, preserve_missing_labels=False, in_place=False):
  """
  Remap the values of `arr` through the dict `table`.  A value missing from `table` is left alone
  (preserve_missing_labels=True) or raises KeyError.  Reference: fastremap.pyx:654-703 and :705-769.
  After a KeyError the contents of an in-place CUDA tensor are unspecified (the reference leaves a
  partially rewritten prefix); a numpy input is left untouched.
  """
  if type(arr) == list:
    arr = np.array(arr)
  is_t = isinstance(arr, torch.Tensor)
  if not is_t:
    arr = np.asarray(arr)
  adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  size = arr.numel() if is_t else arr.size

  original_dtype = adtype
  if len(table):
    min_label, max_label = min(table.values()), max(table.values())
    fit_value = min_label if abs(min_label) > abs(max_label) else max_label
    arr = refit(arr, fit_value, increase_only=True)
    adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  widened = adtype != original_dtype

  if is_t:
    writeable, contiguous = True, None  # decided by ingest
  else:
    writeable = arr.flags.writeable
    contiguous = arr.flags.f_contiguous or arr.flags.c_contiguous

  identity = all([k == v for k, v in table.items()]) and preserve_missing_labels
  if identity or size == 0:
    if is_t:
      return arr if (in_place or widened) else arr.clone()
    make_copy = ((not in_place) or (not writeable) or (not contiguous)) and not widened
    return np.copy(arr, order="F" if arr.flags["F_CONTIGUOUS"] else "C") if make_copy else arr

  _require_int(adtype, "remap")
  keys, vals = _dict_to_arrays(table, adtype)  # OverflowError before any device work

  if is_t:
    v = D.ingest(arr, alias_ok=bool(in_place) or widened)
    in_place_eff = v.aliases_input
  else:
    in_place_eff = (bool(in_place) and writeable and contiguous) or widened
    if arr.nbytes >= STREAM_MIN_BYTES and not (preserve_missing_labels and len(table) == 1):
      if not contiguous:
        arr = np.copy(arr, order="C")
      D.dev()
      kd, vd = D.upload_array(keys), D.upload_array(vals)
      t = table_from_arrays(kd, vd, len(keys), adtype, op=_lib.OP_SET)
      policy = _lib.MISS_KEEP if preserve_missing_labels else _lib.MISS_ERROR
      return _relabel_host_streamed(arr, in_place_eff and arr.flags.writeable, t, policy, 0, len(keys),
                                    "{} was not in the remap table.")
    v = D.ingest(arr)

  if preserve_missing_labels and len(table) == 1:
    before, after = int(keys.view("u%d" % adtype.itemsize)[0]), int(vals.view("u%d" % adtype.itemsize)[0])
    _lib.check(_L().frb_replace_one(D.ptr(v.buf), D.ptr(v.buf), v.n, D.code(adtype), before, after, D.stream()),
               "frb_replace_one")
  else:
    kd, vd = D.upload_array(keys), D.upload_array(vals)
    t = table_from_arrays(kd, vd, len(keys), adtype, op=_lib.OP_SET)
    policy = _lib.MISS_KEEP if preserve_missing_labels else _lib.MISS_ERROR
    _relabel(v.buf, v.n, adtype, t, policy, dst_wide=v.buf, batch_rows=_batch_rows(len(keys)), plane_bytes=v.plane_bytes)
    if not preserve_missing_labels:
      _check_missing(t, v.buf, adtype, "{} was not in the remap table.")

  if is_t:
    return arr if v.aliases_input else D.emit(v.buf, adtype, v.shape, v.order, True)
  if in_place_eff:
    D.download_into(v.buf, v.nbytes, D.flat_view(v.host))
    return v.host
  return D.emit(v.buf, adtype, v.shape, v.order, False)


def _relabel_host_streamed(arr: np.ndarray, in_place_eff: bool, table: LabelTable, policy, fill_bits, n_labels, message):
  """
  Lookup pass (remap / mask / mask_except / remap_from_array_kv) over a large HOST array as a chunk
  pipeline: voxels are independent, so chunk c is relabelled while chunk c+1 is on its way to the
  device and chunk c-1 on its way back (three streams; PCIe busy in both directions).
  With the error policy nothing is written back until the whole array has been checked, so a
  KeyError leaves the caller's array untouched.
  Returns the result array (arr itself when in_place_eff).
  """
  dtype = arr.dtype
  w = dtype.itemsize
  n = arr.size
  nbytes = n * w
  order = "F" if arr.flags["F_CONTIGUOUS"] else "C"
  src_u8 = D.as_torch_bytes(D.flat_view(arr))
  out = arr if in_place_eff else D.host_empty(arr.shape, dtype, order)
  out_u8 = D.as_torch_bytes(D.flat_view(out))
  chunk = max(STREAM_CHUNK_BYTES // 16 * 16, 16)
  bounds = [(a, min(a + chunk, nbytes)) for a in range(0, nbytes, chunk)]
  cur = torch.cuda.current_stream()
  s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
  dev_buf = D.alloc(nbytes)
  s_in.wait_stream(cur)
  s_out.wait_stream(cur)
  overlap_d2h = policy != _lib.MISS_ERROR
  up = D.ChunkUploader(src_u8, dev_buf, bounds, s_in)  # pageable arrays go through a staging ring
  down = D.ChunkDownloader(out_u8, s_out)
  try:
    _relabel_host_streamed_run(bounds, dev_buf, nbytes, w, dtype, table, policy, fill_bits, n_labels, message, overlap_d2h, cur, up, down)
  finally:
    up.abort()
    down.abort()
  return out


def _relabel_host_streamed_run(bounds, dev_buf, nbytes, w, dtype, table, policy, fill_bits, n_labels, message, overlap_d2h, cur, up, down):
  batch = _batch_rows(n_labels)
  for i, (a, b) in enumerate(bounds):
    cur.wait_event(up.event(i))
    view = dev_buf[a:b]
    # the first missing index is relative to the chunk: chunks are visited in memory order and
    # atomicMin keeps the smallest, so a later chunk may only lower it if no earlier chunk set it
    if policy == _lib.MISS_ERROR:
      _relabel_chunk_error(view, (b - a) // w, dtype, table, a // w, batch)
    else:
      _relabel(view, (b - a) // w, dtype, table, policy, fill_bits=fill_bits, dst_wide=view, batch_rows=batch)
    if overlap_d2h:
      e = torch.cuda.Event()
      e.record(cur)
      down.submit(a, b, view, e)
  if not overlap_d2h:
    cur.synchronize()
    _check_missing(table, dev_buf, dtype, message)
    e = torch.cuda.Event()
    e.record(cur)
    down.submit(0, nbytes, dev_buf[:nbytes], e)
  down.finish()
  cur.synchronize()


def _relabel_chunk_error(view, ne, dtype, table: LabelTable, elem_offset, batch):
  """Error-policy lookup of one chunk: the table's MISSING_INDEX word must end up as the GLOBAL first
  missing index, so the chunk-relative index the kernel records is rebased between launches."""
  before = table.meta[_lib.META_MISSING_INDEX].clone()
  table.meta[_lib.META_MISSING_INDEX] = -1  # all-ones: none
  _relabel(view, ne, dtype, table, _lib.MISS_ERROR, dst_wide=view, batch_rows=batch)
  local = table.meta[_lib.META_MISSING_INDEX]
  # earlier chunks win (smaller global index); all-ones (-1 as int64) means none
  rebased = torch.where(local == -1, local, local + elem_offset)
  table.meta[_lib.META_MISSING_INDEX] = torch.where(before == -1, rebased, before)


def mask(arr, labels, in_place=False, value=0):
  """Set every voxel whose label is in `labels` to `value` (reference: fastremap.pyx:438-458)."""
  labels = {lbl: value for lbl in labels}
  return remap(arr, labels, preserve_missing_labels=True, in_place=in_place)


def mask_except(arr, labels, in_place=False, value=0):
  """Set every voxel whose label is NOT in `labels` to `value` (reference: fastremap.pyx:460-490, :492-527)."""
  if not isinstance(labels, list):
    raise TypeError("Argument 'labels' has incorrect type (expected list, got %s)" % type(labels).__name__)
  is_t = isinstance(arr, torch.Tensor)
  adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  if not D.is_int_dtype(adtype):
    raise TypeError("No matching signature found")
  lab = _labels_to_array(labels, adtype)
  fill = _to_bits(value, adtype)

  if is_t:
    v = D.ingest(arr, alias_ok=bool(in_place))
    in_place_eff = v.aliases_input
  else:
    contiguous = arr.flags.f_contiguous or arr.flags.c_contiguous
    in_place_eff = bool(in_place) and arr.flags.writeable and contiguous
    if arr.nbytes >= STREAM_MIN_BYTES:
      if not contiguous:
        arr = np.copy(arr, order="C")
      D.dev()
      kd = D.upload_array(lab)
      t = table_from_arrays(kd, kd, len(lab), adtype, op=_lib.OP_SET)
      return _relabel_host_streamed(arr, in_place_eff, t, _lib.MISS_FILL, fill, len(lab), "")
    v = D.ingest(arr)
  if v.n:
    kd = D.upload_array(lab)
    t = table_from_arrays(kd, kd, len(lab), adtype, op=_lib.OP_SET)
    _relabel(v.buf, v.n, adtype, t, _lib.MISS_FILL, fill_bits=fill, dst_wide=v.buf, batch_rows=_batch_rows(len(lab)),
             plane_bytes=v.plane_bytes)
  if is_t:
    return arr if v.aliases_input else D.emit(v.buf, adtype, v.shape, v.order, True)
  if in_place_eff:
    D.download_into(v.buf, v.nbytes, D.flat_view(v.host))
    return v.host
  return D.emit(v.buf, adtype, v.shape, v.order, False)


def _check_1d_same_dtype(arrs, names):
  dts = []
  for a, nm in zip(arrs, names):
    nd = a.dim() if isinstance(a, torch.Tensor) else np.asarray(a).ndim
    if nd != 1:
      raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % nd)
    dts.append(D.numpy_dtype(a.dtype) if isinstance(a, torch.Tensor) else np.asarray(a).dtype)
  if any(d != dts[0] for d in dts):
    raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (dts[0], [d for d in dts if d != dts[0]][0]))
  return dts[0]


def _finish_1d(arr, v, in_place_eff):
  if v.is_torch:
    return arr if v.aliases_input else D.emit(v.buf, v.dtype, v.shape, v.order, True)
  if in_place_eff:
    D.download_into(v.buf, v.nbytes, D.flat_view(v.host))
    return v.host
  return D.emit(v.buf, v.dtype, v.shape, v.order, False)


def remap_from_array(arr, vals, in_place=True):
  """arr[i] = vals[arr[i]] where arr[i] < len(vals), else unchanged.  1-D, unsigned, same dtype
  (reference: fastremap.pyx:771-791; an empty `vals` is a no-op here, it is an out-of-bounds read there)."""
  dt = _check_1d_same_dtype([arr, vals], ["arr", "vals"])
  if not np.issubdtype(dt, np.unsignedinteger):
    raise TypeError("No matching signature found")
  is_t = isinstance(arr, torch.Tensor)
  if is_t:
    v = D.ingest(arr, alias_ok=bool(in_place))
    in_place_eff = v.aliases_input
  else:
    in_place_eff = bool(in_place) and arr.flags.writeable and arr.flags.c_contiguous
    v = D.ingest(arr)
  vv = D.ingest(vals, alias_ok=True)
  if v.n and vv.n:
    _lib.check(_L().frb_remap_from_array(D.ptr(v.buf), D.ptr(v.buf), v.n, D.ptr(vv.buf), vv.n, D.code(dt), D.stream()),
               "frb_remap_from_array")
  return _finish_1d(arr, v, in_place_eff)


def remap_from_array_kv(arr, keys, vals, preserve_missing_labels=True, in_place=True):
  """Remap through parallel key / value arrays (later duplicate keys win).  1-D, one integer dtype
  (reference: fastremap.pyx:793-827)."""
  dt = _check_1d_same_dtype([arr, keys, vals], ["arr", "keys", "vals"])
  if not D.is_int_dtype(dt):
    raise TypeError("No matching signature found")
  nk = keys.numel() if isinstance(keys, torch.Tensor) else keys.size
  nv = vals.numel() if isinstance(vals, torch.Tensor) else vals.size
  assert nk == nv
  is_t = isinstance(arr, torch.Tensor)
  if is_t:
    v = D.ingest(arr, alias_ok=bool(in_place))
    in_place_eff = v.aliases_input
  else:
    in_place_eff = bool(in_place) and arr.flags.writeable and arr.flags.c_contiguous
    if arr.nbytes >= STREAM_MIN_BYTES:
      if not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr)
      kv, vv = D.ingest(keys, alias_ok=True), D.ingest(vals, alias_ok=True)
      t = table_from_arrays(kv.buf, vv.buf, nk, dt, op=_lib.OP_MAX, positions=True) if nk else LabelTable(dt, 1, 0)
      policy = _lib.MISS_KEEP if preserve_missing_labels else _lib.MISS_ERROR
      return _relabel_host_streamed(arr, in_place_eff, t, policy, 0, nk, "{} was not in the remap keys.")
    v = D.ingest(arr)
  if v.n:
    kv, vv = D.ingest(keys, alias_ok=True), D.ingest(vals, alias_ok=True)
    # positions + max => the last duplicate of a key supplies its value (fastremap.pyx:810-811)
    t = table_from_arrays(kv.buf, vv.buf, nk, dt, op=_lib.OP_MAX, positions=True) if nk else LabelTable(dt, 1, 0)
    policy = _lib.MISS_KEEP if preserve_missing_labels else _lib.MISS_ERROR
    _relabel(v.buf, v.n, dt, t, policy, dst_wide=v.buf, batch_rows=_batch_rows(nk))
    if not preserve_missing_labels:
      _check_missing(t, v.buf, dt, "{} was not in the remap keys.")
  return _finish_1d(arr, v, in_place_eff)


# ------------------------------------------------------------------------------------- component_map


def component_map_arrays(component_labels, parent_labels, max_labels=None):
  """Device part of component_map: returns (keys, parents) as numpy arrays (unordered)."""
  is_t = isinstance(component_labels, torch.Tensor)
  if is_t:
    vc = D.ingest(component_labels, alias_ok=True)
    vp = D.ingest(parent_labels, alias_ok=True)
    if vc.order != vp.order:
      vp = D.ingest(parent_labels.contiguous() if vc.order == "C" else parent_labels, alias_ok=True)
      if vc.order != vp.order:
        raise FastremapB200Error("component_map: tensors must share one memory order")
  else:
    # both arrays take the memory order of the first (fastremap.pyx:95-102)
    if component_labels.flags.c_contiguous:
      component_labels, parent_labels = np.ascontiguousarray(component_labels), np.ascontiguousarray(parent_labels)
    else:
      component_labels, parent_labels = np.asfortranarray(component_labels), np.asfortranarray(parent_labels)
    vc, vp = D.ingest(component_labels), D.ingest(parent_labels)
  _require_int(vc.dtype, "component_map")
  _require_int(vp.dtype, "component_map")
  keys, pars, nl = component_map_device(vc.buf, vc.n, vc.dtype, vp.buf, vp.dtype, max_labels=max_labels, row_bytes=vc.row_bytes,
                                        plane_bytes=vc.plane_bytes)
  return D.download(keys, nl, vc.dtype), D.download(pars, nl, vp.dtype)


def component_map_device(cc_buf, n, cc_dtype, parent_buf, parent_dtype, max_labels=None, row_bytes=0, plane_bytes=0, prev_label_bits=None):
  """K7 on flat device buffers: returns (keys_buf, parents_buf, n_labels), unordered.
  prev_label_bits (sharded volumes): the label the previous slab ended with -- a first voxel that carries it continues a
  run and does not (re)assign its label (fastremap.pyx:581)."""
  cc_dtype, parent_dtype = np.dtype(cc_dtype), np.dtype(parent_dtype)
  L = _L()
  table = LabelTable(cc_dtype, _default_entries(n, max_labels), 0)
  # the build lists the slots it fills, so that the gather walks 10^7 entries instead of scanning (and compacting) a
  # table of 3 * 10^7 mostly empty slots: 0.58 -> 0.15 ms at 10^7 labels.  Not for sharded volumes: frb_component_unhead
  # may withdraw a label, and the list has no way to skip it.
  listed = prev_label_bits is None and not table.direct and table.cap >= (1 << 20)
  lists = {}

  def build(t):
    cap_l = 2 * t.cap if listed else 0   # uint32 words: per-warp regions need slack over the cap / 2 labels the table may hold
    lists["buf"], lists["cap"] = (D.alloc(cap_l * 4), cap_l) if listed else (None, 0)
    _lib.check(L.frb_component_build(D.ptr(cc_buf), n, D.code(cc_dtype), 0, D.ptr(t.slots), t.cap, D.ptr(t.meta),
                                     row_bytes, plane_bytes, D.ptr(lists["buf"]) if listed else None, cap_l, D.stream()),
               "frb_component_build")

  meta = _run_build(build, table)
  nl = int(meta[_lib.META_COUNT])
  n_listed = int(meta[_lib.META_AUX1])
  listed = listed and n_listed > 0 and n_listed + (1 if meta[_lib.META_SIDE_PRESENT] else 0) == nl
  if prev_label_bits is not None and n:
    _lib.check(L.frb_component_unhead(D.ptr(cc_buf), D.code(cc_dtype), int(prev_label_bits) & _lib.EMPTY, 0, D.ptr(table.slots), table.cap,
                                      D.ptr(table.meta), D.stream()), "frb_component_unhead")
  keys = D.alloc(nl * cc_dtype.itemsize)
  pars = D.alloc(nl * parent_dtype.itemsize)
  _lib.check(L.frb_component_gather(D.ptr(table.slots), table.cap, D.ptr(table.meta), D.code(cc_dtype), D.ptr(parent_buf),
                                    D.code(parent_dtype), 0, D.ptr(keys), D.ptr(pars), nl,
                                    D.ptr(lists["buf"]) if listed else None, n_listed if listed else 0, D.stream()),
             "frb_component_gather")
  if prev_label_bits is not None and n:
    nl = int(table.read_meta()[_lib.META_LIST_COUNT])  # one label fewer if the slab's first voxel continued a run
  return keys, pars, nl


def component_map(component_labels, parent_labels):
  """{component label: parent label}; on conflicts the last run start in memory order wins
  (reference: fastremap.pyx:529-563 and :565-587)."""
  if not isinstance(component_labels, torch.Tensor):
    if not isinstance(component_labels, np.ndarray):
      component_labels = np.array(component_labels)
    if not isinstance(parent_labels, np.ndarray):
      parent_labels = np.array(parent_labels)
    size = component_labels.size
  else:
    size = component_labels.numel()
  if size == 0:
    return {}
  if tuple(component_labels.shape) != tuple(parent_labels.shape):
    raise ValueError("The size of the inputs must match: {} vs {}".format(
      tuple(component_labels.shape), tuple(parent_labels.shape)))
  keys, pars = component_map_arrays(component_labels, parent_labels)
  return dict(zip(keys.tolist(), pars.tolist()))


def point_cloud(arr):
  """
  { label: (N, 2 or 3) uint16 array of the positions of its voxels } for the non-zero labels of a 2-D
  or 3-D integer image (reference: fastremap.pyx:1423-1540).  Voxels are visited in C order whatever
  the memory layout (the reference indexes arr[i, j, k]); inside a label the rows are in scan order
  (the reference's argsort is not stable, so its order inside a label is unspecified).
  The values are views into one array, like the reference's.
  """
  is_t = isinstance(arr, torch.Tensor)
  if not is_t:
    arr = np.asarray(arr)
  adtype = D.numpy_dtype(arr.dtype) if is_t else arr.dtype
  if adtype == np.dtype("bool"):
    arr = arr.view(torch.uint8) if is_t else arr.view(np.uint8)
    adtype = np.dtype("u1")
  if not D.is_int_dtype(adtype):
    raise TypeError("No matching signature found")
  ndim = arr.dim() if is_t else arr.ndim
  if ndim not in (2, 3):
    raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % ndim)
  shape = tuple(int(x) for x in arr.shape)
  n = int(np.prod(shape, dtype=np.int64))
  if n == 0:
    return {}
  if any(d > 65536 for d in shape):
    raise ValueError("point_cloud: coordinates are uint16; no axis may be longer than 65536")
  arr = arr.contiguous() if is_t else np.ascontiguousarray(arr)  # the scan is in C order
  v = D.ingest(arr, alias_ok=True)
  L = _L()
  c = D.code(adtype)
  meta = D.alloc_meta()
  ws_bytes = int(L.frb_nonzero_ws_bytes(n))
  ws = D.alloc(ws_bytes)
  _lib.check(L.frb_nonzero_count(D.ptr(v.buf), n, c, D.ptr(meta), D.ptr(ws), ws_bytes, D.stream()), "frb_nonzero_count")
  nfg = int(D.read_u64(meta)[_lib.META_LIST_COUNT])
  if nfg == 0:
    return {}
  labels, index = D.alloc(nfg * 8), D.alloc(nfg * 8)
  _lib.check(L.frb_nonzero_emit(D.ptr(v.buf), n, c, D.ptr(labels), D.ptr(index), nfg, D.ptr(ws), D.stream()), "frb_nonzero_emit")
  sws_bytes = int(L.frb_sort_pairs_u64_ws_bytes(nfg))
  sws = D.alloc(sws_bytes)
  _lib.check(L.frb_sort_pairs_u64(D.ptr(labels), D.ptr(index), nfg, 8 * adtype.itemsize, D.ptr(sws), sws_bytes, D.stream()),
             "frb_sort_pairs_u64")
  coords = D.alloc(nfg * ndim * 2)
  d1, d2 = (shape[1], 1) if ndim == 2 else (shape[1], shape[2])
  _lib.check(L.frb_unravel_u16(D.ptr(index), nfg, ndim, d1, d2, D.ptr(coords), D.stream()), "frb_unravel_u16")
  # label boundaries: the sorted labels and their counts (ascending in the dtype's order, like the sort above)
  uniq, cts, nu = unique_counts_device(v.buf, n, adtype)
  u_np = D.download(uniq, nu, adtype)
  c_np = D.download(cts, nu, np.uint64)
  keep = u_np != 0
  u_np, c_np = u_np[keep], c_np[keep]
  ends = np.cumsum(c_np.astype(np.int64))
  starts = ends - c_np.astype(np.int64)
  if is_t:
    cloud = coords[: nfg * ndim * 2].view(torch.uint16).reshape(nfg, ndim)
  else:
    cloud = D.download(coords, nfg * ndim, np.uint16).reshape(nfg, ndim)
  return {int(lbl): cloud[a:b] for lbl, a, b in zip(u_np.tolist(), starts.tolist(), ends.tolist())}


def inverse_component_map(parent_labels, component_labels):
  """
  { parent label: [component labels ...] } for every (parent, component) pair that occurs at the
  same voxel (reference: fastremap.pyx:589-652, which collects the pairs into Python sets; the lists
  here are sorted ascending, the reference's are in set order).
  The pair of a voxel is packed into one 64-bit label and the distinct pairs are a `unique` of the
  packed volume; labels wider than 4 bytes are renumbered to dense 32-bit ids first.
  """
  is_t = isinstance(component_labels, torch.Tensor)
  if not is_t:
    if not isinstance(component_labels, np.ndarray):
      component_labels = np.array(component_labels)
    if not isinstance(parent_labels, np.ndarray):
      parent_labels = np.array(parent_labels)
    size = component_labels.size
  else:
    size = component_labels.numel()
  if size == 0:
    return {}
  if tuple(component_labels.shape) != tuple(parent_labels.shape):
    raise ValueError("The size of the inputs must match: {} vs {}".format(
      tuple(component_labels.shape), tuple(parent_labels.shape)))
  if is_t:
    vc = D.ingest(component_labels)
    vp = D.ingest(parent_labels)
    if vc.order != vp.order:
      vp = D.ingest(parent_labels.contiguous() if vc.order == "C" else parent_labels)
      if vc.order != vp.order:
        raise FastremapB200Error("inverse_component_map: tensors must share one memory order")
  else:
    if component_labels.flags.c_contiguous:  # both arrays take the memory order of the first (fastremap.pyx:95-102)
      component_labels, parent_labels = np.ascontiguousarray(component_labels), np.ascontiguousarray(parent_labels)
    else:
      component_labels, parent_labels = np.asfortranarray(component_labels), np.asfortranarray(parent_labels)
    vc, vp = D.ingest(component_labels), D.ingest(parent_labels)
  _require_int(vc.dtype, "inverse_component_map")
  _require_int(vp.dtype, "inverse_component_map")
  n = vc.n

  def dense32(v):
    """(buffer of <= 4-byte raw labels, width, decode) with decode: raw value -> original Python int"""
    if v.dtype.itemsize <= 4:
      return v.buf, v.dtype.itemsize, None
    # 8-byte labels: dense ids 0 .. L-1 in encounter order (fewer than 2^32 of them in any array this size)
    res = renumber_device(v.buf, n, v.dtype, start=0, preserve_zero=False, write_wide=False)
    if res.out_dtype.itemsize > 4:
      raise FastremapB200Error("inverse_component_map: more than 2^32 distinct labels")
    keys = D.download(res.keys, res.n_labels, v.dtype)
    return res.out, res.out_dtype.itemsize, keys

  pbuf, wp, pkeys = dense32(vp)
  cbuf, wc, ckeys = dense32(vc)
  packed = D.alloc(n * 8)
  _lib.check(_L().frb_pack_pairs(D.ptr(pbuf), wp, D.ptr(cbuf), wc, D.ptr(packed), n, D.stream()), "frb_pack_pairs")
  uniq, _, nu = unique_counts_device(packed, n, np.uint64) if n > 1 else (packed, None, 1)
  pairs = D.download(uniq, nu, np.uint64)
  par = pairs >> np.uint64(8 * wc)
  com = pairs & np.uint64((1 << (8 * wc)) - 1)

  def decode(raw, keys, dtype):
    if keys is not None:
      return keys[raw.astype(np.int64)]
    return raw.astype("u%d" % dtype.itemsize).view(dtype) if dtype.itemsize < 8 else raw.view(dtype)

  par, com = decode(par, pkeys, vp.dtype), decode(com, ckeys, vc.dtype)
  order = np.lexsort((com, par))
  par, com = par[order], com[order]
  heads = np.flatnonzero(np.concatenate(([True], par[1:] != par[:-1])))
  ends = np.append(heads[1:], par.size)
  return {int(par[a]): com[a:b].tolist() for a, b in zip(heads, ends)}


# ------------------------------------------------------------------------------------- unique

# strategy thresholds (they change speed only, never results -- SURVEY App. A-11)
DENSE_MAX_BINS = 1 << 27        # 1 GiB of uint64 bins
PLANE_WALK_MIN_VOXELS_PER_LABEL = 4096   # plane walk while the previous call on this shape saw more than n / 4096 labels
_unique_density = {}            # (voxels, dtype, plane bytes) -> distinct labels of the previous call
HASH_MAX_FRACTION = 8           # hash strategy while distinct labels <= n / 8, else sort + RLE


def unique_counts_device(buf, n, dtype, strategy=None, max_labels=None, row_bytes=0, plane_bytes=0):
  """
  Sorted distinct labels and their counts of a flat device buffer.
  Returns (uniq_buf, counts_buf, n_unique); buffers are flat uint8 CUDA tensors (labels in `dtype`,
  counts uint64).  strategy in {None, "dense", "hash", "sort"}.
  """
  dtype = np.dtype(dtype)
  w = dtype.itemsize
  c = D.code(dtype)
  L = _L()
  if strategy is None:
    if n <= SMALL_MAX_VOXELS:
      small = _unique_small(buf, n, dtype)
      if small is not None:
        return small
    # The per-warp aggregating cache in front of the table makes the hash strategy as fast as the
    # dense histogram, and it needs no min/max pre-pass over the volume.
    strategy = "hash"
  if strategy == "dense":
    mn, mx, _, _ = _scan(buf, n, dtype)
    bins = mx - mn + 1
    if bins > DENSE_MAX_BINS:
      raise ValueError("unique: dense strategy needs max - min < %d" % DENSE_MAX_BINS)
    min_bits = _to_bits(mn, dtype)
    counts = torch.zeros(bins, dtype=torch.int64, device=D.dev())
    _lib.check(L.frb_hist_dense(D.ptr(buf), n, c, min_bits, D.ptr(counts), bins, row_bytes, D.stream()), "frb_hist_dense")
    meta = D.alloc_meta()
    ws_bytes = int(L.frb_compact_bins_ws_bytes(bins))
    ws = D.alloc(ws_bytes)
    _lib.check(L.frb_compact_bins_count(D.ptr(counts), bins, D.ptr(meta), D.ptr(ws), ws_bytes, D.stream()),
               "frb_compact_bins_count")
    nu = int(D.read_u64(meta)[_lib.META_LIST_COUNT])
    uniq, cts = D.alloc(nu * w), D.alloc(nu * 8)
    _lib.check(L.frb_compact_bins_emit(D.ptr(counts), bins, c, min_bits, D.ptr(uniq), D.ptr(cts), nu, D.ptr(ws), D.stream()),
               "frb_compact_bins_emit")
    return uniq, cts, nu
  if strategy == "hash":
    table = LabelTable(dtype, _default_entries(n, max_labels, voxels_per_label=512), 0)
    if table.direct:
      # 8- / 16-bit labels: the slot index is the label, so the table needs no log, no export and no sort -- count, then
      # an ordered compaction of its 2^8 / 2^16 slots (4 launches instead of 9: uint8 1024^3 0.34 -> 0.26 ms)
      _lib.check(L.frb_hist_hash(D.ptr(buf), n, c, D.ptr(table.slots), table.cap, D.ptr(table.meta), row_bytes, plane_bytes,
                                 None, 0, D.stream()), "frb_hist_hash")
      uniq, cts, ws = D.alloc(table.cap * w), D.alloc(table.cap * 8), D.alloc(256)
      _lib.check(L.frb_direct_export_sorted(D.ptr(table.slots), table.cap, c, D.ptr(table.meta), D.ptr(uniq), D.ptr(cts), table.cap,
                                            D.ptr(ws), 256, D.stream()), "frb_direct_export_sorted")
      return uniq, cts, int(table.read_meta()[_lib.META_LIST_COUNT])
    limit = max(n // HASH_MAX_FRACTION, 1 << 16)
    hws_bytes = int(L.frb_hist_ws_bytes(n))
    hws = D.alloc(hws_bytes)           # per-warp eviction logs of the counting pass
    have_logs = False                  # the logs of a finished pass can be folded into a bigger table without re-reading the volume
    while True:
      # count -> export -> sort, queued back to back: the number of labels stays on the device (the sort
      # kernel reads it there) until the one host round trip at the end
      room = table.entries_max()
      k64, v64 = D.alloc(room * 8), D.alloc(room * 8)
      ws_bytes = int(L.frb_sort_pairs_ws_bytes(room))
      ws = D.alloc(ws_bytes)
      uniq, cts = D.alloc(room * w), D.alloc(room * 8)
      # plane walk or linear walk?  The plane walk pays when segments are small (many labels per plane: 1.40 vs 1.64 ms at
      # 900 k labels in 1024^3) and costs a little when they are large (0.97 vs 0.93 ms at 88 k labels).  Volumes of one
      # job look alike, so the label density of the previous call on the same shape decides; first call: plane walk.
      last = _unique_density.get((n, dtype.str, plane_bytes))
      walk_plane = plane_bytes if (last is None or last * PLANE_WALK_MIN_VOXELS_PER_LABEL > n) else 0
      if have_logs:
        _lib.check(L.frb_hist_apply_log(D.ptr(hws), hws_bytes, n, c, D.ptr(table.slots), table.cap, D.ptr(table.meta), D.stream()),
                   "frb_hist_apply_log")
      else:
        _lib.check(L.frb_hist_hash(D.ptr(buf), n, c, D.ptr(table.slots), table.cap, D.ptr(table.meta), row_bytes, walk_plane,
                                   D.ptr(hws), hws_bytes, D.stream()), "frb_hist_hash")
      _lib.check(L.frb_table_export(D.ptr(table.slots), table.cap, D.ptr(table.meta), D.ptr(k64), D.ptr(v64), room, D.stream()),
                 "frb_table_export")
      count_dev = ctypes.c_void_p(table.meta.data_ptr() + 8 * _lib.META_LIST_COUNT)
      _lib.check(L.frb_sort_table_pairs(D.ptr(k64), D.ptr(v64), room, count_dev, c, D.ptr(uniq), D.ptr(cts), D.ptr(ws), ws_bytes,
                                        D.stream()), "frb_sort_table_pairs")
      meta = table.read_meta()
      if not table.overfull(meta):
        break
      if table.cap // 2 > limit:  # label set too large to hash profitably
        return unique_counts_device(buf, n, dtype, strategy="sort")
      off = int(L.frb_hist_ws_flags_offset())
      have_logs = int(hws[off:off + 4].view(torch.int32).item()) == 0  # no warp bypassed its log: they hold the whole pass
      table.grow(int(meta[_lib.META_COUNT]))
    nu = int(meta[_lib.META_COUNT])
    if len(_unique_density) > 64:
      _unique_density.clear()
    _unique_density[(n, dtype.str, plane_bytes)] = nu
    return uniq, cts, nu
  if strategy == "sort":
    meta = D.alloc_meta()
    ws_bytes = int(L.frb_unique_sort_ws_bytes(n, c))
    ws = D.alloc(ws_bytes)
    _lib.check(L.frb_unique_sort_count(D.ptr(buf), n, c, D.ptr(meta), D.ptr(ws), ws_bytes, D.stream()), "frb_unique_sort_count")
    nu = int(D.read_u64(meta)[_lib.META_LIST_COUNT])
    uniq, cts = D.alloc(nu * w), D.alloc(nu * 8)
    _lib.check(L.frb_unique_sort_emit(n, c, D.ptr(uniq), D.ptr(cts), nu, D.ptr(ws), D.stream()), "frb_unique_sort_emit")
    return uniq, cts, nu
  raise ValueError("unknown unique strategy %r" % (strategy,))


def _first_indices_sorted(buf, n, dtype, nu):
  """First flat index (memory order) of every label, ordered like the sorted labels: K1 with no
  label skipped, export, sort by label.  Returns a uint64 device buffer of nu entries."""
  L = _L()
  dtype = np.dtype(dtype)
  table = LabelTable(dtype, nu, _lib.EMPTY)

  def build(t):
    _lib.check(L.frb_renumber_build(D.ptr(buf), n, D.code(dtype), 0, 0, D.ptr(t.slots), t.cap, D.ptr(t.meta), 0, D.stream()),
               "frb_renumber_build")

  meta = _run_build(build, table)
  cnt = int(meta[_lib.META_COUNT])
  if cnt != nu:
    raise FastremapB200Error("unique: first-index pass found %d labels, counting pass %d" % (cnt, nu))
  k64, v64 = D.alloc(nu * 8), D.alloc(nu * 8)
  _lib.check(L.frb_table_export(D.ptr(table.slots), table.cap, D.ptr(table.meta), D.ptr(k64), D.ptr(v64), nu, D.stream()),
             "frb_table_export")
  ws_bytes = int(L.frb_sort_pairs_ws_bytes(nu))
  ws = D.alloc(ws_bytes)
  uniq, idx = D.alloc(nu * dtype.itemsize), D.alloc(nu * 8)
  _lib.check(L.frb_sort_table_pairs(D.ptr(k64), D.ptr(v64), nu, None, D.code(dtype), D.ptr(uniq), D.ptr(idx), D.ptr(ws), ws_bytes,
                                    D.stream()), "frb_sort_table_pairs")
  return idx  # values still carry FRB_RANK_TAG: the caller masks it


def _inverse_positions(buf, n, dtype, uniq_buf, nu):
  """For every voxel the position of its label in the sorted label list, as a uint64 device buffer:
  a lookup table label -> position and one relabel pass that stores 8-byte results."""
  dtype = np.dtype(dtype)
  w = dtype.itemsize
  t = table_from_arrays(uniq_buf, None, nu, dtype, op=_lib.OP_SET, positions=True, resolve=False)
  out = D.alloc(n * 8)
  if w == 8:
    _relabel(buf, n, dtype, t, _lib.MISS_ERROR, dst_wide=out, batch_rows=_batch_rows(nu))
  else:  # the widening store is fused into the lookup pass (a separate cast cost as much as the lookup)
    _relabel(buf, n, dtype, t, _lib.MISS_ERROR, dst_narrow=out, narrow_width=8)
  return out


def _two_axis_unique(labels):
  """np.unique(labels, axis=0) for a C-contiguous (N, 2) array of < 8-byte integers, by packing every
  row into one integer of twice the width (reference: fastremap.pyx:967-982)."""
  from .dtypes import widen_dtype
  is_t = isinstance(labels, torch.Tensor)
  dtype = D.numpy_dtype(labels.dtype) if is_t else labels.dtype
  wide = np.dtype(widen_dtype(dtype))
  nrows = int(labels.shape[0])
  if nrows == 0:
    return labels
  v = D.ingest(labels, alias_ok=True)
  L = _L()
  packed = D.alloc(nrows * wide.itemsize)
  # row (a, b) read as one little-endian wide integer has b in its high half; the reference sorts
  # with a in the high half (it swaps the columns first, :976)
  _lib.check(L.frb_swap_halves(D.ptr(v.buf), D.ptr(packed), nrows, wide.itemsize, D.stream()), "frb_swap_halves")
  if nrows == 1:
    uniq, nu = packed, 1
  else:
    uniq, _, nu = unique_counts_device(packed, nrows, wide)
  _lib.check(L.frb_swap_halves(D.ptr(uniq), D.ptr(uniq), nu, wide.itemsize, D.stream()), "frb_swap_halves")
  if is_t:
    return uniq[: nu * wide.itemsize].view(D.torch_dtype(dtype)).reshape(nu, 2)
  # the reference's result is the fancy-indexed labels[:, [1, 0]], which numpy lays out column-major
  return np.asfortranarray(D.download(uniq, nu * 2, dtype).reshape((nu, 2)))


def unique(labels, return_index=False, return_inverse=False, return_counts=False, axis=None, *, strategy=None):
  """
  Sorted unique labels, optionally with first index, inverse and counts, in numpy's order
  (uniq, index, inverse, counts).  Reference: fastremap.pyx:855-965.

  The reference picks one of four strategies from the label range and the fraction of equal
  neighbours (:941-952); that choice is visible in the dtypes of the optional outputs and, for
  Fortran-ordered input, in which occurrence counts as "first" -- so the same rule is evaluated
  here (one fused min / max / pixel-pairs pass) whenever index or inverse are requested.  Where the
  reference defers to numpy (floats, axis other than the (N,2) edge-list case) this build raises:
  there is no CPU fallback.
  """
  is_t = isinstance(labels, torch.Tensor)
  if not is_t and not isinstance(labels, np.ndarray):
    labels = np.array(labels)
  adtype = D.numpy_dtype(labels.dtype) if is_t else labels.dtype
  is_int = D.is_int_dtype(adtype)
  if axis is not None or not is_int:
    c_contig = labels.is_contiguous() if is_t else labels.flags.c_contiguous
    if (axis == 0 and labels.ndim == 2 and labels.shape[1] == 2 and adtype.itemsize < 8 and is_int
        and not (return_index or return_inverse or return_counts) and c_contig):
      return _two_axis_unique(labels)
    if not is_int:
      _require_int(adtype, "unique")
    raise NotImplementedError("fastremap_b200.unique: axis=%r is handled by numpy in the reference "
                              "(fastremap.pyx:902-908); no CPU fallback here" % (axis,))
  voxels = labels.numel() if is_t else labels.size
  shape = tuple(labels.shape)

  def to_out(a_np):
    return torch.from_numpy(a_np).to(labels.device) if is_t else a_np

  if voxels == 0 or voxels == 1:
    if voxels == 0:
      uniq = np.array([], dtype=adtype)
      counts, index, inverse = np.array([], dtype=np.uint32), np.array([], dtype=np.uint64), np.array([], dtype=np.uintp)
    else:
      v = D.ingest(labels, alias_ok=True)
      uniq = D.download(v.buf, 1, adtype)
      counts, index, inverse = np.array([1], dtype=np.uint32), np.array([0], dtype=np.uint64), np.array([0], dtype=np.uintp)
    results = [to_out(uniq)]
    if return_index:
      fortran = labels.permute(tuple(reversed(range(labels.dim())))).is_contiguous() if is_t else labels.flags.f_contiguous
      if len(shape) > 1 and fortran:  # :923-929 applies to the degenerate sizes too (and makes the dtype intp)
        index = np.ravel_multi_index(np.unravel_index(index, shape, order="F"), shape, order="C")
      results.append(to_out(index))
    if return_inverse:
      results.append(to_out(inverse.reshape(shape)))
    if return_counts:
      results.append(to_out(counts))
    return tuple(results) if len(results) > 1 else results[0]

  v = D.ingest(labels, alias_ok=True)
  branch = "fast"
  if return_index or return_inverse:
    mn, mx, pairs, _ = _scan(v.buf, v.n, v.dtype)
    if mn >= 0 and mx < voxels:
      branch = "array"
    elif mx - mn <= voxels:
      branch = "array"
    elif float(pairs) / float(voxels) > 0.66:
      branch = "renumber"
    else:
      branch = "numpy"  # :947-948: numpy semantics -- first occurrence in C order, intp outputs
      if v.order == "F" and len(shape) > 1:
        v = D.ingest(labels.contiguous() if is_t else np.ascontiguousarray(labels), alias_ok=True)

  uniq, cts, nu = unique_counts_device(v.buf, v.n, v.dtype, strategy=strategy, row_bytes=v.row_bytes, plane_bytes=v.plane_bytes)
  intp = branch == "numpy"
  results = []
  if is_t:
    results.append(uniq[: nu * adtype.itemsize].view(D.torch_dtype(adtype)))
  else:
    results.append(D.download(uniq, nu, adtype))

  if return_index:
    idx_buf = _first_indices_sorted(v.buf, v.n, v.dtype, nu)
    idx = D.download(idx_buf, nu, np.uint64) & np.uint64(_lib.RANK_TAG - 1)
    if v.order == "F" and len(shape) > 1:  # :923-929: report the C-order index of that voxel
      idx = np.ravel_multi_index(np.unravel_index(idx, shape, order="F"), shape, order="C")
    elif intp:
      idx = idx.astype(np.intp)
    results.append(to_out(idx))
  if return_inverse:
    inv_buf = _inverse_positions(v.buf, v.n, v.dtype, uniq, nu)
    inv_dtype = np.dtype(np.uintp) if branch == "array" else np.dtype(np.intp)
    order = "C" if branch == "numpy" else v.order
    results.append(D.emit(inv_buf, inv_dtype, shape, order, is_t))
  if return_counts:
    cdtype = np.dtype(np.intp) if intp else np.dtype(np.uint64)
    if is_t:
      results.append(cts[: nu * 8].view(D.torch_dtype(cdtype)))
    else:
      results.append(D.download(cts, nu, cdtype))
  return tuple(results) if len(results) > 1 else results[0]
