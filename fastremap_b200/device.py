"""
Device plumbing: torch owns device memory and streams, ctypes hands raw pointers to the C ABI.

A `Vol` is a flat device buffer holding an array's elements in MEMORY order (C arrays are walked
C-wise, F arrays F-wise -- fastremap.pyx:150-157), plus what is needed to hand the result back in
the caller's own container (numpy array or CUDA torch tensor), shape and order.
"""
from __future__ import annotations

import collections
import ctypes
import os
import queue
import threading
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from ._lib import FastremapB200Error

INT_DTYPES = tuple(np.dtype(c) for c in ("u1", "u2", "u4", "u8", "i1", "i2", "i4", "i8"))
_CODE = {np.dtype("u1"): _lib.U8, np.dtype("u2"): _lib.U16, np.dtype("u4"): _lib.U32, np.dtype("u8"): _lib.U64,
         np.dtype("i1"): _lib.I8, np.dtype("i2"): _lib.I16, np.dtype("i4"): _lib.I32, np.dtype("i8"): _lib.I64}
_NP2TORCH = {np.dtype("u1"): torch.uint8, np.dtype("u2"): torch.uint16, np.dtype("u4"): torch.uint32,
             np.dtype("u8"): torch.uint64, np.dtype("i1"): torch.int8, np.dtype("i2"): torch.int16,
             np.dtype("i4"): torch.int32, np.dtype("i8"): torch.int64, np.dtype("bool"): torch.bool}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}

_bound_device = None


def code(dtype) -> int:
  try:
    return _CODE[np.dtype(dtype)]
  except KeyError:
    raise TypeError("fastremap_b200 supports integer dtypes uint8..uint64 / int8..int64, got %s" % np.dtype(dtype))


def is_int_dtype(dtype) -> bool:
  return np.dtype(dtype) in _CODE


def torch_dtype(dtype):
  return _NP2TORCH[np.dtype(dtype)]


def numpy_dtype(tdtype) -> np.dtype:
  try:
    return _TORCH2NP[tdtype]
  except KeyError:
    raise TypeError("unsupported tensor dtype %s" % tdtype)


def dev() -> torch.device:
  """Current CUDA device; binds the library's own CUDA runtime to it.  No GPU -> error (no CPU fallback)."""
  global _bound_device
  if not torch.cuda.is_available():
    raise FastremapB200Error("fastremap_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
  idx = torch.cuda.current_device()
  if _bound_device != idx:
    torch.cuda.init()
    torch.zeros(1, device="cuda:%d" % idx)  # make sure the primary context exists
    _lib.check(_lib.load().frb_set_device(idx), "frb_set_device")
    _bound_device = idx
  return torch.device("cuda", idx)


def stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_side_streams = {}


def side_stream() -> torch.cuda.Stream:
  """A per-device helper stream for small reads that must not wait for the main stream's tail."""
  idx = torch.cuda.current_device()
  if idx not in _side_streams:
    _side_streams[idx] = torch.cuda.Stream(device=idx)
  return _side_streams[idx]


def bind_host_to_gpu(index: int | None = None) -> dict:
  """
  Pin the calling process to the CPU cores of the NUMA node the GPU hangs off, so that page-locked
  buffers (first touch) and the memcpy threads live next to the PCIe root the DMA goes through.
  One process per GPU should call this before it allocates host memory.  Returns what was found
  ({"numa_node": .., "cpus": n} or {"skipped": reason}); never raises -- it is an optimisation.
  """
  try:
    import pynvml
    if index is None:
      index = torch.cuda.current_device()
    pynvml.nvmlInit()
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(visible.split(",")[index]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else index
    bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
    bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
    if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
      bus = bus[4:]
    with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
      node = int(f.read().strip())
    if node < 0:
      return {"skipped": "no NUMA information for %s" % bus}
    with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
      cpus = set()
      for part in f.read().strip().split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    allowed = os.sched_getaffinity(0) & cpus
    if not allowed:
      return {"skipped": "none of node %d's cpus is in this process's affinity mask" % node, "numa_node": node}
    os.sched_setaffinity(0, allowed)
    return {"numa_node": node, "cpus": len(allowed)}
  except Exception as e:  # noqa: BLE001 -- best effort by design
    return {"skipped": "%s: %s" % (type(e).__name__, e)}


def ptr(t):
  if t is None:
    return None
  return ctypes.c_void_p(t.data_ptr())


def alloc(nbytes: int) -> torch.Tensor:
  """Uninitialised device bytes (the caching allocator returns >= 512-byte aligned blocks)."""
  return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev())


def alloc_meta() -> torch.Tensor:
  return torch.zeros(_lib.META_WORDS, dtype=torch.int64, device=dev())


def read_u64(t: torch.Tensor) -> np.ndarray:
  """Synchronising read of a small int64 device tensor as uint64."""
  return t.cpu().numpy().view(np.uint64)


def as_torch_bytes(flat_np: np.ndarray) -> torch.Tensor:
  """Zero-copy uint8 torch view of a contiguous 1-D numpy array (read-only arrays included)."""
  b = flat_np.view(np.uint8) if flat_np.dtype != np.uint8 else flat_np
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    return torch.from_numpy(b)


# ---------------------------------------------------------------------------------- pageable host memory
# An ordinary numpy array is pageable: the driver moves it through its own bounce buffers with one
# thread (measured on the B200 box: 11.5 GB/s up, 20.7 GB/s down, against 55 / 54 GB/s for page-locked
# memory).  A ring of page-locked staging buffers filled / drained by a few memcpy threads (ctypes
# memmove runs without the GIL: 14 GB/s with one thread, 29 GB/s with eight, 38 with sixteen) keeps the DMA engines
# busy instead.  Both helpers take the fast path (plain async copies) when the host side is pinned.
STAGE_BYTES = 32 << 20
STAGE_SLOTS = 4
STAGED_MIN_BYTES = 64 << 20
STAGE_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))  # 12 threads were measured: no faster than 8
MEMCPY_MIN_PIECE = 1 << 20
_copy_pool = None
_copy_pool_lock = threading.Lock()


def is_pinned(t: torch.Tensor) -> bool:
  try:
    return bool(t.is_pinned())
  except RuntimeError:
    return False


def host_memcpy(dst_ptr: int, src_ptr: int, nbytes: int):
  """memcpy split over the copy threads (pieces of at least 1 MiB)."""
  global _copy_pool
  parts = max(1, min(STAGE_THREADS, nbytes // MEMCPY_MIN_PIECE))
  if parts == 1:
    ctypes.memmove(dst_ptr, src_ptr, nbytes)
    return
  with _copy_pool_lock:
    if _copy_pool is None:
      _copy_pool = ThreadPoolExecutor(max_workers=STAGE_THREADS, thread_name_prefix="frb-memcpy")
  per = (nbytes + parts - 1) // parts // 4096 * 4096 + 4096
  list(_copy_pool.map(lambda k: ctypes.memmove(dst_ptr + k * per, src_ptr + k * per, max(0, min(nbytes, (k + 1) * per) - k * per)),
                      range(parts)))


def _stage_buffers():
  return [torch.empty(STAGE_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(STAGE_SLOTS)]


class ChunkUploader:
  """
  Host -> device copy of the byte ranges `bounds` of host_u8 into the same ranges of dev_buf on
  `stream`; event(i) is the CUDA event that marks range i as landed.  Pinned source: everything is
  queued at once.  Pageable source: a background thread stages 32 MiB pieces through page-locked
  buffers (memcpy of piece k+1 overlaps the DMA of piece k).
  """

  def __init__(self, host_u8: torch.Tensor, dev_buf: torch.Tensor, bounds, stream):
    self.host, self.dev_buf, self.bounds, self.stream = host_u8, dev_buf, list(bounds), stream
    self.events = [None] * len(self.bounds)
    self.ready = [threading.Event() for _ in self.bounds]
    self.error = None
    self.thread = None
    total = sum(b - a for a, b in self.bounds)
    if total < STAGED_MIN_BYTES or is_pinned(host_u8):
      with torch.cuda.stream(stream):
        for i, (a, b) in enumerate(self.bounds):
          dev_buf[a:b].copy_(host_u8[a:b], non_blocking=True)
          self.events[i] = torch.cuda.Event()
          self.events[i].record(stream)
          self.ready[i].set()
    else:
      self.device = torch.cuda.current_device()
      self.thread = threading.Thread(target=self._run, name="frb-upload", daemon=True)
      self.thread.start()

  def _run(self):
    try:
      torch.cuda.set_device(self.device)
      stages = _stage_buffers()
      slot_ev = [None] * STAGE_SLOTS
      base = self.host.data_ptr()
      k = 0
      for i, (a, b) in enumerate(self.bounds):
        ev = None
        for p in range(a, b, STAGE_BYTES):
          q = min(b, p + STAGE_BYTES)
          s = k % STAGE_SLOTS
          k += 1
          if slot_ev[s] is not None:
            slot_ev[s].synchronize()
          host_memcpy(stages[s].data_ptr(), base + p, q - p)
          with torch.cuda.stream(self.stream):
            self.dev_buf[p:q].copy_(stages[s][: q - p], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
          slot_ev[s] = ev
        if ev is None:  # an empty range
          ev = torch.cuda.Event()
          ev.record(self.stream)
        self.events[i] = ev
        self.ready[i].set()
    except BaseException as e:  # surfaced by event() / close()
      self.error = e
      for r in self.ready:
        r.set()

  def event(self, i: int):
    self.ready[i].wait()
    if self.error is not None:
      raise self.error
    return self.events[i]

  def close(self):
    if self.thread is not None:
      self.thread.join()
      self.thread = None
    if self.error is not None:
      raise self.error

  def abort(self):
    """Wait for the worker and for the copies it queued, without raising (cleanup path): the device buffer and the
    staging buffers go back to the caching allocators right after, and a copy still in flight on this stream would
    be writing into memory that the next allocation on another stream may be handed."""
    if self.thread is not None:
      self.thread.join()
      self.thread = None
    try:
      self.stream.synchronize()
    except Exception:
      pass


class ChunkDownloader:
  """
  Device -> host copies into byte ranges of host_u8 on `stream`, each after a CUDA event.  Pinned
  destination: plain async copies.  Pageable destination: a background thread downloads 32 MiB pieces
  into page-locked staging buffers and the memcpy threads move them on (DMA of piece k+1 overlaps the
  memcpy of piece k).  finish() returns when every byte is in place.
  """

  def __init__(self, host_u8: torch.Tensor, stream):
    self.host, self.stream = host_u8, stream
    self.direct = is_pinned(host_u8)
    self.error = None
    self.thread = None
    if not self.direct:
      self.device = torch.cuda.current_device()
      self.q = queue.Queue()
      self.thread = threading.Thread(target=self._run, name="frb-download", daemon=True)
      self.thread.start()

  def submit(self, a: int, b: int, dev_view: torch.Tensor, after_event):
    """host[a:b] <- dev_view (b - a bytes) once after_event has completed."""
    if self.direct:
      self.stream.wait_event(after_event)
      with torch.cuda.stream(self.stream):
        self.host[a:b].copy_(dev_view, non_blocking=True)
    else:
      self.q.put((a, b, dev_view, after_event))

  def _run(self):
    try:
      torch.cuda.set_device(self.device)
      stages = _stage_buffers()
      free = list(range(STAGE_SLOTS))
      pending, inflight = collections.deque(), collections.deque()
      base = self.host.data_ptr()
      closed = False
      while True:
        block = not pending and not inflight and not closed
        while True:
          try:
            item = self.q.get(block=block)
          except queue.Empty:
            break
          block = False
          if item is None:
            closed = True
            break
          a, b, view, after = item
          for p in range(a, b, STAGE_BYTES):
            q = min(b, p + STAGE_BYTES)
            pending.append((p, q, view[p - a:q - a], after))
        while free and pending:
          p, q, view, after = pending.popleft()
          s = free.pop()
          self.stream.wait_event(after)
          with torch.cuda.stream(self.stream):
            stages[s][: q - p].copy_(view, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
          inflight.append((ev, s, p, q))
        if inflight:
          ev, s, p, q = inflight.popleft()
          ev.synchronize()
          host_memcpy(base + p, stages[s].data_ptr(), q - p)
          free.append(s)
        elif closed and not pending:
          break
    except BaseException as e:
      self.error = e

  def finish(self):
    if self.direct:
      self.stream.synchronize()
    elif self.thread is not None:
      self.q.put(None)
      self.thread.join()
      self.thread = None
    if self.error is not None:
      raise self.error

  def abort(self):
    """Stop the worker and drain the copies it queued, without raising (cleanup path; cheap after finish()): see
    ChunkUploader.abort."""
    if self.thread is not None:
      self.q.put(None)
      self.thread.join()
      self.thread = None
    try:
      self.stream.synchronize()
    except Exception:
      pass


def upload_bytes(flat_np: np.ndarray) -> torch.Tensor:
  """Contiguous 1-D numpy array -> fresh 1-D uint8 device tensor (async when the source is pinned;
  a large pageable source goes through the staging ring below)."""
  b = flat_np.view(np.uint8) if flat_np.dtype != np.uint8 else flat_np
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    src = torch.from_numpy(b)
  out = alloc(b.size)
  if b.size >= STAGED_MIN_BYTES and not is_pinned(src):
    ChunkUploader(src, out, [(0, b.size)], torch.cuda.current_stream()).close()
  elif b.size:
    out[: b.size].copy_(src, non_blocking=True)
  return out


def upload_array(a: np.ndarray) -> torch.Tensor:
  return upload_bytes(np.ascontiguousarray(a).reshape(-1))


def download_into(buf: torch.Tensor, nbytes: int, flat_np: np.ndarray):
  """Device bytes -> existing contiguous numpy memory."""
  if nbytes == 0:
    return
  b = flat_np.view(np.uint8) if flat_np.dtype != np.uint8 else flat_np
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    dst = torch.from_numpy(b)
  if nbytes >= STAGED_MIN_BYTES and not is_pinned(dst):
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    down = ChunkDownloader(dst, cur)
    down.submit(0, nbytes, buf[:nbytes], ev)
    down.finish()
    return
  dst.copy_(buf[:nbytes], non_blocking=False)


def download(buf: torch.Tensor, n: int, dtype) -> np.ndarray:
  dtype = np.dtype(dtype)
  out = host_empty((n,), dtype)
  download_into(buf, n * dtype.itemsize, out)
  return out


def flat_view(arr: np.ndarray) -> np.ndarray:
  """Zero-copy 1-D view of a C- or F-contiguous array in memory order (fastremap.pyx:157)."""
  return np.lib.stride_tricks.as_strided(arr, shape=(arr.size,), strides=(arr.dtype.itemsize,),
                                         writeable=arr.flags.writeable)


class Vol:
  """Flat device copy (or alias) of an input array, in memory order."""
  __slots__ = ("buf", "dtype", "n", "shape", "order", "is_torch", "host", "aliases_input")

  def __init__(self, buf, dtype, n, shape, order, is_torch, host, aliases_input):
    self.buf, self.dtype, self.n, self.shape, self.order = buf, np.dtype(dtype), int(n), tuple(shape), order
    self.is_torch, self.host, self.aliases_input = is_torch, host, aliases_input

  @property
  def nbytes(self):
    return self.n * self.dtype.itemsize

  @property
  def row_bytes(self):
    """Length in bytes of the fastest-varying axis (a hint for the kernels' row filters)."""
    return row_bytes(self.shape, self.order, self.dtype)


  @property
  def plane_bytes(self):
    """Bytes per plane of the two fastest-varying axes (0 for arrays of fewer than 3 axes): a hint for the
    plane walk of the counting kernel."""
    return plane_bytes(self.shape, self.order, self.dtype)


def row_bytes(shape, order, dtype) -> int:
  if len(shape) == 0:
    return 0
  return int(shape[0] if order == "F" else shape[-1]) * np.dtype(dtype).itemsize


def plane_bytes(shape, order, dtype) -> int:
  if len(shape) < 3:
    return 0
  a, b = (shape[0], shape[1]) if order == "F" else (shape[-1], shape[-2])
  return int(a) * int(b) * np.dtype(dtype).itemsize


def ingest(arr, alias_ok: bool = False) -> Vol:
  """
  Bring `arr` (numpy array or CUDA tensor) to the device as a flat memory-order buffer.
  numpy input is always uploaded into a fresh buffer.  A CUDA tensor is aliased only when
  alias_ok (the caller asked for in-place work) and its layout is C- or F-contiguous; otherwise
  it is cloned, which restates the reference's "copy unless in_place and contiguous" rule.
  """
  if isinstance(arr, torch.Tensor):
    if not arr.is_cuda:
      raise FastremapB200Error("fastremap_b200: tensors must live on a CUDA device (no CPU fallback)")
    dev()
    dt = numpy_dtype(arr.dtype)
    shape = tuple(arr.shape)
    rev = tuple(reversed(range(arr.dim())))
    aliased = True
    if arr.is_contiguous():
      order, flat = "C", arr.reshape(-1)
    elif arr.dim() > 1 and arr.permute(rev).is_contiguous():
      order, flat = "F", arr.permute(rev).reshape(-1)
    else:
      order, flat, aliased = "C", arr.contiguous().reshape(-1), False
    if dt == np.dtype("bool"):
      buf = flat.view(torch.uint8)
    else:
      buf = flat.view(torch.uint8)
    if aliased and not alias_ok:
      buf, aliased = buf.clone(), False
    if buf.numel() and buf.data_ptr() % 16 != 0:
      buf, aliased = buf.clone(), False
    return Vol(buf, dt, flat.numel(), shape, order, True, None, aliased)

  arr = np.asarray(arr)
  order = "F" if arr.flags["F_CONTIGUOUS"] else "C"
  if not (arr.flags["F_CONTIGUOUS"] or arr.flags["C_CONTIGUOUS"]):
    arr = np.copy(arr, order=order)
  flat = flat_view(arr)
  dev()
  return Vol(upload_bytes(flat), arr.dtype, arr.size, arr.shape, order, False, arr, False)


def emit(buf: torch.Tensor, dtype, shape, order, as_torch: bool):
  """Flat device buffer -> array in the caller's container with the given shape and memory order."""
  dtype = np.dtype(dtype)
  n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
  if as_torch:
    t = buf[: n * dtype.itemsize].view(torch_dtype(dtype))
    if order == "F" and len(shape) > 1:
      return t.reshape(tuple(reversed(shape))).permute(tuple(reversed(range(len(shape)))))
    return t.reshape(shape)
  out = host_empty(shape, dtype, order)
  download_into(buf, n * dtype.itemsize, flat_view(out))
  return out


PINNED_RESULT_MIN_BYTES = 1 << 20


def host_empty(shape, dtype, order="C") -> np.ndarray:
  """Result array on the host.  Large results live in page-locked memory from torch's caching
  pinned allocator, so the device->host copy runs at PCIe speed and the block is recycled once the
  caller drops the array."""
  dtype = np.dtype(dtype)
  n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
  nbytes = n * dtype.itemsize
  if nbytes < PINNED_RESULT_MIN_BYTES:
    return np.empty(shape, dtype=dtype, order=order)
  t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
  flat = t.numpy().view(dtype)  # keeps `t` alive through .base
  return flat.reshape(shape, order=order)
