"""
Device plumbing: torch owns device memory and streams, ctypes hands raw pointers to the C ABI.

A `Vol` is a flat device buffer holding an array's elements in MEMORY order (C arrays are walked
C-wise, F arrays F-wise -- fastremap.pyx:150-157), plus what is needed to hand the result back in
the caller's own container (numpy array or CUDA torch tensor), shape and order.
"""
from __future__ import annotations

import ctypes
import warnings

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


def upload_bytes(flat_np: np.ndarray) -> torch.Tensor:
  """Contiguous 1-D numpy array -> fresh 1-D uint8 device tensor (async when the source is pinned)."""
  b = flat_np.view(np.uint8) if flat_np.dtype != np.uint8 else flat_np
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    src = torch.from_numpy(b)
  out = alloc(b.size)
  if b.size:
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


def row_bytes(shape, order, dtype) -> int:
  if len(shape) == 0:
    return 0
  return int(shape[0] if order == "F" else shape[-1]) * np.dtype(dtype).itemsize


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
