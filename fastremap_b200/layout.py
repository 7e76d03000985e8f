"""
Memory-layout operations either side of the relabel path:
`transpose`, `asfortranarray`, `ascontiguousarray` (reference: fastremap.pyx:1140-1273 over
ipt.hpp) and `tobytes` (fastremap.pyx:1547-1654), on the GPU through the C ABI's
frb_transpose_tiles / frb_copy_rows (fastremap_b200/csrc/frb_layout.cu).

The reference permutes the caller's buffer IN PLACE and returns a re-strided view of it.  The
kernels here move the data out of place on the device (one pass through HBM) and the result is
written back into the caller's buffer, so the caller sees the same thing: its array's memory holds
the transposed data and the returned array shares that memory.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from . import device as D

# COMPLEX_NUMBER (fastremap.pyx:30-59): the integer types, float, double, float complex (+ bool via cast=True)
_LAYOUT_DTYPES = frozenset(np.dtype(c) for c in ("u1", "u2", "u4", "u8", "i1", "i2", "i4", "i8", "f4", "f8", "c8", "bool"))
# NUMBER (fastremap.pyx:30-55); the compiled reference also lets bool through its fused-type dispatch (fixture tobytes_bool)
_NUMBER_DTYPES = frozenset(np.dtype(c) for c in ("u1", "u2", "u4", "u8", "i1", "i2", "i4", "i8", "f4", "f8", "bool"))
_TORCH_LAYOUT = {torch.uint8: "u1", torch.uint16: "u2", torch.uint32: "u4", torch.uint64: "u8", torch.int8: "i1",
                 torch.int16: "i2", torch.int32: "i4", torch.int64: "i8", torch.float32: "f4", torch.float64: "f8",
                 torch.complex64: "c8", torch.bool: "bool"}
MAX_AXES = 8  # two tile axes + six batch axes


def _u64_array(values):
  return (ctypes.c_uint64 * max(len(values), 1))(*[int(v) for v in values])


def _prod(values) -> int:
  out = 1
  for v in values:
    out *= int(v)
  return out


def reverse_axes_device(buf: torch.Tensor, mem_dims, width: int):
  """
  Physically reverse the axis order of a dense array held in `buf` (uint8 device bytes).
  mem_dims: axis lengths in MEMORY order, fastest-varying first.  Returns a fresh buffer, or
  `buf` itself when no element moves (at most one axis longer than 1).  This is what
  pyipt::_ipt2d/3d/4d do to the caller's memory (ipt.hpp:330-352).
  """
  dims = [int(d) for d in mem_dims if int(d) != 1]
  n = _prod(dims)
  if len(dims) <= 1 or n == 0:
    return buf
  if len(dims) > MAX_AXES:
    raise NotImplementedError("fastremap_b200: transposition supports at most %d axes longer than 1" % MAX_AXES)
  out = D.alloc(n * width)
  k = len(dims)
  mids = list(range(1, k - 1))
  rc = _lib.load().frb_transpose_tiles(
    D.ptr(buf), D.ptr(out), width, dims[0], dims[-1], _prod(dims[:-1]), _prod(dims[1:]), len(mids),
    _u64_array([dims[j] for j in mids]), _u64_array([_prod(dims[:j]) for j in mids]),
    _u64_array([_prod(dims[j + 1:]) for j in mids]), D.stream())
  _lib.check(rc, "frb_transpose_tiles")
  return out


def _np_layout_dtype(arr, strict=True):
  """The typed paths (:1275-1405) take COMPLEX_NUMBER only; where the reference goes through numpy
  instead (strided input, other ranks) any 1 / 2 / 4 / 8-byte element will do."""
  if (arr.dtype not in _LAYOUT_DTYPES) if strict else (arr.dtype.itemsize not in (1, 2, 4, 8)):
    raise TypeError("No matching signature found")
  return arr.dtype


def _mem_dims(shape, order):
  return tuple(shape) if order == "F" else tuple(reversed(shape))


def _strides_for(shape, order, nbytes):
  strides, acc = [], nbytes
  for d in (shape if order == "F" else reversed(shape)):
    strides.append(acc)
    acc *= int(d)
  return tuple(strides if order == "F" else reversed(strides))


def _permute_numpy(arr: np.ndarray, order: str, in_place: bool):
  """Reverse the physical axis order of the C- or F-contiguous `arr`; returns the flat host memory
  (the array's own when in_place) that now holds the permuted elements."""
  width = arr.dtype.itemsize
  if in_place and not arr.flags.writeable:
    raise ValueError("buffer source array is read-only")
  flat = D.flat_view(arr)
  D.dev()
  src = D.upload_bytes(flat)
  out = reverse_axes_device(src, _mem_dims(arr.shape, order), width)
  if in_place:
    if out is not src:
      D.download_into(out, arr.size * width, flat)
    return flat
  res = D.host_empty((arr.size,), arr.dtype)
  D.download_into(out, arr.size * width, res)
  return res


def _restride(flat: np.ndarray, shape, order):
  return np.lib.stride_tricks.as_strided(flat, shape=shape, strides=_strides_for(shape, order, flat.dtype.itemsize))


def _tensor_flat(t: torch.Tensor):
  """(flat memory-order alias, order) of a C- or F-contiguous CUDA tensor, else (contiguous copy, 'C')."""
  if not t.is_cuda:
    raise D.FastremapB200Error("fastremap_b200: tensors must live on a CUDA device (no CPU fallback)")
  if t.dtype not in _TORCH_LAYOUT:
    raise TypeError("No matching signature found")
  D.dev()
  rev = tuple(reversed(range(t.dim())))
  if t.is_contiguous():
    return t.reshape(-1), "C", True
  if t.dim() > 1 and t.permute(rev).is_contiguous():
    return t.permute(rev).reshape(-1), "F", True
  return t.contiguous().reshape(-1), "C", False


def _tensor_restride(flat: torch.Tensor, shape, order):
  if order == "F" and len(shape) > 1:
    return flat.reshape(tuple(reversed(shape))).permute(tuple(reversed(range(len(shape)))))
  return flat.reshape(shape)


def _permute_tensor(flat: torch.Tensor, shape, order):
  """In-place (as the caller sees it) reversal of the physical axis order of a flat CUDA tensor."""
  width = flat.element_size()
  raw = flat.view(torch.uint8)
  src = raw
  if raw.numel() and raw.data_ptr() % 16 != 0:
    src = raw.clone()
  out = reverse_axes_device(src, _mem_dims(shape, order), width)
  if out is not src:
    raw.copy_(out[: raw.numel()])
  return flat


def _layout_change(arr, want: str):
  """asfortranarray (want='F') / ascontiguousarray (want='C')."""
  have = "C" if want == "F" else "F"
  if isinstance(arr, torch.Tensor):
    flat, order, aliased = _tensor_flat(arr)
    shape = tuple(arr.shape)
    rev = tuple(reversed(range(arr.dim())))
    is_want = arr.is_contiguous() if want == "C" else (arr.dim() <= 1 or arr.permute(rev).is_contiguous())
    if is_want:
      return arr
    if arr.dim() <= 1:
      return flat.reshape(shape)
    if not aliased and want == "C":
      return flat.reshape(shape)  # a strided tensor: the contiguous copy is already the answer
    _permute_tensor(flat, shape, have)  # `flat` aliases the input, or is the fresh C copy of a strided one
    return _tensor_restride(flat, shape, want)

  arr = np.asarray(arr)
  if arr.flags["F_CONTIGUOUS" if want == "F" else "C_CONTIGUOUS"]:
    return arr
  if not arr.flags["C_CONTIGUOUS" if want == "F" else "F_CONTIGUOUS"]:
    # :1188-1189 / :1236-1237 hand a strided array to numpy; here: normalise the layout on the host
    # (as every other entry point does for strided input), then move the data on the GPU
    arr = np.copy(arr, order=have)
    fresh = True
  else:
    fresh = False
  if arr.ndim <= 1:
    return arr
  in_place = fresh or arr.ndim in (2, 3, 4)  # :1202-1222: other ranks go through numpy's out-of-place copy
  dtype = _np_layout_dtype(arr, strict=not fresh and arr.ndim in (2, 3, 4))
  work = arr.view(np.uint8) if dtype == np.dtype("bool") else arr
  flat = _permute_numpy(work, have, in_place)
  return _restride(flat, arr.shape, want).view(dtype)


def asfortranarray(arr):
  """Fortran-ordered array with the same shape and values; a C-contiguous 2-D .. 4-D input is
  permuted in its own memory and the result is a view of it (reference: fastremap.pyx:1175-1221)."""
  return _layout_change(arr, "F")


def ascontiguousarray(arr):
  """C-ordered array with the same shape and values; an F-contiguous 2-D .. 4-D input is permuted
  in its own memory and the result is a view of it (reference: fastremap.pyx:1223-1273)."""
  return _layout_change(arr, "C")


def transpose(arr):
  """
  Physical transposition (reverse the axis order of the memory) of a 2-D .. 4-D array in its own
  buffer (reference: fastremap.pyx:1140-1173).  Like the reference, the returned array keeps the
  INPUT's shape and strides: for square / cubic arrays that is the transposed array, for other shapes
  it is the transposed data seen through the old shape.  Other ranks return `arr.T`.
  """
  if isinstance(arr, torch.Tensor):
    if arr.dim() not in (2, 3, 4):
      return arr.permute(tuple(reversed(range(arr.dim()))))
    flat, order, aliased = _tensor_flat(arr)
    shape = tuple(arr.shape)
    _permute_tensor(flat, shape, order)
    return arr if aliased else _tensor_restride(flat, shape, order)
  arr = np.asarray(arr)
  if not arr.flags["F_CONTIGUOUS"] and not arr.flags["C_CONTIGUOUS"]:
    arr = np.copy(arr, order="C")
  if arr.ndim not in (2, 3, 4):
    return arr.T
  dtype = _np_layout_dtype(arr)
  work = arr.view(np.uint8) if dtype == np.dtype("bool") else arr
  order = "F" if arr.flags["F_CONTIGUOUS"] else "C"  # :1281-1286: F wins when both hold
  _permute_numpy(work, order, True)
  return arr


# ---------------------------------------------------------------------------------- tobytes

def chunk_layout_device(buf: torch.Tensor, shape, in_order: str, chunk, out_order: str, width: int) -> torch.Tensor:
  """
  Device part of tobytes: rearrange a 3-D volume (flat memory-order bytes in `buf`) into its grid
  of chunks, chunk after chunk in Fortran order of the grid, each chunk laid out in `out_order`.
  One pass: a batched row copy when the contiguous axis is the same on both sides, a batched tile
  transposition otherwise.
  """
  sx, sy, sz = (int(s) for s in shape)
  cx, cy, cz = (int(c) for c in chunk)
  g = (sx // cx, sy // cy, sz // cz)
  ce = cx * cy * cz
  cdim = (cx, cy, cz)
  in_s = (sy * sz, sz, 1) if in_order == "C" else (1, sx, sx * sy)           # element strides of (x, y, z)
  out_s = (cy * cz, cz, 1) if out_order == "C" else (1, cx, cx * cy)         # inside a chunk
  gin = tuple(cdim[i] * in_s[i] for i in range(3))                           # one grid step
  gout = (ce, g[0] * ce, g[0] * g[1] * ce)
  out = D.alloc(sx * sy * sz * width)
  fi = 2 if in_order == "C" else 0
  fo = 2 if out_order == "C" else 0
  L = _lib.load()
  if fi == fo:
    within = [i for i in range(3) if i != fo]
    dims = [cdim[i] for i in within] + list(g)
    bis = [in_s[i] * width for i in within] + [s * width for s in gin]
    bos = [out_s[i] * width for i in within] + [s * width for s in gout]
    # walk the rows in INPUT order: the reads are then one sequential stream (a short row's neighbour in
    # the same 128-byte line is consumed at once; in output order it was fetched twice -- ncu showed 1.63x
    # the algorithmic read bytes for 64-byte rows), and scattered full-sector writes cost nothing extra
    keep = sorted((j for j, d in enumerate(dims) if d != 1), key=lambda j: bis[j])
    rc = L.frb_copy_rows(D.ptr(buf), D.ptr(out), cdim[fo] * width, len(keep), _u64_array([dims[j] for j in keep]),
                         _u64_array([bis[j] for j in keep]), _u64_array([bos[j] for j in keep]), D.stream())
    _lib.check(rc, "frb_copy_rows")
  else:
    dims = [cdim[1]] + list(g)
    bis = [in_s[1]] + list(gin)
    bos = [out_s[1]] + list(gout)
    keep = [j for j, d in enumerate(dims) if d != 1]
    rc = L.frb_transpose_tiles(D.ptr(buf), D.ptr(out), width, cdim[fi], cdim[fo], in_s[fo], out_s[fi], len(keep),
                               _u64_array([dims[j] for j in keep]), _u64_array([bis[j] for j in keep]),
                               _u64_array([bos[j] for j in keep]), D.stream())
    _lib.check(rc, "frb_transpose_tiles")
  return out


def tobytes(image, chunk_size, order: str = "C"):
  """
  `cutout.tobytes(order)` for every cutout of the chunk grid, listed in Fortran order of the grid
  (reference: fastremap.pyx:1547-1654).  The rearrangement runs on the GPU in one pass; the host only
  slices the downloaded buffer into `bytes` objects.
  """
  is_t = isinstance(image, torch.Tensor)
  if is_t:
    if image.dtype not in _TORCH_LAYOUT or np.dtype(_TORCH_LAYOUT[image.dtype]) not in _NUMBER_DTYPES:
      raise TypeError("No matching signature found")
    nd, shape_t = image.dim(), tuple(image.shape)
  else:
    image = np.asarray(image)
    if image.dtype not in _NUMBER_DTYPES:
      raise TypeError("No matching signature found")
    nd, shape_t = image.ndim, image.shape
  if nd != 3:
    raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % nd)
  if order not in ["C", "F"]:
    raise ValueError(f"order must be C or F. Got: {order}")
  chunk_f = np.array(chunk_size, dtype=float)
  shape_f = np.array(shape_t, dtype=float)
  if chunk_f.shape != (3,) or np.any(chunk_f < 1) or np.any(chunk_f != np.floor(chunk_f)):
    raise ValueError(f"chunk_size ({chunk_f}) must be three positive integers.")
  if np.any(np.remainder(shape_f, chunk_f)):
    raise ValueError(f"chunk_size ({chunk_f}) must evenly divide the image shape ({shape_f}).")
  chunk = tuple(int(c) for c in chunk_f)
  n = _prod(shape_t)
  if n == 0:
    return []
  if is_t:
    flat, in_order, _ = _tensor_flat(image)
    width = flat.element_size()
    buf = flat.view(torch.uint8)
    if buf.data_ptr() % 16 != 0:
      buf = buf.clone()
  else:
    v = D.ingest(image)
    buf, in_order, width = v.buf, v.order, v.dtype.itemsize
  out = chunk_layout_device(buf, shape_t, in_order, chunk, order, width)
  host = D.host_empty((n * width,), np.uint8)
  D.download_into(out, n * width, host)
  cb = _prod(chunk) * width
  mv = memoryview(host)  # (a thread pool filling preallocated bytes objects was measured: no faster, page faults dominate)
  return [bytes(mv[i * cb:(i + 1) * cb]) for i in range(n * width // cb)]
