"""
Pure-Python dtype ladder: fit_dtype / widen_dtype / narrow_dtype.

These are host-side bookkeeping of the reference's API (fastremap/fastremap.pyx:303-436) and
touch no array data, so they stay Python here as they are Python there.
"""
from __future__ import annotations

import numpy as np


def fit_dtype(dtype, value, exotics=False):
  """Smallest dtype of the same kind that can hold `value` (reference: fastremap.pyx:303-359)."""
  dtype = np.dtype(dtype)
  if np.issubdtype(dtype, np.floating):
    sequence = [np.float16, np.float32, np.float64] if exotics else [np.float32, np.float64]
    info = np.finfo
  elif np.issubdtype(dtype, np.unsignedinteger):
    sequence = [np.uint8, np.uint16, np.uint32, np.uint64]
    info = np.iinfo
    if value < 0:
      raise ValueError(str(value) + " is negative but unsigned data type {} is selected.".format(dtype))
  elif np.issubdtype(dtype, np.complexfloating):
    sequence = [np.csingle, np.cdouble] if exotics else [np.csingle]
    info = np.finfo
  elif np.issubdtype(dtype, np.integer):
    sequence = [np.int8, np.int16, np.int32, np.int64]
    info = np.iinfo
  else:
    raise ValueError(
      "Unsupported dtype: {} Only standard floats, integers, and complex types are supported.".format(dtype))

  test_value = np.real(value)
  if abs(np.real(value)) < abs(np.imag(value)):
    test_value = np.imag(value)

  for candidate in sequence:
    lim = info(candidate)
    if test_value >= 0 and lim.max >= test_value:
      return candidate
    if test_value < 0 and lim.min <= test_value:
      return candidate
  raise ValueError("Unable to find a compatible dtype for {} that can fit {}".format(dtype, value))


def widen_dtype(dtype, exotics: bool = False):
  """Next wider dtype of the same kind; the widest maps to itself (reference: fastremap.pyx:361-399)."""
  dtype = np.dtype(dtype)
  if np.issubdtype(dtype, np.floating):
    sequence = [np.float16, np.float32, np.float64] + ([np.longdouble] if exotics else [])
  elif np.issubdtype(dtype, np.unsignedinteger):
    sequence = [np.uint8, np.uint16, np.uint32, np.uint64]
  elif np.issubdtype(dtype, np.complexfloating):
    sequence = [np.complex64] + ([np.complex128, np.clongdouble] if exotics else [])
  elif np.issubdtype(dtype, np.integer):
    sequence = [np.int8, np.int16, np.int32, np.int64]
  elif exotics:
    raise ValueError(f"Unsupported dtype: {dtype}\n")
  else:
    raise ValueError(
      f"Unsupported dtype: {dtype}\n"
      f"Only standard floats, integers, and complex types are supported."
      f"For additional types (e.g. long double, complex128, clongdouble), enable exotics.")
  idx = sequence.index(dtype)
  return sequence[min(idx + 1, len(sequence) - 1)]


def narrow_dtype(dtype, exotics: bool = False):
  """Next narrower dtype of the same kind; 8-bit maps to itself (reference: fastremap.pyx:401-436)."""
  dtype = np.dtype(dtype)
  if dtype.itemsize == 1:
    return dtype
  if np.issubdtype(dtype, np.floating):
    sequence = ([np.float16] if exotics else []) + [np.float32, np.float64, np.longdouble]
  elif np.issubdtype(dtype, np.unsignedinteger):
    sequence = [np.uint8, np.uint16, np.uint32, np.uint64]
  elif np.issubdtype(dtype, np.complexfloating):
    sequence = [np.complex64, np.complex128, np.clongdouble]
  elif np.issubdtype(dtype, np.integer):
    sequence = [np.int8, np.int16, np.int32, np.int64]
  else:
    raise ValueError(
      f"Unsupported dtype: {dtype}\n"
      f"Only standard floats, integers, and complex types are supported.")
  idx = sequence.index(dtype)
  return sequence[max(idx - 1, 0)]
