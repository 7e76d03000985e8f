"""CPU-only checks of the host side: C-ABI surface, dtype ladder, generator twin, loud failure without a GPU."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

import fastremap_b200 as fr
from fastremap_b200 import _lib
from fastremap_b200.synth import synth_numpy

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_c_abi_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, "include", "fastremap_b200.h")).read()
  declared = set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", header))
  assert declared, "no declarations parsed"
  assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
  lib = ctypes.CDLL(_lib.SO_PATH)
  for name in declared:
    assert hasattr(lib, name), "libfastremap_b200.so does not export %s" % name
  _lib.load()
  assert _lib.launch_count() == 0  # nothing was launched by loading


def test_reference_surface_present():
  names = ["ascontiguousarray", "asfortranarray", "component_map", "fit_dtype", "foreground", "indices",
           "inverse_component_map", "mask", "mask_except", "minmax", "narrow_dtype", "pixel_pairs", "point_cloud",
           "refit", "remap", "remap_from_array", "remap_from_array_kv", "renumber", "tobytes", "transpose", "unique",
           "widen_dtype"]  # reference fastremap/__init__.py:1-24
  for n in names:
    assert callable(getattr(fr, n))


def test_signatures_match_reference_stub():
  import inspect
  sig = inspect.signature
  assert list(sig(fr.renumber).parameters)[:4] == ["arr", "start", "preserve_zero", "in_place"]
  assert [p.default for p in sig(fr.renumber).parameters.values()][1:4] == [1, True, False]
  assert list(sig(fr.remap).parameters) == ["arr", "table", "preserve_missing_labels", "in_place"]
  assert [p.default for p in sig(fr.remap).parameters.values()][2:] == [False, False]
  assert list(sig(fr.mask).parameters) == ["arr", "labels", "in_place", "value"]
  assert list(sig(fr.mask_except).parameters) == ["arr", "labels", "in_place", "value"]
  assert list(sig(fr.unique).parameters)[:5] == ["labels", "return_index", "return_inverse", "return_counts", "axis"]
  assert list(sig(fr.remap_from_array).parameters) == ["arr", "vals", "in_place"]
  assert sig(fr.remap_from_array).parameters["in_place"].default is True
  assert list(sig(fr.remap_from_array_kv).parameters) == ["arr", "keys", "vals", "preserve_missing_labels", "in_place"]
  assert list(sig(fr.refit).parameters) == ["arr", "value", "increase_only", "exotics"]
  assert list(sig(fr.component_map).parameters) == ["component_labels", "parent_labels"]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_fit_dtype_uint(dtype):  # reference automated_test.py:323-343
  assert fr.fit_dtype(dtype, 0) == np.uint8
  assert fr.fit_dtype(dtype, 255) == np.uint8
  assert fr.fit_dtype(dtype, 256) == np.uint16
  assert fr.fit_dtype(dtype, 2 ** 16 - 1) == np.uint16
  assert fr.fit_dtype(dtype, 2 ** 16) == np.uint32
  assert fr.fit_dtype(dtype, 2 ** 32) == np.uint64
  assert fr.fit_dtype(dtype, 2 ** 64 - 1) == np.uint64
  with pytest.raises(ValueError):
    fr.fit_dtype(dtype, -1)
  with pytest.raises(ValueError):
    fr.fit_dtype(dtype, 2 ** 64)


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.int32, np.int64])
def test_fit_dtype_int(dtype):  # reference automated_test.py:345-365
  assert fr.fit_dtype(dtype, 0) == np.int8
  assert fr.fit_dtype(dtype, 127) == np.int8
  assert fr.fit_dtype(dtype, -128) == np.int8
  assert fr.fit_dtype(dtype, 128) == np.int16
  assert fr.fit_dtype(dtype, 2 ** 15 - 1) == np.int16
  assert fr.fit_dtype(dtype, 2 ** 15) == np.int32
  assert fr.fit_dtype(dtype, 2 ** 32) == np.int64
  assert fr.fit_dtype(dtype, 2 ** 63 - 1) == np.int64
  with pytest.raises(ValueError):
    fr.fit_dtype(dtype, 2 ** 63)


def test_fit_dtype_float_and_complex():  # reference automated_test.py:367-412
  for dtype in (np.float16, np.float32, np.float64):
    assert fr.fit_dtype(dtype, 2 ** 63 - 1) == np.float32
    assert fr.fit_dtype(dtype, 2 ** 128) == np.float64
    assert fr.fit_dtype(dtype, 10000, exotics=True) == np.float16
    assert fr.fit_dtype(dtype, 2 ** 32, exotics=True) == np.float32
  for dtype in (np.csingle, np.cdouble):
    for sign in (1, -1, 1j, -1j):
      assert fr.fit_dtype(dtype, sign * 2 ** 32) == np.csingle
      with pytest.raises(ValueError):
        fr.fit_dtype(dtype, sign * 2 ** 128)
      assert fr.fit_dtype(dtype, sign * 2 ** 128, exotics=True) == np.cdouble


def test_narrow_widen_dtype():  # reference automated_test.py:669-718
  assert fr.narrow_dtype(np.uint64) == np.uint32 and fr.narrow_dtype(np.uint8) == np.uint8
  assert fr.narrow_dtype(np.int16) == np.int8 and fr.narrow_dtype(np.int64) == np.int32
  assert fr.narrow_dtype(np.float64) == np.float32 and fr.narrow_dtype(np.float32) == np.float32
  assert fr.narrow_dtype(np.float32, exotics=True) == np.float16
  assert fr.narrow_dtype(np.complex128) == np.complex64
  assert fr.widen_dtype(np.uint64) == np.uint64 and fr.widen_dtype(np.uint8) == np.uint16
  assert fr.widen_dtype(np.int32) == np.int64 and fr.widen_dtype(np.float16) == np.float32
  assert fr.widen_dtype(np.float64) == np.float64
  assert fr.widen_dtype(np.float64, exotics=True) == np.longdouble
  assert fr.widen_dtype(np.complex64) == np.complex64
  assert fr.widen_dtype(np.complex64, exotics=True) == np.complex128


def test_synth_twin_is_a_function_of_global_index():
  full = synth_numpy((12, 10, 16), (3, 2, 4), np.uint64, seed=5, zero_frac=0.2, noise=0.05)
  lo = synth_numpy((5, 10, 16), (3, 2, 4), np.uint64, seed=5, zero_frac=0.2, noise=0.05, z0=0, nz_total=12)
  hi = synth_numpy((7, 10, 16), (3, 2, 4), np.uint64, seed=5, zero_frac=0.2, noise=0.05, z0=5, nz_total=12)
  assert np.array_equal(full, np.concatenate([lo, hi], axis=0))
  assert (full == 0).any() and np.unique(full).size > 10
  again = synth_numpy((12, 10, 16), (3, 2, 4), np.uint64, seed=5, zero_frac=0.2, noise=0.05)
  assert np.array_equal(full, again)
  for dt in (np.int8, np.uint16, np.int64):
    v = synth_numpy((4, 4, 4), (2, 2, 2), dt, seed=1)
    assert v.dtype == np.dtype(dt) and v.min() >= 0


def test_host_side_errors_precede_device_work():
  a = np.arange(10, dtype=np.uint8)
  with pytest.raises(TypeError):
    fr.mask_except(a, (1, 2))  # labels must be a list (reference: typed `list labels`, fastremap.pyx:495)
  with pytest.raises(OverflowError):
    fr.mask_except(a, [300])
  with pytest.raises(OverflowError):
    fr.remap(a, {300: 1, 1: 2}, preserve_missing_labels=True)
  with pytest.raises(ValueError):
    fr.remap(a, {1: -1}, preserve_missing_labels=True)  # negative into unsigned (fit_dtype, fastremap.pyx:331-332)
  with pytest.raises(ValueError):
    fr.component_map(np.zeros((2, 3), np.uint8), np.zeros((3, 2), np.uint8))
  with pytest.raises(ValueError):
    fr.remap_from_array(np.zeros(4, np.uint8), np.zeros(4, np.uint16))
  assert fr.renumber(np.array([], dtype=np.uint32))[1] == {}
  assert fr.component_map([], []) == {}
  assert fr.minmax(np.array([], dtype=np.int32)) == (None, None)
  assert fr.pixel_pairs(np.array([], dtype=np.int32)) == 0
  out, m = fr.renumber(np.array([True, False]), start=7)
  assert m == {0: 0, 1: 7} and out.dtype == bool


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
  with pytest.raises(fr.FastremapB200Error):
    fr.renumber(np.arange(10, dtype=np.uint32))
  with pytest.raises(fr.FastremapB200Error):
    fr.unique(np.arange(10, dtype=np.uint32), return_counts=True)
  with pytest.raises(fr.FastremapB200Error):
    fr.remap(np.arange(10, dtype=np.uint32), {1: 2, 3: 4}, preserve_missing_labels=True)


def test_product_does_not_import_oracle():
  pkg = os.path.join(ROOT, "fastremap_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh")):
        src = open(os.path.join(dirpath, f)).read()
        assert "import oracle" not in src and "oracle/" not in src and "_ref" not in src, f


def test_layout_host_helpers():
  from fastremap_b200 import layout
  assert layout._strides_for((3, 4, 5), "C", 2) == (40, 10, 2)
  assert layout._strides_for((3, 4, 5), "F", 2) == (2, 6, 24)
  assert layout._mem_dims((3, 4, 5), "C") == (5, 4, 3) and layout._mem_dims((3, 4, 5), "F") == (3, 4, 5)
  a = np.arange(24, dtype=np.uint16).reshape(2, 3, 4)
  flat = np.ascontiguousarray(a.T).reshape(-1)                   # memory of the F layout
  f = layout._restride(flat, a.shape, "F")
  assert f.flags.f_contiguous and np.array_equal(f, a)
  assert np.array_equal(np.asfortranarray(a), f)


def test_sizing_rules():
  from fastremap_b200 import api, sharded
  assert api._lookup_slots_per_entry(1000) == api.LOOKUP_SLOTS_PER_ENTRY_MAX
  assert api._lookup_slots_per_entry(5_000_000) == 16                      # 32 slots per label would pass 2 GiB
  assert api._lookup_slots_per_entry(50_000_000) == api.LOOKUP_SLOTS_PER_ENTRY
  m, total, cap = sharded._guess_from([1000, 3000, 2000], 1 << 20)
  assert m >= 3000 * 1.25 and m % 1024 == 0 and total >= 6000 * 1.25 and cap == 1 << 20   # the grown table size is remembered
  assert api._default_entries(1 << 30, None) == (1 << 30) // 256 and api._default_entries(1 << 30, None, 512) == (1 << 30) // 512
  assert api._default_entries(1 << 30, 1000) == 1001 and api._default_entries(1000, None) == 1 << 15


def test_host_memcpy_threads():
  """The staging ring's parallel memcpy (pure host code)."""
  from fastremap_b200 import device as D
  rng = np.random.default_rng(0)
  src = rng.integers(0, 255, size=(5 << 20) + 123, dtype=np.uint8)
  dst = np.zeros_like(src)
  D.host_memcpy(dst.ctypes.data, src.ctypes.data, src.size)
  assert np.array_equal(src, dst)
  dst[:] = 0
  D.host_memcpy(dst.ctypes.data + 7, src.ctypes.data + 7, 1000)    # single piece
  assert np.array_equal(dst[7:1007], src[7:1007]) and dst[:7].sum() == 0 and dst[1007:].sum() == 0


def test_type_stub_matches_the_package():
  """fastremap_b200/__init__.pyi (shipped with py.typed) lists every public name with the parameters the functions really take."""
  import ast
  import inspect
  tree = ast.parse(open(os.path.join(ROOT, "fastremap_b200", "__init__.pyi")).read())
  stubs = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
  assert os.path.exists(os.path.join(ROOT, "fastremap_b200", "py.typed"))
  for name in fr.__all__:
    assert name in stubs, name
    real = inspect.signature(getattr(fr, name)).parameters
    a = stubs[name].args
    assert [x.arg for x in a.args + a.kwonlyargs] == list(real), name


def test_import_fastremap_shim():
  """fastremap_b200/compat on sys.path makes `import fastremap` resolve to this package (same function objects)."""
  import importlib
  compat = os.path.join(ROOT, "fastremap_b200", "compat")
  saved = sys.modules.pop("fastremap", None)
  sys.path.insert(0, compat)
  try:
    shim = importlib.import_module("fastremap")
    assert os.path.dirname(os.path.dirname(shim.__file__)) == compat
    for name in fr.__all__:
      assert getattr(shim, name) is getattr(fr, name), name
  finally:
    sys.path.remove(compat)
    sys.modules.pop("fastremap", None)
    if saved is not None:
      sys.modules["fastremap"] = saved


def test_own_table_sparsity_rule():
  """sharded renumber: the rank's own lookup table gets up to 64 slots per label while it stays at 2^23 slots, never fewer than 8"""
  from fastremap_b200 import sharded, api
  assert sharded._own_slots_per_entry(1) == 64
  assert sharded._own_slots_per_entry(88_000) == 64
  assert sharded._own_slots_per_entry(400_000) == (1 << 23) // 400_000
  assert sharded._own_slots_per_entry(5_000_000) == api.LOOKUP_SLOTS_PER_ENTRY
