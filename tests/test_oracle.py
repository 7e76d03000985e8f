"""
Pins the CPU oracle (oracle/oracle.py + fastremap_oracle.c):
  1. against the golden vectors of the reference's own test-suite (automated_test.py, cited per test),
  2. against the committed fixtures produced by the compiled reference (tests/golden/),
  3. live against the compiled reference on seeded random inputs when oracle/_ref is present.
CPU only.
"""
import numpy as np
import pytest

import golden_util as G

DTYPES = (np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64)


# ------------------------------------------------------------------ 1. reference test-suite goldens


def test_empty_renumber(oracle):  # automated_test.py:12-18
  for dtype in DTYPES:
    data2, remapdict = oracle.renumber(np.array([], dtype=dtype), preserve_zero=False)
    assert np.all(data2 == []) and remapdict == {}


@pytest.mark.parametrize("in_place", [False, True])
def test_1d_renumber(oracle, in_place):  # automated_test.py:20-57
  for dtype in DTYPES:
    data = np.flip(np.arange(8).astype(dtype))
    data2, remapdict = oracle.renumber(np.copy(data), preserve_zero=False, in_place=in_place)
    assert np.all(data2 == np.arange(1, 9)) and len(remapdict) > 0
    data2, remapdict = oracle.renumber(np.copy(data), preserve_zero=True, in_place=in_place)
    assert np.all(data2 == [1, 2, 3, 4, 5, 6, 7, 0])
  data = np.flip(np.arange(8).astype(bool))
  data2, _ = oracle.renumber(np.copy(data), preserve_zero=False, in_place=in_place)
  assert np.all(data2 == [1, 1, 1, 1, 1, 1, 1, 2])
  data2, _ = oracle.renumber(np.copy(data), preserve_zero=True, in_place=in_place)
  assert np.all(data2 == [1, 1, 1, 1, 1, 1, 1, 0])
  ro = np.copy(data)
  ro.flags.writeable = False
  oracle.renumber(ro, preserve_zero=True, in_place=in_place)


def test_2d_renumber(oracle):  # automated_test.py:60-87 -- numbering follows MEMORY order
  for dtype in DTYPES:
    data = np.array([[5, 5, 5, 2], [3, 5, 5, 0], [1, 2, 4, 1], [20, 19, 20, 1]], dtype=dtype)
    data2, _ = oracle.renumber(np.copy(data, order="C"), preserve_zero=True)
    assert np.all(data2 == [[1, 1, 1, 2], [3, 1, 1, 0], [4, 2, 5, 4], [6, 7, 6, 4]])
    data2, _ = oracle.renumber(np.copy(data, order="F"), preserve_zero=True)
    assert np.all(data2 == [[1, 1, 1, 5], [2, 1, 1, 0], [3, 5, 7, 3], [4, 6, 4, 3]])


@pytest.mark.parametrize("dtype", DTYPES)
def test_3d_renumber(oracle, dtype):  # automated_test.py:89-130
  bits = np.dtype(dtype).itemsize * 8
  big = (2 ** (bits - 1)) - 1
  data = np.array([[[big, 0], [2, big]], [[big - 5, big - 1], [big - 7, big - 3]]], dtype=dtype)
  data2, _ = oracle.renumber(np.copy(data, order="C"), preserve_zero=False)
  assert np.all(data2 == [[[1, 2], [3, 1]], [[4, 5], [6, 7]]])
  data2, _ = oracle.renumber(np.copy(data, order="F"), preserve_zero=False)
  assert np.all(data2 == [[[1, 5], [3, 1]], [[2, 6], [4, 7]]])


def test_renumber_no_preserve_zero(oracle):  # automated_test.py:138-152
  res, remap = oracle.renumber(np.array([[0, 1], [1, 2]]), start=0, preserve_zero=False)
  assert remap == {0: 0, 1: 1, 2: 2} and np.all(res == [[0, 1], [1, 2]])


def test_renumber_in_place_broken(oracle):  # automated_test.py:222-227
  data = np.array([5])
  result, mapping = oracle.renumber(data, in_place=True)
  assert result[0] == 1 and data[0] == 1 and mapping == {0: 0, 5: 1}


def test_readme_example(oracle):  # README.md:141-153
  arr = np.array([283732875, 439238823, 283732875, 182812404, 0], dtype=np.int64)
  out, mapping = oracle.renumber(arr, preserve_zero=True)
  assert out.tolist() == [1, 2, 1, 3, 0] and out.dtype == np.int8
  assert mapping == {0: 0, 283732875: 1, 439238823: 2, 182812404: 3}


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("in_place", [False, True])
def test_remap_1d(oracle, dtype, in_place):  # automated_test.py:155-181
  assert len(oracle.remap([], {})) == 0
  data = np.array([1, 2, 2, 2, 3, 4, 5], dtype=dtype)
  remap = {1: 10, 2: 30, 3: 15, 4: 0, 5: 5}
  result = oracle.remap(np.copy(data), remap, preserve_missing_labels=False, in_place=in_place)
  assert np.all(result == [10, 30, 30, 30, 15, 0, 5])
  del remap[2]
  with pytest.raises(KeyError):
    oracle.remap(np.copy(data), remap, preserve_missing_labels=False, in_place=in_place)
  result = oracle.remap(np.copy(data), remap, preserve_missing_labels=True, in_place=in_place)
  assert np.all(result == [10, 2, 2, 2, 15, 0, 5])


@pytest.mark.parametrize("dtype", DTYPES)
def test_remap_2d(oracle, dtype):  # automated_test.py:183-205
  data = np.array([[1, 2, 3, 4, 5], [5, 4, 3, 2, 1]], dtype=dtype)
  remap = {1: 10, 2: 30, 3: 15, 4: 0, 5: 5}
  result = oracle.remap(np.copy(data), remap, preserve_missing_labels=False)
  assert np.all(result == [[10, 30, 15, 0, 5], [5, 0, 15, 30, 10]])
  del remap[2]
  with pytest.raises(KeyError):
    oracle.remap(np.copy(data), remap, preserve_missing_labels=False)
  result = oracle.remap(np.copy(data), remap, preserve_missing_labels=True)
  assert np.all(result == [[10, 2, 15, 0, 5], [5, 0, 15, 2, 10]])


def test_remap_in_place_broken(oracle):  # automated_test.py:216-220
  data = np.array([0])
  result = oracle.remap(data, {0: 1}, in_place=True)
  assert result[0] == 1 and data[0] == 1


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("in_place", [True, False])
def test_mask(oracle, dtype, in_place):  # automated_test.py:229-239
  data = oracle.mask(np.arange(100, dtype=dtype), [5, 10, 15, 20], in_place=in_place)
  labels, cts = np.unique(data, return_counts=True)
  assert cts[0] == 5 and labels[0] == 0 and np.all(cts[1:] == 1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("in_place", [True, False])
def test_mask_except(oracle, dtype, in_place):  # automated_test.py:241-260
  for value in (0, 7, np.iinfo(dtype).max):
    data = oracle.mask_except(np.arange(100, dtype=dtype), [5, 10, 15, 20], in_place=in_place, value=value)
    labels, cts = np.unique(data, return_counts=True)
    assert {l: c for l, c in zip(labels, cts)} == {value: 96, 5: 1, 10: 1, 15: 1, 20: 1}


def test_renumber_remap(oracle):  # automated_test.py:556-563
  rng = np.random.default_rng(0)
  labels = rng.integers(-500, 500, size=(32, 32, 32)).astype(np.int64)
  new_labels, remap = oracle.renumber(labels, in_place=False)
  new_labels = oracle.remap(new_labels, {v: k for k, v in remap.items()}, in_place=True)
  assert np.all(labels == new_labels) and new_labels.dtype in (np.int8, np.int16)


@pytest.mark.parametrize("dtype_cc", DTYPES)
@pytest.mark.parametrize("dtype_p", DTYPES)
def test_component_map(oracle, dtype_cc, dtype_p):  # automated_test.py:565-579
  rng = np.random.default_rng(1)
  cc = rng.integers(0, 100, size=(16, 16, 16)).astype(dtype_cc)
  mapping = oracle.component_map(cc, (cc + 1).astype(dtype_p))
  assert all(k == v - 1 for k, v in mapping.items()) and len(mapping) == len(np.unique(cc))
  assert oracle.component_map([1, 2, 3, 4], [5, 5, 6, 7]) == {1: 5, 2: 5, 3: 6, 4: 7}
  assert oracle.component_map([], []) == {}


def test_unique_strategies_vs_numpy(oracle):  # automated_test.py:443-553
  rng = np.random.default_rng(2)
  for labels in (rng.integers(0, 500, size=(32, 32, 32)), rng.integers(-500, 500, size=(32, 32, 32)),
                 rng.integers(32 ** 3 - 500, 32 ** 3 + 500, size=(32, 32, 32)), rng.integers(-1000, 128 ** 3, size=(20, 20, 20))):
    u_np, i_np, inv_np, c_np = np.unique(labels, return_counts=True, return_index=True, return_inverse=True)
    u, i, inv, c = oracle.unique(labels, return_counts=True, return_index=True, return_inverse=True)
    assert np.all(u == u_np) and np.all(c == c_np) and np.all(inv.reshape(-1) == inv_np.reshape(-1))
    assert np.all(labels.flatten()[i] == labels.flatten()[i_np])
  flat = rng.integers(-1000, 128 ** 3, size=8000)
  u, c = oracle._unique_via_sort(flat)
  u_np, c_np = np.unique(flat, return_counts=True)
  assert np.all(u == u_np) and np.all(c == c_np)
  lab = np.zeros((16, 16, 16), dtype=np.int64)
  lab[0, 0, 0] = 16 ** 3 + 10
  u, i, c, inv = oracle._unique_via_renumber(lab.flatten(), return_index=True, return_inverse=True)
  u_np, c_np = np.unique(lab, return_counts=True)
  assert np.all(u == u_np) and np.all(c == c_np)
  assert np.all(oracle.unique([1, 1, 2, 3, 2, 4, 3]) == [1, 2, 3, 4])


def test_minmax(oracle):  # automated_test.py:414-418
  vol = np.random.default_rng(3).integers(-500, 500, size=(32, 32, 32))
  assert oracle.minmax(vol) == (np.min(vol), np.max(vol))


# ------------------------------------------------------------------ 2. fixtures from the compiled reference


@pytest.mark.parametrize("name", G.case_names())
def test_golden_fixture(oracle, name):
  G.check_case(G.get_case(name), oracle)


# ------------------------------------------------------------------ 3. live differential vs oracle/_ref


@pytest.mark.parametrize("dtype", DTYPES)
def test_live_vs_reference(oracle, reference, dtype):
  from fastremap_b200.synth import synth_numpy
  a = synth_numpy((20, 24, 28), (5, 6, 7), dtype, seed=int(np.dtype(dtype).num), zero_frac=0.15, noise=0.1)
  for arr in (a, np.asfortranarray(a)):
    for pz in (True, False):
      r0, m0 = reference.renumber(np.copy(arr, order="K"), preserve_zero=pz)
      r1, m1 = oracle.renumber(np.copy(arr, order="K"), preserve_zero=pz)
      assert r0.dtype == r1.dtype and np.array_equal(r0, r1) and m0 == m1
    u0, c0 = reference.unique(arr, return_counts=True)
    u1, c1 = oracle.unique(arr, return_counts=True)
    assert np.array_equal(u0, u1) and np.array_equal(c0, c1) and c0.dtype == c1.dtype
    assert reference.pixel_pairs(arr) == oracle.pixel_pairs(arr)
    assert reference.minmax(arr) == oracle.minmax(arr)
  labels = [int(x) for x in np.unique(a)]
  table = {l: (l % 97) for l in labels[::3]}
  assert np.array_equal(reference.remap(a, table, preserve_missing_labels=True), oracle.remap(a, table, preserve_missing_labels=True))
  assert np.array_equal(reference.mask_except(a, labels[::2]), oracle.mask_except(a, labels[::2]))
  par = (a.astype(np.int64) % 50).astype(np.int16)
  assert reference.component_map(a, par) == oracle.component_map(a, par)


def test_inverse_component_map(oracle):  # automated_test.py:580-594
  assert oracle.inverse_component_map([1, 2, 1, 3], [4, 4, 5, 6]) == {1: [4, 5], 2: [4], 3: [6]}
  assert oracle.inverse_component_map([1, 1, 1, 3], [4, 4, 5, 6]) == {1: [4, 5], 3: [6]}
  assert oracle.inverse_component_map([1, 1, 1, 1], [4, 2, 5, 6]) == {1: [2, 4, 5, 6]}
  assert oracle.inverse_component_map([], []) == {}


def test_inverse_component_map_vs_reference(oracle, reference):
  from fastremap_b200.synth import synth_numpy
  cc = synth_numpy((12, 20, 24), (4, 5, 6), np.uint32, seed=9, zero_frac=0.1, noise=0.1)
  par = synth_numpy((12, 20, 24), (4, 5, 6), np.int16, seed=9, zero_frac=0.1, noise=0.0, cell_div=3)
  ref = reference.inverse_component_map(par, cc)
  assert {k: sorted(v) for k, v in ref.items()} == oracle.inverse_component_map(par, cc)


def _cloud_sets(pc):
  return {k: sorted(map(tuple, np.asarray(v).tolist())) for k, v in pc.items()}


def test_point_cloud(oracle):  # automated_test.py:596-628
  x = np.ones((2, 2, 2), dtype=np.uint8)
  gt = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]])
  ptc = oracle.point_cloud(x)
  assert len(ptc) == 1 and np.all(ptc[1] == gt)
  x[0, 0, 0] = 2
  x[1, 1, 1] = 3
  ptc = oracle.point_cloud(x)
  assert len(ptc) == 3 and np.all(ptc[1] == gt[1:7]) and np.all(ptc[2] == gt[:1]) and np.all(ptc[3] == gt[7:])
  assert oracle.point_cloud(np.ones((0, 0, 0), dtype=np.uint8)) == {}


def test_point_cloud_vs_reference(oracle, reference):
  from fastremap_b200.synth import synth_numpy
  for shape, grid in (((9, 14, 20), (3, 4, 5)), ((1, 30, 26), (1, 5, 4))):
    a = synth_numpy(shape, grid, np.int32, seed=4, zero_frac=0.3, noise=0.1)
    a[a % 3 == 0] *= -1
    for arr in (a, np.asfortranarray(a), a[0]):
      assert _cloud_sets(reference.point_cloud(arr)) == _cloud_sets(oracle.point_cloud(arr))
