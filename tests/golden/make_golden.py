#!/usr/bin/env python
"""
Generate tests/golden/golden.npz + golden_manifest.json by running the compiled, UNMODIFIED
reference (oracle/_ref, built by oracle/build_ref.sh from /root/reference) on seeded inputs.

Run here (the container that has /root/reference); the outputs are committed so that the oracle
and the CUDA path can be pinned against the real reference anywhere, including the GPU box.

Every case is {fn, args, kwargs} -> result.  Arrays live in the npz under "<case>/<slot>";
dicts are stored as sorted key / value arrays; an expected exception is stored by type + message.
Large cases store a sha256 digest of the result bytes instead of the bytes.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(1, ROOT)

import fastremap as ref  # noqa: E402  (the compiled reference)
from fastremap_b200.synth import synth_numpy  # noqa: E402

DTYPES = ["uint8", "uint16", "uint32", "uint64", "int8", "int16", "int32", "int64"]

arrays = {}
manifest = []


def digest(a: np.ndarray) -> str:
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Synth:
  """An input given by its synth_numpy recipe (not stored in the npz)."""

  def __init__(self, order="C", **params):
    self.params, self.order = params, order

  def make(self):
    a = synth_numpy(**self.params)
    return np.asfortranarray(a) if self.order == "F" else a


def put_arg(case, slot, value):
  if isinstance(value, Synth):
    return {"synth": value.params, "order": value.order}
  if isinstance(value, np.ndarray):
    key = "%s/%s" % (case, slot)
    arrays[key] = value
    return {"array": key, "order": "F" if (value.flags.f_contiguous and not value.flags.c_contiguous) else "C"}
  if isinstance(value, dict):
    ks = list(value.keys())
    arrays["%s/%s.k" % (case, slot)] = np.array(ks, dtype=object).astype(np.int64 if all(-2**63 <= k < 2**63 for k in ks) else np.uint64) if ks else np.zeros(0, np.int64)
    vs = list(value.values())
    arrays["%s/%s.v" % (case, slot)] = np.array(vs, dtype=object).astype(np.int64 if all(-2**63 <= k < 2**63 for k in vs) else np.uint64) if vs else np.zeros(0, np.int64)
    return {"dict": "%s/%s" % (case, slot)}
  if isinstance(value, list):
    return {"list": [int(v) for v in value]}
  return {"value": value}


def put_result(case, res, big=False):
  def one(slot, r):
    if isinstance(r, np.ndarray):
      if big:
        return {"digest": digest(r), "dtype": r.dtype.name, "shape": list(r.shape)}
      arrays["%s/%s" % (case, slot)] = r
      return {"array": "%s/%s" % (case, slot), "dtype": r.dtype.name,
              "order": "F" if (r.flags.f_contiguous and not r.flags.c_contiguous) else "C"}
    if isinstance(r, dict):
      ks = sorted(r.keys())
      kd = np.uint64 if any(k >= 2**63 for k in ks) else np.int64
      vs = [r[k] for k in ks]
      vd = np.uint64 if any(v >= 2**63 for v in vs) else np.int64
      ka, va = np.array(ks, dtype=kd), np.array(vs, dtype=vd)
      if big:
        return {"dict_digest": [digest(ka), digest(va)], "len": len(ks), "kdtype": np.dtype(kd).name, "vdtype": np.dtype(vd).name}
      arrays["%s/%s.k" % (case, slot)] = ka
      arrays["%s/%s.v" % (case, slot)] = va
      return {"dict": "%s/%s" % (case, slot)}
    if r is None:
      return {"value": None}
    if isinstance(r, list):  # tobytes: list[bytes]
      arrays["%s/%s" % (case, slot)] = np.frombuffer(b"".join(r), dtype=np.uint8)
      return {"bytes_list": "%s/%s" % (case, slot), "lengths": [len(b) for b in r]}
    return {"value": int(r)}
  if isinstance(res, tuple):
    return {"tuple": [one("out%d" % i, r) for i, r in enumerate(res)]}
  return one("out0", res)


def case(name, fn, args, kwargs=None, big=False):
  kwargs = kwargs or {}
  entry = {"name": name, "fn": fn, "args": [put_arg(name, "arg%d" % i, a) for i, a in enumerate(args)],
           "kwargs": kwargs}
  call_args = [a.make() if isinstance(a, Synth) else (np.copy(a, order="K") if isinstance(a, np.ndarray) else a) for a in args]
  try:
    res = getattr(ref, fn)(*call_args, **kwargs)
    entry["result"] = put_result(name, res, big=big)
  except Exception as e:  # expected failures are part of the contract
    entry["raises"] = {"type": type(e).__name__, "message": str(e)}
  manifest.append(entry)


def vol(dtype, seed, shape=(6, 7, 8), grid=(2, 3, 3), zero_frac=0.2, noise=0.1, small_ids=None):
  a = synth_numpy(shape, grid, dtype, seed=seed, zero_frac=zero_frac, noise=noise)
  if small_ids is not None:  # keep labels in a small non-negative range
    a = (a.astype(np.uint64) % np.uint64(small_ids)).astype(dtype)
  return a


# ---------------------------------------------------------------- renumber
k = 0
for dt in DTYPES:
  for order in "CF":
    for pz in (True, False):
      a = vol(dt, seed=k)
      if order == "F":
        a = np.asfortranarray(a)
      case("renumber_%s_%s_pz%d" % (dt, order, pz), "renumber", [a], {"preserve_zero": pz})
      k += 1
case("renumber_start0", "renumber", [vol("uint32", 100)], {"start": 0, "preserve_zero": False})
case("renumber_start0_pz", "renumber", [vol("uint32", 101)], {"start": 0, "preserve_zero": True})
case("renumber_start5", "renumber", [vol("int64", 102)], {"start": 5})
case("renumber_start300_u8", "renumber", [vol("uint8", 103)], {"start": 300})
case("renumber_start70000_u16", "renumber", [vol("uint16", 104)], {"start": 70000})
case("renumber_neg_start_i16", "renumber", [vol("int16", 105)], {"start": -3})
case("renumber_neg_start_i64", "renumber", [vol("int64", 106)], {"start": -300, "preserve_zero": False})
perm = np.argsort(synth_numpy((1, 1, 256), (1, 1, 256), "uint64", seed=7, zero_frac=0.0)[0, 0]).astype(np.uint8)
case("renumber_u8_all256_wrap", "renumber", [perm], {"preserve_zero": False})
case("renumber_u8_all256_pz", "renumber", [perm], {"preserve_zero": True})
case("renumber_1d_u64_big", "renumber", [np.array([5, 7, 10**12, 0, 2**64 - 1, 7, 2**63], dtype=np.uint64)], {})
case("renumber_i64_allones", "renumber", [np.array([-1, 3, -1, 0, -2**63, 3], dtype=np.int64)], {})
case("renumber_255_labels_u64", "renumber", [np.arange(1, 256, dtype=np.uint64)[::-1].copy()], {})
case("renumber_256_labels_u64", "renumber", [np.arange(1, 257, dtype=np.uint64)[::-1].copy()], {})
case("renumber_bool_pz", "renumber", [np.array([True, False, True])], {"preserve_zero": True, "start": 4})
case("renumber_bool_nopz", "renumber", [np.array([True, False, True])], {"preserve_zero": False})
case("renumber_noncontig", "renumber", [vol("uint32", 107, shape=(6, 8, 8))[:, ::2, :]], {})
bigspec = dict(shape=[64, 64, 64], grid=[50, 50, 50], dtype="uint64", seed=11, zero_frac=0.05, noise=0.02)
case("renumber_big_u64_100k", "renumber", [Synth(**bigspec)], {}, big=True)
case("renumber_big_u64_100k_F", "renumber", [Synth(order="F", **bigspec)], {"preserve_zero": False, "start": 3}, big=True)

# ---------------------------------------------------------------- remap / mask / mask_except
for i, dt in enumerate(DTYPES):
  a = vol(dt, seed=200 + i, small_ids=50)
  labels = [int(x) for x in np.unique(a)]
  full = {l: (l * 3 + 1) % 120 for l in labels}
  part = {l: (l * 5 + 2) % 120 for l in labels[::2]}
  case("remap_full_%s" % dt, "remap", [a, full], {"preserve_missing_labels": False})
  case("remap_part_keep_%s" % dt, "remap", [a, part], {"preserve_missing_labels": True})
  case("remap_part_err_%s" % dt, "remap", [a, part], {"preserve_missing_labels": False})
  case("remap_F_%s" % dt, "remap", [np.asfortranarray(a), part], {"preserve_missing_labels": True})
  case("remap_single_%s" % dt, "remap", [a, {labels[1]: 77}], {"preserve_missing_labels": True})
  case("remap_identity_%s" % dt, "remap", [a, {l: l for l in labels}], {"preserve_missing_labels": True})
  case("mask_%s" % dt, "mask", [a, labels[1:6]], {})
  case("mask_value_%s" % dt, "mask", [a, labels[1:6]], {"value": 99})
  case("mask_except_%s" % dt, "mask_except", [a, labels[2:9]], {})
  case("mask_except_max_%s" % dt, "mask_except", [a, labels[2:9]], {"value": int(np.iinfo(dt).max)})
  case("mask_except_empty_%s" % dt, "mask_except", [a, []], {"value": 7})
case("remap_widen_u8", "remap", [vol("uint8", 230, small_ids=20), {1: 1000, 2: 5}], {"preserve_missing_labels": True})
case("remap_widen_i8_neg", "remap", [vol("int8", 231, small_ids=20), {1: -1000, 2: 5}], {"preserve_missing_labels": True})
case("remap_neg_into_unsigned", "remap", [vol("uint8", 232, small_ids=20), {1: -1}], {"preserve_missing_labels": True})
case("remap_key_overflow", "remap", [vol("uint8", 233, small_ids=20), {300: 1, 1: 2}], {"preserve_missing_labels": True})
case("remap_nochain", "remap", [np.array([1, 2, 3, 1], dtype=np.uint16), {1: 2, 2: 3}], {"preserve_missing_labels": True})
case("remap_empty_table_err", "remap", [np.array([4, 4], dtype=np.uint16), {}], {"preserve_missing_labels": False})
case("remap_u64_huge", "remap", [np.array([2**64 - 1, 5, 2**63, 5], dtype=np.uint64), {2**64 - 1: 1, 2**63: 2**64 - 1}],
     {"preserve_missing_labels": True})
case("mask_except_overflow", "mask_except", [vol("uint8", 234, small_ids=20), [300]], {})

# ---------------------------------------------------------------- remap_from_array(_kv)
for i, dt in enumerate(DTYPES):
  a = vol(dt, seed=300 + i, small_ids=40).reshape(-1)
  if dt.startswith("u"):
    vals = ((np.arange(30) * 7 + 3) % 100).astype(dt)
    case("remap_from_array_%s" % dt, "remap_from_array", [a, vals], {"in_place": False})
  keys = np.array([1, 2, 3, 5, 8, 2, 13, 1], dtype=dt)
  vals = np.array([10, 20, 30, 50, 80, 21, 33, 11], dtype=dt)
  case("remap_kv_keep_%s" % dt, "remap_from_array_kv", [a, keys, vals], {"in_place": False})
  case("remap_kv_err_%s" % dt, "remap_from_array_kv", [a, keys, vals], {"preserve_missing_labels": False, "in_place": False})
  allk = np.unique(a)
  case("remap_kv_full_%s" % dt, "remap_from_array_kv", [a, allk, ((allk.astype(np.int64) * 2 + 1) % 100).astype(dt)],
       {"preserve_missing_labels": False, "in_place": False})

# ---------------------------------------------------------------- component_map
for i, (dc, dp) in enumerate([("uint8", "uint64"), ("uint64", "uint32"), ("int32", "int8"), ("uint16", "uint16"),
                              ("int64", "int64"), ("uint32", "int16")]):
  cc = vol(dc, seed=400 + i, small_ids=60)
  par = (cc.astype(np.int64) // 4 + 1).astype(dp)
  case("component_map_%s_%s" % (dc, dp), "component_map", [cc, par])
  case("component_map_F_%s_%s" % (dc, dp), "component_map", [np.asfortranarray(cc), np.asfortranarray(par)])
case("component_map_conflict", "component_map", [np.array([1, 1, 2, 1, 1], dtype=np.uint32), np.array([5, 6, 7, 8, 9], dtype=np.uint32)])
case("component_map_conflict_2d_F", "component_map",
     [np.asfortranarray(np.array([[1, 2, 1], [1, 1, 2]], dtype=np.uint64)), np.asfortranarray(np.array([[5, 6, 7], [8, 9, 10]], dtype=np.uint8))])
case("component_map_mixed_order", "component_map",
     [np.array([[1, 2, 1], [1, 1, 2]], dtype=np.uint64), np.asfortranarray(np.array([[5, 6, 7], [8, 9, 10]], dtype=np.uint8))])

# ---------------------------------------------------------------- unique / minmax / pixel_pairs / refit
for i, dt in enumerate(DTYPES):
  dense = vol(dt, seed=500 + i, small_ids=100)
  sparse = vol(dt, seed=520 + i)
  case("unique_dense_%s" % dt, "unique", [dense], {"return_counts": True})
  case("unique_sparse_%s" % dt, "unique", [sparse], {"return_counts": True})
  case("unique_sparse_F_%s" % dt, "unique", [np.asfortranarray(sparse)], {"return_counts": True})
  case("unique_nocounts_%s" % dt, "unique", [sparse], {})
  case("minmax_%s" % dt, "minmax", [sparse])
  case("pixel_pairs_%s" % dt, "pixel_pairs", [sparse])
  case("pixel_pairs_F_%s" % dt, "pixel_pairs", [np.asfortranarray(sparse)])
  case("refit_%s" % dt, "refit", [dense])
  case("refit_value_%s" % dt, "refit", [dense], {"value": 300})
if True:
  signed = (vol("int32", 540).astype(np.int64) % 1001 - 500).astype(np.int32)
  case("unique_shifted_i32", "unique", [signed], {"return_counts": True})
  case("unique_single", "unique", [np.array([777], dtype=np.int64)], {"return_counts": True})
  case("unique_empty", "unique", [np.array([], dtype=np.uint8)], {"return_counts": True})
  rnd = (synth_numpy((1, 50, 50), (1, 50, 50), "uint64", seed=541, zero_frac=0.0)).astype(np.uint64)
  case("unique_random_u64", "unique", [rnd], {"return_counts": True})
  case("unique_random_i64", "unique", [rnd.view(np.int64)], {"return_counts": True})
  case("unique_big_u32", "unique", [Synth(shape=[64, 64, 64], grid=[20, 20, 20], dtype="uint32", seed=542, zero_frac=0.1)],
       {"return_counts": True}, big=True)
  case("unique_big_u64_sparse", "unique", [Synth(shape=[64, 64, 64], grid=[30, 30, 30], dtype="uint64", seed=543, zero_frac=0.1, noise=0.05)],
       {"return_counts": True}, big=True)

# ---------------------------------------------------------------- unique: index / inverse / axis=0
for i, dt in enumerate(DTYPES):
  dense = vol(dt, seed=600 + i, small_ids=100)
  sparse = vol(dt, seed=620 + i)
  allf = {"return_index": True, "return_inverse": True, "return_counts": True}
  case("unique_all_dense_%s" % dt, "unique", [dense], allf)
  case("unique_all_dense_F_%s" % dt, "unique", [np.asfortranarray(dense)], allf)
  case("unique_all_sparse_%s" % dt, "unique", [sparse], allf)
  case("unique_all_sparse_F_%s" % dt, "unique", [np.asfortranarray(sparse)], allf)
  case("unique_index_%s" % dt, "unique", [sparse], {"return_index": True})
  case("unique_inverse_F_%s" % dt, "unique", [np.asfortranarray(sparse)], {"return_inverse": True})
  if np.dtype(dt).itemsize < 8:
    info = np.iinfo(dt)
    edges = (synth_numpy((1, 300, 2), (1, 300, 2), "uint64", seed=640 + i, zero_frac=0.0) % 23).astype(np.int64)
    edges = (edges + max(int(info.min), -11)).astype(dt).reshape(300, 2)
    case("unique_axis0_%s" % dt, "unique", [edges], {"axis": 0})
if True:
  rnd = (synth_numpy((9, 11, 13), (9, 11, 13), "uint64", seed=660, zero_frac=0.0)).astype(np.uint64)  # iid: the numpy branch (:947-948)
  allf = {"return_index": True, "return_inverse": True, "return_counts": True}
  case("unique_all_random_u64", "unique", [rnd], allf)
  case("unique_all_random_u64_F", "unique", [np.asfortranarray(rnd)], allf)
  case("unique_all_random_i64", "unique", [rnd.view(np.int64)], allf)
  case("unique_all_single", "unique", [np.array([[777]], dtype=np.int64)], allf)
  case("unique_all_empty", "unique", [np.array([], dtype=np.uint8)], allf)

# ---------------------------------------------------------------- foreground / indices ("next" rows)
for i, dt in enumerate(DTYPES):
  sparse = vol(dt, seed=700 + i)
  case("foreground_%s" % dt, "foreground", [sparse])
  case("foreground_F_%s" % dt, "foreground", [np.asfortranarray(sparse)])
  flat = vol(dt, seed=720 + i, small_ids=20).reshape(-1)
  case("indices_%s" % dt, "indices", [flat, int(flat[len(flat) // 2])])
  case("indices_absent_%s" % dt, "indices", [flat, 21])
case("indices_overflow", "indices", [np.array([1, 2, 3], dtype=np.uint8), 300])
case("indices_empty", "indices", [np.array([], dtype=np.int32), 0])

# ---------------------------------------------------------------- layout changes ("next" rank 4)
LAYOUT_DTYPES = ["uint8", "uint16", "uint32", "uint64", "int32", "float32", "float64", "bool", "complex64"]


def ramp(shape, dt, order):
  n = int(np.prod(shape))
  a = (np.arange(n, dtype=np.int64) * 7 + 3) % 251
  if dt == "complex64":
    a = (a + 1j * (a[::-1] % 17)).astype(np.complex64)
  else:
    a = (a % 2).astype(bool) if dt == "bool" else a.astype(dt)
  return np.copy(a.reshape(shape), order=order)


for dt in LAYOUT_DTYPES:
  for shape in [(4, 6), (4, 5, 6), (3, 4, 2, 5)]:
    tag = "%s_%s" % (dt, "x".join(map(str, shape)))
    case("asfortranarray_%s" % tag, "asfortranarray", [ramp(shape, dt, "C")])
    case("ascontiguousarray_%s" % tag, "ascontiguousarray", [ramp(shape, dt, "F")])
    case("transpose_C_%s" % tag, "transpose", [ramp(shape, dt, "C")])
    case("transpose_F_%s" % tag, "transpose", [ramp(shape, dt, "F")])
for shape in [(7,), (1, 9), (9, 1), (5, 5), (70, 33), (6, 6, 6), (1, 8, 5), (33, 2, 65), (3, 3, 3, 3), (2, 1, 3, 4), (2, 3, 2, 2, 3)]:
  tag = "u2_%s" % "x".join(map(str, shape))
  case("asfortranarray_%s" % tag, "asfortranarray", [ramp(shape, "uint16", "C")])
  case("asfortranarray_noop_%s" % tag, "asfortranarray", [ramp(shape, "uint16", "F")])
  case("ascontiguousarray_%s" % tag, "ascontiguousarray", [ramp(shape, "uint16", "F")])
  case("transpose_C_%s" % tag, "transpose", [ramp(shape, "uint16", "C")])
  case("transpose_F_%s" % tag, "transpose", [ramp(shape, "uint16", "F")])
case("asfortranarray_f16", "asfortranarray", [ramp((3, 4), "float16", "C")])

for dt in ["uint8", "uint16", "uint32", "uint64", "int16", "float32", "float64"]:
  for io in "CF":
    for oo in "CF":
      case("tobytes_%s_%s%s" % (dt, io, oo), "tobytes", [ramp((8, 12, 4), dt, io), [4, 3, 2]], {"order": oo})
for io in "CF":
  for oo in "CF":
    case("tobytes_whole_%s%s" % (io, oo), "tobytes", [ramp((5, 6, 7), "int32", io), [5, 6, 7]], {"order": oo})
    case("tobytes_slabs_%s%s" % (io, oo), "tobytes", [ramp((6, 4, 70), "uint8", io), [3, 4, 35]], {"order": oo})
    case("tobytes_unit_%s%s" % (io, oo), "tobytes", [ramp((3, 2, 4), "uint16", io), [1, 1, 1]], {"order": oo})
case("tobytes_misaligned", "tobytes", [ramp((8, 8, 8), "uint8", "C"), [3, 3, 3]])
case("tobytes_bad_order", "tobytes", [ramp((8, 8, 8), "uint8", "C"), [4, 4, 4]], {"order": "K"})
case("tobytes_bool", "tobytes", [ramp((4, 4, 4), "bool", "C"), [2, 2, 2]])

np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
with open(os.path.join(HERE, "golden_manifest.json"), "w") as f:
  json.dump(manifest, f, indent=1, sort_keys=True)
print("cases:", len(manifest), "arrays:", len(arrays),
      "npz bytes:", os.path.getsize(os.path.join(HERE, "golden.npz")))
