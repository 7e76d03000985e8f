"""Replay of the committed golden cases (tests/golden/) against any implementation module."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load():
  if not _cache:
    with open(os.path.join(HERE, "golden", "golden_manifest.json")) as f:
      _cache["manifest"] = json.load(f)
    _cache["npz"] = np.load(os.path.join(HERE, "golden", "golden.npz"))
  return _cache["manifest"], _cache["npz"]


def case_names(fns=None):
  manifest, _ = load()
  return [e["name"] for e in manifest if fns is None or e["fn"] in fns]


def get_case(name):
  manifest, _ = load()
  for e in manifest:
    if e["name"] == name:
      return e
  raise KeyError(name)


def digest(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _as_order(a, order):
  return np.asfortranarray(a) if order == "F" else np.ascontiguousarray(a)


def build_args(entry):
  _, npz = load()
  from fastremap_b200.synth import synth_numpy
  out = []
  for spec in entry["args"]:
    if "synth" in spec:
      out.append(_as_order(synth_numpy(**spec["synth"]), spec["order"]))
    elif "array" in spec:
      out.append(_as_order(np.array(npz[spec["array"]]), spec["order"]))
    elif "dict" in spec:
      ks, vs = npz[spec["dict"] + ".k"].tolist(), npz[spec["dict"] + ".v"].tolist()
      out.append(dict(zip(ks, vs)))
    elif "list" in spec:
      out.append(list(spec["list"]))
    else:
      out.append(spec["value"])
  return out


def _check_one(spec, got, what):
  _, npz = load()
  if "array" in spec:
    exp = npz[spec["array"]]
    got = np.asarray(got.cpu().numpy() if hasattr(got, "cpu") else got)
    assert got.dtype == np.dtype(spec["dtype"]), "%s: dtype %s != %s" % (what, got.dtype, spec["dtype"])
    assert got.shape == exp.shape, "%s: shape %s != %s" % (what, got.shape, exp.shape)
    assert np.array_equal(got, exp), "%s: values differ" % what
    if got.ndim > 1 and spec.get("order") == "F":
      assert got.flags.f_contiguous, "%s: expected an F-ordered result" % what
    elif got.ndim > 1:
      assert got.flags.c_contiguous, "%s: expected a C-ordered result" % what
  elif "digest" in spec:
    got = np.asarray(got)
    assert got.dtype == np.dtype(spec["dtype"]), "%s: dtype %s != %s" % (what, got.dtype, spec["dtype"])
    assert list(got.shape) == spec["shape"], "%s: shape" % what
    assert digest(got) == spec["digest"], "%s: digest differs" % what
  elif "dict" in spec:
    ks, vs = npz[spec["dict"] + ".k"].tolist(), npz[spec["dict"] + ".v"].tolist()
    assert isinstance(got, dict), "%s: expected a dict" % what
    assert got == dict(zip(ks, vs)), "%s: dict differs" % what
  elif "dict_digest" in spec:
    assert isinstance(got, dict) and len(got) == spec["len"], "%s: dict length" % what
    ks = sorted(got.keys())
    ka = np.array(ks, dtype=spec["kdtype"])
    va = np.array([got[k] for k in ks], dtype=spec["vdtype"])
    assert [digest(ka), digest(va)] == spec["dict_digest"], "%s: dict digest differs" % what
  elif "bytes_list" in spec:
    assert isinstance(got, list) and all(isinstance(b, bytes) for b in got), "%s: expected list[bytes]" % what
    assert [len(b) for b in got] == spec["lengths"], "%s: chunk lengths differ" % what
    assert b"".join(got) == npz[spec["bytes_list"]].tobytes(), "%s: bytes differ" % what
  else:
    assert got == spec["value"], "%s: %r != %r" % (what, got, spec["value"])


def check_case(entry, impl):
  """Call impl.<fn>(*args, **kwargs) and compare with the reference's recorded outcome."""
  args = build_args(entry)
  fn = getattr(impl, entry["fn"])
  if "raises" in entry:
    exc = {"KeyError": KeyError, "ValueError": ValueError, "OverflowError": OverflowError, "TypeError": TypeError}[
      entry["raises"]["type"]]
    try:
      fn(*args, **entry["kwargs"])
    except exc as e:
      if exc in (KeyError, ValueError):
        assert str(e) == entry["raises"]["message"], "%s: message %r != %r" % (entry["name"], str(e), entry["raises"]["message"])
      return
    raise AssertionError("%s: expected %s" % (entry["name"], entry["raises"]["type"]))
  got = fn(*args, **entry["kwargs"])
  spec = entry["result"]
  if "tuple" in spec:
    assert isinstance(got, tuple) and len(got) == len(spec["tuple"]), "%s: tuple arity" % entry["name"]
    for i, (s, g) in enumerate(zip(spec["tuple"], got)):
      _check_one(s, g, "%s[%d]" % (entry["name"], i))
  else:
    _check_one(spec, got, entry["name"])
