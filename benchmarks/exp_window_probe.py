#!/usr/bin/env python
"""One pass through L2?  Times frb_stream_probe's windowed mode (ONE persistent cooperative kernel: per window a read pass,
a grid barrier, a read + 8 B + 4 B write pass, a grid barrier) against the two full passes it would replace, on 8 GiB.
  FRB_PROBE_WINDOW_MIB=32 python benchmarks/exp_window_probe.py [inplace]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device

S = 1024
n = S ** 3
L = _lib.load()
D.dev()
vol = synth_device((S, S, S), (46, 46, 46), np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
inplace = len(sys.argv) > 1 and sys.argv[1] == "inplace"
wide = vol if inplace else torch.empty_like(vol)
narrow = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
out = torch.zeros(2, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
  ms = []
  for i in range(reps + 2):
    flush.fill_(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    if i >= 2: ms.append(e0.elapsed_time(e1))
  return min(ms), float(np.mean(ms))


def two_pass():
  _lib.check(L.frb_stream_probe(D.ptr(vol), n * 8, None, None, 0, D.ptr(out), D.stream()))
  _lib.check(L.frb_stream_probe(D.ptr(vol), n * 8, D.ptr(wide), D.ptr(narrow), 2, D.ptr(out), D.stream()))


def windowed():
  _lib.check(L.frb_stream_probe(D.ptr(vol), n * 8, D.ptr(wide), D.ptr(narrow), 8 | 2, D.ptr(out), D.stream()))


mib = int(os.environ.get("FRB_PROBE_WINDOW_MIB", "64"))
mn, mean = timed(two_pass)
print(json.dumps({"case": "two full passes (read probe + read/write probe)", "in_place": inplace, "ms_min": mn, "ms": mean}), flush=True)
mn, mean = timed(windowed)
print(json.dumps({"case": "windowed persistent kernel", "window_MiB": mib, "in_place": inplace, "ms_min": mn, "ms": mean,
                  "alg_GB/s_at_20B_per_voxel": 20 * n / (mean / 1e3) / 1e9}), flush=True)
