#!/usr/bin/env python
"""Feasibility probe for an L2-resident chunked renumber: per chunk, K1 (build) followed by a
read + wide write + narrow write pass over the SAME chunk (the traffic shape of K3), all chunks queued
on one stream (optionally captured in a CUDA graph).  Compares with the two full passes."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device

S = 1024
n = S ** 3
L = _lib.load()
D.dev()
vol = synth_device((S, S, S), (46, 46, 46), np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
work = vol.clone()
narrow = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
out = torch.zeros(1, dtype=torch.int64, device="cuda")
table = api.LabelTable(np.uint64, n // 256, _lib.EMPTY)
c = D.code(np.uint64)

def run_chunks(chunk_bytes, use_probe_only=False):
  st = D.stream()
  for a in range(0, n * 8, chunk_bytes):
    b = min(a + chunk_bytes, n * 8)
    ne = (b - a) // 8
    src = ctypes.c_void_p(work.data_ptr() + a)
    nar = ctypes.c_void_p(narrow.data_ptr() + a // 2)
    if not use_probe_only:
      _lib.check(L.frb_renumber_build(src, ne, c, a // 8, 1, D.ptr(table.slots), table.cap, D.ptr(table.meta), S * 8, st))
    else:
      _lib.check(L.frb_stream_probe(src, b - a, None, None, 0, D.ptr(out), st))
    _lib.check(L.frb_stream_probe(src, b - a, src, nar, 2, D.ptr(out), st))

def timed(fn, reps=4):
  ms = []
  for i in range(reps + 1):
    work.copy_(vol); table.clear(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    if i: ms.append(e0.elapsed_time(e1))
  return min(ms)

print(json.dumps({"case": "two full passes (K1 + probe2)", "ms": timed(lambda: run_chunks(n * 8))}), flush=True)
for mib in (16, 24, 32, 48, 64, 128):
  for probe_only in (False, True):
    cb = mib << 20
    t_eager = timed(lambda: run_chunks(cb, probe_only))
    g = torch.cuda.CUDAGraph()
    work.copy_(vol); table.clear(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
      run_chunks(cb, probe_only)
    t_graph = timed(lambda: g.replay())
    print(json.dumps({"case": "chunked", "chunk_MiB": mib, "first_pass": "read probe" if probe_only else "K1", "eager_ms": t_eager, "graph_ms": t_graph}), flush=True)
