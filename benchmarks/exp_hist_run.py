#!/usr/bin/env python
"""unique(return_counts) on the 1024^3 uint32 configs for several tile-run lengths (frb_set_tile_run): runs that are
not a multiple of the volume's cell period take the CTAs out of lockstep."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import _lib, api  # noqa: E402
from fastremap_b200 import device as D  # noqa: E402
from fastremap_b200.synth import synth_device  # noqa: E402
from benchmarks.bench_ops import timed, flat_u8  # noqa: E402

S = 1024
n = S ** 3
D.dev()
L = _lib.load()
runs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 61, 67, 31, 127, 16]
sparse = flat_u8(synth_device((S, S, S), (100,) * 3, np.uint32, seed=1, zero_frac=0.1))
vol = synth_device((S, S, S), (46,) * 3, np.uint64, seed=0, zero_frac=0.1)
dense = api.renumber_device(flat_u8(vol).clone(), n, np.uint64).out
del vol
for r in runs:
  L.frb_set_tile_run(r)
  ms, mn, out = timed(lambda: api.unique_counts_device(sparse, n, np.uint32, row_bytes=S * 4, plane_bytes=S * S * 4), 5)
  md, mnd, outd = timed(lambda: api.unique_counts_device(dense, n, np.uint32, row_bytes=S * 4, plane_bytes=S * S * 4), 5)
  print(json.dumps({"tile_run": r, "unique_sparse_ms": ms, "unique_dense_ms": md, "n_sparse": out[2], "n_dense": outd[2]}), flush=True)
