#!/usr/bin/env python
"""ncu driver: renumber of a 1024^3 uint64 volume with ~720k labels (the K2 load of the 8-GPU merge)."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
S = 1024
n = S ** 3
D.dev()
vol = synth_device((S, S, S), (93, 93, 93), np.uint64, seed=3, zero_frac=0.1).reshape(-1).view(torch.uint8)
work = torch.empty_like(vol)
for _ in range(3):
  work.copy_(vol)
  timers = {}
  res = api.renumber_device(work, n, np.uint64, write_wide=True, timers=timers)
  torch.cuda.synchronize()
  print("labels", res.n_labels, {k: round(a.elapsed_time(b), 4) for k, (a, b) in timers.items()})
