#!/usr/bin/env python
"""Driver for a launch list of config 1: renumber_device of the 512^3 uint64 volume, three times."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import api
from fastremap_b200 import device as D
from fastremap_b200.synth import synth_device
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = S ** 3
D.dev()
vol = synth_device((S, S, S), (46, 46, 46), np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
work = torch.empty_like(vol)
for _ in range(3):
  work.copy_(vol)
  res = api.renumber_device(work, n, np.uint64, write_wide=True, row_bytes=S * 8)
torch.cuda.synchronize()
print("labels", res.n_labels)
