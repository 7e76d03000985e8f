#!/usr/bin/env python
"""Small driver for ncu captures: runs renumber (K1,K2,K3) and unique (hash strategy) twice on a synthetic volume."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastremap_b200 import api  # noqa: E402
from fastremap_b200 import device as D  # noqa: E402
from fastremap_b200.synth import synth_device  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = S ** 3
D.dev()
vol = synth_device((S, S, S), (46 * S // 1024,) * 3, np.uint64, seed=0, zero_frac=0.1).reshape(-1).view(torch.uint8)
work = torch.empty_like(vol)
for _ in range(2):
  work.copy_(vol)
  res = api.renumber_device(work, n, np.uint64, write_wide=True)
dense32 = res.out
for _ in range(2):
  api.unique_counts_device(dense32, n, np.uint32, strategy="hash")
torch.cuda.synchronize()
print("ok", res.n_labels)
