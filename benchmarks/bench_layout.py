#!/usr/bin/env python
"""
Device-resident timing of the layout kernels (frb_transpose_tiles / frb_copy_rows): C <-> F
transposition of a volume and the tobytes chunk rearrangement, per element width.  One JSON line per
case: ms, algorithmic GB/s (one read + one write per element) and the fraction of the measured HBM peak.

  python benchmarks/bench_layout.py [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from fastremap_b200 import device as D  # noqa: E402
from fastremap_b200 import layout  # noqa: E402
from bench_ops import peak, timed  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--reps", type=int, default=5)
  args = ap.parse_args()
  D.dev()
  g = torch.Generator(device="cuda").manual_seed(0)
  cases = [("1024x1024x1024", (1024, 1024, 1024), 1), ("1024x1024x1024", (1024, 1024, 1024), 2), ("1024x1024x512", (1024, 1024, 512), 4),
           ("1024x1024x512", (1024, 1024, 512), 8), ("1000x1000x500", (1000, 1000, 500), 8), ("16384x16384", (16384, 16384), 4),
           ("512x512x64x16", (512, 512, 64, 16), 4), ("512x512x128x3", (512, 512, 128, 3), 4)]
  for name, shape, w in cases:
    n = int(np.prod(shape))
    buf = torch.randint(0, 255, (n * w,), dtype=torch.uint8, device="cuda", generator=g)
    m, mn, out = timed(lambda: layout.reverse_axes_device(buf, tuple(reversed(shape)), w), args.reps)
    gbs = 2 * w * n / (m / 1e3) / 1e9
    print(json.dumps({"op": "transpose C->F %s width %d" % (name, w), "ms": m, "ms_min": mn, "alg_GB/s": gbs,
                      "frac_of_measured_hbm_peak": gbs / peak()}), flush=True)
    del out
    if len(shape) == 4:  # the way back: F -> C
      m, mn, out = timed(lambda: layout.reverse_axes_device(buf, shape, w), args.reps)
      gbs = 2 * w * n / (m / 1e3) / 1e9
      print(json.dumps({"op": "transpose F->C %s width %d" % (name, w), "ms": m, "ms_min": mn, "alg_GB/s": gbs,
                        "frac_of_measured_hbm_peak": gbs / peak()}), flush=True)
      del out
    if len(shape) == 3:
      for io, oo, chunk in [("C", "C", (64, 64, 64)), ("F", "F", (128, 128, 64)), ("C", "F", (64, 64, 64)), ("F", "C", (128, 128, 64))]:
        if any(s % c for s, c in zip(shape, chunk)):
          continue
        m, mn, out = timed(lambda: layout.chunk_layout_device(buf, shape, io, chunk, oo, w), args.reps)
        gbs = 2 * w * n / (m / 1e3) / 1e9
        print(json.dumps({"op": "tobytes rearrangement %s width %d %s->%s chunk %s" % (name, w, io, oo, "x".join(map(str, chunk))),
                          "ms": m, "ms_min": mn, "alg_GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak()}), flush=True)
        del out
    del buf
    torch.cuda.empty_cache()


if __name__ == "__main__":
  main()
