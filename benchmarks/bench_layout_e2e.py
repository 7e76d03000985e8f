#!/usr/bin/env python
"""End-to-end (host array in, host array out) timing of asfortranarray / tobytes against numpy's own
out-of-place equivalents on the same host (context only: numpy is not the reference's algorithm)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import fastremap_b200 as fr


def best_of(fn, setup, reps=3):
  best = None
  for i in range(reps + 1):
    arg = setup()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(arg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if i:
      best = dt if best is None else min(best, dt)
  return best


for shape, dtype in (((512, 512, 512), np.float32), ((1024, 1024, 1024), np.uint8), ((1024, 1024, 512), np.uint64)):
  n = int(np.prod(shape))
  pinned = torch.empty(n * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
  base = pinned.numpy().view(dtype).reshape(shape)
  src = (np.arange(n, dtype=np.uint32) * 2654435761 >> 7).astype(dtype).reshape(shape)

  def fresh():
    base[...] = src
    return base
  t_gpu = best_of(fr.asfortranarray, fresh)
  t_np = best_of(np.asfortranarray, fresh, reps=1)
  print(json.dumps({"op": "asfortranarray %s %s" % ("x".join(map(str, shape)), np.dtype(dtype).name), "fastremap_b200_s": t_gpu,
                    "numpy_s": t_np, "GB": n * np.dtype(dtype).itemsize / 1e9}), flush=True)
  if np.dtype(dtype).itemsize == 1:
    t_gpu = best_of(lambda a: fr.tobytes(a, (64, 64, 64), order="C"), fresh)
    def np_tobytes(a):
      return [a[x:x + 64, y:y + 64, z:z + 64].tobytes("C") for z in range(0, a.shape[2], 64) for y in range(0, a.shape[1], 64) for x in range(0, a.shape[0], 64)]
    t_np = best_of(np_tobytes, fresh, reps=1)
    print(json.dumps({"op": "tobytes %s %s chunks 64^3 C->C" % ("x".join(map(str, shape)), np.dtype(dtype).name), "fastremap_b200_s": t_gpu,
                      "numpy_s": t_np}), flush=True)
  del pinned, base, src
