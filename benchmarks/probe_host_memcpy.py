#!/usr/bin/env python
"""Host memcpy scaling with threads (pageable -> page-locked), and the driver's own pageable H2D rate."""
import ctypes, json, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
n = 1 << 30
src = np.ones(n, dtype=np.uint8)
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dst = pin.numpy()
sp, dp = src.ctypes.data, dst.ctypes.data


def run(T, piece=32 << 20):
  pool = ThreadPoolExecutor(T)
  t = time.perf_counter()
  for a in range(0, n, piece):
    b = min(n, a + piece)
    per = (b - a + T - 1) // T
    list(pool.map(lambda k: ctypes.memmove(dp + a + k * per, sp + a + k * per, max(0, min(b, a + (k + 1) * per) - (a + k * per))), range(T)))
  dt = time.perf_counter() - t
  pool.shutdown()
  return n / dt / 1e9


for T in (1, 2, 4, 6, 8, 12, 16):
  run(T)
  print(json.dumps({"threads": T, "pageable_to_pinned_GBps": round(run(T), 1)}), flush=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, h in (("pageable", torch.from_numpy(src)), ("pinned", pin)):
  for _ in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
  print(json.dumps({"h2d_from": name, "GBps": round(n / dt / 1e9, 1)}), flush=True)
  for _ in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
  print(json.dumps({"d2h_to": name, "GBps": round(n / dt / 1e9, 1)}), flush=True)
t = time.perf_counter()
x = torch.empty(128 << 20, dtype=torch.uint8, pin_memory=True)
print(json.dumps({"pin_alloc_128MiB_s": round(time.perf_counter() - t, 4)}), flush=True)
