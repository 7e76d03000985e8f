"""Build libfastremap_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
SO = os.path.join(HERE, "libfastremap_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
EXTRA = os.environ.get("FRB_CFLAGS", "").split()          # experiments: e.g. -DFRB_STREAM_ROWS=8
if EXTRA:
  tag = "_" + "_".join(f.replace("-D", "").replace("=", "") for f in EXTRA)
  OBJ = os.path.join(HERE, "csrc", "_obj" + tag)
  SO = os.path.join(HERE, "..", "build_variants", "libfastremap_b200%s.so" % tag)
  os.makedirs(os.path.dirname(SO), exist_ok=True)
  FLAGS = FLAGS + EXTRA


def sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
  hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
  hdrs.append(os.path.join(HERE, "..", "include", "fastremap_b200.h"))
  return max(os.path.getmtime(h) for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OBJ, exist_ok=True)
  dep = _deps_mtime()
  jobs = []
  objs = []
  for src in sources():
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep):
      cmd = [NVCC] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
      jobs.append(cmd)

  def run(cmd):
    p = subprocess.run(cmd, capture_output=True, text=True)
    return cmd, p.returncode, p.stdout + p.stderr

  if jobs:
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
      for cmd, rc, out in ex.map(run, jobs):
        if verbose or rc != 0:
          sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        if rc != 0:
          raise RuntimeError("nvcc failed for %s" % cmd[-3])
  if jobs or force or not os.path.exists(SO):
    cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", SO] + objs
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
      sys.stderr.write(p.stdout + p.stderr)
      raise RuntimeError("link failed")
  return SO


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
