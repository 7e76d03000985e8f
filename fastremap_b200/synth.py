"""
Synthetic "connectomics-like" label volumes: a coarse grid of cells with
splitmix64 ids, nearest-neighbour upsampled.  Two bit-identical forms: `synth_numpy` on the host
(for the CPU reference / oracle) and `synth_device` (CUDA kernel frb_synth_volume, no PCIe traffic).
Both are functions of (seed, global voxel index) only, so a z-slab generated on one rank equals
the same planes of the full volume.
"""
from __future__ import annotations

import numpy as np

_M = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
  with np.errstate(over="ignore"):
    z = x + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _mix(seed: int, idx: np.ndarray) -> np.ndarray:
  with np.errstate(over="ignore"):
    return _splitmix64(np.uint64((seed * 0xD1342543DE82EF95) & _M) + idx.astype(np.uint64))


def _idmask(dtype) -> int:
  dtype = np.dtype(dtype)
  bits = 8 * dtype.itemsize
  idbits = bits - 1 if np.issubdtype(dtype, np.signedinteger) else min(bits, 63)
  return (1 << idbits) - 1


def synth_numpy(shape, grid, dtype, seed=0, zero_frac=0.1, noise=0.0, z0=0, nz_total=None, cell_div=1) -> np.ndarray:
  """(nz, ny, nx) C-order volume; planes [z0, z0+nz) of a (nz_total, ny, nx) volume."""
  nz, ny, nx = (int(s) for s in shape)
  gz, gy, gx = (int(g) for g in grid)
  nz_total = int(nz_total) if nz_total is not None else nz
  dtype = np.dtype(dtype)
  zg = np.arange(nz, dtype=np.uint64) + np.uint64(z0)
  y = np.arange(ny, dtype=np.uint64)
  x = np.arange(nx, dtype=np.uint64)
  cz = (zg * np.uint64(gz)) // np.uint64(nz_total)
  cy = (y * np.uint64(gy)) // np.uint64(ny)
  cx = (x * np.uint64(gx)) // np.uint64(nx)
  c = (cz[:, None, None] * np.uint64(gy) + cy[None, :, None]) * np.uint64(gx) + cx[None, None, :]
  zero_thresh = int(zero_frac * 16777216.0)
  noise_thresh = int(noise * 16777216.0)
  if not noise_thresh and nz * ny * nx > 4 * gz * gy * gx:
    # no per-voxel noise: a voxel's label is its cell's label, so hash the cells once and upsample;
    # z-planes inside one layer of cells are copies of each other
    table = synth_cell_labels(grid, dtype, seed, zero_frac, cell_div)
    out = np.empty((nz, ny, nx), dtype=dtype)
    iy, ix = cy.astype(np.intp)[:, None], cx.astype(np.intp)[None, :]
    prev, plane = None, None
    for z in range(nz):
      layer = int(cz[z])
      if layer != prev:
        plane, prev = table[layer][iy, ix], layer
      out[z] = plane
    return out
  if noise_thresh:
    flat = (zg[:, None, None] * np.uint64(ny) + y[None, :, None]) * np.uint64(nx) + x[None, None, :]
    hn = _mix(seed + 2, flat)
    hit = (hn >> np.uint64(40)) < np.uint64(noise_thresh)
    ncells = np.uint64(gz * gy * gx)
    c = np.where(hit, (c + np.uint64(1) + (hn & np.uint64(7))) % ncells, c)
  c = c // np.uint64(cell_div)
  label = _mix(seed, c) & np.uint64(_idmask(dtype))
  label = np.where(label == 0, np.uint64(1), label)
  if zero_thresh:
    label = np.where((_mix(seed + 1, c) >> np.uint64(40)) < np.uint64(zero_thresh), np.uint64(0), label)
  return np.ascontiguousarray(label.astype("u%d" % dtype.itemsize).view(dtype))


def synth_cell_labels(grid, dtype, seed=0, zero_frac=0.1, cell_div=1) -> np.ndarray:
  """(gz, gy, gx) array: the label every voxel of cell (cz, cy, cx) carries in a noise-free volume."""
  gz, gy, gx = (int(g) for g in grid)
  return synth_numpy((gz, gy, gx), (gz, gy, gx), dtype, seed=seed, zero_frac=zero_frac, noise=0.0, cell_div=cell_div)


def synth_device(shape, grid, dtype, seed=0, zero_frac=0.1, noise=0.0, z0=0, nz_total=None, cell_div=1):
  """Same volume generated on the current CUDA device; returns a torch tensor of shape (nz, ny, nx)."""
  from . import _lib
  from . import device as D
  nz, ny, nx = (int(s) for s in shape)
  gz, gy, gx = (int(g) for g in grid)
  nz_total = int(nz_total) if nz_total is not None else nz
  dtype = np.dtype(dtype)
  buf = D.alloc(nz * ny * nx * dtype.itemsize)
  _lib.check(_lib.load().frb_synth_volume(D.ptr(buf), D.code(dtype), nz, ny, nx, int(z0), nz_total, gz, gy, gx,
                                          int(seed) & _M, float(zero_frac), float(noise), int(cell_div), D.stream()),
             "frb_synth_volume")
  return buf[: nz * ny * nx * dtype.itemsize].view(D.torch_dtype(dtype)).reshape(nz, ny, nx)
