"""
N>1 host logic on CPU: world_size-2 gloo run of the sharded-renumber exchange
(fastremap_b200/sharded.py: shard_offsets + allgather_varlen + merge-by-min + rank), with the
per-shard kernels stood in by numpy, checked against the oracle's renumber of the whole volume.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, tmpdir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  sys.path.insert(0, ROOT)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from fastremap_b200 import sharded
    from fastremap_b200.synth import synth_numpy
    nz_total, ny, nx = 24, 20, 28
    bounds = [0, 10, 24]  # unequal slabs: offsets must come from the exchanged counts
    z0, z1 = bounds[rank], bounds[rank + 1]
    slab = synth_numpy((z1 - z0, ny, nx), (5, 4, 6), np.uint64, seed=3, zero_frac=0.2, noise=0.1, z0=z0, nz_total=nz_total)
    flat = slab.reshape(-1)
    offset, total = sharded.shard_offsets(flat.size)
    assert total == nz_total * ny * nx and offset == z0 * ny * nx

    # stand-in for K1 + export: local (label, first global index), label 0 left out (preserve_zero)
    keys, first = np.unique(flat, return_index=True)
    keep = keys != 0
    keys, first = keys[keep], first[keep] + offset
    gk = sharded.allgather_varlen(torch.from_numpy(keys.view(np.int64)))
    gv = sharded.allgather_varlen(torch.from_numpy(first.astype(np.int64)))
    assert [t.numel() for t in gk] == [t.numel() for t in gv]
    # stand-in for the merge (atomicMin per label) + K2 (rank by first index)
    allk = np.concatenate([t.numpy().view(np.uint64) for t in gk])
    allv = np.concatenate([t.numpy() for t in gv])
    order = np.lexsort((allv, allk))
    allk, allv = allk[order], allv[order]
    lead = np.ones(allk.size, dtype=bool)
    lead[1:] = allk[1:] != allk[:-1]
    mk, mv = allk[lead], allv[lead]          # min first index per label
    rank_order = np.argsort(mv, kind="stable")
    ids = np.empty(mk.size, dtype=np.uint64)
    ids[rank_order] = np.arange(1, mk.size + 1, dtype=np.uint64)
    mapping = dict(zip(mk.tolist(), ids.tolist()))
    mapping[0] = 0
    # stand-in for K3 on the own slab
    out = np.vectorize(mapping.get, otypes=[np.uint64])(flat)
    np.save(os.path.join(tmpdir, "out_%d.npy" % rank), out)
    if rank == 0:
      np.save(os.path.join(tmpdir, "map_k.npy"), np.array(sorted(mapping), dtype=np.uint64))
      np.save(os.path.join(tmpdir, "map_v.npy"), np.array([mapping[k] for k in sorted(mapping)], dtype=np.uint64))
  finally:
    dist.destroy_process_group()


def test_sharded_exchange_world2(oracle, tmp_path):
  port = _free_port()
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  from fastremap_b200.synth import synth_numpy
  full = synth_numpy((24, 20, 28), (5, 4, 6), np.uint64, seed=3, zero_frac=0.2, noise=0.1)
  e_out, e_map = oracle.renumber(np.copy(full))
  got = np.concatenate([np.load(tmp_path / ("out_%d.npy" % r)) for r in range(2)]).reshape(full.shape)
  assert np.array_equal(got.astype(e_out.dtype), e_out)
  mk, mv = np.load(tmp_path / "map_k.npy"), np.load(tmp_path / "map_v.npy")
  assert dict(zip(mk.tolist(), mv.tolist())) == e_map


def test_allgather_varlen_handles_empty_rank(tmp_path):
  port = _free_port()
  mp.spawn(_empty_worker, args=(2, port), nprocs=2, join=True)


def _empty_worker(rank, world, port):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  sys.path.insert(0, ROOT)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from fastremap_b200 import sharded
    t = torch.arange(5 * rank, dtype=torch.int64)  # rank 0 contributes nothing
    parts = sharded.allgather_varlen(t)
    assert [p.numel() for p in parts] == [0, 5]
    assert parts[1].tolist() == [0, 1, 2, 3, 4]
    ks, vs = sharded.allgather_pairs(t, t * 10)
    assert [k.tolist() for k in ks] == [[], [0, 1, 2, 3, 4]] and vs[1].tolist() == [0, 10, 20, 30, 40]
  finally:
    dist.destroy_process_group()


def _reduce_worker(rank, world, port):
  """unique (sum merge), component_map (last shard wins), KeyError status and minmax / pixel_pairs
  reductions of fastremap_b200/sharded.py with the per-shard kernels stood in by numpy."""
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  sys.path.insert(0, ROOT)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from fastremap_b200 import sharded
    rng = np.random.default_rng(5)
    full = rng.integers(-40, 40, size=3000).astype(np.int64)
    full[1000:1500] = full[999]          # a run that straddles the slab boundary at 1200
    bounds = [0, 1200, 3000]
    mine = full[bounds[rank]:bounds[rank + 1]]
    dev = torch.device("cpu")
    # minmax / pixel_pairs / nonzero
    pairs = int(np.count_nonzero(mine[1:] == mine[:-1]))
    got = sharded.reduce_scan(int(mine.min()), int(mine.max()), pairs, int(np.count_nonzero(mine)),
                              int(mine[0]) & 0xFFFFFFFFFFFFFFFF, int(mine[-1]) & 0xFFFFFFFFFFFFFFFF, mine.size, True, dev)
    assert got == (int(full.min()), int(full.max()), int(np.count_nonzero(full[1:] == full[:-1])), int(np.count_nonzero(full)))
    u64 = np.array([2 ** 64 - 1, 5, 2 ** 63], dtype=np.uint64) if rank else np.array([7, 7], dtype=np.uint64)
    got = sharded.reduce_scan(int(u64.min()), int(u64.max()), int(np.count_nonzero(u64[1:] == u64[:-1])), int(u64.size),
                              int(u64[0]), int(u64[-1]), u64.size, False, dev)
    assert got == (5, 2 ** 64 - 1, 1, 5)
    # an empty shard takes no part
    got = sharded.reduce_scan(3 if rank else 0, 9 if rank else 0, 2 if rank else 0, 4 if rank else 0, 3, 9, 4 if rank else 0, False, dev)
    assert got == (3, 9, 2, 4)
    # unique: per-shard (label, count) lists, merged by sum
    k, c = np.unique(mine, return_counts=True)
    ks, cs = sharded.allgather_pairs(torch.from_numpy(k), torch.from_numpy(c.astype(np.int64)))
    merged = {}
    for kk, cc in zip(ks, cs):
      for a, b in zip(kk.tolist(), cc.tolist()):
        merged[a] = merged.get(a, 0) + b
    eu, ec = np.unique(full, return_counts=True)
    assert merged == dict(zip(eu.tolist(), ec.tolist()))
    # KeyError status: the globally first missing voxel wins on every rank
    missing_local = np.flatnonzero(mine == 13)
    gi = int(missing_local[0]) + bounds[rank] if missing_local.size else -1
    first = sharded.allreduce_first_missing(gi, 13, None)
    assert (first == 13) == bool(np.any(full == 13))
    assert sharded.allreduce_first_missing(-1, 0, None) is None
    assert sharded.allreduce_first_missing(10 - rank, (2 ** 64 - 1) if rank else 3, None) == 2 ** 64 - 1
  finally:
    dist.destroy_process_group()


def test_sharded_reductions_world2():
  port = _free_port()
  mp.spawn(_reduce_worker, args=(2, port), nprocs=2, join=True)
