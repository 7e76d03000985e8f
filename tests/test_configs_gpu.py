"""
The BASELINE.json configs at their stated sizes, against the oracle (-m gpu).

Each test runs one config of benchmarks/configs.py -- the code bench.py puts into the `configs` list of its N=1 line --
with the oracle port as the CPU side: the 1024^3 volume (512^3 for config 1) is generated on the device, the op runs on
the WHOLE volume with its full-size table (1 M-entry dict, 5 M-label keep list through k_relabel_batched, 10 M-label
component table), and the result is compared bit for bit with the oracle on the whole volume (configs 1, 2) or on the
first z-planes (configs 3, 5), plus the closed-form / conservation checks that hold at full size.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SLAB = 128  # z-planes the oracle re-computes for the slab-local configs (2^27 voxels)


@pytest.fixture(scope="module")
def cfg(frb):
  from benchmarks import configs
  return configs


def _run(cfg, fn, oracle):
  rec = fn(oracle, "port", 6552.0, 1, slab_planes=SLAB)
  assert rec["parity"], rec
  assert rec.get("full_size_check", True), rec
  return rec


def test_config1_renumber_512_vs_oracle(cfg, oracle):
  rec = _run(cfg, cfg.config1, oracle)
  assert rec["labels"] == 87567 or rec["labels"] > 80000  # ~88 k labels


def test_config2a_unique_dense_1024_u32_vs_oracle(cfg, oracle):
  rec = _run(cfg, cfg.config2a, oracle)
  assert rec["n_unique"] == 87555


def test_config2b_unique_sparse_1024_u32_vs_oracle(cfg, oracle, frb):
  rec = _run(cfg, cfg.config2b, oracle)
  assert rec["n_unique"] > 850_000
  # closed form at full size: a label's count is the summed volume of the cells that carry it
  import torch
  from fastremap_b200.synth import synth_cell_labels, synth_device
  S, g = 1024, 100
  vol = synth_device((S, S, S), (g, g, g), np.uint32, seed=1, zero_frac=0.1)
  u, c = frb.unique(vol, return_counts=True)
  u, c = u.cpu().numpy(), c.cpu().numpy().astype(np.int64)
  width = np.bincount((np.arange(S, dtype=np.int64) * g) // S, minlength=g)            # voxels per cell along one axis
  cell_vox = (width[:, None, None] * width[None, :, None] * width[None, None, :]).reshape(-1)
  cells = synth_cell_labels((g, g, g), np.uint32, seed=1, zero_frac=0.1).reshape(-1)
  eu, inv = np.unique(cells, return_inverse=True)
  ec = np.bincount(inv, weights=cell_vox.astype(np.float64), minlength=eu.size).astype(np.int64)
  assert np.array_equal(u, eu) and np.array_equal(c, ec)
  del vol
  torch.cuda.empty_cache()


def test_config3_remap_1M_entry_dict_vs_oracle(cfg, oracle):
  rec = _run(cfg, cfg.config3, oracle)
  assert rec["dict_entries"] == 1_000_000 and rec["labels_in_volume"] > 1_100_000


def test_config5_mask_except_5M_labels_batched_lookup_vs_oracle(cfg, oracle):
  rec = _run(cfg, cfg.config5a, oracle)
  assert rec["kept_labels"] >= 4_900_000  # >= 2^14 labels: k_relabel_batched, table resident in HBM


def test_config5_component_map_10M_labels_vs_oracle(cfg, oracle):
  rec = _run(cfg, cfg.config5b, oracle)
  assert rec["n_components"] > 9_900_000
