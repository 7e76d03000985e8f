"""
ctypes binding of libfastremap_b200.so (the C ABI declared in include/fastremap_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises, loudly.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("FASTREMAP_B200_SO") or os.path.join(HERE, "libfastremap_b200.so")

# dtype codes (frb_dtype)
U8, U16, U32, U64, I8, I16, I32, I64 = range(8)
OK, ERR_CUDA, ERR_BAD_ARG = 0, 3, 4
META_WORDS = 16
META_COUNT, META_OVERFLOW, META_SIDE_PRESENT, META_SIDE_VAL, META_MISSING_INDEX, META_LIST_COUNT, META_ASSIGNED, META_AUX1 = 0, 1, 2, 3, 4, 5, 6, 7
RANK_TAG = 1 << 63
OP_MIN, OP_MAX, OP_ADD, OP_SET = 0, 1, 2, 3
MISS_KEEP, MISS_ERROR, MISS_FILL = 0, 1, 2
EMPTY = 0xFFFFFFFFFFFFFFFF


class FastremapB200Error(RuntimeError):
  pass


_vp, _sz, _i, _u64, _i64, _dbl = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64,
                                  ctypes.c_int64, ctypes.c_double)

# name -> (restype, argtypes); must list every symbol include/fastremap_b200.h declares
SIGNATURES = {
  "frb_table_bytes": (_sz, [_u64]),
  "frb_table_direct_cap": (_u64, [_i]),
  "frb_table_init": (_i, [_vp, _u64, _vp, _u64, _vp]),
  "frb_table_insert": (_i, [_vp, _i, _vp, _i, _u64, _sz, _i, _i, _vp, _u64, _vp, _vp]),
  "frb_table_insert_gathered": (_i, [_vp, _u64, _i, _vp, _i, _i, _vp, _u64, _vp, _vp]),
  "frb_table_insert_packed": (_i, [_vp, _u64, _i, _u64, _i, _i, _vp, _u64, _vp, _vp]),
  "frb_table_resolve": (_i, [_vp, _i, _vp, _u64, _vp, _vp]),
  "frb_table_project": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _u64, _vp, _vp]),
  "frb_table_export": (_i, [_vp, _u64, _vp, _vp, _vp, _sz, _vp]),
  "frb_minmax_pairs": (_i, [_vp, _sz, _i, _vp, _vp]),
  "frb_renumber_build": (_i, [_vp, _sz, _i, _u64, _i, _vp, _u64, _vp, _u64, _vp]),
  "frb_renumber_rank_ws_bytes": (_sz, [_sz]),
  "frb_renumber_rank": (_i, [_vp, _u64, _vp, _i, _i64, _u64, _u64, _vp, _vp, _sz, _vp, _sz, _vp]),
  "frb_relabel": (_i, [_vp, _vp, _vp, _i, _sz, _i, _vp, _u64, _i, _vp, _i, _u64, _i, _u64, _i, _i, _u64, _vp]),
  "frb_relabel_auto": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _u64, _i, _vp, _i, _u64, _i, _u64, _u64, _u64, _vp]),
  "frb_replace_one": (_i, [_vp, _vp, _sz, _i, _u64, _u64, _vp]),
  "frb_remap_from_array": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _vp]),
  "frb_swap_halves": (_i, [_vp, _vp, _sz, _i, _vp]),
  "frb_pack_pairs": (_i, [_vp, _i, _vp, _i, _vp, _sz, _vp]),
  "frb_cast": (_i, [_vp, _i, _vp, _i, _sz, _vp]),
  "frb_transpose_tiles": (_i, [_vp, _vp, _i, _u64, _u64, _u64, _u64, _i, _vp, _vp, _vp, _vp]),
  "frb_copy_rows": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _vp, _vp]),
  "frb_hist_dense": (_i, [_vp, _sz, _i, _u64, _vp, _u64, _u64, _vp]),
  "frb_compact_bins_ws_bytes": (_sz, [_u64]),
  "frb_compact_bins_count": (_i, [_vp, _u64, _vp, _vp, _sz, _vp]),
  "frb_compact_bins_emit": (_i, [_vp, _u64, _i, _u64, _vp, _vp, _sz, _vp, _vp]),
  "frb_hist_ws_bytes": (_sz, [_sz]),
  "frb_hist_ws_flags_offset": (_sz, []),
  "frb_hist_hash": (_i, [_vp, _sz, _i, _vp, _u64, _vp, _u64, _u64, _vp, _sz, _vp]),
  "frb_hist_apply_log": (_i, [_vp, _sz, _sz, _i, _vp, _u64, _vp, _vp]),
  "frb_sort_pairs_ws_bytes": (_sz, [_sz]),
  "frb_sort_table_pairs": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
  "frb_direct_export_sorted": (_i, [_vp, _u64, _i, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
  "frb_unique_sort_ws_bytes": (_sz, [_sz, _i]),
  "frb_unique_sort_count": (_i, [_vp, _sz, _i, _vp, _vp, _sz, _vp]),
  "frb_unique_sort_emit": (_i, [_sz, _i, _vp, _vp, _sz, _vp, _vp]),
  "frb_indices_ws_bytes": (_sz, [_sz]),
  "frb_indices_count": (_i, [_vp, _sz, _i, _u64, _vp, _vp, _sz, _vp]),
  "frb_indices_emit": (_i, [_vp, _sz, _i, _u64, _vp, _sz, _vp, _vp]),
  "frb_nonzero_ws_bytes": (_sz, [_sz]),
  "frb_nonzero_count": (_i, [_vp, _sz, _i, _vp, _vp, _sz, _vp]),
  "frb_nonzero_emit": (_i, [_vp, _sz, _i, _vp, _vp, _sz, _vp, _vp]),
  "frb_sort_pairs_u64_ws_bytes": (_sz, [_sz]),
  "frb_sort_pairs_u64": (_i, [_vp, _vp, _sz, _i, _vp, _sz, _vp]),
  "frb_unravel_u16": (_i, [_vp, _sz, _i, _u64, _u64, _vp, _vp]),
  "frb_component_build": (_i, [_vp, _sz, _i, _u64, _vp, _u64, _vp, _u64, _u64, _vp, _u64, _vp]),
  "frb_component_unhead": (_i, [_vp, _i, _u64, _u64, _vp, _u64, _vp, _vp]),
  "frb_component_gather": (_i, [_vp, _u64, _vp, _i, _vp, _i, _u64, _vp, _vp, _sz, _vp, _u64, _vp]),
  "frb_synth_volume": (_i, [_vp, _i, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _dbl, _dbl, _u64, _vp]),
  "frb_small_table_cap": (_u64, [_sz, _i]),
  "frb_small_list_capacity": (_sz, [_sz, _i]),
  "frb_small_ws_bytes": (_sz, [_sz, _i]),
  "frb_renumber_small": (_i, [_vp, _sz, _i, _i64, _i, _vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _sz, _vp]),
  "frb_unique_small": (_i, [_vp, _sz, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
  "frb_stream_probe": (_i, [_vp, _sz, _vp, _vp, _i, _vp, _vp]),
  "frb_set_device": (_i, [_i]),
  "frb_set_tile_run": (None, [_i]),
  "frb_set_relabel_variant": (None, [_i]),
  "frb_launch_count": (_u64, []),
  "frb_last_error": (ctypes.c_char_p, []),
}

_lib = None


def load():
  """Load the CUDA library (once).  Raises FastremapB200Error if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(SO_PATH):
    raise FastremapB200Error(
      "fastremap_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
      "(or `python fastremap_b200/_build.py`). There is no CPU fallback." % SO_PATH)
  L = ctypes.CDLL(SO_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(L, name)  # AttributeError here means the .so is stale: rebuild
    fn.restype = res
    fn.argtypes = args
  _lib = L
  return L


def check(rc: int, what: str = ""):
  if rc != OK:
    msg = load().frb_last_error()
    raise FastremapB200Error("%s failed (code %d): %s" % (what or "fastremap_b200 call", rc, (msg or b"").decode()))


def launch_count() -> int:
  return int(load().frb_launch_count())
