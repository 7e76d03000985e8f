// frb_build.cuh -- the shared "run heads -> label table" pass behind K1 (renumber, atomicMin of the
// first index) and K7 (component_map, atomicMax of the last run start).
//
// Streaming structure (stream_rows, frb_common.cuh): a CTA of 8 warps owns a CONTIGUOUS range of
// tiles, so consecutive tiles of one warp are spatially adjacent rows of the volume and see the
// same few labels; a warp keeps 4 x 512 B of 16-byte streaming loads in flight, finds the run
// heads (a[i] != a[i-1], exact across lanes / rows / tiles / CTAs) and offers (label, index) to
// the table.
//
// Two things keep the table off the critical path:
//  1. A per-warp shared-memory set of recently offered labels.  A warp walks its voxels in index
//     order (ascending for MIN, descending for MAX), so once it has offered a label every later
//     offer of that label by the same warp is dominated and is dropped without touching the table
//     (~90 % of the run heads on segmentation data).  The set is race free by construction: within
//     a row all reads (phase 1) are separated from all writes (phase 2) by __syncwarp, entries are
//     single 8-byte words, and a dropped offer is always dominated by one this warp already sent.
//  2. Surviving offers are deferred: the 16-byte slot read is issued when the head is found and
//     consumed at the end of the tile, after the next tile's streaming loads are in flight, so the
//     L2 round trip overlaps the HBM latency the warp has to wait for anyway.  The common outcome
//     ("label present with a better index") ends there without an atomic.
#pragma once
#include "frb_common.cuh"

namespace frb {

static constexpr int BUILD_TILE_VECS = STREAM_TILE_VECS;      // 16-byte vectors per CTA tile
static constexpr int BUILD_CACHE = 256;                       // entries in the per-warp label set

template <int OP>
__device__ __noinline__ void offer_slow(Slot* slots, u64 mask, int direct, u64* meta, u64 key, u64 v, unsigned int* inserted) {
  table_update<OP>(slots, mask, direct, meta, key, v, inserted);
}

template <int OP>
__device__ __noinline__ bool table_update_slow(Slot* slots, u64 mask, int direct, u64* meta, u64 key, u64 v, unsigned int* inserted) {
  return table_update<OP>(slots, mask, direct, meta, key, v, inserted);
}

// value offered for element i is i + bias  (MIN / forward: first occurrence; MAX / reverse: last run start)
template <int W, int OP, bool REVERSE>
__device__ __forceinline__ void build_pass(const uint4* __restrict__ a, size_t n, u64 bias, int skip_zero,
                                           Slot* __restrict__ slots, u64 mask, int direct, u64* __restrict__ meta, int run) {
  constexpr int V = 16 / W;
  __shared__ u64 s_seen[8][BUILD_CACHE];
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int inserted = 0;

  if (warp < STREAM_WARPS)
    for (int i = lane; i < BUILD_CACHE; i += 32) s_seen[warp][i] = FRB_EMPTY;
  __syncwarp();

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // scalar tail (< V elements at the very end); never filtered, so its order does not matter
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(a, i);
      if ((i == 0 || key != ld_elem<W>(a, i - 1)) && !(skip_zero && key == 0))
        offer_slow<OP>(slots, mask, direct, meta, key, (u64)i + bias, &inserted);
    }
  }

  const TileMap map(nvec, (size_t)run);
  // a table that overflowed in an earlier launch attempt cannot happen (the host re-inits), but a
  // table that overflows during this pass makes the rest of the pass pointless: checked per tile
  bool aborted = false;

  // one deferred offer per lane per tile
  bool pend = false;
  u64 pkey = 0, pval = 0;
  Slot pslot;
  pslot.key = 0;
  pslot.val = 0;

  auto row_fn = [&](int, const uint4& cur, bool valid, size_t i0, u64 prev, bool has_prev) {
    if (aborted) return;
    // phase 1: heads of this row that the warp has not offered before
    unsigned int todo = 0;
    if (valid) {
#pragma unroll
      for (int j = 0; j < V; j++) {
        const u64 key = vec_get<W>(cur, j);
        const bool head = (!has_prev || key != prev) && !(skip_zero && key == 0);
        prev = key;
        has_prev = true;
        if (head) {
          const bool hit = key != FRB_EMPTY && s_seen[warp][(hash_key(key) >> 48) & (BUILD_CACHE - 1)] == key;
          if (!hit) todo |= 1u << j;
        }
      }
    }
    __syncwarp();
    // phase 2: remember, and queue (or send) the offer
    if (todo) {
#pragma unroll
      for (int j = 0; j < V; j++) {
        if (todo & (1u << j)) {
          const u64 key = vec_get<W>(cur, j);
          const u64 h = hash_key(key);
          const u64 val = (u64)(i0 + j) + bias;
          s_seen[warp][(h >> 48) & (BUILD_CACHE - 1)] = key;
          if (!pend && key != FRB_EMPTY) {
            pend = true;
            pkey = key;
            pval = val;
            pslot = ld_slot_cg(slots + (direct ? key : (h & mask & ~1ull)));  // consumed in tile_end
          } else {
            offer_slow<OP>(slots, mask, direct, meta, key, val, &inserted);
          }
        }
      }
    }
    __syncwarp();
  };

  auto tile_end = [&](const uint4*, size_t) {
    if (pend) {
      if (!(pslot.key == pkey && op_settled<OP>(pslot.val, pval)))
        offer_slow<OP>(slots, mask, direct, meta, pkey, pval, &inserted);
      pend = false;
    }
    u64 ovf = 0;
    if (lane == 0) ovf = *((volatile u64*)&meta[FRB_META_OVERFLOW]);
    if (__shfl_sync(FRB_FULL_MASK, ovf, 0)) aborted = true;
  };

  stream_rows<W, REVERSE, true, false>(a, nvec, map, row_fn, tile_end);
  block_add_to(&meta[FRB_META_COUNT], inserted);
}

}  // namespace frb
