// frb_scan.cuh -- scan a slot array and compact the selected slots into output lists.
//
// Used by table export (unique, the multi-GPU merges), K2's "labels not ranked yet" list and
// component_map's gather.  The slot array streams through the TMA-staged loop of frb_common.cuh
// (one 16-byte vector = one slot); the indices of selected slots are staged in shared memory and a
// CTA reserves output space with a single global atomic per few thousand hits -- one atomic per warp
// row would put 10^5..10^6 atomics on the same address, and even one per tile leaves the CTA waiting
// for an L2 round trip every 16 KiB.  Tables hold at most 2^32 slots (64 GiB).
#pragma once
#include "frb_common.cuh"

namespace frb {

// Sel: bool pick(u64 key, u64 val) const;  u64 fetch(u64 key, u64 val) const;  void emit(u64 pos, u64 key, u64 val, u64 slot, u64 fetched) const;
// fetch is the (optional) dependent read of an emit -- component_map's parent[winning index]: the flush issues the slot
// re-reads of four hits, then their four fetches, then the four emits, so the round trips of a thread overlap (with one
// hit at a time the gather of 10^7 labels spent 0.6 ms on 16 serial L2 + HBM round trips per thread and flush).
static constexpr int SCAN_BUF = 4096;  // selected slot indices staged per CTA between two flushes

template <typename Sel>
__global__ void __launch_bounds__(STREAM_THREADS)
k_table_scan(const uint4* __restrict__ slots, u64 cap, u64* counter, u64 capacity, Sel sel, int run) {
  __shared__ unsigned int s_slot[SCAN_BUF];
  __shared__ unsigned int s_count;
  __shared__ u64 s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  auto row_fn = [&](int, const uint4& cur, bool valid, size_t i0, u64, bool) {
    const u64 key = (u64)cur.x | ((u64)cur.y << 32), val = (u64)cur.z | ((u64)cur.w << 32);
    const bool hit = valid && sel.pick(key, val);
    const unsigned int b = __ballot_sync(FRB_FULL_MASK, hit);
    if (b == 0) return;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(&s_count, (unsigned int)__popc(b));
    base = __shfl_sync(FRB_FULL_MASK, base, 0);
    if (hit) s_slot[base + __popc(b & ((1u << lane) - 1u))] = (unsigned int)(i0 >> 1);
  };
  // all consumer threads: reserve output space with ONE global atomic, then re-read the selected
  // slots (they have just streamed through L2) and emit them
  auto flush = [&]() {
    const unsigned int count = s_count;
    if (threadIdx.x == 0 && count) s_base = atomicAdd(counter, (u64)count);
    consumer_bar();
    const u64 base = s_base;
    constexpr unsigned int T = STREAM_WARPS * 32, B = 4;
    for (unsigned int i0 = threadIdx.x; i0 < count; i0 += T * B) {
      Slot sl[B];
      u64 slot[B], got[B];
      bool live[B];
#pragma unroll
      for (unsigned int b = 0; b < B; b++) {
        const unsigned int i = i0 + b * T;
        live[b] = i < count && base + i < capacity;
        slot[b] = live[b] ? (u64)s_slot[i] : 0;
        sl[b].key = 0;
        sl[b].val = 0;
        if (live[b]) sl[b] = ld_slot_cg_batch(reinterpret_cast<const Slot*>(slots) + slot[b]);
      }
#pragma unroll
      for (unsigned int b = 0; b < B; b++) got[b] = live[b] ? sel.fetch(sl[b].key, sl[b].val) : 0;
#pragma unroll
      for (unsigned int b = 0; b < B; b++)
        if (live[b]) sel.emit(base + i0 + b * T, sl[b].key, sl[b].val, slot[b], got[b]);
    }
    consumer_bar();
    if (threadIdx.x == 0) s_count = 0;
    consumer_bar();
  };
  auto after_tile = [&](size_t) {
    consumer_bar();  // all rows of this tile are staged; s_count is stable
    if (s_count + STREAM_TILE_VECS > SCAN_BUF) flush();  // the same decision in every consumer warp
    else consumer_bar();                                 // nobody may stage the next tile before everyone has read s_count
  };
  const TileMap map((size_t)cap, (size_t)run);
  stream_rows<8, false, false, 0>(slots, (size_t)cap, map, row_fn, [](const uint4*, size_t) {}, after_tile);
  if (warp < STREAM_WARPS) {
    consumer_bar();
    flush();
  }
}

template <typename Sel>
inline int launch_table_scan(const void* slots, u64 cap, u64* counter, u64 capacity, const Sel& sel, const char* what, cudaStream_t st) {
  if (cap > (1ull << 32)) {
    set_error("label tables hold at most 2^32 slots (64 GiB)");
    return FRB_ERR_BAD_ARG;
  }
  const size_t tiles = ((size_t)cap + STREAM_TILE_VECS - 1) / STREAM_TILE_VECS;
  k_table_scan<Sel><<<stream_grid(k_table_scan<Sel>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>(
      (const uint4*)slots, cap, counter, capacity, sel, fair_run(stream_grid(k_table_scan<Sel>, tiles), tiles));
  return check_launch(what);
}

}  // namespace frb
