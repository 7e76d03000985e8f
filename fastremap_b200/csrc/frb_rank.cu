// frb_rank.cu -- K2 of renumber: rank the labels that have no id yet by their first index.
//
// Replaces the implicit `remap_id += 1` ordering of _renumber (fastremap.pyx:244-246).  The
// reference hands out ids with a serial counter; here the table value of a label is either
//   TAG | first flat index   (label seen by K1, not ranked yet; TAG = bit 63), or
//   rank                     (0-based sequence number of the label in encounter order),
// and id = (T)(start + rank) is formed by the relabel kernel.  Ranked values are smaller than every
// tagged value, so later K1 launches (atomicMin) never disturb them: ranking is INCREMENTAL.  A
// volume can be renumbered chunk by chunk in memory order (K1, K2, K3 per chunk) because the id of
// a label depends only on the prefix of the array in front of its first occurrence -- that is what
// lets the host pipeline overlap the PCIe transfers in both directions with the kernels.
//
// Two launches, no host round trip (the list length stays on the device):
//   k_export_pending    scans the slots, compacts (first index - index_base, slot) of tagged values
//   k_rank_sort_assign  ONE persistent kernel: stable LSD radix sort (8-bit digits) of the list with
//                       a grid barrier between the histogram and scatter phases of every pass,
//                       then rank -> table value, (label, id) -> mapping arrays.
#include "frb_internal.cuh"

namespace frb {

static constexpr u64 RANK_TAG = 0x8000000000000000ull;
static constexpr int RS_THREADS = 256;
static constexpr int RS_ITEMS = 8;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048
static constexpr int RS_WARPS = RS_THREADS / 32;

__device__ __forceinline__ u64 ldcg64(const u64* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------- export
__global__ void __launch_bounds__(256)
k_export_pending(const Slot* __restrict__ slots, u64 cap, u64* meta, u64 index_base, u64 index_count,
                 u64* __restrict__ key_out, u64* __restrict__ slot_out, size_t capacity) {
  constexpr int U = 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const size_t cap_r = (cap + 31) & ~(size_t)31;
  for (size_t s0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s0 < cap_r; s0 += stride * U) {
    Slot cur[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t s = s0 + (size_t)u * stride;
      cur[u].key = FRB_EMPTY;
      cur[u].val = 0;
      if (s < cap) cur[u] = ld_slot_cg(slots + s);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t s = s0 + (size_t)u * stride;
      // pending = tagged first index inside [index_base, index_base + index_count): a chunk pipeline that
      // restarts after a table overflow may already hold (incomplete) entries of LATER chunks
      const bool pending = cur[u].key != FRB_EMPTY && (cur[u].val & RANK_TAG) &&
                           ((cur[u].val & ~RANK_TAG) - index_base) < index_count;
      const unsigned int b = __ballot_sync(FRB_FULL_MASK, pending);
      if (b == 0) continue;
      u64 base = 0;
      if (lane == 0) base = atomicAdd(&meta[FRB_META_LIST_COUNT], (u64)__popc(b));
      base = __shfl_sync(FRB_FULL_MASK, base, 0);
      if (pending) {
        const u64 pos = base + __popc(b & ((1u << lane) - 1u));
        if (pos < capacity) {
          key_out[pos] = (cur[u].val & ~RANK_TAG) - index_base;
          slot_out[pos] = (u64)s;
        }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && meta[FRB_META_SIDE_PRESENT]) {
    const u64 v = meta[FRB_META_SIDE_VAL];
    if ((v & RANK_TAG) && ((v & ~RANK_TAG) - index_base) < index_count) {
      const u64 pos = atomicAdd(&meta[FRB_META_LIST_COUNT], 1ull);
      if (pos < capacity) {
        key_out[pos] = (v & ~RANK_TAG) - index_base;
        slot_out[pos] = cap;  // "slot cap" = the out-of-band all-ones key
      }
    }
  }
}

// ---------------------------------------------------------------------------------- sort + assign
// All CTAs of the (cooperatively launched, hence co-resident) grid meet here.  Data produced by
// other CTAs is read with ld.global.cg afterwards (an L1 line could be stale).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    generation++;
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = generation * gridDim.x;
    while (*((volatile unsigned int*)counter) < target) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}

struct RankArgs {
  u64* keys;      // sort key (first index - base)        } list written by k_export_pending
  u64* pay;       // payload (slot index)                  }
  u64* keys_alt;  // ping-pong buffers
  u64* pay_alt;
  unsigned int* hist;     // [gridDim.x][256]
  unsigned int* barrier;  // zeroed before the launch
  Slot* slots;
  u64 cap;
  u64* meta;
  u64 start_bits;
  void* keys_out;
  void* ids_out;
  u64 capacity;  // entries the list buffers AND the mapping arrays can hold
  int index_bits;
};

template <int W>
__global__ void __launch_bounds__(RS_THREADS) k_rank_sort_assign(RankArgs A) {
  __shared__ unsigned int h[256];
  __shared__ unsigned int base[256];
  __shared__ unsigned int wc[RS_WARPS][256];
  __shared__ unsigned int tile_tot[256];
  __shared__ unsigned int warp_sum[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;

  const u64 n = ldcg64(&A.meta[FRB_META_LIST_COUNT]);
  const u64 assigned = ldcg64(&A.meta[FRB_META_ASSIGNED]);
  // an overflowed build (K1 aborted) must not be ranked: the list is incomplete.  Every CTA takes the
  // same decision: n / assigned are stable, and OVERFLOW is only ever raised below when the second
  // condition already holds.
  if (ldcg64(&A.meta[FRB_META_OVERFLOW]) != 0 || n > A.capacity || assigned + n > A.capacity) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(&A.meta[FRB_META_OVERFLOW], 1ull);
    return;
  }

  u64 chunk = (n + G - 1) / G;
  chunk = align_up(chunk < 1 ? 1 : chunk, RS_TILE);
  u64 beg = (u64)blockIdx.x * chunk;
  if (beg > n) beg = n;
  u64 end = beg + chunk;
  if (end > n) end = n;
  const unsigned int active = (unsigned int)((n + chunk - 1) / chunk);  // CTAs that own keys (<= G)

  u64* kin = A.keys;
  u64* vin = A.pay;
  u64* kout = A.keys_alt;
  u64* vout = A.pay_alt;
  unsigned int generation = 0;

  for (int shift = 0; shift < A.index_bits && n > 1; shift += 8) {
    const int nb = (A.index_bits - shift) < 8 ? (A.index_bits - shift) : 8;
    const unsigned int dmask = (1u << nb) - 1u;
    // ---- digit histogram of this CTA's chunk
    h[threadIdx.x] = 0;
    __syncthreads();
    if (beg < end) {
      const u64 end_r = beg + align_up(end - beg, 32);
      for (u64 i = beg + threadIdx.x; i < end_r; i += RS_THREADS) {
        const bool valid = i < end;
        const unsigned int m = __ballot_sync(FRB_FULL_MASK, valid);
        if (valid) {
          const unsigned int d = (unsigned int)(ldcg64(kin + i) >> shift) & dmask;
          const unsigned int peers = __match_any_sync(m, d);
          if (lane == __ffs(peers) - 1) atomicAdd(&h[d], (unsigned int)__popc(peers));
        }
      }
    }
    __syncthreads();
    if (blockIdx.x < active) A.hist[(size_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];
    grid_barrier(A.barrier, generation);
    // ---- global position of this CTA's first key of every digit (digit-major, CTA-minor order)
    {
      // thread d sums digit d over the CTAs that own keys (coalesced rows of 256 counters)
      unsigned int tot = 0, part = 0;
      const unsigned int* col = A.hist + threadIdx.x;
      unsigned int b = 0;
      for (; b + 8 <= active; b += 8) {
        unsigned int v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = __ldcg(col + (size_t)(b + q) * 256);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          part += (b + q) < blockIdx.x ? v[q] : 0u;
          tot += v[q];
        }
      }
      for (; b < active; b++) {
        const unsigned int v = __ldcg(col + (size_t)b * 256);
        part += b < blockIdx.x ? v : 0u;
        tot += v;
      }
      unsigned int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(FRB_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) warp_sum[warp] = incl;
      __syncthreads();
      unsigned int woff = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++) woff += w < warp ? warp_sum[w] : 0u;
      base[threadIdx.x] = woff + incl - tot + part;
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++) wc[w][threadIdx.x] = 0;
    }
    __syncthreads();
    // ---- stable scatter, tile by tile
    for (u64 tile = beg; tile < end; tile += RS_TILE) {
      u64 key[RS_ITEMS];
      unsigned int rank[RS_ITEMS];
      unsigned int dig[RS_ITEMS];
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * RS_ITEMS) + k * 32 + lane;
        const bool valid = i < end;
        key[k] = valid ? ldcg64(kin + i) : 0ull;
        const unsigned int d = (unsigned int)(key[k] >> shift) & dmask;
        dig[k] = d;
        const unsigned int m = __ballot_sync(FRB_FULL_MASK, valid);
        unsigned int peers = 0, pre = 0;
        if (valid) {
          peers = __match_any_sync(m, d);
          pre = wc[warp][d];
        }
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) wc[warp][d] = pre + (unsigned int)__popc(peers);
        __syncwarp();
        rank[k] = pre + (unsigned int)__popc(peers & ((1u << lane) - 1u));
      }
      __syncthreads();
      {
        unsigned int run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
          const unsigned int t = wc[w][threadIdx.x];
          wc[w][threadIdx.x] = run;
          run += t;
        }
        tile_tot[threadIdx.x] = run;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * RS_ITEMS) + k * 32 + lane;
        if (i < end) {
          const u64 pos = (u64)base[dig[k]] + wc[warp][dig[k]] + rank[k];
          kout[pos] = key[k];
          vout[pos] = ldcg64(vin + i);
        }
      }
      __syncthreads();
      base[threadIdx.x] += tile_tot[threadIdx.x];
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++) wc[w][threadIdx.x] = 0;
      __syncthreads();
    }
    grid_barrier(A.barrier, generation);
    u64* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }

  // ---- rank r of the sorted list -> table value `assigned + r`; mapping arrays in id order
  for (u64 r = (u64)blockIdx.x * RS_THREADS + threadIdx.x; r < n; r += (u64)G * RS_THREADS) {
    const u64 s = ldcg64(vin + r);
    const u64 rank = assigned + r;
    u64 key;
    if (s == A.cap) {
      key = FRB_EMPTY;
      A.meta[FRB_META_SIDE_VAL] = rank;
    } else {
      key = A.slots[s].key;
      A.slots[s].val = rank;
    }
    st_elem<W>(A.keys_out, rank, key);
    st_elem<W>(A.ids_out, rank, trunc_width(A.start_bits + rank, W));  // counter arithmetic wraps in the array dtype (:225)
  }
  grid_barrier(A.barrier, generation);  // every CTA has read `assigned` long before this point
  if (blockIdx.x == 0 && threadIdx.x == 0) A.meta[FRB_META_ASSIGNED] = assigned + n;
}

static int rank_grid(const void* kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RS_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  (void)per_sm;
  return sm_count();  // one CTA per SM: the barrier cost grows with the grid, the work is tiny
}

}  // namespace frb

using namespace frb;

extern "C" {

// workspace: 4 lists of `capacity` uint64 (sort key, payload, 2 ping-pong) + histograms + barrier word
size_t frb_renumber_rank_ws_bytes(size_t capacity) {
  const size_t c = capacity < 1 ? 1 : capacity;
  return 4 * align_up(c * sizeof(u64), 256) + align_up((size_t)256 * 2 * (size_t)sm_count() * sizeof(unsigned int), 256) + 256;
}

int frb_renumber_rank(void* slots, uint64_t cap, uint64_t* meta, int dtype, int64_t start, uint64_t index_base,
                      uint64_t index_count, void* keys_out, void* ids_out, size_t capacity, void* ws, size_t ws_bytes,
                      void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || !ws || ws_bytes < frb_renumber_rank_ws_bytes(capacity) || capacity < 1 ||
      !keys_out || !ids_out || index_count < 1 || index_count > (1ull << 62)) {
    set_error("frb_renumber_rank: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t lb = align_up(capacity * sizeof(u64), 256);
  const size_t hb = align_up((size_t)256 * 2 * (size_t)sm_count() * sizeof(unsigned int), 256);
  char* base = (char*)ws;
  RankArgs A;
  A.keys = (u64*)(base);
  A.pay = (u64*)(base + lb);
  A.keys_alt = (u64*)(base + 2 * lb);
  A.pay_alt = (u64*)(base + 3 * lb);
  A.hist = (unsigned int*)(base + 4 * lb);
  A.barrier = (unsigned int*)(base + 4 * lb + hb);
  A.slots = (Slot*)slots;
  A.cap = cap;
  A.meta = (u64*)meta;
  const int w = dtype_width(dtype);
  A.start_bits = trunc_width((u64)start, w);
  A.keys_out = keys_out;
  A.ids_out = ids_out;
  A.capacity = capacity;
  int index_bits = 1;
  while (index_bits < 63 && ((index_count - 1) >> index_bits) != 0) index_bits++;
  A.index_bits = index_bits;

  cudaMemsetAsync(&A.meta[FRB_META_LIST_COUNT], 0, sizeof(u64), st);
  cudaMemsetAsync(A.barrier, 0, 256, st);
  Grid g = grid_for((size_t)cap, 256 * 4, 8);
  k_export_pending<<<g.blocks, g.threads, 0, st>>>((const Slot*)slots, cap, A.meta, index_base, index_count, A.keys, A.pay, capacity);
  int rc = check_launch("k_export_pending");
  if (rc) return rc;

  const void* kernel;
  switch (w) {
    case 1: kernel = (const void*)k_rank_sort_assign<1>; break;
    case 2: kernel = (const void*)k_rank_sort_assign<2>; break;
    case 4: kernel = (const void*)k_rank_sort_assign<4>; break;
    default: kernel = (const void*)k_rank_sort_assign<8>; break;
  }
  void* params[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(rank_grid(kernel)), dim3(RS_THREADS), params, 0, st);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return FRB_ERR_CUDA;
  }
  return check_launch("k_rank_sort_assign");
}

}  // extern "C"
