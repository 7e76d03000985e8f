// frb_coopsort.cuh -- stable LSD radix sort of (key, payload) pairs inside ONE persistent kernel.
//
// The O(L) lists of the hot path (L = number of distinct labels: 10^4 .. 10^7) are sorted by a
// cooperatively launched grid of one CTA per SM: every pass is histogram -> grid barrier -> scatter
// -> grid barrier, the list length is read from device memory (no host round trip), and a pass whose
// digit is the same for every key is skipped after its histogram.  At 10^5 keys a pass costs about
// ten microseconds -- a separate launch per phase would cost that much in launch latency alone.
#pragma once
#include "frb_common.cuh"

namespace frb {

static constexpr int CS_THREADS = 256;
static constexpr int CS_ITEMS = 8;
static constexpr int CS_TILE = CS_THREADS * CS_ITEMS;  // 2048
static constexpr int CS_WARPS = CS_THREADS / 32;

// All CTAs of the (cooperatively launched, hence co-resident) grid meet here.  Data produced by
// other CTAs is read with ld.global.cg afterwards (an L1 line could be stale).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    generation++;
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = generation * gridDim.x;
    while (*((volatile unsigned int*)counter) < target) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

struct CoopSortBuffers {
  u64* keys;      // in: the list; the sorted list ends up in *sorted_keys / *sorted_pay below
  u64* pay;
  u64* keys_alt;  // ping-pong buffers
  u64* pay_alt;
  u64* pay2;      // optional second payload (NULL: none) and its ping-pong buffer
  u64* pay2_alt;
  unsigned int* hist;     // [2][gridDim.x][256] (double-buffered: a skipped pass has no second barrier)
  unsigned int* barrier;  // one word, zeroed before the launch
};

#ifndef FRB_COOP_CTAS_PER_SM
#define FRB_COOP_CTAS_PER_SM 1
#endif
// host side: CTAs of a cooperative sort launch (co-resident by construction)
inline int coop_sort_grid() { return FRB_COOP_CTAS_PER_SM * sm_count(); }
// host side: bytes of the hist + barrier area for a grid of `ctas` CTAs
inline size_t coop_sort_aux_bytes(int ctas) { return (size_t)2 * ctas * 256 * sizeof(unsigned int) + 256; }

// Sort the n pairs by bits [begin_bit, end_bit) of the key.  Must be called by every thread of every
// CTA with identical arguments.  Shared memory: 3 * 256 + 8 * 256 + 8 words (static, declared here).
__device__ __forceinline__ void coop_radix_sort(const CoopSortBuffers& B, u64 n, int begin_bit, int end_bit,
                                                unsigned int& generation, u64** sorted_keys, u64** sorted_pay,
                                                u64** sorted_pay2 = nullptr) {
  __shared__ unsigned int h[256];
  __shared__ unsigned int base[256];
  __shared__ unsigned int wc[CS_WARPS][256];
  __shared__ unsigned int tile_tot[256];
  __shared__ unsigned int warp_sum[CS_WARPS];
  __shared__ int s_skip;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;

  u64 chunk = (n + G - 1) / G;
  chunk = (chunk < 1 ? 1 : chunk);
  chunk = (chunk + CS_TILE - 1) / CS_TILE * CS_TILE;
  u64 beg = (u64)blockIdx.x * chunk;
  if (beg > n) beg = n;
  u64 end = beg + chunk;
  if (end > n) end = n;
  const unsigned int active = (unsigned int)((n + chunk - 1) / chunk);  // CTAs that own keys (<= G)
  const bool one_tile = chunk == CS_TILE;  // the CTA's keys stay in registers between the two phases

  u64* kin = B.keys;
  u64* vin = B.pay;
  u64* kout = B.keys_alt;
  u64* vout = B.pay_alt;
  u64* win = B.pay2;
  u64* wout = B.pay2_alt;
  const bool two = B.pay2 != nullptr;

  int pass = 0;
  for (int shift = begin_bit; shift < end_bit && n > 1; shift += 8, pass++) {
    unsigned int* hist = B.hist + (size_t)(pass & 1) * G * 256;
    const int nb = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
    const unsigned int dmask = (1u << nb) - 1u;
    u64 key[CS_ITEMS];
    // ---- digit histogram of this CTA's chunk
    h[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_skip = 0;
    __syncthreads();
    for (u64 tile = beg; tile < end; tile += CS_TILE) {
#pragma unroll
      for (int k = 0; k < CS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * CS_ITEMS) + k * 32 + lane;
        key[k] = i < end ? __ldcg(kin + i) : 0ull;
      }
#pragma unroll
      for (int k = 0; k < CS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * CS_ITEMS) + k * 32 + lane;
        const bool valid = i < end;
        const unsigned int m = __ballot_sync(FRB_FULL_MASK, valid);
        if (valid) {
          const unsigned int d = (unsigned int)(key[k] >> shift) & dmask;
          const unsigned int peers = __match_any_sync(m, d);
          if (lane == __ffs(peers) - 1) atomicAdd(&h[d], (unsigned int)__popc(peers));
        }
      }
    }
    __syncthreads();
    if (blockIdx.x < active) hist[(size_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];
    grid_barrier(B.barrier, generation);
    // ---- global position of this CTA's first key of every digit (digit-major, CTA-minor order)
    {
      unsigned int tot = 0, part = 0;
      const unsigned int* col = hist + threadIdx.x;
      unsigned int b = 0;
      for (; b + 32 <= active; b += 32) {  // 32 independent L2 loads in flight per thread
        unsigned int v[32];
#pragma unroll
        for (int q = 0; q < 32; q++) v[q] = __ldcg(col + (size_t)(b + q) * 256);
#pragma unroll
        for (int q = 0; q < 32; q++) {
          part += (b + q) < blockIdx.x ? v[q] : 0u;
          tot += v[q];
        }
      }
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
      if ((u64)tot == n) s_skip = 1;  // every key has this digit: the pass would be the identity
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
      for (int w = 0; w < CS_WARPS; w++) woff += w < warp ? warp_sum[w] : 0u;
      base[threadIdx.x] = woff + incl - tot + part;
#pragma unroll
      for (int w = 0; w < CS_WARPS; w++) wc[w][threadIdx.x] = 0;
    }
    __syncthreads();
    const bool skip = s_skip != 0;  // the same decision in every CTA (it comes from the global histogram)
    __syncthreads();                // s_skip is reset at the top of the next pass
    if (skip) continue;
    // ---- stable scatter, tile by tile
    for (u64 tile = beg; tile < end; tile += CS_TILE) {
      u64 val[CS_ITEMS], val2[CS_ITEMS];
      unsigned int rank[CS_ITEMS];
      unsigned int dig[CS_ITEMS];
#pragma unroll
      for (int k = 0; k < CS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * CS_ITEMS) + k * 32 + lane;
        if (!one_tile) key[k] = i < end ? __ldcg(kin + i) : 0ull;
        val[k] = i < end ? __ldcg(vin + i) : 0ull;
        val2[k] = (two && i < end) ? __ldcg(win + i) : 0ull;
      }
#pragma unroll
      for (int k = 0; k < CS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * CS_ITEMS) + k * 32 + lane;
        const bool valid = i < end;
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
        for (int w = 0; w < CS_WARPS; w++) {
          const unsigned int t = wc[w][threadIdx.x];
          wc[w][threadIdx.x] = run;
          run += t;
        }
        tile_tot[threadIdx.x] = run;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CS_ITEMS; k++) {
        const u64 i = tile + (u64)warp * (32 * CS_ITEMS) + k * 32 + lane;
        if (i < end) {
          const u64 pos = (u64)base[dig[k]] + wc[warp][dig[k]] + rank[k];
          kout[pos] = key[k];
          vout[pos] = val[k];
          if (two) wout[pos] = val2[k];
        }
      }
      __syncthreads();
      base[threadIdx.x] += tile_tot[threadIdx.x];
#pragma unroll
      for (int w = 0; w < CS_WARPS; w++) wc[w][threadIdx.x] = 0;
      __syncthreads();
    }
    grid_barrier(B.barrier, generation);
    u64* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
    t = win; win = wout; wout = t;
  }
  *sorted_keys = kin;
  *sorted_pay = vin;
  if (sorted_pay2) *sorted_pay2 = win;
}

}  // namespace frb
