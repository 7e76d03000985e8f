// frb_sort.cu -- stable LSD radix sort (8-bit digits) and a single-CTA scan.
//
// Used for the O(L) work of the hot path (L = number of distinct labels): ranking labels by
// first index in renumber (K2), ordering (label, count) pairs in unique, and -- over all N
// voxels -- the sort + run-length-encode strategy of unique for label sets too large to hash.
//
// Layout: B CTAs each own one contiguous chunk of the input (a multiple of the 2048-key tile).
// Per pass: (1) per-CTA digit histogram, (2) exclusive scan of the digit-major [256][B] table,
// (3) scatter: each CTA walks its chunk tile by tile; inside a tile keys are ranked with
// warp-level match.any (stable: warp w owns keys [256w, 256w+256) of the tile, item k lane l is
// key 256w + 32k + l), warp digit counters are combined across warps, and keys go straight to
// their global position.
#include "frb_internal.cuh"

namespace frb {

static constexpr int SORT_THREADS = 256;
static constexpr int SORT_ITEMS = 8;
static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048
static constexpr int SORT_WARPS = SORT_THREADS / 32;

static inline int sort_blocks(size_t n) {
  size_t need = (n + SORT_TILE - 1) / SORT_TILE;
  size_t cap = (size_t)sm_count() * 4;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
static inline size_t sort_chunk(size_t n, int blocks) {
  size_t c = (n + blocks - 1) / blocks;
  return align_up(c < 1 ? 1 : c, SORT_TILE);
}

size_t radix_hist_bytes(size_t n) { return (size_t)256 * sort_blocks(n) * sizeof(u64) + 256; }

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const K* __restrict__ keys, size_t n, size_t chunk, int shift, unsigned int dmask, u64* __restrict__ hist) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  const int lane = threadIdx.x & 31;
  if (beg < end) {
    const size_t end_r = beg + align_up(end - beg, 32);
    for (size_t i = beg + threadIdx.x; i < end_r; i += SORT_THREADS) {
      const bool valid = i < end;
      const unsigned int m = __ballot_sync(FRB_FULL_MASK, valid);
      if (valid) {
        const unsigned int d = (unsigned int)(keys[i] >> shift) & dmask;
        const unsigned int peers = __match_any_sync(m, d);
        if (lane == __ffs(peers) - 1) atomicAdd(&h[d], (unsigned int)__popc(peers));
      }
    }
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = (u64)h[threadIdx.x];
}

__global__ void __launch_bounds__(1024) k_scan_exclusive(u64* __restrict__ data, size_t M, u64* total_out) {
  __shared__ u64 warp_excl[32];
  __shared__ u64 s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 carry = 0;  // every thread tracks the same running total
  for (size_t base = 0; base < M; base += 1024) {
    const size_t i = base + threadIdx.x;
    const u64 v = i < M ? data[i] : 0ull;
    u64 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(FRB_FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_excl[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const u64 ws = warp_excl[lane];
      u64 wincl = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(FRB_FULL_MASK, wincl, o);
        if (lane >= o) wincl += t;
      }
      warp_excl[lane] = wincl - ws;
      if (lane == 31) s_total = wincl;
    }
    __syncthreads();
    if (i < M) data[i] = incl - v + warp_excl[warp] + carry;
    carry += s_total;
    __syncthreads();  // warp_excl / s_total are rewritten by the next iteration
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

int scan_exclusive_u64(u64* data, size_t M, u64* total_out, cudaStream_t st) {
  k_scan_exclusive<<<1, 1024, 0, st>>>(data, M, total_out);
  return check_launch("k_scan_exclusive");
}

template <typename K, bool HAS_V>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const K* __restrict__ kin, const u64* __restrict__ vin, K* __restrict__ kout, u64* __restrict__ vout,
               size_t n, size_t chunk, int shift, unsigned int dmask, const u64* __restrict__ offs) {
  __shared__ u64 base[256];
  __shared__ unsigned int wc[SORT_WARPS][256];
  __shared__ unsigned int tile_tot[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  base[threadIdx.x] = offs[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) wc[w][threadIdx.x] = 0;
  __syncthreads();
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  for (size_t tile = beg; tile < end; tile += SORT_TILE) {
    K key[SORT_ITEMS];
    unsigned int rank[SORT_ITEMS];
    unsigned int dig[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
      const size_t i = tile + (size_t)warp * (32 * SORT_ITEMS) + k * 32 + lane;
      const bool valid = i < end;
      key[k] = valid ? kin[i] : (K)0;
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
      for (int w = 0; w < SORT_WARPS; w++) {
        const unsigned int t = wc[w][threadIdx.x];
        wc[w][threadIdx.x] = run;
        run += t;
      }
      tile_tot[threadIdx.x] = run;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
      const size_t i = tile + (size_t)warp * (32 * SORT_ITEMS) + k * 32 + lane;
      if (i < end) {
        const size_t pos = (size_t)(base[dig[k]] + wc[warp][dig[k]] + rank[k]);
        kout[pos] = key[k];
        if (HAS_V) vout[pos] = vin[i];
      }
    }
    __syncthreads();
    base[threadIdx.x] += tile_tot[threadIdx.x];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) wc[w][threadIdx.x] = 0;
    __syncthreads();
  }
}

template <typename K>
static int radix_sort_impl(K* keys, u64* vals, K* keys_alt, u64* vals_alt, size_t n, int begin_bit, int end_bit,
                           void* hist_ws, cudaStream_t st) {
  if (n <= 1 || end_bit <= begin_bit) return FRB_OK;
  const int blocks = sort_blocks(n);
  const size_t chunk = sort_chunk(n, blocks);
  u64* hist = (u64*)hist_ws;
  K* kin = keys;
  K* kout = keys_alt;
  u64* vin = vals;
  u64* vout = vals_alt;
  int rc;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    const int nb = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
    const unsigned int dmask = (1u << nb) - 1u;
    k_sort_hist<K><<<blocks, SORT_THREADS, 0, st>>>(kin, n, chunk, shift, dmask, hist);
    if ((rc = check_launch("k_sort_hist"))) return rc;
    if ((rc = scan_exclusive_u64(hist, (size_t)256 * blocks, nullptr, st))) return rc;
    if (vals)
      k_sort_scatter<K, true><<<blocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, chunk, shift, dmask, hist);
    else
      k_sort_scatter<K, false><<<blocks, SORT_THREADS, 0, st>>>(kin, nullptr, kout, nullptr, n, chunk, shift, dmask, hist);
    if ((rc = check_launch("k_sort_scatter"))) return rc;
    K* tk = kin; kin = kout; kout = tk;
    u64* tv = vin; vin = vout; vout = tv;
  }
  if (kin != keys) {
    cudaMemcpyAsync(keys, kin, n * sizeof(K), cudaMemcpyDeviceToDevice, st);
    if (vals) cudaMemcpyAsync(vals, vin, n * sizeof(u64), cudaMemcpyDeviceToDevice, st);
  }
  return FRB_OK;
}

int radix_sort_u64(u64* keys, u64* vals, u64* keys_alt, u64* vals_alt, size_t n, int begin_bit, int end_bit,
                   void* hist_ws, cudaStream_t st) {
  return radix_sort_impl<u64>(keys, vals, keys_alt, vals_alt, n, begin_bit, end_bit, hist_ws, st);
}
int radix_sort_u32(uint32_t* keys, uint32_t* keys_alt, size_t n, int begin_bit, int end_bit, void* hist_ws,
                   cudaStream_t st) {
  return radix_sort_impl<uint32_t>(keys, nullptr, keys_alt, nullptr, n, begin_bit, end_bit, hist_ws, st);
}

}  // namespace frb
