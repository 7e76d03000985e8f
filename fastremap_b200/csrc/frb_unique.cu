// frb_unique.cu -- K4/K5: unique(return_counts) strategies.
//
// Replaces _unique_via_array (fastremap.pyx:1061-1138), _unique_via_shifted_array (:984-1001),
// _unique_via_renumber (:1003-1022) and _unique_via_sort (:1024-1059).  All strategies produce
// the same (sorted labels, counts); the host picks by label range / cardinality:
//   dense  counts[label - min] += 1, then an ordered compaction of the non-empty bins
//   hash   (label -> count) table, then a sort of the L distinct labels
//   sort   radix sort of all N voxels, then run-length encode
// The counting kernel aggregates before it touches memory (per-lane repeat counting in registers,
// per-warp shared-memory (label, count) hash, then one global atomic per live entry): see k_hist.
#include "frb_internal.cuh"
#include "frb_build.cuh"
#include "frb_coopsort.cuh"

namespace frb {

struct DenseSink {
  u64* counts;
  u64 min_bits;
  int w;
  __device__ __forceinline__ void operator()(u64 key, unsigned int c, unsigned int*) const {
    atomicAdd(&counts[trunc_width(key - min_bits, w)], (u64)c);
  }
  // two-step form used by the drain: nothing to look up for a dense bin
  __device__ __forceinline__ Slot peek(u64) const { Slot s; s.key = 0; s.val = 0; return s; }
  __device__ __forceinline__ void commit(u64 key, unsigned int c, const Slot&, unsigned int* ins) const { (*this)(key, c, ins); }
};
struct HashSink {
  Slot* slots;
  u64 mask;
  int direct;
  u64* meta;
  __device__ __forceinline__ void operator()(u64 key, unsigned int c, unsigned int* inserted) const {
    table_update<FRB_OP_ADD>(slots, mask, direct, meta, key, (u64)c, inserted);
  }
  // two-step form used by the drain: the home slots of all of a lane's entries are read first (the
  // L2 round trips overlap), then a label already sitting in its home slot costs one fire-and-forget
  // atomic; anything else takes the probing path
  __device__ __forceinline__ Slot peek(u64 key) const {
    return ld_slot_cg_batch(slots + (direct ? key : home_slot(key, mask)));
  }
  __device__ __forceinline__ void commit(u64 key, unsigned int c, const Slot& home, unsigned int* inserted) const {
    if (home.key == key && key != FRB_EMPTY) atomicAdd(&slots[direct ? key : home_slot(key, mask)].val, (u64)c);
    else table_update<FRB_OP_ADD>(slots, mask, direct, meta, key, (u64)c, inserted);
  }
};

static constexpr int HIST_CACHE = 2048;        // (label, partial count) entries per CTA
static constexpr int HIST_SYNC_TILES = 8;      // the consumer warps meet this often ...
static constexpr int HIST_FLUSH_SYNCS = 8;     // ... and flush the table to memory at least every 8th time
static constexpr int HIST_FLUSH_CLAIMS = 768;  // or as soon as this many of its entries are taken

// Counting pass.  Three levels of aggregation sit between a voxel and a global atomic:
//   lane   remembers the 16-byte vector it saw directly above (rows of a warp are 4 KiB apart, rows
//          that lie above one another share a pattern slot: hist_slot) and a repeat count; an
//          identical vector just bumps the count
//          (7 integer instructions for V voxels).  A different vector flushes the old one: equal
//          neighbours inside the vector are merged and (label, multiplicity * repeats) is added to
//   CTA    a shared-memory hash of (label, partial count) shared by the 8 consumer warps (they walk
//          adjacent rows of the volume, i.e. the same labels), updated with shared-memory atomics.
//          Entries are claimed with CAS and never change label between flushes, so a lane can add
//          to an entry it has seen without locking; a label that finds both of its two candidate
//          entries taken goes straight to memory.  Every few tiles the consumer warps meet at a
//          named barrier and, on a schedule or when the table is filling up, drain it with one
//   global atomic per live entry (dense bins or the label table, see the sinks above); the home-slot
//          reads of a thread's entries are issued together so their L2 round trips overlap.
// pattern slot of row u: rows that lie directly above one another (q * 4 KiB apart, q = DU) share a slot
template <int DU>
__device__ __forceinline__ constexpr int hist_slot(int u) { return DU == 1 ? 0 : DU == 2 ? (u & 1) : DU == 3 ? (u == 3 ? 0 : u) : u; }

template <int W, typename Sink, int IL, int DU, bool VAR_TILE>
__global__ void __launch_bounds__(STREAM_THREADS, 3) k_hist(const uint4* __restrict__ a, size_t n, Sink sink, u64* count_target, int run,
                                                            int tile_vecs) {
  constexpr int V = 16 / W;
  __shared__ u64 s_key[HIST_CACHE];
  __shared__ unsigned int s_cnt[HIST_CACHE];
  __shared__ unsigned int s_claims;
  const size_t nvec = n / V;
  const int warp = threadIdx.x >> 5;
  const bool consumer = warp < STREAM_WARPS;
  unsigned int inserted = 0;
  for (int i = threadIdx.x; i < HIST_CACHE; i += STREAM_THREADS) { s_key[i] = FRB_EMPTY; s_cnt[i] = 0; }
  if (threadIdx.x == 0) s_claims = 0;
  __syncthreads();

  unsigned int claimed = 0;  // entries this lane has claimed since the consumer warps last met
  auto add = [&](u64 key, unsigned int c) {
    if (key != FRB_EMPTY) {
      unsigned int e = (unsigned int)(hash_key(key) >> 48) & (HIST_CACHE - 1);
#pragma unroll
      for (int attempt = 0; attempt < 2; attempt++, e ^= 1u) {
        u64 ck = s_key[e];
        if (ck == FRB_EMPTY) {
          const u64 old = atomicCAS(&s_key[e], (u64)FRB_EMPTY, key);
          ck = old == FRB_EMPTY ? key : old;
          claimed += old == FRB_EMPTY ? 1u : 0u;
        }
        if (ck == key) {
          atomicAdd(&s_cnt[e], c);
          return;
        }
      }
    }
    sink(key, c, &inserted);
  };
  auto flush_vec = [&](const uint4& p, unsigned int reps) {
    u64 key = vec_get<W>(p, 0);
    unsigned int c = 1;
#pragma unroll
    for (int j = 1; j < V; j++) {
      const u64 kj = vec_get<W>(p, j);
      if (kj == key) {
        c++;
      } else {
        add(key, c * reps);
        key = kj;
        c = 1;
      }
    }
    add(key, c * reps);
  };

  uint4 pat[STREAM_ROWS];
  unsigned int rep[STREAM_ROWS];
#pragma unroll
  for (int u = 0; u < STREAM_ROWS; u++) { pat[u] = make_uint4(0, 0, 0, 0); rep[u] = 0; }
  int syncs_since_flush = 0;

  auto row_fn = [&](int u, const uint4& cur, bool valid, size_t, u64, bool) {
    if (!valid) return;
    const int sl = hist_slot<DU>(u);  // a constant once the row loop is unrolled
    const uint4 l = pat[sl];
    const uint32_t d = (cur.x ^ l.x) | (cur.y ^ l.y) | (cur.z ^ l.z) | (cur.w ^ l.w);
    if (d == 0 && rep[sl] != 0) {
      rep[sl]++;
    } else {
      if (rep[sl] != 0) flush_vec(l, rep[sl]);
      pat[sl] = cur;
      rep[sl] = 1;
    }
  };
  // all consumer threads, between two consumer barriers: nobody is inside add()
  auto drain = [&]() {
    constexpr int BATCH = 4;  // home-slot reads in flight per thread (registers: 7 per entry)
    constexpr int NCONS = STREAM_WARPS * 32;
#pragma unroll 1
    for (int q0 = 0; q0 < HIST_CACHE / NCONS; q0 += BATCH) {
      u64 k[BATCH];
      unsigned int cnt[BATCH];
      Slot home[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; q++) {
        const int e = (q0 + q) * NCONS + threadIdx.x;
        k[q] = s_key[e];
        cnt[q] = s_cnt[e];
        if (k[q] != FRB_EMPTY) {
          home[q] = sink.peek(k[q]);
          s_key[e] = FRB_EMPTY;
          s_cnt[e] = 0;
        }
      }
#pragma unroll
      for (int q = 0; q < BATCH; q++)
        if (k[q] != FRB_EMPTY) sink.commit(k[q], cnt[q], home[q], &inserted);
    }
  };
  auto after_tile = [&](size_t k) {
    if ((k % HIST_SYNC_TILES) != HIST_SYNC_TILES - 1) return;
    if (claimed) { atomicAdd(&s_claims, claimed); claimed = 0; }
    consumer_bar();
    const bool flush = s_claims >= HIST_FLUSH_CLAIMS || ++syncs_since_flush >= HIST_FLUSH_SYNCS;  // the same in every warp
    consumer_bar();
    if (flush) {
      drain();
      if (threadIdx.x == 0) s_claims = 0;
      syncs_since_flush = 0;
      consumer_bar();
    }
  };

  const TileMap map(nvec, (size_t)run, VAR_TILE ? (size_t)tile_vecs : (size_t)STREAM_TILE_VECS);
  stream_rows<W, false, false, IL, VAR_TILE>(a, nvec, map, row_fn, [](const uint4*, size_t) {}, after_tile);
  if (consumer) {
#pragma unroll
    for (int u = 0; u < STREAM_ROWS; u++)
      if (rep[u] != 0) flush_vec(pat[u], rep[u]);
    consumer_bar();
    drain();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = nvec * V; i < n; i++) sink(ld_elem<W>(a, i), 1u, &inserted);
  if (count_target) block_add_to(count_target, inserted);
}

// ------------------------------------------------------------------ ordered compaction of bins
static inline int compact_blocks(size_t n) {
  size_t need = (n + 2047) / 2048;
  size_t cap = (size_t)sm_count() * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
static inline size_t compact_chunk(size_t n, int blocks) {
  size_t c = (n + blocks - 1) / blocks;
  return align_up(c < 1 ? 1 : c, 256);
}

// block-wide exclusive prefix of a 0/1 flag (256 threads); returns the prefix, *total = block sum
__device__ __forceinline__ unsigned int block_flag_scan(bool flag, unsigned int* total) {
  __shared__ unsigned int wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int b = __ballot_sync(FRB_FULL_MASK, flag);
  if (lane == 0) wsum[warp] = __popc(b);
  __syncthreads();
  unsigned int pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const unsigned int s = wsum[w];
    if (w < warp) pre += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return pre + __popc(b & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256)
k_bins_count(const u64* __restrict__ counts, u64 bins, size_t chunk, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > bins) end = bins;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += counts[i] != 0 ? 1 : 0;
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  c = __reduce_add_sync(FRB_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}

template <int W>
__global__ void __launch_bounds__(256)
k_bins_emit(const u64* __restrict__ counts, u64 bins, size_t chunk, const u64* __restrict__ block_offs, u64 min_bits,
            void* __restrict__ uniq_out, u64* __restrict__ counts_out, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > bins) end = bins;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const u64 c = i < end ? counts[i] : 0;
    unsigned int tot;
    const unsigned int pre = block_flag_scan(c != 0, &tot);
    if (c != 0) {
      const u64 pos = run + pre;
      if (pos < capacity) {
        st_elem<W>(uniq_out, pos, min_bits + (u64)i);
        counts_out[pos] = c;
      }
    }
    run += tot;
  }
}

// ------------------------------------------------------------------ indices(arr, value): ordered positions of a value
template <int W>
__global__ void __launch_bounds__(256)
k_match_count(const void* __restrict__ a, size_t n, size_t chunk, u64 value, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += ld_elem<W>(a, i) == value ? 1 : 0;
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  c = __reduce_add_sync(FRB_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}

template <int W>
__global__ void __launch_bounds__(256)
k_match_emit(const void* __restrict__ a, size_t n, size_t chunk, u64 value, const u64* __restrict__ block_offs,
             u64* __restrict__ out, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const bool hit = i < end && ld_elem<W>(a, i) == value;
    unsigned int tot;
    const unsigned int pre = block_flag_scan(hit, &tot);
    if (hit && run + pre < capacity) out[run + pre] = (u64)i;
    run += tot;
  }
}

// ------------------------------------------------------------------ point_cloud: foreground voxels as (label, index) pairs
// ordered compaction of the non-zero voxels: label bits (sign bit flipped for signed dtypes, so that
// unsigned radix order is numeric order) and flat index, both uint64
template <int W>
__global__ void __launch_bounds__(256)
k_nonzero_count(const void* __restrict__ a, size_t n, size_t chunk, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += ld_elem<W>(a, i) != 0 ? 1 : 0;
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  c = __reduce_add_sync(FRB_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}

template <int W>
__global__ void __launch_bounds__(256)
k_nonzero_emit(const void* __restrict__ a, size_t n, size_t chunk, const u64* __restrict__ block_offs, u64 flip,
               u64* __restrict__ labels_out, u64* __restrict__ index_out, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const u64 v = i < end ? ld_elem<W>(a, i) : 0;
    unsigned int tot;
    const unsigned int pre = block_flag_scan(v != 0, &tot);
    if (v != 0 && run + pre < capacity) {
      labels_out[run + pre] = v ^ flip;
      index_out[run + pre] = (u64)i;
    }
    run += tot;
  }
}

// flat C-order index -> (i, j[, k]) as uint16 rows
__global__ void __launch_bounds__(256)
k_unravel_u16(const u64* __restrict__ index, size_t n, int ndim, u64 d1, u64 d2, uint16_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const u64 f = index[p];
    if (ndim == 2) {
      out[2 * p] = (uint16_t)(f / d1);
      out[2 * p + 1] = (uint16_t)(f % d1);
    } else {
      const u64 ij = f / d2;
      out[3 * p] = (uint16_t)(ij / d1);
      out[3 * p + 1] = (uint16_t)(ij % d1);
      out[3 * p + 2] = (uint16_t)(f % d2);
    }
  }
}

// ------------------------------------------------------------------ sorted (label, count) pairs
// One persistent kernel: flip the sign bit of signed labels (so that unsigned digit order is the
// dtype's order), sort the pairs by label (frb_coopsort.cuh), emit typed labels + values.
struct SortPairsArgs {
  CoopSortBuffers sort;
  const u64* n_dev;  // optional: the list length lives on the device (<= n)
  u64 n;
  u64 flip;
  int key_bits;
  void* uniq_out;
  u64* vals_out;
};

template <int W>
__global__ void __launch_bounds__(CS_THREADS) k_sort_emit_pairs(SortPairsArgs A) {
  if (A.n_dev) {
    const u64 nd = __ldcg(A.n_dev);
    A.n = nd < A.n ? nd : A.n;
  }
  const u64 stride = (u64)gridDim.x * CS_THREADS;
  unsigned int generation = 0;
  if (A.flip) {
    for (u64 i = (u64)blockIdx.x * CS_THREADS + threadIdx.x; i < A.n; i += stride) A.sort.keys[i] ^= A.flip;
    grid_barrier(A.sort.barrier, generation);
  }
  u64* kin;
  u64* vin;
  coop_radix_sort(A.sort, A.n, 0, A.key_bits, generation, &kin, &vin);
  for (u64 i = (u64)blockIdx.x * CS_THREADS + threadIdx.x; i < A.n; i += stride) {
    st_elem<W>(A.uniq_out, i, __ldcg(kin + i) ^ A.flip);
    A.vals_out[i] = __ldcg(vin + i);
  }
}

// ------------------------------------------------------------------ sort + run-length encode
template <int W, typename K>
__global__ void __launch_bounds__(256) k_widen_flip(const void* __restrict__ a, size_t n, K* __restrict__ out, K flip) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (K)ld_elem<W>(a, i) ^ flip;
}

template <typename K>
__global__ void __launch_bounds__(256)
k_heads_count(const K* __restrict__ sorted, size_t n, size_t chunk, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  c = __reduce_add_sync(FRB_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}

// A run [start, end) of label k with output position p: the head writes uniq[p] and adds -start
// to counts[p]; the tail adds end.  counts_out is zeroed by the caller, so counts[p] = end - start.
template <int W, typename K>
__global__ void __launch_bounds__(256)
k_heads_emit(const K* __restrict__ sorted, size_t n, size_t chunk, const u64* __restrict__ block_offs, K flip,
             void* __restrict__ uniq_out, u64* __restrict__ counts_out, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const bool in = i < end;
    const K k = in ? sorted[i] : (K)0;
    const bool head = in && (i == 0 || k != sorted[i - 1]);
    const bool tail = in && (i + 1 == n || k != sorted[i + 1]);
    unsigned int tot;
    const unsigned int pre = block_flag_scan(head, &tot);
    // output position of the run this element belongs to = (#heads at or before i) - 1
    const u64 p = run + pre + (head ? 1 : 0) - 1;
    if (head && p < capacity) {
      st_elem<W>(uniq_out, p, (u64)(k ^ flip));
      atomicAdd(&counts_out[p], (u64)0 - (u64)i);
    }
    if (tail && p < capacity) atomicAdd(&counts_out[p], (u64)i + 1);
    run += tot;
  }
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_hist_dense(const void* a, size_t n, int dtype, uint64_t min_bits, uint64_t* counts, uint64_t bins, uint64_t row_bytes, void* stream) {
  if (!dtype_ok(dtype) || !counts || bins == 0 || (n && (!a || !aligned16(a)))) { set_error("frb_hist_dense: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  DenseSink sink;
  sink.counts = (u64*)counts;
  sink.min_bits = trunc_width(min_bits, w);
  sink.w = w;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)a;
  // the dense strategy is only ever forced (tests, comparisons): two geometries are enough
  const RowGeom geo = row_geom((size_t)row_bytes);
  const int tv = geo.tv;
#define FRB_KHD(WW, VV) k_hist<WW, DenseSink, 8, 4, VV><<<stream_grid(k_hist<WW, DenseSink, 8, 4, VV>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>( \
    p, n, sink, nullptr, fair_run(stream_grid(k_hist<WW, DenseSink, 8, 4, VV>, tiles), tiles), tv)
#define FRB_KHD_GEO(WW) if (tv != STREAM_TILE_VECS) { FRB_KHD(WW, true); } else { FRB_KHD(WW, false); }
  switch (w) {
    case 1: FRB_KHD_GEO(1) break;
    case 2: FRB_KHD_GEO(2) break;
    case 4: FRB_KHD_GEO(4) break;
    default: FRB_KHD_GEO(8) break;
  }
#undef FRB_KHD_GEO
#undef FRB_KHD
  return check_launch("k_hist<dense>");
}

int frb_hist_hash(const void* a, size_t n, int dtype, void* slots, uint64_t cap, uint64_t* meta, uint64_t row_bytes, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || (n && (!a || !aligned16(a)))) { set_error("frb_hist_hash: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  HashSink sink;
  sink.slots = (Slot*)slots;
  sink.mask = cap - 1;
  sink.direct = w <= 2 ? 1 : 0;
  sink.meta = (u64*)meta;
  if (sink.direct && cap != frb_table_direct_cap(dtype)) { set_error("frb_hist_hash: direct table needs cap == 2^bits"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)a;
  u64* ct = &((u64*)meta)[FRB_META_COUNT];
  const RowGeom geo = row_geom((size_t)row_bytes);
  const int tv = geo.tv;
#define FRB_KH(WW, II, DD, VV) k_hist<WW, HashSink, II, DD, VV><<<stream_grid(k_hist<WW, HashSink, II, DD, VV>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>( \
    p, n, sink, ct, fair_run(stream_grid(k_hist<WW, HashSink, II, DD, VV>, tiles), tiles), tv)
#define FRB_KH_GEO(WW)                                                        \
  if (geo.tv != STREAM_TILE_VECS) { FRB_KH(WW, 8, 4, true); }                 \
  else if (geo.il == 4) { FRB_KH(WW, 4, 1, false); }                          \
  else if (geo.il == 2) { FRB_KH(WW, 2, 1, false); }                          \
  else if (geo.il == 1) { FRB_KH(WW, 1, 1, false); }                          \
  else if (geo.du == 1) { FRB_KH(WW, 8, 1, false); }                          \
  else if (geo.du == 2) { FRB_KH(WW, 8, 2, false); }                          \
  else if (geo.du == 3) { FRB_KH(WW, 8, 3, false); }                          \
  else { FRB_KH(WW, 8, 4, false); }
  switch (w) {
    case 1: FRB_KH_GEO(1) break;
    case 2: FRB_KH_GEO(2) break;
    case 4: FRB_KH_GEO(4) break;
    default: FRB_KH_GEO(8) break;
  }
#undef FRB_KH_GEO
#undef FRB_KH
  return check_launch("k_hist<hash>");
}

size_t frb_compact_bins_ws_bytes(uint64_t bins) { return (size_t)compact_blocks((size_t)bins) * sizeof(u64) + 256; }

int frb_compact_bins_count(const uint64_t* counts, uint64_t bins, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!counts || !meta || !ws || bins == 0 || ws_bytes < frb_compact_bins_ws_bytes(bins)) { set_error("frb_compact_bins_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks((size_t)bins);
  const size_t chunk = compact_chunk((size_t)bins, blocks);
  k_bins_count<<<blocks, 256, 0, st>>>((const u64*)counts, bins, chunk, (u64*)ws);
  int rc = check_launch("k_bins_count");
  if (rc) return rc;
  return scan_exclusive_u64((u64*)ws, (size_t)blocks, &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_compact_bins_emit(const uint64_t* counts, uint64_t bins, int dtype, uint64_t min_bits, void* uniq_out,
                          uint64_t* counts_out, size_t capacity, const void* ws, void* stream) {
  if (!dtype_ok(dtype) || !counts || !ws || bins == 0 || (capacity && (!uniq_out || !counts_out))) { set_error("frb_compact_bins_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks((size_t)bins);
  const size_t chunk = compact_chunk((size_t)bins, blocks);
  const int w = dtype_width(dtype);
  const u64 mb = trunc_width(min_bits, w);
  const u64* c = (const u64*)counts;
  const u64* offs = (const u64*)ws;
  u64* co = (u64*)counts_out;
  switch (w) {
    case 1: k_bins_emit<1><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
    case 2: k_bins_emit<2><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
    case 4: k_bins_emit<4><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
    default: k_bins_emit<8><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
  }
  return check_launch("k_bins_emit");
}

size_t frb_indices_ws_bytes(size_t n) { return (size_t)compact_blocks(n) * sizeof(u64) + 256; }

int frb_indices_count(const void* a, size_t n, int dtype, uint64_t value_bits, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !meta || !ws || ws_bytes < frb_indices_ws_bytes(n)) { set_error("frb_indices_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const int w = dtype_width(dtype);
  const u64 v = trunc_width(value_bits, w);
  switch (w) {
    case 1: k_match_count<1><<<blocks, 256, 0, st>>>(a, n, chunk, v, (u64*)ws); break;
    case 2: k_match_count<2><<<blocks, 256, 0, st>>>(a, n, chunk, v, (u64*)ws); break;
    case 4: k_match_count<4><<<blocks, 256, 0, st>>>(a, n, chunk, v, (u64*)ws); break;
    default: k_match_count<8><<<blocks, 256, 0, st>>>(a, n, chunk, v, (u64*)ws); break;
  }
  int rc = check_launch("k_match_count");
  if (rc) return rc;
  return scan_exclusive_u64((u64*)ws, (size_t)blocks, &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_indices_emit(const void* a, size_t n, int dtype, uint64_t value_bits, uint64_t* out, size_t capacity, const void* ws, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !ws || (capacity && !out)) { set_error("frb_indices_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (capacity == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const int w = dtype_width(dtype);
  const u64 v = trunc_width(value_bits, w);
  const u64* offs = (const u64*)ws;
  switch (w) {
    case 1: k_match_emit<1><<<blocks, 256, 0, st>>>(a, n, chunk, v, offs, (u64*)out, capacity); break;
    case 2: k_match_emit<2><<<blocks, 256, 0, st>>>(a, n, chunk, v, offs, (u64*)out, capacity); break;
    case 4: k_match_emit<4><<<blocks, 256, 0, st>>>(a, n, chunk, v, offs, (u64*)out, capacity); break;
    default: k_match_emit<8><<<blocks, 256, 0, st>>>(a, n, chunk, v, offs, (u64*)out, capacity); break;
  }
  return check_launch("k_match_emit");
}

size_t frb_nonzero_ws_bytes(size_t n) { return (size_t)compact_blocks(n) * sizeof(u64) + 256; }

int frb_nonzero_count(const void* a, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !meta || !ws || ws_bytes < frb_nonzero_ws_bytes(n)) { set_error("frb_nonzero_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  switch (dtype_width(dtype)) {
    case 1: k_nonzero_count<1><<<blocks, 256, 0, st>>>(a, n, chunk, (u64*)ws); break;
    case 2: k_nonzero_count<2><<<blocks, 256, 0, st>>>(a, n, chunk, (u64*)ws); break;
    case 4: k_nonzero_count<4><<<blocks, 256, 0, st>>>(a, n, chunk, (u64*)ws); break;
    default: k_nonzero_count<8><<<blocks, 256, 0, st>>>(a, n, chunk, (u64*)ws); break;
  }
  int rc = check_launch("k_nonzero_count");
  if (rc) return rc;
  return scan_exclusive_u64((u64*)ws, (size_t)blocks, &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_nonzero_emit(const void* a, size_t n, int dtype, uint64_t* labels_out, uint64_t* index_out, size_t capacity,
                     const void* ws, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !ws || (capacity && (!labels_out || !index_out))) { set_error("frb_nonzero_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (capacity == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const int w = dtype_width(dtype);
  const u64 flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  const u64* offs = (const u64*)ws;
  switch (w) {
    case 1: k_nonzero_emit<1><<<blocks, 256, 0, st>>>(a, n, chunk, offs, flip, (u64*)labels_out, (u64*)index_out, capacity); break;
    case 2: k_nonzero_emit<2><<<blocks, 256, 0, st>>>(a, n, chunk, offs, flip, (u64*)labels_out, (u64*)index_out, capacity); break;
    case 4: k_nonzero_emit<4><<<blocks, 256, 0, st>>>(a, n, chunk, offs, flip, (u64*)labels_out, (u64*)index_out, capacity); break;
    default: k_nonzero_emit<8><<<blocks, 256, 0, st>>>(a, n, chunk, offs, flip, (u64*)labels_out, (u64*)index_out, capacity); break;
  }
  return check_launch("k_nonzero_emit");
}

// stable sort of n (key, value) uint64 pairs by the low key_bits bits of the key (many-launch LSD
// radix sort: lists as long as the volume).  workspace: 2 ping-pong lists + histograms
size_t frb_sort_pairs_u64_ws_bytes(size_t n) {
  const size_t c = n < 1 ? 1 : n;
  return 2 * align_up(c * sizeof(u64), 256) + radix_hist_bytes(c);
}
int frb_sort_pairs_u64(uint64_t* keys, uint64_t* vals, size_t n, int key_bits, void* ws, size_t ws_bytes, void* stream) {
  if ((n && (!keys || !vals)) || !ws || ws_bytes < frb_sort_pairs_u64_ws_bytes(n) || key_bits < 1 || key_bits > 64) { set_error("frb_sort_pairs_u64: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n <= 1) return FRB_OK;
  const size_t lb = align_up(n * sizeof(u64), 256);
  return radix_sort_u64((u64*)keys, (u64*)vals, (u64*)ws, (u64*)((char*)ws + lb), n, 0, key_bits, (char*)ws + 2 * lb, (cudaStream_t)stream);
}

int frb_unravel_u16(const uint64_t* index, size_t n, int ndim, uint64_t d1, uint64_t d2, uint16_t* out, void* stream) {
  if ((n && (!index || !out)) || (ndim != 2 && ndim != 3) || d1 == 0 || (ndim == 3 && d2 == 0)) { set_error("frb_unravel_u16: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  Grid g = grid_for(n, 256 * 4, 16);
  k_unravel_u16<<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>((const u64*)index, n, ndim, d1, d2, out);
  return check_launch("k_unravel_u16");
}

// workspace: 2 ping-pong lists of L uint64 + the cooperative sort's histograms and barrier word
size_t frb_sort_pairs_ws_bytes(size_t L) {
  const size_t c = L < 1 ? 1 : L;
  return 2 * align_up(c * sizeof(u64), 256) + align_up(coop_sort_aux_bytes(coop_sort_grid()), 256);
}

int frb_sort_table_pairs(uint64_t* keys, uint64_t* vals, size_t L, const uint64_t* count_dev, int dtype, void* uniq_out,
                         uint64_t* vals_out, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !ws || ws_bytes < frb_sort_pairs_ws_bytes(L) || (L && (!keys || !vals || !uniq_out || !vals_out))) {
    set_error("frb_sort_table_pairs: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (L == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const size_t lb = align_up(L * sizeof(u64), 256);
  SortPairsArgs A;
  A.sort.keys = (u64*)keys;
  A.sort.pay = (u64*)vals;
  A.sort.keys_alt = (u64*)ws;
  A.sort.pay_alt = (u64*)((char*)ws + lb);
  A.sort.pay2 = nullptr;
  A.sort.pay2_alt = nullptr;
  A.sort.hist = (unsigned int*)((char*)ws + 2 * lb);
  A.sort.barrier = (unsigned int*)((char*)ws + 2 * lb + (size_t)2 * coop_sort_grid() * 256 * sizeof(unsigned int));
  A.n = L;
  A.n_dev = (const u64*)count_dev;
  A.flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  A.key_bits = 8 * w;
  A.uniq_out = uniq_out;
  A.vals_out = (u64*)vals_out;
  cudaMemsetAsync(A.sort.barrier, 0, 256, st);
  const void* kernel;
  switch (w) {
    case 1: kernel = (const void*)k_sort_emit_pairs<1>; break;
    case 2: kernel = (const void*)k_sort_emit_pairs<2>; break;
    case 4: kernel = (const void*)k_sort_emit_pairs<4>; break;
    default: kernel = (const void*)k_sort_emit_pairs<8>; break;
  }
  void* params[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(coop_sort_grid()), dim3(CS_THREADS), params, 0, st);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return FRB_ERR_CUDA;
  }
  return check_launch("k_sort_emit_pairs");
}

// workspace: sorted keys (n * kw) | ping-pong (n * kw) | block counts | radix histograms;  kw = 4 (W<=4) or 8
static inline size_t usort_key_bytes(int dtype) { return dtype_width(dtype) == 8 ? 8 : 4; }
size_t frb_unique_sort_ws_bytes(size_t n, int dtype) {
  if (!dtype_ok(dtype)) return 0;
  const size_t c = n < 1 ? 1 : n;
  return 2 * align_up(c * usort_key_bytes(dtype), 256) + align_up((size_t)compact_blocks(c) * sizeof(u64), 256) + radix_hist_bytes(c);
}

int frb_unique_sort_count(const void* a, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !meta || !ws || ws_bytes < frb_unique_sort_ws_bytes(n, dtype)) { set_error("frb_unique_sort_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const size_t kb = usort_key_bytes(dtype);
  const size_t lb = align_up(n * kb, 256);
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  char* base = (char*)ws;
  u64* block_counts = (u64*)(base + 2 * lb);
  void* hist = base + 2 * lb + align_up((size_t)blocks * sizeof(u64), 256);
  const u64 flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  Grid g = grid_for(n, 256 * 8, 16);
  int rc;
  if (kb == 8) {
    u64* keys = (u64*)base;
    u64* alt = (u64*)(base + lb);
    k_widen_flip<8, u64><<<g.blocks, g.threads, 0, st>>>(a, n, keys, flip);
    if ((rc = check_launch("k_widen_flip"))) return rc;
    if ((rc = radix_sort_u64(keys, nullptr, alt, nullptr, n, 0, 64, hist, st))) return rc;
    k_heads_count<u64><<<blocks, 256, 0, st>>>(keys, n, chunk, block_counts);
  } else {
    uint32_t* keys = (uint32_t*)base;
    uint32_t* alt = (uint32_t*)(base + lb);
    switch (w) {
      case 1: k_widen_flip<1, uint32_t><<<g.blocks, g.threads, 0, st>>>(a, n, keys, (uint32_t)flip); break;
      case 2: k_widen_flip<2, uint32_t><<<g.blocks, g.threads, 0, st>>>(a, n, keys, (uint32_t)flip); break;
      default: k_widen_flip<4, uint32_t><<<g.blocks, g.threads, 0, st>>>(a, n, keys, (uint32_t)flip); break;
    }
    if ((rc = check_launch("k_widen_flip"))) return rc;
    if ((rc = radix_sort_u32(keys, alt, n, 0, 8 * w, hist, st))) return rc;
    k_heads_count<uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_counts);
  }
  if ((rc = check_launch("k_heads_count"))) return rc;
  return scan_exclusive_u64(block_counts, (size_t)blocks, &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_unique_sort_emit(size_t n, int dtype, void* uniq_out, uint64_t* counts_out, size_t capacity, const void* ws, void* stream) {
  if (!dtype_ok(dtype) || n == 0 || !ws || (capacity && (!uniq_out || !counts_out))) { set_error("frb_unique_sort_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (capacity == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const size_t kb = usort_key_bytes(dtype);
  const size_t lb = align_up(n * kb, 256);
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const char* base = (const char*)ws;
  const u64* block_offs = (const u64*)(base + 2 * lb);
  const u64 flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  cudaMemsetAsync(counts_out, 0, capacity * sizeof(u64), st);
  u64* co = (u64*)counts_out;
  if (kb == 8) {
    k_heads_emit<8, u64><<<blocks, 256, 0, st>>>((const u64*)base, n, chunk, block_offs, flip, uniq_out, co, capacity);
  } else {
    const uint32_t* keys = (const uint32_t*)base;
    switch (w) {
      case 1: k_heads_emit<1, uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_offs, (uint32_t)flip, uniq_out, co, capacity); break;
      case 2: k_heads_emit<2, uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_offs, (uint32_t)flip, uniq_out, co, capacity); break;
      default: k_heads_emit<4, uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_offs, (uint32_t)flip, uniq_out, co, capacity); break;
    }
  }
  return check_launch("k_heads_emit");
}

}  // extern "C"
