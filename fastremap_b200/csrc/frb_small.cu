// frb_small.cu -- small arrays: renumber and unique(return_counts) as ONE cooperative kernel each.
//
// The volume-sized path of renumber is a dozen launches (table init, K1, two scans, the cooperative sort, K3) and two
// host round trips: fine for a gigabyte, but a 32^3 chunk then costs 0.4-0.5 ms where the reference's serial loop
// (fastremap.pyx:196-257) needs 0.1 ms -- and a drop-in gets called on small chunks too.  Below ~2^21 voxels the
// whole operation is one cooperatively launched kernel with grid barriers between its phases
//     init table | first index of every label (atomicMin) | compact | sort by first index | assign ranks | relabel
// writing every result (narrowed ids, ids at the input width, the two mapping arrays, the counters) into ONE buffer
// that the host fetches with one copy.  Same table format, same sort, same id arithmetic as the big path; only the
// streaming machinery (TMA ring, filters, logs) is left out: the data fits L2.
// A table that turns out too small raises meta[OVERFLOW] and the caller takes the volume-sized path.
#include <string.h>

#include "frb_internal.cuh"
#include "frb_coopsort.cuh"

namespace frb {

struct SmallArgs {
  const void* src;
  u64 n;
  Slot* slots;  // cap slots (direct tables: cap = 256 / 65536)
  u64 cap;
  int direct;
  u64* meta;
  CoopSortBuffers sort;  // three lists (+ ping-pong) of `capacity` entries
  u64 capacity;
  // renumber
  u64 start_bits;
  int skip_zero;
  void* wide_out;    // nullable: ids at the input width
  void* narrow_out;  // ids at the refit width (n * W bytes reserved)
  u64 lim1, lim2, lim4;
  void* keys_out;
  void* ids_out;
  int index_bits;
  // unique
  void* uniq_out;
  u64* counts_out;
  u64 flip;
  int key_bits;
};

template <int W>
__device__ __forceinline__ bool small_find(const SmallArgs& A, u64 key, u64* val) {
  if (A.direct) {
    const Slot s = ld_slot_cg(A.slots + key);
    *val = s.val;
    return s.key == key;
  }
  return table_find_from(A.slots, A.cap - 1, A.meta, key, home_slot(key, A.cap - 1), val);
}

// phases shared by both kernels ------------------------------------------------------------------------------------
__device__ __forceinline__ void small_init(const SmallArgs& A, u64 init_val, unsigned int& gen) {
  const u64 stride = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 s = tid; s < A.cap; s += stride) {
    A.slots[s].key = FRB_EMPTY;
    A.slots[s].val = init_val;
  }
  if (tid < FRB_META_WORDS) A.meta[tid] = tid == FRB_META_MISSING_INDEX ? FRB_EMPTY : (tid == FRB_META_SIDE_VAL ? init_val : 0ull);
  grid_barrier(A.sort.barrier, gen);
}

// every occupied slot (and the out-of-band all-ones key) -> (sort key, slot, label); returns the list length
template <typename KeyOf>
__device__ __forceinline__ u64 small_compact(const SmallArgs& A, KeyOf key_of, unsigned int& gen) {
  const u64 stride = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 s = tid; s <= A.cap; s += stride) {
    u64 k, v;
    bool live;
    if (s < A.cap) {
      const Slot sl = ld_slot_cg(A.slots + s);
      k = sl.key;
      v = sl.val;
      live = k != FRB_EMPTY;
    } else {
      k = FRB_EMPTY;
      v = __ldcg(&A.meta[FRB_META_SIDE_VAL]);
      live = __ldcg(&A.meta[FRB_META_SIDE_PRESENT]) != 0;
    }
    if (live) {
      const u64 pos = atomicAdd(&A.meta[FRB_META_LIST_COUNT], 1ull);
      if (pos < A.capacity) {
        A.sort.keys[pos] = key_of(k, v);
        A.sort.pay[pos] = s;
        A.sort.pay2[pos] = k;
      }
    }
  }
  grid_barrier(A.sort.barrier, gen);
  return __ldcg(&A.meta[FRB_META_LIST_COUNT]);
}

// ------------------------------------------------------------------------------------------ renumber
template <int W>
__global__ void __launch_bounds__(CS_THREADS) k_renumber_small(SmallArgs A) {
  const u64 stride = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int gen = 0;
  small_init(A, FRB_EMPTY, gen);
  // first flat index of every label; a voxel equal to its predecessor is dominated (fastremap.pyx:236-251 has the same shortcut)
  {
    unsigned int inserted = 0;
    for (u64 i = tid; i < A.n; i += stride) {
      const u64 key = ld_elem<W>(A.src, i);
      if ((A.skip_zero && key == 0) || (i > 0 && key == ld_elem<W>(A.src, i - 1))) continue;
      table_update<FRB_OP_MIN>(A.slots, A.cap - 1, A.direct, A.meta, key, FRB_RANK_TAG | i, &inserted);
    }
    if (inserted) atomicAdd(&A.meta[FRB_META_COUNT], (u64)inserted);
  }
  grid_barrier(A.sort.barrier, gen);
  const u64 count = __ldcg(&A.meta[FRB_META_COUNT]);
  if (__ldcg(&A.meta[FRB_META_OVERFLOW]) != 0 || count > A.capacity || (!A.direct && 2 * count > A.cap)) {  // the same in every thread
    if (tid == 0) A.meta[FRB_META_OVERFLOW] = 1;
    return;
  }
  const u64 nl = small_compact(A, [](u64, u64 v) { return v & ~FRB_RANK_TAG; }, gen);
  u64 *kin, *vin, *win;
  coop_radix_sort(A.sort, nl, 0, A.index_bits, gen, &kin, &vin, &win);
  for (u64 r = tid; r < nl; r += stride) {  // rank r -> table value; mapping arrays in id order
    const u64 s = __ldcg(vin + r), key = __ldcg(win + r);
    if (s == A.cap) A.meta[FRB_META_SIDE_VAL] = r;
    else A.slots[s].val = r;
    st_elem<W>(A.keys_out, r, key);
    st_elem<W>(A.ids_out, r, trunc_width(A.start_bits + r, W));  // counter arithmetic wraps in the array dtype (:225)
  }
  if (tid == 0) A.meta[FRB_META_ASSIGNED] = nl;
  grid_barrier(A.sort.barrier, gen);
  // relabel; the refit width follows from the label count (thresholds from fit_dtype, like frb_relabel_auto)
  int wo = nl < A.lim1 ? 1 : (nl < A.lim2 ? 2 : (nl < A.lim4 ? 4 : 0));
  if (wo >= W) wo = 0;
  for (u64 i = tid; i < A.n; i += stride) {
    const u64 key = ld_elem<W>(A.src, i);
    u64 id = 0;
    if (!(A.skip_zero && key == 0)) {
      u64 rank = 0;
      small_find<W>(A, key, &rank);
      id = trunc_width(A.start_bits + rank, W);
    }
    if (A.wide_out) st_elem<W>(A.wide_out, i, id);
    if (wo) st_elem_rt(A.narrow_out, wo, i, id);
  }
}

// ------------------------------------------------------------------------------------------ unique(return_counts)
template <int W>
__global__ void __launch_bounds__(CS_THREADS) k_unique_small(SmallArgs A) {
  const u64 threads = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int gen = 0;
  small_init(A, 0ull, gen);
  // a thread owns a contiguous piece of the array and adds one (label, run length) pair per run
  {
    const u64 per = (A.n + threads - 1) / threads;
    const u64 beg = tid * per < A.n ? tid * per : A.n;
    const u64 end = beg + per < A.n ? beg + per : A.n;
    unsigned int inserted = 0;
    u64 cur = 0, len = 0;
    for (u64 i = beg; i < end; i++) {
      const u64 key = ld_elem<W>(A.src, i);
      if (len && key == cur) {
        len++;
        continue;
      }
      if (len) table_update<FRB_OP_ADD>(A.slots, A.cap - 1, A.direct, A.meta, cur, len, &inserted);
      cur = key;
      len = 1;
    }
    if (len) table_update<FRB_OP_ADD>(A.slots, A.cap - 1, A.direct, A.meta, cur, len, &inserted);
    if (inserted) atomicAdd(&A.meta[FRB_META_COUNT], (u64)inserted);
  }
  grid_barrier(A.sort.barrier, gen);
  const u64 count = __ldcg(&A.meta[FRB_META_COUNT]);
  if (__ldcg(&A.meta[FRB_META_OVERFLOW]) != 0 || count > A.capacity || (!A.direct && 2 * count > A.cap)) {
    if (tid == 0) A.meta[FRB_META_OVERFLOW] = 1;
    return;
  }
  const u64 flip = A.flip;
  const u64 nl = small_compact(A, [flip](u64 k, u64) { return k ^ flip; }, gen);  // sort key: the label, sign bit flipped for signed dtypes
  // the payload the sort carries is the COUNT, not the slot
  for (u64 r = tid; r < nl; r += threads) {
    const u64 s = A.sort.pay[r];
    A.sort.pay[r] = s == A.cap ? __ldcg(&A.meta[FRB_META_SIDE_VAL]) : ld_slot_cg(A.slots + s).val;
  }
  grid_barrier(A.sort.barrier, gen);
  u64 *kin, *vin, *win;
  coop_radix_sort(A.sort, nl, 0, A.key_bits, gen, &kin, &vin, &win);
  for (u64 r = tid; r < nl; r += threads) {
    st_elem<W>(A.uniq_out, r, __ldcg(win + r));
    A.counts_out[r] = __ldcg(vin + r);
  }
}

static int small_grid(size_t n) {
  const size_t want = (n + 4 * CS_THREADS - 1) / (4 * CS_THREADS);  // >= 4 voxels per thread
  const size_t cap = (size_t)coop_sort_grid();
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace frb

using namespace frb;

extern "C" {

// table capacity used for an n-voxel array of this dtype: direct for 8 / 16-bit labels, else up to 2^17 slots
uint64_t frb_small_table_cap(size_t n, int dtype) {
  if (!dtype_ok(dtype)) return 0;
  const uint64_t direct = frb_table_direct_cap(dtype);
  if (direct) return direct;
  uint64_t cap = 1024;
  while (cap < 2 * (uint64_t)n && cap < (1ull << 17)) cap <<= 1;
  return cap;
}

// list entries the kernels can hold (labels beyond it raise meta[OVERFLOW])
size_t frb_small_list_capacity(size_t n, int dtype) {
  const uint64_t cap = frb_small_table_cap(n, dtype);
  if (frb_table_direct_cap(dtype)) return (size_t)cap + 2;  // a direct table can be full
  return (size_t)cap / 2 + 2;
}

static int small_setup(SmallArgs& A, const void* src, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, cudaStream_t st) {
  const uint64_t cap = frb_small_table_cap(n, dtype);
  size_t capacity = frb_small_list_capacity(n, dtype);
  const size_t need = align_up((size_t)cap * sizeof(Slot), 256) + 6 * align_up(capacity * sizeof(u64), 256) + align_up(coop_sort_aux_bytes(coop_sort_grid()), 256);
  if (!src || !meta || !ws || ws_bytes < need || n == 0 || !aligned16(ws)) return FRB_ERR_BAD_ARG;
  char* base = (char*)ws;
  const size_t lb = align_up(capacity * sizeof(u64), 256);
  A.src = src;
  A.n = n;
  A.slots = (Slot*)base;
  base += align_up((size_t)cap * sizeof(Slot), 256);
  A.cap = cap;
  A.direct = frb_table_direct_cap(dtype) ? 1 : 0;
  A.meta = (u64*)meta;
  A.sort.keys = (u64*)(base);
  A.sort.pay = (u64*)(base + lb);
  A.sort.keys_alt = (u64*)(base + 2 * lb);
  A.sort.pay_alt = (u64*)(base + 3 * lb);
  A.sort.pay2 = (u64*)(base + 4 * lb);
  A.sort.pay2_alt = (u64*)(base + 5 * lb);
  A.sort.hist = (unsigned int*)(base + 6 * lb);
  A.sort.barrier = (unsigned int*)(base + 6 * lb + (size_t)2 * coop_sort_grid() * 256 * sizeof(unsigned int));
  A.capacity = capacity;
  cudaMemsetAsync(A.sort.barrier, 0, 256, st);
  return FRB_OK;
}

// workspace: [table slots | 6 sort lists | histograms + barrier]; outputs are separate (the caller lays them out
// contiguously so that few copies bring everything back)
size_t frb_small_ws_bytes(size_t n, int dtype) {
  const uint64_t cap = frb_small_table_cap(n, dtype);
  const size_t capacity = frb_small_list_capacity(n, dtype);
  return align_up((size_t)cap * sizeof(Slot), 256) + 6 * align_up(capacity * sizeof(u64), 256) + align_up(coop_sort_aux_bytes(coop_sort_grid()), 256);
}

int frb_renumber_small(const void* src, size_t n, int dtype, int64_t start, int skip_zero, void* wide_out, void* narrow_out,
                       uint64_t lim1, uint64_t lim2, uint64_t lim4, void* keys_out, void* ids_out, uint64_t* meta, void* ws,
                       size_t ws_bytes, void* stream) {
  SmallArgs A;
  memset(&A, 0, sizeof(A));
  cudaStream_t st = (cudaStream_t)stream;
  if (!dtype_ok(dtype) || !narrow_out || !keys_out || !ids_out || small_setup(A, src, n, dtype, meta, ws, ws_bytes, st) != FRB_OK) {
    set_error("frb_renumber_small: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  const int w = dtype_width(dtype);
  A.start_bits = trunc_width((u64)start, w);
  A.skip_zero = skip_zero ? 1 : 0;
  A.wide_out = wide_out;
  A.narrow_out = narrow_out;
  A.lim1 = lim1;
  A.lim2 = lim2;
  A.lim4 = lim4;
  A.keys_out = keys_out;
  A.ids_out = ids_out;
  int bits = 1;
  while (bits < 63 && ((n - 1) >> bits) != 0) bits++;
  A.index_bits = bits;
  const void* kernel;
  switch (w) {
    case 1: kernel = (const void*)k_renumber_small<1>; break;
    case 2: kernel = (const void*)k_renumber_small<2>; break;
    case 4: kernel = (const void*)k_renumber_small<4>; break;
    default: kernel = (const void*)k_renumber_small<8>; break;
  }
  void* params[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(small_grid(n)), dim3(CS_THREADS), params, 0, st);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FRB_ERR_CUDA; }
  return check_launch("k_renumber_small");
}

int frb_unique_small(const void* src, size_t n, int dtype, void* uniq_out, uint64_t* counts_out, uint64_t* meta, void* ws,
                     size_t ws_bytes, void* stream) {
  SmallArgs A;
  memset(&A, 0, sizeof(A));
  cudaStream_t st = (cudaStream_t)stream;
  if (!dtype_ok(dtype) || !uniq_out || !counts_out || small_setup(A, src, n, dtype, meta, ws, ws_bytes, st) != FRB_OK) {
    set_error("frb_unique_small: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  const int w = dtype_width(dtype);
  A.uniq_out = uniq_out;
  A.counts_out = (u64*)counts_out;
  A.flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  A.key_bits = 8 * w;
  const void* kernel;
  switch (w) {
    case 1: kernel = (const void*)k_unique_small<1>; break;
    case 2: kernel = (const void*)k_unique_small<2>; break;
    case 4: kernel = (const void*)k_unique_small<4>; break;
    default: kernel = (const void*)k_unique_small<8>; break;
  }
  void* params[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(small_grid(n)), dim3(CS_THREADS), params, 0, st);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FRB_ERR_CUDA; }
  return check_launch("k_unique_small");
}

}  // extern "C"
