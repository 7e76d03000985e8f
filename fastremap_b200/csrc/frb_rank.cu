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
//   k_table_scan        scans the slots, compacts (first index - index_base, slot) of tagged values
//   k_rank_sort_assign  ONE persistent kernel: stable LSD radix sort (8-bit digits) of the list with
//                       a grid barrier between the histogram and scatter phases of every pass,
//                       then rank -> table value, (label, id) -> mapping arrays.
#include "frb_internal.cuh"
#include "frb_coopsort.cuh"
#include "frb_scan.cuh"

namespace frb {

static constexpr u64 RANK_TAG = 0x8000000000000000ull;

// ---------------------------------------------------------------------------------- export
// pending = tagged first index inside [index_base, index_base + index_count): a chunk pipeline that
// restarts after a table overflow may already hold (incomplete) entries of LATER chunks
struct PendingSel {
  u64 index_base, index_count;
  u64* key_out;    // first index - index_base (the sort key)
  u64* slot_out;   // slot index (payload 1)
  u64* label_out;  // label bits (payload 2: carried through the sort so that the assign phase has no dependent random read)
  __device__ __forceinline__ bool pick(u64 key, u64 val) const {
    return key != FRB_EMPTY && (val & RANK_TAG) && ((val & ~RANK_TAG) - index_base) < index_count;
  }
  __device__ __forceinline__ u64 fetch(u64, u64) const { return 0; }
  __device__ __forceinline__ void emit(u64 pos, u64 key, u64 val, u64 slot, u64) const {
    key_out[pos] = (val & ~RANK_TAG) - index_base;
    slot_out[pos] = slot;
    label_out[pos] = key;
  }
};
__global__ void k_pending_side(u64 cap, u64* meta, PendingSel sel, u64 capacity) {
  if (threadIdx.x == 0 && meta[FRB_META_SIDE_PRESENT] && sel.pick(0, meta[FRB_META_SIDE_VAL])) {
    const u64 pos = atomicAdd(&meta[FRB_META_LIST_COUNT], 1ull);
    if (pos < capacity) sel.emit(pos, FRB_EMPTY, meta[FRB_META_SIDE_VAL], cap, 0);  // "slot cap" = the out-of-band all-ones key
  }
}

// ---------------------------------------------------------------------------------- sort + assign
struct RankArgs {
  CoopSortBuffers sort;  // keys = first index - base, pay = slot index: the list written by k_export_pending
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
__global__ void __launch_bounds__(CS_THREADS) k_rank_sort_assign(RankArgs A) {
  const u64 n = __ldcg(&A.meta[FRB_META_LIST_COUNT]);
  const u64 assigned = __ldcg(&A.meta[FRB_META_ASSIGNED]);
  // an overflowed build (K1 aborted) must not be ranked: the list is incomplete.  Every CTA takes the
  // same decision: n / assigned are stable, and OVERFLOW is only ever raised below when the second
  // condition already holds.
  if (__ldcg(&A.meta[FRB_META_OVERFLOW]) != 0 || n > A.capacity || assigned + n > A.capacity) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(&A.meta[FRB_META_OVERFLOW], 1ull);
    return;
  }
  unsigned int generation = 0;
  u64* kin;
  u64* vin;
  u64* win;
  coop_radix_sort(A.sort, n, 0, A.index_bits, generation, &kin, &vin, &win);

  // ---- rank r of the sorted list -> table value `assigned + r`; mapping arrays in id order.
  // Coalesced reads of (slot, label), one scattered fire-and-forget store per label.
  for (u64 r = (u64)blockIdx.x * CS_THREADS + threadIdx.x; r < n; r += (u64)gridDim.x * CS_THREADS) {
    const u64 s = __ldcg(vin + r);
    const u64 key = __ldcg(win + r);
    const u64 rank = assigned + r;
    if (s == A.cap) A.meta[FRB_META_SIDE_VAL] = rank;  // the out-of-band all-ones label
    else A.slots[s].val = rank;
    st_elem<W>(A.keys_out, rank, key);
    st_elem<W>(A.ids_out, rank, trunc_width(A.start_bits + rank, W));  // counter arithmetic wraps in the array dtype (:225)
  }
  grid_barrier(A.sort.barrier, generation);  // every CTA has read `assigned` long before this point
  if (blockIdx.x == 0 && threadIdx.x == 0) A.meta[FRB_META_ASSIGNED] = assigned + n;
}

static int rank_grid() { return coop_sort_grid(); }

}  // namespace frb

using namespace frb;

extern "C" {

// workspace: 6 lists of `capacity` uint64 (sort key, 2 payloads, their ping-pong buffers) + histograms + barrier word
size_t frb_renumber_rank_ws_bytes(size_t capacity) {
  const size_t c = capacity < 1 ? 1 : capacity;
  return 6 * align_up(c * sizeof(u64), 256) + align_up(coop_sort_aux_bytes(coop_sort_grid()), 256);
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
  const size_t hb = (size_t)2 * coop_sort_grid() * 256 * sizeof(unsigned int);
  char* base = (char*)ws;
  RankArgs A;
  A.sort.keys = (u64*)(base);
  A.sort.pay = (u64*)(base + lb);
  A.sort.keys_alt = (u64*)(base + 2 * lb);
  A.sort.pay_alt = (u64*)(base + 3 * lb);
  A.sort.pay2 = (u64*)(base + 4 * lb);
  A.sort.pay2_alt = (u64*)(base + 5 * lb);
  A.sort.hist = (unsigned int*)(base + 6 * lb);
  A.sort.barrier = (unsigned int*)(base + 6 * lb + hb);
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
  cudaMemsetAsync(A.sort.barrier, 0, 256, st);
  PendingSel sel;
  sel.index_base = index_base;
  sel.index_count = index_count;
  sel.key_out = A.sort.keys;
  sel.slot_out = A.sort.pay;
  sel.label_out = A.sort.pay2;
  int rc = launch_table_scan(slots, cap, &A.meta[FRB_META_LIST_COUNT], (u64)capacity, sel, "k_table_scan<pending>", st);
  if (rc) return rc;
  k_pending_side<<<1, 32, 0, st>>>(cap, A.meta, sel, (u64)capacity);
  if ((rc = check_launch("k_pending_side"))) return rc;

  const void* kernel;
  switch (w) {
    case 1: kernel = (const void*)k_rank_sort_assign<1>; break;
    case 2: kernel = (const void*)k_rank_sort_assign<2>; break;
    case 4: kernel = (const void*)k_rank_sort_assign<4>; break;
    default: kernel = (const void*)k_rank_sort_assign<8>; break;
  }
  void* params[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(rank_grid()), dim3(CS_THREADS), params, 0, st);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return FRB_ERR_CUDA;
  }
  return check_launch("k_rank_sort_assign");
}

}  // extern "C"
