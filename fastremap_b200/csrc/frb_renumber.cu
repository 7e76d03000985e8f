// frb_renumber.cu -- K6 (fused min / max / pixel-pairs scan), K1 (first-index table build) and
// K2 (rank labels by first index -> ids) of the renumber path.
//
// Reference loops replaced: _minmax fastremap.pyx:77-93, _pixel_pairs :840-853, and the
// find/insert half of _renumber :236-251.  The reference numbers labels in encounter order with
// a serial counter; here id(label) = start + rank(first flat index of label), which is the same
// numbering computed without the serial dependence: K1 records min(first index) per label with
// 64-bit atomicMin, K2 sorts the L labels by that index.
#include "frb_internal.cuh"

namespace frb {

static constexpr int ROWS = 4;                        // 16-byte rows per warp per tile
static constexpr int TILE_VECS = 8 * ROWS * 32;       // 16-byte vectors per CTA tile (256 threads)

// ------------------------------------------------------------------------------------------ K6
template <int W, bool SIGNED>
__global__ void __launch_bounds__(256)
k_minmax_pairs(const uint4* __restrict__ a, size_t n, u64* __restrict__ out) {
  constexpr int V = 16 / W;
  typedef long long i64;
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t tiles = (nvec + TILE_VECS - 1) / TILE_VECS;
  u64 umin = FRB_EMPTY, umax = 0;
  i64 smin = 0x7FFFFFFFFFFFFFFFll, smax = (i64)0x8000000000000000ull;
  unsigned long long pairs = 0, nonzero = 0;

  auto consume = [&](u64 key, u64 prev, bool has_prev) {
    if (SIGNED) {
      const i64 s = (i64)sign_extend(key, W);
      smin = s < smin ? s : smin;
      smax = s > smax ? s : smax;
    } else {
      umin = key < umin ? key : umin;
      umax = key > umax ? key : umax;
    }
    pairs += (has_prev && key == prev) ? 1 : 0;
    nonzero += key != 0 ? 1 : 0;
  };

  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const size_t vbase = t * TILE_VECS + (size_t)warp * (ROWS * 32);
    uint4 v[ROWS];
    bool valid[ROWS];
#pragma unroll
    for (int u = 0; u < ROWS; u++) {
      const size_t vi = vbase + u * 32 + lane;
      valid[u] = vi < nvec;
      v[u] = valid[u] ? ld_stream16(a + vi) : make_uint4(0, 0, 0, 0);
    }
    u64 carry = 0;
    bool carry_ok = false;
    if (vbase > 0 && vbase < nvec) {
      carry = ld_elem<W>(a, vbase * V - 1);
      carry_ok = true;
    }
#pragma unroll
    for (int u = 0; u < ROWS; u++) {
      const u64 last = vec_get<W>(v[u], V - 1);
      u64 prev = __shfl_up_sync(FRB_FULL_MASK, last, 1);
      const u64 row_last = __shfl_sync(FRB_FULL_MASK, last, 31);
      bool has_prev = true;
      if (lane == 0) { prev = carry; has_prev = carry_ok; }
      carry = row_last;
      carry_ok = true;
      if (valid[u]) {
#pragma unroll
        for (int j = 0; j < V; j++) {
          const u64 key = vec_get<W>(v[u], j);
          consume(key, prev, has_prev);
          prev = key;
          has_prev = true;
        }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(a, i);
      consume(key, i > 0 ? ld_elem<W>(a, i - 1) : 0, i > 0);
    }
  }
  // warp reduce, then one atomic set per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (SIGNED) {
      const i64 a1 = __shfl_xor_sync(FRB_FULL_MASK, smin, o), b1 = __shfl_xor_sync(FRB_FULL_MASK, smax, o);
      smin = a1 < smin ? a1 : smin;
      smax = b1 > smax ? b1 : smax;
    } else {
      const u64 a1 = __shfl_xor_sync(FRB_FULL_MASK, umin, o), b1 = __shfl_xor_sync(FRB_FULL_MASK, umax, o);
      umin = a1 < umin ? a1 : umin;
      umax = b1 > umax ? b1 : umax;
    }
    pairs += __shfl_xor_sync(FRB_FULL_MASK, pairs, o);
    nonzero += __shfl_xor_sync(FRB_FULL_MASK, nonzero, o);
  }
  if (lane == 0) {
    if (SIGNED) {
      atomicMin((i64*)&out[0], smin);
      atomicMax((i64*)&out[1], smax);
    } else {
      atomicMin(&out[0], umin);
      atomicMax(&out[1], umax);
    }
    if (pairs) atomicAdd(&out[2], pairs);
    if (nonzero) atomicAdd(&out[3], nonzero);
  }
}

__global__ void k_minmax_init(u64* out, int is_signed) {
  if (threadIdx.x == 0) {
    out[0] = is_signed ? 0x7FFFFFFFFFFFFFFFull : FRB_EMPTY;
    out[1] = is_signed ? 0x8000000000000000ull : 0ull;
    out[2] = 0;
    out[3] = 0;
  }
}

// ------------------------------------------------------------------------------------------ K1
// Every run head (a[i] != a[i-1], exact across lanes, rows and tiles) offers its flat index to
// the table with atomicMin.  A head first reads its slot (16 B, L2-coherent): the common case
// "label already present with a smaller index" ends there without an atomic.
template <int W>
__global__ void __launch_bounds__(256)
k_renumber_build(const uint4* __restrict__ a, size_t n, u64 flat_offset, int skip_zero, Slot* __restrict__ slots,
                 u64 mask, int direct, u64* __restrict__ meta) {
  constexpr int V = 16 / W;
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t tiles = (nvec + TILE_VECS - 1) / TILE_VECS;
  unsigned int inserted = 0;

  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    // a full table aborts the pass (the host grows the table and retries)
    u64 ovf = 0;
    if (lane == 0) ovf = *((volatile u64*)&meta[FRB_META_OVERFLOW]);
    if (__shfl_sync(FRB_FULL_MASK, ovf, 0)) break;

    const size_t vbase = t * TILE_VECS + (size_t)warp * (ROWS * 32);
    uint4 v[ROWS];
    bool valid[ROWS];
#pragma unroll
    for (int u = 0; u < ROWS; u++) {
      const size_t vi = vbase + u * 32 + lane;
      valid[u] = vi < nvec;
      v[u] = valid[u] ? ld_stream16(a + vi) : make_uint4(0, 0, 0, 0);
    }
    u64 carry = 0;
    bool carry_ok = false;
    if (vbase > 0 && vbase < nvec) {
      carry = ld_elem<W>(a, vbase * V - 1);
      carry_ok = true;
    }
#pragma unroll
    for (int u = 0; u < ROWS; u++) {
      const u64 last = vec_get<W>(v[u], V - 1);
      u64 prev = __shfl_up_sync(FRB_FULL_MASK, last, 1);
      const u64 row_last = __shfl_sync(FRB_FULL_MASK, last, 31);
      bool has_prev = true;
      if (lane == 0) { prev = carry; has_prev = carry_ok; }
      carry = row_last;
      carry_ok = true;
      if (valid[u]) {
        const size_t i0 = (vbase + u * 32 + lane) * V;
#pragma unroll
        for (int j = 0; j < V; j++) {
          const u64 key = vec_get<W>(v[u], j);
          const bool head = !has_prev || key != prev;
          prev = key;
          has_prev = true;
          if (head && !(skip_zero && key == 0))
            table_update<FRB_OP_MIN>(slots, mask, direct, meta, key, (u64)(i0 + j) + flat_offset, &inserted);
        }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(a, i);
      const bool head = i == 0 || key != ld_elem<W>(a, i - 1);
      if (head && !(skip_zero && key == 0))
        table_update<FRB_OP_MIN>(slots, mask, direct, meta, key, (u64)i + flat_offset, &inserted);
    }
  }
  block_add_to(&meta[FRB_META_COUNT], inserted);
}

// ------------------------------------------------------------------------------------------ K2
// after the sort: rank r -> id; write the id into the table and emit the (label, id) mapping
template <int W>
__global__ void __launch_bounds__(256)
k_assign_ids(const u64* __restrict__ payload, size_t L, Slot* __restrict__ slots, u64 cap, u64* __restrict__ meta,
             u64 start_bits, void* __restrict__ keys_out, void* __restrict__ ids_out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < L; r += stride) {
    const u64 s = payload[r];
    const u64 id = trunc_width(start_bits + (u64)r, W);  // counter arithmetic wraps in the array dtype
    u64 key;
    if (s == cap) {
      key = FRB_EMPTY;
      meta[FRB_META_SIDE_VAL] = id;
    } else {
      key = slots[s].key;
      slots[s].val = id;
    }
    st_elem<W>(keys_out, r, key);
    st_elem<W>(ids_out, r, id);
  }
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_minmax_pairs(const void* a, size_t n, int dtype, uint64_t* out, void* stream) {
  if (!dtype_ok(dtype) || !out || (n && (!a || !aligned16(a)))) { set_error("frb_minmax_pairs: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const bool sg = dtype_signed(dtype);
  k_minmax_init<<<1, 32, 0, st>>>((u64*)out, sg ? 1 : 0);
  int rc = check_launch("k_minmax_init");
  if (rc || n == 0) return rc;
  Grid g = grid_for(n / (16 / w) + 1, TILE_VECS, 8);
  const uint4* p = (const uint4*)a;
  u64* o = (u64*)out;
#define FRB_LAUNCH(WW)                                                                     \
  if (sg) k_minmax_pairs<WW, true><<<g.blocks, g.threads, 0, st>>>(p, n, o);               \
  else k_minmax_pairs<WW, false><<<g.blocks, g.threads, 0, st>>>(p, n, o)
  switch (w) {
    case 1: FRB_LAUNCH(1); break;
    case 2: FRB_LAUNCH(2); break;
    case 4: FRB_LAUNCH(4); break;
    default: FRB_LAUNCH(8); break;
  }
#undef FRB_LAUNCH
  return check_launch("k_minmax_pairs");
}

int frb_renumber_build(const void* a, size_t n, int dtype, uint64_t flat_offset, int skip_zero, void* slots,
                       uint64_t cap, uint64_t* meta, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || (n && (!a || !aligned16(a)))) {
    set_error("frb_renumber_build: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  const int direct = w <= 2 ? 1 : 0;
  if (direct && cap != frb_table_direct_cap(dtype)) { set_error("frb_renumber_build: direct table needs cap == 2^bits"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = grid_for(n / (16 / w) + 1, TILE_VECS, 8);
  const uint4* p = (const uint4*)a;
  Slot* s = (Slot*)slots;
  u64* m = (u64*)meta;
  switch (w) {
    case 1: k_renumber_build<1><<<g.blocks, g.threads, 0, st>>>(p, n, flat_offset, skip_zero, s, cap - 1, direct, m); break;
    case 2: k_renumber_build<2><<<g.blocks, g.threads, 0, st>>>(p, n, flat_offset, skip_zero, s, cap - 1, direct, m); break;
    case 4: k_renumber_build<4><<<g.blocks, g.threads, 0, st>>>(p, n, flat_offset, skip_zero, s, cap - 1, direct, m); break;
    default: k_renumber_build<8><<<g.blocks, g.threads, 0, st>>>(p, n, flat_offset, skip_zero, s, cap - 1, direct, m); break;
  }
  return check_launch("k_renumber_build");
}

// workspace: 4 lists of `capacity` uint64 (sort key, payload, 2 ping-pong) + the radix histograms
size_t frb_renumber_rank_ws_bytes(size_t capacity) {
  const size_t c = capacity < 1 ? 1 : capacity;
  return 4 * align_up(c * sizeof(u64), 256) + radix_hist_bytes(c);
}

int frb_renumber_rank(void* slots, uint64_t cap, uint64_t* meta, int dtype, int64_t start, int index_bits,
                      void* keys_out, void* ids_out, size_t capacity, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || !ws || ws_bytes < frb_renumber_rank_ws_bytes(capacity) ||
      (capacity && (!keys_out || !ids_out)) || index_bits < 1 || index_bits > 64) {
    set_error("frb_renumber_rank: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t c = capacity < 1 ? 1 : capacity;
  const size_t lb = align_up(c * sizeof(u64), 256);
  char* base = (char*)ws;
  u64* sort_keys = (u64*)(base);
  u64* payload = (u64*)(base + lb);
  u64* keys_alt = (u64*)(base + 2 * lb);
  u64* pay_alt = (u64*)(base + 3 * lb);
  void* hist = (void*)(base + 4 * lb);
  int rc = launch_table_export((const Slot*)slots, cap, (u64*)meta, sort_keys, payload, capacity, 1, st);
  if (rc) return rc;
  // The caller sized `capacity` from meta[COUNT] after K1, so exactly `capacity` entries exist.
  if (capacity == 0) return FRB_OK;
  rc = radix_sort_u64(sort_keys, payload, keys_alt, pay_alt, capacity, 0, index_bits, hist, st);
  if (rc) return rc;
  const int w = dtype_width(dtype);
  const u64 start_bits = trunc_width((u64)start, w);
  Grid g = grid_for(capacity, 256 * 4, 8);
  Slot* s = (Slot*)slots;
  u64* m = (u64*)meta;
  switch (w) {
    case 1: k_assign_ids<1><<<g.blocks, g.threads, 0, st>>>(payload, capacity, s, cap, m, start_bits, keys_out, ids_out); break;
    case 2: k_assign_ids<2><<<g.blocks, g.threads, 0, st>>>(payload, capacity, s, cap, m, start_bits, keys_out, ids_out); break;
    case 4: k_assign_ids<4><<<g.blocks, g.threads, 0, st>>>(payload, capacity, s, cap, m, start_bits, keys_out, ids_out); break;
    default: k_assign_ids<8><<<g.blocks, g.threads, 0, st>>>(payload, capacity, s, cap, m, start_bits, keys_out, ids_out); break;
  }
  return check_launch("k_assign_ids");
}

}  // extern "C"
