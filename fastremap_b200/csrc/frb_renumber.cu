// frb_renumber.cu -- K6 (fused min / max / pixel-pairs scan), K1 (first-index table build) and
// K2 (rank labels by first index -> ids) of the renumber path.
//
// Reference loops replaced: _minmax fastremap.pyx:77-93, _pixel_pairs :840-853, and the
// find/insert half of _renumber :236-251.  The reference numbers labels in encounter order with
// a serial counter; here id(label) = start + rank(first flat index of label), which is the same
// numbering computed without the serial dependence: K1 records min(first index) per label with
// 64-bit atomicMin, K2 (frb_rank.cu) sorts the labels by that index.  First indices are stored with
// bit 63 set (FRB_RANK_TAG) so that they compare above every already assigned rank: ranking is
// incremental and a volume may be built chunk by chunk.
#include <stdlib.h>

#include "frb_internal.cuh"
#include "frb_build.cuh"

namespace frb {

// ------------------------------------------------------------------------------------------ K6
template <int W, bool SIGNED>
__global__ void __launch_bounds__(STREAM_THREADS)
k_minmax_pairs(const uint4* __restrict__ a, size_t n, u64* __restrict__ out, int run) {
  constexpr int V = 16 / W;
  typedef long long i64;
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31;
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
  auto row_fn = [&](int, const uint4& cur, bool valid, size_t, u64 prev, bool has_prev) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < V; j++) {
      const u64 key = vec_get<W>(cur, j);
      consume(key, prev, has_prev);
      prev = key;
      has_prev = true;
    }
  };
  const TileMap map(nvec, (size_t)run);
  stream_rows<W, false, true, 0>(a, nvec, map, row_fn, [](const uint4*, size_t) {});

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
// First-index pass: the table receives, for every label, the minimum flat index at which it occurs
// (tagged: see FRB_RANK_TAG).  The filtered build pass is shared with K7: frb_build.cuh.
template <int W, int IL, int DU, bool VAR_TILE>
__global__ void __launch_bounds__(STREAM_THREADS)
k_renumber_build(const uint4* __restrict__ a, size_t n, u64 flat_offset, int skip_zero, Slot* __restrict__ slots,
                 u64 mask, int direct, u64* __restrict__ meta, int run, int tile_vecs, PlaneWalk pw) {
  build_pass<W, FRB_OP_MIN, false, false, IL, DU, VAR_TILE>(a, n, FRB_RANK_TAG | flat_offset, skip_zero, slots, mask, direct, meta, run, tile_vecs, pw);
}

// ------------------------------------------------------------------------------------------ probes
// Calibration kernels: the same TMA-staged streaming loop with (almost) no arithmetic.
// mode 0: read only (xor checksum); mode 1: read 16 B, write 16 B; mode 2: read 16 B, write 16 B + 8 B
template <bool PREV>
__global__ void __launch_bounds__(STREAM_THREADS)
k_stream_probe(const uint4* __restrict__ a, size_t nvec, void* wide, void* narrow, int mode, u64* out, int run) {
  uint32_t acc = 0;
  auto row_fn = [&](int, const uint4& cur, bool valid, size_t i0, u64, bool) {
    if (!valid) return;
    acc ^= cur.x ^ cur.y ^ cur.z ^ cur.w;
    if (mode >= 1) st_stream16((char*)wide + i0 * 8, cur);
    if (mode >= 2) st_stream8((char*)narrow + i0 * 4, make_uint2(cur.x, cur.z));
  };
  const TileMap map(nvec, (size_t)run);
  stream_rows<8, false, PREV, 0>(a, nvec, map, row_fn, [](const uint4*, size_t) {});
  acc = __reduce_xor_sync(FRB_FULL_MASK, acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicXor(out, (u64)acc);
}

// One pass through L2?  The feasibility probe of a single-pass renumber: ONE persistent, cooperatively
// launched kernel walks the array in windows; per window: pass A reads it (the traffic of K1), grid barrier, pass B reads it
// again (from L2, if the window is still there) and writes 8 B + 4 B per voxel (the traffic of K3), grid barrier.  No table,
// no ranking: what this measures is the ceiling of the idea -- HBM traffic 20 instead of 28 B/voxel if pass B hits in L2.
__global__ void __launch_bounds__(STREAM_THREADS)
k_window_probe(const uint4* __restrict__ a, size_t nvec, size_t win_vecs, void* wide, void* narrow, u64* out, unsigned int* barrier, int run) {
  extern __shared__ __align__(128) unsigned char frb_dyn_smem[];
  uint32_t acc = 0;
  unsigned int generation = 0;
  auto sync_grid = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) {
      generation++;
      __threadfence();
      atomicAdd(barrier, 1u);
      const unsigned int target = generation * gridDim.x;
      while (*((volatile unsigned int*)barrier) < target) __nanosleep(20);
      __threadfence();
      // the ring's mbarriers are re-initialised by the next stream_rows call: retire the old objects first
      u64* bars = reinterpret_cast<u64*>(frb_dyn_smem + (size_t)STREAM_STAGES * STREAM_STAGE_VECS * 16);
      for (int s = 0; s < 2 * STREAM_STAGES; s++) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[s])) : "memory");
    }
    __syncthreads();
  };
  for (size_t w0 = 0; w0 < nvec; w0 += win_vecs) {
    const size_t nv = nvec - w0 < win_vecs ? nvec - w0 : win_vecs;
    const TileMap map(nv, (size_t)run);
    stream_rows<8, false, false, 0>(a + w0, nv, map, [&](int, const uint4& cur, bool valid, size_t, u64, bool) {
      if (valid) acc ^= cur.x ^ cur.y ^ cur.z ^ cur.w;
    }, [](const uint4*, size_t) {});
    sync_grid();
    char* wd = (char*)wide + w0 * 16;
    char* nr = (char*)narrow + w0 * 8;
    stream_rows<8, false, false, 0>(a + w0, nv, map, [&](int, const uint4& cur, bool valid, size_t i0, u64, bool) {
      if (!valid) return;
      acc ^= cur.y;
      st_stream16(wd + i0 * 8, cur);
      st_stream8(nr + i0 * 4, make_uint2(cur.x, cur.z));
    }, [](const uint4*, size_t) {});
    sync_grid();
  }
  acc = __reduce_xor_sync(FRB_FULL_MASK, acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicXor(out, (u64)acc);
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
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)a;
  u64* o = (u64*)out;
#define FRB_LAUNCH(WW)                                                                                              \
  if (sg) k_minmax_pairs<WW, true><<<stream_grid(k_minmax_pairs<WW, true>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>(p, n, o, fair_run(stream_grid(k_minmax_pairs<WW, true>, tiles), tiles));    \
  else k_minmax_pairs<WW, false><<<stream_grid(k_minmax_pairs<WW, false>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>(p, n, o, fair_run(stream_grid(k_minmax_pairs<WW, false>, tiles), tiles))
  switch (w) {
    case 1: FRB_LAUNCH(1); break;
    case 2: FRB_LAUNCH(2); break;
    case 4: FRB_LAUNCH(4); break;
    default: FRB_LAUNCH(8); break;
  }
#undef FRB_LAUNCH
  return check_launch("k_minmax_pairs");
}

int frb_stream_probe(const void* src, size_t nbytes, void* dst_wide, void* dst_narrow, int mode, uint64_t* out, void* stream) {
  if (!src || !out || !aligned16(src) || ((mode & 3) >= 1 && !dst_wide) || ((mode & 3) >= 2 && !dst_narrow)) { set_error("frb_stream_probe: bad arguments"); return FRB_ERR_BAD_ARG; }
  const size_t nvec = nbytes / 16;
  const size_t tiles = (nvec + STREAM_TILE_VECS - 1) / STREAM_TILE_VECS + 1;
  if (mode & 8) {  // windowed two-pass probe in one cooperative kernel; out[1] is the grid barrier word (zeroed here)
    if (!dst_wide || !dst_narrow) { set_error("frb_stream_probe: the window probe needs both destinations"); return FRB_ERR_BAD_ARG; }
    const char* e = getenv("FRB_PROBE_WINDOW_MIB");
    size_t win_vecs = (size_t)(e ? atoi(e) : 64) * (1u << 16);
    if (win_vecs < (size_t)STREAM_TILE_VECS) win_vecs = STREAM_TILE_VECS;
    const size_t wtiles = (win_vecs + STREAM_TILE_VECS - 1) / STREAM_TILE_VECS;
    const int grid = stream_grid(k_window_probe, wtiles);
    int run = fair_run(grid, wtiles);
    const uint4* a = (const uint4*)src;
    u64* o = (u64*)out;
    unsigned int* bar = (unsigned int*)(o + 1);
    cudaMemsetAsync(bar, 0, 8, (cudaStream_t)stream);
    void* params[] = {&a, (void*)&nvec, &win_vecs, &dst_wide, &dst_narrow, &o, &bar, &run};
    cudaError_t err = cudaLaunchCooperativeKernel((const void*)k_window_probe, dim3(grid), dim3(STREAM_THREADS), params, STREAM_SMEM_BYTES, (cudaStream_t)stream);
    if (err != cudaSuccess) { set_error(cudaGetErrorString(err)); return FRB_ERR_CUDA; }
    return check_launch("k_window_probe");
  }
  if (mode & 4)
    k_stream_probe<true><<<stream_grid(k_stream_probe<true>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, (cudaStream_t)stream>>>(
        (const uint4*)src, nvec, dst_wide, dst_narrow, mode & 3, (u64*)out, fair_run(stream_grid(k_stream_probe<true>, tiles), tiles));
  else
    k_stream_probe<false><<<stream_grid(k_stream_probe<false>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, (cudaStream_t)stream>>>(
        (const uint4*)src, nvec, dst_wide, dst_narrow, mode & 3, (u64*)out, fair_run(stream_grid(k_stream_probe<false>, tiles), tiles));
  return check_launch("k_stream_probe");
}

int frb_renumber_build(const void* a, size_t n, int dtype, uint64_t flat_offset, int skip_zero, void* slots,
                       uint64_t cap, uint64_t* meta, uint64_t row_bytes, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || (n && (!a || !aligned16(a)))) {
    set_error("frb_renumber_build: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  const int direct = w <= 2 ? 1 : 0;
  if (direct && cap != frb_table_direct_cap(dtype)) { set_error("frb_renumber_build: direct table needs cap == 2^bits"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)a;
  Slot* s = (Slot*)slots;
  u64* m = (u64*)meta;
  const RowGeom geo = row_geom((size_t)row_bytes);
  const int tv = geo.tv;
#define FRB_K1(WW, II, DD, VV) k_renumber_build<WW, II, DD, VV><<<stream_grid(k_renumber_build<WW, II, DD, VV>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>( \
    p, n, flat_offset, skip_zero, s, cap - 1, direct, m, fair_run(stream_grid(k_renumber_build<WW, II, DD, VV>, tiles), tiles), tv, PlaneWalk{0, 0, 0})
#define FRB_K1_GEO(WW)                                                        \
  if (geo.tv != STREAM_TILE_VECS) { FRB_K1(WW, 8, 4, true); }                 \
  else if (geo.il == 4) { FRB_K1(WW, 4, 1, false); }                          \
  else if (geo.il == 2) { FRB_K1(WW, 2, 1, false); }                          \
  else if (geo.il == 1) { FRB_K1(WW, 1, 1, false); }                          \
  else if (geo.du == 1) { FRB_K1(WW, 8, 1, false); }                          \
  else if (geo.du == 2) { FRB_K1(WW, 8, 2, false); }                          \
  else if (geo.du == 3) { FRB_K1(WW, 8, 3, false); }                          \
  else { FRB_K1(WW, 8, 4, false); }
  switch (w) {
    case 1: FRB_K1_GEO(1) break;
    case 2: FRB_K1_GEO(2) break;
    case 4: FRB_K1_GEO(4) break;
    default: FRB_K1_GEO(8) break;
  }
#undef FRB_K1_GEO
#undef FRB_K1
  return check_launch("k_renumber_build");
}

}  // extern "C"
