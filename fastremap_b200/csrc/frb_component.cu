// frb_component.cu -- K7 (component_map) and the synthetic-volume generator.
//
// K7 replaces _component_map (fastremap.pyx:565-587).  The reference stores
// remap[cc[i]] = parent[i] at every run start of cc in memory order, so for each label the LAST
// run start wins.  Here every run start offers (i + 1) to the label's slot with atomicMax; the
// gather pass then reads parent[] at the winning index.  Warps walk their tiles from the back, so
// the first offer a warp makes for a label is the largest it will ever make, and the register /
// shared-memory / home-slot filters of frb_build.cuh drop the rest.
#include "frb_internal.cuh"
#include "frb_build.cuh"
#include "frb_scan.cuh"

namespace frb {

// gather: scan the table (frb_scan.cuh) and emit (label, parent[winning run start]) for every label
struct GatherSel {
  const void* parent;
  void* keys_out;
  void* parents_out;
  u64 flat_offset;
  int wc, wp;
  __device__ __forceinline__ bool pick(u64 key, u64 val) const { return key != FRB_EMPTY && val != 0; }  // 0: withdrawn (frb_component_unhead)
  __device__ __forceinline__ u64 fetch(u64, u64 val) const { return ld_elem_rt(parent, wp, (size_t)(val - 1 - flat_offset)); }
  __device__ __forceinline__ void emit(u64 pos, u64 key, u64, u64, u64 fetched) const {
    st_elem_rt(keys_out, wc, pos, key);
    st_elem_rt(parents_out, wp, pos, fetched);
  }
};
template <int W, int IL, int DU, bool VAR_TILE>
__global__ void __launch_bounds__(STREAM_THREADS, 3)  // 72 registers: three CTAs per SM (the pass is latency-bound at 10^7 labels)
k_component_build(const uint4* __restrict__ cc, size_t n, u64 flat_offset, Slot* __restrict__ slots, u64 mask, int direct,
                  u64* __restrict__ meta, int run, int tile_vecs, PlaneWalk pw, unsigned int* slot_list, u64 region_cap) {
  build_pass<W, FRB_OP_MAX, true, true, IL, DU, VAR_TILE>(cc, n, flat_offset + 1, 0, slots, mask, direct, meta, run, tile_vecs, pw, slot_list, region_cap);
}
// the same pass over a balanced plane walk (whole 16 KiB tiles, rows of 4 KiB multiples: IL = 8)
template <int W, int DU>
__global__ void __launch_bounds__(STREAM_THREADS, 3)
k_component_build_z(const uint4* __restrict__ cc, size_t n, u64 flat_offset, Slot* __restrict__ slots, u64 mask, int direct,
                    u64* __restrict__ meta, PlaneWalkZ pwz, unsigned int* slot_list, u64 region_cap) {
  build_pass<W, FRB_OP_MAX, true, true, 8, DU, false, true>(cc, n, flat_offset + 1, 0, slots, mask, direct, meta, 1, STREAM_TILE_VECS, PlaneWalk{0, 0, 0},
                                                             slot_list, region_cap, pwz);
}

// gather through the list of filled slots the build pass kept (layout: frb_build.cuh; complete -- the host checked
// meta[AUX1]): no scan of a mostly empty table, no compaction.  A CTA takes one region at a time; its first output
// position is the sum of the counts of the regions in front (a few thousand words, L2-resident); four independent
// (slot, parent) read chains per thread.  A region lists its slots in the order its warp walked the volume, so the
// reads of parent[] of one region stay inside a few 2 MB pages.
__global__ void __launch_bounds__(256)
k_component_gather_list(const unsigned int* __restrict__ list, const Slot* slots, GatherSel sel) {
  constexpr int B = 4;
  __shared__ unsigned long long s_base;
  const unsigned int R = list[0], rcap = list[1];
  const unsigned int* counts = list + 4;
  const unsigned int* data = list + 4 + R;
  for (unsigned int r = blockIdx.x; r < R; r += gridDim.x) {
    const unsigned int cnt = counts[r];
    if (cnt == 0) continue;  // (the same in every thread)
    unsigned long long part = 0;
    for (unsigned int q = threadIdx.x; q < r; q += 256) part += counts[q];
    __syncthreads();
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    part = __reduce_add_sync(FRB_FULL_MASK, (unsigned int)part);  // a region holds < 2^32 / 32 entries: the warp sum fits
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_base, part);
    __syncthreads();
    const u64 base = s_base;
    const unsigned int* mine = data + (u64)r * rcap;
    for (unsigned int i0 = threadIdx.x; i0 < cnt; i0 += 256 * B) {
      Slot sl[B];
      u64 got[B];
      bool live[B];
#pragma unroll
      for (int b = 0; b < B; b++) {
        const unsigned int i = i0 + b * 256;
        live[b] = i < cnt;
        sl[b].key = 0;
        sl[b].val = 0;
        if (live[b]) sl[b] = ld_slot_cg_batch(slots + mine[i]);
      }
#pragma unroll
      for (int b = 0; b < B; b++) got[b] = (live[b] && sl[b].val != 0) ? sel.fetch(sl[b].key, sl[b].val) : 0;
#pragma unroll
      for (int b = 0; b < B; b++)
        if (live[b]) sel.emit(base + i0 + b * 256, sl[b].key, sl[b].val, 0, got[b]);
    }
  }
}

// Sharded volumes: a slab's first voxel was offered as a run start; if it carries the label the previous slab ended
// with it is not one (fastremap.pyx:581: only cc[i] != cc[i-1] starts a run).  Withdraw that offer -- unless a later
// run start of the label in this slab has overtaken it anyway.
__global__ void k_component_unhead(const void* cc, int w, u64 prev_bits, u64 flat_offset, Slot* slots, u64 mask, int direct, u64* meta) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const u64 key = ld_elem_rt(cc, w, 0);
  if (key != prev_bits) return;
  if (key == FRB_EMPTY && !direct) { atom_cas_u64(&meta[FRB_META_SIDE_VAL], flat_offset + 1, 0ull); return; }
  u64 h = direct ? key : home_slot(key, mask);
  for (int probe = 0; probe < FRB_MAX_PROBE; probe++, h = (h + 1) & mask) {
    const Slot cur = ld_slot_cg(slots + h);
    if (cur.key == key) { atom_cas_u64(&slots[h].val, flat_offset + 1, 0ull); return; }
    if (cur.key == FRB_EMPTY || direct) return;
  }
}

__global__ void k_store_u64(u64* p, u64 v) { *p = v; }
__global__ void k_gather_side(u64* meta, GatherSel sel, u64 capacity) {
  if (threadIdx.x == 0 && meta[FRB_META_SIDE_PRESENT] && meta[FRB_META_SIDE_VAL] != 0) {  // the out-of-band all-ones label
    const u64 pos = atomicAdd(&meta[FRB_META_LIST_COUNT], 1ull);
    if (pos < capacity) sel.emit(pos, FRB_EMPTY, meta[FRB_META_SIDE_VAL], 0, sel.fetch(FRB_EMPTY, meta[FRB_META_SIDE_VAL]));
  }
}

// ------------------------------------------------------------------------------ synthetic volume
__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
  u64 z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 mix(u64 seed, u64 idx) { return splitmix64(seed * 0xD1342543DE82EF95ull + idx); }

struct SynthArgs {
  u64 nz, ny, nx, z0, nz_total, gz, gy, gx, seed, zero_thresh, noise_thresh, idmask, cell_div;
};

template <int W>
__global__ void __launch_bounds__(256) k_synth(void* __restrict__ out, SynthArgs S) {
  constexpr int V = 16 / W;
  const u64 n = S.nz * S.ny * S.nx;
  const u64 ncells = S.gz * S.gy * S.gx;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = (n + V - 1) / V;
  for (size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec; vi += stride) {
    u64 r[V];
#pragma unroll
    for (int j = 0; j < V; j++) {
      const u64 i = (u64)vi * V + j;
      u64 label = 0;
      if (i < n) {
        const u64 x = i % S.nx;
        const u64 yz = i / S.nx;
        const u64 y = yz % S.ny;
        const u64 zg = yz / S.ny + S.z0;
        u64 c = (((zg * S.gz) / S.nz_total) * S.gy + (y * S.gy) / S.ny) * S.gx + (x * S.gx) / S.nx;
        if (S.noise_thresh) {
          const u64 hn = mix(S.seed + 2, (zg * S.ny + y) * S.nx + x);
          if ((hn >> 40) < S.noise_thresh) c = (c + 1 + (hn & 7)) % ncells;
        }
        c /= S.cell_div;
        label = mix(S.seed, c) & S.idmask;
        if (label == 0) label = 1;
        if ((mix(S.seed + 1, c) >> 40) < S.zero_thresh) label = 0;
      }
      r[j] = label;
    }
    const size_t i0 = vi * V;
    if (i0 + V <= n) {
      store_results<V, W>(out, i0, r);
    } else {
#pragma unroll
      for (int j = 0; j < V; j++)
        if (i0 + j < n) st_elem<W>(out, i0 + j, r[j]);
    }
  }
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_component_build(const void* cc, size_t n, int cc_dtype, uint64_t flat_offset, void* slots, uint64_t cap,
                        uint64_t* meta, uint64_t row_bytes, uint64_t plane_bytes, uint32_t* slot_list, uint64_t slot_list_capacity,
                        void* stream) {
  if (!dtype_ok(cc_dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || (n && (!cc || !aligned16(cc)))) { set_error("frb_component_build: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(cc_dtype);
  const int direct = w <= 2 ? 1 : 0;
  if (direct && cap != frb_table_direct_cap(cc_dtype)) { set_error("frb_component_build: direct table needs cap == 2^bits"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)cc;
  Slot* s = (Slot*)slots;
  u64* m = (u64*)meta;
  const RowGeom geo = row_geom((size_t)row_bytes);
  const int tv = geo.tv;
  if (slot_list) {  // header + per-region counts (at most 8 warps x 4 CTAs x SMs regions): regions whose warp never runs stay empty
    const size_t head = (size_t)(4 + 8 * 4 * sm_count() + 1) * sizeof(uint32_t);
    if ((size_t)slot_list_capacity * sizeof(uint32_t) < 2 * head) { set_error("frb_component_build: slot list too small"); return FRB_ERR_BAD_ARG; }
    cudaMemsetAsync(slot_list, 0, head, st);
  }
#define FRB_K7(WW, II, DD, VV) k_component_build<WW, II, DD, VV><<<stream_grid(k_component_build<WW, II, DD, VV>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>( \
    p, n, flat_offset, s, cap - 1, direct, m, fair_run(stream_grid(k_component_build<WW, II, DD, VV>, tiles), tiles), tv, \
    (VV) ? PlaneWalk{0, 0, 0} : plane_walk_for(n * (size_t)WW, (n % (16 / WW)) ? 0 : (size_t)plane_bytes, stream_grid(k_component_build<WW, II, DD, VV>, tiles)), \
    (unsigned int*)slot_list, slot_list ? slot_list_region_cap((u64)slot_list_capacity, (u64)stream_grid(k_component_build<WW, II, DD, VV>, tiles) * STREAM_WARPS) : 0ull)
#define FRB_K7Z(WW, DD) { \
    const int gz = stream_grid(k_component_build_z<WW, DD>, tiles); \
    const PlaneWalkZ pz = plane_walk_z_for(n * (size_t)WW, (n % (16 / WW)) ? 0 : (size_t)plane_bytes, gz, (size_t)plane_band()); \
    if (pz.band) { \
      k_component_build_z<WW, DD><<<gz, STREAM_THREADS, STREAM_SMEM_BYTES, st>>>(p, n, flat_offset, s, cap - 1, direct, m, pz, (unsigned int*)slot_list, \
          slot_list ? slot_list_region_cap((u64)slot_list_capacity, (u64)gz * STREAM_WARPS) : 0ull); \
      return check_launch("k_component_build_z"); \
    } }
#define FRB_K7_GEO(WW)                                                        \
  if (geo.tv == STREAM_TILE_VECS && geo.il == 8) {                            \
    if (geo.du == 1) FRB_K7Z(WW, 1) else if (geo.du == 2) FRB_K7Z(WW, 2) else if (geo.du == 3) FRB_K7Z(WW, 3) else FRB_K7Z(WW, 4) \
  }                                                                           \
  if (geo.tv != STREAM_TILE_VECS) { FRB_K7(WW, 8, 4, true); }                 \
  else if (geo.il == 4) { FRB_K7(WW, 4, 1, false); }                          \
  else if (geo.il == 2) { FRB_K7(WW, 2, 1, false); }                          \
  else if (geo.il == 1) { FRB_K7(WW, 1, 1, false); }                          \
  else if (geo.du == 1) { FRB_K7(WW, 8, 1, false); }                          \
  else if (geo.du == 2) { FRB_K7(WW, 8, 2, false); }                          \
  else if (geo.du == 3) { FRB_K7(WW, 8, 3, false); }                          \
  else { FRB_K7(WW, 8, 4, false); }
  switch (w) {
    case 1: FRB_K7_GEO(1) break;
    case 2: FRB_K7_GEO(2) break;
    case 4: FRB_K7_GEO(4) break;
    default: FRB_K7_GEO(8) break;
  }
#undef FRB_K7_GEO
#undef FRB_K7Z
#undef FRB_K7
  return check_launch("k_component_build");
}

int frb_component_unhead(const void* cc, int cc_dtype, uint64_t prev_bits, uint64_t flat_offset, void* slots, uint64_t cap, uint64_t* meta,
                         void* stream) {
  if (!dtype_ok(cc_dtype) || !cc || !slots || !meta || (cap & (cap - 1)) != 0) { set_error("frb_component_unhead: bad arguments"); return FRB_ERR_BAD_ARG; }
  const int w = dtype_width(cc_dtype);
  k_component_unhead<<<1, 32, 0, (cudaStream_t)stream>>>(cc, w, trunc_width(prev_bits, w), flat_offset, (Slot*)slots, cap - 1, w <= 2 ? 1 : 0, (u64*)meta);
  return check_launch("k_component_unhead");
}

int frb_component_gather(const void* slots, uint64_t cap, uint64_t* meta, int cc_dtype, const void* parent, int p_dtype,
                         uint64_t flat_offset, void* keys_out, void* parents_out, size_t capacity, const uint32_t* slot_list,
                         uint64_t slot_list_count, void* stream) {
  if (!dtype_ok(cc_dtype) || !dtype_ok(p_dtype) || !slots || !meta || !parent || (capacity && (!keys_out || !parents_out)) ||
      (slot_list && slot_list_count > capacity)) { set_error("frb_component_gather: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(&((u64*)meta)[FRB_META_LIST_COUNT], 0, sizeof(u64), st);
  GatherSel sel;
  sel.parent = parent;
  sel.keys_out = keys_out;
  sel.parents_out = parents_out;
  sel.flat_offset = flat_offset;
  sel.wc = dtype_width(cc_dtype);
  sel.wp = dtype_width(p_dtype);
  u64* m = (u64*)meta;
  if (slot_list) {  // the complete list of filled slots: entry i of the output is slot slot_list[i]; then the all-ones label
    if (slot_list_count) {
      k_component_gather_list<<<sm_count() * 8, 256, 0, st>>>((const unsigned int*)slot_list, (const Slot*)slots, sel);
      int rc = check_launch("k_component_gather_list");
      if (rc) return rc;
    }
    k_store_u64<<<1, 1, 0, st>>>(&m[FRB_META_LIST_COUNT], (u64)slot_list_count);
    k_gather_side<<<1, 32, 0, st>>>(m, sel, (u64)capacity);
    return check_launch("k_gather_side");
  }
  int rc = launch_table_scan(slots, cap, &m[FRB_META_LIST_COUNT], (u64)capacity, sel, "k_table_scan<gather>", st);
  if (rc) return rc;
  k_gather_side<<<1, 32, 0, st>>>(m, sel, (u64)capacity);
  return check_launch("k_gather_side");
}

int frb_synth_volume(void* out, int dtype, uint64_t nz, uint64_t ny, uint64_t nx, uint64_t z0, uint64_t nz_total,
                     uint64_t gz, uint64_t gy, uint64_t gx, uint64_t seed, double zero_frac, double noise,
                     uint64_t cell_div, void* stream) {
  if (!dtype_ok(dtype) || !out || !aligned16(out) || !nz || !ny || !nx || !nz_total || !gz || !gy || !gx || !cell_div) { set_error("frb_synth_volume: bad arguments"); return FRB_ERR_BAD_ARG; }
  const int w = dtype_width(dtype);
  SynthArgs S;
  S.nz = nz; S.ny = ny; S.nx = nx; S.z0 = z0; S.nz_total = nz_total; S.gz = gz; S.gy = gy; S.gx = gx; S.seed = seed;
  S.zero_thresh = (u64)(zero_frac * 16777216.0);
  S.noise_thresh = (u64)(noise * 16777216.0);
  const int idbits = dtype_signed(dtype) ? 8 * w - 1 : (8 * w > 63 ? 63 : 8 * w);
  S.idmask = (1ull << idbits) - 1ull;
  S.cell_div = cell_div;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)(nz * ny * nx);
  Grid g = grid_for(n / (16 / w) + 1, 256 * 2, 16);
  switch (w) {
    case 1: k_synth<1><<<g.blocks, g.threads, 0, st>>>(out, S); break;
    case 2: k_synth<2><<<g.blocks, g.threads, 0, st>>>(out, S); break;
    case 4: k_synth<4><<<g.blocks, g.threads, 0, st>>>(out, S); break;
    default: k_synth<8><<<g.blocks, g.threads, 0, st>>>(out, S); break;
  }
  return check_launch("k_synth");
}

}  // extern "C"
