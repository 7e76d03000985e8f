// frb_common.cuh -- shared device helpers for the fastremap_b200 kernels (sm_100a).
//
// Every kernel in this library is an HBM-bound integer streaming kernel.  The common shape (stream_rows below): the
// volume is staged through shared memory by the TMA engine -- a producer warp issues one 1-D bulk copy
// (cp.async.bulk, SASS UBLKCP) per 16 KiB tile into a 3-deep ring guarded by mbarriers -- and 8 consumer warps read
// rows of 32 x 16 B = 512 contiguous bytes with conflict-free LDS.128, each lane unpacking V = 16 / W elements of
// width W.  Grids are persistent (co-resident CTAs x 148 SMs on B200); tiles are dealt to CTAs in runs (TileMap) or,
// for the kernels that aggregate per label, band by band through stacks of planes (PlaneWalk).  Label tables are
// open-addressing arrays of 16-byte slots with even home slots (one 32-byte sector = the first two probes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fastremap_b200.h"

#define FRB_EMPTY 0xFFFFFFFFFFFFFFFFull
#define FRB_FULL_MASK 0xFFFFFFFFu
#define FRB_MAX_PROBE 512

namespace frb {

typedef unsigned long long u64;

struct __align__(16) Slot {
  u64 key;
  u64 val;
};

// ---- launch bookkeeping -----------------------------------------------------------------
extern unsigned long long g_launches;  // defined in frb_runtime.cu
void set_error(const char* msg);
int check_launch(const char* what);
int sm_count();

struct Grid {
  int blocks;
  int threads;
};
// persistent-style grid: `per_sm` CTAs per SM, capped by the amount of work
inline Grid grid_for(size_t work_items, size_t items_per_block, int per_sm, int threads = 256) {
  size_t need = (work_items + items_per_block - 1) / items_per_block;
  size_t cap = (size_t)sm_count() * (size_t)per_sm;
  if (need < 1) need = 1;
  Grid g;
  g.blocks = (int)(need < cap ? need : cap);
  g.threads = threads;
  return g;
}

inline int dtype_width(int dtype) { return 1 << (dtype & 3); }
inline bool dtype_signed(int dtype) { return dtype >= 4; }
inline bool dtype_ok(int dtype) { return dtype >= 0 && dtype <= 7; }
inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// ---- element access -----------------------------------------------------------------------
template <int W> struct UIntOf;
template <> struct UIntOf<1> { typedef uint8_t type; };
template <> struct UIntOf<2> { typedef uint16_t type; };
template <> struct UIntOf<4> { typedef uint32_t type; };
template <> struct UIntOf<8> { typedef uint64_t type; };

template <int W>
__device__ __forceinline__ u64 ld_elem(const void* p, size_t i) {
  return (u64)(((const typename UIntOf<W>::type*)p)[i]);
}
template <int W>
__device__ __forceinline__ void st_elem(void* p, size_t i, u64 v) {
  ((typename UIntOf<W>::type*)p)[i] = (typename UIntOf<W>::type)v;
}
__device__ __forceinline__ u64 ld_elem_rt(const void* p, int w, size_t i) {
  switch (w) {
    case 1: return ld_elem<1>(p, i);
    case 2: return ld_elem<2>(p, i);
    case 4: return ld_elem<4>(p, i);
    default: return ld_elem<8>(p, i);
  }
}
__device__ __forceinline__ void st_elem_rt(void* p, int w, size_t i, u64 v) {
  switch (w) {
    case 1: st_elem<1>(p, i, v); break;
    case 2: st_elem<2>(p, i, v); break;
    case 4: st_elem<4>(p, i, v); break;
    default: st_elem<8>(p, i, v); break;
  }
}

// streaming 16-byte load: read-only path, do not allocate in L1
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// same, for buffers that the kernel also writes (in-place relabel): no .nc
__device__ __forceinline__ uint4 ld_stream16_rw(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// streaming stores (written once, never re-read by this kernel)
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream8(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(void* p, uint16_t v) {
  asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}

// element j (compile-time after unrolling) of a 16-byte vector, zero-extended
template <int W>
__device__ __forceinline__ u64 vec_get(const uint4& v, int j) {
  const uint32_t w32[4] = {v.x, v.y, v.z, v.w};
  if (W == 8) return (u64)w32[2 * j] | ((u64)w32[2 * j + 1] << 32);
  if (W == 4) return (u64)w32[j];
  if (W == 2) return (u64)((w32[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
  return (u64)((w32[j >> 2] >> ((j & 3) * 8)) & 0xFFu);
}

// Pack V results (V = 16 / W_IN elements per lane) at output width WO and store them at
// element index `i0` of `dst` (lane-contiguous, so a warp writes 32 * V * WO contiguous bytes).
template <int V, int WO>
__device__ __forceinline__ void store_results(void* dst, size_t i0, const u64 (&r)[V]) {
  constexpr int BYTES = V * WO;
  char* p = (char*)dst + i0 * (size_t)WO;
  if constexpr (BYTES > 16 && WO == 8) {  // widening store to 8-byte results: V / 2 16-byte pieces, lane-contiguous
#pragma unroll
    for (int j = 0; j < V; j += 2)
      st_stream16(p + 8 * j, make_uint4((uint32_t)r[j], (uint32_t)(r[j] >> 32), (uint32_t)r[j + 1], (uint32_t)(r[j + 1] >> 32)));
  } else if constexpr (BYTES > 16) {  // widening store to 2- or 4-byte results (refit's astype): BYTES / 16 pieces of 16 / WO results
    constexpr int PER = 16 / WO;
#pragma unroll
    for (int q = 0; q < V / PER; q++) {
      uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < PER; j++) {
        const uint32_t x = (uint32_t)r[q * PER + j];
        if constexpr (WO == 4) { w[j] = x; }
        else { w[j >> 1] |= (x & 0xFFFFu) << ((j & 1) * 16); }
      }
      st_stream16(p + 16 * q, make_uint4(w[0], w[1], w[2], w[3]));
    }
  } else if constexpr (BYTES == 16) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < V; j++) {
      if constexpr (WO == 8) { w[2 * j] = (uint32_t)r[j]; w[2 * j + 1] = (uint32_t)(r[j] >> 32); }
      else if constexpr (WO == 4) { w[j] = (uint32_t)r[j]; }
      else if constexpr (WO == 2) { w[j >> 1] |= ((uint32_t)r[j] & 0xFFFFu) << ((j & 1) * 16); }
      else { w[j >> 2] |= ((uint32_t)r[j] & 0xFFu) << ((j & 3) * 8); }
    }
    st_stream16(p, make_uint4(w[0], w[1], w[2], w[3]));
  } else if constexpr (BYTES == 8) {
    uint32_t w[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < V; j++) {
      if constexpr (WO == 4) { w[j] = (uint32_t)r[j]; }
      else if constexpr (WO == 2) { w[j >> 1] |= ((uint32_t)r[j] & 0xFFFFu) << ((j & 1) * 16); }
      else { w[j >> 2] |= ((uint32_t)r[j] & 0xFFu) << ((j & 3) * 8); }
    }
    st_stream8(p, make_uint2(w[0], w[1]));
  } else if constexpr (BYTES == 4) {
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < V; j++) {
      if constexpr (WO == 2) { w |= ((uint32_t)r[j] & 0xFFFFu) << ((j & 1) * 16); }
      else { w |= ((uint32_t)r[j] & 0xFFu) << ((j & 3) * 8); }
    }
    st_stream4(p, w);
  } else {  // BYTES == 2: V == 2, WO == 1
    uint16_t w = (uint16_t)(((uint32_t)r[0] & 0xFFu) | (((uint32_t)r[1] & 0xFFu) << 8));
    st_stream2(p, w);
  }
}

// ---- hashing / table -----------------------------------------------------------------------
// Fibonacci multiply + xor-fold: cheap, and mixes the high (chunk id) and low (local id) bits
// that real segmentation ids are made of.
__device__ __forceinline__ u64 hash_key(u64 k) {
  k *= 0x9E3779B97F4A7C15ull;
  return k ^ (k >> 32);
}
// Home slot of a key.  Homes are EVEN: a probe sequence starts on a 32-byte pair of slots, so the
// first two probes of a lookup are one 32-byte sector (one 256-bit load) -- the table behaves like
// a bucketed table with 2 slots per bucket and linear probing over buckets.
__device__ __forceinline__ u64 home_slot(u64 key, u64 mask) { return hash_key(key) & mask & ~1ull; }
__host__ __device__ __forceinline__ u64 sign_extend(u64 v, int w) {
  const int sh = 64 - 8 * w;
  return (u64)(((long long)(v << sh)) >> sh);
}
__host__ __device__ __forceinline__ u64 trunc_width(u64 v, int w) {
  return w == 8 ? v : (v & ((1ull << (8 * w)) - 1ull));
}

// L2-coherent 16-byte slot read (the build kernels mutate the table concurrently; a stale
// L1 line would only cost extra atomics, but could pin a hot key on the slow path forever)
__device__ __forceinline__ Slot ld_slot_cg(const Slot* s) {
  Slot r;
  asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(r.key), "=l"(r.val) : "l"(s));
  return r;
}
// read-only 16-byte slot read, L1-cached (relabel: the table is constant during the kernel)
__device__ __forceinline__ Slot ld_slot_ro(const Slot* s) {
  Slot r;
  asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(r.key), "=l"(r.val) : "l"(s));
  return r;
}

// L2-coherent one-shot slot read that the compiler may schedule early / batch (never use it to poll)
__device__ __forceinline__ Slot ld_slot_cg_batch(const Slot* s) {
  Slot r;
  asm("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(r.key), "=l"(r.val) : "l"(s));
  return r;
}
// same, but free to be scheduled early / batched by the compiler (pure load of a constant table)
__device__ __forceinline__ Slot ld_slot_ro_batch(const Slot* s) {
  Slot r;
  asm("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(r.key), "=l"(r.val) : "l"(s));
  return r;
}

// both slots of an (even-aligned) pair with one 256-bit load
struct SlotPair {
  Slot a, b;
};
__device__ __forceinline__ SlotPair ld_pair_ro(const Slot* s) {
  SlotPair r;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a.key), "=l"(r.a.val), "=l"(r.b.key), "=l"(r.b.val) : "l"(s));
  return r;
}

// same, but free to be scheduled early / batched by the compiler (pure load of a constant table)
__device__ __forceinline__ SlotPair ld_pair_ro_batch(const Slot* s) {
  SlotPair r;
  asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a.key), "=l"(r.a.val), "=l"(r.b.key), "=l"(r.b.val) : "l"(s));
  return r;
}

// Global-memory atomics spelled out in PTX: inside out-of-line (__noinline__) helpers the compiler cannot
// prove that a table pointer is global memory and emits generic ATOM.E + address-space checks instead of
// REDG / ATOMG (seen in the SASS of the build and counting kernels' slow paths).
__device__ __forceinline__ void red_add_u64(u64* p, u64 v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
}
__device__ __forceinline__ void red_min_u64(u64* p, u64 v) {
  asm volatile("red.global.min.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
}
__device__ __forceinline__ void red_max_u64(u64* p, u64 v) {
  asm volatile("red.global.max.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
}
__device__ __forceinline__ u64 atom_cas_u64(u64* p, u64 cmp, u64 v) {
  u64 old;
  asm volatile("atom.global.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "l"(__cvta_generic_to_global(p)), "l"(cmp), "l"(v) : "memory");
  return old;
}
__device__ __forceinline__ u64 atom_exch_u64(u64* p, u64 v) {
  u64 old;
  asm volatile("atom.global.exch.b64 %0, [%1], %2;" : "=l"(old) : "l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
  return old;
}

template <int OP>
__device__ __forceinline__ bool op_settled(u64 cur, u64 v) {
  if (OP == FRB_OP_MIN) return cur <= v;
  if (OP == FRB_OP_MAX) return cur >= v;
  return false;
}
template <int OP>
__device__ __forceinline__ void op_apply(u64* addr, u64 v) {
  if (OP == FRB_OP_MIN) red_min_u64(addr, v);
  else if (OP == FRB_OP_MAX) red_max_u64(addr, v);
  else if (OP == FRB_OP_ADD) red_add_u64(addr, v);
  else atom_exch_u64(addr, v);
}

// Find-or-insert `key` and combine `v` into its value.  `first` is an optional already-loaded
// copy of the home slot (speculative probe issued by the caller), else pass probe=false.
// Returns false if the probe limit was hit (table too full): meta[OVERFLOW] is raised.
// ins_slot (optional): receives the slot index if THIS call inserted the key into the slot array (not for the
// out-of-band all-ones key) -- K7 lists the slots it fills so that its gather need not scan the table.
template <int OP>
__device__ __forceinline__ bool table_update(Slot* slots, u64 mask, int direct, u64* meta, u64 key, u64 v,
                                             unsigned int* inserted, u64* ins_slot = nullptr) {
  if (direct) {
    Slot* s = slots + key;
    Slot cur = ld_slot_cg(s);
    if (cur.key != key) {
      u64 prev = atom_cas_u64(&s->key, FRB_EMPTY, key);
      if (prev == FRB_EMPTY) {
        (*inserted)++;
        if (ins_slot) *ins_slot = key;
      }
    } else if (op_settled<OP>(cur.val, v)) {
      return true;
    }
    op_apply<OP>(&s->val, v);
    return true;
  }
  if (key == FRB_EMPTY) {  // the one key the slot array cannot hold lives in the meta block
    if (atom_cas_u64(&meta[FRB_META_SIDE_PRESENT], 0ull, 1ull) == 0ull) (*inserted)++;
    op_apply<OP>(&meta[FRB_META_SIDE_VAL], v);
    return true;
  }
  u64 h = home_slot(key, mask);
#pragma unroll 1
  for (int probe = 0; probe < FRB_MAX_PROBE; probe++, h = (h + 1) & mask) {
    Slot* s = slots + h;
    Slot cur = ld_slot_cg(s);
    if (cur.key == key) {
      if (!op_settled<OP>(cur.val, v)) op_apply<OP>(&s->val, v);
      return true;
    }
    if (cur.key == FRB_EMPTY) {
      u64 prev = atom_cas_u64(&s->key, FRB_EMPTY, key);
      if (prev == FRB_EMPTY) {
        (*inserted)++;
        if (ins_slot) *ins_slot = h;
        op_apply<OP>(&s->val, v);
        return true;
      }
      if (prev == key) {
        op_apply<OP>(&s->val, v);
        return true;
      }
    }
  }
  atom_exch_u64(&meta[FRB_META_OVERFLOW], 1ull);
  return false;
}

// Read-only lookup, continuing at slot h of the key's (linear) probe sequence.  Returns true and
// *val if found.
__device__ __forceinline__ bool table_find_from(const Slot* slots, u64 mask, const u64* meta, u64 key, u64 h, u64* val) {
  if (key == FRB_EMPTY) {
    if (meta[FRB_META_SIDE_PRESENT]) { *val = meta[FRB_META_SIDE_VAL]; return true; }
    return false;
  }
#pragma unroll 1
  for (int probe = 0; probe < FRB_MAX_PROBE; probe++) {
    Slot cur = ld_slot_ro(slots + h);
    if (cur.key == key) { *val = cur.val; return true; }
    if (cur.key == FRB_EMPTY) return false;
    h = (h + 1) & mask;
  }
  return false;
}

// block-wide sum of a per-thread counter into *target (one atomic per CTA)
__device__ __forceinline__ void block_add_to(u64* target, unsigned int v) {
  __shared__ unsigned int s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  v = __reduce_add_sync(FRB_FULL_MASK, v);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_sum, v);
  __syncthreads();
  if (threadIdx.x == 0 && s_sum) atomicAdd(target, (u64)s_sum);
}

}  // namespace frb

// ------------------------------------------------------------------------------------------------
// stream_rows: the common streaming loop of the volume kernels.
//
// HBM3e needs well over 100 KB of reads in flight per SM to run at full rate; registers cannot
// hold that, so the volume is staged through shared memory by the TMA engine.  A CTA is 8 consumer
// warps plus one producer warp.  The producer's elected lane issues one 1-D bulk copy
// (cp.async.bulk, SASS UBLKCP) per 16 KiB tile -- large requests are what the TMA engine is fast
// at -- into a STAGES-deep ring and arms the stage's "full" mbarrier with the byte count.  A
// consumer warp waits on "full", reads its 4 rows of 32 x 16 B with conflict-free shared loads,
// and arrives on the stage's "empty" mbarrier; the producer refills a stage once all 8 warps have
// released it.  At 3-4 CTAs/SM x STAGES x 16 KiB this keeps 150-200 KB in flight per SM at no
// register cost, and the copy carries the 16 bytes in front of the tile so that "the element
// before mine" is always in the same staged bytes.
//
// Tile -> CTA assignment (TileMap): a CTA takes runs of `run` consecutive tiles (spatially
// adjacent rows of the volume, which the per-warp label caches live on); consecutive runs of one
// CTA are gridDim.x * run tiles apart, so all CTAs together stream through one compact window of
// the array (DRAM page locality).
//
//   row_fn(u, cur, valid, i0, prev, has_prev)
//     u        row of the warp-tile (0..STREAM_ROWS-1, a compile-time constant after unrolling)
//     cur      the lane's 16-byte vector (V = 16/W elements)
//     valid    false for lanes past the end of the array
//     i0       flat element index of the lane's first element
//     prev     the element preceding the lane's first element in memory order (exact across
//              lanes, rows, tiles and CTAs: it comes out of the same staged bytes), valid iff has_prev
//   after_tile(k) (optional) is called by every consumer warp after it has released tile number k
//     of the CTA -- also by warps whose part of the last tile lies past the end of the array.
//   tile_end(stage, vb) is called once per warp and tile after its rows.  INTERLEAVE: `stage` holds the
//     CTA tile (vector stream_vec(u, warp, lane) of it is what row_fn(u) saw), vb is the tile's first
//     vector index.  Contiguous rows: `stage` is the warp's own 2 KiB of the tile (vector u*32+lane),
//     vb the index of the warp's first vector.
//
// REVERSE walks tiles and rows from the back (component_map keeps the LAST run start).
// Kernels are launched with STREAM_THREADS threads; every lane of a consumer warp calls row_fn
// (it may use warp-wide primitives); the producer warp returns from stream_rows without calling it.
// ------------------------------------------------------------------------------------------------
namespace frb {

#ifndef FRB_STREAM_STAGES
#define FRB_STREAM_STAGES 3
#endif
static constexpr int STREAM_ROWS = 4;
static constexpr int STREAM_WARPS = 8;                                     // consumer warps
static constexpr int STREAM_THREADS = (STREAM_WARPS + 1) * 32;             // + 1 producer warp
static constexpr int STREAM_TILE_VECS = STREAM_WARPS * STREAM_ROWS * 32;   // 16-byte vectors per CTA tile (16 KiB)
static constexpr int STREAM_STAGES = FRB_STREAM_STAGES;
static constexpr int STREAM_STAGE_VECS = STREAM_TILE_VECS + 8;             // 128 B in front: predecessor vector + pad
static constexpr int STREAM_SMEM_BYTES = STREAM_STAGES * STREAM_STAGE_VECS * 16 + 2 * STREAM_STAGES * 8;

// Vector (inside the 16 KiB CTA tile) that lane `lane` of consumer warp `warp` handles in its row u.
// Rows of one warp are INTERLEAVED with those of the other warps (512-byte chunk number u * 8 + warp),
// so a warp's rows are 4 KiB apart: for volumes whose fastest axis is a multiple of 4 KiB long, a
// lane meets the voxels directly above / below its previous row one or two rows later -- what the
// "same as the row above" filters of the build and counting kernels feed on.
// IL = how many 512-byte chunks apart the rows of one warp are: 8 (all warps interleaved, rows 4 KiB
// apart), 4 / 2 / 1 (groups of IL warps interleaved: rows 2 KiB / 1 KiB / 512 B apart, for volumes with
// rows that short), 0 = contiguous rows with the legacy consumer body (kernels that write the volume).
template <int IL>
__device__ __forceinline__ int stream_vec(int u, int warp, int lane) {
  if (IL == 0) return (warp * STREAM_ROWS + u) * 32 + lane;
  if (IL == STREAM_WARPS) return (u * STREAM_WARPS + warp) * 32 + lane;  // one group: spelled out, the general form costs K1 14 %
  return ((warp / IL) * (STREAM_ROWS * IL) + (warp % IL) + u * IL) * 32 + lane;
}

// A lane remembers the last vector it saw in each of its 4 row slots.  For a volume whose fastest
// axis is q * 4 KiB long (q = 1..4) the voxels directly above row u (below, when tiles are walked
// backwards) are what the lane saw q rows earlier: slot (u -/+ q) mod 4, of this tile or the previous
// one.  DU = q, or 4 when the row length is unknown / no such slot exists ("same position, previous
// tile": a row or two further up for typical volumes -- still a good predictor, just a weaker one).
// DU is a template parameter of the kernels: selecting the slot at run time cost K1 30 %.
struct RowGeom {
  int il;  // chunk distance between the rows of a warp (stream_vec)
  int du;  // row slot distance to the voxels directly above (ref_slot_ct)
  int tv;  // vectors per tile
};
inline RowGeom row_geom(size_t row_bytes) {
  RowGeom g = {8, 4, 1024};
  if (row_bytes == 0 || (row_bytes % 16) != 0) return g;  // unknown: "same position, previous tile"
  if ((row_bytes % 512) == 0) {
    const size_t rc = row_bytes / 512;  // 512-byte chunks per row
    if (rc == 1 || rc == 2 || rc == 4) { g.il = (int)rc; g.du = 1; return g; }      // short rows: a warp's rows ARE consecutive volume rows
    if (rc == 8 || rc == 16 || rc == 24 || rc == 32) { g.du = (int)(rc / 8); return g; }  // 4 .. 16 KiB
  }
  const size_t rv = row_bytes / 16;
  if (rv <= 1024) g.tv = (int)((1024 / rv) * rv);  // a whole number of rows per tile
  return g;
}
template <bool REVERSE, int DU>
__device__ __forceinline__ constexpr int ref_slot_ct(int u) {
  return DU == 1 ? (REVERSE ? ((u + 1) & 3) : ((u + 3) & 3)) : DU == 2 ? ((u + 2) & 3) : DU == 3 ? (REVERSE ? ((u + 3) & 3) : ((u + 1) & 3)) : (u & 3);
}
// Tile bookkeeping comes in two index widths.  32-bit (TileMap): a tile is 16 KiB, so even a 180 GB volume has
// ~10^7 of them, and the read-only streaming loops are issue-bound (64-bit adds are two instructions and two
// registers each).  64-bit (TileMap64): what K3 was tuned with -- its code generation is touchy, and with the
// 32-bit iterator k_relabel_auto<8> measured 3.75 ms instead of 3.52 ms.
struct PlaneWalk {  // see plane_walk_for (host) and TileMapT's second constructor
  unsigned int plane_tiles, planes, band;
};
template <typename IDX> struct TileIterT;
template <typename IDX>
struct TileMapT {
  typedef IDX index_type;
  typedef TileIterT<IDX> iter_type;
  IDX tiles;   // tiles in the whole array
  IDX run;     // consecutive tiles per run
  IDX count;   // tiles owned by this CTA
  IDX tv;      // vectors per tile: STREAM_TILE_VECS, or less when a tile is a whole number of volume rows
  IDX first;   // first tile of this CTA (forward walks)
  IDX jump;    // distance from the last tile of a run to the first tile of this CTA's next run, minus 1
  bool planar; // plane walk (second constructor)
  // linear walk: runs of `run` tiles dealt round-robin to the CTAs
  __device__ __forceinline__ TileMapT(size_t nvec, size_t run_, size_t tile_vecs = STREAM_TILE_VECS) {
    tv = tile_vecs < 32 || tile_vecs > (size_t)STREAM_TILE_VECS ? (IDX)STREAM_TILE_VECS : (IDX)tile_vecs;
    tiles = (IDX)((nvec + tv - 1) / tv);
    if (sizeof(IDX) == 8) run = run_ < 1 ? (IDX)1 : (IDX)run_;  // (exactly the code K3 was tuned with)
    else run = run_ < 1 ? (IDX)1 : (run_ > (size_t)tiles ? (tiles ? tiles : (IDX)1) : (IDX)run_);  // clamp: products below stay in 32 bits
    if (sizeof(IDX) != 8) {  // the 64-bit iterator derives these itself (K3's tuned code path, left as it was)
      first = (IDX)blockIdx.x * run;
      jump = ((IDX)gridDim.x - 1) * run;
      planar = false;
    }
    const IDX nruns = (tiles + run - 1) / run;
    const IDX b = blockIdx.x, g = gridDim.x;
    const IDX mine = nruns > b ? (nruns - b + g - 1) / g : 0;  // runs owned by this CTA
    if (mine == 0) { count = 0; return; }
    const IDX last_start = ((mine - 1) * g + b) * run;
    const IDX last_len = tiles - last_start < run ? tiles - last_start : run;
    count = (mine - 1) * run + last_len;
  }
  // Plane walk (forward only; the host checks the preconditions, see plane_walk_ok): the array is `planes` planes of
  // `plane_tiles` full tiles; a CTA owns ONE band of `band` consecutive tiles (the same rows) of a stack of
  // consecutive planes and walks it plane after plane.  What a warp has aggregated for the segments of its band
  // is still in its shared-memory map when the same segments come round again one plane deeper, so a segment
  // costs one global update per CTA instead of one per plane.  gridDim.x >= plane_tiles / band.
  __device__ __forceinline__ explicit TileMapT(const PlaneWalk& pw) {
    const IDX plane_tiles = pw.plane_tiles, planes = pw.planes, band = pw.band;
    tv = (IDX)STREAM_TILE_VECS;
    tiles = plane_tiles * planes;
    run = band;
    jump = plane_tiles - band;
    planar = true;
    const IDX bands = plane_tiles / band;                  // bands per plane
    const IDX b = (IDX)blockIdx.x % bands, g = (IDX)blockIdx.x / bands;
    const IDX stacks = ((IDX)gridDim.x - b + bands - 1) / bands;  // CTAs that share band b: they split the planes
    const IDX z0 = (IDX)(((size_t)planes * g) / stacks), z1 = (IDX)(((size_t)planes * (g + 1)) / stacks);
    first = z0 * plane_tiles + b * band;
    count = (z1 - z0) * band;
  }
};
// walks the tiles of one CTA in ascending (or descending) order with additions only
template <typename IDX>
struct TileIterT {
  IDX t;      // current tile
  IDX o;      // offset of the current tile inside its run
  IDX run;    // tiles per run
  IDX jump;   // distance from the last tile of a run to the first tile of this CTA's next run, minus 1
  template <bool REVERSE>
  __device__ __forceinline__ void init(const TileMapT<IDX>& m) {
    run = m.run;
    jump = sizeof(IDX) == 8 ? ((IDX)gridDim.x - 1) * m.run : m.jump;
    const IDX nruns_mine = (m.count + m.run - 1) / m.run;
    const IDX last_len = m.count - (nruns_mine ? (nruns_mine - 1) * m.run : 0);
    if (REVERSE && sizeof(IDX) != 8 && m.planar) {  // last tile of the band in the deepest plane of the stack (count > 0 here)
      o = run - 1;
      t = m.first + (m.count / run - 1) * (run + jump) + o;
    } else if (REVERSE) {
      const IDX r = nruns_mine ? nruns_mine - 1 : 0;
      o = last_len ? last_len - 1 : 0;
      t = (r * gridDim.x + blockIdx.x) * run + o;
    } else {
      o = 0;
      t = sizeof(IDX) == 8 ? (IDX)blockIdx.x * run : m.first;
    }
  }
  __device__ __forceinline__ size_t tile() const { return (size_t)t; }
  template <bool REVERSE>
  __device__ __forceinline__ void step() {
    if (REVERSE) {
      if (o == 0) { o = run - 1; t -= jump + 1; } else { o--; t--; }
    } else {
      o++;
      t++;
      if (o == run) { o = 0; t += jump; }
    }
  }
};
typedef TileMapT<unsigned int> TileMap;
typedef TileMapT<size_t> TileMap64;

// Balanced plane walk (round 2, second session).  The plane walk above gives every CTA ONE band and splits the planes of
// a band among the CTAs that share it: with 444 CTAs and 64 bands, 60 bands have 7 CTAs (146 planes each) and 4 bands
// have 6 (171 planes each) -- the kernel ends 16 % later than its average CTA.  Here the work items are (band, plane)
// pairs in band-major order -- all planes of band 0, then all planes of band 1 ... -- and CTA c takes the contiguous
// range [total * c / G, total * (c + 1) / G): every CTA still walks a stack of consecutive planes of one band (and
// steps over to the next band at most a few times), and loads differ by one item at most.  The array must be exactly
// planes * plane_tiles full tiles.
struct PlaneWalkZ {
  unsigned int plane_tiles, planes, band;  // band == 0: not applicable
};
struct TileIterZ;
struct TileMapZ {
  typedef unsigned int index_type;
  typedef TileIterZ iter_type;
  unsigned int count;  // tiles owned by this CTA
  unsigned int tv;     // (interface of TileMapT; always the full tile)
  PlaneWalkZ pw;
  unsigned int t_first, z_first;  // first item of this CTA: its first tile and its plane
  unsigned int t_last, z_last;    // last item: ITS first tile and plane (reverse walks start there)
  __device__ __forceinline__ explicit TileMapZ(const PlaneWalkZ& w) {
    tv = (unsigned int)STREAM_TILE_VECS;
    pw = w;
    const unsigned long long total = (unsigned long long)(w.plane_tiles / w.band) * w.planes;  // items
    const unsigned long long i0 = total * blockIdx.x / gridDim.x, i1 = total * (blockIdx.x + 1ull) / gridDim.x;
    count = (unsigned int)(i1 - i0) * w.band;
    const unsigned long long il = i1 > i0 ? i1 - 1 : i0;
    z_first = (unsigned int)(i0 % w.planes);
    t_first = z_first * w.plane_tiles + (unsigned int)(i0 / w.planes) * w.band;
    z_last = (unsigned int)(il % w.planes);
    t_last = z_last * w.plane_tiles + (unsigned int)(il / w.planes) * w.band;
  }
};
struct TileIterZ {
  unsigned int t, o, zc;  // current tile, its offset inside the band, planes left in this band (forward) / plane index (reverse)
  PlaneWalkZ pw;          // (kernel parameters: read from the constant bank where they are used)
  template <bool REVERSE>
  __device__ __forceinline__ void init(const TileMapZ& m) {
    pw = m.pw;
    if (REVERSE) { o = pw.band - 1; t = m.t_last + o; zc = m.z_last; }
    else { o = 0; t = m.t_first; zc = pw.planes - m.z_first; }
  }
  __device__ __forceinline__ size_t tile() const { return (size_t)t; }
  template <bool REVERSE>
  __device__ __forceinline__ void step() {
    if (REVERSE) {
      if (o == 0) {
        o = pw.band - 1;
        if (zc == 0) { zc = pw.planes - 1; t += (pw.planes - 1u) * pw.plane_tiles - 1u; }  // first tile of the band in plane 0 -> last tile of the previous band in the last plane
        else { zc--; t -= pw.plane_tiles - pw.band + 1u; }
      } else { o--; t--; }
    } else {
      o++;
      t++;
      if (o == pw.band) {
        o = 0;
        if (--zc == 0) { zc = pw.planes; t -= (pw.planes - 1u) * pw.plane_tiles; }  // past the band's last plane -> the next band in plane 0
        else t += pw.plane_tiles - pw.band;
      }
    }
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Polling an mbarrier is a shared-memory access; warps that spin on it steal shared-memory bandwidth from the TMA
// engine that is trying to deliver their data, so back off between probes.
#ifdef FRB_MBAR_HINT
// experiment: let the hardware park the thread (try_wait with a suspend-time hint) instead of a nanosleep loop,
// whose instructions were 30 % of everything k_hist issued
__device__ __forceinline__ bool mbar_try_wait_hint(void* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait_hint(bar, parity, FRB_MBAR_HINT)) {}
}
#else
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned int ns = 32;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (ns < 256) ns <<= 1;
  }
}
#endif
// 1-D bulk copy global -> shared through the TMA engine, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 eviction priority of streamed data.  A volume that is read once (and written once) should not push the label
// table out of the 126 MB L2: between two uses of a label's table sector -- one row, one plane further on -- tens to
// hundreds of MB of volume stream through.  Lines brought in (or written) with evict_first are the first to go, so
// whatever else the kernel reads with the default priority stays.
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar, u64 policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
// Used by the kernels that only READ the volume (interleaved rows: K1, K7, k_hist, K6), for arrays of 256 MB and more
// (shorter ones -- pipeline chunks, small volumes -- may well be read again out of L2 by the next kernel).  Measured
// at 1024^3 (profiles/r02i_evict_first.jsonl): component_map 3.73 -> 3.57 ms, unique on 900 k labels 1.40 -> 1.37 ms,
// K1 1.285 -> 1.257 ms.  NOT for K3: it writes the lines it has just read, and with evict_first on its reads
// k_relabel_auto<8> went from 3.53 to 3.62 ms; evict_first on its stores cost another 4-6 %.
#ifndef FRB_EF_MIN_VECS
#define FRB_EF_MIN_VECS (1ull << 24)
#endif

__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// INTERLEAVE: see stream_vec.  Kernels that write the volume back (K3) keep a warp's rows contiguous:
// measured, interleaving costs them 8 % (stores of one warp 4 KiB apart), while the kernels that only
// read gain from the vertical alignment of a lane's rows.
// VAR_TILE: the tile length is map.tv (a whole number of volume rows) instead of the constant 16 KiB; a
// template flag because the constant folds into shifts in loops that are issue-bound (K1 + 18 % otherwise).
template <int W, bool REVERSE, bool NEED_PREV, int IL, bool VAR_TILE, typename MAP, typename RowFn, typename TileEndFn, typename AfterTileFn>
__device__ __forceinline__ void stream_rows_impl(const uint4* a, size_t nvec, const MAP& map, RowFn row_fn,
                                                 TileEndFn tile_end, AfterTileFn after_tile) {
  typedef typename MAP::index_type IDX;
  typedef typename MAP::iter_type ITER;
  constexpr int V = 16 / W;
  constexpr int R = STREAM_ROWS;
  constexpr int S = STREAM_STAGES;
  extern __shared__ __align__(128) unsigned char frb_dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* ring = reinterpret_cast<uint4*>(frb_dyn_smem);
  u64* full = reinterpret_cast<u64*>(frb_dyn_smem + (size_t)S * STREAM_STAGE_VECS * 16);
  u64* empty = full + S;
  const IDX ntiles = map.count;
  const size_t TV = VAR_TILE ? (size_t)map.tv : (size_t)STREAM_TILE_VECS;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], STREAM_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (ntiles == 0) return;

  if (warp == STREAM_WARPS) {
    // ---------------- producer: one elected lane feeds the ring
    if (lane == 0) {
      ITER it;
      it.template init<REVERSE>(map);
      int s = 0;
      uint32_t round = 0;
      const bool evict_first = IL != 0 && nvec >= (size_t)FRB_EF_MIN_VECS;
      const u64 policy = evict_first ? l2_policy_evict_first() : 0;
      for (IDX k = 0; k < ntiles; k++) {
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1u);  // all 8 warps released the previous fill
        const size_t vb = it.tile() * TV;
        it.template step<REVERSE>();
        size_t nv = nvec - vb < TV ? nvec - vb : TV;
        uint4* dst = ring + (size_t)s * STREAM_STAGE_VECS + 8;
        const uint4* src = a + vb;
        if (NEED_PREV && vb > 0) { dst -= 1; src -= 1; nv += 1; }
        const uint32_t bytes = (uint32_t)(nv * 16);
        mbar_expect_tx(&full[s], bytes);
        if (evict_first) bulk_g2s_hint(dst, src, bytes, &full[s], policy);
        else bulk_g2s(dst, src, bytes, &full[s]);
        if (++s == S) { s = 0; round++; }
      }
    }
    return;
  }

  // ---------------- consumers
  constexpr int ROW_STRIDE = IL ? IL * 32 : 32;  // vectors between two rows of one warp
  const int first = stream_vec<IL>(0, warp, lane);
  ITER it;
  it.template init<REVERSE>(map);
  int s = 0;
  uint32_t round = 0;
  for (IDX k = 0; k < ntiles; k++) {
    const size_t t = it.tile();
    it.template step<REVERSE>();
    if constexpr (IL != 0) {
      const size_t vb = t * TV;                // first vector of the CTA tile
      mbar_wait(&full[s], round & 1u);
      const uint4* stage = ring + (size_t)s * STREAM_STAGE_VECS + 8;
      const uint4* mine = stage + first;       // this lane's vector of row 0; row u is u * ROW_STRIDE further
      const size_t vfirst = vb + first;
      if (!VAR_TILE && vb + STREAM_TILE_VECS <= nvec) {
        // whole tile inside the array (all but the last tile): no per-lane bounds checks
#pragma unroll
        for (int uu = 0; uu < R; uu++) {
          const int u = REVERSE ? R - 1 - uu : uu;
          const uint4 cur = mine[u * ROW_STRIDE];
          u64 prev = 0;
          bool has_prev = false;
          if (NEED_PREV) {
            has_prev = (vfirst + u * ROW_STRIDE) != 0;
            if (has_prev) prev = (u64)*(reinterpret_cast<const typename UIntOf<W>::type*>(mine + u * ROW_STRIDE) - 1);
          }
          row_fn(u, cur, true, (vfirst + u * ROW_STRIDE) * V, prev, has_prev);
        }
        tile_end(stage, vb);
      } else if ((size_t)stream_vec<IL>(0, warp, 0) < TV && vb + (size_t)stream_vec<IL>(0, warp, 0) < nvec) {  // the warp's first row starts inside the tile and the array
#pragma unroll
        for (int uu = 0; uu < R; uu++) {
          const int u = REVERSE ? R - 1 - uu : uu;
          const size_t vi = vfirst + u * ROW_STRIDE;
          const bool valid = (size_t)(first + u * ROW_STRIDE) < TV && vi < nvec;
          uint4 cur = make_uint4(0, 0, 0, 0);
          u64 prev = 0;
          bool has_prev = false;
          if (valid) {
            cur = mine[u * ROW_STRIDE];
            if (NEED_PREV && vi > 0) {
              prev = (u64)*(reinterpret_cast<const typename UIntOf<W>::type*>(mine + u * ROW_STRIDE) - 1);  // element in front
              has_prev = true;
            }
          }
          row_fn(u, cur, valid, vi * V, prev, has_prev);
        }
        tile_end(stage, vb);
      }
    } else {
      // contiguous rows: tile_end gets the WARP's part of the stage and its first vector index
      // (kept as it was before the interleaved mode existed: K3 is 5 % slower through the other body)
      const size_t vb = t * STREAM_TILE_VECS + (size_t)warp * (R * 32);
      mbar_wait(&full[s], round & 1u);
      const uint4* wstage = ring + (size_t)s * STREAM_STAGE_VECS + 8 + (size_t)warp * (R * 32);
      if (vb + R * 32 <= nvec) {
        // whole warp-tile inside the array (all but the last tile): no per-lane bounds checks
#pragma unroll
        for (int uu = 0; uu < R; uu++) {
          const int u = REVERSE ? R - 1 - uu : uu;
          const int off = u * 32 + lane;
          const uint4 cur = wstage[off];
          u64 prev = 0;
          bool has_prev = false;
          if (NEED_PREV) {
            has_prev = (vb + off) != 0;
            if (has_prev) prev = (u64)*(reinterpret_cast<const typename UIntOf<W>::type*>(wstage + off) - 1);
          }
          row_fn(u, cur, true, (vb + off) * V, prev, has_prev);
        }
        tile_end(wstage, vb);
      } else if (vb < nvec) {
#pragma unroll
        for (int uu = 0; uu < R; uu++) {
          const int u = REVERSE ? R - 1 - uu : uu;
          const int off = u * 32 + lane;
          const size_t vi = vb + off;
          const bool valid = vi < nvec;
          uint4 cur = make_uint4(0, 0, 0, 0);
          u64 prev = 0;
          bool has_prev = false;
          if (valid) {
            cur = wstage[off];
            if (NEED_PREV && vi > 0) {
              prev = (u64)*(reinterpret_cast<const typename UIntOf<W>::type*>(wstage + off) - 1);  // element in front
              has_prev = true;
            }
          }
          row_fn(u, cur, valid, vi * V, prev, has_prev);
        }
        tile_end(wstage, vb);
      }
    }
    __syncwarp();  // every lane of this warp is done with the stage
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == S) { s = 0; round++; }
    after_tile(k);  // every consumer warp, every tile of the CTA (k is the same in all of them)
  }
}
template <int W, bool REVERSE, bool NEED_PREV, int IL, bool VAR_TILE, typename IDX, typename RowFn, typename TileEndFn, typename AfterTileFn>
__device__ __forceinline__ void stream_rows(const uint4* a, size_t nvec, const TileMapT<IDX>& map, RowFn row_fn,
                                            TileEndFn tile_end, AfterTileFn after_tile) {
  stream_rows_impl<W, REVERSE, NEED_PREV, IL, VAR_TILE>(a, nvec, map, row_fn, tile_end, after_tile);
}
template <int W, bool REVERSE, bool NEED_PREV, int IL, typename IDX, typename RowFn, typename TileEndFn, typename AfterTileFn>
__device__ __forceinline__ void stream_rows(const uint4* a, size_t nvec, const TileMapT<IDX>& map, RowFn row_fn,
                                            TileEndFn tile_end, AfterTileFn after_tile) {
  stream_rows_impl<W, REVERSE, NEED_PREV, IL, false>(a, nvec, map, row_fn, tile_end, after_tile);
}
template <int W, bool REVERSE, bool NEED_PREV, int IL, typename IDX, typename RowFn, typename TileEndFn>
__device__ __forceinline__ void stream_rows(const uint4* a, size_t nvec, const TileMapT<IDX>& map, RowFn row_fn,
                                            TileEndFn tile_end) {
  stream_rows_impl<W, REVERSE, NEED_PREV, IL, false>(a, nvec, map, row_fn, tile_end, [](size_t) {});
}
// the same loop over a balanced plane walk (TileMapZ): full tiles only
template <int W, bool REVERSE, bool NEED_PREV, int IL, typename RowFn, typename TileEndFn>
__device__ __forceinline__ void stream_rows_z(const uint4* a, size_t nvec, const TileMapZ& map, RowFn row_fn, TileEndFn tile_end) {
  stream_rows_impl<W, REVERSE, NEED_PREV, IL, false>(a, nvec, map, row_fn, tile_end, [](size_t) {});
}
// barrier among the consumer warps of a streaming CTA (the producer warp never joins it)
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(STREAM_WARPS * 32) : "memory"); }

// host side: tiles in an n-element array (+1 so that the CTA owning the scalar tail exists)
inline size_t stream_tiles(size_t n, int w) { return (n / (16 / w) + STREAM_TILE_VECS - 1) / STREAM_TILE_VECS + 1; }
int tile_run();  // tuning knob (frb_runtime.cu): consecutive tiles per CTA run
// Run length for a launch of `grid` CTAs over `tiles` tiles: short arrays (label tables, pipeline
// chunks) get shorter runs, so that every CTA of the grid has work.  Decided on the host: the same
// clamp inside the kernels cost k_hist 8 % (register allocation of its hot loop).
inline int fair_run(int grid, size_t tiles) {
  const size_t g = (size_t)(grid < 1 ? 1 : grid);
  const size_t fair = tiles / (2 * g);
  const size_t rmax = (size_t)tile_run();
  if (rmax > fair) return (int)(fair < 1 ? 1 : fair);
  // Mid-sized arrays: the number of runs should be (close to) a multiple of the grid.  With runs of exactly 64 tiles a
  // 512^3 uint64 volume is 1024 runs for 444 CTAs -- 136 CTAs get three runs, the others two, and the kernel ends 30 %
  // later than its average CTA: K1 took 0.243 ms there (4.4 TB/s) and takes 0.183 ms with k = 3 runs of 50 tiles per CTA.
  // Only while a CTA has few runs (k <= 8): at 1024^3 (k = 19) the imbalance is 3 %, and runs of 63 instead of 64 tiles
  // -- 1008 KiB pieces instead of 1 MiB ones -- cost the big-table K3 8 % (mask_except 4.14 -> 4.46 ms), so long arrays
  // keep the power of two.
  const size_t k = (tiles + g * rmax - 1) / (g * rmax);
  if (k > 8) return (int)rmax;
  const size_t run = (tiles + k * g - 1) / (k * g);
  return (int)(run < 1 ? 1 : run);
}

// Plane walk of a read-only streaming kernel (TileMapT's second constructor): possible when the array is a whole
// number of planes, a plane a whole number of 16 KiB tiles, the grid has a CTA for every band of a plane, and
// every CTA gets a stack of at least 4 planes (otherwise nothing comes round again).  Returns the band length in
// tiles (0: walk linearly).  Bands are 8 tiles (128 KiB contiguous, 32 rows of a 1024-wide uint32 volume): short
// enough for a warp's 128-entry map to hold the segments that cross its columns of the band.
inline PlaneWalk plane_walk_for(size_t nbytes, size_t plane_bytes, int grid) {
  PlaneWalk pw = {0, 0, 0};
  const size_t tile_bytes = (size_t)STREAM_TILE_VECS * 16;
  if (plane_bytes == 0 || plane_bytes % tile_bytes != 0 || nbytes % plane_bytes != 0) return pw;
  const size_t pt = plane_bytes / tile_bytes, planes = nbytes / plane_bytes;
  const size_t band = 8;
  if (pt % band != 0 || pt / band > (size_t)grid || pt > (1u << 20) || planes > (1u << 20)) return pw;
  const size_t stacks = (size_t)grid / (pt / band);   // CTAs per band (at least)
  if (planes / (stacks + 1) < 4) return pw;
  pw.plane_tiles = (unsigned int)pt;
  pw.planes = (unsigned int)planes;
  pw.band = (unsigned int)band;
  return pw;
}

// Balanced plane walk (TileMapZ): possible when the array is a whole number of planes, a plane a whole number of bands of
// `band` 16 KiB tiles, and every CTA gets a stack of at least 4 planes.
inline PlaneWalkZ plane_walk_z_for(size_t nbytes, size_t plane_bytes, int grid, size_t band) {
  PlaneWalkZ pw = {0, 0, 0};
  const size_t tile_bytes = (size_t)STREAM_TILE_VECS * 16;
  if (plane_bytes == 0 || band == 0 || grid < 1 || plane_bytes % (tile_bytes * band) != 0 || nbytes % plane_bytes != 0) return pw;
  const size_t pt = plane_bytes / tile_bytes, planes = nbytes / plane_bytes;
  if (pt > (1u << 20) || planes > (1u << 20) || pt * planes > 0xFFFFFFF0ull || (pt / band) * planes < 4 * (size_t)grid) return pw;
  pw.plane_tiles = (unsigned int)pt;
  pw.planes = (unsigned int)planes;
  pw.band = (unsigned int)band;
  return pw;
}

// Persistent grid for a streaming kernel (STREAM_THREADS threads, STREAM_SMEM_BYTES of dynamic shared memory):
// as many CTAs as are co-resident on the device, capped by the number of tiles.  Also opts the
// kernel in to > 48 KiB of dynamic shared memory.
template <typename Kernel>
inline int stream_grid(Kernel kernel, size_t tiles) {
  static int per_sm = 0;  // one instantiation per kernel: the occupancy query runs once
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM_BYTES);  // per device: always
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, STREAM_THREADS, STREAM_SMEM_BYTES) != cudaSuccess || per_sm < 1)
      per_sm = 2;
  }
  size_t b = (size_t)per_sm * (size_t)sm_count();
  if (tiles < 1) tiles = 1;
  return (int)(b < tiles ? b : tiles);
}

}  // namespace frb
