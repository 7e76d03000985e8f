// frb_common.cuh -- shared device helpers for the fastremap_b200 kernels (sm_100a).
//
// Every kernel in this library is an HBM-bound integer streaming kernel.  The common
// shape is: a warp reads one "row" of 32 x 16 B = 512 contiguous bytes per load
// instruction (ld.global.nc.L1::no_allocate.v4 -- the volume is read once, keep it out of
// L1 so the label table can live there), each lane unpacks V = 16 / W elements of width W,
// and a CTA of 256 threads owns a tile of ROWS_PER_WARP rows per warp.  Grids are sized to
// a multiple of the SM count (148 on B200) and walk tiles with a grid stride.
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
  if constexpr (BYTES == 16) {
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

template <int OP>
__device__ __forceinline__ bool op_settled(u64 cur, u64 v) {
  if (OP == FRB_OP_MIN) return cur <= v;
  if (OP == FRB_OP_MAX) return cur >= v;
  return false;
}
template <int OP>
__device__ __forceinline__ void op_apply(u64* addr, u64 v) {
  if (OP == FRB_OP_MIN) atomicMin(addr, v);
  else if (OP == FRB_OP_MAX) atomicMax(addr, v);
  else if (OP == FRB_OP_ADD) atomicAdd(addr, v);
  else atomicExch(addr, v);
}

// Find-or-insert `key` and combine `v` into its value.  `first` is an optional already-loaded
// copy of the home slot (speculative probe issued by the caller), else pass probe=false.
// Returns false if the probe limit was hit (table too full): meta[OVERFLOW] is raised.
template <int OP>
__device__ __forceinline__ bool table_update(Slot* slots, u64 mask, int direct, u64* meta, u64 key, u64 v,
                                             unsigned int* inserted) {
  if (direct) {
    Slot* s = slots + key;
    Slot cur = ld_slot_cg(s);
    if (cur.key != key) {
      u64 prev = atomicCAS(&s->key, FRB_EMPTY, key);
      if (prev == FRB_EMPTY) (*inserted)++;
    } else if (op_settled<OP>(cur.val, v)) {
      return true;
    }
    op_apply<OP>(&s->val, v);
    return true;
  }
  if (key == FRB_EMPTY) {  // the one key the slot array cannot hold lives in the meta block
    if (atomicCAS(&meta[FRB_META_SIDE_PRESENT], 0ull, 1ull) == 0ull) (*inserted)++;
    op_apply<OP>(&meta[FRB_META_SIDE_VAL], v);
    return true;
  }
  u64 h = hash_key(key) & mask;
#pragma unroll 1
  for (int probe = 0; probe < FRB_MAX_PROBE; probe++) {
    Slot* s = slots + h;
    Slot cur = ld_slot_cg(s);
    if (cur.key == key) {
      if (!op_settled<OP>(cur.val, v)) op_apply<OP>(&s->val, v);
      return true;
    }
    if (cur.key == FRB_EMPTY) {
      u64 prev = atomicCAS(&s->key, FRB_EMPTY, key);
      if (prev == FRB_EMPTY) {
        (*inserted)++;
        op_apply<OP>(&s->val, v);
        return true;
      }
      if (prev == key) {
        op_apply<OP>(&s->val, v);
        return true;
      }
    }
    h = (h + 1) & mask;
  }
  atomicExch(&meta[FRB_META_OVERFLOW], 1ull);
  return false;
}

// Read-only lookup.  Returns true and *val if found.
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
