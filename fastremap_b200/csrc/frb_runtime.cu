// frb_runtime.cu -- launch bookkeeping, error reporting, table init/insert/export entry points.
#include <stdio.h>
#include <string.h>

#include "frb_common.cuh"
#include "frb_scan.cuh"

namespace frb {

unsigned long long g_launches = 0;
static char g_error[512] = "";
int g_sm_count = 0;

void set_error(const char* msg) {
  strncpy(g_error, msg, sizeof(g_error) - 1);
  g_error[sizeof(g_error) - 1] = 0;
}

int check_launch(const char* what) {
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    set_error(buf);
    return FRB_ERR_CUDA;
  }
  return FRB_OK;
}

int g_tile_run = 64;
int tile_run() { return g_tile_run; }
int plane_band() {
  static const int band = getenv("FRB_PLANE_BAND") ? atoi(getenv("FRB_PLANE_BAND")) : 16;  // measured 4 .. 64: profiles/r02t_plane_walk_bands.jsonl
  return band < 0 ? 0 : band;
}
int g_relabel_variant = -1;
int relabel_variant() { return g_relabel_variant; }

int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count = n;
  }
  return g_sm_count;
}

// ---------------------------------------------------------------------------------- table init
__device__ __forceinline__ void meta_init(u64* meta, u64 init_val) {
  if (blockIdx.x == 0 && threadIdx.x < FRB_META_WORDS) {
    u64 v = 0;
    if (threadIdx.x == FRB_META_MISSING_INDEX) v = FRB_EMPTY;
    if (threadIdx.x == FRB_META_SIDE_VAL) v = init_val;
    meta[threadIdx.x] = v;
  }
}
// WIDE: two slots (32 bytes, one 256-bit store) per thread and iteration -- needs 32-byte aligned slots and an even cap
template <bool WIDE>
__global__ void __launch_bounds__(256) k_table_init(Slot* slots, u64 cap, u64* meta, u64 init_val) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (WIDE) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < cap / 2; p += stride)
      asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(slots + 2 * p), "l"(FRB_EMPTY), "l"(init_val), "l"(FRB_EMPTY), "l"(init_val) : "memory");
  } else {
    const ulonglong2 e = make_ulonglong2(FRB_EMPTY, init_val);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < cap; p += stride) *reinterpret_cast<ulonglong2*>(slots + p) = e;
  }
  meta_init(meta, init_val);
}
__global__ void __launch_bounds__(32) k_meta_init(u64* meta, u64 init_val) { meta_init(meta, init_val); }

// ---------------------------------------------------------------------------------- insert
template <int OP>
__global__ void __launch_bounds__(256)
k_table_insert(const void* __restrict__ keys, int kw, const void* __restrict__ vals, int vw, u64 val_offset, size_t L,
               int direct, Slot* slots, u64 mask, u64* meta) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned int inserted = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += stride) {
    u64 k = ld_elem_rt(keys, kw, i);
    u64 v = vals ? ld_elem_rt(vals, vw, i) : (u64)i + val_offset;
    table_update<OP>(slots, mask, direct, meta, k, v, &inserted);
  }
  block_add_to(&meta[FRB_META_COUNT], inserted);
}

// the multi-GPU merge: `world` segments of an all-gathered buffer, segment r = [m keys | m values]
// (uint64 each) of which the first metas[r][COUNT] pairs are valid
template <int OP>
__global__ void __launch_bounds__(256)
k_table_insert_gathered(const u64* __restrict__ gathered, u64 m, int world, const u64* __restrict__ metas, int direct,
                        Slot* slots, u64 mask, u64* meta) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t total = (size_t)m * (size_t)world;
  unsigned int inserted = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / m, j = i - r * m;
    if (j >= metas[r * FRB_META_WORDS + FRB_META_COUNT]) continue;
    const u64* seg = gathered + r * 2 * m;
    table_update<OP>(slots, mask, direct, meta, seg[j], seg[m + j], &inserted);
  }
  block_add_to(&meta[FRB_META_COUNT], inserted);
}

// One block of `block` uint64 per rank: [FRB_META_WORDS words of the rank's table meta | m keys | m values];
// the exporter dropped pairs beyond m.  Used when m was GUESSED before the counts were known (one
// collective, no host round trip): a rank whose count exceeds m, whose own build overflowed, or whose
// table is over its load limit makes the destination's OVERFLOW flag go up, which parks the kernels
// queued behind this one (K2 / K3 check it) and sends the host down the two-phase path.
template <int OP>
__global__ void __launch_bounds__(256)
k_table_insert_packed(const u64* __restrict__ packed, u64 m, int world, u64 local_cap, int direct, Slot* slots, u64 mask, u64* meta) {
  const u64 block = FRB_META_WORDS + 2 * m;
  bool bad = false;
  for (int r = 0; r < world; r++) {
    const u64* mt = packed + (size_t)r * block;
    const u64 cnt = mt[FRB_META_COUNT];
    bad |= cnt > m || mt[FRB_META_OVERFLOW] != 0 || (local_cap != 0 && 2 * cnt > local_cap);
  }
  if (bad) {  // the same decision in every thread
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(&meta[FRB_META_OVERFLOW], 1ull);
    return;
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t total = (size_t)m * (size_t)world;
  unsigned int inserted = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / m, j = i - r * m;
    const u64* seg = packed + r * block;
    if (j >= seg[FRB_META_COUNT]) continue;
    table_update<OP>(slots, mask, direct, meta, seg[FRB_META_WORDS + j], seg[FRB_META_WORDS + m + j], &inserted);
  }
  block_add_to(&meta[FRB_META_COUNT], inserted);
}

// Project a list of keys through one (hashed) table into another: dst[key] = src[key] for every listed key that
// src holds.  Used by the multi-GPU renumber: after K2 every rank copies the global ranks of ITS OWN labels from the
// merged table into a compact lookup table for K3.  The list length may still be on the device (count_dev).
__global__ void __launch_bounds__(256)
k_table_project(const u64* __restrict__ keys, const u64* __restrict__ count_dev, u64 max_count, const Slot* __restrict__ src, u64 src_mask,
                const u64* __restrict__ src_meta, Slot* dst, u64 dst_mask, u64* dst_meta) {
  if (__ldcg(&src_meta[FRB_META_OVERFLOW]) != 0) {  // the merge failed (sizes guessed too small): park what is queued behind
    if (blockIdx.x == 0 && threadIdx.x == 0) atom_exch_u64(&dst_meta[FRB_META_OVERFLOW], 1ull);
    return;
  }
  u64 n = max_count;
  if (count_dev) {
    const u64 c = __ldcg(count_dev);
    n = c < n ? c : n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) dst_meta[FRB_META_ASSIGNED] = __ldcg(&src_meta[FRB_META_ASSIGNED]);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned int inserted = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 key = keys[i];
    u64 val = 0;
    if (table_find_from(src, src_mask, src_meta, key, home_slot(key, src_mask), &val))
      table_update<FRB_OP_SET>(dst, dst_mask, 0, dst_meta, key, val, &inserted);
  }
  block_add_to(&dst_meta[FRB_META_COUNT], inserted);
}

__global__ void __launch_bounds__(256)
k_table_resolve(const void* __restrict__ vals, int vw, Slot* slots, u64 cap, u64* meta) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += stride) {
    if (slots[s].key != FRB_EMPTY) slots[s].val = ld_elem_rt(vals, vw, (size_t)slots[s].val);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && meta[FRB_META_SIDE_PRESENT])
    meta[FRB_META_SIDE_VAL] = ld_elem_rt(vals, vw, (size_t)meta[FRB_META_SIDE_VAL]);
}

// ---------------------------------------------------------------------------------- export
// mode 0: (key bits, value); mode 1: (value, slot index)
struct ExportSel {
  u64* a_out;
  u64* b_out;
  int mode;
  __device__ __forceinline__ bool pick(u64 key, u64) const { return key != FRB_EMPTY; }
  __device__ __forceinline__ u64 fetch(u64, u64) const { return 0; }
  __device__ __forceinline__ void emit(u64 pos, u64 key, u64 val, u64 slot, u64) const {
    if (mode == 0) { a_out[pos] = key; b_out[pos] = val; }
    else { a_out[pos] = val; b_out[pos] = slot; }
  }
};
__global__ void k_export_side(u64 cap, u64* meta, ExportSel sel, u64 capacity) {
  if (threadIdx.x == 0 && meta[FRB_META_SIDE_PRESENT]) {  // the out-of-band all-ones key ("slot cap")
    const u64 pos = atomicAdd(&meta[FRB_META_LIST_COUNT], 1ull);
    if (pos < capacity) sel.emit(pos, FRB_EMPTY, meta[FRB_META_SIDE_VAL], cap, 0);
  }
}

int launch_table_export(const Slot* slots, u64 cap, u64* meta, u64* a_out, u64* b_out, size_t capacity, int mode,
                        cudaStream_t st) {
  cudaMemsetAsync(&meta[FRB_META_LIST_COUNT], 0, sizeof(u64), st);
  ExportSel sel;
  sel.a_out = a_out;
  sel.b_out = b_out;
  sel.mode = mode;
  int rc = launch_table_scan(slots, cap, &meta[FRB_META_LIST_COUNT], (u64)capacity, sel, "k_table_scan<export>", st);
  if (rc) return rc;
  k_export_side<<<1, 32, 0, st>>>(cap, meta, sel, (u64)capacity);
  return check_launch("k_export_side");
}

}  // namespace frb

using namespace frb;

namespace frb { extern int g_sm_count; extern int g_tile_run; extern int g_relabel_variant; }

extern "C" {

size_t frb_table_bytes(uint64_t cap) { return (size_t)cap * sizeof(Slot); }

uint64_t frb_table_direct_cap(int dtype) {
  if (!dtype_ok(dtype)) return 0;
  const int w = dtype_width(dtype);
  return w == 1 ? 256ull : (w == 2 ? 65536ull : 0ull);
}

int frb_table_init(void* slots, uint64_t cap, uint64_t* meta, uint64_t init_val, void* stream) {
  if (!slots || !meta || cap == 0 || (cap & (cap - 1)) != 0 || !aligned16(slots)) {
    set_error("frb_table_init: bad arguments (cap must be a power of two, slots 16-byte aligned)");
    return FRB_ERR_BAD_ARG;
  }
  if (init_val == FRB_EMPTY && cap >= (1ull << 16)) {  // every byte is 0xFF: the driver's memset is the fastest fill there is
    if (cudaMemsetAsync(slots, 0xFF, (size_t)cap * sizeof(Slot), (cudaStream_t)stream) != cudaSuccess) {
      set_error("frb_table_init: cudaMemsetAsync failed");
      return FRB_ERR_CUDA;
    }
    k_meta_init<<<1, 32, 0, (cudaStream_t)stream>>>((u64*)meta, init_val);
    return check_launch("k_meta_init");
  }
  Grid g = grid_for((size_t)cap, 256 * 8, 8);
  if (cap >= 2 && ((uintptr_t)slots & 31) == 0) k_table_init<true><<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>((Slot*)slots, cap, (u64*)meta, init_val);
  else k_table_init<false><<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>((Slot*)slots, cap, (u64*)meta, init_val);
  return check_launch("k_table_init");
}

int frb_table_insert(const void* keys, int key_width, const void* vals, int val_width, uint64_t val_offset, size_t L,
                     int op, int direct, void* slots, uint64_t cap, uint64_t* meta, void* stream) {
  if (!slots || !meta || (L && !keys) || (cap & (cap - 1)) != 0) {
    set_error("frb_table_insert: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (L == 0) return FRB_OK;
  Grid g = grid_for(L, 256 * 4, 8);
  cudaStream_t st = (cudaStream_t)stream;
  Slot* s = (Slot*)slots;
  u64* m = (u64*)meta;
  switch (op) {
    case FRB_OP_MIN: k_table_insert<FRB_OP_MIN><<<g.blocks, g.threads, 0, st>>>(keys, key_width, vals, val_width, val_offset, L, direct, s, cap - 1, m); break;
    case FRB_OP_MAX: k_table_insert<FRB_OP_MAX><<<g.blocks, g.threads, 0, st>>>(keys, key_width, vals, val_width, val_offset, L, direct, s, cap - 1, m); break;
    case FRB_OP_ADD: k_table_insert<FRB_OP_ADD><<<g.blocks, g.threads, 0, st>>>(keys, key_width, vals, val_width, val_offset, L, direct, s, cap - 1, m); break;
    case FRB_OP_SET: k_table_insert<FRB_OP_SET><<<g.blocks, g.threads, 0, st>>>(keys, key_width, vals, val_width, val_offset, L, direct, s, cap - 1, m); break;
    default: set_error("frb_table_insert: bad op"); return FRB_ERR_BAD_ARG;
  }
  return check_launch("k_table_insert");
}

int frb_table_insert_gathered(const uint64_t* gathered, uint64_t m, int world, const uint64_t* metas, int op, int direct,
                              void* slots, uint64_t cap, uint64_t* meta, void* stream) {
  if (!gathered || !metas || !slots || !meta || m == 0 || world < 1 || (cap & (cap - 1)) != 0) {
    set_error("frb_table_insert_gathered: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  Grid g = grid_for((size_t)m * (size_t)world, 256 * 4, 8);
  cudaStream_t st = (cudaStream_t)stream;
  Slot* s = (Slot*)slots;
  u64* mt = (u64*)meta;
  const u64* gp = (const u64*)gathered;
  const u64* ms = (const u64*)metas;
  switch (op) {
    case FRB_OP_MIN: k_table_insert_gathered<FRB_OP_MIN><<<g.blocks, g.threads, 0, st>>>(gp, m, world, ms, direct, s, cap - 1, mt); break;
    case FRB_OP_MAX: k_table_insert_gathered<FRB_OP_MAX><<<g.blocks, g.threads, 0, st>>>(gp, m, world, ms, direct, s, cap - 1, mt); break;
    case FRB_OP_ADD: k_table_insert_gathered<FRB_OP_ADD><<<g.blocks, g.threads, 0, st>>>(gp, m, world, ms, direct, s, cap - 1, mt); break;
    case FRB_OP_SET: k_table_insert_gathered<FRB_OP_SET><<<g.blocks, g.threads, 0, st>>>(gp, m, world, ms, direct, s, cap - 1, mt); break;
    default: set_error("frb_table_insert_gathered: bad op"); return FRB_ERR_BAD_ARG;
  }
  return check_launch("k_table_insert_gathered");
}

int frb_table_insert_packed(const uint64_t* packed, uint64_t m, int world, uint64_t local_cap, int op, int direct, void* slots,
                            uint64_t cap, uint64_t* meta, void* stream) {
  if (!packed || !slots || !meta || m == 0 || world < 1 || (cap & (cap - 1)) != 0) {
    set_error("frb_table_insert_packed: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  Grid g = grid_for((size_t)m * (size_t)world, 256 * 4, 8);
  cudaStream_t st = (cudaStream_t)stream;
  Slot* s = (Slot*)slots;
  u64* mt = (u64*)meta;
  const u64* pp = (const u64*)packed;
  switch (op) {
    case FRB_OP_MIN: k_table_insert_packed<FRB_OP_MIN><<<g.blocks, g.threads, 0, st>>>(pp, m, world, local_cap, direct, s, cap - 1, mt); break;
    case FRB_OP_MAX: k_table_insert_packed<FRB_OP_MAX><<<g.blocks, g.threads, 0, st>>>(pp, m, world, local_cap, direct, s, cap - 1, mt); break;
    case FRB_OP_ADD: k_table_insert_packed<FRB_OP_ADD><<<g.blocks, g.threads, 0, st>>>(pp, m, world, local_cap, direct, s, cap - 1, mt); break;
    case FRB_OP_SET: k_table_insert_packed<FRB_OP_SET><<<g.blocks, g.threads, 0, st>>>(pp, m, world, local_cap, direct, s, cap - 1, mt); break;
    default: set_error("frb_table_insert_packed: bad op"); return FRB_ERR_BAD_ARG;
  }
  return check_launch("k_table_insert_packed");
}

int frb_table_project(const uint64_t* keys, const uint64_t* count_dev, uint64_t max_count, const void* src_slots, uint64_t src_cap,
                      const uint64_t* src_meta, void* dst_slots, uint64_t dst_cap, uint64_t* dst_meta, void* stream) {
  if (!keys || !src_slots || !src_meta || !dst_slots || !dst_meta || (src_cap & (src_cap - 1)) != 0 || (dst_cap & (dst_cap - 1)) != 0 || max_count == 0) {
    set_error("frb_table_project: bad arguments (hashed tables, power-of-two capacities)");
    return FRB_ERR_BAD_ARG;
  }
  Grid g = grid_for((size_t)max_count, 256 * 2, 8);
  k_table_project<<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>((const u64*)keys, (const u64*)count_dev, max_count, (const Slot*)src_slots,
                                                                    src_cap - 1, (const u64*)src_meta, (Slot*)dst_slots, dst_cap - 1, (u64*)dst_meta);
  return check_launch("k_table_project");
}

int frb_table_resolve(const void* vals, int val_width, void* slots, uint64_t cap, uint64_t* meta, void* stream) {
  if (!vals || !slots || !meta) { set_error("frb_table_resolve: bad arguments"); return FRB_ERR_BAD_ARG; }
  Grid g = grid_for((size_t)cap, 256 * 8, 8);
  k_table_resolve<<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>(vals, val_width, (Slot*)slots, cap, (u64*)meta);
  return check_launch("k_table_resolve");
}

int frb_table_export(const void* slots, uint64_t cap, uint64_t* meta, uint64_t* keys_out, uint64_t* vals_out,
                     size_t capacity, void* stream) {
  if (!slots || !meta || (capacity && (!keys_out || !vals_out))) { set_error("frb_table_export: bad arguments"); return FRB_ERR_BAD_ARG; }
  return launch_table_export((const Slot*)slots, cap, (u64*)meta, (u64*)keys_out, (u64*)vals_out, capacity, 0, (cudaStream_t)stream);
}

int frb_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FRB_ERR_CUDA; }
  g_sm_count = 0;
  return FRB_OK;
}

void frb_set_tile_run(int run) { g_tile_run = run < 1 ? 1 : run; }
void frb_set_relabel_variant(int variant) { g_relabel_variant = variant < 0 ? -1 : (variant ? 1 : 0); }

uint64_t frb_launch_count(void) { return g_launches; }
const char* frb_last_error(void) { return g_error; }

}  // extern "C"
