// frb_relabel.cu -- K3: the per-voxel table lookup + store, with the refit narrowing fused in.
//
// Replaces the write half of _renumber (fastremap.pyx:240/:244/:248), _remap (:749-767),
// _mask_except (:516-525), remap_from_array_kv (:817-825), remap_from_array (:785-789) and the
// astype of refit (:301).
//
// Every lane looks up all V = 16/W elements of its 16-byte vector.  Equal neighbours inside a
// warp hit the same 16-byte slot, so the hardware coalesces them into one L1 request and the
// reference's serial "same as previous voxel" shortcut is not needed; the first probe of each
// element is issued unconditionally so all V loads are in flight before the first compare.
// The table is read-only here (ld.global.nc, L1-allocating); the volume streams past L1.
// Tables beyond L1 and 64-bit labels: k_relabel_batched (two rows of home-pair loads in flight per warp) or
// k_relabel_runs (one lookup per run of equal labels, a warp-tile's runs in flight together), chosen per call by a
// run-density sampler.
#include "frb_internal.cuh"
#include "frb_build.cuh"

namespace frb {

struct RelabelArgs {
  const Slot* slots;
  u64 mask;
  const u64* meta_ro;
  u64* meta;
  u64 fill;
  u64 add;
  int direct;
  int policy;
  int zero_fixed;
  int guard;
};

// what to store for a label that is / is not in the table
__device__ __forceinline__ u64 relabel_hit(const RelabelArgs& A, u64 key, u64 val) {
  return A.policy == FRB_MISS_FILL ? key : val + A.add;
}
__device__ __forceinline__ u64 relabel_miss(const RelabelArgs& A, u64 key, size_t idx) {
  if (A.policy == FRB_MISS_KEEP) return key;
  if (A.policy == FRB_MISS_FILL) return A.fill;
  atomicMin(&A.meta[FRB_META_MISSING_INDEX], (u64)idx);
  return key;
}

// The first load of a lookup: one 16-byte slot of a direct table, or the 32-byte home PAIR of a
// hashed table (homes are even, see home_slot): the first two probes for the price of one sector.
struct Probe {
  Slot a, b;
};
template <bool HASHED>
__device__ __forceinline__ Probe first_probe(const RelabelArgs& A, u64 key, u64* h) {
  Probe p;
  if (!HASHED || A.direct) {
    *h = key;
    p.a = ld_slot_ro(A.slots + key);
    p.b.key = FRB_EMPTY;
    p.b.val = 0;
  } else {
    *h = home_slot(key, A.mask);
    const SlotPair sp = ld_pair_ro(A.slots + *h);
    p.a = sp.a;
    p.b = sp.b;
  }
  return p;
}

// result for one element whose first probe has already been loaded
template <bool HASHED>
__device__ __forceinline__ u64 relabel_one(const RelabelArgs& A, u64 key, u64 h, const Probe& p, size_t idx) {
  if (A.zero_fixed && key == 0) return 0;
  if (key != FRB_EMPTY) {
    if (p.a.key == key) return relabel_hit(A, key, p.a.val);
    if (!HASHED || A.direct || p.a.key == FRB_EMPTY) return relabel_miss(A, key, idx);
    if (p.b.key == key) return relabel_hit(A, key, p.b.val);
    if (p.b.key == FRB_EMPTY) return relabel_miss(A, key, idx);
  }
  // rare: the all-ones key (meta block) or a probe sequence longer than the home pair
  u64 val = 0;
  if (table_find_from(A.slots, A.mask, A.meta_ro, key, key == FRB_EMPTY ? h : ((h + 2) & A.mask), &val)) return relabel_hit(A, key, val);
  return relabel_miss(A, key, idx);
}

// All V results of one 16-byte vector of 8-, 16- or 32-bit labels.  Neighbouring voxels mostly carry the
// same label, and a lookup (one divergent 16- or 32-byte load per lane) costs far more than a compare:
// with one lookup per voxel K3 cost ~2.1 ms per 2^30 voxels at EVERY width, i.e. 8-, 16- and 32-bit
// volumes ran at 0.15 / 0.27 / 0.61 of the HBM roofline.  So the voxels of a vector are resolved one after
// the other with a one-entry cache: a voxel equal to its predecessor reuses the predecessor's result.
// (Batching 16 probes up front also needed 119 registers.)  64-bit vectors hold 2 voxels, are HBM-bound
// already and their code is codegen-sensitive: they keep both probes issued up front (see the kernels).
template <int W, bool HASHED>
__device__ __forceinline__ void relabel_vector_cached(const RelabelArgs& A, const uint4& cur, size_t i0, u64 (&r)[16 / W]) {
  constexpr int V = 16 / W;
  u64 prev_key = 0, prev_r = 0;
#pragma unroll
  for (int j = 0; j < V; j++) {
    const u64 key = vec_get<W>(cur, j);
    if (j == 0 || key != prev_key) {
      u64 h;
      const Probe first = first_probe<HASHED>(A, key, &h);
      prev_r = relabel_one<HASHED>(A, key, h, first, i0 + j);
      prev_key = key;
    }
    r[j] = prev_r;
  }
}

// 32-bit vectors (4 voxels): one probe per DISTINCT label of the vector.  The one-entry cache above still
// makes the warp walk the probe path at every position where any lane changes label (ncu: 209 warp
// instructions per row, 14 of 32 lanes active, issue-bound at 72 %); here the warp iterates max-over-lanes
// of the number of distinct labels in a lane's vector (mostly 1 or 2) and each probe's result is handed to
// every position that carries the label.
template <bool HASHED>
__device__ __forceinline__ void relabel_vector_distinct4(const RelabelArgs& A, const uint4& cur, size_t i0, u64 (&r)[4]) {
  const unsigned int k32[4] = {cur.x, cur.y, cur.z, cur.w};
  unsigned int todo = 0xFu;
#pragma unroll 1
  while (todo) {
    const int j = __ffs(todo) - 1;
    const unsigned int k = j == 0 ? k32[0] : (j == 1 ? k32[1] : (j == 2 ? k32[2] : k32[3]));
    u64 h;
    const Probe first = first_probe<HASHED>(A, (u64)k, &h);
    const u64 res = relabel_one<HASHED>(A, (u64)k, h, first, i0 + j);
#pragma unroll
    for (int t = 0; t < 4; t++)
      if (k32[t] == k) {
        r[t] = res;
        todo &= ~(1u << t);
      }
  }
}

template <int W, int WO, bool WIDE>
__global__ void __launch_bounds__(STREAM_THREADS)
k_relabel(const uint4* src, void* dst_wide, void* dst_narrow, size_t n, RelabelArgs A, int run) {
  constexpr int V = 16 / W;
  constexpr bool HASHED = W >= 4;  // 8/16-bit dtypes use direct tables
  const size_t nvec = n / V;
  if (A.guard && __ldcg(&A.meta_ro[FRB_META_OVERFLOW]) != 0) return;  // uniform: the flag is stable while this kernel runs
  auto row_fn = [&](int, const uint4& cur, bool valid, size_t i0, u64, bool) {
    if (!valid) return;
    u64 r[V];
    if constexpr (W == 4) {
      relabel_vector_distinct4<HASHED>(A, cur, i0, r);
    } else if constexpr (W <= 2) {
      relabel_vector_cached<W, HASHED>(A, cur, i0, r);
    } else {
      u64 key[V], h[V];
      Probe first[V];
#pragma unroll
      for (int j = 0; j < V; j++) {
        key[j] = vec_get<W>(cur, j);
        first[j] = first_probe<HASHED>(A, key[j], &h[j]);
      }
#pragma unroll
      for (int j = 0; j < V; j++) r[j] = relabel_one<HASHED>(A, key[j], h[j], first[j], i0 + j);
    }
    if (WIDE) store_results<V, W>(dst_wide, i0, r);
    if (WO > 0) store_results<V, (WO > 0 ? WO : 1)>(dst_narrow, i0, r);
  };
  const TileMap64 map(nvec, (size_t)run);
  stream_rows<W, false, false, 0>(src, nvec, map, row_fn, [](const uint4*, size_t) {});

  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(src, i);
      u64 h;
      const Probe first = first_probe<HASHED>(A, key, &h);
      const u64 r = relabel_one<HASHED>(A, key, h, first, i);
      if (WIDE) st_elem<W>(dst_wide, i, r);
      if (WO > 0) st_elem<(WO > 0 ? WO : 1)>(dst_narrow, i, r);
    }
  }
}

// renumber's K3 when the refit width is still on the device: the number of labels (meta[ASSIGNED])
// picks the narrow width from thresholds the host derived from fit_dtype, so that K1 -> K2 -> K3 can
// be queued back to back without a host round trip.  wide_always: in_place semantics (ids also
// written at the input width); otherwise the wide store happens only when nothing narrower fits.
template <int W>
__global__ void __launch_bounds__(STREAM_THREADS)
k_relabel_auto(const uint4* src, void* dst_wide, void* dst_narrow, size_t n, RelabelArgs A, u64 lim1, u64 lim2, u64 lim4,
               int wide_always, int run) {
  constexpr int V = 16 / W;
  constexpr bool HASHED = W >= 4;
  const size_t nvec = n / V;
  if (__ldcg(&A.meta_ro[FRB_META_OVERFLOW]) != 0) return;  // the build overflowed: the host grows the table and starts over
  const u64 nl = __ldcg(&A.meta_ro[FRB_META_ASSIGNED]);
  int wo = nl < lim1 ? 1 : (nl < lim2 ? 2 : (nl < lim4 ? 4 : 0));
  if (wo >= W) wo = 0;
  const bool wide = wide_always || wo == 0;
  auto row_fn = [&](int, const uint4& cur, bool valid, size_t i0, u64, bool) {
    if (!valid) return;
    u64 r[V];
    if constexpr (W == 4) {
      relabel_vector_distinct4<HASHED>(A, cur, i0, r);
    } else if constexpr (W <= 2) {
      relabel_vector_cached<W, HASHED>(A, cur, i0, r);
    } else {
      u64 key[V], h[V];
      Probe first[V];
#pragma unroll
      for (int j = 0; j < V; j++) {
        key[j] = vec_get<W>(cur, j);
        first[j] = first_probe<HASHED>(A, key[j], &h[j]);
      }
#pragma unroll
      for (int j = 0; j < V; j++) r[j] = relabel_one<HASHED>(A, key[j], h[j], first[j], i0 + j);
    }
    if (wide) store_results<V, W>(dst_wide, i0, r);
    if (W > 1 && wo == 1) store_results<V, 1>(dst_narrow, i0, r);
    if (W > 2 && wo == 2) store_results<V, (W > 2 ? 2 : 1)>(dst_narrow, i0, r);
    if (W > 4 && wo == 4) store_results<V, (W > 4 ? 4 : 1)>(dst_narrow, i0, r);
  };
  const TileMap64 map(nvec, (size_t)run);
  stream_rows<W, false, false, 0>(src, nvec, map, row_fn, [](const uint4*, size_t) {});

  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(src, i);
      u64 h;
      const Probe first = first_probe<HASHED>(A, key, &h);
      const u64 r = relabel_one<HASHED>(A, key, h, first, i);
      if (wide) st_elem<W>(dst_wide, i, r);
      if (wo) st_elem_rt(dst_narrow, wo, i, r);
    }
  }
}

// Batched variant for tables that do not fit L1 (remap / mask / mask_except with 10^5+ labels): the
// home-pair loads of NB rows of the warp-tile (NB * V per lane) are issued before any of them is
// consumed, so their L2 / HBM round trips overlap instead of queueing row after row.  In place or
// out of place at the input width only (what the remap family needs).
template <int W, int NB>
__global__ void __launch_bounds__(STREAM_THREADS)
k_relabel_batched(const uint4* src, void* dst_wide, size_t n, RelabelArgs A, int run, PlaneWalk pw) {
  constexpr int V = 16 / W;
  constexpr int R = STREAM_ROWS;
  static_assert(W == 8 && R % NB == 0, "64-bit labels, hashed tables only");
  if (__ldcg(&A.meta_ro[FRB_META_AUX1]) != 0) return;  // the sampler chose k_relabel_runs (uniform: the word is stable)
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31;
  if (A.guard && __ldcg(&A.meta_ro[FRB_META_OVERFLOW]) != 0) return;
  auto tile_end = [&](const uint4* stage, size_t vb) {
#pragma unroll
    for (int u0 = 0; u0 < R; u0 += NB) {
      u64 key[NB][V], h[NB][V];
      Probe first[NB][V];
      bool valid[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int u = u0 + q;
        valid[q] = vb + (u * 32 + lane) < nvec;
        const uint4 cur = valid[q] ? stage[(u * 32 + lane)] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < V; j++) {
          key[q][j] = vec_get<W>(cur, j);
          h[q][j] = home_slot(key[q][j], A.mask);
          const SlotPair sp = ld_pair_ro_batch(A.slots + h[q][j]);
          first[q][j].a = sp.a;
          first[q][j].b = sp.b;
        }
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        if (!valid[q]) continue;
        const size_t i0 = (vb + ((u0 + q) * 32 + lane)) * V;
        u64 r[V];
#pragma unroll
        for (int j = 0; j < V; j++) r[j] = relabel_one<true>(A, key[q][j], h[q][j], first[q][j], i0 + j);
        store_results<V, W>(dst_wide, i0, r);
      }
    }
    // (tried: prefetch.global.L2 of the table sectors the NEXT tile will want while it waits in the ring: 3-4x SLOWER,
    // remap 3.07 -> 12.4 ms -- two prefetch instructions per voxel swamp the LSU.  Not used.)
  };
  // pw: a plane walk (frb_common.cuh) is possible here but measured SLOWER than the linear walk (remap through 10^6
  // labels 3.49 vs 3.07 ms, mask_except with 5 M labels 5.04 vs 4.18 ms at 1024^3): the hint is accepted and ignored
  // unless the experiment flag is set
#ifdef FRB_K3_PLANE_WALK
  const TileMap map = pw.band ? TileMap(pw) : TileMap(nvec, (size_t)run);
#else
  const TileMap map(nvec, (size_t)run);
#endif
  stream_rows<W, false, false, 0>(src, nvec, map, [](int, const uint4&, bool, size_t, u64, bool) {}, tile_end);

  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(src, i);
      u64 h;
      const Probe first = first_probe<true>(A, key, &h);
      st_elem<W>(dst_wide, i, relabel_one<true>(A, key, h, first, i));
    }
  }
}

// Tried in round 2 and removed (profiles/r02b_k3_variants.jsonl): "probe run heads only" -- only the first voxel of a run
// inside a warp row loads its home pair, everyone else takes the head's result through one 64-bit shuffle.  The
// hypothesis came from ncu (l1tex__data_pipe_lsu_wavefronts at 70 %: a divergent LDG.256 with all 32 lanes active costs
// ~20 wavefronts however many lanes share a sector).  Bit-identical results, but SLOWER: remap 3.05 -> 3.92 ms,
// mask_except 4.14 -> 4.80 ms (two rows per batch; 3.87 / 5.53 ms with one).  The extra ~20 instructions per row (two
// shuffles up, ballot, clz, two shuffles, selects, divergent compare paths) cost more than the wavefronts saved: the
// kernel is bound by issue slots and dependent latency at 24 warps per SM, not by the L1 data pipe.  Forcing 4 CTAs
// per SM (56 registers, 108 bytes of spills) was slower still (3.86 / 5.90 ms), and so was a 4-deep TMA ring (4.20 /
// 5.37 ms: more streaming reads in flight delay the table probes).
// One lookup per RUN, and the runs of a whole warp-tile looked up together (round 2).  k_relabel_batched keeps 2 rows
// (128 voxels) of a warp in flight per table round trip, and registers (16 per row for the home pairs) cap it there; at
// 10^6..10^7 labels the pass is bound by that latency (ncu: no unit above 70 %, long-scoreboard stalls, 54 % issue).
// Here the 256 voxels of a warp-tile are first cut into runs of equal labels (a lane compares with the element in front
// of it in memory, two ballots per row number the runs), the run heads' labels go to a list in shared memory, and then
// LANE i LOOKS UP RUN i: every lane of a load instruction asks for a different label, and the ~50 runs of a tile
// (4.8-voxel runs) are in flight together with two loads per lane -- the whole tile per round trip, at 16 registers.
// Results go back through shared memory (run number -> result).  Exact for every policy: a result depends on the label
// alone, and the first voxel missing from the table (error policy: atomicMin of its index) is the head of its run.
static constexpr int RUNS_MAX = STREAM_ROWS * 64;  // runs in a warp-tile of 64-bit labels (every voxel its own run)
template <int W>
__global__ void __launch_bounds__(STREAM_THREADS)
k_relabel_runs(const uint4* src, void* dst_wide, size_t n, RelabelArgs A, int run) {
  if (__ldcg(&A.meta_ro[FRB_META_AUX1]) == 0) return;  // the sampler chose k_relabel_batched (uniform: the word is stable)
  constexpr int V = 16 / W;
  constexpr int R = STREAM_ROWS;
  static_assert(W == 8 && V == 2 && R == 4, "64-bit labels, hashed tables only");
  __shared__ u64 s_key[STREAM_WARPS][RUNS_MAX];            // label of run i; overwritten by its result
  __shared__ unsigned char s_pos[STREAM_WARPS][RUNS_MAX];  // element position (0..255) of the run's head in the warp-tile
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int below = (1u << lane) - 1u;
  if (A.guard && __ldcg(&A.meta_ro[FRB_META_OVERFLOW]) != 0) return;
  auto tile_end = [&](const uint4* stage, size_t vb) {
    u64* const keys = s_key[warp];
    unsigned char* const pos = s_pos[warp];
    // pass 1: number the runs.  first[u]: run number of the lane's element 0 in row u; split bit u: element 1 starts a run
    unsigned int first[R];
    unsigned int split = 0, total = 0;
#pragma unroll
    for (int u = 0; u < R; u++) {
      const bool valid = vb + (u * 32 + lane) < nvec;
      const uint4 cur = valid ? stage[u * 32 + lane] : make_uint4(0, 0, 0, 0);
      const u64 k0 = vec_get<W>(cur, 0), k1 = vec_get<W>(cur, 1);
      // the element in front of mine in memory: the staged bytes hold it for every lane but the tile's first
      const u64 before = (u == 0 && lane == 0) ? 0 : reinterpret_cast<const u64*>(stage + (u * 32 + lane))[-1];
      const bool h0 = valid && ((u == 0 && lane == 0) || k0 != before);
      const bool h1 = valid && k1 != k0;
      const unsigned int b0 = __ballot_sync(FRB_FULL_MASK, h0), b1 = __ballot_sync(FRB_FULL_MASK, h1);
      // runs opened in this row in front of my element 0 (heads are ordered: lane-major, element 0 before element 1)
      const unsigned int mine = total + __popc(b0 & below) + __popc(b1 & below) + (h0 ? 1u : 0u) - 1u;
      first[u] = mine;
      if (h0) { keys[mine] = k0; pos[mine] = (unsigned char)(u * 64 + lane * 2); }
      if (h1) { keys[mine + 1] = k1; pos[mine + 1] = (unsigned char)(u * 64 + lane * 2 + 1); split |= 1u << u; }
      total += __popc(b0) + __popc(b1);
    }
    __syncwarp();
    // pass 2: lane i looks up runs i, i + 32 (, i + 64 ...): all of a tile's lookups in flight together
    for (unsigned int base = 0; base < total; base += 64) {
      const unsigned int i0 = base + lane, i1 = base + 32 + lane;
      const bool a0 = i0 < total, a1 = i1 < total;
      const u64 k0 = a0 ? keys[i0] : 0, k1 = a1 ? keys[i1] : 0;
      const u64 h0 = home_slot(k0, A.mask), h1 = home_slot(k1, A.mask);
      Probe p0, p1;
      p0.a.key = p0.b.key = p1.a.key = p1.b.key = FRB_EMPTY;
      p0.a.val = p0.b.val = p1.a.val = p1.b.val = 0;
      if (a0) { const SlotPair sp = ld_pair_ro_batch(A.slots + h0); p0.a = sp.a; p0.b = sp.b; }
      if (a1) { const SlotPair sp = ld_pair_ro_batch(A.slots + h1); p1.a = sp.a; p1.b = sp.b; }
      if (a0) keys[i0] = relabel_one<true>(A, k0, h0, p0, vb * V + pos[i0]);
      if (a1) keys[i1] = relabel_one<true>(A, k1, h1, p1, vb * V + pos[i1]);
    }
    __syncwarp();
    // pass 3: every voxel takes the result of its run
#pragma unroll
    for (int u = 0; u < R; u++) {
      if (vb + (u * 32 + lane) < nvec) {
        u64 r[V];
        r[0] = keys[first[u]];
        r[1] = (split >> u) & 1u ? keys[first[u] + 1] : r[0];
        store_results<V, W>(dst_wide, (vb + (u * 32 + lane)) * V, r);
      }
    }
    __syncwarp();  // the lists are reused by the warp's next tile
  };
  const TileMap64 map(nvec, (size_t)run);
  stream_rows<W, false, false, 0>(src, nvec, map, [](int, const uint4&, bool, size_t, u64, bool) {}, tile_end);

  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(src, i);
      u64 h;
      const Probe first = first_probe<true>(A, key, &h);
      st_elem<W>(dst_wide, i, relabel_one<true>(A, key, h, first, i));
    }
  }
}

// Which of the two?  k_relabel_runs wins when runs are long (remap of 9.5-voxel runs through 10^6 labels: 3.25 -> 2.92
// ms) and loses when they are short (mask_except on 4.8-voxel runs, 10^7 labels: 4.13 -> 5.39 ms: its lookups no longer
// coalesce inside a load instruction, and the table's L2 misses go up from 62 M to 101 M sectors).  The host cannot
// know, so a sampler looks at 64 blocks of 4 KiB spread over the volume, counts the run heads, and leaves the choice in
// the table's meta[AUX1] (1: runs), where both kernels -- launched back to back -- read it; the one not chosen returns
// at once (~3 us).  meta words 8..10 are the sampler's scratch (left zero again).
static constexpr int SAMPLE_BLOCKS = 64;
#ifndef FRB_RUNS_MAX_HEADS
#define FRB_RUNS_MAX_HEADS 40  // run heads per 256 voxels up to which k_relabel_runs is chosen
#endif
__global__ void __launch_bounds__(256) k_run_density_sample(const uint4* __restrict__ src, size_t nvec, u64* meta, int force) {
  const size_t per = nvec / SAMPLE_BLOCKS;
  size_t v = (size_t)blockIdx.x * per + threadIdx.x;
  unsigned int heads = 0, seen = 0;
  if (per >= 256 && v < nvec) {
    const uint4 cur = ld_stream16(src + v);
    const u64 k0 = vec_get<8>(cur, 0), k1 = vec_get<8>(cur, 1);
    heads = (k1 != k0) ? 1u : 0u;
    if (v > 0) heads += reinterpret_cast<const u64*>(src + v)[-1] != k0 ? 1u : 0u;
    seen = 2;
  }
  heads = __reduce_add_sync(FRB_FULL_MASK, heads);
  seen = __reduce_add_sync(FRB_FULL_MASK, seen);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&meta[8], (u64)heads);
    atomicAdd(&meta[9], (u64)seen);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&meta[10], 1ull) == (u64)gridDim.x - 1) {  // last CTA: decide, then leave the scratch words zero
      __threadfence();
      const u64 h = atomicExch(&meta[8], 0ull), s = atomicExch(&meta[9], 0ull);
      atomicExch(&meta[10], 0ull);
      u64 choose = (s > 0 && h * 256 <= (u64)FRB_RUNS_MAX_HEADS * s) ? 1ull : 0ull;
      if (force >= 0) choose = (u64)force;
      meta[FRB_META_AUX1] = choose;
    }
  }
}

static int launch_relabel_runs(const void* src, void* dst_wide, size_t n, const RelabelArgs& A, cudaStream_t st) {
  const size_t tiles = stream_tiles(n, 8);
  const int grid = stream_grid(k_relabel_runs<8>, tiles);
  k_relabel_runs<8><<<grid, STREAM_THREADS, STREAM_SMEM_BYTES, st>>>((const uint4*)src, dst_wide, n, A, fair_run(grid, tiles));
  return check_launch("k_relabel_runs");
}

template <int W, int NB>
static int launch_relabel_batched(const void* src, void* dst_wide, size_t n, const RelabelArgs& A, size_t plane_bytes, cudaStream_t st) {
  const size_t tiles = stream_tiles(n, W);
  const int grid = stream_grid(k_relabel_batched<W, NB>, tiles);
  const PlaneWalk pw = plane_walk_for(n * (size_t)W, (n % (16 / W)) ? 0 : plane_bytes, grid);
  k_relabel_batched<W, NB><<<grid, STREAM_THREADS, STREAM_SMEM_BYTES, st>>>((const uint4*)src, dst_wide, n, A, fair_run(grid, tiles), pw);
  return check_launch("k_relabel_batched");
}

template <int W, int WO, bool WIDE>
static int launch_relabel(const void* src, void* dst_wide, void* dst_narrow, size_t n, const RelabelArgs& A, cudaStream_t st) {
  const size_t tiles = stream_tiles(n, W);
  k_relabel<W, WO, WIDE><<<stream_grid(k_relabel<W, WO, WIDE>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>((const uint4*)src, dst_wide, dst_narrow, n, A, fair_run(stream_grid(k_relabel<W, WO, WIDE>, tiles), tiles));
  return check_launch("k_relabel");
}

template <int W>
static int dispatch_relabel(const void* src, void* dst_wide, void* dst_narrow, int wo, size_t n, const RelabelArgs& A, cudaStream_t st) {
  const bool wide = dst_wide != nullptr;
  if (!dst_narrow) wo = 0;
#define FRB_CASE(WO_)                                                                                 \
  if (wo == WO_) {                                                                                    \
    if (wide) return launch_relabel<W, WO_, true>(src, dst_wide, dst_narrow, n, A, st);               \
    if (WO_ > 0) return launch_relabel<W, WO_, false>(src, dst_wide, dst_narrow, n, A, st);           \
  }
  FRB_CASE(0)
  FRB_CASE(1)
  if constexpr (W >= 2) { FRB_CASE(2) }
  if constexpr (W >= 4) { FRB_CASE(4) }
  if constexpr (W >= 8) { FRB_CASE(8) }
#undef FRB_CASE
  // the one widening combination: 8-byte results from narrower labels (positions for unique's inverse), second destination only
  if constexpr (W < 8) {
    if (wo == 8 && !wide) return launch_relabel<W, 8, false>(src, dst_wide, dst_narrow, n, A, st);
  }
  set_error("frb_relabel: unsupported width combination (narrow_width must be <= input width, or 8 without dst_wide; one destination needed)");
  return FRB_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------- element-wise passes on the streaming loop
// replace_one, remap_from_array and refit's cast are "read a vector, write a vector": the volume comes in through the
// TMA ring of stream_rows (it may be the destination too: a tile is staged in shared memory before any of it is
// written), results leave with 16-byte streaming stores.  Fn: void operator()(const uint4& cur, size_t i0, u64 (&r)[V])
template <int W, int WO, typename Fn>
__global__ void __launch_bounds__(STREAM_THREADS) k_map_stream(const uint4* src, void* dst, size_t n, Fn fn, int run) {
  constexpr int V = 16 / W;
  const size_t nvec = n / V;
  auto row_fn = [&](int, const uint4& cur, bool valid, size_t i0, u64, bool) {
    if (!valid) return;
    u64 r[V];
#pragma unroll
    for (int j = 0; j < V; j++) r[j] = fn(vec_get<W>(cur, j));
    store_results<V, WO>(dst, i0, r);
  };
  const TileMap map(nvec, (size_t)run);
  stream_rows<W, false, false, 0>(src, nvec, map, row_fn, [](const uint4*, size_t) {});
  if (blockIdx.x == 0 && threadIdx.x == 0)  // scalar tail (< V elements)
    for (size_t i = nvec * V; i < n; i++) st_elem<WO>(dst, i, fn(ld_elem<W>(src, i)));
}
template <int W, int WO, typename Fn>
static int launch_map_stream(const void* src, void* dst, size_t n, const Fn& fn, const char* what, cudaStream_t st) {
  const size_t tiles = stream_tiles(n, W);
  const int grid = stream_grid(k_map_stream<W, WO, Fn>, tiles);
  k_map_stream<W, WO, Fn><<<grid, STREAM_THREADS, STREAM_SMEM_BYTES, st>>>((const uint4*)src, dst, n, fn, fair_run(grid, tiles));
  return check_launch(what);
}

// single pair replace: remap's one-entry fast path (fastremap.pyx:721-729)
struct ReplaceOne {
  u64 before, after;
  __device__ __forceinline__ u64 operator()(u64 k) const { return k == before ? after : k; }
};
// remap_from_array (fastremap.pyx:785-789): arr[i] = vals[arr[i]] where arr[i] < len(vals)
template <int W>
struct FromArray {
  const void* vals;
  u64 nvals;
  __device__ __forceinline__ u64 operator()(u64 k) const { return k < nvals ? (u64)__ldg(((const typename UIntOf<W>::type*)vals) + k) : k; }
};
// refit's astype (fastremap.pyx:301): widen with the source's signedness, narrow by truncation
template <int WI, bool SIGNED_IN>
struct CastElem {
  __device__ __forceinline__ u64 operator()(u64 k) const { return SIGNED_IN ? sign_extend(k, WI) : k; }
};

// ---------------------------------------------------------------------- swap halves (edge lists)
// dst[i] = src[i] with its low and high halves exchanged: an (N, 2) C-order array of w-byte ints,
// read as N ints of 2w bytes, becomes "column 0 is the high half" -- sorting those ints sorts the
// rows lexicographically.  The (N,2) packing trick of _two_axis_unique (fastremap.pyx:967-982).
template <int W>
__global__ void __launch_bounds__(256) k_swap_halves(const void* src, void* dst, size_t n) {  // dst may be src (no __restrict__)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  constexpr int H = 4 * W;  // bits in a half
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 v = ld_elem<W>(src, i);
    const u64 lo = v & ((1ull << H) - 1ull);
    st_elem<W>(dst, i, (lo << H) | (v >> H));
  }
}

// ---------------------------------------------------------------------- pack pairs (inverse_component_map)
// out[i] = (hi[i] << lo_bits) | lo[i], raw bits, as uint64: the (parent, component) pair of a voxel
// as ONE label, so that the distinct pairs are a `unique` of the packed volume
// (_inverse_component_map, fastremap.pyx:623-652, collects the same pairs into Python sets).
__global__ void __launch_bounds__(256)
k_pack_pairs(const void* __restrict__ hi, int whi, const void* __restrict__ lo, int wlo, u64* __restrict__ out, size_t n, int lo_bits) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (ld_elem_rt(hi, whi, i) << lo_bits) | ld_elem_rt(lo, wlo, i);
}

// ---------------------------------------------------------------------- cast (refit astype): unaligned fallback
// (buffers that are not 16-byte aligned cannot go through the TMA ring)
template <int WI, int WO, bool SIGNED_IN>
__global__ void __launch_bounds__(256) k_cast_elems(const void* __restrict__ src, void* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 v = ld_elem<WI>(src, i);
    st_elem<WO>(dst, i, SIGNED_IN ? sign_extend(v, WI) : v);
  }
}

template <int WI, int WO>
static int launch_cast(const void* src, void* dst, size_t n, bool signed_in, cudaStream_t st) {
  if (aligned16(src) && aligned16(dst)) {
    if (signed_in) return launch_map_stream<WI, WO>(src, dst, n, CastElem<WI, true>(), "k_map_stream<cast>", st);
    return launch_map_stream<WI, WO>(src, dst, n, CastElem<WI, false>(), "k_map_stream<cast>", st);
  }
  Grid g = grid_for(n, 256 * 8, 16);
  if (signed_in) k_cast_elems<WI, WO, true><<<g.blocks, g.threads, 0, st>>>(src, dst, n);
  else k_cast_elems<WI, WO, false><<<g.blocks, g.threads, 0, st>>>(src, dst, n);
  return check_launch("k_cast_elems");
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_relabel(const void* src, void* dst_wide, void* dst_narrow, int narrow_width, size_t n, int dtype,
                const void* slots, uint64_t cap, int direct, uint64_t* meta, int policy, uint64_t fill, int zero_fixed,
                uint64_t add, int guard_overflow, int batch_rows, uint64_t plane_bytes, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || (!dst_wide && !dst_narrow) ||
      (n && (!src || !aligned16(src) || (dst_wide && !aligned16(dst_wide)) || (dst_narrow && !aligned16(dst_narrow)))) ||
      policy < 0 || policy > 2) {
    set_error("frb_relabel: bad arguments (pointers must be 16-byte aligned)");
    return FRB_ERR_BAD_ARG;
  }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  RelabelArgs A;
  A.slots = (const Slot*)slots;
  A.mask = cap - 1;
  A.meta_ro = (const u64*)meta;
  A.meta = (u64*)meta;
  A.fill = trunc_width(fill, w);
  A.direct = direct ? 1 : 0;
  A.policy = policy;
  A.zero_fixed = zero_fixed ? 1 : 0;
  A.add = add;
  A.guard = guard_overflow ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  // 64-bit labels only: with 4 labels per vector the batch is 8+ probes per lane (80+ registers) and was
  // measured 2.5-3x SLOWER than the plain kernel's one-probe-per-distinct-label loop at every table size
  if (batch_rows > 1 && !A.direct && w == 8 && dst_wide && !dst_narrow) {
    k_run_density_sample<<<SAMPLE_BLOCKS, 256, 0, st>>>((const uint4*)src, n / 2, A.meta, relabel_variant());
    int rc = check_launch("k_run_density_sample");
    if (rc) return rc;
    if ((rc = launch_relabel_runs(src, dst_wide, n, A, st))) return rc;
    return batch_rows >= 4 ? launch_relabel_batched<8, 4>(src, dst_wide, n, A, (size_t)plane_bytes, st)
                           : launch_relabel_batched<8, 2>(src, dst_wide, n, A, (size_t)plane_bytes, st);
  }
  switch (w) {
    case 1: return dispatch_relabel<1>(src, dst_wide, dst_narrow, narrow_width, n, A, st);
    case 2: return dispatch_relabel<2>(src, dst_wide, dst_narrow, narrow_width, n, A, st);
    case 4: return dispatch_relabel<4>(src, dst_wide, dst_narrow, narrow_width, n, A, st);
    default: return dispatch_relabel<8>(src, dst_wide, dst_narrow, narrow_width, n, A, st);
  }
}

int frb_relabel_auto(const void* src, void* dst_wide, void* dst_narrow, size_t n, int dtype, const void* slots, uint64_t cap,
                     int direct, uint64_t* meta, int zero_fixed, uint64_t add, int wide_always, uint64_t lim1, uint64_t lim2,
                     uint64_t lim4, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || !dst_wide || !dst_narrow ||
      (n && (!src || !aligned16(src) || !aligned16(dst_wide) || !aligned16(dst_narrow)))) {
    set_error("frb_relabel_auto: bad arguments (both destinations needed, 16-byte aligned)");
    return FRB_ERR_BAD_ARG;
  }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  RelabelArgs A;
  A.slots = (const Slot*)slots;
  A.mask = cap - 1;
  A.meta_ro = (const u64*)meta;
  A.meta = (u64*)meta;
  A.fill = 0;
  A.direct = direct ? 1 : 0;
  A.policy = FRB_MISS_ERROR;
  A.zero_fixed = zero_fixed ? 1 : 0;
  A.add = add;
  A.guard = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)src;
#define FRB_AUTO(WW) k_relabel_auto<WW><<<stream_grid(k_relabel_auto<WW>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>(p, dst_wide, dst_narrow, n, A, lim1, lim2, lim4, wide_always, fair_run(stream_grid(k_relabel_auto<WW>, tiles), tiles))
  switch (w) {
    case 1: FRB_AUTO(1); break;
    case 2: FRB_AUTO(2); break;
    case 4: FRB_AUTO(4); break;
    default: FRB_AUTO(8); break;
  }
#undef FRB_AUTO
  return check_launch("k_relabel_auto");
}

int frb_replace_one(const void* src, void* dst, size_t n, int dtype, uint64_t before, uint64_t after, void* stream) {
  if (!dtype_ok(dtype) || (n && (!src || !dst || !aligned16(src) || !aligned16(dst)))) { set_error("frb_replace_one: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  cudaStream_t st = (cudaStream_t)stream;
  ReplaceOne fn;
  fn.before = trunc_width(before, w);
  fn.after = trunc_width(after, w);
  switch (w) {
    case 1: return launch_map_stream<1, 1>(src, dst, n, fn, "k_map_stream<replace_one>", st);
    case 2: return launch_map_stream<2, 2>(src, dst, n, fn, "k_map_stream<replace_one>", st);
    case 4: return launch_map_stream<4, 4>(src, dst, n, fn, "k_map_stream<replace_one>", st);
    default: return launch_map_stream<8, 8>(src, dst, n, fn, "k_map_stream<replace_one>", st);
  }
}

int frb_remap_from_array(const void* src, void* dst, size_t n, const void* vals, size_t nvals, int dtype, void* stream) {
  if (!dtype_ok(dtype) || dtype_signed(dtype) || (n && (!src || !dst || !aligned16(src) || !aligned16(dst))) || (nvals && !vals)) {
    set_error("frb_remap_from_array: bad arguments (unsigned dtypes only)");
    return FRB_ERR_BAD_ARG;
  }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  cudaStream_t st = (cudaStream_t)stream;
#define FRB_FA(WW) { FromArray<WW> fn; fn.vals = vals; fn.nvals = nvals; return launch_map_stream<WW, WW>(src, dst, n, fn, "k_map_stream<remap_from_array>", st); }
  switch (w) {
    case 1: FRB_FA(1)
    case 2: FRB_FA(2)
    case 4: FRB_FA(4)
    default: FRB_FA(8)
  }
#undef FRB_FA
}

int frb_swap_halves(const void* src, void* dst, size_t n, int width, void* stream) {
  if ((n && (!src || !dst)) || (width != 2 && width != 4 && width != 8)) { set_error("frb_swap_halves: bad arguments (width 2, 4 or 8)"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = grid_for(n, 256 * 4, 16);
  switch (width) {
    case 2: k_swap_halves<2><<<g.blocks, g.threads, 0, st>>>(src, dst, n); break;
    case 4: k_swap_halves<4><<<g.blocks, g.threads, 0, st>>>(src, dst, n); break;
    default: k_swap_halves<8><<<g.blocks, g.threads, 0, st>>>(src, dst, n); break;
  }
  return check_launch("k_swap_halves");
}

int frb_pack_pairs(const void* hi, int hi_width, const void* lo, int lo_width, uint64_t* out, size_t n, void* stream) {
  const bool wok = (hi_width == 1 || hi_width == 2 || hi_width == 4) && (lo_width == 1 || lo_width == 2 || lo_width == 4);
  if (!wok || (n && (!hi || !lo || !out))) { set_error("frb_pack_pairs: bad arguments (widths 1, 2 or 4)"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  Grid g = grid_for(n, 256 * 4, 16);
  k_pack_pairs<<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>(hi, hi_width, lo, lo_width, (u64*)out, n, 8 * lo_width);
  return check_launch("k_pack_pairs");
}

int frb_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, void* stream) {
  if (!dtype_ok(src_dtype) || !dtype_ok(dst_dtype) || (n && (!src || !dst))) { set_error("frb_cast: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  const int wi = dtype_width(src_dtype), wo = dtype_width(dst_dtype);
  const bool sg = dtype_signed(src_dtype);
  cudaStream_t st = (cudaStream_t)stream;
#define FRB_CAST(WI_, WO_) if (wi == WI_ && wo == WO_) return launch_cast<WI_, WO_>(src, dst, n, sg, st);
  FRB_CAST(1, 1) FRB_CAST(1, 2) FRB_CAST(1, 4) FRB_CAST(1, 8)
  FRB_CAST(2, 1) FRB_CAST(2, 2) FRB_CAST(2, 4) FRB_CAST(2, 8)
  FRB_CAST(4, 1) FRB_CAST(4, 2) FRB_CAST(4, 4) FRB_CAST(4, 8)
  FRB_CAST(8, 1) FRB_CAST(8, 2) FRB_CAST(8, 4) FRB_CAST(8, 8)
#undef FRB_CAST
  set_error("frb_cast: unsupported widths");
  return FRB_ERR_BAD_ARG;
}

}  // extern "C"
