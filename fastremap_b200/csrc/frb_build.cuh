// frb_build.cuh -- the filtered "volume -> label table" pass behind K1 (renumber: atomicMin of the
// first index of every label) and K7 (component_map: atomicMax of the last run start).
//
// Offering ANY occurrence (K1) / ANY run start (K7) of a label is correct -- the table combines
// with MIN / MAX -- so the work is in offering as few as possible: segmentation volumes repeat the
// same label for tens of voxels in every direction.  Three filters sit between a voxel and the
// table, each exact:
//   1. registers: a lane remembers the 16-byte vectors of its last four rows; a warp's rows are
//      4 KiB apart (stream_vec), so for the usual row lengths one of them is the voxels DIRECTLY above
//      (rows_up / pick_slot; below, when tiles are walked backwards for MAX -- "remembered" always
//      means "already offered something at least as good").  K1: an element equal to the remembered one, or to its predecessor in
//      memory, is dominated.  K7 (HEADS): if the vector AND the element in front of it are the same
//      as remembered, every run start in it has a twin one tile later.  Testing a whole vector costs
//      7 integer instructions for all V elements.
//   2. a per-warp shared-memory set of labels the warp has already offered (phase-separated: all
//      reads of a tile before all writes, so a hit always refers to an earlier tile).
//   3. the 16-byte home slot itself: read first (L2-coherent), issued one tile ahead of its use so
//      the round trip overlaps the streaming; "label present with a better value" ends there.
// Only what survives all three reaches atomicCAS / atomicMin / atomicMax.
// K1 keeps one such deferred offer per lane (labels are few where it is timed); K7 queues the survivors per warp and
// settles them 32 at a time straight from the pre-read slots, and lists the slots it fills (QUEUED, below).
#pragma once
#include "frb_common.cuh"

namespace frb {

static constexpr int BUILD_CACHE = 256;  // entries in the per-warp set of labels already offered
static constexpr int BUILD_QUEUE = 64;   // K7: offers a warp has queued for its next batch of 32 (see build_pass)

// out-of-line find-or-insert + combine (the rare path of the build kernels; keeps their hot loops small)
template <int OP>
__device__ __noinline__ bool table_update_slow(Slot* slots, u64 mask, int direct, u64* meta, u64 key, u64 v, unsigned int* inserted) {
  return table_update<OP>(slots, mask, direct, meta, key, v, inserted);
}
// K7's list of filled slots (so that its gather need not scan the table).  Layout, in 32-bit words:
//   [0] R = regions (consumer warps of the build grid)   [1] region capacity   [2], [3] unused
//   [4 .. 4 + R)   entries per region (written by each warp when it is done)
//   then R regions of `region capacity` slot indices, one per consumer warp: appended to through a counter in shared
//   memory (no global atomic, no round trip), in the order the warp walks the volume -- so the gather's reads of
//   parent[] stay inside a few pages.
// meta[AUX1] receives the number of entries listed; a region that overflowed adds 2^40 (the host then scans instead).
static constexpr u64 LIST_OVERFLOW = 1ull << 40;
__host__ __device__ __forceinline__ u64 slot_list_region_cap(u64 list_words, u64 regions) {
  return list_words > 4 + regions ? (list_words - 4 - regions) / regions : 0;
}
// append slot index `at` to the calling warp's region (any subset of its lanes may call this, in any order)
__device__ __forceinline__ void slot_list_append(unsigned int* list, u64 region_cap, unsigned int* warp_count, u64 at) {
  const unsigned int pos = atomicAdd(warp_count, 1u);
  if (pos < region_cap) {
    const u64 regions = (u64)gridDim.x * STREAM_WARPS;
    list[4 + regions + ((u64)blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5)) * region_cap + pos] = (unsigned int)at;
  }
}
template <int OP>
__device__ __noinline__ bool table_update_slow_listed(Slot* slots, u64 mask, int direct, u64* meta, u64 key, u64 v, unsigned int* inserted,
                                                      unsigned int* list, u64 region_cap, unsigned int* warp_count) {
  u64 at = FRB_EMPTY;
  const bool ok = table_update<OP>(slots, mask, direct, meta, key, v, inserted, &at);
  if (at != FRB_EMPTY) slot_list_append(list, region_cap, warp_count, at);
  return ok;
}

// value offered for element i is bias + i.  HEADS: only run starts (a[i] != a[i-1]) are offered.
// IL / DU: how a warp's rows are laid out in the tile and which of the lane's row slots holds the row
// directly above / below (row_geom; compile-time: the selection must not cost instructions, this loop
// is issue-bound)
// PZ: walk the volume as a balanced plane walk (pwz, frb_common.cuh) instead of `run` / `pw`
template <int W, int OP, bool REVERSE, bool HEADS, int IL, int DU, bool VAR_TILE, bool PZ = false>
__device__ __forceinline__ void build_pass(const uint4* __restrict__ a, size_t n, u64 bias, int skip_zero,
                                           Slot* __restrict__ slots, u64 mask, int direct, u64* __restrict__ meta, int run,
                                           int tile_vecs, PlaneWalk pw, unsigned int* slot_list = nullptr, u64 region_cap = 0,
                                           PlaneWalkZ pwz = PlaneWalkZ{0, 0, 0}) {
  constexpr int V = 16 / W;
  typedef typename UIntOf<W>::type T;
  // QUEUED (K7, round 2): what survives the filters is not offered by the lane that found it but appended to a per-warp
  // queue; whenever 32 offers wait, every free lane takes ONE, reads its home slot now and settles it at the next tile
  // end straight from that copy (present: one fire-and-forget reduction; empty: one CAS; another label: the lane reads
  // the second slot of the home pair and tries again a tile later -- at load 0.3 a quarter of all first probes).  Only
  // what is still unsettled after the pair takes the out-of-line probing path.  The old scheme below -- one
  // deferred offer per lane, the rest synchronously -- left the warp waiting for two dependent round trips (re-read +
  // CAS) in most tiles at 10^7 labels, because with 0.35 offers per lane and tile SOME lane nearly always had two
  // (ncu, profiles/r02_ncu_full_component_map.csv: no unit above 35 % of its peak, 2 us per tile and CTA).
  constexpr bool QUEUED = HEADS;
  __shared__ u64 s_seen[8][BUILD_CACHE];
  __shared__ Slot s_queue[QUEUED ? 8 : 1][QUEUED ? BUILD_QUEUE : 1];
  __shared__ unsigned int s_listed[8];  // QUEUED: entries this warp has appended to its region of the slot list
  unsigned int qn = 0;  // entries in this warp's queue (the same value in every lane)
  constexpr u64 SECOND = 1ull << 62;  // QUEUED: flag in pval -- the pending offer looks at the second slot of its home pair
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int inserted = 0;  // labels this lane has inserted ...
  unsigned int reported = 0;  // ... of which this many have been added to meta[COUNT] already
  unsigned int tiles_since_report = 0;
  // a hashed table that is more than half full is too small: stop early (the host grows it and starts
  // over) instead of running into the probe limit at ~100 % load, where every offer walks 512 slots
  const u64 count_limit = direct ? ~0ull : (mask + 1) / 2 + 1;
  if (warp < STREAM_WARPS) {
    for (int i = lane; i < BUILD_CACHE; i += 32) s_seen[warp][i] = FRB_EMPTY;
    if (lane == 0) s_listed[warp] = 0;
  }
  __syncwarp();

  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar tail (< V elements), unfiltered
    for (size_t i = nvec * V; i < n; i++) {
      const u64 key = ld_elem<W>(a, i);
      const bool head = !HEADS || i == 0 || key != ld_elem<W>(a, i - 1);
      if (head && !(skip_zero && key == 0)) table_update_slow<OP>(slots, mask, direct, meta, key, bias + (u64)i, &inserted);
    }
  }

  static_assert(STREAM_ROWS == 4, "pick_slot assumes 4 row slots per lane");
  uint4 last[4];
  u64 last_prev[4];  // HEADS only
  unsigned int have = 0;       // bit u: row slot u of this lane holds a vector it has really seen
  u64 cand = 0;                // bit (u * V + j): element j of row u needs a closer look
  // A full table makes the rest of the pass pointless (the host grows the table and retries): a lane
  // stops offering once one of its own offers failed, or if the table had already overflowed when
  // this launch started (chunk pipelines queue launches ahead of the host's overflow check).
  bool aborted = *((volatile u64*)&meta[FRB_META_OVERFLOW]) != 0;
  bool pend = false;           // one deferred offer per lane, resolved a tile later
  u64 pkey = 0, pval = 0;
  Slot pslot;
  pslot.key = 0;
  pslot.val = 0;

  // slot_list (K7, optional): every slot this pass fills is listed (layout above)
  if (QUEUED && slot_list && blockIdx.x == 0 && threadIdx.x == 0) {
    slot_list[0] = gridDim.x * STREAM_WARPS;
    slot_list[1] = (unsigned int)region_cap;
  }
  auto offer = [&](u64 key, u64 val) {
    if (QUEUED && slot_list) {
      if (!table_update_slow_listed<OP>(slots, mask, direct, meta, key, val, &inserted, slot_list, region_cap, &s_listed[warp])) aborted = true;
    } else if (!table_update_slow<OP>(slots, mask, direct, meta, key, val, &inserted)) {
      aborted = true;
    }
  };

  auto row_fn = [&](int u, const uint4& cur, bool valid, size_t, u64 prev, bool has_prev) {
    if (!valid) return;
    const uint4 l = last[ref_slot_ct<REVERSE, DU>(u)];  // the row above (below) this one, if the row length is known
    const bool have_last = (have >> ref_slot_ct<REVERSE, DU>(u)) & 1u;
    last[u] = cur;
    have |= 1u << u;
    const uint32_t d = (cur.x ^ l.x) | (cur.y ^ l.y) | (cur.z ^ l.z) | (cur.w ^ l.w);
    if (HEADS) {
      const u64 lp = last_prev[ref_slot_ct<REVERSE, DU>(u)];
      last_prev[u] = prev;
      if (d != 0 || !have_last || !has_prev || prev != lp) {
        // not a copy of the remembered row: every run start in it is a candidate
#pragma unroll
        for (int j = 0; j < V; j++) {
          const u64 key = vec_get<W>(cur, j);
          if ((!has_prev || key != prev) && !(skip_zero && key == 0)) cand |= 1ull << (u * V + j);
          prev = key;
          has_prev = true;
        }
      }
    } else if (d != 0 || !have_last) {
      // rare path (a label boundary one row up, or the warp's first tile): keep the elements that
      // differ from the remembered vector AND from their predecessor in memory -- an element equal
      // to its predecessor is dominated by whatever happened to that predecessor
#pragma unroll
      for (int j = 0; j < V; j++) {
        const u64 key = vec_get<W>(cur, j);
        const bool differs_up = !have_last || key != vec_get<W>(l, j);
        const bool differs_prev = !has_prev || key != prev;
        if (differs_up && differs_prev && !(skip_zero && key == 0)) cand |= 1ull << (u * V + j);
        prev = key;
        has_prev = true;
      }
    }
  };

  auto resolve_pending = [&]() {
    if (pend) {
      if (QUEUED) {
        const bool second = (pval & SECOND) != 0;
        const u64 val = pval & ~SECOND;
        const u64 h = direct ? pkey : (home_slot(pkey, mask) + (second ? 1 : 0));
        Slot* s = slots + h;
        pend = false;
        if (pslot.key == pkey) {
          if (!op_settled<OP>(pslot.val, val)) op_apply<OP>(&s->val, val);
        } else if (pslot.key == FRB_EMPTY) {
          const u64 old = atom_cas_u64(&s->key, FRB_EMPTY, pkey);
          if (old == FRB_EMPTY) {
            inserted++;
            if (slot_list) slot_list_append(slot_list, region_cap, &s_listed[warp], h);
          }
          if (old == FRB_EMPTY || old == pkey) op_apply<OP>(&s->val, val);
          else offer(pkey, val);
        } else if (!second && !direct) {
          pval |= SECOND;  // another label lives in the home slot: look at the other slot of the pair, settle a tile later
          pslot = ld_slot_cg(s + 1);
          pend = true;
        } else {
          offer(pkey, val);
        }
      } else {
        if (!(pslot.key == pkey && op_settled<OP>(pslot.val, pval))) offer(pkey, pval);
        pend = false;
      }
    }
  };
  // QUEUED: every free lane takes one queued offer (from the tail) and reads its home slot; settled by resolve_pending
  auto take_batch = [&]() {
    const unsigned int free_lanes = __ballot_sync(FRB_FULL_MASK, !pend);
    const unsigned int nfree = __popc(free_lanes);
    const unsigned int take = qn < nfree ? qn : nfree;
    const unsigned int mine = __popc(free_lanes & ((1u << lane) - 1u));
    if (!pend && mine < take) {
      const Slot e = s_queue[warp][qn - 1u - mine];
      pend = true;
      pkey = e.key;
      pval = e.val;
      pslot = ld_slot_cg(slots + (direct ? pkey : home_slot(pkey, mask)));
    }
    qn -= take;
    __syncwarp();
  };

  auto tile_end = [&](const uint4* stage, size_t vb) {
    resolve_pending();
    if (aborted) cand = 0;
    // phase 1: which candidates has this warp not offered before?
    u64 todo = 0;
    for (u64 m = cand; m; m &= m - 1) {
      const unsigned int b = (unsigned int)(__ffsll((long long)m) - 1);
      const unsigned int u = b / V, j = b % V;
      const u64 key = (u64)reinterpret_cast<const T*>(stage + stream_vec<IL>(u, warp, lane))[j];
      const bool hit = key != FRB_EMPTY && s_seen[warp][(hash_key(key) >> 48) & (BUILD_CACHE - 1)] == key;
      if (!hit) todo |= 1ull << b;
    }
    cand = 0;
    __syncwarp();
    // phase 2: remember them and queue / send the offers
    if constexpr (QUEUED) {
      u64 t = todo;
      while (true) {
        const bool has = t != 0;
        if (!__any_sync(FRB_FULL_MASK, has)) break;
        u64 key = 0, val = 0;
        if (has) {
          const unsigned int b = (unsigned int)(__ffsll((long long)t) - 1);
          t &= t - 1;
          const unsigned int u = b / V, j = b % V;
          key = (u64)reinterpret_cast<const T*>(stage + stream_vec<IL>(u, warp, lane))[j];
          val = bias + (u64)((vb + stream_vec<IL>(u, warp, lane)) * V + j);
          s_seen[warp][(hash_key(key) >> 48) & (BUILD_CACHE - 1)] = key;
        }
        const bool queueable = has && (direct || key != FRB_EMPTY);  // the out-of-band all-ones label goes straight to the table
        const unsigned int bal = __ballot_sync(FRB_FULL_MASK, queueable);
        const unsigned int pos = qn + __popc(bal & ((1u << lane) - 1u));
        if (queueable && pos < (unsigned int)BUILD_QUEUE) {
          Slot e;
          e.key = key;
          e.val = val;
          s_queue[warp][pos] = e;
        } else if (has) {
          offer(key, val);  // queue full (noise volumes): straight to the table
        }
        qn += __popc(bal);
        if (qn > (unsigned int)BUILD_QUEUE) qn = BUILD_QUEUE;
      }
      __syncwarp();
      if (qn >= 32u) take_batch();
    } else {
    for (u64 m = todo; m; m &= m - 1) {
      const unsigned int b = (unsigned int)(__ffsll((long long)m) - 1);
      const unsigned int u = b / V, j = b % V;
      const u64 key = (u64)reinterpret_cast<const T*>(stage + stream_vec<IL>(u, warp, lane))[j];
      const u64 h = hash_key(key);
      const u64 val = bias + (u64)((vb + stream_vec<IL>(u, warp, lane)) * V + j);
      s_seen[warp][(h >> 48) & (BUILD_CACHE - 1)] = key;
      if (!pend && key != FRB_EMPTY) {
        pend = true;
        pkey = key;
        pval = val;
        pslot = ld_slot_cg(slots + (direct ? key : (h & mask & ~1ull)));  // consumed at the next tile_end
      } else {
        offer(key, val);
      }
    }
    }
    __syncwarp();
    // every 16th tile the warp adds what it has inserted to the global count (one atomic per warp) and
    // learns from the returned value whether the table has passed half full
    if (++tiles_since_report == 16) {
      tiles_since_report = 0;
      const unsigned int delta = __reduce_add_sync(FRB_FULL_MASK, inserted - reported);
      reported = inserted;
      if (delta) {
        u64 before = 0;
        if (lane == 0) before = atomicAdd(&meta[FRB_META_COUNT], (u64)delta);
        before = __shfl_sync(FRB_FULL_MASK, before, 0);
        if (before + delta > count_limit) {
          if (lane == 0) atomicExch(&meta[FRB_META_OVERFLOW], 1ull);
          aborted = true;
        }
      }
    }
  };

  // plane walk (frb_common.cuh): a CTA follows one band of rows through a stack of planes, so the labels of the band
  // are still in the warps' sets of "already offered" when they come round again -- with tables too big for L2 an
  // offer is a random HBM access, and those, not the streaming, are what bounds the pass at 10^7 labels
  if constexpr (PZ) {
    stream_rows_z<W, REVERSE, true, IL>(a, nvec, TileMapZ(pwz), row_fn, tile_end);
  } else {
    const TileMap map = (!VAR_TILE && pw.band) ? TileMap(pw) : TileMap(nvec, (size_t)run, VAR_TILE ? (size_t)tile_vecs : (size_t)STREAM_TILE_VECS);
    stream_rows<W, REVERSE, true, IL, VAR_TILE>(a, nvec, map, row_fn, tile_end, [](size_t) {});
  }
  resolve_pending();
  if (QUEUED && warp < STREAM_WARPS) {  // what is still queued or waiting for its second probe (the producer warp has nothing)
    while (qn || __any_sync(FRB_FULL_MASK, pend)) {
      take_batch();
      resolve_pending();
    }
    __syncwarp();
    if (slot_list && lane == 0) {
      const unsigned int ln = s_listed[warp];
      slot_list[4 + (u64)blockIdx.x * STREAM_WARPS + warp] = (u64)ln <= region_cap ? ln : 0u;
      if (ln) atomicAdd(&meta[FRB_META_AUX1], (u64)ln <= region_cap ? (u64)ln : LIST_OVERFLOW);
    }
  }
  block_add_to(&meta[FRB_META_COUNT], inserted - reported);
}

}  // namespace frb
