// frb_unique.cu -- K4/K5: unique(return_counts) strategies.
//
// Replaces _unique_via_array (fastremap.pyx:1061-1138), _unique_via_shifted_array (:984-1001),
// _unique_via_renumber (:1003-1022) and _unique_via_sort (:1024-1059).  All strategies produce
// the same (sorted labels, counts); the host picks by label range / cardinality:
//   dense  counts[label - min] += 1, then an ordered compaction of the non-empty bins
//   hash   (label -> count) table, then a sort of the L distinct labels
//   sort   radix sort of all N voxels, then run-length encode
// The counting kernel aggregates before it touches memory (per-lane repeat counting in registers,
// per-warp shared-memory (label, count) hash, then one global atomic per live entry): see k_hist.
#include "frb_internal.cuh"
#include "frb_build.cuh"
#include "frb_coopsort.cuh"

namespace frb {

struct DenseSink {
  static constexpr bool USE_LOG = false;  // a dense bin takes a fire-and-forget reduction as it is
  __device__ __forceinline__ bool logs() const { return false; }
  u64* counts;
  u64 min_bits;
  int w;
  __device__ __forceinline__ void operator()(u64 key, unsigned int c, unsigned int*) const {
    red_add_u64(&counts[trunc_width(key - min_bits, w)], (u64)c);
  }
  // two-step form used by the batched sends: nothing to look up for a dense bin
  __device__ __forceinline__ u64 peek(u64) const { return 0; }
  __device__ __forceinline__ void commit(u64 key, unsigned int c, u64, unsigned int* ins) const { (*this)(key, c, ins); }
};
struct HashSink {
  static constexpr bool USE_LOG = true;
  Slot* slots;
  u64 mask;
  int direct;
  u64* meta;
  // Direct tables (8- and 16-bit labels: slot = label) need no log and no look before the update: the key is
  // stored as it is (every writer stores the same value) and the count takes a reduction, both fire and forget;
  // meta[COUNT] is then made by a scan of the 2^8 / 2^16 slots after the pass (k_direct_count).
  __device__ __forceinline__ bool logs() const { return !direct; }
  __device__ __forceinline__ void direct_add(u64 key, unsigned int c) const {
    asm volatile("st.global.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(&slots[key].key)), "l"(key) : "memory");
    red_add_u64(&slots[key].val, (u64)c);
  }
  __device__ __forceinline__ void operator()(u64 key, unsigned int c, unsigned int* inserted) const {
    if (direct) direct_add(key, c);
    else table_update<FRB_OP_ADD>(slots, mask, direct, meta, key, (u64)c, inserted);
  }
  // two-step form used by the batched sends: the home-slot keys of all pairs a warp is about to send are
  // read first (one L2 round trip for the lot), then a label already sitting in its home slot costs one
  // fire-and-forget reduction; anything else takes the probing path
  __device__ __forceinline__ u64 peek(u64 key) const {
    u64 k;
    asm("ld.global.cg.u64 %0, [%1];" : "=l"(k) : "l"(&slots[direct ? key : home_slot(key, mask)].key));
    return k;
  }
  __device__ __forceinline__ void commit(u64 key, unsigned int c, u64 home_key, unsigned int* inserted) const {
    if (direct) direct_add(key, c);
    else if (home_key == key && key != FRB_EMPTY) red_add_u64(&slots[home_slot(key, mask)].val, (u64)c);
    else table_update<FRB_OP_ADD>(slots, mask, direct, meta, key, (u64)c, inserted);
  }
};

static constexpr int HIST_QUEUE = 64;   // flush requests a warp can hold (it handles 32 at a time: 31 left over + 32 new fit)
static constexpr int HIST_WMAP = 128;   // (label, partial count) entries of a warp's private map

// Counting pass.  Three levels of aggregation sit between a voxel and a global atomic:
//   lane   remembers the 16-byte vector it saw directly above (rows of a warp are 4 KiB apart, rows
//          that lie above one another share a pattern slot: hist_slot) and a repeat count; an
//          identical vector just bumps the count (7 integer instructions for V voxels).  A different
//          vector retires the old one: the lane pushes (vector, repeats) into its warp's
//   queue  in shared memory (ballot + popc, one predicated 16-byte store).  Retiring lanes are a
//          minority of every row, so the retirement work is NOT done in place, where it would run at
//          10 % lane occupancy in a loop that is issue-bound; as soon as 32 requests are queued the
//          warp handles them with all lanes busy: every lane splits one vector into (label, count)
//          runs and adds them to the
//   map    a 128-entry direct-mapped (label, partial count) table private to the warp (a warp owns a
//          fixed band of columns, so the labels it meets belong to it and to its two neighbours): a hit
//          is one shared-memory atomic; on a miss one lane per entry (match.any) replaces the resident
//          label, and only the evicted (label, count) pair goes to
//   global memory: dense bins or the label table (the sinks above), a handful of lanes at a time,
//          whose probes are in flight together.
// Nothing in the loop is CTA-wide: no barrier, no scheduled drain -- the previous version shared one map
// among the 8 warps and drained it behind a named barrier, which cost a third of all stall samples
// (ncu, profiles/r02a_*: all 8 warps parked on the L2 round trip of the drain while the TMA ring ran dry).
// pattern slot of row u: rows that lie directly above one another (q * 4 KiB apart, q = DU) share a slot
template <int DU>
__device__ __forceinline__ constexpr int hist_slot(int u) { return DU == 1 ? 0 : DU == 2 ? (u & 1) : DU == 3 ? (u == 3 ? 0 : u) : u; }

// per-warp queues and maps of k_hist (file scope, so that the out-of-line helpers below address them as
// shared memory with 32-bit offsets instead of carrying generic pointers through the hot loop)
__shared__ uint4 hs_qpat[STREAM_WARPS][HIST_QUEUE];
__shared__ unsigned int hs_qrep[STREAM_WARPS][HIST_QUEUE];
__shared__ __align__(16) u64 hs_key[STREAM_WARPS][HIST_WMAP];
__shared__ unsigned int hs_cnt[STREAM_WARPS][HIST_WMAP];
__shared__ unsigned int hs_logpos[STREAM_WARPS];  // records the warp has appended to its log region
__shared__ unsigned char hs_claim[STREAM_WARPS][HIST_WMAP / 2];  // per set: the lane that may install a new label this round
__shared__ unsigned char hs_mru[STREAM_WARPS][HIST_WMAP / 2];    // per set: the way that was used last

// Eviction log.  What a warp's map evicts does not go to the label table from inside the streaming loop:
// an update of the table is a dependent L2 round trip (read the home slot, then add), and a warp that
// waits for one holds back the CTA's whole TMA ring (measured: a third of all stall samples, whichever
// way the updates were batched).  Instead every consumer warp APPENDS (label, count) records to a region
// of its own in global memory -- plain 16-byte stores, fire and forget -- and k_hist_apply_log folds all
// regions into the table afterwards with every lane busy and four probes in flight per thread.  A warp
// whose region is full falls back to updating the table directly (and says so in flags[0]); a table that
// turns out too small then only costs a second apply pass, not a second pass over the volume.
struct HistLog {
  ulonglong2* records;    // regions * cap records; region r = records + r * cap
  unsigned int* counts;   // records written per region
  unsigned int* flags;    // [0] != 0: some warp ran out of log space and updated the table directly
  unsigned int cap;       // records per region (0: no log, always update the table directly)
};

// Add (key, c) of the lanes with `act` to the warp's map; called by all 32 lanes of a consumer warp.
// A lane that had to take over an entry gets the (label, count) pair it evicted back (key == FRB_EMPTY:
// nothing evicted); the caller sends those to global memory in batches.
struct HistEvicted {
  u64 key;
  unsigned int cnt;
};
__device__ __forceinline__ HistEvicted hist_map_add(u64 key, unsigned int c, bool act) {  // one call site, inside the out-of-line hist_process
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64* const s_key = hs_key[warp];
  unsigned int* const s_cnt = hs_cnt[warp];
  HistEvicted ev;
  ev.key = FRB_EMPTY;
  ev.cnt = 0;
  // two-way sets (an aligned pair of entries, both keys in one 16-byte shared load)
  const unsigned int e0 = (unsigned int)(hash_key(key) >> 48) & (HIST_WMAP - 2);
  while (__any_sync(FRB_FULL_MASK, act)) {
    u64 k0 = 0, k1 = 0;
    if (act) {
      const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(&s_key[e0]);
      k0 = kk.x;
      k1 = kk.y;
    }
    if (act && (k0 == key || k1 == key)) {
      const unsigned int way = k0 == key ? 0u : 1u;
      atomicAdd(&s_cnt[e0 + way], c);
      hs_mru[warp][e0 >> 1] = (unsigned char)way;
      act = false;
    }
    __syncwarp();
    // the lanes left want a set that does not hold their label: one of them per set installs it.  The winner is
    // whoever's lane number survives in the set's claim byte (match.any did the same job at ~10x the latency)
    if (act) hs_claim[warp][e0 >> 1] = (unsigned char)lane;
    __syncwarp();
    if (act && hs_claim[warp][e0 >> 1] == (unsigned char)lane) {
      // victim: a free way, else the way that was NOT used last.  Labels that keep coming back (the background,
      // segments much larger than a tile, and -- on a plane walk -- everything that continues one plane deeper)
      // stay; what has ended leaves.  (Plain direct mapping evicted the background's entry with every 128th new
      // label, and the ~650 k background records of a 1024^3 volume then queued up on ONE table slot: same-address
      // atomics serialise in L2.  "Keep the larger count" protected the background but let dead segments with
      // large counts squat in the sets for the rest of a plane walk.)
      const unsigned int c0 = s_cnt[e0], c1 = s_cnt[e0 + 1];
      const unsigned int v = k0 == FRB_EMPTY ? 0u : (k1 == FRB_EMPTY ? 1u : (hs_mru[warp][e0 >> 1] ? 0u : 1u));
      hs_mru[warp][e0 >> 1] = (unsigned char)v;
      ev.key = v ? k1 : k0;
      ev.cnt = v ? c1 : c0;
      s_key[e0 + v] = key;
      s_cnt[e0 + v] = c;
      act = false;
    }
    __syncwarp();
  }
  return ev;
}

// element j (run-time index) of a 16-byte vector, zero-extended
template <int W>
__device__ __forceinline__ u64 vec_get_dyn(const uint4& v, int j) {
  constexpr int PER = 4 / (W < 4 ? W : 4);  // elements per 32-bit word (W <= 4)
  if (W == 8) return j ? ((u64)v.z | ((u64)v.w << 32)) : ((u64)v.x | ((u64)v.y << 32));
  const int wi = W == 4 ? j : j / PER;
  const uint32_t w32 = wi == 0 ? v.x : (wi == 1 ? v.y : (wi == 2 ? v.z : v.w));
  if (W == 4) return (u64)w32;
  return (u64)((w32 >> ((j % PER) * 8 * W)) & ((1u << (8 * W)) - 1u));
}

// Send the (key, cnt) pairs of the lanes with `ev` on their way to global memory; called by all 32 lanes.
template <typename Sink>
__device__ __forceinline__ void hist_emit(u64 key, unsigned int cnt, bool ev, const Sink& sink, const HistLog& log, unsigned int* inserted) {
  if (Sink::USE_LOG && sink.logs()) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int m = __ballot_sync(FRB_FULL_MASK, ev);
    if (m == 0) return;
    const unsigned int cur = hs_logpos[warp];
    const unsigned int add = __popc(m);
    if (cur + add <= log.cap) {
      if (ev) {
        ulonglong2* dst = log.records + (size_t)(blockIdx.x * STREAM_WARPS + warp) * log.cap + cur + __popc(m & ((1u << lane) - 1u));
        asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(__cvta_generic_to_global(dst)), "l"(key), "l"((u64)cnt) : "memory");
      }
      __syncwarp();
      if (lane == 0) hs_logpos[warp] = cur + add;
      __syncwarp();
      return;
    }
    if (lane == 0 && log.cap) atomicOr(log.flags, 1u);
  }
  if (ev) sink(key, cnt, inserted);
}

// Handle `count` (<= 32) queued (vector, repeats) requests starting at ring position qhead, one per lane.
//  1. every lane splits its vector into runs of equal labels;
//  2. runs that continue across lane boundaries (the requests of one row are queued in lane order, and a
//     segment of a volume is usually wider than one 16-byte vector) are merged with a segmented warp scan,
//     so a label that spans several lanes costs ONE map update, made by the lane in which its run ends;
//  3. the (label, weighted length) pairs go to the warp's map (hist_map_add), one call per "k-th run that ends
//     in a lane" -- one or two calls per batch; what the map evicts is appended to the warp's log (hist_emit).
// Counts are 32-bit: a warp never sees 2^32 voxels (that would take a 15-terabyte volume).
template <int W, typename Sink>
__device__ __noinline__ unsigned int hist_process(int qhead, int count, const Sink sink, const HistLog log) {
  constexpr int V = 16 / W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool has = lane < count;
  const int pos = (qhead + lane) & (HIST_QUEUE - 1);
  const uint4 p = has ? hs_qpat[warp][pos] : make_uint4(0, 0, 0, 0);
  const unsigned int reps = has ? hs_qrep[warp][pos] : 0u;
  unsigned int inserted = 0;

  // runs inside the lane: bit j of `ends` = the run holding element j ends there (j < V - 1)
  const u64 kf = vec_get<W>(p, 0), kl = vec_get<W>(p, V - 1);
  unsigned int ends = 0, cl = 1;
#pragma unroll
  for (int j = 1; j < V; j++) {
    if (vec_get<W>(p, j) != vec_get<W>(p, j - 1)) {
      ends |= 1u << (j - 1);
      cl = 1;
    } else {
      cl++;
    }
  }
  // across lanes (lanes below `count` are contiguous): T = weighted length of the run open at my right edge
  const u64 kprev = __shfl_up_sync(FRB_FULL_MASK, kl, 1);
  const u64 knext = __shfl_down_sync(FRB_FULL_MASK, kf, 1);
  const bool joins = has && lane > 0 && kprev == kf;  // my first run continues the previous lane's last one
  unsigned int T = cl * reps;
  bool seg = !(ends == 0 && joins);                   // my last run starts inside me (or at my left edge)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int v = __shfl_up_sync(FRB_FULL_MASK, T, d);
    const bool f = __shfl_up_sync(FRB_FULL_MASK, (int)seg, d) != 0;
    if (lane >= d && !seg) {
      T += v;
      seg = f;
    }
  }
  const unsigned int Tprev = __shfl_up_sync(FRB_FULL_MASK, T, 1);
  const unsigned int carry = joins ? Tprev : 0u;      // what my first run inherits from the lanes to the left
  const bool tail_ends = has && (lane + 1 >= count || knext != kl);
  unsigned int pend = has ? (ends | (tail_ends ? 1u << (V - 1) : 0u)) : 0u;  // element positions at which a run ends in this lane

  int start = 0;  // first element of the lane's next run
  bool first_round = true;
  while (__any_sync(FRB_FULL_MASK, pend != 0)) {
    bool act = pend != 0;
    const int j = act ? __ffs(pend) - 1 : 0;
    const u64 key = vec_get_dyn<W>(p, j);
    // a run that reaches the right edge carries the scan's total (which includes what it inherited if the
    // whole lane is one run); a run that ends inside the lane is its own length, plus the carry if it is the first
    const unsigned int total = j == V - 1 ? T : (unsigned int)(j + 1 - start) * reps + (start == 0 ? carry : 0u);
    pend &= pend - 1;
    start = j + 1;
    HistEvicted ev;
    if (key == FRB_EMPTY) {  // the map's "free" marker: only a 64-bit all-ones label can collide with it -- bypass the map
      ev.key = FRB_EMPTY;
      ev.cnt = 0;
      if (act) sink(key, total, &inserted);
      act = false;
    }
#ifdef FRB_HIST_NOMAP  // experiment: merged runs straight to the log, no per-warp map
    ev.key = act ? key : FRB_EMPTY;
    ev.cnt = total;
#else
    if (first_round) {
      ev = hist_map_add(key, total, act);
    } else {
      // a lane's second, third ... run end: segments narrower than a 16-byte vector.  They bypass the map (a map
      // call costs ~100 warp instructions whether 1 or 32 lanes take part) and go to the log as they are
      ev.key = act ? key : FRB_EMPTY;
      ev.cnt = total;
    }
    first_round = false;
#endif
    hist_emit(ev.key, ev.cnt, ev.key != FRB_EMPTY, sink, log, &inserted);
  }
  return inserted;
}

// PZ: the volume is walked as a balanced plane walk (pwz, frb_common.cuh) instead of `run` / `pw`
template <int W, typename Sink, int IL, int DU, bool VAR_TILE, bool PZ = false>
__global__ void __launch_bounds__(STREAM_THREADS, 3) k_hist(const uint4* __restrict__ a, size_t n, const __grid_constant__ Sink sink,
                                                            const __grid_constant__ HistLog log, u64* count_target, int run, int tile_vecs,
                                                            PlaneWalk pw, PlaneWalkZ pwz) {
  constexpr int V = 16 / W;
  const size_t nvec = n / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool consumer = warp < STREAM_WARPS;
  const int wq = consumer ? warp : 0;  // the producer warp never touches the queues
  unsigned int inserted = 0;
  if (consumer) {
    for (int i = lane; i < HIST_WMAP; i += 32) { hs_key[warp][i] = FRB_EMPTY; hs_cnt[warp][i] = 0; }
    if (lane == 0) hs_logpos[warp] = 0;
  }
  __syncwarp();

  int qhead = 0, qn = 0;  // warp-uniform: ring position of the oldest request, number of requests queued
  const unsigned int lanes_below = (1u << lane) - 1u;

  uint4 pat[STREAM_ROWS];
  unsigned int rep[STREAM_ROWS];
#pragma unroll
  for (int u = 0; u < STREAM_ROWS; u++) { pat[u] = make_uint4(0, 0, 0, 0); rep[u] = 0; }

  // queue the lanes' retired patterns; once 32 are waiting, handle 32
  auto retire = [&](bool push, const uint4& p, unsigned int r) {
    const unsigned int m = __ballot_sync(FRB_FULL_MASK, push);
    if (m == 0) return;
    if (push) {
      const int pos = (qhead + qn + __popc(m & lanes_below)) & (HIST_QUEUE - 1);
      hs_qpat[wq][pos] = p;
      hs_qrep[wq][pos] = r;
    }
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) {
      inserted += hist_process<W>(qhead, 32, sink, log);
      qhead = (qhead + 32) & (HIST_QUEUE - 1);
      qn -= 32;
      __syncwarp();
    }
  };

  auto row_fn = [&](int u, const uint4& cur, bool valid, size_t, u64, bool) {
    const int sl = hist_slot<DU>(u);  // a constant once the row loop is unrolled
    const uint4 l = pat[sl];
    const uint32_t d = (cur.x ^ l.x) | (cur.y ^ l.y) | (cur.z ^ l.z) | (cur.w ^ l.w);
    const bool same = valid && d == 0 && rep[sl] != 0;
    // the common case first, and cheap (this loop is issue-bound): every lane saw the row above again
    if (__all_sync(FRB_FULL_MASK, same)) {
      rep[sl]++;
      return;
    }
    retire(valid && !same && rep[sl] != 0, l, rep[sl]);
    if (valid) {
      if (same) {
        rep[sl]++;
      } else {
        pat[sl] = cur;
        rep[sl] = 1;
      }
    }
  };

  if constexpr (PZ) {
    stream_rows_z<W, false, false, IL>(a, nvec, TileMapZ(pwz), row_fn, [](const uint4*, size_t) {});
  } else {
    const TileMap map = (!VAR_TILE && pw.band) ? TileMap(pw)
                                               : TileMap(nvec, (size_t)run, VAR_TILE ? (size_t)tile_vecs : (size_t)STREAM_TILE_VECS);
    stream_rows<W, false, false, IL, VAR_TILE>(a, nvec, map, row_fn, [](const uint4*, size_t) {}, [](size_t) {});
  }
  if (consumer) {
    // what the lanes still hold, what is still queued, then the map itself
#pragma unroll
    for (int u = 0; u < STREAM_ROWS; u++) retire(rep[u] != 0, pat[u], rep[u]);
    if (qn > 0) inserted += hist_process<W>(qhead, qn, sink, log);
    __syncwarp();
    for (int i = lane; i < HIST_WMAP; i += 32) {
      const u64 k = hs_key[warp][i];
      hist_emit(k, hs_cnt[warp][i], k != FRB_EMPTY, sink, log, &inserted);
    }
    if (Sink::USE_LOG && log.cap && lane == 0) log.counts[blockIdx.x * STREAM_WARPS + warp] = hs_logpos[warp];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = nvec * V; i < n; i++) sink(ld_elem<W>(a, i), 1u, &inserted);
  if (count_target) block_add_to(count_target, inserted);
}

// meta[COUNT] of a direct table after a counting pass: the slots that hold a label
__global__ void __launch_bounds__(256) k_direct_count(const Slot* slots, u64 cap, u64* meta) {
  unsigned int mine = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) mine += ld_slot_cg(slots + i).key != FRB_EMPTY;
  mine = __reduce_add_sync(FRB_FULL_MASK, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&meta[FRB_META_COUNT], (u64)mine);
}

// Direct tables in the dtype's own order: slot index = label bits, so walking the slots (from the sign bit for signed
// dtypes) IS the sorted order -- an ordered compaction replaces export + radix sort (five launches, one of them
// cooperative).  Two launches of cap / 2048 CTAs: per-CTA counts, then every CTA adds up the counts in front of it and
// emits its labels and counts.
static constexpr int DIRECT_CHUNK = 2048;  // slots per CTA (256 threads x 8 consecutive slots)
__global__ void __launch_bounds__(256) k_direct_block_counts(const Slot* slots, u64 cap, u64 flip, unsigned int* block_counts) {
  unsigned int mine = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const u64 i = (u64)blockIdx.x * DIRECT_CHUNK + threadIdx.x * 8 + q;
    if (i < cap) mine += ld_slot_cg(slots + (i ^ flip)).key != FRB_EMPTY;
  }
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  mine = __reduce_add_sync(FRB_FULL_MASK, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s, mine);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}
template <int W>
__global__ void __launch_bounds__(256) k_direct_emit(const Slot* slots, u64 cap, u64 flip, const unsigned int* block_counts, u64* meta,
                                                     void* uniq_out, u64* counts_out, u64 capacity) {
  __shared__ unsigned int s_warp[8];
  __shared__ unsigned int s_base;
  if (threadIdx.x < 32) {  // CTAs in front of this one (at most 32 of them), and the total for meta
    unsigned int before = 0, total = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) {
      const unsigned int c = __ldcg(block_counts + b);
      total += c;
      if (b < blockIdx.x) before += c;
    }
    before = __reduce_add_sync(FRB_FULL_MASK, before);
    total = __reduce_add_sync(FRB_FULL_MASK, total);
    if (threadIdx.x == 0) {
      s_base = before;
      if (blockIdx.x == 0) meta[FRB_META_LIST_COUNT] = total;
    }
  }
  Slot sl[8];
  unsigned int mine = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const u64 i = (u64)blockIdx.x * DIRECT_CHUNK + threadIdx.x * 8 + q;
    sl[q].key = FRB_EMPTY;
    sl[q].val = 0;
    if (i < cap) sl[q] = ld_slot_cg(slots + (i ^ flip));
    mine += sl[q].key != FRB_EMPTY;
  }
  // exclusive scan of `mine` over the CTA's 256 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int v = __shfl_up_sync(FRB_FULL_MASK, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned int pos = s_base + incl - mine;
  for (int w2 = 0; w2 < warp; w2++) pos += s_warp[w2];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    if (sl[q].key != FRB_EMPTY) {
      if (pos < capacity) {
        st_elem<W>(uniq_out, pos, sl[q].key);
        counts_out[pos] = sl[q].val;
      }
      pos++;
    }
  }
}

// Pull a table into L2 (one prefetch per 128-byte line, nothing lands in a register).  The counting pass has
// just streamed the whole volume through L2, so the table initialised before it is back in HBM; folding the
// logs into it is millions of RANDOM 32-byte accesses, which HBM serves at a small fraction of its streaming
// rate (measured: 0.31 ms for 3.1 M records into a 64 MiB table, against 0.05 ms when the table is cache-resident).
__global__ void __launch_bounds__(256) k_prefetch_l2(const char* base, size_t bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 128;
  for (size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes; off += stride)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
}

// Fold the eviction logs of all warps into the label table.  A CTA takes one region (one warp's log) at a time:
//   1. its records are aggregated in a shared-memory (label, count) table first.  Labels that are everywhere -- the
//      background, segments that cross many warps' bands -- are evicted from thousands of warp maps, and their
//      records would otherwise queue up on ONE slot of the label table: same-address atomics serialise in L2
//      (measured: 79 k records of label 0 among 3.1 M made this kernel take 0.3 ms whatever the table's size);
//   2. the distinct labels of the region then go to the table, four per thread with their home-slot keys read
//      together (one L2 round trip for the four); a label already sitting in its home slot costs one
//      fire-and-forget reduction, first sightings and collisions take the probing path.
static constexpr int APPLY_SLOTS = 2048;  // >= 2 x the 4096-record cap of a region / 4 ... regions hold ~10^3 records
__global__ void __launch_bounds__(256) k_hist_apply_log(HistLog log, unsigned int regions, HashSink sink, u64* count_target) {
  __shared__ u64 a_key[APPLY_SLOTS];
  __shared__ unsigned int a_cnt[APPLY_SLOTS];
  unsigned int inserted = 0;
  for (unsigned int r = blockIdx.x; r < regions; r += gridDim.x) {
    const unsigned int cnt = log.counts[r];
    if (cnt == 0) continue;  // (the same in every thread)
    const ulonglong2* rec = log.records + (size_t)r * log.cap;
    for (int e = threadIdx.x; e < APPLY_SLOTS; e += 256) { a_key[e] = FRB_EMPTY; a_cnt[e] = 0; }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < cnt; i += 256) {
      ulonglong2 v;
      asm("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(rec + i));
      bool done = false;
      if (v.x != FRB_EMPTY) {
        unsigned int e = (unsigned int)(hash_key(v.x) >> 40) & (APPLY_SLOTS - 1);
#pragma unroll 1
        for (int probe = 0; probe < 8 && !done; probe++, e = (e + 1) & (APPLY_SLOTS - 1)) {
          u64 k = a_key[e];
          if (k == FRB_EMPTY) {
            const u64 old = atomicCAS((unsigned long long*)&a_key[e], (unsigned long long)FRB_EMPTY, (unsigned long long)v.x);
            k = old == FRB_EMPTY ? v.x : old;
          }
          if (k == v.x) {
            atomicAdd(&a_cnt[e], (unsigned int)v.y);
            done = true;
          }
        }
      }
      if (!done) sink(v.x, (unsigned int)v.y, &inserted);  // the all-ones label, or a crowded corner of the shared table
    }
    __syncthreads();
    for (int e0 = threadIdx.x; e0 < APPLY_SLOTS; e0 += 4 * 256) {
      u64 k[4], home[4];
      unsigned int c[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        k[q] = a_key[e0 + q * 256];
        c[q] = a_cnt[e0 + q * 256];
        home[q] = 0;
        if (k[q] != FRB_EMPTY) home[q] = sink.peek(k[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (k[q] != FRB_EMPTY) sink.commit(k[q], c[q], home[q], &inserted);
    }
    __syncthreads();
  }
  if (count_target) block_add_to(count_target, inserted);
}

// ------------------------------------------------------------------ ordered compaction of bins
static inline int compact_blocks(size_t n) {
  size_t need = (n + 2047) / 2048;
  size_t cap = (size_t)sm_count() * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
static inline size_t compact_chunk(size_t n, int blocks) {
  size_t c = (n + blocks - 1) / blocks;
  return align_up(c < 1 ? 1 : c, 256);
}

// block-wide exclusive prefix of a 0/1 flag (256 threads); returns the prefix, *total = block sum
__device__ __forceinline__ unsigned int block_flag_scan(bool flag, unsigned int* total) {
  __shared__ unsigned int wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int b = __ballot_sync(FRB_FULL_MASK, flag);
  if (lane == 0) wsum[warp] = __popc(b);
  __syncthreads();
  unsigned int pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const unsigned int s = wsum[w];
    if (w < warp) pre += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return pre + __popc(b & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256)
k_bins_count(const u64* __restrict__ counts, u64 bins, size_t chunk, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > bins) end = bins;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += counts[i] != 0 ? 1 : 0;
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  c = __reduce_add_sync(FRB_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}

template <int W>
__global__ void __launch_bounds__(256)
k_bins_emit(const u64* __restrict__ counts, u64 bins, size_t chunk, const u64* __restrict__ block_offs, u64 min_bits,
            void* __restrict__ uniq_out, u64* __restrict__ counts_out, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > bins) end = bins;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const u64 c = i < end ? counts[i] : 0;
    unsigned int tot;
    const unsigned int pre = block_flag_scan(c != 0, &tot);
    if (c != 0) {
      const u64 pos = run + pre;
      if (pos < capacity) {
        st_elem<W>(uniq_out, pos, min_bits + (u64)i);
        counts_out[pos] = c;
      }
    }
    run += tot;
  }
}

// ------------------------------------------------------------------ ordered selection: indices(arr, value), point_cloud's foreground
// "positions (and labels) of the voxels that satisfy a predicate, in memory order" in two passes, because the caller
// sizes the output: _count leaves per-CTA counts (scanned into offsets afterwards), _emit writes.  A CTA owns a
// contiguous chunk of the array; a thread takes one 16-byte vector per step (V voxels), so a block-wide scan and its
// two barriers are paid once per 256 * V voxels (the first version scanned one flag per voxel: indices ran at 0.36
// of the HBM roofline).  Buffers that are not 16-byte aligned take the element-wise kernels below.
struct MatchPred {   // indices: arr[i] == value
  u64 value;
  __device__ __forceinline__ bool operator()(u64 k) const { return k == value; }
};
struct NonzeroPred {  // point_cloud: foreground voxels
  __device__ __forceinline__ bool operator()(u64 k) const { return k != 0; }
};
struct IndexEmit {  // out[pos] = flat index
  u64* out;
  __device__ __forceinline__ void operator()(u64 pos, u64, u64 i) const { out[pos] = i; }
};
struct LabelIndexEmit {  // (label bits with the sign bit flipped, flat index)
  u64* labels_out;
  u64* index_out;
  u64 flip;
  __device__ __forceinline__ void operator()(u64 pos, u64 k, u64 i) const { labels_out[pos] = k ^ flip; index_out[pos] = i; }
};

// block-wide exclusive prefix of a per-thread count (256 threads); returns the prefix, *total = block sum
__device__ __forceinline__ unsigned int block_count_scan(unsigned int c, unsigned int* total) {
  __shared__ unsigned int wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(FRB_FULL_MASK, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  unsigned int pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const unsigned int s = wsum[w];
    if (w < warp) pre += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return pre + incl - c;
}

template <int W, typename Pred>
__global__ void __launch_bounds__(256)
k_select_count(const uint4* __restrict__ a, size_t n, size_t chunk, Pred pred, u64* __restrict__ block_counts) {
  constexpr int V = 16 / W;
  const size_t beg = (size_t)blockIdx.x * chunk;  // chunks are multiples of 256 elements: whole vectors
  size_t end = beg + chunk;
  if (end > n) end = n;
  const size_t vend = end / V;
  unsigned int c = 0;
  size_t vi = beg / V + threadIdx.x;
  for (; vi + 3 * 256 < vend; vi += 4 * 256) {  // four vectors in flight per thread
    uint4 v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = ld_stream16(a + vi + q * 256);
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int j = 0; j < V; j++) c += pred(vec_get<W>(v[q], j)) ? 1u : 0u;
  }
  for (; vi < vend; vi += 256) {
    const uint4 v = ld_stream16(a + vi);
#pragma unroll
    for (int j = 0; j < V; j++) c += pred(vec_get<W>(v, j)) ? 1u : 0u;
  }
  if (threadIdx.x == 0)
    for (size_t i = vend * V > beg ? vend * V : beg; i < end; i++) c += pred(ld_elem<W>(a, i)) ? 1u : 0u;  // < V elements at the very end
  unsigned int tot;
  block_count_scan(c, &tot);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = tot;
}

template <int W, typename Pred, typename Emit>
__global__ void __launch_bounds__(256)
k_select_emit(const uint4* __restrict__ a, size_t n, size_t chunk, Pred pred, const u64* __restrict__ block_offs, Emit emit, size_t capacity) {
  constexpr int V = 16 / W;
  // positions (inside the step's 256 * V voxels) of the hits, in order: the hits of a step leave the CTA through
  // this list so that consecutive threads write consecutive output elements -- written straight from the vector
  // loop, a dense selection (point_cloud: 90 % foreground) stored with a stride of V elements per lane, 0.30 of peak
  __shared__ unsigned short s_off[256 * V];
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  const size_t vend = end / V;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg / V; base < vend; base += 256) {
    const size_t vi = base + threadIdx.x;
    unsigned int m = 0;
    if (vi < vend) {
      const uint4 v = ld_stream16(a + vi);
#pragma unroll
      for (int j = 0; j < V; j++) m |= pred(vec_get<W>(v, j)) ? (1u << j) : 0u;
    }
    unsigned int tot;
    unsigned int p = block_count_scan((unsigned int)__popc(m), &tot);
#pragma unroll
    for (int j = 0; j < V; j++)
      if ((m >> j) & 1u) s_off[p++] = (unsigned short)(threadIdx.x * V + j);
    __syncthreads();
    for (unsigned int q = threadIdx.x; q < tot; q += 256) {
      const u64 pos = run + q;
      if (pos < capacity) {
        const size_t i = base * V + s_off[q];
        emit(pos, ld_elem<W>(a, i), (u64)i);  // the element has just been read: an L1 / L2 hit
      }
    }
    __syncthreads();
    run += tot;
  }
  if (threadIdx.x == 0)
    for (size_t i = vend * V > beg ? vend * V : beg; i < end; i++) {
      const u64 k = ld_elem<W>(a, i);
      if (pred(k)) {
        if (run < capacity) emit(run, k, (u64)i);
        run++;
      }
    }
}

// element-wise fallbacks (unaligned buffers)
template <int W, typename Pred>
__global__ void __launch_bounds__(256)
k_select_count_elems(const void* __restrict__ a, size_t n, size_t chunk, Pred pred, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += pred(ld_elem<W>(a, i)) ? 1u : 0u;
  unsigned int tot;
  block_count_scan(c, &tot);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = tot;
}

template <int W, typename Pred, typename Emit>
__global__ void __launch_bounds__(256)
k_select_emit_elems(const void* __restrict__ a, size_t n, size_t chunk, Pred pred, const u64* __restrict__ block_offs, Emit emit, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const u64 k = i < end ? ld_elem<W>(a, i) : 0;
    const bool hit = i < end && pred(k);
    unsigned int tot;
    const unsigned int pre = block_flag_scan(hit, &tot);
    if (hit && run + pre < capacity) emit(run + pre, k, (u64)i);
    run += tot;
  }
}

template <typename Pred>
static int launch_select_count(const void* a, size_t n, int w, const Pred& pred, u64* block_counts, cudaStream_t st) {
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const bool vec = aligned16(a);
#define FRB_SC(WW) if (vec) k_select_count<WW, Pred><<<blocks, 256, 0, st>>>((const uint4*)a, n, chunk, pred, block_counts); \
                   else k_select_count_elems<WW, Pred><<<blocks, 256, 0, st>>>(a, n, chunk, pred, block_counts);
  switch (w) {
    case 1: FRB_SC(1) break;
    case 2: FRB_SC(2) break;
    case 4: FRB_SC(4) break;
    default: FRB_SC(8) break;
  }
#undef FRB_SC
  return check_launch("k_select_count");
}
template <typename Pred, typename Emit>
static int launch_select_emit(const void* a, size_t n, int w, const Pred& pred, const u64* block_offs, const Emit& emit, size_t capacity, cudaStream_t st) {
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const bool vec = aligned16(a);
#define FRB_SE(WW) if (vec) k_select_emit<WW, Pred, Emit><<<blocks, 256, 0, st>>>((const uint4*)a, n, chunk, pred, block_offs, emit, capacity); \
                   else k_select_emit_elems<WW, Pred, Emit><<<blocks, 256, 0, st>>>(a, n, chunk, pred, block_offs, emit, capacity);
  switch (w) {
    case 1: FRB_SE(1) break;
    case 2: FRB_SE(2) break;
    case 4: FRB_SE(4) break;
    default: FRB_SE(8) break;
  }
#undef FRB_SE
  return check_launch("k_select_emit");
}

// flat C-order index -> (i, j[, k]) as uint16 rows
__global__ void __launch_bounds__(256)
k_unravel_u16(const u64* __restrict__ index, size_t n, int ndim, u64 d1, u64 d2, uint16_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const u64 f = index[p];
    if (ndim == 2) {
      out[2 * p] = (uint16_t)(f / d1);
      out[2 * p + 1] = (uint16_t)(f % d1);
    } else {
      const u64 ij = f / d2;
      out[3 * p] = (uint16_t)(ij / d1);
      out[3 * p + 1] = (uint16_t)(ij % d1);
      out[3 * p + 2] = (uint16_t)(f % d2);
    }
  }
}

// ------------------------------------------------------------------ sorted (label, count) pairs
// One persistent kernel: flip the sign bit of signed labels (so that unsigned digit order is the
// dtype's order), sort the pairs by label (frb_coopsort.cuh), emit typed labels + values.
struct SortPairsArgs {
  CoopSortBuffers sort;
  const u64* n_dev;  // optional: the list length lives on the device (<= n)
  u64 n;
  u64 flip;
  int key_bits;
  void* uniq_out;
  u64* vals_out;
};

template <int W>
__global__ void __launch_bounds__(CS_THREADS) k_sort_emit_pairs(SortPairsArgs A) {
  if (A.n_dev) {
    const u64 nd = __ldcg(A.n_dev);
    A.n = nd < A.n ? nd : A.n;
  }
  const u64 stride = (u64)gridDim.x * CS_THREADS;
  unsigned int generation = 0;
  if (A.flip) {
    for (u64 i = (u64)blockIdx.x * CS_THREADS + threadIdx.x; i < A.n; i += stride) A.sort.keys[i] ^= A.flip;
    grid_barrier(A.sort.barrier, generation);
  }
  u64* kin;
  u64* vin;
  coop_radix_sort(A.sort, A.n, 0, A.key_bits, generation, &kin, &vin);
  for (u64 i = (u64)blockIdx.x * CS_THREADS + threadIdx.x; i < A.n; i += stride) {
    st_elem<W>(A.uniq_out, i, __ldcg(kin + i) ^ A.flip);
    A.vals_out[i] = __ldcg(vin + i);
  }
}

// ------------------------------------------------------------------ sort + run-length encode
template <int W, typename K>
__global__ void __launch_bounds__(256) k_widen_flip(const void* __restrict__ a, size_t n, K* __restrict__ out, K flip) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (K)ld_elem<W>(a, i) ^ flip;
}

template <typename K>
__global__ void __launch_bounds__(256)
k_heads_count(const K* __restrict__ sorted, size_t n, size_t chunk, u64* __restrict__ block_counts) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  unsigned int c = 0;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) c += (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  c = __reduce_add_sync(FRB_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s;
}

// A run [start, end) of label k with output position p: the head writes uniq[p] and adds -start
// to counts[p]; the tail adds end.  counts_out is zeroed by the caller, so counts[p] = end - start.
template <int W, typename K>
__global__ void __launch_bounds__(256)
k_heads_emit(const K* __restrict__ sorted, size_t n, size_t chunk, const u64* __restrict__ block_offs, K flip,
             void* __restrict__ uniq_out, u64* __restrict__ counts_out, size_t capacity) {
  const size_t beg = (size_t)blockIdx.x * chunk;
  size_t end = beg + chunk;
  if (end > n) end = n;
  u64 run = block_offs[blockIdx.x];
  for (size_t base = beg; base < end; base += 256) {
    const size_t i = base + threadIdx.x;
    const bool in = i < end;
    const K k = in ? sorted[i] : (K)0;
    const bool head = in && (i == 0 || k != sorted[i - 1]);
    const bool tail = in && (i + 1 == n || k != sorted[i + 1]);
    unsigned int tot;
    const unsigned int pre = block_flag_scan(head, &tot);
    // output position of the run this element belongs to = (#heads at or before i) - 1
    const u64 p = run + pre + (head ? 1 : 0) - 1;
    if (head && p < capacity) {
      st_elem<W>(uniq_out, p, (u64)(k ^ flip));
      atomicAdd(&counts_out[p], (u64)0 - (u64)i);
    }
    if (tail && p < capacity) atomicAdd(&counts_out[p], (u64)i + 1);
    run += tot;
  }
}

}  // namespace frb

using namespace frb;

extern "C" {

// eviction-log workspace: [counts: regions x u32 | flags: 256 B | records: regions x cap x 16 B]
static inline unsigned int hist_regions_max() { return (unsigned int)sm_count() * 4u * STREAM_WARPS; }  // >= any k_hist grid x 8
static inline size_t hist_counts_bytes() { return align_up((size_t)hist_regions_max() * sizeof(unsigned int), 256); }
static inline unsigned int hist_log_cap(size_t n) {
  // records per warp: a record stands for a segment cross-section, tens to hundreds of voxels; 1/16 of the
  // warp's share of the volume, between 256 and 4096 records (64 KiB) -- beyond that the warp updates the table itself
  const size_t share = n / hist_regions_max() / 16;
  return (unsigned int)(share < 256 ? 256 : (share > 4096 ? 4096 : share));
}
size_t frb_hist_ws_bytes(size_t n) { return hist_counts_bytes() + 256 + (size_t)hist_regions_max() * hist_log_cap(n) * sizeof(ulonglong2); }

static HistLog hist_log_of(void* ws, size_t ws_bytes, size_t n) {
  HistLog log;
  log.records = nullptr;
  log.counts = nullptr;
  log.flags = nullptr;
  log.cap = 0;
  if (ws && ws_bytes >= frb_hist_ws_bytes(n)) {
    log.counts = (unsigned int*)ws;
    log.flags = (unsigned int*)((char*)ws + hist_counts_bytes());
    log.records = (ulonglong2*)((char*)ws + hist_counts_bytes() + 256);
    log.cap = hist_log_cap(n);
  }
  return log;
}

int frb_hist_dense(const void* a, size_t n, int dtype, uint64_t min_bits, uint64_t* counts, uint64_t bins, uint64_t row_bytes, void* stream) {
  if (!dtype_ok(dtype) || !counts || bins == 0 || (n && (!a || !aligned16(a)))) { set_error("frb_hist_dense: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  DenseSink sink;
  sink.counts = (u64*)counts;
  sink.min_bits = trunc_width(min_bits, w);
  sink.w = w;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)a;
  // the dense strategy is only ever forced (tests, comparisons): two geometries are enough
  const HistLog nolog = hist_log_of(nullptr, 0, n);
  const PlaneWalk nowalk = {0, 0, 0};
  const RowGeom geo = row_geom((size_t)row_bytes);
  const int tv = geo.tv;
#define FRB_KHD(WW, VV) k_hist<WW, DenseSink, 8, 4, VV><<<stream_grid(k_hist<WW, DenseSink, 8, 4, VV>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>( \
    p, n, sink, nolog, nullptr, fair_run(stream_grid(k_hist<WW, DenseSink, 8, 4, VV>, tiles), tiles), tv, nowalk, PlaneWalkZ{0, 0, 0})
#define FRB_KHD_GEO(WW) if (tv != STREAM_TILE_VECS) { FRB_KHD(WW, true); } else { FRB_KHD(WW, false); }
  switch (w) {
    case 1: FRB_KHD_GEO(1) break;
    case 2: FRB_KHD_GEO(2) break;
    case 4: FRB_KHD_GEO(4) break;
    default: FRB_KHD_GEO(8) break;
  }
#undef FRB_KHD_GEO
#undef FRB_KHD
  return check_launch("k_hist<dense>");
}

int frb_hist_hash(const void* a, size_t n, int dtype, void* slots, uint64_t cap, uint64_t* meta, uint64_t row_bytes,
                  uint64_t plane_bytes, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !slots || !meta || (cap & (cap - 1)) != 0 || (n && (!a || !aligned16(a))) || (ws && (!aligned16(ws) || ws_bytes < frb_hist_ws_bytes(n)))) {
    set_error("frb_hist_hash: bad arguments (ws must hold frb_hist_ws_bytes(n), or be NULL)");
    return FRB_ERR_BAD_ARG;
  }
  if (n == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  HashSink sink;
  sink.slots = (Slot*)slots;
  sink.mask = cap - 1;
  sink.direct = w <= 2 ? 1 : 0;
  sink.meta = (u64*)meta;
  if (sink.direct && cap != frb_table_direct_cap(dtype)) { set_error("frb_hist_hash: direct table needs cap == 2^bits"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tiles = stream_tiles(n, w);
  const uint4* p = (const uint4*)a;
  u64* ct = &((u64*)meta)[FRB_META_COUNT];
  const RowGeom geo = row_geom((size_t)row_bytes);
  const int tv = geo.tv;
  HistLog log = hist_log_of(ws, ws_bytes, n);
  if (sink.direct) log.cap = 0;  // direct tables are updated fire-and-forget (HashSink::direct_add): no log, no fold
  if (log.cap) cudaMemsetAsync(ws, 0, hist_counts_bytes() + 256, st);  // regions that no warp owns in this launch stay empty
  bool launched_z = false;
#define FRB_KH(WW, II, DD, VV) k_hist<WW, HashSink, II, DD, VV><<<stream_grid(k_hist<WW, HashSink, II, DD, VV>, tiles), STREAM_THREADS, STREAM_SMEM_BYTES, st>>>( \
    p, n, sink, log, ct, fair_run(stream_grid(k_hist<WW, HashSink, II, DD, VV>, tiles), tiles), tv, \
    (VV) ? PlaneWalk{0, 0, 0} : plane_walk_for(n * (size_t)WW, (n % (16 / WW)) ? 0 : (size_t)plane_bytes, stream_grid(k_hist<WW, HashSink, II, DD, VV>, tiles)), \
    PlaneWalkZ{0, 0, 0})
#define FRB_KHZ(WW, DD) { \
    const int gz = stream_grid(k_hist<WW, HashSink, 8, DD, false, true>, tiles); \
    const PlaneWalkZ pz = plane_walk_z_for(n * (size_t)WW, (n % (16 / WW)) ? 0 : (size_t)plane_bytes, gz, (size_t)plane_band()); \
    if (pz.band) { \
      k_hist<WW, HashSink, 8, DD, false, true><<<gz, STREAM_THREADS, STREAM_SMEM_BYTES, st>>>(p, n, sink, log, ct, 1, tv, PlaneWalk{0, 0, 0}, pz); \
      launched_z = true; \
    } }
#define FRB_KH_GEO(WW)                                                        \
  if (geo.tv == STREAM_TILE_VECS && geo.il == 8) {                            \
    if (geo.du == 1) FRB_KHZ(WW, 1) else if (geo.du == 2) FRB_KHZ(WW, 2) else if (geo.du == 3) FRB_KHZ(WW, 3) else FRB_KHZ(WW, 4) \
  }                                                                           \
  if (launched_z) {}                                                          \
  else if (geo.tv != STREAM_TILE_VECS) { FRB_KH(WW, 8, 4, true); }            \
  else if (geo.il == 4) { FRB_KH(WW, 4, 1, false); }                          \
  else if (geo.il == 2) { FRB_KH(WW, 2, 1, false); }                          \
  else if (geo.il == 1) { FRB_KH(WW, 1, 1, false); }                          \
  else if (geo.du == 1) { FRB_KH(WW, 8, 1, false); }                          \
  else if (geo.du == 2) { FRB_KH(WW, 8, 2, false); }                          \
  else if (geo.du == 3) { FRB_KH(WW, 8, 3, false); }                          \
  else { FRB_KH(WW, 8, 4, false); }
  switch (w) {
    case 1: FRB_KH_GEO(1) break;
    case 2: FRB_KH_GEO(2) break;
    case 4: FRB_KH_GEO(4) break;
    default: FRB_KH_GEO(8) break;
  }
#undef FRB_KH_GEO
#undef FRB_KHZ
#undef FRB_KH
  int rc = check_launch("k_hist<hash>");
  if (rc) return rc;
  if (sink.direct) {  // COUNT = labels in the table (the whole table: COUNT of a direct table is never partial)
    cudaMemsetAsync(ct, 0, sizeof(u64), st);
    k_direct_count<<<cap >= 65536 ? 32 : 1, 256, 0, st>>>((const Slot*)slots, cap, (u64*)meta);
    return check_launch("k_direct_count");
  }
  if (!log.cap) return rc;
  return frb_hist_apply_log(ws, ws_bytes, n, dtype, slots, cap, meta, stream);
}

// second step of frb_hist_hash (which calls it itself): fold the eviction logs in ws into the table.  Callable again
// on a bigger, freshly initialised table when the first one overflowed -- provided ws says that no warp bypassed its
// log (frb_hist_log_complete): the logs then hold everything the pass over the volume produced.
int frb_hist_apply_log(const void* ws, size_t ws_bytes, size_t n, int dtype, void* slots, uint64_t cap, uint64_t* meta, void* stream) {
  if (!dtype_ok(dtype) || !ws || ws_bytes < frb_hist_ws_bytes(n) || !slots || !meta || (cap & (cap - 1)) != 0) { set_error("frb_hist_apply_log: bad arguments"); return FRB_ERR_BAD_ARG; }
  const int w = dtype_width(dtype);
  HashSink sink;
  sink.slots = (Slot*)slots;
  sink.mask = cap - 1;
  sink.direct = w <= 2 ? 1 : 0;
  sink.meta = (u64*)meta;
  const HistLog log = hist_log_of((void*)ws, ws_bytes, n);
  const unsigned int regions = hist_regions_max();
  const size_t table_bytes = (size_t)cap * sizeof(Slot);
  if (table_bytes <= ((size_t)96 << 20)) {  // fits the 126 MB L2 next to the log records that stream through
    k_prefetch_l2<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>((const char*)slots, table_bytes);
    int rc = check_launch("k_prefetch_l2");
    if (rc) return rc;
  }
  k_hist_apply_log<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(log, regions, sink, &((u64*)meta)[FRB_META_COUNT]);
  return check_launch("k_hist_apply_log");
}

int frb_direct_export_sorted(const void* slots, uint64_t cap, int dtype, uint64_t* meta, void* uniq_out, uint64_t* vals_out,
                             size_t capacity, void* ws, size_t ws_bytes, void* stream) {
  const int w = dtype_ok(dtype) ? dtype_width(dtype) : 0;
  if (w < 1 || w > 2 || !slots || !meta || cap != frb_table_direct_cap(dtype) || (capacity && (!uniq_out || !vals_out)) || !ws ||
      ws_bytes < 256) {
    set_error("frb_direct_export_sorted: bad arguments (8- / 16-bit dtypes, a direct table, 256 bytes of workspace)");
    return FRB_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)((cap + DIRECT_CHUNK - 1) / DIRECT_CHUNK);  // 1 or 32
  const u64 flip = dtype_signed(dtype) ? cap / 2 : 0;
  k_direct_block_counts<<<blocks, 256, 0, st>>>((const Slot*)slots, cap, flip, (unsigned int*)ws);
  int rc = check_launch("k_direct_block_counts");
  if (rc) return rc;
  if (w == 1) k_direct_emit<1><<<blocks, 256, 0, st>>>((const Slot*)slots, cap, flip, (const unsigned int*)ws, (u64*)meta, uniq_out, (u64*)vals_out, (u64)capacity);
  else k_direct_emit<2><<<blocks, 256, 0, st>>>((const Slot*)slots, cap, flip, (const unsigned int*)ws, (u64*)meta, uniq_out, (u64*)vals_out, (u64)capacity);
  return check_launch("k_direct_emit");
}

// byte offset inside ws of the 32-bit flag word: 0 after a pass = every eviction went through the logs
size_t frb_hist_ws_flags_offset(void) { return hist_counts_bytes(); }

size_t frb_compact_bins_ws_bytes(uint64_t bins) { return (size_t)compact_blocks((size_t)bins) * sizeof(u64) + 256; }

int frb_compact_bins_count(const uint64_t* counts, uint64_t bins, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!counts || !meta || !ws || bins == 0 || ws_bytes < frb_compact_bins_ws_bytes(bins)) { set_error("frb_compact_bins_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks((size_t)bins);
  const size_t chunk = compact_chunk((size_t)bins, blocks);
  k_bins_count<<<blocks, 256, 0, st>>>((const u64*)counts, bins, chunk, (u64*)ws);
  int rc = check_launch("k_bins_count");
  if (rc) return rc;
  return scan_exclusive_u64((u64*)ws, (size_t)blocks, &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_compact_bins_emit(const uint64_t* counts, uint64_t bins, int dtype, uint64_t min_bits, void* uniq_out,
                          uint64_t* counts_out, size_t capacity, const void* ws, void* stream) {
  if (!dtype_ok(dtype) || !counts || !ws || bins == 0 || (capacity && (!uniq_out || !counts_out))) { set_error("frb_compact_bins_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compact_blocks((size_t)bins);
  const size_t chunk = compact_chunk((size_t)bins, blocks);
  const int w = dtype_width(dtype);
  const u64 mb = trunc_width(min_bits, w);
  const u64* c = (const u64*)counts;
  const u64* offs = (const u64*)ws;
  u64* co = (u64*)counts_out;
  switch (w) {
    case 1: k_bins_emit<1><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
    case 2: k_bins_emit<2><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
    case 4: k_bins_emit<4><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
    default: k_bins_emit<8><<<blocks, 256, 0, st>>>(c, bins, chunk, offs, mb, uniq_out, co, capacity); break;
  }
  return check_launch("k_bins_emit");
}

size_t frb_indices_ws_bytes(size_t n) { return (size_t)compact_blocks(n) * sizeof(u64) + 256; }

int frb_indices_count(const void* a, size_t n, int dtype, uint64_t value_bits, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !meta || !ws || ws_bytes < frb_indices_ws_bytes(n)) { set_error("frb_indices_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  MatchPred pred;
  pred.value = trunc_width(value_bits, w);
  int rc = launch_select_count(a, n, w, pred, (u64*)ws, st);
  if (rc) return rc;
  return scan_exclusive_u64((u64*)ws, (size_t)compact_blocks(n), &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_indices_emit(const void* a, size_t n, int dtype, uint64_t value_bits, uint64_t* out, size_t capacity, const void* ws, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !ws || (capacity && !out)) { set_error("frb_indices_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (capacity == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  MatchPred pred;
  pred.value = trunc_width(value_bits, w);
  IndexEmit emit;
  emit.out = (u64*)out;
  return launch_select_emit(a, n, w, pred, (const u64*)ws, emit, capacity, (cudaStream_t)stream);
}

size_t frb_nonzero_ws_bytes(size_t n) { return (size_t)compact_blocks(n) * sizeof(u64) + 256; }

int frb_nonzero_count(const void* a, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !meta || !ws || ws_bytes < frb_nonzero_ws_bytes(n)) { set_error("frb_nonzero_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_select_count(a, n, dtype_width(dtype), NonzeroPred(), (u64*)ws, st);
  if (rc) return rc;
  return scan_exclusive_u64((u64*)ws, (size_t)compact_blocks(n), &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_nonzero_emit(const void* a, size_t n, int dtype, uint64_t* labels_out, uint64_t* index_out, size_t capacity,
                     const void* ws, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !ws || (capacity && (!labels_out || !index_out))) { set_error("frb_nonzero_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (capacity == 0) return FRB_OK;
  const int w = dtype_width(dtype);
  LabelIndexEmit emit;
  emit.labels_out = (u64*)labels_out;
  emit.index_out = (u64*)index_out;
  emit.flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  return launch_select_emit(a, n, w, NonzeroPred(), (const u64*)ws, emit, capacity, (cudaStream_t)stream);
}

// stable sort of n (key, value) uint64 pairs by the low key_bits bits of the key (many-launch LSD
// radix sort: lists as long as the volume).  workspace: 2 ping-pong lists + histograms
size_t frb_sort_pairs_u64_ws_bytes(size_t n) {
  const size_t c = n < 1 ? 1 : n;
  return 2 * align_up(c * sizeof(u64), 256) + radix_hist_bytes(c);
}
int frb_sort_pairs_u64(uint64_t* keys, uint64_t* vals, size_t n, int key_bits, void* ws, size_t ws_bytes, void* stream) {
  if ((n && (!keys || !vals)) || !ws || ws_bytes < frb_sort_pairs_u64_ws_bytes(n) || key_bits < 1 || key_bits > 64) { set_error("frb_sort_pairs_u64: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n <= 1) return FRB_OK;
  const size_t lb = align_up(n * sizeof(u64), 256);
  return radix_sort_u64((u64*)keys, (u64*)vals, (u64*)ws, (u64*)((char*)ws + lb), n, 0, key_bits, (char*)ws + 2 * lb, (cudaStream_t)stream);
}

int frb_unravel_u16(const uint64_t* index, size_t n, int ndim, uint64_t d1, uint64_t d2, uint16_t* out, void* stream) {
  if ((n && (!index || !out)) || (ndim != 2 && ndim != 3) || d1 == 0 || (ndim == 3 && d2 == 0)) { set_error("frb_unravel_u16: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (n == 0) return FRB_OK;
  Grid g = grid_for(n, 256 * 4, 16);
  k_unravel_u16<<<g.blocks, g.threads, 0, (cudaStream_t)stream>>>((const u64*)index, n, ndim, d1, d2, out);
  return check_launch("k_unravel_u16");
}

// workspace: 2 ping-pong lists of L uint64 + the cooperative sort's histograms and barrier word
size_t frb_sort_pairs_ws_bytes(size_t L) {
  const size_t c = L < 1 ? 1 : L;
  return 2 * align_up(c * sizeof(u64), 256) + align_up(coop_sort_aux_bytes(coop_sort_grid()), 256);
}

int frb_sort_table_pairs(uint64_t* keys, uint64_t* vals, size_t L, const uint64_t* count_dev, int dtype, void* uniq_out,
                         uint64_t* vals_out, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !ws || ws_bytes < frb_sort_pairs_ws_bytes(L) || (L && (!keys || !vals || !uniq_out || !vals_out))) {
    set_error("frb_sort_table_pairs: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (L == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const size_t lb = align_up(L * sizeof(u64), 256);
  SortPairsArgs A;
  A.sort.keys = (u64*)keys;
  A.sort.pay = (u64*)vals;
  A.sort.keys_alt = (u64*)ws;
  A.sort.pay_alt = (u64*)((char*)ws + lb);
  A.sort.pay2 = nullptr;
  A.sort.pay2_alt = nullptr;
  A.sort.hist = (unsigned int*)((char*)ws + 2 * lb);
  A.sort.barrier = (unsigned int*)((char*)ws + 2 * lb + (size_t)2 * coop_sort_grid() * 256 * sizeof(unsigned int));
  A.n = L;
  A.n_dev = (const u64*)count_dev;
  A.flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  A.key_bits = 8 * w;
  A.uniq_out = uniq_out;
  A.vals_out = (u64*)vals_out;
  cudaMemsetAsync(A.sort.barrier, 0, 256, st);
  const void* kernel;
  switch (w) {
    case 1: kernel = (const void*)k_sort_emit_pairs<1>; break;
    case 2: kernel = (const void*)k_sort_emit_pairs<2>; break;
    case 4: kernel = (const void*)k_sort_emit_pairs<4>; break;
    default: kernel = (const void*)k_sort_emit_pairs<8>; break;
  }
  void* params[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(coop_sort_grid()), dim3(CS_THREADS), params, 0, st);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return FRB_ERR_CUDA;
  }
  return check_launch("k_sort_emit_pairs");
}

// workspace: sorted keys (n * kw) | ping-pong (n * kw) | block counts | radix histograms;  kw = 4 (W<=4) or 8
static inline size_t usort_key_bytes(int dtype) { return dtype_width(dtype) == 8 ? 8 : 4; }
size_t frb_unique_sort_ws_bytes(size_t n, int dtype) {
  if (!dtype_ok(dtype)) return 0;
  const size_t c = n < 1 ? 1 : n;
  return 2 * align_up(c * usort_key_bytes(dtype), 256) + align_up((size_t)compact_blocks(c) * sizeof(u64), 256) + radix_hist_bytes(c);
}

int frb_unique_sort_count(const void* a, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, void* stream) {
  if (!dtype_ok(dtype) || !a || n == 0 || !meta || !ws || ws_bytes < frb_unique_sort_ws_bytes(n, dtype)) { set_error("frb_unique_sort_count: bad arguments"); return FRB_ERR_BAD_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const size_t kb = usort_key_bytes(dtype);
  const size_t lb = align_up(n * kb, 256);
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  char* base = (char*)ws;
  u64* block_counts = (u64*)(base + 2 * lb);
  void* hist = base + 2 * lb + align_up((size_t)blocks * sizeof(u64), 256);
  const u64 flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  Grid g = grid_for(n, 256 * 8, 16);
  int rc;
  if (kb == 8) {
    u64* keys = (u64*)base;
    u64* alt = (u64*)(base + lb);
    k_widen_flip<8, u64><<<g.blocks, g.threads, 0, st>>>(a, n, keys, flip);
    if ((rc = check_launch("k_widen_flip"))) return rc;
    if ((rc = radix_sort_u64(keys, nullptr, alt, nullptr, n, 0, 64, hist, st))) return rc;
    k_heads_count<u64><<<blocks, 256, 0, st>>>(keys, n, chunk, block_counts);
  } else {
    uint32_t* keys = (uint32_t*)base;
    uint32_t* alt = (uint32_t*)(base + lb);
    switch (w) {
      case 1: k_widen_flip<1, uint32_t><<<g.blocks, g.threads, 0, st>>>(a, n, keys, (uint32_t)flip); break;
      case 2: k_widen_flip<2, uint32_t><<<g.blocks, g.threads, 0, st>>>(a, n, keys, (uint32_t)flip); break;
      default: k_widen_flip<4, uint32_t><<<g.blocks, g.threads, 0, st>>>(a, n, keys, (uint32_t)flip); break;
    }
    if ((rc = check_launch("k_widen_flip"))) return rc;
    if ((rc = radix_sort_u32(keys, alt, n, 0, 8 * w, hist, st))) return rc;
    k_heads_count<uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_counts);
  }
  if ((rc = check_launch("k_heads_count"))) return rc;
  return scan_exclusive_u64(block_counts, (size_t)blocks, &((u64*)meta)[FRB_META_LIST_COUNT], st);
}

int frb_unique_sort_emit(size_t n, int dtype, void* uniq_out, uint64_t* counts_out, size_t capacity, const void* ws, void* stream) {
  if (!dtype_ok(dtype) || n == 0 || !ws || (capacity && (!uniq_out || !counts_out))) { set_error("frb_unique_sort_emit: bad arguments"); return FRB_ERR_BAD_ARG; }
  if (capacity == 0) return FRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int w = dtype_width(dtype);
  const size_t kb = usort_key_bytes(dtype);
  const size_t lb = align_up(n * kb, 256);
  const int blocks = compact_blocks(n);
  const size_t chunk = compact_chunk(n, blocks);
  const char* base = (const char*)ws;
  const u64* block_offs = (const u64*)(base + 2 * lb);
  const u64 flip = dtype_signed(dtype) ? (1ull << (8 * w - 1)) : 0ull;
  cudaMemsetAsync(counts_out, 0, capacity * sizeof(u64), st);
  u64* co = (u64*)counts_out;
  if (kb == 8) {
    k_heads_emit<8, u64><<<blocks, 256, 0, st>>>((const u64*)base, n, chunk, block_offs, flip, uniq_out, co, capacity);
  } else {
    const uint32_t* keys = (const uint32_t*)base;
    switch (w) {
      case 1: k_heads_emit<1, uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_offs, (uint32_t)flip, uniq_out, co, capacity); break;
      case 2: k_heads_emit<2, uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_offs, (uint32_t)flip, uniq_out, co, capacity); break;
      default: k_heads_emit<4, uint32_t><<<blocks, 256, 0, st>>>(keys, n, chunk, block_offs, (uint32_t)flip, uniq_out, co, capacity); break;
    }
  }
  return check_launch("k_heads_emit");
}

}  // extern "C"
