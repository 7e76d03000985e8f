/*
 * fastremap_oracle.c -- CPU restatement of the reference's relabel hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker the CUDA path is compared
 * against; it is never shipped, never imported by fastremap_b200/, and never the
 * thing that is measured as the product.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It restates, in plain sequential C, the algorithms of the reference's typed inner
 * loops (reference = seung-lab/fastremap, file fastremap/fastremap.pyx; line numbers
 * below refer to that file).  Parity of this restatement is PINNED: tests/test_oracle.py
 * checks it against the golden vectors of the reference's own test-suite
 * (automated_test.py) and against fixtures produced by the compiled, unmodified
 * reference (tests/golden/, generator script committed beside them).
 *
 * Design: the reference instantiates every loop once per dtype (Cython fused types).
 * Here every loop is written once over "raw element bits": elements are loaded as
 * 64-bit patterns (zero-extended) by item width, which is all that equality, hashing
 * and re-storing need; the few places where ORDER matters (minmax, unique, the
 * remap_from_array bound check) take an is_signed flag.  Storing at the item width
 * truncates, which reproduces the reference's "counter arithmetic is in the array
 * dtype" wrap-around (fastremap.pyx:225 `cdef NUMBER remap_id = start`).
 *
 * The hash map is a plain open-addressing table (linear probing, power-of-two size,
 * load <= 0.5, separate occupancy byte so that every 64-bit pattern is a legal key).
 * The reference uses ska::flat_hash_map (robin hood); only the key->value association
 * is observable, so any correct map restates it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ elements */

static inline uint64_t ld(const void *p, int w, size_t i) {
  switch (w) {
    case 1: return ((const uint8_t *)p)[i];
    case 2: return ((const uint16_t *)p)[i];
    case 4: return ((const uint32_t *)p)[i];
    default: return ((const uint64_t *)p)[i];
  }
}
static inline void st(void *p, int w, size_t i, uint64_t v) {
  switch (w) {
    case 1: ((uint8_t *)p)[i] = (uint8_t)v; break;
    case 2: ((uint16_t *)p)[i] = (uint16_t)v; break;
    case 4: ((uint32_t *)p)[i] = (uint32_t)v; break;
    default: ((uint64_t *)p)[i] = v; break;
  }
}
static inline uint64_t trunc_w(uint64_t v, int w) {
  return w == 8 ? v : (v & ((1ull << (8 * w)) - 1));
}
static inline int64_t sext(uint64_t v, int w) {
  switch (w) {
    case 1: return (int8_t)v;
    case 2: return (int16_t)v;
    case 4: return (int32_t)v;
    default: return (int64_t)v;
  }
}
/* a < b in the dtype's own ordering */
static inline int lt(uint64_t a, uint64_t b, int w, int is_signed) {
  return is_signed ? (sext(a, w) < sext(b, w)) : (a < b);
}

/* ------------------------------------------------------------------ hash map */

typedef struct {
  uint64_t *keys, *vals;
  uint8_t *used;
  size_t cap, count; /* cap is a power of two */
} omap;

static int omap_init(omap *m, size_t cap) {
  size_t c = 16;
  while (c < cap) c <<= 1;
  m->cap = c;
  m->count = 0;
  m->keys = (uint64_t *)malloc(c * sizeof(uint64_t));
  m->vals = (uint64_t *)malloc(c * sizeof(uint64_t));
  m->used = (uint8_t *)calloc(c, 1);
  return (m->keys && m->vals && m->used) ? 0 : -1;
}
static void omap_free(omap *m) {
  free(m->keys); free(m->vals); free(m->used);
  m->keys = m->vals = NULL; m->used = NULL;
}
static inline size_t omap_h(uint64_t k, size_t cap) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (size_t)k & (cap - 1);
}
/* returns slot index or (size_t)-1 */
static inline size_t omap_find(const omap *m, uint64_t k) {
  size_t s = omap_h(k, m->cap);
  while (m->used[s]) {
    if (m->keys[s] == k) return s;
    s = (s + 1) & (m->cap - 1);
  }
  return (size_t)-1;
}
static int omap_put(omap *m, uint64_t k, uint64_t v);
static int omap_grow(omap *m) {
  omap n;
  if (omap_init(&n, m->cap * 2)) return -1;
  for (size_t i = 0; i < m->cap; i++)
    if (m->used[i]) omap_put(&n, m->keys[i], m->vals[i]);
  omap_free(m);
  *m = n;
  return 0;
}
/* insert or overwrite (operator[] = v) */
static int omap_put(omap *m, uint64_t k, uint64_t v) {
  if ((m->count + 1) * 2 > m->cap && omap_grow(m)) return -1;
  size_t s = omap_h(k, m->cap);
  while (m->used[s]) {
    if (m->keys[s] == k) { m->vals[s] = v; return 0; }
    s = (s + 1) & (m->cap - 1);
  }
  m->used[s] = 1; m->keys[s] = k; m->vals[s] = v; m->count++;
  return 0;
}

/* ------------------------------------------------------------------ scans */

/* fastremap.pyx:77-93 `_minmax`: one pass, strict comparisons, first element seeds both. */
int orc_minmax(const void *arr, size_t n, int w, int is_signed, uint64_t *out_min, uint64_t *out_max) {
  if (n == 0) return 1; /* reference returns (None, None) */
  uint64_t mn = ld(arr, w, 0), mx = mn;
  for (size_t i = 1; i < n; i++) {
    uint64_t v = ld(arr, w, i);
    if (lt(v, mn, w, is_signed)) mn = v;
    if (lt(mx, v, w, is_signed)) mx = v;
  }
  *out_min = mn; *out_max = mx;
  return 0;
}

/* fastremap.pyx:840-853 `_pixel_pairs`: number of i>=1 with a[i]==a[i-1] in memory order. */
uint64_t orc_pixel_pairs(const void *arr, size_t n, int w) {
  if (n == 0) return 0;
  uint64_t pairs = 0, label = ld(arr, w, 0);
  for (size_t i = 1; i < n; i++) {
    uint64_t v = ld(arr, w, i);
    if (v == label) pairs++; else label = v;
  }
  return pairs;
}

/* ------------------------------------------------------------------ renumber */

/*
 * fastremap.pyx:196-257 `_renumber`.
 *   :220-221  preserve_zero seeds the map with 0 -> 0
 *   :225      remap_id = (NUMBER)start        (counter lives in the array dtype -> wraps)
 *   :230      last_elem = ~arr[0]             (first voxel never takes the run shortcut)
 *   :236-251  run shortcut / find / insert-with-next-id
 *   :253      factor = remap_id (the NEXT unused id) -- returned through *next_id
 * Writes the new ids in place at the input width.  keys_out/vals_out (caller-allocated,
 * capacity n+1 entries of 8 bytes) receive the map in insertion order, as raw bits at
 * the array width.  Returns the number of map entries (including the seeded 0->0).
 */
int64_t orc_renumber(void *arr, size_t n, int w, int64_t start, int preserve_zero,
                     uint64_t *keys_out, uint64_t *vals_out, uint64_t *next_id) {
  omap m;
  if (omap_init(&m, 2048)) return -1;
  size_t nk = 0;
  if (preserve_zero) { omap_put(&m, 0, 0); keys_out[nk] = 0; vals_out[nk] = 0; nk++; }
  uint64_t remap_id = trunc_w((uint64_t)start, w);
  if (n == 0) { *next_id = remap_id; omap_free(&m); return (int64_t)nk; }
  uint64_t last_elem = trunc_w(~ld(arr, w, 0), w);
  uint64_t last_remap_id = remap_id;
  for (size_t i = 0; i < n; i++) {
    uint64_t elem = ld(arr, w, i);
    if (elem == last_elem) { st(arr, w, i, last_remap_id); continue; }
    size_t s = omap_find(&m, elem);
    uint64_t id;
    if (s == (size_t)-1) {
      id = remap_id;
      if (omap_put(&m, elem, id)) { omap_free(&m); return -1; }
      keys_out[nk] = elem; vals_out[nk] = id; nk++;
      remap_id = trunc_w(remap_id + 1, w);
    } else {
      id = m.vals[s];
    }
    st(arr, w, i, id);
    last_elem = elem;
    last_remap_id = id;
  }
  *next_id = remap_id;
  omap_free(&m);
  return (int64_t)nk;
}

/* ------------------------------------------------------------------ remap */

/*
 * fastremap.pyx:705-769 `_remap`.
 *   :721-729  single-entry + preserve fast path (compare and replace)
 *   :733-734  table -> map (later duplicates overwrite; a dict has none)
 *   :740-747  first voxel handled before the loop
 *   :749-767  run shortcut; miss -> keep (preserve) or KeyError naming the label
 * Return 0, or 1 on a missing label with *missing_index / *missing_label set; in that
 * case the prefix [0, missing_index) has already been rewritten, exactly like the
 * reference leaves it.
 */
int orc_remap(void *arr, size_t n, int w, const uint64_t *keys, const uint64_t *vals, size_t nkv,
              int preserve_missing, uint64_t *missing_index, uint64_t *missing_label) {
  if (n == 0) return 0;
  if (preserve_missing && nkv == 1) {
    uint64_t before = trunc_w(keys[0], w), after = trunc_w(vals[0], w);
    if (before == after) return 0;
    for (size_t i = 0; i < n; i++)
      if (ld(arr, w, i) == before) st(arr, w, i, after);
    return 0;
  }
  omap m;
  if (omap_init(&m, nkv * 2 + 16)) return -1;
  for (size_t j = 0; j < nkv; j++) omap_put(&m, trunc_w(keys[j], w), trunc_w(vals[j], w));

  uint64_t last_elem = ld(arr, w, 0), last_remap_id = 0;
  size_t s = omap_find(&m, last_elem);
  if (s == (size_t)-1) {
    if (!preserve_missing) { *missing_index = 0; *missing_label = last_elem; omap_free(&m); return 1; }
    last_remap_id = last_elem;
  } else {
    st(arr, w, 0, m.vals[s]);
    last_remap_id = m.vals[s];
  }
  for (size_t i = 1; i < n; i++) {
    uint64_t elem = ld(arr, w, i);
    if (elem == last_elem) { st(arr, w, i, last_remap_id); continue; }
    s = omap_find(&m, elem);
    if (s == (size_t)-1) {
      if (preserve_missing) { last_elem = elem; last_remap_id = elem; continue; }
      *missing_index = i; *missing_label = elem; omap_free(&m); return 1;
    }
    st(arr, w, i, m.vals[s]);
    last_elem = elem;
    last_remap_id = m.vals[s];
  }
  omap_free(&m);
  return 0;
}

/* fastremap.pyx:492-527 `_mask_except`: keep listed labels, everything else := value. */
int orc_mask_except(void *arr, size_t n, int w, const uint64_t *labels, size_t nlabels, uint64_t value) {
  if (n == 0) return 0;
  omap m;
  if (omap_init(&m, nlabels * 2 + 16)) return -1;
  for (size_t j = 0; j < nlabels; j++) omap_put(&m, trunc_w(labels[j], w), trunc_w(labels[j], w));
  value = trunc_w(value, w);
  uint64_t last_elem = ld(arr, w, 0);
  uint64_t last_elem_value = (omap_find(&m, last_elem) == (size_t)-1) ? value : last_elem;
  for (size_t i = 0; i < n; i++) {
    uint64_t v = ld(arr, w, i);
    if (v == last_elem) {
      st(arr, w, i, last_elem_value);
    } else if (omap_find(&m, v) == (size_t)-1) {
      last_elem = v; last_elem_value = value; st(arr, w, i, value);
    } else {
      last_elem = v; last_elem_value = v;
    }
  }
  omap_free(&m);
  return 0;
}

/*
 * fastremap.pyx:771-791 `remap_from_array`: arr[i] = vals[arr[i]] unless arr[i] > len(vals)-1.
 * Unsigned dtypes only.  The reference computes maxkey = vals.size - 1 in size_t, so
 * nvals == 0 reads out of bounds there; this restatement treats nvals == 0 as a no-op
 * ("do not reproduce").
 */
void orc_remap_from_array(void *arr, size_t n, int w, const void *vals, size_t nvals) {
  if (nvals == 0) return;
  uint64_t maxkey = (uint64_t)nvals - 1;
  for (size_t i = 0; i < n; i++) {
    uint64_t e = ld(arr, w, i);
    if (e > maxkey) continue;
    st(arr, w, i, ld(vals, w, (size_t)e));
  }
}

/*
 * fastremap.pyx:793-827 `remap_from_array_kv`: map built from keys[i] -> vals[i] in order
 * (later duplicates win, :810-811); every voxel looked up (no run shortcut, :817-825);
 * miss -> keep, or KeyError naming the label (:823).
 */
int orc_remap_from_array_kv(void *arr, size_t n, int w, const void *keys, const void *vals, size_t nkv,
                            int preserve_missing, uint64_t *missing_index, uint64_t *missing_label) {
  omap m;
  if (omap_init(&m, nkv * 2 + 16)) return -1;
  for (size_t j = 0; j < nkv; j++) omap_put(&m, ld(keys, w, j), ld(vals, w, j));
  for (size_t i = 0; i < n; i++) {
    uint64_t e = ld(arr, w, i);
    size_t s = omap_find(&m, e);
    if (s == (size_t)-1) {
      if (preserve_missing) continue;
      *missing_index = i; *missing_label = e; omap_free(&m); return 1;
    }
    st(arr, w, i, m.vals[s]);
  }
  omap_free(&m);
  return 0;
}

/* ------------------------------------------------------------------ component_map */

/*
 * fastremap.pyx:565-587 `_component_map`: a store remap[cc[i]] = parent[i] at every run
 * start of cc in memory order; the last store per key wins; dict insertion order is
 * first-occurrence order.  keys_out / vals_out (capacity n) receive the entries in
 * first-occurrence order as raw bits at widths wc / wp.  Returns the entry count.
 */
int64_t orc_component_map(const void *cc, int wc, const void *parent, int wp, size_t n,
                          uint64_t *keys_out, uint64_t *vals_out) {
  if (n == 0) return 0;
  omap m; /* key -> position in keys_out */
  if (omap_init(&m, 2048)) return -1;
  size_t nk = 0;
  uint64_t last_label = ld(cc, wc, 0);
  omap_put(&m, last_label, 0); keys_out[0] = last_label; vals_out[0] = ld(parent, wp, 0); nk = 1;
  for (size_t i = 0; i < n; i++) {
    uint64_t c = ld(cc, wc, i);
    if (c == last_label) continue;
    size_t s = omap_find(&m, c);
    if (s == (size_t)-1) {
      if (omap_put(&m, c, nk)) { omap_free(&m); return -1; }
      keys_out[nk] = c; vals_out[nk] = ld(parent, wp, i); nk++;
    } else {
      vals_out[m.vals[s]] = ld(parent, wp, i);
    }
    last_label = c;
  }
  omap_free(&m);
  return (int64_t)nk;
}

/* ------------------------------------------------------------------ unique */

/*
 * fastremap.pyx:1061-1138 `_unique_via_array`: dense histogram over [0, max_label]
 * (:1085-1086), optional first-index table seeded with the uintp-max sentinel (:1077-1091),
 * ascending compaction of the non-empty bins (:1093-1119), optional inverse through a
 * label -> position table (:1123-1130).
 * counts / index are caller-zeroed / caller-sentinelled scratch of max_label+1 entries.
 * uniq_out (raw bits), cts_out, idx_out have capacity max_label+1; inverse_out capacity n.
 * idx_out / inverse_out may be NULL.  Returns the number of uniques.
 */
int64_t orc_unique_via_array(const void *labels, size_t n, int w, uint64_t max_label,
                             uint64_t *counts, uint64_t *index,
                             uint64_t *uniq_out, uint64_t *cts_out, uint64_t *idx_out, uint64_t *inverse_out) {
  for (size_t i = 0; i < n; i++) counts[ld(labels, w, i)] += 1;
  if (idx_out) {
    for (size_t i = 0; i < n; i++) {
      uint64_t l = ld(labels, w, i);
      if (index[l] == UINT64_MAX) index[l] = i;
    }
  }
  size_t j = 0;
  for (uint64_t b = 0; b <= max_label; b++) {
    if (counts[b] > 0) {
      uniq_out[j] = b; cts_out[j] = counts[b];
      if (idx_out) idx_out[j] = index[b];
      j++;
    }
    if (b == UINT64_MAX) break;
  }
  if (inverse_out && j) {
    /* reuse `counts` as the label -> position table (the reference allocates `mapping`) */
    for (size_t k = 0; k < j; k++) counts[uniq_out[k]] = k;
    for (size_t i = 0; i < n; i++) inverse_out[i] = counts[ld(labels, w, i)];
  }
  return (int64_t)j;
}

/*
 * fastremap.pyx:1024-1059 `_unique_via_sort`, the run-length-encode half (:1042-1054).
 * `sorted` must already be sorted in the dtype's order (the reference calls ndarray.sort()
 * at :1030; the Python wrapper of this oracle does the same with numpy).
 * uniq_out / cts_out have capacity n.  Returns the number of uniques.
 */
int64_t orc_rle_sorted(const void *sorted, size_t n, int w, uint64_t *uniq_out, uint64_t *cts_out) {
  if (n == 0) return 0;
  size_t j = 0;
  uint64_t cur = ld(sorted, w, 0), accum = 1;
  for (size_t i = 1; i < n; i++) {
    uint64_t v = ld(sorted, w, i);
    if (v == cur) { accum++; }
    else { uniq_out[j] = cur; cts_out[j] = accum; j++; accum = 1; cur = v; }
  }
  uniq_out[j] = cur; cts_out[j] = accum; j++;
  return (int64_t)j;
}
