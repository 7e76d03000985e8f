/*
 * fastremap_b200.h -- C ABI of the B200-native relabel hot path.
 *
 * The reference (seung-lab/fastremap) has no FFI: its boundary is the Python module
 * surface (fastremap/__init__.py:1-24, signatures in fastremap/fastremap.pyi).  The host
 * side of this library (fastremap_b200/api.py) keeps that surface verbatim; the entry
 * points below are what it binds with ctypes, and each one names the typed inner loop of
 * fastremap/fastremap.pyx that it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller; the library never allocates
 *    device memory (workspace sizes come from the *_bytes() queries);
 *  - volume pointers (a, src, dst*) must be 16-byte aligned; n is in elements (voxels);
 *  - arrays are walked in MEMORY order (the caller flattens C or F arrays by stride);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - return value: FRB_OK or an FRB_ERR_* code for argument / launch errors.  Data
 *    dependent outcomes (table overflow, missing label) are reported through the table's
 *    `meta` block, which the caller reads back after synchronising;
 *  - dtype codes are frb_dtype; "raw bits" means the element zero-extended to 64 bits.
 */
#ifndef FASTREMAP_B200_H
#define FASTREMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FRB_U8 = 0, FRB_U16 = 1, FRB_U32 = 2, FRB_U64 = 3,
  FRB_I8 = 4, FRB_I16 = 5, FRB_I32 = 6, FRB_I64 = 7
} frb_dtype;

enum { FRB_OK = 0, FRB_ERR_BAD_ARG = 4, FRB_ERR_CUDA = 3 };

/* ---- label table ---------------------------------------------------------------------
 * Replaces ska::flat_hash_map<T,T> (fastremap/ska_flat_hash_map.hpp:1373-1490) as used by
 * _renumber / _remap / _mask_except / remap_from_array_kv.
 * A table is `cap` slots of 16 bytes {uint64 key bits, uint64 value} plus a meta block of
 * FRB_META_WORDS uint64.  cap is a power of two.  Tables for 8/16-bit dtypes are DIRECT
 * (slot index == key, cap == 2^bits); wider dtypes use open addressing with linear probing from an
 * EVEN home slot, so the first two probes of a lookup are one 32-byte sector (one 256-bit load).
 * Tables that are only looked up (remap / mask / mask_except) are sized for a load factor <= 1/8:
 * on this workload a lookup's cost is the length of its dependent probe chain, not the table's size.
 * The all-ones key pattern is the EMPTY marker; a real all-ones key lives in the meta block.
 */
#define FRB_META_WORDS 16
enum {
  FRB_META_COUNT = 0,        /* number of distinct keys inserted                          */
  FRB_META_OVERFLOW = 1,     /* !=0: a probe sequence exceeded the limit -> grow and retry */
  FRB_META_SIDE_PRESENT = 2, /* the all-ones key is present                               */
  FRB_META_SIDE_VAL = 3,     /* ... and this is its value                                 */
  FRB_META_MISSING_INDEX = 4,/* smallest flat index whose label was not in the table      */
  FRB_META_LIST_COUNT = 5,   /* entries written by export / rank / unique                 */
  FRB_META_ASSIGNED = 6,     /* renumber: labels ranked so far (next rank to hand out)    */
  FRB_META_AUX1 = 7          /* component_map: slots appended to the build's slot list   */
};
/* renumber table values: FRB_RANK_TAG | first flat index while a label is unranked, its 0-based
 * rank in encounter order afterwards (ranks compare below every tagged index). */
#define FRB_RANK_TAG 0x8000000000000000ull
enum { FRB_OP_MIN = 0, FRB_OP_MAX = 1, FRB_OP_ADD = 2, FRB_OP_SET = 3 };

size_t frb_table_bytes(uint64_t cap);                 /* = cap * 16 */
uint64_t frb_table_direct_cap(int dtype);             /* 256 / 65536 for 8/16-bit dtypes, else 0 */
/* keys := EMPTY, values := init_val, meta := 0 (MISSING_INDEX := all-ones, SIDE_VAL := init_val) */
int frb_table_init(void* slots, uint64_t cap, uint64_t* meta, uint64_t init_val, void* stream);
/* insert L (key, value) pairs, combining values of equal keys with `op`.
 * keys: elements of width key_width bytes (1,2,4,8).  vals: width val_width, or NULL to use
 * the pair's position (0..L-1) + val_offset as the value.  direct != 0 selects a DIRECT table. */
int frb_table_insert(const void* keys, int key_width, const void* vals, int val_width, uint64_t val_offset,
                     size_t L, int op, int direct, void* slots, uint64_t cap, uint64_t* meta, void* stream);
/* the multi-GPU key-table merge in one launch: `gathered` is the result of an all-gather of per-rank
 * buffers [m uint64 keys | m uint64 values]; metas is the all-gathered meta blocks (world x
 * FRB_META_WORDS, device memory) whose COUNT word says how many pairs of each segment are valid. */
int frb_table_insert_gathered(const uint64_t* gathered, uint64_t m, int world, const uint64_t* metas, int op, int direct,
                              void* slots, uint64_t cap, uint64_t* meta, void* stream);
/* The same merge when the per-rank capacity m was fixed BEFORE the counts were known (one collective,
 * no host round trip): packed holds, per rank, [FRB_META_WORDS words: the rank's table meta | m keys |
 * m values] (the exporter dropped pairs beyond m).  If a rank's COUNT exceeds m, its OVERFLOW word is
 * set, or 2 * COUNT > local_cap (local_cap != 0: the load limit of the per-rank tables), nothing is
 * inserted and meta[OVERFLOW] of the destination is raised: kernels queued behind (frb_renumber_rank,
 * frb_relabel with guard_overflow, frb_relabel_auto) do nothing and the host takes the two-phase path. */
int frb_table_insert_packed(const uint64_t* packed, uint64_t m, int world, uint64_t local_cap, int op, int direct, void* slots,
                            uint64_t cap, uint64_t* meta, void* stream);
/* dst[key] = src[key] for every key of the list that the hashed table src holds (values copied as they are; dst must be
 * initialised, sized for max_count keys).  count_dev != NULL: the list holds min(*count_dev, max_count) keys, the count
 * still being on the device.  Copies meta[FRB_META_ASSIGNED]; a raised OVERFLOW flag in src_meta is passed on instead.
 * Multi-GPU renumber: a rank's own labels with their global ranks, as a compact lookup table for K3. */
int frb_table_project(const uint64_t* keys, const uint64_t* count_dev, uint64_t max_count, const void* src_slots, uint64_t src_cap,
                      const uint64_t* src_meta, void* dst_slots, uint64_t dst_cap, uint64_t* dst_meta, void* stream);
/* replace every stored value v (a position) by vals[v] -- "later duplicate wins" for
 * remap_from_array_kv (fastremap.pyx:810-811) after an FRB_OP_MAX insert of positions. */
int frb_table_resolve(const void* vals, int val_width, void* slots, uint64_t cap, uint64_t* meta, void* stream);
/* write all (key bits, value) pairs to keys_out/vals_out (uint64 each, capacity `capacity`);
 * meta[FRB_META_LIST_COUNT] receives the number of pairs (unordered). */
int frb_table_export(const void* slots, uint64_t cap, uint64_t* meta, uint64_t* keys_out, uint64_t* vals_out,
                     size_t capacity, void* stream);

/* ---- scans ---------------------------------------------------------------------------
 * Replaces _minmax (fastremap.pyx:77-93) and _pixel_pairs (:840-853) with ONE fused pass.
 * out[0]=min, out[1]=max (raw bits sign-extended to 64 for signed dtypes), out[2]=#(a[i]==a[i-1]),
 * out[3]=#(a[i]!=0).  out is 4 uint64 on the device, initialised by the call. */
int frb_minmax_pairs(const void* a, size_t n, int dtype, uint64_t* out, void* stream);

/* ---- renumber ------------------------------------------------------------------------
 * K1 replaces the find/insert half of _renumber's loop (fastremap.pyx:236-251): for every
 * run head a[i] != a[i-1] the table receives key a[i] with value min(i + flat_offset).
 * skip_zero != 0 (preserve_zero) leaves label 0 out of the table.
 * row_bytes: length in bytes of the volume's fastest axis, or 0 if unknown -- a performance hint only
 * (the "same as the voxels directly above" filter can be aimed exactly when it is 4, 8, 12 or 16 KiB). */
int frb_renumber_build(const void* a, size_t n, int dtype, uint64_t flat_offset, int skip_zero,
                       void* slots, uint64_t cap, uint64_t* meta, uint64_t row_bytes, void* stream);
/* K2 replaces the implicit `remap_id += 1` ordering (fastremap.pyx:244-246).  Every label whose
 * table value is still a tagged first index is ranked by that index; the label with the r-th
 * smallest index gets rank meta[FRB_META_ASSIGNED] + r as its table value, and ASSIGNED advances.
 * Ranking is incremental: build + rank may be repeated chunk by chunk in memory order (index_base =
 * flat index of the chunk's first voxel, index_count = voxels in the chunk; only labels whose first
 * index lies in [index_base, index_base + index_count) are ranked).
 * keys_out / ids_out (elements of the array dtype, `capacity` entries, indexed by rank) receive
 * label and id = (T)(start + rank) (arithmetic wraps in the array dtype, :225).  The count of newly
 * ranked labels is left in meta[FRB_META_LIST_COUNT]; nothing is read back by the host.  If the
 * table overflowed during the build, or capacity is too small, nothing is ranked and
 * meta[FRB_META_OVERFLOW] is (left) set. */
size_t frb_renumber_rank_ws_bytes(size_t capacity);
int frb_renumber_rank(void* slots, uint64_t cap, uint64_t* meta, int dtype, int64_t start, uint64_t index_base,
                      uint64_t index_count, void* keys_out, void* ids_out, size_t capacity, void* ws, size_t ws_bytes,
                      void* stream);

/* ---- relabel -------------------------------------------------------------------------
 * K3 replaces the write half of _renumber (fastremap.pyx:240/:244/:248), _remap (:749-767),
 * _mask_except (:516-525) and remap_from_array_kv (:817-825), with the refit astype
 * (:301) fused in as an optional narrow store.
 *   dst_wide   (nullable) receives results at the input width (may alias src: in place)
 *   dst_narrow (nullable) receives results truncated to narrow_width bytes (<= input width; or 8 with
 *   dst_wide NULL: zero-extended 8-byte results from narrower labels, e.g. positions for unique's inverse)
 *   policy: FRB_MISS_KEEP   missing label stays as it is           (preserve_missing_labels)
 *           FRB_MISS_ERROR  missing label: meta[MISSING_INDEX] = min(flat index); value kept
 *           FRB_MISS_FILL   missing label := fill; present labels stay as they are (mask_except)
 *   zero_fixed != 0: label 0 maps to 0 without a lookup (renumber's preserve_zero).
 *   add: added to every value found in the table (renumber: table value = rank, add = start).
 *   guard_overflow != 0: do nothing if meta[FRB_META_OVERFLOW] is set (chunk pipelines launch the
 *   relabel before the host has seen whether the build of that chunk overflowed).
 *   batch_rows: 0 / 1 = look up row by row (a table whose live part sits in L1: renumber); 2 or 4 =
 *   issue the table loads of that many rows of a warp-tile before consuming any (tables of 10^5+
 *   labels that live in L2 / HBM; wide destination only).
 *   plane_bytes: bytes per plane of the two fastest axes or 0 (performance hint for the batched form: walk the
 *   volume band by band through stacks of planes, so that table sectors are still in L2 one plane deeper). */
enum { FRB_MISS_KEEP = 0, FRB_MISS_ERROR = 1, FRB_MISS_FILL = 2 };
int frb_relabel(const void* src, void* dst_wide, void* dst_narrow, int narrow_width, size_t n, int dtype,
                const void* slots, uint64_t cap, int direct, uint64_t* meta, int policy, uint64_t fill,
                int zero_fixed, uint64_t add, int guard_overflow, int batch_rows, uint64_t plane_bytes, void* stream);
/* renumber's K3 queued straight behind K2: the refit width is chosen ON THE DEVICE from the number of
 * labels nl = meta[FRB_META_ASSIGNED]: 1 byte if nl < lim1, else 2 if nl < lim2, else 4 if nl < lim4
 * (widths >= the input width are ignored), else no narrow copy.  The host derives lim* from fit_dtype
 * and reads nl afterwards to know which it was.  dst_narrow must hold n * (input width / 2) bytes.
 * wide_always != 0: ids are also written at the input width (in_place); otherwise only when nothing
 * narrower fits.  Does nothing if meta[FRB_META_OVERFLOW] is set.  Policy: missing label = error. */
int frb_relabel_auto(const void* src, void* dst_wide, void* dst_narrow, size_t n, int dtype, const void* slots, uint64_t cap,
                     int direct, uint64_t* meta, int zero_fixed, uint64_t add, int wide_always, uint64_t lim1, uint64_t lim2,
                     uint64_t lim4, void* stream);
/* single-pair compare-and-replace, the :721-729 fast path of _remap */
int frb_replace_one(const void* src, void* dst, size_t n, int dtype, uint64_t before, uint64_t after, void* stream);
/* remap_from_array (fastremap.pyx:771-791): dst[i] = src[i] < nvals ? vals[src[i]] : src[i] (unsigned dtypes) */
int frb_remap_from_array(const void* src, void* dst, size_t n, const void* vals, size_t nvals, int dtype, void* stream);
/* exchange the low and high halves of n elements of `width` bytes (2, 4, 8): the (N,2) edge-list
 * packing of _two_axis_unique (fastremap.pyx:967-982), src may equal dst */
int frb_swap_halves(const void* src, void* dst, size_t n, int width, void* stream);
/* out[i] = (hi[i] << 8*lo_width) | lo[i] (raw bits, widths 1 / 2 / 4 bytes) as uint64: a voxel's
 * (parent, component) pair as one label -- inverse_component_map (fastremap.pyx:589-652) is then the
 * `unique` of the packed volume. */
int frb_pack_pairs(const void* hi, int hi_width, const void* lo, int lo_width, uint64_t* out, size_t n, void* stream);
/* refit's astype (fastremap.pyx:301): integer width change, truncating or sign/zero extending */
int frb_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, void* stream);

/* ---- memory layout: transposition and chunk serialisation ------------------------------
 * Replaces pyipt::_ipt2d / _ipt3d / _ipt4d (fastremap.pyx:61-68, ipt.hpp:330-352) behind transpose /
 * asfortranarray / ascontiguousarray (fastremap.pyx:1140-1273) and the strip loops of tobytes
 * (fastremap.pyx:1547-1654).  Out of place: `in` and `out` must not overlap.
 *
 * Batched 2-D transposition of `width`-byte elements (1, 2, 4, 8): for every batch index and
 * a < dim_a, b < dim_b:
 *   out[out_off + a * out_stride_a + b] = in[in_off + b * in_stride_b + a]
 * where in_off / out_off = sum over the nbatch (<= 6) batch axes of coordinate * stride.  All
 * strides in elements; batch_dims[0] is the fastest-varying batch axis of the tile order. */
int frb_transpose_tiles(const void* in, void* out, int width, uint64_t dim_a, uint64_t dim_b, uint64_t in_stride_b,
                        uint64_t out_stride_a, int nbatch, const uint64_t* batch_dims, const uint64_t* batch_in_strides,
                        const uint64_t* batch_out_strides, void* stream);
/* Batched copy of rows of row_bytes contiguous bytes; offsets as above but in BYTES. */
int frb_copy_rows(const void* in, void* out, uint64_t row_bytes, int nbatch, const uint64_t* batch_dims,
                  const uint64_t* batch_in_strides_bytes, const uint64_t* batch_out_strides_bytes, void* stream);

/* ---- unique(return_counts) -----------------------------------------------------------
 * Replaces _unique_via_array (fastremap.pyx:1061-1138), _unique_via_shifted_array (:984-1001),
 * _unique_via_renumber (:1003-1022) and _unique_via_sort (:1024-1059).
 * dense:  counts[a[i]-min] += 1 over `bins` uint64 bins (caller-zeroed), then an ordered
 *         compaction of the non-empty bins;
 * hash:   (key,count) table + sort of the distinct keys;
 * sort:   radix sort of all voxels + run-length encode.
 * row_bytes: fastest-axis length in bytes or 0 (performance hint, see frb_renumber_build). */
int frb_hist_dense(const void* a, size_t n, int dtype, uint64_t min_bits, uint64_t* counts, uint64_t bins, uint64_t row_bytes,
                   void* stream);
/* ordered compaction of the non-empty bins, in two steps because the caller sizes the outputs:
 * _count leaves the number of non-empty bins in meta[FRB_META_LIST_COUNT] (and per-CTA offsets in
 * ws); _emit writes label = min + bin (array dtype) and its count for every non-empty bin. */
size_t frb_compact_bins_ws_bytes(uint64_t bins);
int frb_compact_bins_count(const uint64_t* counts, uint64_t bins, uint64_t* meta, void* ws, size_t ws_bytes, void* stream);
int frb_compact_bins_emit(const uint64_t* counts, uint64_t bins, int dtype, uint64_t min_bits, void* uniq_out,
                          uint64_t* counts_out, size_t capacity, const void* ws, void* stream);
/* hash strategy (fastremap.pyx:1003-1022 builds the same label -> count map through renumber + a dense histogram).
 * The pass over the volume aggregates in registers and shared memory and APPENDS what it evicts to per-warp logs in
 * ws (frb_hist_ws_bytes(n) bytes; NULL: update the table from inside the pass instead -- slower); the logs are then
 * folded into the table (frb_hist_apply_log, called by frb_hist_hash itself).  If the table overflows and the flag
 * word at ws + frb_hist_ws_flags_offset() is 0, frb_hist_apply_log on a bigger table finishes the job without a
 * second pass over the volume.
 * plane_bytes: bytes per plane of the two fastest axes, or 0 (performance hint): when the volume is a stack of planes
 * of whole 16 KiB tiles, a CTA walks one band of rows through a stack of consecutive planes, so the segments it has
 * aggregated in shared memory come round again one plane deeper. */
size_t frb_hist_ws_bytes(size_t n);
size_t frb_hist_ws_flags_offset(void);
int frb_hist_hash(const void* a, size_t n, int dtype, void* slots, uint64_t cap, uint64_t* meta, uint64_t row_bytes,
                  uint64_t plane_bytes, void* ws, size_t ws_bytes, void* stream);
int frb_hist_apply_log(const void* ws, size_t ws_bytes, size_t n, int dtype, void* slots, uint64_t cap, uint64_t* meta, void* stream);
size_t frb_sort_pairs_ws_bytes(size_t L);
/* sort exported (key bits, value) pairs ascending in the dtype's own order (keys / vals are
 * clobbered) and emit typed keys + values.  L = number of pairs, or -- count_dev != NULL -- the
 * capacity of the lists whose actual length min(*count_dev, L) is still on the device (e.g.
 * &meta[FRB_META_LIST_COUNT] right after frb_table_export: no host round trip in between). */
int frb_sort_table_pairs(uint64_t* keys, uint64_t* vals, size_t L, const uint64_t* count_dev, int dtype, void* uniq_out,
                         uint64_t* vals_out, void* ws, size_t ws_bytes, void* stream);
/* direct tables (8- / 16-bit dtypes: slot index = label bits): the labels in the dtype's own order with their
 * values, by an ordered compaction of the slot array -- replaces frb_table_export + frb_sort_table_pairs.
 * meta[FRB_META_LIST_COUNT] receives the number of labels; ws: 256 bytes. */
int frb_direct_export_sorted(const void* slots, uint64_t cap, int dtype, uint64_t* meta, void* uniq_out, uint64_t* vals_out,
                             size_t capacity, void* ws, size_t ws_bytes, void* stream);
/* sort strategy, two steps: _count copies + radix sorts the voxels inside ws and leaves the number
 * of distinct labels in meta[FRB_META_LIST_COUNT]; _emit run-length encodes the sorted copy. */
size_t frb_unique_sort_ws_bytes(size_t n, int dtype);
int frb_unique_sort_count(const void* a, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, void* stream);
int frb_unique_sort_emit(size_t n, int dtype, void* uniq_out, uint64_t* counts_out, size_t capacity, const void* ws, void* stream);

/* ---- indices (SURVEY 8(f) "next" row) -------------------------------------------------
 * Replaces indices (fastremap.pyx:104-119): the ascending flat positions i with a[i] == value, in
 * two steps because the caller sizes the output: _count leaves the number of matches in
 * meta[FRB_META_LIST_COUNT] (and per-CTA offsets in ws), _emit writes the positions. */
size_t frb_indices_ws_bytes(size_t n);
int frb_indices_count(const void* a, size_t n, int dtype, uint64_t value_bits, uint64_t* meta, void* ws, size_t ws_bytes, void* stream);
int frb_indices_emit(const void* a, size_t n, int dtype, uint64_t value_bits, uint64_t* out, size_t capacity, const void* ws, void* stream);

/* ---- point_cloud (SURVEY 8(f) "next" row) ------------------------------------------------
 * Replaces _point_cloud_2d / _3d (fastremap.pyx:1444-1540): the non-zero voxels of a C-order image
 * as (label, flat index) uint64 pairs in scan order (labels with the sign bit flipped for signed
 * dtypes), a stable sort by label, and the indices unravelled to uint16 coordinate rows. */
size_t frb_nonzero_ws_bytes(size_t n);
int frb_nonzero_count(const void* a, size_t n, int dtype, uint64_t* meta, void* ws, size_t ws_bytes, void* stream);
int frb_nonzero_emit(const void* a, size_t n, int dtype, uint64_t* labels_out, uint64_t* index_out, size_t capacity,
                     const void* ws, void* stream);
size_t frb_sort_pairs_u64_ws_bytes(size_t n);
int frb_sort_pairs_u64(uint64_t* keys, uint64_t* vals, size_t n, int key_bits, void* ws, size_t ws_bytes, void* stream);
/* out[p] = unravel(index[p]) for a C-order shape (d0, d1) or (d0, d1, d2): rows of ndim uint16 */
int frb_unravel_u16(const uint64_t* index, size_t n, int ndim, uint64_t d1, uint64_t d2, uint16_t* out, void* stream);

/* ---- component_map -------------------------------------------------------------------
 * Replaces _component_map (fastremap.pyx:565-587): at every run start of cc the table
 * receives key cc[i] with value max(i+1); frb_component_gather then emits (cc label, parent[i])
 * for the winning run start of every label.  count in meta[FRB_META_LIST_COUNT].
 * row_bytes / plane_bytes: performance hints (0 = unknown), see frb_renumber_build / frb_hist_hash.
 * slot_list (optional, slot_list_capacity uint32 words): the build lists the index of every slot it fills -- per
 * consumer warp in a region of its own, no atomics (header, per-region counts, regions: csrc/frb_build.cuh) -- and adds
 * the number of listed slots to meta[FRB_META_AUX1]; a region that ran out of space adds 2^40.  When
 * AUX1 + SIDE_PRESENT == COUNT the list is complete and frb_component_gather can walk it (slot_list,
 * slot_list_count = AUX1) instead of scanning and compacting the whole slot array (no label may have been withdrawn
 * by frb_component_unhead).  Size it at 4 words per expected label or more.  NULL: no list, the gather scans. */
int frb_component_build(const void* cc, size_t n, int cc_dtype, uint64_t flat_offset, void* slots, uint64_t cap,
                        uint64_t* meta, uint64_t row_bytes, uint64_t plane_bytes, uint32_t* slot_list, uint64_t slot_list_capacity,
                        void* stream);
/* z-sharded volumes: withdraw the offer frb_component_build made for the slab's first voxel if that voxel carries
 * prev_bits, the label the previous slab ended with (it continues a run, :581) and no later run start of the label in
 * this slab has overtaken it; frb_component_gather skips withdrawn labels. */
int frb_component_unhead(const void* cc, int cc_dtype, uint64_t prev_bits, uint64_t flat_offset, void* slots, uint64_t cap, uint64_t* meta,
                         void* stream);
int frb_component_gather(const void* slots, uint64_t cap, uint64_t* meta, int cc_dtype, const void* parent, int p_dtype,
                         uint64_t flat_offset, void* keys_out, void* parents_out, size_t capacity, const uint32_t* slot_list,
                         uint64_t slot_list_count, void* stream);

/* ---- synthetic volumes (benchmark / test input) -------------------
 * Blocky "connectomics-like" labels: a gz*gy*gx grid of cells with splitmix64 ids, nearest
 * neighbour upsampled to (nz_total, ny, nx); this call fills planes [z0, z0+nz) in C order.
 * cell_div > 1 merges cell_div consecutive cells into one label (a "parent" volume that is a
 * function of the cell_div == 1 volume with the same grid). */
int frb_synth_volume(void* out, int dtype, uint64_t nz, uint64_t ny, uint64_t nx, uint64_t z0, uint64_t nz_total,
                     uint64_t gz, uint64_t gy, uint64_t gx, uint64_t seed, double zero_frac, double noise,
                     uint64_t cell_div, void* stream);

/* calibration: the library's TMA-staged streaming loop with no arithmetic.  mode 0 reads nbytes
 * (xor checksum into *out); mode 1 also writes them to dst_wide; mode 2 also writes half of them
 * to dst_narrow (the 8 B read / 8 B + 4 B write shape of the renumber relabel pass).  mode | 4: the copy
 * also carries the 16 bytes in front of every tile.  mode | 8 (with mode 2): the windowed variant -- ONE
 * persistent cooperative kernel, per window (FRB_PROBE_WINDOW_MIB, default 64) a read pass, a grid barrier,
 * the read + write pass, a grid barrier; out must then hold two words (out[1] is the barrier word). */
int frb_stream_probe(const void* src, size_t nbytes, void* dst_wide, void* dst_narrow, int mode, uint64_t* out, void* stream);
/* bind this library's (statically linked) CUDA runtime to `device`; call once per process/thread
 * before the first launch when the caller's framework selected a device other than 0 */
int frb_set_device(int device);
/* tuning knob: consecutive 16 KiB tiles a CTA processes before it jumps ahead by gridDim.x runs
 * (1 = grid-stride interleave, large = one contiguous chunk per CTA).  Default 64. */
void frb_set_tile_run(int run);
/* tuning knob: which lookup kernel frb_relabel uses for 64-bit labels when batch_rows > 1.  -1 (default): a sampler
 * counts run heads in 64 blocks of the volume and chooses per call; 0: k_relabel_batched (two rows of home-pair loads
 * in flight per warp); 1: k_relabel_runs (one lookup per run of equal labels, a warp-tile's runs in flight together). */
void frb_set_relabel_variant(int variant);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t frb_launch_count(void);
const char* frb_last_error(void);

/* ---- small arrays (up to ~2^21 voxels): one cooperative kernel per call ---------------------------
 * renumber (fastremap.pyx:196-257 + the refit astype :301) and unique(return_counts) (:855-965) with every phase --
 * table init, build, compaction, sort, rank / emit, relabel -- inside ONE cooperatively launched kernel, grid barriers
 * in between: a 32^3 chunk costs one launch instead of a dozen.  Same results as the volume-sized entry points.
 * meta[FRB_META_OVERFLOW] != 0 afterwards: more labels than the small table holds (frb_small_list_capacity) -- nothing
 * usable was written, take the volume-sized path.  meta[FRB_META_ASSIGNED] (renumber) / meta[FRB_META_LIST_COUNT]
 * (unique): number of labels.  narrow_out (n * width bytes) receives the ids at the refit width picked from
 * lim1 / lim2 / lim4 like frb_relabel_auto (nothing if no narrower dtype fits: use wide_out, which may alias src). */
uint64_t frb_small_table_cap(size_t n, int dtype);
size_t frb_small_list_capacity(size_t n, int dtype);
size_t frb_small_ws_bytes(size_t n, int dtype);
int frb_renumber_small(const void* src, size_t n, int dtype, int64_t start, int skip_zero, void* wide_out, void* narrow_out,
                       uint64_t lim1, uint64_t lim2, uint64_t lim4, void* keys_out, void* ids_out, uint64_t* meta, void* ws,
                       size_t ws_bytes, void* stream);
int frb_unique_small(const void* src, size_t n, int dtype, void* uniq_out, uint64_t* counts_out, uint64_t* meta, void* ws,
                     size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
