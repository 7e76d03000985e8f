// frb_internal.cuh -- host-side helpers shared between translation units.
#pragma once
#include "frb_common.cuh"

namespace frb {

// frb_runtime.cu
int launch_table_export(const Slot* slots, u64 cap, u64* meta, u64* a_out, u64* b_out, size_t capacity, int mode,
                        cudaStream_t st);

int plane_band();       // tiles per band of the balanced plane walk (K7, k_hist): 16; FRB_PLANE_BAND overrides it for A/B runs (0: the first session's one-band-per-CTA walk)
int relabel_variant();  // tuning knob (frb_set_relabel_variant): -1 sampler decides, 0 k_relabel_batched, 1 k_relabel_runs

// frb_sort.cu -- LSD radix sort, 8 bits per pass, stable.
// Sorts by bits [begin_bit, end_bit) of the key.  keys_alt / vals_alt are ping-pong buffers of
// n elements; the sorted result always ends in keys / vals.  vals may be NULL (keys only).
// hist_ws must hold radix_hist_bytes(n) bytes.
size_t radix_hist_bytes(size_t n);
int radix_sort_u64(u64* keys, u64* vals, u64* keys_alt, u64* vals_alt, size_t n, int begin_bit, int end_bit,
                   void* hist_ws, cudaStream_t st);
int radix_sort_u32(uint32_t* keys, uint32_t* keys_alt, size_t n, int begin_bit, int end_bit, void* hist_ws,
                   cudaStream_t st);
// single-CTA exclusive scan of M uint64 values in place; total written to *total_out if not NULL
int scan_exclusive_u64(u64* data, size_t M, u64* total_out, cudaStream_t st);

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace frb
