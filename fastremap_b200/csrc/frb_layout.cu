// frb_layout.cu -- memory-layout changes either side of the relabel path:
// physical transposition (C order <-> Fortran order) and the chunk-grid serialisation of `tobytes`.
//
// Replaces pyipt::_ipt2d/_ipt3d/_ipt4d (fastremap.pyx:61-68, ipt.hpp:330-352: in-place square swaps and
// follow-the-cycles rectangular permutations) and the strip loops of tobytes (fastremap.pyx:1628-1654).
// On the GPU both are ONE pass through HBM, out of place (the host side hands the result back in the
// caller's buffer, which is what "in place" means to a caller):
//
//   k_transpose_tiles  batched 2-D transposition through shared memory.  Axis A is contiguous in the
//                      input, axis B is contiguous in the output; up to 6 further axes are batch axes
//                      with independent input / output strides.  "Reverse all axes" (C <-> F) of a
//                      2-D .. 8-D array and "cut into chunks AND switch the order inside each chunk"
//                      are both instances.  Tiles of T x T elements, rows padded so that the column
//                      reads are bank-conflict free for every element width.
//   k_copy_rows        batched row copy for layouts whose contiguous axis is the same on both sides
//                      (tobytes C->C / F->F): one 1..16-byte granule per thread.
//
// Roofline: HBM, 2 * W bytes per element (one read, one write).
#include "frb_internal.cuh"

namespace frb {

static constexpr int LAYOUT_MAX_BATCH = 6;

struct LayoutArgs {
  const void* in;
  void* out;
  u64 A, B;            // transposition: tile axes; row copy: A = granules per row, B unused
  u64 in_sb, out_sa;   // input stride of axis B, output stride of axis A (elements)
  u64 tiles_a, tiles_b, total;
  int nb;
  unsigned int bd[LAYOUT_MAX_BATCH];  // batch axis lengths (fastest-decoded first)
  u64 bis[LAYOUT_MAX_BATCH];          // input strides (elements / granules)
  u64 bos[LAYOUT_MAX_BATCH];          // output strides
};

template <int W> struct ElemOf;
template <> struct ElemOf<1> { using type = uint8_t; };
template <> struct ElemOf<2> { using type = uint16_t; };
template <> struct ElemOf<4> { using type = uint32_t; };
template <> struct ElemOf<8> { using type = u64; };
template <> struct ElemOf<16> { using type = uint4; };

// decode a batch index into the two base offsets
__device__ __forceinline__ void batch_offsets(const LayoutArgs& P, u64 b, u64& in_off, u64& out_off) {
  in_off = 0;
  out_off = 0;
#pragma unroll
  for (int j = 0; j < LAYOUT_MAX_BATCH; j++) {
    if (j < P.nb) {
      u64 c;
      if (b <= 0xFFFFFFFFull) {  // 32-bit division whenever it is enough
        const unsigned int b32 = (unsigned int)b;
        c = b32 % P.bd[j];
        b = b32 / P.bd[j];
      } else {
        c = b % P.bd[j];
        b = b / P.bd[j];
      }
      in_off += c * P.bis[j];
      out_off += c * P.bos[j];
    }
  }
}

template <int W, int T>
__global__ void __launch_bounds__(256) k_transpose_tiles(const __grid_constant__ LayoutArgs P) {
  using E = typename ElemOf<W>::type;
  constexpr int PAD = W >= 4 ? 1 : 4 / W;  // row stride = 4 * odd bytes: a column read touches 32 banks
  __shared__ E tile[T][T + PAD];
  const E* __restrict__ in = (const E*)P.in;
  E* __restrict__ out = (E*)P.out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int RPW = T / 8;   // tile rows per warp
  constexpr int CPL = T / 32;  // columns per lane
  for (u64 t = blockIdx.x; t < P.total; t += gridDim.x) {
    u64 r = t;
    const u64 ta = r % P.tiles_a;
    r /= P.tiles_a;
    const u64 tb = r % P.tiles_b;
    r /= P.tiles_b;
    u64 in_off, out_off;
    batch_offsets(P, r, in_off, out_off);
    const u64 a0 = ta * T, b0 = tb * T;
    const bool full = a0 + T <= P.A && b0 + T <= P.B;
    if (full) {
      E v[RPW][CPL];
#pragma unroll
      for (int i = 0; i < RPW; i++) {
        const E* row = in + in_off + (b0 + warp + 8 * i) * P.in_sb + a0;
#pragma unroll
        for (int c = 0; c < CPL; c++) v[i][c] = __ldcs(row + lane + 32 * c);
      }
#pragma unroll
      for (int i = 0; i < RPW; i++)
#pragma unroll
        for (int c = 0; c < CPL; c++) tile[warp + 8 * i][lane + 32 * c] = v[i][c];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < RPW; i++) {
        E* row = out + out_off + (a0 + warp + 8 * i) * P.out_sa + b0;
#pragma unroll
        for (int c = 0; c < CPL; c++) __stcs(row + lane + 32 * c, tile[lane + 32 * c][warp + 8 * i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < RPW; i++) {
        const u64 b = b0 + warp + 8 * i;
#pragma unroll
        for (int c = 0; c < CPL; c++) {
          const u64 a = a0 + lane + 32 * c;
          if (a < P.A && b < P.B) tile[warp + 8 * i][lane + 32 * c] = in[in_off + b * P.in_sb + a];
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < RPW; i++) {
        const u64 a = a0 + warp + 8 * i;
#pragma unroll
        for (int c = 0; c < CPL; c++) {
          const u64 b = b0 + lane + 32 * c;
          if (a < P.A && b < P.B) out[out_off + a * P.out_sa + b] = tile[lane + 32 * c][warp + 8 * i];
        }
      }
    }
    __syncthreads();
  }
}

// one granule of G bytes per thread; A = granules per row, rows decoded as batch indices
template <int G>
__global__ void __launch_bounds__(256) k_copy_rows(const __grid_constant__ LayoutArgs P) {
  using E = typename ElemOf<G>::type;
  const E* __restrict__ in = (const E*)P.in;
  E* __restrict__ out = (E*)P.out;
  const u64 step = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < P.total; i += step) {
    u64 row, col;
    if (i <= 0xFFFFFFFFull && P.A <= 0xFFFFFFFFull) {
      row = (unsigned int)i / (unsigned int)P.A;
      col = (unsigned int)i % (unsigned int)P.A;
    } else {
      row = i / P.A;
      col = i % P.A;
    }
    u64 in_off, out_off;
    batch_offsets(P, row, in_off, out_off);
    out[out_off + col] = in[in_off + col];
  }
}

static bool fill_batch(LayoutArgs& P, int nbatch, const uint64_t* dims, const uint64_t* in_strides, const uint64_t* out_strides,
                       u64& count) {
  if (nbatch < 0 || nbatch > LAYOUT_MAX_BATCH || (nbatch > 0 && (!dims || !in_strides || !out_strides))) return false;
  P.nb = nbatch;
  count = 1;
  for (int j = 0; j < LAYOUT_MAX_BATCH; j++) {
    P.bd[j] = 1;
    P.bis[j] = P.bos[j] = 0;
  }
  for (int j = 0; j < nbatch; j++) {
    if (dims[j] < 1 || dims[j] > 0xFFFFFFFFull) return false;
    P.bd[j] = (unsigned int)dims[j];
    P.bis[j] = in_strides[j];
    P.bos[j] = out_strides[j];
    count *= dims[j];
  }
  return true;
}

template <int W>
static int launch_transpose(LayoutArgs& P, u64 batches, cudaStream_t st) {
  const bool small = P.A <= 32 || P.B <= 32;  // short axes: 32 x 32 tiles waste less
  const int T = small ? 32 : 64;
  P.tiles_a = (P.A + T - 1) / T;
  P.tiles_b = (P.B + T - 1) / T;
  P.total = P.tiles_a * P.tiles_b * batches;
  const u64 max_grid = (u64)sm_count() * 64;
  const unsigned int grid = (unsigned int)(P.total < max_grid ? P.total : max_grid);
  if (small) k_transpose_tiles<W, 32><<<grid, 256, 0, st>>>(P);
  else k_transpose_tiles<W, 64><<<grid, 256, 0, st>>>(P);
  return check_launch("k_transpose_tiles");
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_transpose_tiles(const void* in, void* out, int width, uint64_t dim_a, uint64_t dim_b, uint64_t in_stride_b,
                        uint64_t out_stride_a, int nbatch, const uint64_t* batch_dims, const uint64_t* batch_in_strides,
                        const uint64_t* batch_out_strides, void* stream) {
  LayoutArgs P;
  u64 batches;
  if (!fill_batch(P, nbatch, batch_dims, batch_in_strides, batch_out_strides, batches) ||
      (width != 1 && width != 2 && width != 4 && width != 8) || in == out) {
    set_error("frb_transpose_tiles: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (dim_a == 0 || dim_b == 0) return FRB_OK;
  if (!in || !out) {
    set_error("frb_transpose_tiles: null buffer");
    return FRB_ERR_BAD_ARG;
  }
  P.in = in;
  P.out = out;
  P.A = dim_a;
  P.B = dim_b;
  P.in_sb = in_stride_b;
  P.out_sa = out_stride_a;
  cudaStream_t st = (cudaStream_t)stream;
  switch (width) {
    case 1: return launch_transpose<1>(P, batches, st);
    case 2: return launch_transpose<2>(P, batches, st);
    case 4: return launch_transpose<4>(P, batches, st);
    default: return launch_transpose<8>(P, batches, st);
  }
}

int frb_copy_rows(const void* in, void* out, uint64_t row_bytes, int nbatch, const uint64_t* batch_dims,
                  const uint64_t* batch_in_strides_bytes, const uint64_t* batch_out_strides_bytes, void* stream) {
  LayoutArgs P;
  u64 rows;
  if (!fill_batch(P, nbatch, batch_dims, batch_in_strides_bytes, batch_out_strides_bytes, rows) || in == out) {
    set_error("frb_copy_rows: bad arguments");
    return FRB_ERR_BAD_ARG;
  }
  if (row_bytes == 0) return FRB_OK;
  if (!in || !out) {
    set_error("frb_copy_rows: null buffer");
    return FRB_ERR_BAD_ARG;
  }
  // granule: the largest power of two (<= 16) that divides the row length, every stride and both base addresses
  u64 bits = row_bytes | (u64)(uintptr_t)in | (u64)(uintptr_t)out | 16;
  for (int j = 0; j < nbatch; j++) bits |= P.bis[j] | P.bos[j];
  const int g = (int)(bits & (~bits + 1));
  for (int j = 0; j < nbatch; j++) {
    P.bis[j] /= g;
    P.bos[j] /= g;
  }
  P.in = in;
  P.out = out;
  P.A = row_bytes / g;
  P.B = 1;
  P.in_sb = P.out_sa = 0;
  P.tiles_a = P.tiles_b = 0;
  P.total = P.A * rows;
  const u64 want = (P.total + 255) / 256;
  const u64 max_grid = (u64)sm_count() * 32;
  const unsigned int grid = (unsigned int)(want < max_grid ? want : max_grid);
  cudaStream_t st = (cudaStream_t)stream;
  switch (g) {
    case 1: k_copy_rows<1><<<grid, 256, 0, st>>>(P); break;
    case 2: k_copy_rows<2><<<grid, 256, 0, st>>>(P); break;
    case 4: k_copy_rows<4><<<grid, 256, 0, st>>>(P); break;
    case 8: k_copy_rows<8><<<grid, 256, 0, st>>>(P); break;
    default: k_copy_rows<16><<<grid, 256, 0, st>>>(P); break;
  }
  return check_launch("k_copy_rows");
}

}  // extern "C"
