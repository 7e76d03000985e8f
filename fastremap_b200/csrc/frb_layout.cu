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
//                      (tobytes C->C / F->F): 1..16-byte granules, four in flight per thread, index
//                      decoding by multiply-high division with host-made reciprocals.
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
  // merged tile axes (k_transpose_tiles<.., true>): A = Ai * A1 where the A1 axis continues A contiguously
  // in the INPUT; its output stride is out_sa1.  Likewise B = Bi * B1, contiguous in the OUTPUT, input
  // stride in_sb1.  Row offsets are split with the reciprocals ai_magic / bi_magic.
  u64 in_sb1, out_sa1, ai_magic, bi_magic;
  unsigned int Ai, Bi;
  u64 tiles_a, tiles_b, total;
  int nb;
  int fast;                           // every index that is divided fits 32 bits: multiply-high division
  u64 amagic;                         // ceil(2^64 / A)
  u64 bmagic[LAYOUT_MAX_BATCH];       // ceil(2^64 / bd[j])
  unsigned int bd[LAYOUT_MAX_BATCH];  // batch axis lengths (fastest-decoded first), none of them 1
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

// x / d for x, d < 2^32 with magic = ceil(2^64 / d): exact, because x * (magic - 2^64 / d) < 2^32 <= 2^64 / d
__device__ __forceinline__ unsigned int fast_div(unsigned int x, u64 magic) { return (unsigned int)__umul64hi((u64)x, magic); }

__device__ __forceinline__ void batch_offsets_fast(const LayoutArgs& P, unsigned int b, u64& in_off, u64& out_off) {
  in_off = 0;
  out_off = 0;
#pragma unroll
  for (int j = 0; j < LAYOUT_MAX_BATCH; j++) {
    if (j < P.nb) {
      const unsigned int q = fast_div(b, P.bmagic[j]);
      const unsigned int c = b - q * P.bd[j];
      b = q;
      in_off += (u64)c * P.bis[j];
      out_off += (u64)c * P.bos[j];
    }
  }
}

template <bool M>
__device__ __forceinline__ u64 in_row_offset(const LayoutArgs& P, u64 b) {
  if (!M) return b * P.in_sb;
  const unsigned int q = fast_div((unsigned int)b, P.bi_magic);
  return (u64)((unsigned int)b - q * P.Bi) * P.in_sb + (u64)q * P.in_sb1;
}
template <bool M>
__device__ __forceinline__ u64 out_row_offset(const LayoutArgs& P, u64 a) {
  if (!M) return a * P.out_sa;
  const unsigned int q = fast_div((unsigned int)a, P.ai_magic);
  return (u64)((unsigned int)a - q * P.Ai) * P.out_sa + (u64)q * P.out_sa1;
}

template <int W, int T, bool M>
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
        const E* row = in + in_off + in_row_offset<M>(P, b0 + warp + 8 * i) + a0;
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
        E* row = out + out_off + out_row_offset<M>(P, a0 + warp + 8 * i) + b0;
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
          if (a < P.A && b < P.B) tile[warp + 8 * i][lane + 32 * c] = in[in_off + in_row_offset<M>(P, b) + a];
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < RPW; i++) {
        const u64 a = a0 + warp + 8 * i;
#pragma unroll
        for (int c = 0; c < CPL; c++) {
          const u64 b = b0 + lane + 32 * c;
          if (a < P.A && b < P.B) out[out_off + out_row_offset<M>(P, a) + b] = tile[lane + 32 * c][warp + 8 * i];
        }
      }
    }
    __syncthreads();
  }
}

// 1- and 2-byte elements, R = 4 / W of them per 32-bit word: every global access is a full word, the
// R x R element blocks are transposed in registers with byte permutes, and shared memory holds words
// [a][b / R] with the column rotated by a / R so that both the scattered writes (32 different a, one
// column) and the row reads are bank-conflict free.  The tile is 32HR x 32HR elements: with H = 2 every
// row segment read or written is 256 bytes, which is what DRAM wants (128-byte segments reach ~0.6 of
// the copy bandwidth, 256-byte ones ~0.9).  Needs A, B and every stride divisible by R and 4-byte
// aligned bases.  Dynamic shared memory: (32HR) * (32H) words.
// M: merged tile axes, as in k_transpose_tiles -- what makes the chunk grids of tobytes (64-byte rows inside a chunk)
// eligible: A = (z in chunk) x (chunk index along z), B = (x in chunk) x (y in chunk) are both thousands long.
template <int W, int H, bool M>
__global__ void __launch_bounds__(256) k_transpose_packed(const __grid_constant__ LayoutArgs P) {
  constexpr int R = 4 / W;
  constexpr int TW = 32 * H;  // tile edge in words
  constexpr int T = TW * R;   // tile edge in elements
  extern __shared__ __align__(16) uint32_t S[];  // [T][TW]
  const uint32_t* __restrict__ in = (const uint32_t*)P.in;
  uint32_t* __restrict__ out = (uint32_t*)P.out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (u64 t = blockIdx.x; t < P.total; t += gridDim.x) {
    u64 r = t;
    const u64 ta = r % P.tiles_a;
    r /= P.tiles_a;
    const u64 tb = r % P.tiles_b;
    r /= P.tiles_b;
    u64 in_off, out_off;
    batch_offsets(P, r, in_off, out_off);
    const u64 a0 = ta * T, b0 = tb * T;
    const uint32_t* src = in + (in_off + a0) / R + lane;
    bool a_ok[H];
#pragma unroll
    for (int h = 0; h < H; h++) a_ok[h] = a0 + (u64)R * (lane + 32 * h) < P.A;
    // row groups rb = warp + 8 * i, i < 4H; NB of them (NB * H * R loads per thread) in flight at a time
    constexpr int NB = (4 * H) < (32 / (H * R)) ? (4 * H) : (32 / (H * R));  // 32 words per thread
#pragma unroll
    for (int i0 = 0; i0 < 4 * H; i0 += NB) {
      uint32_t w[NB][H][R];
#pragma unroll
      for (int ii = 0; ii < NB; ii++) {
        const int rb = warp + 8 * (i0 + ii);
        // the R rows of a group lie in one segment of a merged B axis (Bi, b0 and R * rb are multiples of R): one
        // reciprocal multiply per group, not per row
        const u64 bg = b0 + (u64)R * rb;
        const u64 group_w = bg < P.B ? in_row_offset<M>(P, bg) / R : 0;
        const u64 step_w = P.in_sb / R;
#pragma unroll
        for (int j = 0; j < R; j++) {
          const u64 b = bg + j;
          const u64 row_w = group_w + j * step_w;
#pragma unroll
          for (int h = 0; h < H; h++) w[ii][h][j] = (a_ok[h] && b < P.B) ? __ldcs(src + row_w + 32 * h) : 0u;
        }
      }
#pragma unroll
      for (int ii = 0; ii < NB; ii++) {
        const int rb = warp + 8 * (i0 + ii);
#pragma unroll
        for (int h = 0; h < H; h++) {
          uint32_t tr[R];
          if constexpr (R == 4) {
            const uint32_t x0 = __byte_perm(w[ii][h][0], w[ii][h][1], 0x5140), x1 = __byte_perm(w[ii][h][0], w[ii][h][1], 0x7362);
            const uint32_t y0 = __byte_perm(w[ii][h][2], w[ii][h][3], 0x5140), y1 = __byte_perm(w[ii][h][2], w[ii][h][3], 0x7362);
            tr[0] = __byte_perm(x0, y0, 0x5410);
            tr[1] = __byte_perm(x0, y0, 0x7632);
            tr[2] = __byte_perm(x1, y1, 0x5410);
            tr[3] = __byte_perm(x1, y1, 0x7632);
          } else {
            tr[0] = __byte_perm(w[ii][h][0], w[ii][h][1], 0x5410);
            tr[1] = __byte_perm(w[ii][h][0], w[ii][h][1], 0x7632);
          }
          const int ac = lane + 32 * h;
#pragma unroll
          for (int k = 0; k < R; k++) S[(R * ac + k) * TW + ((rb + ac) & (TW - 1))] = tr[k];
        }
      }
    }
    __syncthreads();
    uint32_t* dst = out + (out_off + b0) / R + lane;
    bool b_ok[H];
#pragma unroll
    for (int h = 0; h < H; h++) b_ok[h] = b0 + (u64)R * (lane + 32 * h) < P.B;
#pragma unroll 4
    for (int i = 0; i < 4 * H * R; i++) {
      const int a = warp + 8 * i;
      if (a0 + a < P.A) {
        const u64 row_w = out_row_offset<M>(P, a0 + a) / R;
#pragma unroll
        for (int h = 0; h < H; h++)
          if (b_ok[h]) __stcs(dst + row_w + 32 * h, S[a * TW + ((lane + 32 * h + a / R) & (TW - 1))]);
      }
    }
    __syncthreads();
  }
}

// U granules of G bytes per thread and iteration; A = granules per row, rows decoded as batch indices
template <int G, int U>
__global__ void __launch_bounds__(256) k_copy_rows(const __grid_constant__ LayoutArgs P) {
  using E = typename ElemOf<G>::type;
  const E* __restrict__ in = (const E*)P.in;
  E* __restrict__ out = (E*)P.out;
  const u64 step = (u64)gridDim.x * blockDim.x;
  for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < P.total; i0 += step * U) {
    E v[U];
    u64 dst[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const u64 i = i0 + (u64)u * step;
      dst[u] = ~0ull;
      if (i < P.total) {
        u64 in_off, out_off, col;
        if (P.fast) {
          const unsigned int row = fast_div((unsigned int)i, P.amagic);
          col = (unsigned int)i - row * (unsigned int)P.A;
          batch_offsets_fast(P, row, in_off, out_off);
        } else {
          const u64 row = i / P.A;
          col = i % P.A;
          batch_offsets(P, row, in_off, out_off);
        }
        v[u] = __ldcs(in + in_off + col);
        dst[u] = out_off + col;
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (dst[u] != ~0ull) __stcs(out + dst[u], v[u]);
  }
}

static bool fill_batch(LayoutArgs& P, int nbatch, const uint64_t* dims, const uint64_t* in_strides, const uint64_t* out_strides,
                       u64& count) {
  if (nbatch < 0 || nbatch > LAYOUT_MAX_BATCH || (nbatch > 0 && (!dims || !in_strides || !out_strides))) return false;
  P.nb = 0;
  P.fast = 0;
  P.amagic = 0;
  count = 1;
  for (int j = 0; j < LAYOUT_MAX_BATCH; j++) {
    P.bd[j] = 1;
    P.bis[j] = P.bos[j] = P.bmagic[j] = 0;
  }
  for (int j = 0; j < nbatch; j++) {
    if (dims[j] < 1 || dims[j] > 0xFFFFFFFFull) return false;
    count *= dims[j];
    if (dims[j] == 1) continue;  // a unit axis contributes nothing
    P.bd[P.nb] = (unsigned int)dims[j];
    P.bmagic[P.nb] = ~0ull / dims[j] + 1;
    P.bis[P.nb] = in_strides[j];
    P.bos[P.nb] = out_strides[j];
    P.nb++;
  }
  return true;
}

template <int W>
static int launch_transpose(LayoutArgs& P, u64 batches, cudaStream_t st) {
  const u64 max_grid = (u64)sm_count() * 64;
  // A short tile axis makes short row segments (DRAM wants >= 256 bytes).  If a batch axis continues it
  // contiguously (in the input for A, in the output for B) the two are walked as ONE tile axis, e.g. the
  // channel axis of an (x, y, z, c) array together with z, or a chunk's rows across the chunk grid.
  bool merged = false;
  P.Ai = (unsigned int)P.A;
  P.Bi = (unsigned int)P.B;
  P.in_sb1 = P.out_sa1 = 0;
  auto drop_axis = [&](int j) {
    for (int k = j; k + 1 < P.nb; k++) {
      P.bd[k] = P.bd[k + 1];
      P.bmagic[k] = P.bmagic[k + 1];
      P.bis[k] = P.bis[k + 1];
      P.bos[k] = P.bos[k + 1];
    }
    P.nb--;
  };
  constexpr u64 SHORT = W >= 4 ? 64 : 512 / W;  // elements: rows under 256 bytes (4-/8-byte: 64 elements, as tuned), 512 bytes for 1-/2-byte (one packed tile edge is 256)
  if (P.A < SHORT && P.A > 1 && P.A <= 0xFFFFFFFFull)
    for (int j = 0; j < P.nb; j++)
      if (P.bis[j] == P.A && P.A * P.bd[j] <= 0xFFFFFFFFull) {
        P.out_sa1 = P.bos[j];
        P.A *= P.bd[j];
        batches /= P.bd[j];
        drop_axis(j);
        merged = true;
        break;
      }
  if (P.B < SHORT && P.B > 1 && P.B <= 0xFFFFFFFFull)
    for (int j = 0; j < P.nb; j++)
      if (P.bos[j] == P.B && P.B * P.bd[j] <= 0xFFFFFFFFull) {
        P.in_sb1 = P.bis[j];
        P.B *= P.bd[j];
        batches /= P.bd[j];
        drop_axis(j);
        merged = true;
        break;
      }
  P.ai_magic = P.Ai > 1 ? ~0ull / P.Ai + 1 : 0;
  P.bi_magic = P.Bi > 1 ? ~0ull / P.Bi + 1 : 0;
  if constexpr (W <= 2) {
    constexpr u64 R = 4 / W;
    u64 bits = P.A | P.B | P.in_sb | P.out_sa | P.Ai | P.Bi | P.in_sb1 | P.out_sa1;
    for (int j = 0; j < P.nb; j++) bits |= P.bis[j] | P.bos[j];
    const bool aligned = bits % R == 0 && (uintptr_t)P.in % 4 == 0 && (uintptr_t)P.out % 4 == 0;
    // 256-byte segments (H = 2) when both axes fill most of such a tile; 128-byte ones only pay for bytes
    const int H = (P.A >= 48 * R && P.B >= 48 * R) ? 2 : ((W == 1 && P.A >= 32 * R && P.B >= 32 * R) ? 1 : 0);
    if (aligned && H) {
      const u64 T = 32 * H * R;
      P.tiles_a = (P.A + T - 1) / T;
      P.tiles_b = (P.B + T - 1) / T;
      P.total = P.tiles_a * P.tiles_b * batches;
      const unsigned int grid = (unsigned int)(P.total < max_grid ? P.total : max_grid);
      const size_t smem = (size_t)T * 32 * H * sizeof(uint32_t);
#define FRB_PACKED(HH, MM)                                                                                              \
  {                                                                                                                     \
    cudaFuncSetAttribute(k_transpose_packed<W, HH, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); /* per device: always */ \
    k_transpose_packed<W, HH, MM><<<grid, 256, smem, st>>>(P);                                                          \
  }
      if (H == 2) { if (merged) FRB_PACKED(2, true) else FRB_PACKED(2, false) }
      else { if (merged) FRB_PACKED(1, true) else FRB_PACKED(1, false) }
#undef FRB_PACKED
      return check_launch("k_transpose_packed");
    }
  }
  const bool small = P.A <= 32 || P.B <= 32;  // short axes: 32 x 32 tiles waste less
  const int T = small ? 32 : 64;
  P.tiles_a = (P.A + T - 1) / T;
  P.tiles_b = (P.B + T - 1) / T;
  P.total = P.tiles_a * P.tiles_b * batches;
  const unsigned int grid = (unsigned int)(P.total < max_grid ? P.total : max_grid);
  if (merged) {
    if (small) k_transpose_tiles<W, 32, true><<<grid, 256, 0, st>>>(P);
    else k_transpose_tiles<W, 64, true><<<grid, 256, 0, st>>>(P);
  } else {
    if (small) k_transpose_tiles<W, 32, false><<<grid, 256, 0, st>>>(P);
    else k_transpose_tiles<W, 64, false><<<grid, 256, 0, st>>>(P);
  }
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
  for (int j = 0; j < P.nb; j++) bits |= P.bis[j] | P.bos[j];
  const int g = (int)(bits & (~bits + 1));
  for (int j = 0; j < P.nb; j++) {
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
  P.fast = P.total <= 0xFFFFFFFFull && P.A > 1;
  P.amagic = P.A > 1 ? ~0ull / P.A + 1 : 0;
  constexpr int U = 4;
  const u64 want = (P.total + 256 * U - 1) / (256 * U);
  const u64 max_grid = (u64)sm_count() * 8;
  const unsigned int grid = (unsigned int)(want < max_grid ? want : max_grid);
  cudaStream_t st = (cudaStream_t)stream;
  switch (g) {
    case 1: k_copy_rows<1, U><<<grid, 256, 0, st>>>(P); break;
    case 2: k_copy_rows<2, U><<<grid, 256, 0, st>>>(P); break;
    case 4: k_copy_rows<4, U><<<grid, 256, 0, st>>>(P); break;
    case 8: k_copy_rows<8, U><<<grid, 256, 0, st>>>(P); break;
    default: k_copy_rows<16, U><<<grid, 256, 0, st>>>(P); break;
  }
  return check_launch("k_copy_rows");
}

}  // extern "C"
