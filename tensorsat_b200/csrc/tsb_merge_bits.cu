// tsb_merge_bits.cu — launchers of the bit-packed boolean merge and of the byte <-> bit conversions.
#include <type_traits>

#include "lower.h"
#include "k_merge_bits.cuh"

namespace tsb {

template <int KB, int HL, int HR>
static void bits_w_one(tsb_ctx* ctx, const MergeDesc& d, const void* l, uint64_t bl, const void* r, uint64_t br, void* o,
                       int grid, cudaStream_t stream) {
  auto k = k_merge_bits_w<KB, HL, HR>;
  TSB_LAUNCH_PDL(k, dim3(grid), dim3(kBitsThreads), 0, stream, d, (const uint32_t*)l, bl, (const uint32_t*)r, br, (uint4*)o);
}

template <int KB, int HL>
static void bits_w_r(tsb_ctx* ctx, const MergeDesc& d, const void* l, uint64_t bl, const void* r, uint64_t br, void* o,
                     int grid, cudaStream_t s) {
  switch (d.nibR >> 5) {
    case 0: bits_w_one<KB, HL, 0>(ctx, d, l, bl, r, br, o, grid, s); break;
    case 1: bits_w_one<KB, HL, 1>(ctx, d, l, bl, r, br, o, grid, s); break;
    case 2: bits_w_one<KB, HL, 2>(ctx, d, l, bl, r, br, o, grid, s); break;
    default: bits_w_one<KB, HL, 3>(ctx, d, l, bl, r, br, o, grid, s); break;
  }
}

template <int KB>
static void bits_w_l(tsb_ctx* ctx, const MergeDesc& d, const void* l, uint64_t bl, const void* r, uint64_t br, void* o,
                     int grid, cudaStream_t s) {
  switch (d.nibL >> 5) {
    case 0: bits_w_r<KB, 0>(ctx, d, l, bl, r, br, o, grid, s); break;
    case 1: bits_w_r<KB, 1>(ctx, d, l, bl, r, br, o, grid, s); break;
    case 2: bits_w_r<KB, 2>(ctx, d, l, bl, r, br, o, grid, s); break;
    default: bits_w_r<KB, 3>(ctx, d, l, bl, r, br, o, grid, s); break;
  }
}

void launch_merge_bits(tsb_ctx* ctx, const MergeDesc& d, const void* l, uint64_t bitoff_l, const void* r,
                       uint64_t bitoff_r, void* o, cudaStream_t stream) {
  TSB_CHECK(d.kbits <= 8, TSB_ERR_UNSUPPORTED, "internal: bit-packed merge with more than 8 summed wires");
  TSB_CHECK(((uintptr_t)l & 15) == 0 && (!d.has_rhs || ((uintptr_t)r & 15) == 0) && ((uintptr_t)o & 15) == 0, TSB_ERR_INVALID,
            "internal: bit-packed operands must be 16-byte aligned");
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  // the general kernel builds its lane tables per CTA: give every CTA a few tiles to amortise that — unless there are
  // too few tiles to fill the SMs anyway (the broadcast twins of the once-only schedule are on its critical path)
  const int64_t fill = (int64_t)ctx->sm_count * 2;
  const int grid = (int)std::max<int64_t>(1, n_tiles <= fill ? n_tiles : std::min<int64_t>(std::max<int64_t>((n_tiles + 3) / 4, fill), (int64_t)ctx->sm_count * 8));
  if (d.has_rhs && d.kbits <= 1 && ((d.nibL == 127 && bits_lane_class(d.nibR) == 2) || (d.nibR == 127 && bits_lane_class(d.nibL) == 2))) {
    // word-aligned fast path, one operand lacking a single in-word lane (too large for a broadcast twin): its fetch is
    // an 8-byte load whose 16-bit quarters are doubled block-wise in registers
    const int gw = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 8);
    const bool lhs_half = d.nibL != 127;
    if (d.kbits == 0) { if (lhs_half) bits_w_one<0, 4, 3>(ctx, d, l, bitoff_l, r, bitoff_r, o, gw, stream); else bits_w_one<0, 3, 4>(ctx, d, l, bitoff_l, r, bitoff_r, o, gw, stream); }
    else { if (lhs_half) bits_w_one<1, 4, 3>(ctx, d, l, bitoff_l, r, bitoff_r, o, gw, stream); else bits_w_one<1, 3, 4>(ctx, d, l, bitoff_l, r, bitoff_r, o, gw, stream); }
    ctx->note("k_merge_bits_w");
    TSB_CUDA(cudaGetLastError());
    return;
  }
  if (d.has_rhs && d.kbits <= 1 && (d.nibL & 31) == 31 && (d.nibR & 31) == 31) {
    // word-aligned fast path: both operands hold the five in-word lanes
    const int gw = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 8);
    if (d.kbits == 0) bits_w_l<0>(ctx, d, l, bitoff_l, r, bitoff_r, o, gw, stream);
    else bits_w_l<1>(ctx, d, l, bitoff_l, r, bitoff_r, o, gw, stream);
    ctx->note("k_merge_bits_w");
    TSB_CUDA(cudaGetLastError());
    return;
  }
  // lane classes known at compile time where it pays (k_merge_bits.cuh): a full-lane operand against anything, and
  // against one that lacks a single in-word lane
  int cl = bits_lane_class(d.nibL), cr = d.has_rhs ? bits_lane_class(d.nibR) : 0;
  if (cl == 2 && cr != 1) cl = 0;
  if (cr == 2 && cl != 1) cr = 0;
  const int kb = d.kbits <= 1 ? d.kbits : -1;
  auto go = [&](auto kern) {
    TSB_LAUNCH_PDL(kern, dim3(grid), dim3(kBitsThreads), 0, stream, d, (const uint32_t*)l, bitoff_l, (const uint32_t*)r, bitoff_r, (uint4*)o);
  };
  auto by_class = [&](auto kbc) {
    constexpr int KBC = decltype(kbc)::value;
    switch (cl * 3 + cr) {
      case 1 * 3 + 0: go(k_merge_bits<KBC, 1, 0>); break;
      case 0 * 3 + 1: go(k_merge_bits<KBC, 0, 1>); break;
      case 1 * 3 + 1: go(k_merge_bits<KBC, 1, 1>); break;
      case 1 * 3 + 2: go(k_merge_bits<KBC, 1, 2>); break;
      case 2 * 3 + 1: go(k_merge_bits<KBC, 2, 1>); break;
      default: go(k_merge_bits<KBC, 0, 0>); break;
    }
  };
  if (kb == 0) by_class(std::integral_constant<int, 0>{});
  else if (kb == 1) by_class(std::integral_constant<int, 1>{});
  else by_class(std::integral_constant<int, -1>{});
  ctx->note("k_merge_bits");
  TSB_CUDA(cudaGetLastError());
}

template <int NOPS, int KB>
static void bits_n_one(const BitsNDesc& d, const BitsNPtrs& pp, void* o, int grid, cudaStream_t stream) {
  auto k = k_merge_bits_n<NOPS, KB>;
  TSB_LAUNCH_PDL(k, dim3(grid), dim3(kBitsThreads), 0, stream, d, pp, (uint4*)o);
}

template <int NOPS>
static void bits_n_k(tsb_ctx* ctx, const BitsNDesc& d, const BitsNPtrs& pp, void* o, int grid, cudaStream_t stream) {
  (void)ctx;
  switch (d.kbits) {
    case 0: bits_n_one<NOPS, 0>(d, pp, o, grid, stream); break;
    case 1: bits_n_one<NOPS, 1>(d, pp, o, grid, stream); break;
    case 2: bits_n_one<NOPS, 2>(d, pp, o, grid, stream); break;
    default: bits_n_one<NOPS, 3>(d, pp, o, grid, stream); break;
  }
}

void launch_merge_bits_n(tsb_ctx* ctx, const BitsNDesc& d, const void* const* ptrs, const uint64_t* bitoffs, void* o,
                         cudaStream_t stream) {
  TSB_CHECK(d.n_ops >= 2 && d.n_ops <= kBitsMaxOps && d.kbits >= 0 && d.kbits <= kBitsMaxKB, TSB_ERR_INVALID,
            "internal: bad fused merge descriptor");
  BitsNPtrs pp{};
  for (int i = 0; i < d.n_ops; i++) {
    TSB_CHECK(((uintptr_t)ptrs[i] & 15) == 0, TSB_ERR_INVALID, "internal: bit-packed operands must be 16-byte aligned");
    pp.p[i] = (const uint32_t*)ptrs[i];
    pp.bitoff[i] = bitoffs[i];
  }
  TSB_CHECK(((uintptr_t)o & 15) == 0, TSB_ERR_INVALID, "internal: bit-packed result must be 16-byte aligned");
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 6));
  if (d.n_ops == 2) bits_n_k<2>(ctx, d, pp, o, grid, stream);
  else if (d.n_ops == 3) bits_n_k<3>(ctx, d, pp, o, grid, stream);
  else bits_n_k<4>(ctx, d, pp, o, grid, stream);
  ctx->note("k_merge_bits_n");
  TSB_CUDA(cudaGetLastError());
}

void launch_pack_bits(tsb_ctx* ctx, const void* src, void* dst, int64_t n_entries, cudaStream_t stream) {
  const int64_t n_words = (n_entries + 31) >> 5;
  const int grid = (int)std::min<int64_t>((n_words + 255) / 256, (int64_t)ctx->sm_count * 16);
  TSB_LAUNCH_PDL(k_pack_bits, dim3(std::max(grid, 1)), dim3(256), 0, stream, (const uint8_t*)src, (uint32_t*)dst, n_entries);
  ctx->note("k_pack_bits");
  TSB_CUDA(cudaGetLastError());
}

void launch_pack_map(tsb_ctx* ctx, const void* src, bool src_packed, void* dst, int64_t n_entries, const BitRuns& lo,
                     const BitRuns& hi, cudaStream_t stream) {
  const int64_t n_words = (n_entries + 31) >> 5;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_words + 255) / 256, (int64_t)ctx->sm_count * 16));
  if (src_packed)
    TSB_LAUNCH_PDL(k_pack_map<true>, dim3(grid), dim3(256), 0, stream, src, (uint32_t*)dst, n_entries, lo, hi);
  else
    TSB_LAUNCH_PDL(k_pack_map<false>, dim3(grid), dim3(256), 0, stream, src, (uint32_t*)dst, n_entries, lo, hi);
  ctx->note("k_pack_map");
  TSB_CUDA(cudaGetLastError());
}

void launch_unpack_bits(tsb_ctx* ctx, const void* src, void* dst, int64_t n_entries, bool accumulate,
                        cudaStream_t stream) {
  const int64_t n_vec = (n_entries + 15) >> 4;
  const int grid = (int)std::min<int64_t>((n_vec + 255) / 256, (int64_t)ctx->sm_count * 16);
  TSB_LAUNCH_PDL(k_unpack_bits, dim3(std::max(grid, 1)), dim3(256), 0, stream, (const uint32_t*)src, (uint8_t*)dst,
             n_entries, accumulate ? 1 : 0);
  ctx->note("k_unpack_bits");
  TSB_CUDA(cudaGetLastError());
}

}  // namespace tsb
