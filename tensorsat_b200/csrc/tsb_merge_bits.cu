// tsb_merge_bits.cu — launchers of the bit-packed boolean merge and of the byte <-> bit conversions.
#include "lower.h"
#include "k_merge_bits.cuh"

namespace tsb {

void launch_merge_bits(tsb_ctx* ctx, const MergeDesc& d, const void* l, uint64_t bitoff_l, const void* r,
                       uint64_t bitoff_r, void* o, cudaStream_t stream) {
  TSB_CHECK(d.kbits <= 8, TSB_ERR_UNSUPPORTED, "internal: bit-packed merge with more than 8 summed wires");
  TSB_CHECK(((uintptr_t)l & 15) == 0 && (!d.has_rhs || ((uintptr_t)r & 15) == 0) && ((uintptr_t)o & 15) == 0, TSB_ERR_INVALID,
            "internal: bit-packed operands must be 16-byte aligned");
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 8);
  if (d.kbits == 0)
    TSB_LAUNCH(k_merge_bits<0>, dim3(grid), dim3(kBitsThreads), 0, stream, d, (const uint32_t*)l, bitoff_l,
               (const uint32_t*)r, bitoff_r, (uint4*)o);
  else if (d.kbits == 1)
    TSB_LAUNCH(k_merge_bits<1>, dim3(grid), dim3(kBitsThreads), 0, stream, d, (const uint32_t*)l, bitoff_l,
               (const uint32_t*)r, bitoff_r, (uint4*)o);
  else
    TSB_LAUNCH(k_merge_bits<-1>, dim3(grid), dim3(kBitsThreads), 0, stream, d, (const uint32_t*)l, bitoff_l,
               (const uint32_t*)r, bitoff_r, (uint4*)o);
  ctx->launches++;
  TSB_CUDA(cudaGetLastError());
}

void launch_pack_bits(tsb_ctx* ctx, const void* src, void* dst, int64_t n_entries, cudaStream_t stream) {
  const int64_t n_words = (n_entries + 31) >> 5;
  const int grid = (int)std::min<int64_t>((n_words + 255) / 256, (int64_t)ctx->sm_count * 16);
  TSB_LAUNCH(k_pack_bits, dim3(std::max(grid, 1)), dim3(256), 0, stream, (const uint8_t*)src, (uint32_t*)dst, n_entries);
  ctx->launches++;
  TSB_CUDA(cudaGetLastError());
}

void launch_pack_map(tsb_ctx* ctx, const void* src, bool src_packed, void* dst, int64_t n_entries, const BitRuns& lo,
                     const BitRuns& hi, cudaStream_t stream) {
  const int64_t n_words = (n_entries + 31) >> 5;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_words + 255) / 256, (int64_t)ctx->sm_count * 16));
  if (src_packed)
    TSB_LAUNCH(k_pack_map<true>, dim3(grid), dim3(256), 0, stream, src, (uint32_t*)dst, n_entries, lo, hi);
  else
    TSB_LAUNCH(k_pack_map<false>, dim3(grid), dim3(256), 0, stream, src, (uint32_t*)dst, n_entries, lo, hi);
  ctx->launches++;
  TSB_CUDA(cudaGetLastError());
}

void launch_unpack_bits(tsb_ctx* ctx, const void* src, void* dst, int64_t n_entries, bool accumulate,
                        cudaStream_t stream) {
  const int64_t n_vec = (n_entries + 15) >> 4;
  const int grid = (int)std::min<int64_t>((n_vec + 255) / 256, (int64_t)ctx->sm_count * 16);
  TSB_LAUNCH(k_unpack_bits, dim3(std::max(grid, 1)), dim3(256), 0, stream, (const uint32_t*)src, (uint8_t*)dst,
             n_entries, accumulate ? 1 : 0);
  ctx->launches++;
  TSB_CUDA(cudaGetLastError());
}

}  // namespace tsb
