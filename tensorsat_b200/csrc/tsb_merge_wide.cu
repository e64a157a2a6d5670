// tsb_merge_wide.cu — instantiations of k_merge_wide (int64 results) for the nine operand type pairs.
#include "lower.h"
#include "k_merge_wide.cuh"

namespace tsb {

template <typename TL, typename TR>
static void wide_one(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t stream) {
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 8);
  auto k = k_merge_wide<TL, TR>;
  TSB_LAUNCH(k, dim3(grid), dim3(kWideThreads), 0, stream, d, (const TL*)l, (const TR*)r, (int64_t*)o);
  ctx->note("k_merge_wide");
  TSB_CUDA(cudaGetLastError());
}

template <typename TL>
static void wide_r(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, int dr, void* o, cudaStream_t s) {
  if (dr == TSB_U8) wide_one<TL, uint8_t>(ctx, d, l, r, o, s);
  else if (dr == TSB_I32) wide_one<TL, int32_t>(ctx, d, l, r, o, s);
  else wide_one<TL, int64_t>(ctx, d, l, r, o, s);
}

bool merge_wide_supported(const MergeDesc& d, const void* l, int dl, const void* r, int dr, const void* o) {
  if (d.kbits > 8) return false;
  const uintptr_t al = (uintptr_t)dtype_size(dl) * ((d.nibL & 1) ? 2 : 1) - 1;
  const uintptr_t ar = (uintptr_t)dtype_size(dr) * ((d.nibR & 1) ? 2 : 1) - 1;
  if ((uintptr_t)l & al) return false;
  if (d.has_rhs && ((uintptr_t)r & ar)) return false;
  return ((uintptr_t)o & 15) == 0;
}

void launch_merge_wide(tsb_ctx* ctx, const MergeDesc& d, const void* l, int dl, const void* r, int dr, void* o,
                       cudaStream_t s) {
  if (!d.has_rhs) dr = TSB_U8;
  if (dl == TSB_U8) wide_r<uint8_t>(ctx, d, l, r, dr, o, s);
  else if (dl == TSB_I32) wide_r<int32_t>(ctx, d, l, r, dr, o, s);
  else wide_r<int64_t>(ctx, d, l, r, dr, o, s);
}

}  // namespace tsb
