// tsb_merge_cnt_48.cu — instantiations of k_merge_cnt for operand types (int32_t, uint8_t).
#include "lower.h"
#include "k_merge_cnt.cuh"

namespace tsb {

template <int NL, int NR>
static void cnt_one(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t stream) {
  if constexpr (((NL | NR) & 15) == 15) {
    const int64_t n_tiles = d.n_vec >> d.tile_bits;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 6);
    auto k = k_merge_cnt<NL, NR, int32_t, uint8_t>;
    TSB_LAUNCH(k, dim3(grid), dim3(kCntThreads), 0, stream, d, (const int32_t*)l, (const uint8_t*)r, (int32_t*)o);
    ctx->note("k_merge_cnt");
    TSB_CUDA(cudaGetLastError());
  } else {
    TSB_CHECK(false, TSB_ERR_INVALID, "internal: output lane covered by neither operand");
  }
}

template <int NL>
static void cnt_r(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t s) {
  switch (d.nibR) {
#define C(n) case n: cnt_one<NL, n>(ctx, d, l, r, o, s); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    default: TSB_CHECK(false, TSB_ERR_INVALID, "internal: bad lane mask");
  }
}

template <>
void launch_merge_cnt_t<int32_t, uint8_t>(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t s) {
  switch (d.nibL) {
#define C(n) case n: cnt_r<n>(ctx, d, l, r, o, s); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    default: TSB_CHECK(false, TSB_ERR_INVALID, "internal: bad lane mask");
  }
}

}  // namespace tsb
