// tsb_merge_sat.cu — instantiations and dispatch of the specialised boolean merge kernel.
#include "k_merge_sat.cuh"
#include "lower.h"

namespace tsb {

bool merge_sat_supported(const MergeDesc& d) {
  if (!d.has_rhs || d.kbits > 8 || d.hiL.n > 32 || d.hiR.n > 32 || d.tile_bits < 2) return false;
  for (const BitRuns* r : {&d.loL, &d.loR})
    for (int i = 0; i < r->n; i++)
      if (r->out_shift[i] + r->len[i] > 32) return false;
  return true;
}

static int covered_bits(const BitRuns& r) {
  int n = 0;
  for (int i = 0; i < r.n; i++) n += r.len[i];
  return n;
}

template <int NL, int NR>
static void launch_one(tsb_ctx* ctx, const MergeSatDesc& dd, const void* l, const void* r, void* o, cudaStream_t stream) {
  if constexpr (((NL | NR) & 15) == 15) {
    const int64_t n_tiles = dd.m.n_vec >> dd.m.tile_bits;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 8);
    auto k = k_merge_sat<NL, NR>;
    TSB_LAUNCH_PDL(k, dim3(grid), dim3(kSatThreads), 0, stream, dd, (const uint8_t*)l, (const uint8_t*)r, (uint8_t*)o);
    ctx->note("k_merge_sat");
    TSB_CUDA(cudaGetLastError());
  } else {
    TSB_CHECK(false, TSB_ERR_INVALID, "internal: output lane covered by neither operand");
  }
}

template <int NL>
static void dispatch_r(tsb_ctx* ctx, const MergeSatDesc& dd, const void* l, const void* r, void* o, cudaStream_t s) {
  switch (dd.m.nibR) {
#define C(n) case n: launch_one<NL, n>(ctx, dd, l, r, o, s); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    default: TSB_CHECK(false, TSB_ERR_INVALID, "internal: bad lane mask");
  }
}

void launch_merge_sat(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t s) {
  MergeSatDesc dd;
  dd.m = d;
  dd.stream_l = covered_bits(d.loL) + covered_bits(d.hiL) + popc4(d.nibL) == d.R;
  dd.stream_r = covered_bits(d.loR) + covered_bits(d.hiR) + popc4(d.nibR) == d.R;
  // choose the two tile-local bits that save the most load bytes: a bit an operand does not depend on
  // lets the thread reuse that operand's vector for both values of the bit
  auto depends = [](const BitRuns& r, int bit) {
    for (int i = 0; i < r.n; i++)
      if (bit >= r.in_shift[i] && bit < r.in_shift[i] + r.len[i]) return true;
    return false;
  };
  const int costL = 1 << popc4(d.nibL), costR = 1 << popc4(d.nibR);
  int best[2] = {-1, -1}, best_s[2] = {-1, -1};
  for (int bit = 0; bit < d.tile_bits; bit++) {
    const int sv = (depends(d.loL, bit) ? 0 : costL) + (depends(d.loR, bit) ? 0 : costR);
    if (sv > best_s[0]) { best[1] = best[0]; best_s[1] = best_s[0]; best[0] = bit; best_s[0] = sv; }
    else if (sv > best_s[1]) { best[1] = bit; best_s[1] = sv; }
  }
  dd.bit_a = std::min(best[0], best[1]);
  dd.bit_b = std::max(best[0], best[1]);
  switch (d.nibL) {
#define C(n) case n: dispatch_r<n>(ctx, dd, l, r, o, s); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    default: TSB_CHECK(false, TSB_ERR_INVALID, "internal: bad lane mask");
  }
}

}  // namespace tsb
