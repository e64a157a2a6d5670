// tsb_merge_bits_f.cu — launcher of the factored fused merge and the remaining k_merge_bits_f instantiations.
#include "bits_f_dispatch.h"

namespace tsb {

BitsFKernel bits_f_kernel_small(int nops, int kb, int masks, int n_pair) {
  switch (nops * 4 + kb) {
    case 2 * 4 + 0: return bits_f_kernel<2, 0>(masks, n_pair);
    case 2 * 4 + 1: return bits_f_kernel<2, 1>(masks, n_pair);
    case 2 * 4 + 2: return bits_f_kernel<2, 2>(masks, n_pair);
    case 3 * 4 + 0: return bits_f_kernel<3, 0>(masks, n_pair);
    case 3 * 4 + 1: return bits_f_kernel<3, 1>(masks, n_pair);
    case 4 * 4 + 0: return bits_f_kernel<4, 0>(masks, n_pair);
    case 4 * 4 + 1: return bits_f_kernel<4, 1>(masks, n_pair);
    default: return nullptr;
  }
}

// the per-tile offset tables of a descriptor, [n_ops + 1][2^tile_bits]: what the kernel would deposit per CTA
std::vector<uint32_t> bits_f_tables(const BitsFDesc& d) {
  const int nt = 1 << d.tile_bits;
  std::vector<uint32_t> t((size_t)(d.n_ops + 1) * nt);
  for (int o = 0; o < d.n_ops; o++)
    for (int i = 0; i < nt; i++) t[(size_t)o * nt + i] = (uint32_t)(deposit(d.op[o].lo, (uint64_t)i) >> 7);
  for (int i = 0; i < nt; i++) t[(size_t)d.n_ops * nt + i] = (uint32_t)deposit(d.out_lo, (uint64_t)i);
  return t;
}

void launch_merge_bits_f(tsb_ctx* ctx, const BitsFDesc& d, const void* const* ptrs, const uint64_t* bitoffs, void* o,
                         cudaStream_t stream, const uint32_t* tables) {
  TSB_CHECK(d.n_ops >= 2 && d.n_ops <= kBitsMaxOps && d.kbits >= 0 && d.kbits <= kBitsFMaxKB && d.n_pair >= 0 &&
                d.n_pair <= kBitsFMaxPair, TSB_ERR_INVALID, "internal: bad factored merge descriptor");
  BitsNPtrs pp{};
  for (int s = 0; s < d.n_ops; s++) {
    const int i = d.perm[s];
    TSB_CHECK(((uintptr_t)ptrs[i] & 15) == 0, TSB_ERR_INVALID, "internal: bit-packed operands must be 16-byte aligned");
    pp.p[s] = (const uint32_t*)ptrs[i];
    pp.bitoff[s] = bitoffs[i];
  }
  TSB_CHECK(((uintptr_t)o & 15) == 0, TSB_ERR_INVALID, "internal: bit-packed result must be 16-byte aligned");
  BitsFKernel k = nullptr;
  if (d.n_pair >= 2) {
    TSB_CHECK(bits_f_has_np2(d.n_ops, d.kbits), TSB_ERR_INVALID, "internal: no private-streaming kernel for this step");
    TSB_CHECK(d.n_pair2 == 0 || bits_f_has_two_sided(d.n_ops, d.kbits), TSB_ERR_INVALID, "internal: no two-sided kernel for this step");
    if (d.n_pair2 > 0)
      k = d.n_ops == 3 ? bits_f_kernel_two_3_2(d.masks, d.n_pair, d.n_pair2)
          : d.n_pair2 == 1 ? bits_f_kernel_two1_4_2(d.masks, d.n_pair) : bits_f_kernel_two2_4_2(d.masks, d.n_pair);
    else
      k = d.kbits == 1 ? bits_f_kernel_np2_x_1(d.n_ops, d.masks, d.n_pair)
          : d.n_ops == 4 ? bits_f_kernel_np2_4_2(d.masks, d.n_pair) : bits_f_kernel_np2_3_2(d.masks, d.n_pair);
  } else {
    k = d.n_ops == 4 && d.kbits == 2 ? bits_f_kernel_4_2(d.masks, d.n_pair)
        : d.n_ops == 3 && d.kbits == 2 ? bits_f_kernel_3_2(d.masks, d.n_pair)
                                       : bits_f_kernel_small(d.n_ops, d.kbits, d.masks, d.n_pair);
  }
  TSB_CHECK(k != nullptr, TSB_ERR_INVALID, "internal: no factored merge kernel for masks %d", d.masks);
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  // tuning knobs (A/B runs): persistent CTAs per SM, shared-memory carve-out (percent; the rest of the 256 KB is L1)
  static const int ctas_env = [] { const char* e = getenv("TSB_BITSF_CTAS"); return e ? atoi(e) : 0; }();
  const int per_sm = ctas_env > 0 ? ctas_env : (d.n_pair == 0 ? 3 : 2);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * per_sm));
#ifndef TSB_HOST_EMU
  static const int carve_env = [] { const char* e = getenv("TSB_BITSF_CARVEOUT"); return e ? atoi(e) : -1; }();
  if (carve_env >= 0) cudaFuncSetAttribute((const void*)k, cudaFuncAttributePreferredSharedMemoryCarveout, carve_env);
#endif
  TSB_LAUNCH_PDL(k, dim3(grid), dim3(kBitsFThreads), 0, stream, d, pp, (uint4*)o, tables);
  ctx->note("k_merge_bits_f");
  TSB_CUDA(cudaGetLastError());
}

}  // namespace tsb
