// k_merge_cnt.cuh — the counting-semiring monotone merge with int32 results, specialised at compile
// time on the lane masks (NIBL / NIBR) and on the operand types (u8 / i32).  Same scheme as
// k_merge_sat.cuh; the per-lane operand selection is a compile-time register choice, the arithmetic
// is one IMAD per lane and summed-label combination.  Replaces the un-clamped integer einsum of the
// #SAT prototype (notebooks/prototypes/2025-02-13-#sat-tensor-contraction.ipynb:215) on the large
// middle of a counting plan, where intermediates hold counts < 2^31 (plan-time bound).  Bound: HBM.
#pragma once
#include "k_merge_sat.cuh"

namespace tsb {

constexpr int kCntThreads = 256;

constexpr int pext4_c(int e, int nib) {
  int r = 0, o = 0;
  for (int b = 0; b < 4; b++)
    if ((nib >> b) & 1) { r |= ((e >> b) & 1) << o; o++; }
  return r;
}

// the 2^popc(NIB) operand elements one output vector touches
template <typename T, int NIB> struct RawE { int32_t v[1 << popc4(NIB)]; };

template <typename T, int NIB>
TSB_D RawE<T, NIB> raw_load_e(const T* p) {
  constexpr int VB = popc4(NIB);
  RawE<T, NIB> r;
  if constexpr (sizeof(T) == 1) {
    Raw<NIB> b = raw_load<NIB>((const uint8_t*)p, 0);
#pragma unroll
    for (int i = 0; i < (1 << VB); i++) r.v[i] = (int32_t)((b.w[i >> 2] >> (8 * (i & 3))) & 0xffu);
  } else {
    if constexpr (VB == 0) r.v[0] = *p;
    else if constexpr (VB == 1) { uint2 t = *(const uint2*)p; r.v[0] = (int32_t)t.x; r.v[1] = (int32_t)t.y; }
    else {
#pragma unroll
      for (int q = 0; q < (1 << VB) / 4; q++) {
        uint4 t = ((const uint4*)p)[q];
        r.v[4 * q] = (int32_t)t.x; r.v[4 * q + 1] = (int32_t)t.y; r.v[4 * q + 2] = (int32_t)t.z; r.v[4 * q + 3] = (int32_t)t.w;
      }
    }
  }
  return r;
}

template <int NIBL, int NIBR, typename TL, typename TR>
__global__ void __launch_bounds__(kCntThreads, 3)
k_merge_cnt(const __grid_constant__ MergeDesc d, const TL* __restrict__ lhs, const TR* __restrict__ rhs,
            int32_t* __restrict__ out) {
  __shared__ uint32_t tabL[1 << kMergeTileBits], tabR[1 << kMergeTileBits];
  __shared__ int64_t kL[256], kR[256];
  __shared__ RunsSmem runsL, runsR;
  const int nt = 1 << d.tile_bits;
  for (int i = threadIdx.x; i < nt; i += kCntThreads) {
    tabL[i] = (uint32_t)deposit(d.loL, (uint64_t)i);
    tabR[i] = (uint32_t)deposit(d.loR, (uint64_t)i);
  }
  for (int i = threadIdx.x; i < (1 << d.kbits); i += kCntThreads) {
    int64_t a = 0, b = 0;
    for (int q = 0; q < d.kbits; q++)
      if (i >> q & 1) { a += d.ksL[q]; b += d.ksR[q]; }
    kL[i] = a;
    kR[i] = b;
  }
  runs_to_smem(d.hiL, runsL);
  runs_to_smem(d.hiR, runsR);
  __syncthreads();
  const int K = 1 << d.kbits;
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const bool acc_out = d.accumulate != 0;
  int64_t tile = blockIdx.x;
  uint64_t offL = tile < n_tiles ? deposit_warp(runsL, (uint64_t)tile) : 0;
  uint64_t offR = tile < n_tiles ? deposit_warp(runsR, (uint64_t)tile) : 0;
  while (tile < n_tiles) {
    const TL* baseL = lhs + offL;
    const TR* baseR = rhs + offR;
    uint4* outp = (uint4*)(out + ((tile << d.tile_bits) << 4));
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      offL = deposit_warp(runsL, (uint64_t)next);
      offR = deposit_warp(runsR, (uint64_t)next);
    }
    for (int vi = threadIdx.x; vi < nt; vi += kCntThreads) {
      const TL* pl = baseL + tabL[vi];
      const TR* pr = baseR + tabR[vi];
      int32_t a[16];
#pragma unroll
      for (int e = 0; e < 16; e++) a[e] = 0;
#pragma unroll 2
      for (int kk = 0; kk < K; kk++) {
        const RawE<TL, NIBL> rl = raw_load_e<TL, NIBL>(pl + kL[kk]);
        const RawE<TR, NIBR> rr = raw_load_e<TR, NIBR>(pr + kR[kk]);
#pragma unroll
        for (int e = 0; e < 16; e++) a[e] += rl.v[pext4_c(e, NIBL)] * rr.v[pext4_c(e, NIBR)];
      }
      uint4* dst = outp + (vi << 2);
      if (acc_out) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint4 p = dst[q];
          a[4 * q] += (int32_t)p.x; a[4 * q + 1] += (int32_t)p.y; a[4 * q + 2] += (int32_t)p.z; a[4 * q + 3] += (int32_t)p.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        dst[q] = make_uint4((uint32_t)a[4 * q], (uint32_t)a[4 * q + 1], (uint32_t)a[4 * q + 2], (uint32_t)a[4 * q + 3]);
    }
    tile = next;
  }
}

// launchers: one translation unit per operand-type pair (build time); declared in lower.h

}  // namespace tsb
