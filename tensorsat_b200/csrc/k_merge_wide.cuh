// k_merge_wide.cuh — the counting-semiring monotone merge with int64 results.
//
// Exact #SAT over more than ~30 free variables needs 64-bit counts (the notebook prototype of the
// reference multiplies and adds Python-width numpy integers:
// notebooks/prototypes/2025-02-13-#sat-tensor-contraction.ipynb:215); the middle of such a plan is a
// chain of "int64 intermediate x small uint8/int32 factor" merges that stream 8 B in and 8 B out per
// element.  Same MergeDesc as the other merge kernels (16 output elements per vector), but with
// 8-byte elements the thread mapping changes: one thread produces TWO consecutive outputs (one
// 16-byte store), eight threads share a vector, so a warp writes 512 contiguous bytes and every
// operand load is either one aligned element pair (output bit 0 is an operand label) or one element
// used twice.  Bound: HBM (algorithmic bytes = |lhs| + |rhs| + |out|, 1 IMAD.WIDE chain per output).
#pragma once
#include "k_merge_sat.cuh"

namespace tsb {

constexpr int kWideThreads = 256;
constexpr int kWideG = 2;  // output pairs per thread in flight

struct alignas(16) I64x2 { int64_t x, y; };
struct alignas(8) I32x2 { int32_t x, y; };

template <typename T>
TSB_D void load_pair(const T* p, bool two, int64_t& a, int64_t& b) {
  if (two) {
    if constexpr (sizeof(T) == 8) { const I64x2 t = *(const I64x2*)p; a = t.x; b = t.y; }
    else if constexpr (sizeof(T) == 4) { const I32x2 t = *(const I32x2*)p; a = t.x; b = t.y; }
    else { const uint32_t t = *(const uint16_t*)p; a = (int64_t)(t & 0xffu); b = (int64_t)(t >> 8); }
  } else {
    a = b = (int64_t)*p;
  }
}

TSB_HD int pext3(int e, int nib3) {
  int r = 0, o = 0;
  for (int b = 0; b < 3; b++)
    if (nib3 >> b & 1) { r |= ((e >> b) & 1) << o; o++; }
  return r;
}

template <typename TL, typename TR>
__global__ void __launch_bounds__(kWideThreads, 4)
k_merge_wide(const __grid_constant__ MergeDesc d, const TL* __restrict__ lhs, const TR* __restrict__ rhs,
             int64_t* __restrict__ out) {
  __shared__ uint32_t tabL[1 << kMergeTileBits], tabR[1 << kMergeTileBits];
  __shared__ int64_t kL[256], kR[256];
  __shared__ RunsSmem runsL, runsR;
  const int nt = 1 << d.tile_bits;
  for (int i = threadIdx.x; i < nt; i += kWideThreads) {
    tabL[i] = (uint32_t)deposit(d.loL, (uint64_t)i);
    tabR[i] = (uint32_t)deposit(d.loR, (uint64_t)i);
  }
  for (int i = threadIdx.x; i < (1 << d.kbits); i += kWideThreads) {
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
  const bool acc_out = d.accumulate != 0, has_rhs = d.has_rhs != 0;
  // this thread's output pair inside a vector and the operand elements it reads
  const int sub = threadIdx.x & 7;
  const bool twoL = (d.nibL & 1) != 0, twoR = (d.nibR & 1) != 0;
  const int subL = pext3(sub, d.nibL >> 1) << (twoL ? 1 : 0);
  const int subR = pext3(sub, d.nibR >> 1) << (twoR ? 1 : 0);
  const int n_pairs = nt << 3;
  int64_t tile = blockIdx.x;
  uint64_t offL = tile < n_tiles ? deposit_warp(runsL, (uint64_t)tile) : 0;
  uint64_t offR = tile < n_tiles ? deposit_warp(runsR, (uint64_t)tile) : 0;
  while (tile < n_tiles) {
    const TL* baseL = lhs + offL + subL;
    const TR* baseR = rhs + offR + subR;
    I64x2* outp = (I64x2*)(out + ((tile << d.tile_bits) << 4));
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      offL = deposit_warp(runsL, (uint64_t)next);
      offR = deposit_warp(runsR, (uint64_t)next);
    }
    for (int i0 = threadIdx.x; i0 < n_pairs; i0 += kWideThreads * kWideG) {
      const TL* pl[kWideG];
      const TR* pr[kWideG];
      bool on[kWideG];
      int64_t a0[kWideG], a1[kWideG];
#pragma unroll
      for (int g = 0; g < kWideG; g++) {
        const int idx = i0 + g * kWideThreads;
        on[g] = idx < n_pairs;
        const int vi = on[g] ? idx >> 3 : 0;
        pl[g] = baseL + tabL[vi];
        pr[g] = baseR + tabR[vi];
        a0[g] = a1[g] = 0;
      }
      if (acc_out) {
#pragma unroll
        for (int g = 0; g < kWideG; g++)
          if (on[g]) { const I64x2 p = outp[i0 + g * kWideThreads]; a0[g] = p.x; a1[g] = p.y; }
      }
      for (int kk = 0; kk < K; kk++) {
        const int64_t sl = kL[kk], sr = kR[kk];
        int64_t l0[kWideG], l1[kWideG], r0[kWideG], r1[kWideG];
#pragma unroll
        for (int g = 0; g < kWideG; g++) {
          l0[g] = l1[g] = 0; r0[g] = r1[g] = 1;
          if (on[g]) {
            load_pair<TL>(pl[g] + sl, twoL, l0[g], l1[g]);
            if (has_rhs) load_pair<TR>(pr[g] + sr, twoR, r0[g], r1[g]);
          }
        }
#pragma unroll
        for (int g = 0; g < kWideG; g++) { a0[g] += l0[g] * r0[g]; a1[g] += l1[g] * r1[g]; }
      }
#pragma unroll
      for (int g = 0; g < kWideG; g++)
        if (on[g]) outp[i0 + g * kWideThreads] = I64x2{a0[g], a1[g]};
    }
    tile = next;
  }
}

}  // namespace tsb
