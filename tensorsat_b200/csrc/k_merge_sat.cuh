// k_merge_sat.cuh — the boolean-semiring monotone merge, specialised at compile time on which of
// the 16 output lanes of a vector each operand depends on (NIBL / NIBR, see k_merge.cuh for the
// scheme).  This is the kernel that carries > 99 % of the bytes of a SAT contraction plan
// (FinRel._contract2's einsum + sign, tensorsat/lang/fin_rel.py:276-279).
//
// Per 16-byte output vector and summed-label combination a thread issues one vector load per
// operand (1..16 bytes, compile-time width), a compile-time byte expansion (0..4 PRMT / IMUL),
// and 4 LOP3; offsets come from 32-bit shared-memory tables.  Streaming operands / results use
// evict-first cache hints so that the reused operand stays in L1/L2.  Bound: HBM.
#pragma once
#include "k_merge.cuh"

namespace tsb {

constexpr int kSatThreads = 256;

struct MergeSatDesc {
  MergeDesc m;
  int32_t stream_l, stream_r;  // operand bytes are used once: load with evict-first
};

constexpr int popc4(int x) { return (x & 1) + ((x >> 1) & 1) + ((x >> 2) & 1) + ((x >> 3) & 1); }

#ifdef TSB_HOST_EMU
template <typename T> TSB_D T ld_stream(const T* p) { return *p; }
template <typename T> TSB_D void st_stream(T* p, T v) { *p = v; }
#else
template <typename T> TSB_D T ld_stream(const T* p) { return __ldcs(p); }
template <typename T> TSB_D void st_stream(T* p, T v) { __stcs(p, v); }
#endif

// Loads the 2^popc(NIB) operand bytes one output vector touches and expands them to 16 lanes.
template <int NIB, bool STREAM>
TSB_D void load_expand(const uint8_t* p, uint32_t X[4]) {
  constexpr int VB = popc4(NIB), p0 = NIB & 3, p1 = NIB >> 2;
  uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  if constexpr (VB == 0) b0 = STREAM ? ld_stream(p) : *p;
  else if constexpr (VB == 1) b0 = STREAM ? ld_stream((const uint16_t*)p) : *(const uint16_t*)p;
  else if constexpr (VB == 2) b0 = STREAM ? ld_stream((const uint32_t*)p) : *(const uint32_t*)p;
  else if constexpr (VB == 3) { uint2 t = STREAM ? ld_stream((const uint2*)p) : *(const uint2*)p; b0 = t.x; b1 = t.y; }
  else { uint4 t = STREAM ? ld_stream((const uint4*)p) : *(const uint4*)p; b0 = t.x; b1 = t.y; b2 = t.z; b3 = t.w; }
  // groups: the operand bytes selected by lanes 4g..4g+3 of the (at most 4) distinct groups
  uint32_t G0, G1 = 0, G2 = 0, G3 = 0;
  constexpr int NG = 1 << popc4(p1);  // distinct groups needed
  if constexpr (p0 == 3) {
    G0 = b0; G1 = b1; G2 = b2; G3 = b3;
  } else if constexpr (p0 == 0) {
    G0 = (b0 & 0xffu) * 0x01010101u;
    if constexpr (NG > 1) G1 = ((b0 >> 8) & 0xffu) * 0x01010101u;
    if constexpr (NG > 2) { G2 = ((b0 >> 16) & 0xffu) * 0x01010101u; G3 = (b0 >> 24) * 0x01010101u; }
  } else if constexpr (p0 == 1) {
    G0 = __byte_perm(b0, 0u, 0x1010u);
    if constexpr (NG > 1) G1 = __byte_perm(b0, 0u, 0x3232u);
    if constexpr (NG > 2) { G2 = __byte_perm(b1, 0u, 0x1010u); G3 = __byte_perm(b1, 0u, 0x3232u); }
  } else {
    G0 = __byte_perm(b0, 0u, 0x1100u);
    if constexpr (NG > 1) G1 = __byte_perm(b0, 0u, 0x3322u);
    if constexpr (NG > 2) { G2 = __byte_perm(b1, 0u, 0x1100u); G3 = __byte_perm(b1, 0u, 0x3322u); }
  }
  if constexpr (p1 == 3) { X[0] = G0; X[1] = G1; X[2] = G2; X[3] = G3; }
  else if constexpr (p1 == 0) { X[0] = X[1] = X[2] = X[3] = G0; }
  else if constexpr (p1 == 1) { X[0] = G0; X[1] = G1; X[2] = G0; X[3] = G1; }
  else { X[0] = X[1] = G0; X[2] = X[3] = G1; }
}

template <int NIBL, int NIBR>
TSB_D void merge_sat_body(const MergeSatDesc& dd, const uint8_t* __restrict__ lhs, const uint8_t* __restrict__ rhs,
                          uint8_t* __restrict__ out, const uint32_t* tabL, const uint32_t* tabR, const int64_t* kL,
                          const int64_t* kR) {
  const MergeDesc& d = dd.m;
  const int nt = 1 << d.tile_bits;
  const int K = 1 << d.kbits;
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint8_t* baseL = lhs + deposit(d.hiL, (uint64_t)tile);
    const uint8_t* baseR = rhs + deposit(d.hiR, (uint64_t)tile);
    uint4* outp = (uint4*)out + (tile << d.tile_bits);
#pragma unroll 2
    for (int vi = threadIdx.x; vi < nt; vi += kSatThreads) {
      const uint8_t* pl = baseL + tabL[vi];
      const uint8_t* pr = baseR + tabR[vi];
      uint32_t a0, a1, a2, a3;
      {
        uint32_t XL[4], XR[4];
        load_expand<NIBL, false>(pl, XL);
        load_expand<NIBR, false>(pr, XR);
        a0 = XL[0] & XR[0]; a1 = XL[1] & XR[1]; a2 = XL[2] & XR[2]; a3 = XL[3] & XR[3];
      }
      for (int kk = 1; kk < K; kk++) {
        uint32_t XL[4], XR[4];
        load_expand<NIBL, false>(pl + kL[kk], XL);
        load_expand<NIBR, false>(pr + kR[kk], XR);
        a0 |= XL[0] & XR[0]; a1 |= XL[1] & XR[1]; a2 |= XL[2] & XR[2]; a3 |= XL[3] & XR[3];
      }
      if (d.accumulate) { uint4 p = outp[vi]; a0 |= p.x; a1 |= p.y; a2 |= p.z; a3 |= p.w; }
      st_stream(outp + vi, make_uint4(a0, a1, a2, a3));
    }
  }
}

template <int NIBL, int NIBR>
__global__ void __launch_bounds__(kSatThreads)
k_merge_sat(const __grid_constant__ MergeSatDesc dd, const uint8_t* __restrict__ lhs, const uint8_t* __restrict__ rhs,
            uint8_t* __restrict__ out) {
  __shared__ uint32_t tabL[1 << kMergeTileBits], tabR[1 << kMergeTileBits];
  __shared__ int64_t kL[256], kR[256];
  const MergeDesc& d = dd.m;
  const int nt = 1 << d.tile_bits;
  for (int i = threadIdx.x; i < nt; i += kSatThreads) {
    tabL[i] = (uint32_t)deposit(d.loL, (uint64_t)i);
    tabR[i] = (uint32_t)deposit(d.loR, (uint64_t)i);
  }
  for (int i = threadIdx.x; i < (1 << d.kbits); i += kSatThreads) {
    int64_t a = 0, b = 0;
    for (int q = 0; q < d.kbits; q++)
      if (i >> q & 1) { a += d.ksL[q]; b += d.ksR[q]; }
    kL[i] = a;
    kR[i] = b;
  }
  __syncthreads();
  merge_sat_body<NIBL, NIBR>(dd, lhs, rhs, out, tabL, tabR, kL, kR);
}

// Preconditions checked by the host: has_rhs, kbits <= 8, table offsets < 2^32.
bool merge_sat_supported(const MergeDesc& d);
void launch_merge_sat(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t stream);

}  // namespace tsb
