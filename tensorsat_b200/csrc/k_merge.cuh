// k_merge.cuh — the hot kernel: pairwise contraction of binary-extent tensors whose free
// labels appear in the same relative order in lhs, rhs and out ("monotone merge").
//
// Replaces the integer einsum + np.sign of FinRel._contract2 (tensorsat/lang/fin_rel.py:276-279)
// for every step a plan executes: plans keep all tensors in one global label order, so no step
// ever needs a transposed copy (the reference's opt_einsum call re-lays operands out on every
// step).  Per output vector of 16 consecutive entries a thread
//   * finds the lhs / rhs offsets by depositing the output index bits (tables for the low
//     bits, run-wise shifts for the tile index),
//   * loads the 1..16 distinct operand bytes those 16 outputs touch with ONE vector load each,
//   * expands them to 16 lanes with byte permutes, and combines with LOP3 (boolean semiring)
//     or integer MACs (counting semiring),
//   * loops over the summed labels, and writes one 16-byte (or wider) coalesced vector.
// Algorithmic traffic: every operand byte and every output byte crosses HBM once
// (operand re-reads across tiles are served by L1/L2).  Bound: HBM.
#pragma once
#include "tsb_common.h"

namespace tsb {

constexpr int kMaxRuns = 40;
constexpr int kMergeMaxK = 16;     // summed labels handled inside one launch
constexpr int kMergeTileBits = 10; // vectors per tile = 2^tile_bits
constexpr int kMergeThreads = 128;

struct BitRuns {  // address = OR over runs of ((x >> in_shift) & ((1 << len) - 1)) << out_shift
  int32_t n;
  uint8_t in_shift[kMaxRuns], len[kMaxRuns], out_shift[kMaxRuns];
};

TSB_HD uint64_t deposit(const BitRuns& r, uint64_t x) {
  uint64_t a = 0;
  for (int i = 0; i < r.n; i++)
    a |= ((x >> r.in_shift[i]) & ((uint64_t(1) << r.len[i]) - 1)) << r.out_shift[i];
  return a;
}

#ifndef TSB_D
#define TSB_D __device__ __forceinline__
#endif

// Per-tile operand base: the tile-index bits are deposited run by run; one run per lane, OR-reduced
// across the warp (every warp computes its own copy, no block barrier needed).
struct RunsSmem {
  uint32_t in_shift[32], out_shift[32];
  uint64_t mask[32];
};

TSB_D void runs_to_smem(const BitRuns& r, RunsSmem& s) {
  for (int i = threadIdx.x; i < 32; i += blockDim.x) {
    const bool on = i < r.n;
    s.in_shift[i] = on ? r.in_shift[i] : 0;
    s.out_shift[i] = on ? r.out_shift[i] : 0;
    s.mask[i] = on ? ((uint64_t(1) << r.len[i]) - 1) : 0;
  }
}

TSB_D uint64_t deposit_warp(const RunsSmem& s, uint64_t x) {
  const int lane = threadIdx.x & 31;
  const uint64_t c = ((x >> s.in_shift[lane]) & s.mask[lane]) << s.out_shift[lane];
  const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)c);
  const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(c >> 32));
  return ((uint64_t)hi << 32) | lo;
}

struct MergeDesc {
  int32_t R;          // output rank in bits (>= 4)
  int32_t kbits;      // summed labels
  int32_t nibL, nibR; // which of output bits 0..3 are lhs / rhs labels
  int32_t has_rhs;
  int32_t accumulate; // combine with the value already in out (slice accumulation)
  int32_t tile_bits;
  BitRuns loL, loR;   // vector-in-tile index (output bits 4..4+tile_bits) -> operand offset
  BitRuns hiL, hiR;   // tile index -> operand offset
  int64_t ksL[kMergeMaxK], ksR[kMergeMaxK];  // operand stride of each summed label
  int64_t n_vec;      // 2^(R-4)
};

// descriptor of the fused n-ary bit-packed merge (k_merge_bits.cuh: k_merge_bits_n)
constexpr int kBitsMaxOps = 4;
constexpr int kBitsMaxKB = 3;

struct BitsOperand {
  BitRuns lo, hi;               // vector-in-tile index / tile index -> operand offset (entries)
  int64_t ks[kBitsMaxKB];       // operand stride of each summed wire (0: operand does not hold it)
  int32_t hi2;                  // which of output bits 5, 6 the operand holds
  int32_t pad;
};

struct BitsNDesc {
  int32_t R, kbits, n_ops, tile_bits;
  int64_t n_vec;
  BitsOperand op[kBitsMaxOps];
};

struct BitsNPtrs {
  const uint32_t* p[kBitsMaxOps];
  uint64_t bitoff[kBitsMaxOps];
};

// descriptor of the factored fused merge (k_merge_bits_f.cuh): summed-wire masks known at compile time,
// optionally one "pair" output wire handled per thread so that operands that do not hold it are fetched once
constexpr int kBitsFMaxKB = 2;
constexpr int kBitsFMaxPair = 3;

struct BitsFOperand {
  BitRuns lo, hi;   // reduced vector-in-tile index / reduced tile index -> operand offset (entries)
  uint32_t kw[4];   // offset, in 16-byte units, of summed-wire combination kk (wires the operand lacks add 0)
  uint32_t pw[8];   // offset, in 16-byte units, of combination g of the pair wires THIS operand holds
  int32_t kmask;    // summed wires the operand holds
  int32_t held;     // pair wires the operand holds
};

struct BitsFDesc {
  int32_t R, kbits, n_ops, tile_bits, n_pair, masks;
  int32_t n_pair2, pad_;      // private streaming: pair wires private to slot 1 (two-sided form), else 0
  int32_t perm[kBitsMaxOps];  // slot -> index of the caller's operand (slots are sorted by kmask, descending)
  int64_t n_vec;              // work items: 2^(R - 7 - n_pair - n_pair2)
  BitRuns out_lo, out_hi;     // reduced index -> output vector index (the pair wires are holes)
  uint32_t ow[8];             // output vector offset of pair-wire combination g (private streaming: of slot 0's wires)
  uint32_t ow2[4];            // ... of combination h of slot 1's private wires
  BitsFOperand op[kBitsMaxOps];
};

template <typename T> struct VecLoad;

TSB_D void load_sub_u8(const uint8_t* p, int vbits, uint32_t buf[4]) {
  buf[0] = buf[1] = buf[2] = buf[3] = 0;
  switch (vbits) {
    case 0: buf[0] = *p; break;
    case 1: buf[0] = *(const uint16_t*)p; break;
    case 2: buf[0] = *(const uint32_t*)p; break;
    case 3: { uint2 t = *(const uint2*)p; buf[0] = t.x; buf[1] = t.y; } break;
    default: { uint4 t = *(const uint4*)p; buf[0] = t.x; buf[1] = t.y; buf[2] = t.z; buf[3] = t.w; } break;
  }
}

// X byte e (0..15) = buf byte pext(e, nib): the operand byte that output lane e reads.
TSB_D void expand16(const uint32_t buf[4], int nib, uint32_t X[4]) {
  const int p0 = nib & 3, p1 = nib >> 2;
  uint32_t G0, G1, G2, G3;
  switch (p0) {
    case 3: G0 = buf[0]; G1 = buf[1]; G2 = buf[2]; G3 = buf[3]; break;
    case 0:
      G0 = (buf[0] & 0xffu) * 0x01010101u;
      G1 = ((buf[0] >> 8) & 0xffu) * 0x01010101u;
      G2 = ((buf[0] >> 16) & 0xffu) * 0x01010101u;
      G3 = (buf[0] >> 24) * 0x01010101u;
      break;
    case 1:
      G0 = (buf[0] & 0xffffu) * 0x00010001u;
      G1 = (buf[0] >> 16) * 0x00010001u;
      G2 = (buf[1] & 0xffffu) * 0x00010001u;
      G3 = (buf[1] >> 16) * 0x00010001u;
      break;
    default:  // p0 == 2
      G0 = __byte_perm(buf[0], 0u, 0x1100u);
      G1 = __byte_perm(buf[0], 0u, 0x3322u);
      G2 = __byte_perm(buf[1], 0u, 0x1100u);
      G3 = __byte_perm(buf[1], 0u, 0x3322u);
      break;
  }
  switch (p1) {
    case 3: X[0] = G0; X[1] = G1; X[2] = G2; X[3] = G3; break;
    case 0: X[0] = X[1] = X[2] = X[3] = G0; break;
    case 1: X[0] = G0; X[1] = G1; X[2] = G0; X[3] = G1; break;
    default: X[0] = X[1] = G0; X[2] = X[3] = G1; break;
  }
}

TSB_HD int pext4(int e, int nib) {
  int r = 0, o = 0;
  for (int b = 0; b < 4; b++)
    if (nib >> b & 1) { r |= ((e >> b) & 1) << o; o++; }
  return r;
}

template <typename T> struct AccOf { using type = int32_t; };
template <> struct AccOf<int64_t> { using type = int64_t; };

template <typename TL, typename TR, typename TO, int CLAMP>
__global__ void __launch_bounds__(kMergeThreads)
k_merge(const __grid_constant__ MergeDesc d, const TL* __restrict__ lhs, const TR* __restrict__ rhs,
        TO* __restrict__ out) {
  __shared__ uint64_t tabL[1 << kMergeTileBits], tabR[1 << kMergeTileBits];
  __shared__ int64_t kL[256], kR[256];
  const int nt = 1 << d.tile_bits;
  for (int i = threadIdx.x; i < nt; i += blockDim.x) {
    tabL[i] = deposit(d.loL, (uint64_t)i);
    tabR[i] = deposit(d.loR, (uint64_t)i);
  }
  const int klo = d.kbits < 8 ? d.kbits : 8;
  const int khi = d.kbits - klo;
  for (int i = threadIdx.x; i < (1 << klo); i += blockDim.x) {
    int64_t a = 0, b = 0;
    for (int q = 0; q < klo; q++)
      if (i >> q & 1) { a += d.ksL[q]; b += d.ksR[q]; }
    kL[i] = a;
    kR[i] = b;
  }
  __syncthreads();
  const int vlb = __popc((unsigned)d.nibL), vrb = __popc((unsigned)d.nibR);
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  using Acc = typename AccOf<TO>::type;
  constexpr bool kBytePath = sizeof(TL) == 1 && sizeof(TR) == 1;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t baseL = deposit(d.hiL, (uint64_t)tile);
    const uint64_t baseR = deposit(d.hiR, (uint64_t)tile);
    for (int vi = threadIdx.x; vi < nt; vi += blockDim.x) {
      const uint64_t offL = baseL + tabL[vi];
      const uint64_t offR = baseR + tabR[vi];
      const int64_t o = (((int64_t)tile << d.tile_bits) | vi) << 4;
      if constexpr (kBytePath && CLAMP && sizeof(TO) == 1) {
        // boolean semiring on packed bytes: acc |= L & R, 4 lanes per LOP3
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int kh = 0; kh < (1 << khi); kh++) {
          int64_t hL = 0, hR = 0;
          for (int q = 0; q < khi; q++)
            if (kh >> q & 1) { hL += d.ksL[klo + q]; hR += d.ksR[klo + q]; }
          for (int kk = 0; kk < (1 << klo); kk++) {
            uint32_t buf[4], XL[4], XR[4];
            load_sub_u8((const uint8_t*)lhs + offL + hL + kL[kk], vlb, buf);
            expand16(buf, d.nibL, XL);
            if (d.has_rhs) {
              load_sub_u8((const uint8_t*)rhs + offR + hR + kR[kk], vrb, buf);
              expand16(buf, d.nibR, XR);
              a0 |= XL[0] & XR[0]; a1 |= XL[1] & XR[1]; a2 |= XL[2] & XR[2]; a3 |= XL[3] & XR[3];
            } else {
              a0 |= XL[0]; a1 |= XL[1]; a2 |= XL[2]; a3 |= XL[3];
            }
          }
        }
        uint4* dst = (uint4*)((uint8_t*)out + o);
        if (d.accumulate) { uint4 p = *dst; a0 |= p.x; a1 |= p.y; a2 |= p.z; a3 |= p.w; }
        *dst = make_uint4(a0, a1, a2, a3);
      } else {
        Acc acc[16];
#pragma unroll
        for (int e = 0; e < 16; e++) acc[e] = 0;
        int iL[16], iR[16];
#pragma unroll
        for (int e = 0; e < 16; e++) { iL[e] = pext4(e, d.nibL); iR[e] = pext4(e, d.nibR); }
        for (int kh = 0; kh < (1 << khi); kh++) {
          int64_t hL = 0, hR = 0;
          for (int q = 0; q < khi; q++)
            if (kh >> q & 1) { hL += d.ksL[klo + q]; hR += d.ksR[klo + q]; }
          for (int kk = 0; kk < (1 << klo); kk++) {
            const TL* pl = lhs + offL + hL + kL[kk];
            const TR* pr = rhs + offR + hR + kR[kk];
            if constexpr (kBytePath) {
              uint32_t buf[4], XL[4], XR[4];
              load_sub_u8((const uint8_t*)pl, vlb, buf);
              expand16(buf, d.nibL, XL);
              if (d.has_rhs) {
                load_sub_u8((const uint8_t*)pr, vrb, buf);
                expand16(buf, d.nibR, XR);
              } else {
                XR[0] = XR[1] = XR[2] = XR[3] = 0x01010101u;
              }
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const uint32_t a = (XL[e >> 2] >> (8 * (e & 3))) & 0xffu;
                const uint32_t b = (XR[e >> 2] >> (8 * (e & 3))) & 0xffu;
                acc[e] += (Acc)(a * b);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const Acc a = (Acc)pl[iL[e]];
                const Acc b = d.has_rhs ? (Acc)pr[iR[e]] : (Acc)1;
                acc[e] += a * b;
              }
            }
          }
        }
        TO* dst = out + o;
#pragma unroll
        for (int e = 0; e < 16; e++) {
          Acc v = acc[e];
          if (d.accumulate) v += (Acc)dst[e];
          if (CLAMP) v = v != 0;
          dst[e] = (TO)v;
        }
      }
    }
  }
}

}  // namespace tsb
