// k_merge_bits.cuh — the boolean-semiring monotone merge on BIT-PACKED relations (1 bit per entry).
//
// SURVEY.md §8(f) N4 / the reference authors' planned "Binary Packing" (paper/typst/main.typ:336-339):
// the steps of FinRel._contract2 (tensorsat/lang/fin_rel.py:255-288) that dominate a SAT plan are
// bandwidth-bound, and a 0/1 relation stored one bit per entry moves 8x fewer bytes.  Entry i of a
// packed tensor is bit (i & 31) of 32-bit word (i >> 5); tensors smaller than a word are padded.
//
// Same scheme as k_merge_sat (one global wire order => operand address bits are an order-preserving
// subsequence of the output's), one thread per 16-byte output vector = 128 entries = the 7 lowest
// output address bits.  Per operand and summed-wire combination the thread needs the aligned chunk of
// 2^c bits (c = number of those 7 bits the operand depends on) that the vector touches: ONE load of
// 4 / 8 / 16 bytes, then an expansion to 128 lanes.  Word-level lanes (bits 5, 6) are register moves;
// in-word lanes (bits 0..4) go through a 256-entry shared-memory table per operand built once per CTA
// (an operand with 4 of the 5 in-word bits is expanded by bit duplication instead), so any of the 128 x 128 lane-mask pairs
// runs on the same code.  Combine: 4 LOP3 per summed-wire combination; one 16-byte store.
// Bound: HBM at 1/8 byte per entry (algorithmic bytes = (|lhs| + |rhs| + |out|) / 8).
#pragma once
#include "k_merge.cuh"

namespace tsb {

constexpr int kBitsThreads = 256;
constexpr int kBitsG = 2;        // output vectors per thread in flight
constexpr int kBitsVecBits = 7;  // entries per 16-byte vector, log2

TSB_HD int pext_bits(int e, int mask, int nbits) {
  int r = 0, o = 0;
  for (int b = 0; b < nbits; b++)
    if (mask >> b & 1) { r |= ((e >> b) & 1) << o; o++; }
  return r;
}

struct BitsLane {  // uniform per launch and operand
  int c;    // lanes bits the operand depends on (0..7)
  int c5;   // of which in-word (bits 0..4)
  int hi2;  // mask of word-level bits (output bits 5, 6)
  int p4;   // c5 == 4: the one in-word lane bit the operand lacks
};

// 16 bits -> the 32 in-word lanes of an operand that lacks in-word lane bit p: lane e reads bit pext(e, ~(1 << p)),
// i.e. blocks of 2^p bits are doubled.  Pure ALU (a byte permute for p = 3, shifts and masks below that): the
// two-table lookup this replaces made a 2^30-entry step of the n=200 plan run at a quarter of the HBM rate.
TSB_D uint32_t bits_dup16(uint32_t s, int p) {
  if (p == 4) return s | (s << 16);
  if (p == 3) return __byte_perm(s, 0, 0x1100);
  uint32_t t = (s | (s << 8)) & 0x00FF00FFu;   // bytes into 16-bit slots
  t = (t | (t << 4)) & 0x0F0F0F0Fu;            // nibbles into byte slots
  if (p == 2) return t * 0x11u;
  t = (t | (t << 2)) & 0x33333333u;            // bit pairs into nibble slots
  if (p == 1) return t * 5u;
  t = (t | (t << 1)) & 0x55555555u;            // bits into pair slots
  return t * 3u;
}

TSB_HD BitsLane bits_lane(int nib7) {
  BitsLane l;
  int c = 0, c5 = 0;
  for (int b = 0; b < 7; b++)
    if (nib7 >> b & 1) { c++; if (b < 5) c5++; }
  l.c = c; l.c5 = c5; l.hi2 = nib7 >> 5;
  l.p4 = 0;
  for (int b = 0; b < 5; b++) if (!(nib7 >> b & 1)) l.p4 = b;
  return l;
}

// lut[b]: the 32 in-word lanes produced by a sub-chunk (at most 8 bits: at most three in-word lanes held) of value b.
// Built once per CTA and only as far as the kernel will index it (nothing when the operand holds all five
// in-word lanes, table 0 only when it holds at most three).  `pos` is 32 words of shared scratch.
TSB_D void bits_build_lut(int nib7, uint32_t* lut, uint32_t* pos) {
  const int lo5 = nib7 & 31;
  const int c5 = __popc((unsigned)lo5);
  if (threadIdx.x < 32) pos[threadIdx.x] = (uint32_t)pext_bits(threadIdx.x, lo5, 5);
  __syncthreads();
  if (c5 >= 4) return;   // all five lanes: used in place; four: bits_dup16
  const int n = 1 << (1 << c5);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t v = (uint32_t)i;
    uint32_t x = 0;
#pragma unroll 8
    for (int e = 0; e < 32; e++) x |= ((v >> pos[e]) & 1u) << e;
    lut[i] = x;
  }
}

// Lane classes known at compile time (k_merge_bits<KB, LCL, LCR>): 0 = anything (decided per launch at run time),
// 1 = the operand holds all seven lane bits (one 16-byte load, used in place), 2 = it lacks exactly one in-word lane
// bit and holds both word-level bits (one 8-byte load, four bit duplications).  The run-time form costs ~50
// instructions per fetch; a 2^30-entry step of the n=200 plan spent 226 instructions per result vector in it.
TSB_HD int bits_lane_class(int nib7) {
  if (nib7 == 127) return 1;
  int c5 = 0;
  for (int b = 0; b < 5; b++) c5 += nib7 >> b & 1;
  return (nib7 >> 5) == 3 && c5 == 4 ? 2 : 0;
}

// X[w] = the 32 lanes of output word w: lane e of word w reads chunk bit pext(w*32+e, nib7)
template <int LC = 0>
TSB_D void bits_fetch(const uint32_t* __restrict__ base, uint64_t a, const BitsLane ln, const uint32_t* lut,
                      uint32_t X[4]) {
  const uint32_t* p = base + (a >> 5);
  if constexpr (LC == 1) {
    const uint4 t = *(const uint4*)p;
    X[0] = t.x; X[1] = t.y; X[2] = t.z; X[3] = t.w;
    return;
  }
  if constexpr (LC == 2) {
    const uint2 t = *(const uint2*)p;
    X[0] = bits_dup16(t.x & 0xffffu, ln.p4); X[1] = bits_dup16(t.x >> 16, ln.p4);
    X[2] = bits_dup16(t.y & 0xffffu, ln.p4); X[3] = bits_dup16(t.y >> 16, ln.p4);
    return;
  }
  uint32_t e0, e1 = 0, e2 = 0, e3 = 0;
  if (ln.c5 == 5) {
    if (ln.hi2 == 0) { e0 = *p; }
    else if (ln.hi2 == 3) { const uint4 t = *(const uint4*)p; e0 = t.x; e1 = t.y; e2 = t.z; e3 = t.w; }
    else { const uint2 t = *(const uint2*)p; e0 = t.x; e1 = t.y; }
  } else {
    // sub-word sub-chunks: the chunk (<= 64 bits) starts at bit (a & 31) of *p and never straddles
    uint64_t v;
    if (ln.c == 6) { const uint2 t = *(const uint2*)p; v = (uint64_t)t.x | ((uint64_t)t.y << 32); }
    else v = (uint64_t)(*p >> (uint32_t)(a & 31));
    const uint32_t mask = (1u << (1 << ln.c5)) - 1u;
    const int nsub = ln.hi2 == 0 ? 1 : ln.hi2 == 3 ? 4 : 2;
    uint32_t s0 = (uint32_t)v & mask, s1 = (uint32_t)(v >> (1 << ln.c5)) & mask,
             s2 = (uint32_t)(v >> (2 << ln.c5)) & mask, s3 = (uint32_t)(v >> (3 << ln.c5)) & mask;
    if (ln.c5 == 4) {
      e0 = bits_dup16(s0, ln.p4);
      if (nsub > 1) e1 = bits_dup16(s1, ln.p4);
      if (nsub > 2) { e2 = bits_dup16(s2, ln.p4); e3 = bits_dup16(s3, ln.p4); }
    } else {
      e0 = lut[s0];
      if (nsub > 1) e1 = lut[s1];
      if (nsub > 2) { e2 = lut[s2]; e3 = lut[s3]; }
    }
  }
  switch (ln.hi2) {
    case 0: X[0] = X[1] = X[2] = X[3] = e0; break;
    case 1: X[0] = e0; X[1] = e1; X[2] = e0; X[3] = e1; break;
    case 2: X[0] = X[1] = e0; X[2] = X[3] = e1; break;
    default: X[0] = e0; X[1] = e1; X[2] = e2; X[3] = e3; break;
  }
}

// d.nibL / d.nibR hold 7-bit lane masks, d.n_vec = 2^(R-7), all operand offsets are in entries (= bits).
// bitoffL / bitoffR: where the operand (a slice of a shared tensor) starts inside lhs / rhs.
// KB = 0 / 1: that many summed wires, every load of a thread's vectors issued before the first use
// (the steps of a SAT plan sum at most one wire); KB = -1: loop over the 2^kbits combinations.
template <int KB, int LCL = 0, int LCR = 0>
__global__ void __launch_bounds__(kBitsThreads, KB == 1 ? 3 : 4)
k_merge_bits(const __grid_constant__ MergeDesc d, const uint32_t* __restrict__ lhs, uint64_t bitoffL,
             const uint32_t* __restrict__ rhs, uint64_t bitoffR, uint4* __restrict__ out) {
  __shared__ uint32_t tabL[1 << kMergeTileBits], tabR[1 << kMergeTileBits];
  __shared__ int64_t kL[256], kR[256];
  __shared__ uint32_t lutL[256], lutR[256], posL[32], posR[32];
  __shared__ RunsSmem runsL, runsR;
  TSB_PDL_TRIGGER();
  const int nt = 1 << d.tile_bits;
  for (int i = threadIdx.x; i < nt; i += kBitsThreads) {
    tabL[i] = (uint32_t)deposit(d.loL, (uint64_t)i);
    tabR[i] = (uint32_t)deposit(d.loR, (uint64_t)i);
  }
  for (int i = threadIdx.x; i < (1 << d.kbits); i += kBitsThreads) {
    int64_t a = 0, b = 0;
    for (int q = 0; q < d.kbits; q++)
      if (i >> q & 1) { a += d.ksL[q]; b += d.ksR[q]; }
    kL[i] = a;
    kR[i] = b;
  }
  bits_build_lut(d.nibL, lutL, posL);
  if (d.has_rhs) bits_build_lut(d.nibR, lutR, posR);
  runs_to_smem(d.hiL, runsL);
  runs_to_smem(d.hiR, runsR);
  __syncthreads();
  TSB_PDL_WAIT();
  const BitsLane lnL = bits_lane(d.nibL), lnR = bits_lane(d.nibR);
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const bool acc_out = d.accumulate != 0, has_rhs = d.has_rhs != 0;
  const uint64_t sl1 = (uint64_t)d.ksL[0], sr1 = (uint64_t)d.ksR[0];  // KB == 1: stride of the summed wire
  int64_t tile = blockIdx.x;
  uint64_t offL = tile < n_tiles ? deposit_warp(runsL, (uint64_t)tile) : 0;
  uint64_t offR = tile < n_tiles ? deposit_warp(runsR, (uint64_t)tile) : 0;
  while (tile < n_tiles) {
    const uint64_t baseL = bitoffL + offL, baseR = bitoffR + offR;
    uint4* outp = out + (tile << d.tile_bits);
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      offL = deposit_warp(runsL, (uint64_t)next);
      offR = deposit_warp(runsR, (uint64_t)next);
    }
    for (int v0 = threadIdx.x; v0 < nt; v0 += kBitsThreads * kBitsG) {
      uint64_t aL[kBitsG], aR[kBitsG];
      bool on[kBitsG];
      uint32_t acc[kBitsG][4];
#pragma unroll
      for (int g = 0; g < kBitsG; g++) {
        const int vi = v0 + g * kBitsThreads;
        on[g] = vi < nt;
        aL[g] = baseL + tabL[on[g] ? vi : 0];
        aR[g] = baseR + tabR[on[g] ? vi : 0];
        acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0;
      }
      if (acc_out) {
#pragma unroll
        for (int g = 0; g < kBitsG; g++)
          if (on[g]) {
            const uint4 p = outp[v0 + g * kBitsThreads];
            acc[g][0] = p.x; acc[g][1] = p.y; acc[g][2] = p.z; acc[g][3] = p.w;
          }
      }
      if constexpr (KB >= 0) {
        constexpr int KK = 1 << KB;
        uint32_t XL[kBitsG][KK][4], XR[kBitsG][KK][4];
#pragma unroll
        for (int g = 0; g < kBitsG; g++) {
#pragma unroll
          for (int kk = 0; kk < KK; kk++) {
            if (on[g]) {
              bits_fetch<LCL>(lhs, aL[g] + (kk ? sl1 : 0), lnL, lutL, XL[g][kk]);
              if (has_rhs) bits_fetch<LCR>(rhs, aR[g] + (kk ? sr1 : 0), lnR, lutR, XR[g][kk]);
              else XR[g][kk][0] = XR[g][kk][1] = XR[g][kk][2] = XR[g][kk][3] = 0xffffffffu;
            } else {
              XL[g][kk][0] = XL[g][kk][1] = XL[g][kk][2] = XL[g][kk][3] = 0;
              XR[g][kk][0] = XR[g][kk][1] = XR[g][kk][2] = XR[g][kk][3] = 0;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < kBitsG; g++) {
#pragma unroll
          for (int kk = 0; kk < KK; kk++) {
            acc[g][0] |= XL[g][kk][0] & XR[g][kk][0]; acc[g][1] |= XL[g][kk][1] & XR[g][kk][1];
            acc[g][2] |= XL[g][kk][2] & XR[g][kk][2]; acc[g][3] |= XL[g][kk][3] & XR[g][kk][3];
          }
        }
      } else {
        const int K = 1 << d.kbits;
        for (int kk = 0; kk < K; kk++) {
          const uint64_t sl = (uint64_t)kL[kk], sr = (uint64_t)kR[kk];
          uint32_t XL[kBitsG][4], XR[kBitsG][4];
#pragma unroll
          for (int g = 0; g < kBitsG; g++) {
            if (on[g]) {
              bits_fetch<LCL>(lhs, aL[g] + sl, lnL, lutL, XL[g]);
              if (has_rhs) bits_fetch<LCR>(rhs, aR[g] + sr, lnR, lutR, XR[g]);
              else XR[g][0] = XR[g][1] = XR[g][2] = XR[g][3] = 0xffffffffu;
            } else {
              XL[g][0] = XL[g][1] = XL[g][2] = XL[g][3] = 0;
              XR[g][0] = XR[g][1] = XR[g][2] = XR[g][3] = 0;
            }
          }
#pragma unroll
          for (int g = 0; g < kBitsG; g++) {
            acc[g][0] |= XL[g][0] & XR[g][0]; acc[g][1] |= XL[g][1] & XR[g][1];
            acc[g][2] |= XL[g][2] & XR[g][2]; acc[g][3] |= XL[g][3] & XR[g][3];
          }
        }
      }
#pragma unroll
      for (int g = 0; g < kBitsG; g++)
        if (on[g]) outp[v0 + g * kBitsThreads] = make_uint4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
    }
    tile = next;
  }
}

// ---- word-aligned fast path ---------------------------------------------------------------------------
// Both operands hold all five in-word lanes (true for every large step of a plan once small operands are
// stored broadcast, see tsb_plan.cu): every offset is a multiple of 32 entries, so the kernel works in
// 32-bit WORD offsets, the word-level lane masks HL / HR (output bits 5, 6) are template parameters and an
// operand fetch is exactly one 4 / 8 / 16-byte load whose words are used in place.
// H = 0..3: which of the two word-level lane bits the operand holds (all five in-word lanes held); H = 4: the operand
// holds both word-level bits and lacks ONE in-word lane bit p4 (8-byte load, each 16-bit quarter doubled block-wise)
template <int H>
TSB_D void bits_load_words(const uint32_t* __restrict__ p, uint32_t X[4], int p4 = 0) {
  if constexpr (H == 4) {
    const uint2 t = *(const uint2*)p;
    X[0] = bits_dup16(t.x & 0xffffu, p4); X[1] = bits_dup16(t.x >> 16, p4);
    X[2] = bits_dup16(t.y & 0xffffu, p4); X[3] = bits_dup16(t.y >> 16, p4);
  }
  else if constexpr (H == 0) { const uint32_t t = *p; X[0] = X[1] = X[2] = X[3] = t; }
  else if constexpr (H == 1) { const uint2 t = *(const uint2*)p; X[0] = t.x; X[1] = t.y; X[2] = t.x; X[3] = t.y; }
  else if constexpr (H == 2) { const uint2 t = *(const uint2*)p; X[0] = X[1] = t.x; X[2] = X[3] = t.y; }
  else { const uint4 t = *(const uint4*)p; X[0] = t.x; X[1] = t.y; X[2] = t.z; X[3] = t.w; }
}

template <int KB, int HL, int HR>
__global__ void __launch_bounds__(kBitsThreads, 4)
k_merge_bits_w(const __grid_constant__ MergeDesc d, const uint32_t* __restrict__ lhs, uint64_t bitoffL,
               const uint32_t* __restrict__ rhs, uint64_t bitoffR, uint4* __restrict__ out) {
  constexpr int KK = 1 << KB;
  constexpr int G = KB == 0 ? 4 : 2;  // eight operand fetches in flight per thread either way
  __shared__ uint32_t tabL[1 << kMergeTileBits], tabR[1 << kMergeTileBits];
  __shared__ RunsSmem runsL, runsR;
  TSB_PDL_TRIGGER();
  const int nt = 1 << d.tile_bits;
  for (int i = threadIdx.x; i < nt; i += kBitsThreads) {
    tabL[i] = (uint32_t)(deposit(d.loL, (uint64_t)i) >> 5);
    tabR[i] = (uint32_t)(deposit(d.loR, (uint64_t)i) >> 5);
  }
  runs_to_smem(d.hiL, runsL);
  runs_to_smem(d.hiR, runsR);
  __syncthreads();
  TSB_PDL_WAIT();
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const bool acc_out = d.accumulate != 0;
  const uint64_t kwL = KB ? (uint64_t)d.ksL[0] >> 5 : 0, kwR = KB ? (uint64_t)d.ksR[0] >> 5 : 0;
  const int p4L = HL == 4 ? bits_lane(d.nibL).p4 : 0, p4R = HR == 4 ? bits_lane(d.nibR).p4 : 0;
  int64_t tile = blockIdx.x;
  uint64_t offL = tile < n_tiles ? deposit_warp(runsL, (uint64_t)tile) : 0;
  uint64_t offR = tile < n_tiles ? deposit_warp(runsR, (uint64_t)tile) : 0;
  while (tile < n_tiles) {
    const uint32_t* pL = lhs + ((bitoffL + offL) >> 5);
    const uint32_t* pR = rhs + ((bitoffR + offR) >> 5);
    uint4* outp = out + (tile << d.tile_bits);
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      offL = deposit_warp(runsL, (uint64_t)next);
      offR = deposit_warp(runsR, (uint64_t)next);
    }
    for (int v0 = threadIdx.x; v0 < nt; v0 += kBitsThreads * G) {
      uint32_t XL[G][KK][4], XR[G][KK][4];
      bool on[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int vi = v0 + g * kBitsThreads;
        on[g] = vi < nt;
        const uint32_t* ql = pL + tabL[on[g] ? vi : 0];
        const uint32_t* qr = pR + tabR[on[g] ? vi : 0];
#pragma unroll
        for (int kk = 0; kk < KK; kk++) {
          bits_load_words<HL>(ql + (kk ? kwL : 0), XL[g][kk], p4L);
          bits_load_words<HR>(qr + (kk ? kwR : 0), XR[g][kk], p4R);
        }
      }
#pragma unroll
      for (int g = 0; g < G; g++) {
        if (!on[g]) continue;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (acc_out) r = outp[v0 + g * kBitsThreads];
#pragma unroll
        for (int kk = 0; kk < KK; kk++) {
          r.x |= XL[g][kk][0] & XR[g][kk][0]; r.y |= XL[g][kk][1] & XR[g][kk][1];
          r.z |= XL[g][kk][2] & XR[g][kk][2]; r.w |= XL[g][kk][3] & XR[g][kk][3];
        }
        outp[v0 + g * kBitsThreads] = r;
      }
    }
    tile = next;
  }
}

// ---- fused n-ary merge ("bucket" step) -----------------------------------------------------------------
// out[o] = OR over the summed wires of AND_i operand_i[...]: several consecutive steps of the tree in ONE
// pass, so the intermediates between them never reach HBM (classic bucket elimination: multiply every
// tensor of the bucket, sum the variable out, without materialising the product).  For a chain
// T1 = sum_k1 X*Y, T2 = sum_k2 T1*Z the fused step reads X, Y, Z and writes T2; T1 (the widest tensor of
// the plan) is neither written nor read back.  Word-aligned operands only (see k_merge_bits_w).
// Every operand of a fused step holds all seven lane bits of the output (small operands are stored broadcast
// over the ones they do not depend on, see tsb_plan.cu), so an operand fetch is one 16-byte load.
template <int NOPS, int KB>
__global__ void __launch_bounds__(kBitsThreads, 3)
k_merge_bits_n(const __grid_constant__ BitsNDesc d, const __grid_constant__ BitsNPtrs ptrs, uint4* __restrict__ out) {
  constexpr int KK = 1 << KB;
  constexpr int G = 2;   // operand fetches in flight per thread: G * NOPS (times KU)
  [[maybe_unused]] constexpr int KU = NOPS * KK <= 8 ? KK : (NOPS <= 2 ? 2 : 1);  // unrolling further only spills
  __shared__ uint32_t tab[NOPS][1 << kMergeTileBits];
  __shared__ uint32_t kw[NOPS][KK];
  __shared__ RunsSmem runs[NOPS];
  TSB_PDL_TRIGGER();
  const int nt = 1 << d.tile_bits;
#pragma unroll
  for (int o = 0; o < NOPS; o++) {
    // offsets in 16-byte units: every operand offset is a multiple of 128 entries
    for (int i = threadIdx.x; i < nt; i += kBitsThreads) tab[o][i] = (uint32_t)(deposit(d.op[o].lo, (uint64_t)i) >> 7);
    if (threadIdx.x < KK) {
      int64_t a = 0;
      for (int q = 0; q < KB; q++)
        if (threadIdx.x >> q & 1) a += d.op[o].ks[q];
      kw[o][threadIdx.x] = (uint32_t)(a >> 7);
    }
    runs_to_smem(d.op[o].hi, runs[o]);
  }
  __syncthreads();
  TSB_PDL_WAIT();
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  int64_t tile = blockIdx.x;
  uint64_t off[NOPS];
#pragma unroll
  for (int o = 0; o < NOPS; o++) off[o] = tile < n_tiles ? deposit_warp(runs[o], (uint64_t)tile) : 0;
  while (tile < n_tiles) {
    const uint4* base[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; o++) base[o] = (const uint4*)ptrs.p[o] + ((ptrs.bitoff[o] + off[o]) >> 7);
    uint4* outp = out + (tile << d.tile_bits);
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
#pragma unroll
      for (int o = 0; o < NOPS; o++) off[o] = deposit_warp(runs[o], (uint64_t)next);
    }
    for (int v0 = threadIdx.x; v0 < nt; v0 += kBitsThreads * G) {
      uint32_t q[G][NOPS];
      bool on[G];
      uint4 acc[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int vi = v0 + g * kBitsThreads;
        on[g] = vi < nt;
#pragma unroll
        for (int o = 0; o < NOPS; o++) q[g][o] = tab[o][on[g] ? vi : 0];
        acc[g] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll KU
      for (int kk = 0; kk < KK; kk++) {
        uint4 X[G][NOPS];
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int o = 0; o < NOPS; o++) X[g][o] = base[o][q[g][o] + kw[o][kk]];
#pragma unroll
        for (int g = 0; g < G; g++) {
          uint4 t = X[g][0];
#pragma unroll
          for (int o = 1; o < NOPS; o++) { t.x &= X[g][o].x; t.y &= X[g][o].y; t.z &= X[g][o].z; t.w &= X[g][o].w; }
          acc[g].x |= t.x; acc[g].y |= t.y; acc[g].z |= t.z; acc[g].w |= t.w;
        }
      }
#pragma unroll
      for (int g = 0; g < G; g++)
        if (on[g]) outp[v0 + g * kBitsThreads] = acc[g];
    }
    tile = next;
  }
}

// ---- conversions between the byte (0/1 per uint8) and the bit-packed representation -----------------

// four 0/1 bytes -> 4 bits (byte j -> bit j): the partial products land on distinct bits, no carries
TSB_D uint32_t pack4(uint32_t w) { return ((w * 0x01020408u) >> 24) & 0xfu; }
// 4 bits -> four 0/1 bytes
TSB_D uint32_t unpack4(uint32_t n) { return ((n & 0xfu) * 0x00204081u) & 0x01010101u; }

static __global__ void k_pack_bits(const uint8_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n_entries) {
  TSB_PDL_TRIGGER();
  TSB_PDL_WAIT();
  const int64_t n_words = (n_entries + 31) >> 5;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = w << 5;
    uint32_t bits = 0;
    if (e0 + 32 <= n_entries) {
      const uint4 a = *(const uint4*)(src + e0), b = *(const uint4*)(src + e0 + 16);
      bits = pack4(a.x) | pack4(a.y) << 4 | pack4(a.z) << 8 | pack4(a.w) << 12 | pack4(b.x) << 16 | pack4(b.y) << 20 |
             pack4(b.z) << 24 | pack4(b.w) << 28;
    } else {
      for (int64_t j = e0; j < n_entries; j++) bits |= (uint32_t)(src[j] != 0) << (j - e0);
    }
    dst[w] = bits;
  }
}

// Packing with an index map: broadcast over wires the source does not have, or re-layout of a packed source.
template <bool SRC_PACKED>
__global__ void k_pack_map(const void* __restrict__ src, uint32_t* __restrict__ dst, int64_t n_entries,
                           const __grid_constant__ BitRuns lo, const __grid_constant__ BitRuns hi) {
  __shared__ uint32_t lane_off[32];
  TSB_PDL_TRIGGER();
  if (threadIdx.x < 32) lane_off[threadIdx.x] = (uint32_t)deposit(lo, (uint64_t)threadIdx.x);
  __syncthreads();
  TSB_PDL_WAIT();
  const int64_t n_words = (n_entries + 31) >> 5;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t base = deposit(hi, (uint64_t)w);
    const int nl = (int)(n_entries - (w << 5) < 32 ? n_entries - (w << 5) : 32);
    uint32_t bits = 0;
    for (int e = 0; e < nl; e++) {
      const uint64_t idx = base + lane_off[e];
      uint32_t b;
      if constexpr (SRC_PACKED) b = (((const uint32_t*)src)[idx >> 5] >> (uint32_t)(idx & 31)) & 1u;
      else b = ((const uint8_t*)src)[idx] != 0;
      bits |= b << e;
    }
    dst[w] = bits;
  }
}

static __global__ void k_unpack_bits(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n_entries,
                                     int accumulate) {
  TSB_PDL_TRIGGER();
  TSB_PDL_WAIT();
  const int64_t n_vec = (n_entries + 15) >> 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t h = (src[v >> 1] >> ((v & 1) * 16)) & 0xffffu;
    const int64_t e0 = v << 4;
    if (e0 + 16 <= n_entries) {
      uint4 r = make_uint4(unpack4(h), unpack4(h >> 4), unpack4(h >> 8), unpack4(h >> 12));
      uint4* q = (uint4*)(dst + e0);
      if (accumulate) { const uint4 p = *q; r.x |= p.x; r.y |= p.y; r.z |= p.z; r.w |= p.w; }
      *q = r;
    } else {
      for (int64_t j = e0; j < n_entries; j++) {
        const uint8_t b = (uint8_t)((h >> (j - e0)) & 1u);
        dst[j] = accumulate ? (uint8_t)(dst[j] | b) : b;
      }
    }
  }
}

}  // namespace tsb
