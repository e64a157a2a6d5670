// k_merge_bits_f.cuh — the fused n-ary bit-packed merge, factored.
//
// Same step as k_merge_bits_n (k_merge_bits.cuh): out[o] = OR over the summed wires of AND_i operand_i[...],
// i.e. a run of consecutive FinRel._contract2 steps (tensorsat/lang/fin_rel.py:255-288) whose intermediates are
// never stored.  On a sliced SAT plan the operands of these steps are small (L2 / L1 resident broadcast twins)
// and the step is bound by the number of 16-byte operand fetches per output vector, not by HBM.  Two things
// cut that number:
//   * which summed wires an operand holds is a template parameter (MASKS, KB bits per operand slot, slots
//     sorted by mask so that only the sorted tuples are instantiated): an operand that lacks a summed wire is
//     fetched once for both of its values instead of once per combination;
//   * one thread produces 2^NP output vectors that differ in NP "pair" wires chosen by the plan among the
//     output wires that most operands do NOT hold: such an operand is fetched once for all of them (which
//     pair wires an operand holds is uniform per launch and operand: a switch, no divergence).
// The last operand's AND is folded into the OR over the summed-wire combinations.
#pragma once
#include "k_merge.cuh"

namespace tsb {

constexpr int kBitsFThreads = 256;

TSB_HD constexpr int bits_f_mask(int masks, int kb, int o) { return (masks >> (o * kb)) & ((1 << kb) - 1); }

// one operand: fetch the distinct vectors (summed-wire subsets of M x pair subsets of H) and fold them in
template <int KB, int NP, int M, int H, bool FIRST, bool LAST>
TSB_D void bits_f_fold(const uint4* __restrict__ q, const BitsFOperand& op, uint4 (&prod)[1 << NP][1 << KB],
                       uint4 (&acc)[1 << NP]) {
  constexpr int KK = 1 << KB, GG = 1 << NP;
#pragma unroll
  for (int s = 0; s < KK; s++) {
    if ((s & ~M) != 0) continue;
    uint4 x[GG];
#pragma unroll
    for (int g = 0; g < GG; g++)
      if ((g & ~H) == 0) x[g] = q[op.kw[s] + op.pw[g]];
#pragma unroll
    for (int g = 0; g < GG; g++) {
      const uint4 v = x[g & H];
#pragma unroll
      for (int kk = 0; kk < KK; kk++) {
        if ((kk & M) != s) continue;
        if constexpr (LAST) {
          if constexpr (FIRST) {
            acc[g].x |= v.x; acc[g].y |= v.y; acc[g].z |= v.z; acc[g].w |= v.w;
          } else {
            acc[g].x |= prod[g][kk].x & v.x; acc[g].y |= prod[g][kk].y & v.y;
            acc[g].z |= prod[g][kk].z & v.z; acc[g].w |= prod[g][kk].w & v.w;
          }
        } else if constexpr (FIRST) {
          prod[g][kk] = v;
        } else {
          prod[g][kk].x &= v.x; prod[g][kk].y &= v.y; prod[g][kk].z &= v.z; prod[g][kk].w &= v.w;
        }
      }
    }
  }
}

template <int KB, int NP, int M, bool FIRST, bool LAST>
TSB_D void bits_f_operand(const uint4* __restrict__ q, const BitsFOperand& op, uint4 (&prod)[1 << NP][1 << KB],
                          uint4 (&acc)[1 << NP]) {
  if constexpr (NP == 0) {
    bits_f_fold<KB, NP, M, 0, FIRST, LAST>(q, op, prod, acc);
  } else if constexpr (NP == 1) {
    if (op.held) bits_f_fold<KB, NP, M, 1, FIRST, LAST>(q, op, prod, acc);
    else bits_f_fold<KB, NP, M, 0, FIRST, LAST>(q, op, prod, acc);
  } else {
    switch (op.held) {
      case 0: bits_f_fold<KB, NP, M, 0, FIRST, LAST>(q, op, prod, acc); break;
      case 1: bits_f_fold<KB, NP, M, 1, FIRST, LAST>(q, op, prod, acc); break;
      case 2: bits_f_fold<KB, NP, M, 2, FIRST, LAST>(q, op, prod, acc); break;
      default: bits_f_fold<KB, NP, M, 3, FIRST, LAST>(q, op, prod, acc); break;
    }
  }
}

template <int NOPS, int KB, int MASKS, int NP>
__global__ void __launch_bounds__(kBitsFThreads, NP == 2 ? 2 : 3)
k_merge_bits_f(const __grid_constant__ BitsFDesc d, const __grid_constant__ BitsNPtrs ptrs, uint4* __restrict__ out) {
  constexpr int KK = 1 << KB, GG = 1 << NP;
  __shared__ uint32_t tab[NOPS + 1][1 << kMergeTileBits];
  __shared__ RunsSmem runs[NOPS + 1];
  const int nt = 1 << d.tile_bits;
#pragma unroll
  for (int o = 0; o < NOPS; o++) {
    // offsets in 16-byte units: every operand holds the seven lane bits, so they are multiples of 128 entries
    for (int i = threadIdx.x; i < nt; i += kBitsFThreads) tab[o][i] = (uint32_t)(deposit(d.op[o].lo, (uint64_t)i) >> 7);
    runs_to_smem(d.op[o].hi, runs[o]);
  }
  for (int i = threadIdx.x; i < nt; i += kBitsFThreads) tab[NOPS][i] = (uint32_t)deposit(d.out_lo, (uint64_t)i);
  runs_to_smem(d.out_hi, runs[NOPS]);
  __syncthreads();
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  int64_t tile = blockIdx.x;
  uint64_t off[NOPS + 1];
#pragma unroll
  for (int o = 0; o <= NOPS; o++) off[o] = tile < n_tiles ? deposit_warp(runs[o], (uint64_t)tile) : 0;
  while (tile < n_tiles) {
    const uint4* base[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; o++) base[o] = (const uint4*)ptrs.p[o] + ((ptrs.bitoff[o] + off[o]) >> 7);
    uint4* outp = out + off[NOPS];
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
#pragma unroll
      for (int o = 0; o <= NOPS; o++) off[o] = deposit_warp(runs[o], (uint64_t)next);
    }
    for (int vi = threadIdx.x; vi < nt; vi += kBitsFThreads) {
      uint4 prod[GG][KK], acc[GG];
#pragma unroll
      for (int g = 0; g < GG; g++) acc[g] = make_uint4(0, 0, 0, 0);
      bits_f_operand<KB, NP, bits_f_mask(MASKS, KB, 0), true, NOPS == 1>(base[0] + tab[0][vi], d.op[0], prod, acc);
      if constexpr (NOPS > 1)
        bits_f_operand<KB, NP, bits_f_mask(MASKS, KB, 1), false, NOPS == 2>(base[1] + tab[1][vi], d.op[1], prod, acc);
      if constexpr (NOPS > 2)
        bits_f_operand<KB, NP, bits_f_mask(MASKS, KB, 2), false, NOPS == 3>(base[2] + tab[2][vi], d.op[2], prod, acc);
      if constexpr (NOPS > 3)
        bits_f_operand<KB, NP, bits_f_mask(MASKS, KB, 3), false, NOPS == 4>(base[3] + tab[3][vi], d.op[3], prod, acc);
      uint4* o = outp + tab[NOPS][vi];
#pragma unroll
      for (int g = 0; g < GG; g++) o[d.ow[g]] = acc[g];
    }
    tile = next;
  }
}

}  // namespace tsb
