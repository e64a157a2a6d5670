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
//   * optionally (NP = 1) one thread produces two output vectors that differ in a "pair" wire chosen by the
//     plan among the output wires that most operand fetches do NOT depend on: such an operand is fetched once
//     for both.
#pragma once
#include "k_merge.cuh"

namespace tsb {

constexpr int kBitsFThreads = 256;

TSB_HD constexpr int bits_f_mask(int masks, int kb, int o) { return (masks >> (o * kb)) & ((1 << kb) - 1); }

// acc = OR over the summed-wire combinations of AND over the operands; x[o][s] holds operand o at the
// combination s of the summed wires IT holds (s subset of its mask), so combination kk reads x[o][kk & mask_o]
template <int NOPS, int KB, int MASKS>
TSB_D uint4 bits_f_combine(const uint4 (&x)[NOPS][1 << KB]) {
  constexpr int KK = 1 << KB;
  uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int kk = 0; kk < KK; kk++) {
    uint4 t = x[0][kk & bits_f_mask(MASKS, KB, 0)];
#pragma unroll
    for (int o = 1; o < NOPS; o++) {
      const uint4 v = x[o][kk & bits_f_mask(MASKS, KB, o)];
      t.x &= v.x; t.y &= v.y; t.z &= v.z; t.w &= v.w;
    }
    acc.x |= t.x; acc.y |= t.y; acc.z |= t.z; acc.w |= t.w;
  }
  return acc;
}

// NP = 0: one output vector per thread and iteration.  NP = 1: two vectors that differ in the plan-chosen pair
// wire; an operand that does not hold it (uniform per launch) is fetched once for both: its second fetch is
// predicated off and the first one's registers are copied instead.  Straight-line code either way, so every
// fetch of an iteration is in flight before the first use.
//
// NP = 2: four vectors per work item, over two pair wires, visited in Gray-code order g = 0, 1, 3, 2 so that one pair
// wire flips per step; before each step only the operands that HOLD the flipped wire are fetched again, the others
// stay in registers.  Nothing but the current operand values is live (no second copy), so the kernel keeps the
// register budget of NP = 0 while an operand that holds neither pair wire is fetched once for four vectors, one that
// holds only the slow wire (bit 1) twice, only the fast wire (bit 0) three times.
template <int NOPS, int KB, int MASKS, int NP>
__global__ void __launch_bounds__(kBitsFThreads, NP == 1 ? 2 : 3)
k_merge_bits_f(const __grid_constant__ BitsFDesc d, const __grid_constant__ BitsNPtrs ptrs, uint4* __restrict__ out,
               const uint32_t* __restrict__ gtab) {
  constexpr int KK = 1 << KB;
  __shared__ uint32_t tab[NOPS + 1][1 << kMergeTileBits];
  __shared__ RunsSmem runs[NOPS + 1];
  TSB_PDL_TRIGGER();
  const int nt = 1 << d.tile_bits;
  if (gtab != nullptr) {
    // the offset tables depend on the descriptor only: a plan computes them once (bits_f_tables) and every CTA of
    // every launch copies them in, [slot][vector in tile], instead of depositing bit runs again
#pragma unroll
    for (int o = 0; o <= NOPS; o++)
      for (int i = threadIdx.x; i < nt; i += kBitsFThreads) tab[o][i] = gtab[o * nt + i];
#pragma unroll
    for (int o = 0; o < NOPS; o++) runs_to_smem(d.op[o].hi, runs[o]);
  } else {
#pragma unroll
    for (int o = 0; o < NOPS; o++) {
      // offsets in 16-byte units: every operand holds the seven lane bits, so they are multiples of 128 entries
      for (int i = threadIdx.x; i < nt; i += kBitsFThreads) tab[o][i] = (uint32_t)(deposit(d.op[o].lo, (uint64_t)i) >> 7);
      runs_to_smem(d.op[o].hi, runs[o]);
    }
    for (int i = threadIdx.x; i < nt; i += kBitsFThreads) tab[NOPS][i] = (uint32_t)deposit(d.out_lo, (uint64_t)i);
  }
  runs_to_smem(d.out_hi, runs[NOPS]);
  __syncthreads();
  TSB_PDL_WAIT();  // tables came from the parameters; operands and result are touched from here on
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
      uint4 x0[NOPS][KK];
#pragma unroll
      for (int o = 0; o < NOPS; o++) {
        const uint4* q = base[o] + tab[o][vi];
#pragma unroll
        for (int s = 0; s < KK; s++)
          if ((s & ~bits_f_mask(MASKS, KB, o)) == 0) x0[o][s] = q[d.op[o].kw[s]];
      }
      uint4* po = outp + tab[NOPS][vi];
      if constexpr (NP == 2) {
        po[0] = bits_f_combine<NOPS, KB, MASKS>(x0);
#pragma unroll
        for (int step = 1; step < 4; step++) {
          const int g = step == 1 ? 1 : step == 2 ? 3 : 2;   // Gray code: bit 0, bit 1, bit 0 flip
          const int flipped = step == 2 ? 2 : 1;
#pragma unroll
          for (int o = 0; o < NOPS; o++) {
            if (d.op[o].held & flipped) {
              const uint4* q = base[o] + tab[o][vi] + d.op[o].pw[g];
#pragma unroll
              for (int s = 0; s < KK; s++)
                if ((s & ~bits_f_mask(MASKS, KB, o)) == 0) x0[o][s] = q[d.op[o].kw[s]];
            }
          }
          po[d.ow[g]] = bits_f_combine<NOPS, KB, MASKS>(x0);
        }
      } else if constexpr (NP == 1) {
        // the second vector's real fetches are issued before anything waits on the first vector's data
        uint4 x1[NOPS][KK];
#pragma unroll
        for (int o = 0; o < NOPS; o++) {
          const uint4* q = base[o] + tab[o][vi] + d.op[o].pw[1];
          if (d.op[o].held != 0) {
#pragma unroll
            for (int s = 0; s < KK; s++)
              if ((s & ~bits_f_mask(MASKS, KB, o)) == 0) x1[o][s] = q[d.op[o].kw[s]];
          }
        }
        po[0] = bits_f_combine<NOPS, KB, MASKS>(x0);
#pragma unroll
        for (int o = 0; o < NOPS; o++) {
          if (d.op[o].held == 0) {
#pragma unroll
            for (int s = 0; s < KK; s++)
              if ((s & ~bits_f_mask(MASKS, KB, o)) == 0) x1[o][s] = x0[o][s];
          }
        }
        po[d.ow[1]] = bits_f_combine<NOPS, KB, MASKS>(x1);
      } else {
        po[0] = bits_f_combine<NOPS, KB, MASKS>(x0);
      }
    }
    tile = next;
  }
}

}  // namespace tsb
