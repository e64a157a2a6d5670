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
//     for both;
//   * NP = 2 / 3 ("private streaming"): the pair wires are output wires that ONLY the operand in slot 0 holds.  A
//     thread fetches the other operands once, folds them into one vector per summed-wire combination of slot 0, and
//     then streams slot 0 over the 4 / 8 pair-wire combinations: 2^|mask_0| fetches and as many LOP3 per word for
//     every further output vector — the floor of the step, since every vector of slot 0 is used exactly once;
//   * NT = 1 / 2 on top of that ("two-sided"): slot 1 has private output wires of its own.  Its vectors for the 2 / 4
//     combinations of those are folded in as well (one vector per summed-wire combination AND combination of slot 1's
//     wires stays in registers), and every vector streamed from slot 0 then yields 2 / 4 output vectors.
#pragma once
#include "k_merge.cuh"

namespace tsb {

constexpr int kBitsFThreads = 256;

TSB_HD constexpr int bits_f_mask(int masks, int kb, int o) { return (masks >> (o * kb)) & ((1 << kb) - 1); }

// L2 eviction-priority hints (sm_80+): the result of a per-slice step that the next launch streams back in should
// stay in the 126 MB L2 instead of being written through to HBM and read back
TSB_D uint64_t bits_f_policy(bool last) {
#ifndef TSB_HOST_EMU
  uint64_t pol;
  if (last) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#else
  return 0;
#endif
}
TSB_D void bits_f_store_hint(uint4* p, const uint4& v, uint64_t pol) {
#ifndef TSB_HOST_EMU
  // (no "memory" clobber: nothing in these kernels reads the result back, and the read-only fetches of the next
  // combinations must stay free to move above the store)
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol));
#else
  *p = v;
#endif
}
TSB_D uint4 bits_f_ldg_hint(const uint4* p, uint64_t pol) {
#ifndef TSB_HOST_EMU
  uint4 v;
  asm("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
#else
  return *p;
#endif
}

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
// NP = 2 / 3: 4 / 8 vectors per work item over pair wires private to slot 0 (see the header).  w[s] = OR over the
// summed-wire combinations kk that slot 0 sees as s of AND over the OTHER operands; out[g] = OR_s slot0[s][g] & w[s].
#ifndef TSB_BITSF_L2HINT
// private streaming: bit 0 = evict-last stores (the result of a step is what the next launch streams back in), bit 1 =
// evict-first fetches of slot 0 (its last use).  Measured on the n=50 bench: slice loop 0.80 -> 0.75 ms with both.
#define TSB_BITSF_L2HINT 3
#endif
#ifndef TSB_BITSF_NP_CTAS
#define TSB_BITSF_NP_CTAS 2   // resident CTAs per SM the private-streaming kernels are compiled for (A/B builds)
#endif
template <int NOPS, int KB, int MASKS, int NP, int NT = 0>
__global__ void __launch_bounds__(kBitsFThreads, NP == 0 ? 3 : NP == 1 ? 2 : TSB_BITSF_NP_CTAS)
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
  [[maybe_unused]] const uint64_t pol_last = bits_f_policy(true), pol_first = bits_f_policy(false);
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
      if constexpr (NP >= 2) {
        constexpr int M0 = bits_f_mask(MASKS, KB, 0), M1 = bits_f_mask(MASKS, KB, 1);
        constexpr int NH = 1 << NT;
        // w[s][h]: everything but slot 0, per summed-wire combination s as slot 0 sees it and combination h of slot 1's wires
        uint4 w[KK][NH];
        const uint4* q1 = base[1] + tab[1][vi];
#pragma unroll
        for (int h = 0; h < NH; h++) {
          uint4 t1[KK];   // slot 1 at combination h
#pragma unroll
          for (int c = 0; c < KK; c++)
            if ((c & ~M1) == 0) t1[c] = h == 0 ? x0[1][c] : __ldg(q1 + d.op[1].pw[h] + d.op[1].kw[c]);
#pragma unroll
          for (int s = 0; s < KK; s++) {
            if ((s & ~M0) != 0) continue;
            w[s][h] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int kk = 0; kk < KK; kk++) {
              if ((kk & M0) != s) continue;
              uint4 t = t1[kk & M1];
#pragma unroll
              for (int o = 2; o < NOPS; o++) {
                const uint4 v = x0[o][kk & bits_f_mask(MASKS, KB, o)];
                t.x &= v.x; t.y &= v.y; t.z &= v.z; t.w &= v.w;
              }
              w[s][h].x |= t.x; w[s][h].y |= t.y; w[s][h].z |= t.z; w[s][h].w |= t.w;
            }
          }
        }
        const uint4* q0 = base[0] + tab[0][vi];
#pragma unroll
        for (int g = 0; g < (1 << NP); g++) {
          uint4 v[KK];
#pragma unroll
          for (int s = 0; s < KK; s++) {
            if ((s & ~M0) != 0) continue;
            // read-only path: these fetches may be scheduled above the stores of the earlier combinations
#if TSB_BITSF_L2HINT & 2
            v[s] = g == 0 ? x0[0][s] : bits_f_ldg_hint(q0 + d.op[0].pw[g] + d.op[0].kw[s], pol_first);
#else
            v[s] = g == 0 ? x0[0][s] : __ldg(q0 + d.op[0].pw[g] + d.op[0].kw[s]);
#endif
          }
#pragma unroll
          for (int h = 0; h < NH; h++) {
            uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int s = 0; s < KK; s++) {
              if ((s & ~M0) != 0) continue;
              acc.x |= v[s].x & w[s][h].x; acc.y |= v[s].y & w[s][h].y; acc.z |= v[s].z & w[s][h].z; acc.w |= v[s].w & w[s][h].w;
            }
#if TSB_BITSF_L2HINT & 1
            bits_f_store_hint(po + d.ow[g] + d.ow2[h], acc, pol_last);
#else
            po[d.ow[g] + d.ow2[h]] = acc;
#endif
          }
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
