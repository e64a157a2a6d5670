// bits_f_dispatch.h — instantiation tables of k_merge_bits_f (one translation unit per (NOPS, KB) group,
// see tsb_merge_bits_f*.cu): only mask tuples sorted in descending order are instantiated.
#pragma once
#include <utility>

#include "lower.h"
#include "k_merge_bits_f.cuh"

namespace tsb {

using BitsFKernel = void (*)(const BitsFDesc, const BitsNPtrs, uint4*, const uint32_t*);

constexpr bool bits_f_sorted(int masks, int kb, int nops) {
  for (int o = 0; o + 1 < nops; o++)
    if (bits_f_mask(masks, kb, o) < bits_f_mask(masks, kb, o + 1)) return false;
  return true;
}

// internal linkage: a static local of an inline function would be a process-wide unique symbol, shared with any
// other build of these sources loaded into the same process (the test suite loads the emulated build too)
namespace {

template <int NOPS, int KB, int NP, int NT, int MASKS>
constexpr BitsFKernel bits_f_entry() {
  if constexpr (bits_f_sorted(MASKS, KB, NOPS)) return k_merge_bits_f<NOPS, KB, MASKS, NP, NT>;
  else return nullptr;
}

template <int NOPS, int KB, int NP, int NT = 0, int... I>
inline BitsFKernel bits_f_lookup(int masks, std::integer_sequence<int, I...>) {
  static const BitsFKernel table[] = {bits_f_entry<NOPS, KB, NP, NT, I>()...};
  return table[masks];
}

template <int NOPS, int KB>
inline BitsFKernel bits_f_kernel(int masks, int n_pair) {
  constexpr int N = 1 << (KB * NOPS);
  auto seq = std::make_integer_sequence<int, N>{};
  return n_pair == 0 ? bits_f_lookup<NOPS, KB, 0, 0>(masks, seq) : bits_f_lookup<NOPS, KB, 1, 0>(masks, seq);
}

// the private-streaming instantiations (two / three pair wires held by slot 0 only): one translation unit per group
// n_pair in {2, 3}, one-sided
template <int NOPS, int KB>
inline BitsFKernel bits_f_kernel_np2(int masks, int n_pair) {
  constexpr int N = 1 << (KB * NOPS);
  auto seq = std::make_integer_sequence<int, N>{};
  return n_pair == 2 ? bits_f_lookup<NOPS, KB, 2, 0>(masks, seq) : bits_f_lookup<NOPS, KB, 3, 0>(masks, seq);
}

// the two-sided forms, (n_pair, n_pair2) in {2, 3} x {1, 2}: translation units of their own (groups with two summed
// wires; the four-operand group is split by n_pair2, it is the longest compile of the library)
template <int NOPS, int KB, int NT>
inline BitsFKernel bits_f_kernel_two(int masks, int n_pair) {
  constexpr int N = 1 << (KB * NOPS);
  auto seq = std::make_integer_sequence<int, N>{};
  return n_pair == 2 ? bits_f_lookup<NOPS, KB, 2, NT>(masks, seq) : bits_f_lookup<NOPS, KB, 3, NT>(masks, seq);
}

}  // namespace

// one per translation unit
BitsFKernel bits_f_kernel_4_2(int masks, int n_pair);
BitsFKernel bits_f_kernel_3_2(int masks, int n_pair);
BitsFKernel bits_f_kernel_small(int nops, int kb, int masks, int n_pair);  // everything else
BitsFKernel bits_f_kernel_np2_4_2(int masks, int n_pair);   // n_pair = 2, 3 exist for the (4, 2), (3, 2), (4, 1), (3, 1) groups only
BitsFKernel bits_f_kernel_np2_3_2(int masks, int n_pair);
BitsFKernel bits_f_kernel_two1_4_2(int masks, int n_pair);   // two-sided, one / two private wires in slot 1: (4, 2) and (3, 2) only
BitsFKernel bits_f_kernel_two2_4_2(int masks, int n_pair);
BitsFKernel bits_f_kernel_two_3_2(int masks, int n_pair, int n_pair2);
BitsFKernel bits_f_kernel_np2_x_1(int nops, int masks, int n_pair);
inline bool bits_f_has_np2(int nops, int kb) { return (nops == 3 || nops == 4) && (kb == 1 || kb == 2); }
inline bool bits_f_has_two_sided(int nops, int kb) { return (nops == 3 || nops == 4) && kb == 2; }

}  // namespace tsb
