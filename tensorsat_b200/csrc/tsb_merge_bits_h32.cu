// tsb_merge_bits_h32.cu — k_merge_bits_f instantiations, two-sided private streaming (slots 0 and 1 both hold private
// pair wires): three operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_two_3_2(int masks, int n_pair, int n_pair2) {
  return n_pair2 == 1 ? bits_f_kernel_two<3, 2, 1>(masks, n_pair) : bits_f_kernel_two<3, 2, 2>(masks, n_pair);
}
}  // namespace tsb
