// tsb_merge_bits_h42.cu — k_merge_bits_f instantiations, two-sided private streaming (slots 0 and 1 both hold private
// pair wires): four operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_two_4_2(int masks, int n_pair, int n_pair2) { return bits_f_kernel_two<4, 2>(masks, n_pair, n_pair2); }
}  // namespace tsb
