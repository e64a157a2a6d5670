// tsb_merge_bits_h42.cu — k_merge_bits_f instantiations, two-sided private streaming with ONE private wire in slot 1:
// four operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_two1_4_2(int masks, int n_pair) { return bits_f_kernel_two<4, 2, 1>(masks, n_pair); }
}  // namespace tsb
