// tsb_merge_bits_i42.cu — k_merge_bits_f instantiations, two-sided private streaming with TWO private wires in slot 1:
// four operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_two2_4_2(int masks, int n_pair) { return bits_f_kernel_two<4, 2, 2>(masks, n_pair); }
}  // namespace tsb
