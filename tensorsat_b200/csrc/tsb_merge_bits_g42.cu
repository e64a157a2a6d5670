// tsb_merge_bits_g42.cu — k_merge_bits_f instantiations with two / three private pair wires (4 / 8 vectors per work
// item): four operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_np2_4_2(int masks, int n_pair) { return bits_f_kernel_np2<4, 2>(masks, n_pair); }
}  // namespace tsb
