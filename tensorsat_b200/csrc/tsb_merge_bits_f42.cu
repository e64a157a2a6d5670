// tsb_merge_bits_f42.cu — k_merge_bits_f instantiations: four operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_4_2(int masks, int n_pair) { return bits_f_kernel<4, 2>(masks, n_pair); }
}  // namespace tsb
