// tsb_merge_bits_g32.cu — k_merge_bits_f instantiations with two / three private pair wires: three operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_np2_3_2(int masks, int n_pair) { return bits_f_kernel_np2<3, 2>(masks, n_pair); }
}  // namespace tsb
