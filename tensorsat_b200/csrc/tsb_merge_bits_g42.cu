// tsb_merge_bits_g42.cu — k_merge_bits_f instantiations with two pair wires (four vectors per work item, Gray order):
// four operands, two summed wires.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_np2_4_2(int masks) { return bits_f_kernel_np2<4, 2>(masks); }
}  // namespace tsb
