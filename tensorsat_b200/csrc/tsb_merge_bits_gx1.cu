// tsb_merge_bits_gx1.cu — k_merge_bits_f instantiations with two / three private pair wires: three / four operands, one summed wire.
#include "bits_f_dispatch.h"

namespace tsb {
BitsFKernel bits_f_kernel_np2_x_1(int nops, int masks, int n_pair) {
  return nops == 4 ? bits_f_kernel_np2<4, 1>(masks, n_pair) : bits_f_kernel_np2<3, 1>(masks, n_pair);
}
}  // namespace tsb
