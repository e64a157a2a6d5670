// tsb_gemm.cu — tcgen05 / TMEM / TMA int8 GEMM (north-star item (c)).  Placeholder until the
// kernel lands: the entry point exists so that the ABI is complete.
#include "lower.h"

extern "C" int tsb_gemm_u8(tsb_ctx*, const void*, const void*, void*, int, int, int, int, int, int, tsb_stream) {
  tsb::set_error("tsb_gemm_u8: not built yet");
  return TSB_ERR_UNSUPPORTED;
}
