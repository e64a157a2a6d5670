// tsb_merge_gen.cu — instantiations of the type-generic merge kernel (counting semiring, mixed dtypes).
#define TSB_MERGE_GENERIC_IMPL
#include "lower.h"
