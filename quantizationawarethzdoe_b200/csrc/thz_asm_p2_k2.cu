// One kernel family of the static fast path per translation unit (see thz_asm_p2_kernels.inc): K2, the column kernels (general and fast path).
#define THZ_P2_PART 2
#include "thz_asm_p2_kernels.inc"
