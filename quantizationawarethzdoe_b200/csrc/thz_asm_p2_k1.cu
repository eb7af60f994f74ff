// One kernel family of the static fast path per translation unit (see thz_asm_p2_kernels.inc): K1, the row-FFT kernel.
#define THZ_P2_PART 1
#include "thz_asm_p2_kernels.inc"
