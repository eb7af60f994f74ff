// One kernel family of the static fast path per translation unit (see thz_asm_p2_kernels.inc): K3, the row-iFFT kernel.
#define THZ_P2_PART 3
#include "thz_asm_p2_kernels.inc"
