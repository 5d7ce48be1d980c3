// float / bits instantiations of the pass kernels (see inst_body.inc)
#define NBMF_TUNING 1
#define NBMF_REAL float
#define NBMF_DENSE false
#define NBMF_LOOKUP lookup_f32_bits
#include "inst_body.inc"
