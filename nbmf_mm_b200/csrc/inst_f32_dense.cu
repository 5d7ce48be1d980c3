// float / dense instantiations of the pass kernels (see inst_body.inc)
#define NBMF_REAL float
#define NBMF_DENSE true
#define NBMF_LOOKUP lookup_f32_dense
#include "inst_body.inc"
