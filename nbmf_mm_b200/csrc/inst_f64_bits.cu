// double / bits instantiations of the pass kernels (see inst_body.inc)
#define NBMF_REAL double
#define NBMF_DENSE false
#define NBMF_LOOKUP lookup_f64_bits
#include "inst_body.inc"
