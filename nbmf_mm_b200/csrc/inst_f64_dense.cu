// double / dense instantiations of the pass kernels (see inst_body.inc)
#define NBMF_REAL double
#define NBMF_DENSE true
#define NBMF_LOOKUP lookup_f64_dense
#include "inst_body.inc"
