// float / dense instantiations of the pass kernels (see inst_body.inc)
#ifdef NBMF_EXPERIMENTS
#define NBMF_TUNING_DENSE 1
#endif
#define NBMF_REAL float
#define NBMF_DENSE true
#define NBMF_LOOKUP lookup_f32_dense
#include "inst_body.inc"
