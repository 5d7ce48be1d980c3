// float arithmetic / dense V*mask stored as fp16 (see inst_body.inc): halves the HBM bytes of probabilistic V
#include <cuda_fp16.h>
#define NBMF_REAL float
#define NBMF_VT __half
#define NBMF_DENSE true
#define NBMF_LOOKUP lookup_f32_dense16
#include "inst_body.inc"
