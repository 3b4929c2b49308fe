/*
 * nbiot_simt_c1.cu -- the thread-per-cell kernel built for single-carrier cells (see nbiot_simt.cuh).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_TPC 1
/* inlining policy of the thread-per-cell build: single-call-site functions out of line, only the tiny multi-site helpers inlined (+3.5 % on the
 * sweep against the warp-per-cell kernels' policy; inlining everything costs 11 %: profiles/r02_tpc_variants.txt section 11) */
#ifndef NB_INLINE_SINGLE
#define NB_INLINE_SINGLE 0
#endif
#ifndef NB_INLINE_MULTI
#define NB_INLINE_MULTI 1
#endif
#ifdef NB_EXP_TPC_GLOBAL_HDR      /* measurement build: header reached in place in HBM instead of a local-memory copy */
#define NB_TPC_GLOBAL_HDR 1
#endif
#define NB_TASKS 1
#define NB_FIXED_C 1
#define NB_NO_EVLOG 1
#define NB_SIM_VARIANT tpc_c1
#include "nbiot.h"
#include "nbiot_simt.cuh"
