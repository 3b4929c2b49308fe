/*
 * nbiot_simt_c1.cu -- the thread-per-cell kernel built for single-carrier cells (see nbiot_simt.cuh).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_TPC 1
#ifdef NB_EXP_TPC_GLOBAL_HDR      /* measurement build: header reached in place in HBM instead of a local-memory copy */
#define NB_TPC_GLOBAL_HDR 1
#endif
#define NB_TASKS 1
#define NB_FIXED_C 1
#define NB_NO_EVLOG 1
#define NB_SIM_VARIANT tpc_c1
#include "nbiot.h"
#include "nbiot_simt.cuh"
