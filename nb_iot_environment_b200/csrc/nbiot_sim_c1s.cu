/*
 * nbiot_sim_c1s.cu -- the simulation kernels built for single-carrier cells, no debug event log, look-ahead ring always staged in shared memory (see nbiot_simk.cuh).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_FIXED_C 1
#define NB_NO_EVLOG 1
#define NB_GRID_STAGED 1
#define NB_SIM_VARIANT c1s
#include "nbiot.h"
#include "nbiot_simk.cuh"
