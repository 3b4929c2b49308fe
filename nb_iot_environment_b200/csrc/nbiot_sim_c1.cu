/*
 * nbiot_sim_c1.cu -- the simulation kernels built for single-carrier cells without the debug event log (see nbiot_simk.cuh).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_FIXED_C 1
#define NB_NO_EVLOG 1
#define NB_SIM_VARIANT c1
#include "nbiot.h"
#include "nbiot_simk.cuh"
