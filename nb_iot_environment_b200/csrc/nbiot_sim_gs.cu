/*
 * nbiot_sim_gs.cu -- the simulation kernels built with the look-ahead ring always staged in shared memory (see nbiot_simk.cuh).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_GRID_STAGED 1
#define NB_SIM_VARIANT gs
#include "nbiot.h"
#include "nbiot_simk.cuh"
