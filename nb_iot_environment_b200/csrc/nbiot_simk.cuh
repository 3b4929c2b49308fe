/*
 * nbiot_simk.cuh -- the simulation kernels (warp-per-instance): included by one translation unit per BUILD of the core.
 *   nbiot_kernels.cu : generic build (1 or 2 carriers, optional event log)                    NB_SIM_VARIANT = generic
 *   nbiot_sim_gs.cu  : generic + NB_GRID_STAGED (long launches: the ring is always in shared memory)   NB_SIM_VARIANT = gs
 *   nbiot_sim_c1s.cu : c1 + NB_GRID_STAGED                                                             NB_SIM_VARIANT = c1s
 *   nbiot_sim_c1.cu  : NB_FIXED_C = 1, NB_NO_EVLOG -- the carrier loops and ring addressing compiled for single-carrier
 *                      cells: 3 % less SASS and 6-8 % faster on the headline workload (the kernel is bound by instruction
 *                      supply, profiles/r01_icache.txt, so less code pays twice)                NB_SIM_VARIANT = c1
 * Each build exports nb_sim_variant_<name>(), which hands its kernel entry points to the launcher in nbiot_kernels.cu.
 */
#ifndef NBIOT_SIMK_CUH
#define NBIOT_SIMK_CUH

#include "nbiot_core.cuh"
#include "nbiot_launch.h"

#ifndef NB_WPB
#define NB_WPB 4   /* warps (instances in flight) per CTA */
#endif
/* The hot kernel is built twice: MINB = 7 (72 registers: the budget that lets 7 CTAs = 28 warps live on an SM, which is what the
 * shared-memory slice of a grid_w = 2048 instance allows) and MINB = 9 (56 registers, 36 warps: what a grid_w = 1024 slice allows).
 * Instruction supply, not registers, bounds this kernel (profiles/r01_icache.txt), so the extra warps win wherever shared memory
 * lets them be resident: +11 % on the config-5 sweep. nbiot_create picks the build with the most resident warps for the batch. */
#ifndef NB_MIN_CTAS_LO
#define NB_MIN_CTAS_LO 7
#endif
#ifndef NB_MIN_CTAS_HI
#define NB_MIN_CTAS_HI 9
#endif

#define NB_CAT2(a, b) a##b
#define NB_CAT(a, b) NB_CAT2(a, b)
namespace NB_CAT(nbk_, NB_SIM_VARIANT) {     /* the kernels of this build: nbk_generic::nbiot_sim_kernel<7>, nbk_c1::nbiot_sim_kernel<7>, ... */

/* Dealt launches (nbiot_deal_kernel): cells were dealt to SMs so that every SM holds the same predicted work. Take the next cell of the list
 * of the SM this warp runs on; a warp whose list is used up (the block scheduler gave its SM more warps than the list has cells) takes what other
 * lists have left. Returns n_inst when nothing is left anywhere. The loop is WARP-UNIFORM -- lane 0 claims, every lane follows the outcome --
 * so that no lane leaves a loop the others never entered (a lane-0-only probing loop with a break in it left the warp's convergence state
 * in a condition the simulation core did not survive: hangs and illegal-instruction faults as soon as a warp had to probe a second list). */
__device__ __forceinline__ int nb_claim_dealt(const NbKernelArgs &A, int lane, int &probe)
{
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    _Pragma("unroll 1")
    while (probe < A.deal_lists) {
        int got = -1;
        if (lane == 0) {
            const unsigned int s2 = (smid + (unsigned int)probe) % (unsigned int)A.deal_lists;
            const unsigned int slot = atomicAdd(&A.deal_cnt[s2], 1u);
            if ((int)slot < A.deal_len[s2]) got = A.deal_order[(size_t)s2 * A.deal_stride + slot];
        }
        got = __shfl_sync(0xFFFFFFFFu, got, 0);
        if (got >= 0) return got;
        probe += 1;
    }
    return A.L.n_inst;
}

/* Claim instances from the work counter; stage [ctx | header | grid] in this warp's shared-memory slice;
 * run `body`; write header and grid back. */
template <bool LOAD, class Body>
__device__ __forceinline__ void nb_for_each_instance(const NbKernelArgs &A, const uint8_t *active, Body body)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ NbCtaConst s_ctrl;                      /* the controller tables + cold configuration, once per CTA */
    for (int k = threadIdx.x; k < (int)(sizeof(NbCtaConst) / 4); k += blockDim.x) ((uint32_t *)&s_ctrl)[k] = ((const uint32_t *)A.ctrl)[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *my = smem + (size_t)warp * A.smem_per_warp;
    NbCtx &c = *(NbCtx *)my;
    unsigned char *hdr_s = my + sizeof(NbCtx), *grid_s = hdr_s + sizeof(NbHdr);
    NbHdr *hdr_g = (NbHdr *)(A.state + A.L.off_hdr);
    uint16_t *grid_g = (uint16_t *)(A.state + A.L.off_grid);
    const size_t grid_elems = (size_t)A.L.n_carriers * A.L.grid_w;
    if (lane == 0) nb_fill_ctx_launch(c, A, &s_ctrl);
    int probe = 0;                                     /* dealt launches: lists tried so far (own SM's first, then the others') */
    for (;;) {
        int i = 0;
        if (A.deal_order) i = nb_claim_dealt(A, lane, probe);
        else {
            if (lane == 0) i = (int)atomicAdd(A.counter, 1u);
            i = __shfl_sync(0xFFFFFFFFu, i, 0);
        }
        if (i >= A.L.n_inst) break;
        if (active && !active[i]) continue;
        if (lane == 0) nb_fill_ctx(c, A, i);
        if (LOAD) {
            nb_copy16(hdr_s, hdr_g + i, A.hdr_bytes16);
            if (A.stage_grid) nb_copy16(grid_s, grid_g + (size_t)i * grid_elems, A.grid_bytes16);
        } else if (A.stage_grid) {
            /* construction: the ring starts empty (columns are cleared again when they are generated, so this only keeps
             * whatever was in shared memory out of the state buffer: checkpoints of equal batches are equal byte for byte) */
            for (int k = lane; k < A.grid_bytes16; k += 32) ((uint4 *)grid_s)[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        unsigned long long t0 = 0;
        if (A.prof || A.dur) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        body(c, i);
        if (A.dur && lane == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            A.dur[i] = (uint32_t)(t1 - t0);
        }
        if (A.prof && lane == 0) {
            unsigned long long t1; unsigned int smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            A.prof[4 * (size_t)i] = t0; A.prof[4 * (size_t)i + 1] = t1; A.prof[4 * (size_t)i + 2] = smid; A.prof[4 * (size_t)i + 3] = 0ull;
        }
        __syncwarp();
        nb_copy16(hdr_g + i, hdr_s, A.hdr_bytes16);
        if (A.stage_grid) nb_copy16(grid_g + (size_t)i * grid_elems, grid_s, A.grid_bytes16);
        __syncwarp();
    }
}

/* the hot kernel: Controller.step / run_until for every instance */
template <int MINB>
__global__ void __launch_bounds__(NB_WPB * 32, MINB)
nbiot_sim_kernel(NbKernelArgs A, const uint8_t *ext_actions, const uint8_t *active, int until_time, int max_steps)
{
    if (NB_NOT_MY_LOAD(A)) return;                     /* automatic kernel choice: this launch belongs to the other kernel */
    nb_for_each_instance<true>(A, active, [&](NbCtx &c, int i) {
        nbc_controller_advance(c, c.ctrl, ext_actions ? ext_actions + (size_t)i * c.ctrl->ext_k : nullptr, until_time, max_steps);
    });
}

/* create_system + NodeB.reset (construct = 1) or NodeB.reset alone */
__global__ void __launch_bounds__(NB_WPB * 32)
nbiot_reset_kernel(NbKernelArgs A, int construct, const uint64_t *seeds)
{
    if (construct) nb_for_each_instance<false>(A, nullptr, [&](NbCtx &c, int i) { nbc_construct(c, seeds[i], c.ctrl); nbc_reset(c); });
    else nb_for_each_instance<true>(A, nullptr, [&](NbCtx &c, int) { nbc_reset(c); });
}


}   /* namespace nbk_<variant> */

void NB_CAT(nb_sim_variant_, NB_SIM_VARIANT)(NbSimVariant *out)
{
    using namespace NB_CAT(nbk_, NB_SIM_VARIANT);
    out->sim_lo = (const void *)nbiot_sim_kernel<NB_MIN_CTAS_LO>;
    out->sim_hi = (const void *)nbiot_sim_kernel<NB_MIN_CTAS_HI>;
    out->reset = (const void *)nbiot_reset_kernel;
    out->warps_per_block = NB_WPB;
}

#endif /* NBIOT_SIMK_CUH */
