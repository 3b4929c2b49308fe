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
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(A.counter, 1u);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= A.L.n_inst) break;
        if (active && !active[i]) continue;
        if (lane == 0) nb_fill_ctx(c, A, i, &s_ctrl);
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
        if (A.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        body(c, i);
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

/* Time-sliced launch. A launch of N cells on about N warp slots is ONE wave: it ends with its slowest cell on its busiest SM
 * while the warps of the SMs that finished early idle (measured on the headline workload: a quarter of the warp-time,
 * profiles/r01c_tail.txt), and the next launch's cost of a cell cannot be predicted from the last one (the agent's action changes
 * it). So a warp advances a cell by at most `quantum` simulated subframes, writes it back and, if the cell has not reached the end
 * of the launch, appends it to a global ring; a free warp -- mostly one of an SM that has run out of cells -- takes the oldest
 * waiting cell. While every warp is busy a warp simply gets its own cell back; towards the end the remaining cells spread over all
 * SMs, where each of them also finds less competition for the instruction fetch. Cutting an advance into several
 * nbc_controller_advance calls does not change a single result: the controller's loop re-enters exactly where it stopped (the record
 * written at the end of a slice has the side effects of the node_info the loop would have called). NbHdr.tk_slice marks a cell whose
 * launch is under way (its external action has been applied). */
__device__ __forceinline__ void nb_run_quanta(const NbKernelArgs &A, const uint8_t *ext_actions, const uint8_t *active, int until_time, int max_steps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ NbCtaConst s_ctrl;
    for (int k = threadIdx.x; k < (int)(sizeof(NbCtaConst) / 4); k += blockDim.x) ((uint32_t *)&s_ctrl)[k] = ((const uint32_t *)A.ctrl)[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *my = smem + (size_t)warp * A.smem_per_warp;
    NbCtx &c = *(NbCtx *)my;
    unsigned char *hdr_s = my + sizeof(NbCtx), *grid_s = hdr_s + sizeof(NbHdr);
    NbHdr *hdr_g = (NbHdr *)(A.state + A.L.off_hdr);
    uint16_t *grid_g = (uint16_t *)(A.state + A.L.off_grid);
    const size_t grid_elems = (size_t)A.L.n_carriers * A.L.grid_w;
    NbRequeue *rq = A.rq;
    bool fresh = true;                                 /* cells nobody has touched in this launch are left in the work counter */
    for (;;) {
        int i = -1;
        if (fresh) {
            if (lane == 0) i = (int)atomicAdd(A.counter, 1u);
            i = __shfl_sync(0xFFFFFFFFu, i, 0);
            if (i >= A.L.n_inst) { fresh = false; i = -1; }
            else if (active && !active[i]) continue;
        }
        if (i < 0) {
            /* take a ticket of the ring and wait for its slot: filled by the next warp that parks a cell, or never (launch over) */
            unsigned int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&rq->head, 1u);
            ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
            unsigned int *sp = &A.rq_slots[ticket & A.rq_mask];
            unsigned int v = 0; int rem = 1;
            for (;;) {
                if (lane == 0) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sp) : "memory");
                    if (!v) rem = *(volatile int *)&rq->remaining;
                }
                v = __shfl_sync(0xFFFFFFFFu, v, 0); rem = __shfl_sync(0xFFFFFFFFu, rem, 0);
                if (v || rem <= 0) break;
                __nanosleep(1000);
            }
            if (!v) break;
            if (lane == 0) *(volatile unsigned int *)sp = 0u;
            i = (int)v - 1;
        }
        if (lane == 0) nb_fill_ctx(c, A, i, &s_ctrl);
        {   /* the last writer may have been another SM: read through L2 */
            const uint4 *src = (const uint4 *)(hdr_g + i); uint4 *dst = (uint4 *)hdr_s;
            for (int k = lane; k < A.hdr_bytes16; k += 32) dst[k] = __ldcg(src + k);
            src = (const uint4 *)(grid_g + (size_t)i * grid_elems); dst = (uint4 *)grid_s;
            for (int k = lane; k < A.grid_bytes16; k += 32) dst[k] = __ldcg(src + k);
        }
        __syncwarp();
        NbHdr *h = (NbHdr *)hdr_s;
        const int first = h->tk_slice == 0;
        nbc_controller_advance(c, c.ctrl, first && ext_actions ? ext_actions + (size_t)i * c.ctrl->ext_k : nullptr, until_time, max_steps, h->time + A.quantum);
        const int done = (h->fault & (int32_t)NBIOT_FAULT_FATAL_MASK) || ((c.ctrl->ext_state_mask >> h->state) & 1u) ||
                         (until_time >= 0 && h->time >= until_time);
        __syncwarp();
        if (lane == 0) h->tk_slice = done ? 0 : 1;
        __syncwarp();
        nb_copy16(hdr_g + i, hdr_s, A.hdr_bytes16);
        nb_copy16(grid_g + (size_t)i * grid_elems, grid_s, A.grid_bytes16);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            if (done) atomicSub(&rq->remaining, 1);
            else {
                unsigned int t = atomicAdd(&rq->tail, 1u);
                asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(&A.rq_slots[t & A.rq_mask]), "r"((unsigned int)i + 1u) : "memory");
            }
        }
        __syncwarp();
    }
}

/* number of cells a time-sliced launch has to finish */
__global__ void nb_requeue_init_kernel(NbRequeue *rq, const uint8_t *active, int n_inst)
{
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n_inst; i += blockDim.x) mine += active ? (active[i] ? 1 : 0) : 1;
    atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) { rq->head = 0u; rq->tail = 0u; rq->remaining = cnt; }
}

/* the hot kernel: Controller.step / run_until for every instance */
template <int MINB>
__global__ void __launch_bounds__(NB_WPB * 32, MINB)
nbiot_sim_kernel(NbKernelArgs A, const uint8_t *ext_actions, const uint8_t *active, int until_time, int max_steps)
{
    if (A.quantum > 0) { nb_run_quanta(A, ext_actions, active, until_time, max_steps); return; }
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
    out->requeue_init = (const void *)nb_requeue_init_kernel;
    out->warps_per_block = NB_WPB;
}

#endif /* NBIOT_SIMK_CUH */
