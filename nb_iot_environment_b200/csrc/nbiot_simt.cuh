/*
 * nbiot_simt.cuh -- the thread-per-cell simulation kernel: one THREAD advances one cell, a warp carries 32 unrelated cells.
 *
 * The warp-per-cell kernel (nbiot_simk.cuh) spends 31/32 of its issue slots on copies of the scalar control logic of the reference
 * (event selection, NodeB FSM, first-fit search, receptions) and is bound by instruction supply, not by memory
 * (DESIGN.md section 6). With the batches of BASELINE.json configs[2..4] (65 536 ... 4 000 000 cells) there are enough cells to
 * fill the machine with one cell per lane: the same core source is compiled with NB_TPC (a "warp" of ONE lane, nbiot_warp.h), the
 * lane-parallel sections become scalar loops, and 32 cells share every instruction they have in common; cells of a warp that
 * sit in different handlers diverge and reconverge at the event loop.
 *
 * State: the per-cell header (NbHdr, 3.5 KB) is copied into the thread's LOCAL memory for the launch. Local memory is
 * interleaved by the hardware -- word w of lane l sits next to word w of lane l+1 -- so the 32 headers of a warp are read and
 * written with coalesced, L1-cached accesses although the source addresses them as one plain struct; the copy in and out
 * is 2 x 3.5 KB per cell per launch. The look-ahead ring, the lists and the UE records stay in HBM and are reached with the
 * cell's own pointers (scattered 8- and 16-byte accesses through L1/L2).
 *
 * Included by one translation unit per build (single carrier / generic), like nbiot_simk.cuh.
 */
#ifndef NBIOT_SIMT_CUH
#define NBIOT_SIMT_CUH

#ifndef NB_TPC
#error "nbiot_simt.cuh needs NB_TPC (the one-lane build of the core)"
#endif
#include "nbiot_core.cuh"
#include "nbiot_launch.h"

#ifndef NB_TPC_THREADS
#define NB_TPC_THREADS 128
#endif
/* 8 CTAs of 128 threads = 32 warps per SM at 64 registers: the kernel waits on memory (local-memory headers and scattered list
 * accesses miss L1), so resident warps count for more than registers: +26 % over 16 warps at 115 registers; 40 and 48 warps are
 * no better (profiles/r02_tpc_variants.txt) */
#ifndef NB_TPC_MIN_CTAS
#define NB_TPC_MIN_CTAS 8
#endif
/* class affinity of the voted kernel (see nbiot_simtv_kernel): bonus of the hinted class in the vote (0 = off), SMs that share a hint,
 * and how rarely a warp without a lane in the hinted class moves the hint (clock() & mask == 0) */
#ifndef NB_TPC_AFFINITY
#define NB_TPC_AFFINITY 1024
#endif
#ifndef NB_TPC_HINT_SMS
#define NB_TPC_HINT_SMS 2
#endif
#ifndef NB_TPC_HINT_RARE
#define NB_TPC_HINT_RARE 3
#endif

#define NB_CAT2(a, b) a##b
#define NB_CAT(a, b) NB_CAT2(a, b)
namespace NB_CAT(nbk_, NB_SIM_VARIANT) {

struct __align__(16) NbLocal { NbCtx c; NbHdr h; };      /* nbc_hdr(c) = the bytes behind the context */

__global__ void __launch_bounds__(NB_TPC_THREADS, NB_TPC_MIN_CTAS)
nbiot_simt_kernel(NbKernelArgs A, const uint8_t *ext_actions, const uint8_t *active, int until_time, int max_steps)
{
    if (NB_NOT_MY_LOAD(A)) return;
    __shared__ NbCtaConst s_ctrl;                      /* the controller tables + cold configuration, once per CTA */
    for (int k = threadIdx.x; k < (int)(sizeof(NbCtaConst) / 4); k += blockDim.x) ((uint32_t *)&s_ctrl)[k] = ((const uint32_t *)A.ctrl)[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.L.n_inst) return;
    if (active && !active[i]) return;
    NbLocal L;
    nb_fill_ctx_launch(L.c, A, &s_ctrl);
    nb_fill_ctx(L.c, A, i);
    uint4 *hg = (uint4 *)((NbHdr *)(A.state + A.L.off_hdr) + i), *hl = (uint4 *)&L.h;
#ifdef NB_TPC_GLOBAL_HDR
    L.c.hdr_g = (NbHdr *)hg;
#else
    NB_NOUNROLL
    for (int k = 0; k < (int)(sizeof(NbHdr) / 16); ++k) hl[k] = hg[k];
#endif
    nbc_blk_rebuild(L.c);
    unsigned long long t0 = 0;
    if (A.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    nbc_controller_advance(L.c, L.c.ctrl, ext_actions ? ext_actions + (size_t)i * L.c.ctrl->ext_k : nullptr, until_time, max_steps);
    if (A.prof) {
        unsigned long long t1; unsigned int smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        A.prof[4 * (size_t)i] = t0; A.prof[4 * (size_t)i + 1] = t1; A.prof[4 * (size_t)i + 2] = smid; A.prof[4 * (size_t)i + 3] = 0ull;
    }
#ifndef NB_TPC_GLOBAL_HDR
    NB_NOUNROLL
    for (int k = 0; k < (int)(sizeof(NbHdr) / 16); ++k) hg[k] = hl[k];
#endif
}

/* The same with VOTED DISPATCH. Left to themselves the 32 cells of a warp are in as many different event handlers as the event
 * mix has classes, and SIMT divergence runs every handler present in turn with two or three active lanes each (measured: 2.5 of 32,
 * profiles/r02a_*). Here the cell's control flow is the task chain of the core (nbiot_core.cuh "tasks": one heavy handler per
 * task + the glue behind it, control state in NbHdr.tk_*): every round the warp votes for ONE task class -- the one with the most
 * waiting lanes, waiting time breaking ties so that nobody starves -- and only the lanes that wait for it run; the others keep their
 * task for a later round, when more lanes have joined their class. A handler is therefore entered by all its lanes together and
 * walked once for all of them. */
__global__ void __launch_bounds__(NB_TPC_THREADS, NB_TPC_MIN_CTAS)
nbiot_simtv_kernel(NbKernelArgs A, const uint8_t *ext_actions, const uint8_t *active, int until_time, int max_steps)
{
    if (NB_NOT_MY_LOAD(A)) return;                     /* automatic kernel choice: this launch belongs to the other kernel */
    __shared__ NbCtaConst s_ctrl;
#if NB_TPC_AFFINITY > 0
    /* Class affinity. Both kernels of this library are short of instruction FETCH, and the two SMs of a TPC share the instruction cache behind
     * their own: when their 64 warps sit in different handlers each one fetches for itself. So the SM pair keeps a hint -- the task class one
     * of its warps picked a moment ago -- in a word of global memory (L2; one volatile load per vote), and a warp that has ANY lane waiting
     * for that class runs it; a warp that has none runs its own best class and, one time in four, makes that the pair's new hint (a hint
     * that moves at every such vote is worth 6 % less, one that moves 1 time in 64 loses the gain: profiles/r02_tpc_variants.txt section 15).
     * Measured +22 % on the sweep point. The hint only reorders the tasks of independent cells: results are bit-identical with it on or off. */
    unsigned int smid_;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
    volatile unsigned int *sm_hint = A.counter + 1 + ((smid_ / NB_TPC_HINT_SMS) % NB_SM_HINTS);
#endif
    for (int k = threadIdx.x; k < (int)(sizeof(NbCtaConst) / 4); k += blockDim.x) ((uint32_t *)&s_ctrl)[k] = ((const uint32_t *)A.ctrl)[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < A.L.n_inst && !(active && !active[i]);
    NbLocal L;
    uint4 *hg = (uint4 *)((NbHdr *)(A.state + A.L.off_hdr) + (mine ? i : 0)), *hl = (uint4 *)&L.h;
    int k = NB_TK_DONE;
    if (mine) {
        nb_fill_ctx_launch(L.c, A, &s_ctrl);
        nb_fill_ctx(L.c, A, i);
#ifdef NB_TPC_GLOBAL_HDR
        L.c.hdr_g = (NbHdr *)hg;
#else
        NB_NOUNROLL
        for (int w = 0; w < (int)(sizeof(NbHdr) / 16); ++w) hl[w] = hg[w];
#endif
        nbc_blk_rebuild(L.c);
        L.c.tk_until = until_time; L.c.tk_max_steps = max_steps; L.c.tk_internal = ext_actions ? 0 : 1;
        k = nbc_task_begin(L.c, ext_actions ? ext_actions + (size_t)i * L.c.ctrl->ext_k : nullptr);
    }
    int waited = 0;
    for (;;) {
        __syncwarp(NB_FULL);
        if (!__ballot_sync(NB_FULL, k != NB_TK_DONE)) break;
        int best = 0, best_score = 0;
#if NB_TPC_AFFINITY > 0
        const int fav = (int)*sm_hint;
#endif
#pragma unroll
        for (int cl = 1; cl < NB_TK_N; ++cl) {
            int sc = __reduce_add_sync(NB_FULL, k == cl ? 4 + waited : 0);
#if NB_TPC_AFFINITY > 0
            if (sc > 0 && cl == fav) sc += NB_TPC_AFFINITY;
#endif
            if (sc > best_score) { best_score = sc; best = cl; }
        }
#if NB_TPC_AFFINITY > 0
        if ((threadIdx.x & 31u) == 0 && best && best != fav && (clock() & NB_TPC_HINT_RARE) == 0) *sm_hint = (unsigned int)best;
#endif
        if (k == best) { k = nbc_task_run(L.c, k); waited = 0; }
        else if (k != NB_TK_DONE) waited += 1;
    }
#ifndef NB_TPC_GLOBAL_HDR
    if (mine) {
        NB_NOUNROLL
        for (int w = 0; w < (int)(sizeof(NbHdr) / 16); ++w) hg[w] = hl[w];
    }
#endif
}

}   /* namespace nbk_<variant> */

void NB_CAT(nb_sim_variant_, NB_SIM_VARIANT)(NbSimVariant *out)
{
    using namespace NB_CAT(nbk_, NB_SIM_VARIANT);
    out->sim_lo = (const void *)nbiot_simt_kernel;
    out->sim_hi = (const void *)nbiot_simtv_kernel;
    out->reset = nullptr;                               /* construction and reset run on the warp-per-cell kernels */
    out->warps_per_block = NB_TPC_THREADS / 32;
}

#endif /* NBIOT_SIMT_CUH */
