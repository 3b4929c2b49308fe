/*
 * nbiot_launch.h -- what the two translation units of libnbiot_b200.so share (internal, not part of the C ABI):
 * the kernel argument block, the staging helpers, and the entry points of the builds of the simulation kernels.
 */
#ifndef NBIOT_LAUNCH_H
#define NBIOT_LAUNCH_H

#include <cuda_runtime.h>
#include <stdint.h>
#include "nbiot_core.cuh"

struct NbKernelArgs {
    uint8_t *state;
    NbLayout L;
    const nbiot_conf_t *conf;
    const NbCtaConst *ctrl;      /* controller tables + cold configuration */
    const uint16_t *lut2;
    const uint8_t *mcs;
    /* automatic kernel choice, decided on the device so that it is never stale: load_sum = live UEs of the batch (nbiot_load_kernel, launched
     * just before); a kernel launched with load_side 0 runs only while load_sum <= load_max (light cells: thread-per-cell), with load_side 1
     * only above it (warp-per-cell); the other one returns at once. NULL: unconditional. */
    const unsigned long long *load_sum; unsigned long long load_max; int load_side;
    unsigned int *counter;          /* [0] claim counter, [1 .. NB_SM_HINTS] per-SM class hints (-DNB_TPC_AFFINITY_SM builds of the voted kernel) */
    unsigned long long *prof;    /* optional (NBIOT_PROF_INST=1): per instance { start ns, end ns, SM id, 0 } of the last launch (measurement aid) */
    int32_t *evlog; int evlog_cap;
    int smem_per_warp, hdr_bytes16, grid_bytes16;
    int stage_grid;              /* 1: the look-ahead ring is staged in shared memory; 0: used in place in HBM (short launches) */
    uint16_t *blk;               /* thread-per-cell kernel: [n_inst][C][W / 32] block summaries of the rings (derived data, rebuilt every launch) */
    /* one-wave launches of the warp-per-cell kernel: cells dealt to SMs by predicted cost (nbiot_deal_kernel). deal_order[sm][k] = k-th cell of
     * SM sm's list, deal_len[sm] its length, deal_cnt[sm] the claim counter; NULL: cells are claimed in index order from `counter`. */
    const int32_t *deal_order; const int32_t *deal_len; unsigned int *deal_cnt; int deal_lists, deal_stride;
    uint32_t *dur;               /* per cell: duration of its last long launch in ns (the cost predictor's memory); NULL: not recorded */
};


#if defined(__CUDACC__)
__device__ __forceinline__ void nb_copy16(void *dst, const void *src, int n16)
{
    const uint4 *s = (const uint4 *)src; uint4 *d = (uint4 *)dst;
    for (int i = threadIdx.x & 31; i < n16; i += 32) d[i] = s[i];
}

/* the part of the context that is the same for every instance of a launch: once per warp */
__device__ __forceinline__ void nb_fill_ctx_launch(NbCtx &c, const NbKernelArgs &A, const NbCtaConst *s_cc)
{
    const NbLayout &L = A.L;
    c.lut2 = A.lut2; c.mcs = A.mcs; c.conf = *(const nbiot_conf_hot_t *)A.conf; c.ctrl = &s_cc->ctrl;
    c.W = L.grid_w; c.ue_cap = L.ue_cap; c.ra_cap = L.ra_cap; c.hist_cap = L.hist_cap; c.C = L.n_carriers;
    c.evlog = nullptr; c.evlog_cap = A.evlog_cap;
}

/* the per-instance part: where the instance's arrays live in the state buffer */
__device__ __forceinline__ void nb_fill_ctx(NbCtx &c, const NbKernelArgs &A, int i)
{
    const NbLayout &L = A.L;
    uint8_t *s = A.state;
    for (int k = 0; k < 3; ++k) c.ra[k] = (NbRaEntry *)(s + L.off_ra) + ((size_t)i * 3 + k) * L.ra_cap;
    c.conn = (NbConnEntry *)(s + L.off_conn) + (size_t)i * L.ue_cap;
    c.cont = (NbContEntry *)(s + L.off_cont) + (size_t)i * L.ue_cap;
    c.ue = (NbUe *)(s + L.off_ue) + (size_t)i * L.ue_cap;
    c.free_stack = (uint16_t *)(s + L.off_free) + (size_t)i * L.ue_cap;
    c.occ = (NbOccHist *)(s + L.off_occ) + i;
    c.incoming = (double *)(s + L.off_incoming) + (size_t)i * L.hist_cap;
    c.delays = (int32_t *)(s + L.off_delays) + (size_t)i * L.hist_cap;
    c.service = (int32_t *)(s + L.off_service) + (size_t)i * L.hist_cap;
    c.scratch = (uint16_t *)(s + L.off_scratch) + (size_t)i * L.ra_cap;
    c.rec = (nbiot_record_t *)(s + L.off_rec) + i;
    c.spill = (int32_t *)(s + L.off_spill) + (size_t)i * 6u * nb_spill_cap(L.hist_cap);
    c.evlog = i == 0 ? A.evlog : nullptr;
#if NB_LANES == 1
    c.blk = A.blk ? A.blk + (size_t)i * A.L.n_carriers * (A.L.grid_w >> 5) : nullptr;
#endif
    c.grid = A.stage_grid ? (uint16_t *)((unsigned char *)&c + sizeof(NbCtx) + sizeof(NbHdr))
                          : (uint16_t *)(A.state + A.L.off_grid) + (size_t)i * A.L.n_carriers * A.L.grid_w;
}

#endif

/* ---- builds of the warp-per-instance kernels (nbiot_simk.cuh) ---- */
struct NbSimVariant { const void *sim_lo, *sim_hi, *reset; int warps_per_block; };
void nb_sim_variant_generic(NbSimVariant *out);
void nb_sim_variant_c1(NbSimVariant *out);
void nb_sim_variant_gs(NbSimVariant *out);
void nb_sim_variant_c1s(NbSimVariant *out);
/* ---- builds of the thread-per-cell kernel (nbiot_simt.cuh) ---- */
void nb_sim_variant_tpc_gen(NbSimVariant *out);
void nb_sim_variant_tpc_c1(NbSimVariant *out);
#ifndef NB_TPC_MIN_CELLS
#define NB_SM_HINTS 256
#define NB_NOT_MY_LOAD(A) ((A).load_sum && ((*(A).load_sum > (A).load_max) != ((A).load_side == 1)))
#ifndef NB_TPC_MAX_LOAD
#define NB_TPC_MAX_LOAD 140.0     /* mean live UEs per cell above which a long launch runs on the warp-per-cell kernel (27: 2.9e9 against 1.4e9; 166: 0.87e9
                                     against 1.01e9; thresholds 40 / 70 / 100 / 140 / 200 on the bursty scenario at 65 536 cells: 1.23 / 1.28 / 1.30 / 1.36 / 1.35e9) */
#endif
#define NB_TPC_MIN_CELLS 32768      /* batches from this size on advance on the thread-per-cell kernel */
#endif

#endif /* NBIOT_LAUNCH_H */
