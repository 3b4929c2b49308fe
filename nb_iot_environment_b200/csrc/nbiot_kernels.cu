/*
 * nbiot_kernels.cu -- sm_100a kernels and the C ABI (include/nbiot.h) of the batched NB-IoT cell
 * simulator. One warp owns one cell instance for the whole launch:
 *
 *   1. the warp claims an instance index from a global work counter (persistent grid sized to
 *      SMs x resident CTAs, so uneven per-instance work balances itself);
 *   2. it stages the instance's hot state -- the NbHdr header and the carrier look-ahead ring --
 *      from HBM into its slice of shared memory with 16-byte coalesced copies;
 *   3. it runs the simulation core (nbiot_core.cuh) until the instance needs an external decision
 *      (or reaches the time / step bound): the clock jumps from event to event, list scans run
 *      32 entries at a time with ballot / popc ranks, the UE records, lists and info side buffers
 *      stay in HBM (instance-major, contiguous per instance);
 *   4. it writes the header and grid back with the same coalesced copies.
 *
 * The path is integer / control work on HBM-resident state: no tensor cores, no GEMM. The BLER and
 * MCS tables are read through the read-only path and stay L2-resident (they are touched about 0.03
 * times per simulated subframe, so staging 100 KB of tables per CTA would cost more than it saves;
 * see DESIGN.md).
 */
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "nbiot.h"
#include "nbiot_core.cuh"
#include "nbiot_launch.h"

struct nbiot_handle {
    nbiot_conf_t conf;
    nbiot_ctrl_t ctrl;
    NbLayout L;
    int device, n_inst;
    uint8_t *state;
    nbiot_conf_t *conf_dev;
    NbCtaConst *ctrl_dev;
    uint16_t *lut2_dev;
    uint8_t *mcs_dev;
    uint64_t *seeds_dev;
    unsigned int *counter_dev;
    unsigned long long *prof_dev;   /* NBIOT_PROF_INST=1: per-instance timing of the last launch */
    uint16_t *blk_dev;              /* thread-per-cell kernel: block summaries of the look-ahead rings (derived data) */
    uint32_t *dur_dev;              /* per cell: duration of its last long warp-per-cell launch (cost predictor of the dealing) */
    int32_t *deal_order, *deal_len; unsigned int *deal_cnt; int deal_lists, deal_stride, deal_pad, deal;   /* dealing by predicted cost (one-wave launches) */
    int blocks, smem_bytes, sim_hi;
    int blocks_short, smem_short, sim_hi_short, stage_policy;     /* geometry of launches that leave the grid in HBM */
    NbSimVariant k_gen, k_c1, k_gs, k_c1s;   /* builds of the warp-per-instance kernels (nbiot_simk.cuh): generic / single carrier, ring anywhere / staged */
    NbSimVariant k_tpc_gen, k_tpc_c1;        /* builds of the thread-per-cell kernel (nbiot_simt.cuh) */
    int use_tpc;                 /* long advances run on the thread-per-cell kernel: 1 plain, 2 voted dispatch (large batches) */
    int tpc_forced;              /* NBIOT_KERNEL=tpc|tpcv: every advance, short launches included (tests, measurements) */
    /* NBIOT_KERNEL_AUTO on a batch large enough for the thread-per-cell kernel: a tiny kernel sums the live UEs of the batch in front of every
     * long launch, BOTH simulation kernels are launched, and each looks at the sum: the thread-per-cell one runs while the mean is at most
     * max_load live UEs per cell, the warp-per-cell one above it, the other returns at once (cells with long lists are warp-per-cell work at
     * any batch size: profiles/r02_tpc_variants.txt section 16). Decided on the device, so a host that enqueues many steps ahead cannot make
     * it stale; the sum is also copied back asynchronously, for nbiot_kernel_name only. */
    int auto_kernel, load_pending, last_long_tpc;
    double max_load;
    unsigned long long *load_dev, *load_pin;
    cudaEvent_t load_ev;
    int use_staged_builds;       /* measurement knob NBIOT_SIM_STAGED_BUILDS=0: staged launches on the ring-anywhere builds */
    int use_c1;                  /* single-carrier batch: the c1 build (the generic one while an event log is attached) */
    long long launches;
    int32_t *evlog; int evlog_cap;
    /* staging for nbiot_step_host */
    uint8_t *act_dev; double *obs_dev, *rew_dev; nbiot_obs_item_t *items_dev;
    uint8_t *act_pin; double *obs_pin, *rew_pin;
    size_t act_cap, obs_cap, items_cap;
    std::vector<nbiot_obs_item_t> items_host;    /* what items_dev holds */
};

static thread_local char g_err[512] = "";
static int set_err(int code, const char *fmt, const char *a = "", const char *b = "")
{ snprintf(g_err, sizeof g_err, fmt, a, b); return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(NBIOT_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

/* ------------------------------------------------------------------ device side */
/* the simulation kernels live in nbiot_simk.cuh: this translation unit holds their generic build (1 or 2 carriers, event log) */
#define NB_SIM_VARIANT generic
#include "nbiot_simk.cuh"

/* Controller.get_obs (controller.py:105-120) for the whole batch: one thread per (instance, obs element) */
__global__ void nbiot_obs_kernel(const nbiot_record_t *rec, int n_inst, const nbiot_obs_item_t *items, int n_items, int obs_dim,
                                 double *obs, double *reward)
{
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int total = n_inst * (obs_dim > 0 ? obs_dim : 1);
    if (tid >= total) return;
    int i = obs_dim > 0 ? tid / obs_dim : tid, j = obs_dim > 0 ? tid % obs_dim : 0;
    const nbiot_record_t *r = rec + i;
    if (reward && j == 0) reward[i] = r->reward;
    if (!obs || obs_dim == 0) return;
    int k = 0, base = 0;
    while (k < n_items && j >= base + items[k].length) { base += items[k].length; ++k; }
    if (k >= n_items) return;
    nbiot_obs_item_t it = items[k];
    int e = j - base;
    double v = 0.0;
    if ((it.state_mask >> r->state) & 1u) {
        const unsigned char *p = (const unsigned char *)r + it.offset;
        if (it.kind == 0) v = (double)((const int32_t *)p)[e];
        else if (it.kind == 1) v = ((const double *)p)[e];
        else v = 1.0 - (double)p[e] / 12.0;
    }
    v = v / it.norm;
    obs[(size_t)i * obs_dim + j] = v < 1.0 ? v : 1.0;          /* min(1.0, v/norm): no lower clip */
}

__global__ void nbiot_stats_kernel(const NbHdr *hdr, int n_inst, unsigned long long *out)
{
    __shared__ unsigned long long acc[NBIOT_N_STATS];
    for (int k = threadIdx.x; k < NBIOT_N_STATS; k += blockDim.x) acc[k] = 0ull;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_inst; i += gridDim.x * blockDim.x) {
        const NbHdr *h = hdr + i;
        long long v[NBIOT_N_STATS];
        for (int k = 0; k < NBIOT_N_STATS; ++k) v[k] = 0;
        for (int k = 0; k < 3; ++k) { v[k] = h->ue_arrivals[k]; v[3 + k] = h->ue_departures[k]; v[6 + k] = h->backoffs[k]; }
        v[NBIOT_STAT_CONNECTIONS] = h->ue_connections; v[NBIOT_STAT_ERRORS] = h->error_count; v[NBIOT_STAT_ATTEMPTS] = h->attempts_count;
        v[NBIOT_STAT_TOTAL_BITS] = h->total_bits; v[NBIOT_STAT_SUBFRAMES] = h->subframes_done; v[NBIOT_STAT_EVENTS] = h->events;
        v[NBIOT_STAT_DRAWS] = (long long)h->draws; v[NBIOT_STAT_FAULTED] = (h->fault & NBIOT_FAULT_FATAL_MASK) ? 1 : 0;
        v[NBIOT_STAT_LIVE_UES] = h->ue_count - (h->ue_departures[0] + h->ue_departures[1] + h->ue_departures[2]);
        v[NBIOT_STAT_SUM_DELAY_X1000] = (long long)(h->av_delay * (double)(h->ue_departures[0] + h->ue_departures[1] + h->ue_departures[2]) * 1000.0);
        v[NBIOT_STAT_TRUNCATED] = (h->fault & NBIOT_FAULT_LIST_TRUNC) ? 1 : 0;
        for (int k = 0; k < NBIOT_STAT_MAX_LOOKAHEAD; ++k) if (v[k]) atomicAdd(&acc[k], (unsigned long long)v[k]);
        atomicMax(&acc[NBIOT_STAT_MAX_LOOKAHEAD], (unsigned long long)h->max_lookahead);
        atomicMax(&acc[NBIOT_STAT_MAX_LIVE_UES], (unsigned long long)h->max_live);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NBIOT_N_STATS; k += blockDim.x)
        if (acc[k]) { if (k < NBIOT_STAT_MAX_LOOKAHEAD) atomicAdd(&out[k], acc[k]); else atomicMax(&out[k], acc[k]); }
}

/* statistics histories (perf_monitor.py:61-84): entries recorded per instance */
__global__ void nbiot_stat_count_kernel(const NbOccHist *occ, int n_inst, int32_t *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inst) out[i] = occ[i].stat_n;
}

/* sample ring (conf['traces'] / conf['statistics']): samples recorded per instance */
__global__ void nbiot_trace_count_kernel(const NbOccHist *occ, int n_inst, int32_t *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inst) out[i] = occ[i].trace_n;
}

/* histogram over the whole batch of one history (0 delay, 1 connection, 2 access, 3 acked bits): one warp per instance
 * streams its ring (coalesced int32 reads), bins of `bin_width` with the last bin open-ended; HBM-bound, one pass */
__global__ void nbiot_stat_hist_kernel(const NbOccHist *occ, const int32_t *stat, int cap, int n_inst, int which, int bin_width, int n_bins,
                                       unsigned long long *hist)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n_inst; i += n_warps) {
        int n = occ[i].stat_n; if (n > cap) n = cap;
        const int32_t *v = stat + ((size_t)i * 4u + (size_t)which) * (size_t)cap;
        for (int k = lane; k < n; k += 32) {
            int b = v[k] / bin_width; if (b < 0) b = 0; if (b >= n_bins) b = n_bins - 1;
            atomicAdd(&hist[b], 1ull);
        }
    }
}

/* Dealing the cells of a ONE-WAVE launch to SMs by predicted cost. With as many warp slots as cells (4096 cells on 148 x 28 slots) every cell
 * starts at once and the launch ends with the SM that happened to get the heaviest cells (SM finish times 12 % apart, a quarter of the warp-time
 * idle, profiles/r02_tail.txt). A cell's cost is predictable enough from its own previous duration and the length of its random-access lists
 * (R2 0.48, tools/cost_model2.py): one CTA sorts the cells by that prediction (bitonic sort in shared memory) and deals them to the SMs in snake
 * order, so that every SM's list holds the same predicted work; nbiot_sim_kernel then takes a warp's cell from the list of the SM it runs on.
 * Which warp runs a cell never changes a result. */
#define NB_DEAL_MAX 8192
__global__ void __launch_bounds__(1024)
nbiot_deal_kernel(const NbHdr *hdr, const uint32_t *dur, const uint8_t *active, int n_inst, int n_pad, int n_lists, int stride,
                  int32_t *order, int32_t *len, unsigned int *cnt)
{
    extern __shared__ unsigned long long keys[];          /* (predicted cost << 16) | cell index */
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        unsigned long long k = 0ull;                       /* padding and inactive cells sort last */
        if (i < n_inst && !(active && !active[i])) {
            const NbHdr *h = hdr + i;
            const unsigned long long ra = (unsigned long long)(h->ra_n[0] + h->ra_n[1] + h->ra_n[2]);
            const unsigned long long cost = 1000000ull + 8690ull * ra + (dur ? (365ull * dur[i]) / 1000ull : 0ull);     /* ns; tools/cost_model2.py */
            k = (cost << 16) | (unsigned long long)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1)
        for (int stride2 = size >> 1; stride2 > 0; stride2 >>= 1) {
            for (int t = threadIdx.x; t < n_pad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride2 - 1)), hi = lo + stride2;
                const bool desc = (lo & size) == 0;        /* overall descending */
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    for (int s2 = threadIdx.x; s2 < n_lists; s2 += blockDim.x) { len[s2] = 0; cnt[s2] = 0u; }
    __syncthreads();
    for (int p = threadIdx.x; p < n_pad; p += blockDim.x) {
        const unsigned long long k = keys[p];
        if (!k) continue;
        const int round = p / n_lists, pos = p % n_lists, sm = (round & 1) ? n_lists - 1 - pos : pos;       /* snake: heavy and light alternate */
        order[(size_t)sm * stride + round] = (int32_t)(k & 0xFFFFull);
        atomicMax(&len[sm], round + 1);
    }
}

/* live UEs of the batch (the load proxy of the automatic kernel choice) */
__global__ void nbiot_load_kernel(const NbHdr *hdr, int n_inst, unsigned long long *out)
{
    unsigned long long acc = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_inst; i += gridDim.x * blockDim.x) {
        const NbHdr *h = hdr + i;
        int live = h->ue_count - (h->ue_departures[0] + h->ue_departures[1] + h->ue_departures[2]);
        acc += (unsigned long long)(live > 0 ? live : 0);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

/* ------------------------------------------------------------------ host side */
static NbKernelArgs make_args(nbiot_handle_t *h, int stage_grid = 1)
{
    NbKernelArgs A;
    A.state = h->state; A.L = h->L; A.conf = h->conf_dev; A.ctrl = h->ctrl_dev; A.lut2 = h->lut2_dev; A.mcs = h->mcs_dev;
    A.load_sum = nullptr; A.load_max = 0ull; A.load_side = 0;
    A.counter = h->counter_dev; A.prof = h->prof_dev; A.evlog = h->evlog; A.evlog_cap = h->evlog_cap;
    A.hdr_bytes16 = (int)(sizeof(NbHdr) / 16);
    A.grid_bytes16 = (int)((size_t)h->L.n_carriers * h->L.grid_w * 2 / 16);
    A.stage_grid = stage_grid; A.blk = h->blk_dev;
    A.deal_order = nullptr; A.deal_len = nullptr; A.deal_cnt = nullptr; A.deal_lists = 0; A.deal_stride = 0; A.dur = nullptr;
    A.smem_per_warp = (int)sizeof(NbCtx) + A.hdr_bytes16 * 16 + (stage_grid ? A.grid_bytes16 * 16 : 0);
    return A;
}

static const NbSimVariant *sim_variant(const nbiot_handle_t *h, int staged)
{
    int c1 = h->use_c1 && !h->evlog;
    if (staged && h->use_staged_builds) return c1 ? &h->k_c1s : &h->k_gs;
    return c1 ? &h->k_c1 : &h->k_gen;
}

static int launch_sim(nbiot_handle_t *h, int mode, const uint8_t *ext_actions, int until_time, int max_steps, cudaStream_t st, const uint8_t *active = nullptr)
{
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->counter_dev, 0, sizeof(unsigned int), st));
    /* Short launches -- an external agent that acts at Scheduling / RAR decisions gets control back after one or two NodeB steps
     * (configs 3 and 4) -- touch a few dozen columns of the look-ahead ring: it stays in HBM (read through L1), which saves staging
     * 2 x C x W x 2 bytes per instance per launch and leaves room for more CTAs per SM. Long launches stage it. */
    const unsigned frequent = (1u << NB_ST_SCHEDULING) | (1u << NB_ST_RAR_WINDOW) | (1u << NB_ST_RAR_WINDOW_END);
    int stage = h->stage_policy >= 0 ? h->stage_policy : !(mode == 2 && ext_actions && (h->ctrl.ext_state_mask & frequent));
    NbKernelArgs A = make_args(h, stage);
    int blocks = stage ? h->blocks : h->blocks_short, smem = stage ? h->smem_bytes : h->smem_short, hi = stage ? h->sim_hi : h->sim_hi_short;
    int use_tpc = h->use_tpc;
    const int both = mode == 2 && h->auto_kernel && stage && !h->evlog;      /* automatic choice: both kernels are launched, the load decides on the device */
    if (both) {
        if (h->load_pending && cudaEventQuery(h->load_ev) == cudaSuccess) {   /* for nbiot_kernel_name only: what an earlier long launch saw */
            h->last_long_tpc = (double)*h->load_pin > h->max_load * (double)h->n_inst ? 0 : h->use_tpc;
            h->load_pending = 0;
        }
        (void)cudaGetLastError();
        CK(cudaMemsetAsync(h->load_dev, 0, sizeof(unsigned long long), st));
        nbiot_load_kernel<<<296, 256, 0, st>>>((const NbHdr *)(h->state + h->L.off_hdr), h->n_inst, h->load_dev);
        CK(cudaGetLastError());
        h->launches += 1;
        if (!h->load_pending) {
            CK(cudaMemcpyAsync(h->load_pin, h->load_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(h->load_ev, st));
            h->load_pending = 1;
        }
    }
    if (mode == 2 && use_tpc && !h->evlog && (stage || h->tpc_forced)) {
        /* one thread per cell: header in local memory, ring / lists / UE records in place in HBM */
        NbKernelArgs T = make_args(h, 0);
        if (both) { T.load_sum = h->load_dev; T.load_max = (unsigned long long)(h->max_load * (double)h->n_inst); T.load_side = 0; }
        const NbSimVariant *v = h->use_c1 ? &h->k_tpc_c1 : &h->k_tpc_gen;
        int threads = v->warps_per_block * 32;
        void *args[] = { (void *)&T, (void *)&ext_actions, (void *)&active, (void *)&until_time, (void *)&max_steps };
        CK(cudaLaunchKernel(use_tpc == 2 ? v->sim_hi : v->sim_lo, dim3((h->n_inst + threads - 1) / threads), dim3(threads), args, 0, st));
    }
    if (mode == 2 && both) { CK(cudaGetLastError()); h->launches += 1; }
    if (mode == 2 && (both || !(use_tpc && !h->evlog && (stage || h->tpc_forced)))) {
        if (both) { A.load_sum = h->load_dev; A.load_max = (unsigned long long)(h->max_load * (double)h->n_inst); A.load_side = 1; }
        const NbSimVariant *v = sim_variant(h, stage);
        if (stage && h->deal && blocks * NB_WPB >= h->n_inst) {     /* a long launch in one wave: deal the cells to the SMs by predicted cost */
            nbiot_deal_kernel<<<1, 1024, (size_t)h->deal_pad * sizeof(unsigned long long), st>>>((const NbHdr *)(h->state + h->L.off_hdr), h->dur_dev, active, h->n_inst,
                                                                                                 h->deal_pad, h->deal_lists, h->deal_stride, h->deal_order, h->deal_len, h->deal_cnt);
            CK(cudaGetLastError());
            h->launches += 1;
            A.deal_order = h->deal_order; A.deal_len = h->deal_len; A.deal_cnt = h->deal_cnt; A.deal_lists = h->deal_lists; A.deal_stride = h->deal_stride;
        }
        if (stage) A.dur = h->dur_dev;
        void *args[] = { (void *)&A, (void *)&ext_actions, (void *)&active, (void *)&until_time, (void *)&max_steps };
        CK(cudaLaunchKernel(hi ? v->sim_hi : v->sim_lo, dim3(blocks), dim3(NB_WPB * 32), args, (size_t)smem, st));
    } else if (mode != 2) {
        int construct = mode == 0 ? 1 : 0;
        const uint64_t *seeds = h->seeds_dev;
        void *args[] = { (void *)&A, (void *)&construct, (void *)&seeds };
        CK(cudaLaunchKernel(sim_variant(h, stage)->reset, dim3(h->blocks), dim3(NB_WPB * 32), args, (size_t)h->smem_bytes, st));
    }
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

extern "C" {

int nbiot_abi_version(void) { return NBIOT_ABI_VERSION; }
const char *nbiot_last_error(void) { return g_err; }

size_t nbiot_state_bytes(const nbiot_conf_t *conf, int n_inst)
{
    if (!conf || n_inst <= 0) return 0;
    return (size_t)nb_make_layout(conf, n_inst).total;
}
size_t nbiot_hdr_bytes(void) { return sizeof(NbHdr); }
size_t nbiot_hdr_task_bytes(void) { return sizeof(NbHdr) - offsetof(NbHdr, tk_steps); }

int nbiot_create(const nbiot_conf_t *conf, int n_inst, int device, void *state_dev, size_t state_bytes,
                 const uint64_t *seeds, const uint16_t *lut2, const uint8_t *mcs, const struct nbiot_ctrl *ctrl,
                 void *stream, nbiot_handle_t **out)
{
    static_assert(sizeof(NbHdr) % 16 == 0, "NbHdr must be a multiple of 16 bytes for the staged copies");
    if (!conf || !state_dev || !seeds || !lut2 || !mcs || !ctrl || !out || n_inst <= 0) return set_err(NBIOT_E_INVALID, "nbiot_create: null or empty argument");
    if (conf->ue_cap < 1 || conf->ue_cap > 65535) return set_err(NBIOT_E_INVALID, "nbiot_create: ue_cap must be in 1..65535");
    if (conf->grid_w < 256 || conf->grid_w > 32 * NB_BLK_MAX || (conf->grid_w & (conf->grid_w - 1))) return set_err(NBIOT_E_INVALID, "nbiot_create: grid_w must be a power of two in 256..4096");
    if (conf->hist_cap < 1 || conf->n_carriers < 1 || conf->n_carriers > 2) return set_err(NBIOT_E_INVALID, "nbiot_create: bad hist_cap / n_carriers");
    if (conf->M < 1 || !(conf->ratio >= 0.0)) return set_err(NBIOT_E_INVALID, "nbiot_create: M must be >= 1 and ratio >= 0");
    if (conf->buffer_lo < 1 || conf->buffer_lo > 65535 || conf->buffer_hi > 65535) return set_err(NBIOT_E_INVALID, "nbiot_create: buffer range must lie in 1..65535 (a UE's buffer is held in 16 bits)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_err(NBIOT_E_NOGPU, "no CUDA device: this library has no CPU path");
    if (device < 0 || device >= ndev) return set_err(NBIOT_E_INVALID, "nbiot_create: bad device index");
    CK(cudaSetDevice(device));
    NbLayout L = nb_make_layout(conf, n_inst);
    if (state_bytes < L.total) return set_err(NBIOT_E_SIZE, "nbiot_create: state buffer smaller than nbiot_state_bytes()");
    cudaStream_t st = (cudaStream_t)stream;
    nbiot_handle_t *h = new nbiot_handle_t();      /* value-initialised: every scalar member is zero */
    struct Guard { nbiot_handle_t *h; ~Guard() { if (h) nbiot_destroy(h); } } guard{ h };     /* every error return below frees what was allocated */
    h->conf = *conf; h->ctrl = *ctrl; h->L = L; h->device = device; h->n_inst = n_inst; h->state = (uint8_t *)state_dev;
    CK(cudaMalloc(&h->conf_dev, sizeof(nbiot_conf_t)));
    CK(cudaMalloc(&h->ctrl_dev, sizeof(NbCtaConst)));
    CK(cudaMalloc(&h->lut2_dev, NB_LUT_CURVES * NB_LUT_SNR_PTS * sizeof(uint16_t)));
    CK(cudaMalloc(&h->mcs_dev, NB_MCS_LOSS_ROWS * NB_N_TBS * 3));
    CK(cudaMalloc(&h->seeds_dev, sizeof(uint64_t) * (size_t)n_inst));
    CK(cudaMalloc(&h->counter_dev, sizeof(unsigned int) * (1 + NB_SM_HINTS)));     /* [0]: claim counter of a launch; [1 + sm]: class hint of the voted kernel (measurement builds) */
    CK(cudaMemsetAsync(h->counter_dev, 0, sizeof(unsigned int) * (1 + NB_SM_HINTS), st));
    if (getenv("NBIOT_PROF_INST") && atoi(getenv("NBIOT_PROF_INST"))) CK(cudaMalloc(&h->prof_dev, sizeof(unsigned long long) * 4 * (size_t)n_inst));
    CK(cudaMemcpyAsync(h->conf_dev, conf, sizeof(nbiot_conf_t), cudaMemcpyHostToDevice, st));
    static_assert(sizeof(NbCtaConst) % 4 == 0, "NbCtaConst is staged word by word");
    NbCtaConst cd; memset(&cd, 0, sizeof cd);
    cd.ctrl = nb_make_ctrl_dev(ctrl); cd.cold = nb_make_cold(conf, h->state, L.off_occ, L.off_stat, L.off_trace, L.off_sclog);
    CK(cudaMemcpyAsync(h->ctrl_dev, &cd, sizeof(NbCtaConst), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->lut2_dev, lut2, NB_LUT_CURVES * NB_LUT_SNR_PTS * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->mcs_dev, mcs, NB_MCS_LOSS_ROWS * NB_N_TBS * 3, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->seeds_dev, seeds, sizeof(uint64_t) * (size_t)n_inst, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));     /* the host arrays may go away after this call */
    /* launch geometry: persistent grid = SMs x resident CTAs, capped by the work available */
    NbKernelArgs A = make_args(h);
    h->smem_bytes = NB_WPB * A.smem_per_warp;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if ((size_t)h->smem_bytes + sizeof(NbCtaConst) + 64 > prop.sharedMemPerBlockOptin) return set_err(NBIOT_E_INVALID, "grid_w too large for the shared memory of one CTA");
    nb_sim_variant_generic(&h->k_gen); nb_sim_variant_c1(&h->k_c1); nb_sim_variant_gs(&h->k_gs); nb_sim_variant_c1s(&h->k_c1s);
    nb_sim_variant_tpc_gen(&h->k_tpc_gen); nb_sim_variant_tpc_c1(&h->k_tpc_c1);
    if (const char *e = getenv("NBIOT_TPC_CARVEOUT"))      /* measurement knob: shared-memory carve-out of the thread-per-cell kernels, % */
        for (const NbSimVariant *v : { &h->k_tpc_gen, &h->k_tpc_c1 })
            for (const void *f : { v->sim_lo, v->sim_hi }) CK(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
    /* Which kernel advances the batch. One warp per cell keeps a cell's latency low but repeats the scalar control logic on 32
     * lanes; one thread per cell shares every common instruction between 32 cells but needs the cells to fill the machine
     * (148 SMs x 16 warps x 32 lanes = 75 776) -- DESIGN.md section 6c has the measured crossover. */
    h->use_tpc = n_inst >= NB_TPC_MIN_CELLS ? 2 : 0;
    /* dealing by predicted cost for batches that run their long launches in one wave: measured 1.7 % SLOWER on configs[1] (a cell's
     * duration in a step is mostly that step's fresh randomness, and the launch ends with its slowest CELL, not its most loaded SM:
     * profiles/r02_tail.txt), so it is off unless NBIOT_DEAL=1 */
    h->deal = 0;
    if (n_inst <= NB_DEAL_MAX && n_inst >= 2 * prop.multiProcessorCount && getenv("NBIOT_DEAL") && atoi(getenv("NBIOT_DEAL"))) {
        h->deal_lists = prop.multiProcessorCount;
        h->deal_pad = 2; while (h->deal_pad < n_inst) h->deal_pad <<= 1;
        h->deal_stride = (h->deal_pad + h->deal_lists - 1) / h->deal_lists + 1;
        CK(cudaMalloc(&h->dur_dev, sizeof(uint32_t) * (size_t)n_inst));
        CK(cudaMemsetAsync(h->dur_dev, 0, sizeof(uint32_t) * (size_t)n_inst, st));
        CK(cudaMalloc(&h->deal_order, sizeof(int32_t) * (size_t)h->deal_lists * h->deal_stride));
        CK(cudaMalloc(&h->deal_len, sizeof(int32_t) * (size_t)h->deal_lists));
        CK(cudaMalloc(&h->deal_cnt, sizeof(unsigned int) * (size_t)h->deal_lists));
        CK(cudaFuncSetAttribute(nbiot_deal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NB_DEAL_MAX * (int)sizeof(unsigned long long)));
        h->deal = 1;
    }
    /* block summaries of the look-ahead ring for the thread-per-cell kernel (nbiot_core.cuh): 8 % fewer executed instructions and 3-4 %
     * SLOWER on the sweep configuration (the kernel waits on memory, and the summaries are one more scattered region per cell):
     * off unless NBIOT_TPC_BLK=1 (profiles/r02_tpc_variants.txt) */
    if (getenv("NBIOT_TPC_BLK") && atoi(getenv("NBIOT_TPC_BLK"))) CK(cudaMalloc(&h->blk_dev, sizeof(uint16_t) * (size_t)n_inst * L.n_carriers * (L.grid_w >> 5)));
    if (conf->kernel < NBIOT_KERNEL_AUTO || conf->kernel > NBIOT_KERNEL_THREAD_VOTED) return set_err(NBIOT_E_INVALID, "nbiot_create: conf.kernel must be one of NBIOT_KERNEL_*");
    if (conf->kernel == NBIOT_KERNEL_WARP) h->use_tpc = 0;
    else if (conf->kernel == NBIOT_KERNEL_THREAD) h->use_tpc = 1;
    else if (conf->kernel == NBIOT_KERNEL_THREAD_VOTED) h->use_tpc = 2;
    h->auto_kernel = conf->kernel == NBIOT_KERNEL_AUTO && h->use_tpc && !getenv("NBIOT_KERNEL");
    h->last_long_tpc = h->use_tpc;
    h->max_load = getenv("NBIOT_TPC_MAX_LOAD") ? atof(getenv("NBIOT_TPC_MAX_LOAD")) : NB_TPC_MAX_LOAD;      /* measurement knob */
    if (h->auto_kernel) {
        CK(cudaMalloc(&h->load_dev, sizeof(unsigned long long)));
        CK(cudaMallocHost(&h->load_pin, sizeof(unsigned long long)));
        CK(cudaEventCreateWithFlags(&h->load_ev, cudaEventDisableTiming));
    }
    if (const char *e = getenv("NBIOT_KERNEL")) {              /* measurement override: tpc / tpcv also take the short launches */
        if (!strcmp(e, "tpc")) { h->use_tpc = 1; h->tpc_forced = 1; }
        else if (!strcmp(e, "tpcv")) { h->use_tpc = 2; h->tpc_forced = 1; }
        else if (!strcmp(e, "warp")) h->use_tpc = 0;
    }
    h->use_staged_builds = getenv("NBIOT_SIM_STAGED_BUILDS") ? atoi(getenv("NBIOT_SIM_STAGED_BUILDS")) : 1;
    h->use_c1 = L.n_carriers == 1;
    if (const char *e = getenv("NBIOT_SIM_VARIANT")) h->use_c1 = h->use_c1 && strcmp(e, "generic") != 0;   /* measurement knob */
    /* the attribute belongs to the function, not to the handle: always the device's maximum, so that a later, smaller batch cannot
     * lower the limit under an earlier one */
    for (const NbSimVariant *v : { &h->k_gen, &h->k_c1, &h->k_gs, &h->k_c1s })
        for (const void *f : { v->sim_lo, v->sim_hi, v->reset }) {
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, f));                 /* the opt-in maximum covers static + dynamic shared memory */
            CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
        }
    const NbSimVariant *kv = h->use_c1 ? &h->k_c1 : &h->k_gen;       /* both builds have the same launch bounds and register budgets */
    int per_sm = 0;
    int per_sm_hi = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kv->sim_lo, NB_WPB * 32, h->smem_bytes));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_hi, kv->sim_hi, NB_WPB * 32, h->smem_bytes));
    /* the leaner-register build only when it really puts more CTAs on an SM AND the batch has the instances to fill them */
    h->sim_hi = per_sm_hi > per_sm && (long long)n_inst > (long long)prop.multiProcessorCount * per_sm * NB_WPB;
    if (const char *e = getenv("NBIOT_SIM_BUILD")) h->sim_hi = !strcmp(e, "hi");      /* measurement knob: "lo" / "hi" */
    if (h->sim_hi) per_sm = per_sm_hi;
    if (per_sm < 1) per_sm = 1;
    if (const char *e = getenv("NBIOT_MAX_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   /* measurement knob */
    {   /* the same choice for launches that do not stage the grid */
        NbKernelArgs B = make_args(h, 0);
        h->smem_short = NB_WPB * B.smem_per_warp;
        int lo = 0, hi2 = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lo, kv->sim_lo, NB_WPB * 32, h->smem_short));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&hi2, kv->sim_hi, NB_WPB * 32, h->smem_short));
        h->sim_hi_short = hi2 > lo && (long long)n_inst > (long long)prop.multiProcessorCount * lo * NB_WPB;
        if (const char *e = getenv("NBIOT_SIM_BUILD")) h->sim_hi_short = !strcmp(e, "hi");
        int ps = h->sim_hi_short ? hi2 : lo; if (ps < 1) ps = 1;
        int need_s = (n_inst + NB_WPB - 1) / NB_WPB, cap_s = prop.multiProcessorCount * ps;
        h->blocks_short = need_s < cap_s ? need_s : cap_s;
        h->stage_policy = -1;
        if (const char *e = getenv("NBIOT_STAGE_GRID")) h->stage_policy = atoi(e) ? 1 : 0;       /* measurement knob */
    }
    int need = (n_inst + NB_WPB - 1) / NB_WPB;
    int cap = prop.multiProcessorCount * per_sm;
    h->blocks = need < cap ? need : cap;
    int rc = launch_sim(h, 0, nullptr, -1, 0, st);     /* construct + reset every instance */
    if (rc) return rc;
    guard.h = nullptr;
    *out = h;
    return 0;
}

int nbiot_destroy(nbiot_handle_t *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->load_ev) cudaEventDestroy(h->load_ev);
    cudaFree(h->load_dev); if (h->load_pin) cudaFreeHost(h->load_pin);
    cudaFree(h->conf_dev); cudaFree(h->ctrl_dev); cudaFree(h->lut2_dev); cudaFree(h->mcs_dev); cudaFree(h->seeds_dev); cudaFree(h->counter_dev); cudaFree(h->prof_dev); cudaFree(h->blk_dev); cudaFree(h->dur_dev); cudaFree(h->deal_order); cudaFree(h->deal_len); cudaFree(h->deal_cnt);
    cudaFree(h->act_dev); cudaFree(h->obs_dev); cudaFree(h->rew_dev); cudaFree(h->items_dev);
    cudaFreeHost(h->act_pin); cudaFreeHost(h->obs_pin); cudaFreeHost(h->rew_pin);
    delete h;
    return 0;
}

int nbiot_set_controller(nbiot_handle_t *h, const struct nbiot_ctrl *ctrl, void *stream)
{
    if (!h || !ctrl) return set_err(NBIOT_E_INVALID, "nbiot_set_controller: null argument");
    CK(cudaSetDevice(h->device));
    h->ctrl = *ctrl;
    NbCtrlDev cd = nb_make_ctrl_dev(&h->ctrl);
    CK(cudaMemcpyAsync(&h->ctrl_dev->ctrl, &cd, sizeof(NbCtrlDev), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int nbiot_reset(nbiot_handle_t *h, void *stream)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_reset: null handle");
    return launch_sim(h, 1, nullptr, -1, 0, (cudaStream_t)stream);
}

int nbiot_advance(nbiot_handle_t *h, const uint8_t *ext_actions_dev, int until_time, int max_steps, void *stream)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_advance: null handle");
    if (max_steps < 1) return set_err(NBIOT_E_INVALID, "nbiot_advance: max_steps must be >= 1");
    return launch_sim(h, 2, ext_actions_dev, until_time, max_steps, (cudaStream_t)stream);
}

int nbiot_advance_masked(nbiot_handle_t *h, const uint8_t *ext_actions_dev, const uint8_t *active_dev, int until_time, int max_steps, void *stream)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_advance_masked: null handle");
    if (max_steps < 1) return set_err(NBIOT_E_INVALID, "nbiot_advance_masked: max_steps must be >= 1");
    return launch_sim(h, 2, ext_actions_dev, until_time, max_steps, (cudaStream_t)stream, active_dev);
}

int nbiot_records_ptr(nbiot_handle_t *h, void **records_dev)
{
    if (!h || !records_dev) return set_err(NBIOT_E_INVALID, "nbiot_records_ptr: null argument");
    *records_dev = h->state + h->L.off_rec;
    return 0;
}

int nbiot_hist_ptrs(nbiot_handle_t *h, void **incoming_dev, void **delays_dev, void **service_dev, int *hist_cap)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_hist_ptrs: null handle");
    if (incoming_dev) *incoming_dev = h->state + h->L.off_incoming;
    if (delays_dev) *delays_dev = h->state + h->L.off_delays;
    if (service_dev) *service_dev = h->state + h->L.off_service;
    if (hist_cap) *hist_cap = h->L.hist_cap;
    return 0;
}

static int ensure_items(nbiot_handle_t *h, const nbiot_obs_item_t *items, int n_items, cudaStream_t st, int *obs_dim)
{
    int d = 0;
    for (int k = 0; k < n_items; ++k) d += items[k].length;
    *obs_dim = d;
    if (n_items <= 0) return 0;
    if ((size_t)n_items > h->items_cap) {
        cudaFree(h->items_dev); h->items_dev = nullptr; h->items_cap = 0; h->items_host.clear();
        CK(cudaMalloc(&h->items_dev, sizeof(nbiot_obs_item_t) * (size_t)n_items));
        h->items_cap = (size_t)n_items;
    }
    /* the table only changes when the external agent does: upload it when it differs from what the device already holds */
    if (h->items_host.size() != (size_t)n_items || memcmp(h->items_host.data(), items, sizeof(nbiot_obs_item_t) * (size_t)n_items) != 0) {
        h->items_host.assign(items, items + n_items);
        CK(cudaMemcpyAsync(h->items_dev, h->items_host.data(), sizeof(nbiot_obs_item_t) * (size_t)n_items, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

int nbiot_gather_obs(nbiot_handle_t *h, const nbiot_obs_item_t *items, int n_items, double *obs_dev, double *reward_dev, void *stream)
{
    if (!h || (n_items > 0 && !items)) return set_err(NBIOT_E_INVALID, "nbiot_gather_obs: null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int obs_dim = 0;
    int rc = ensure_items(h, items, n_items, st, &obs_dim);
    if (rc) return rc;
    int total = h->n_inst * (obs_dim > 0 ? obs_dim : 1);
    nbiot_obs_kernel<<<(total + 255) / 256, 256, 0, st>>>((const nbiot_record_t *)(h->state + h->L.off_rec), h->n_inst, h->items_dev, n_items, obs_dim,
                                                       obs_dim > 0 ? obs_dev : nullptr, reward_dev);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int nbiot_reduce_stats(nbiot_handle_t *h, int64_t *out_dev, void *stream)
{
    if (!h || !out_dev) return set_err(NBIOT_E_INVALID, "nbiot_reduce_stats: null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(out_dev, 0, sizeof(int64_t) * NBIOT_N_STATS, st));
    int blocks = (h->n_inst + 127) / 128; if (blocks > 592) blocks = 592;
    nbiot_stats_kernel<<<blocks, 128, 0, st>>>((const NbHdr *)(h->state + h->L.off_hdr), h->n_inst, (unsigned long long *)out_dev);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int nbiot_step_host(nbiot_handle_t *h, const uint8_t *actions_host, int until_time, int max_steps,
                    const nbiot_obs_item_t *items, int n_items, double *obs_host, double *reward_host, void *stream)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_step_host: null handle");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int k = h->ctrl.ext_k;
    size_t act_bytes = actions_host ? (size_t)h->n_inst * (size_t)k : 0;
    if (act_bytes > h->act_cap) {
        cudaFree(h->act_dev); cudaFreeHost(h->act_pin); h->act_dev = nullptr; h->act_pin = nullptr; h->act_cap = 0;
        CK(cudaMalloc(&h->act_dev, act_bytes)); CK(cudaMallocHost(&h->act_pin, act_bytes)); h->act_cap = act_bytes;
    }
    int obs_dim = 0;
    int rc = ensure_items(h, items, n_items, st, &obs_dim);
    if (rc) return rc;
    size_t obs_elems = (size_t)h->n_inst * (size_t)(obs_dim > 0 ? obs_dim : 1);
    if (obs_elems > h->obs_cap) {
        cudaFree(h->obs_dev); cudaFree(h->rew_dev); cudaFreeHost(h->obs_pin); cudaFreeHost(h->rew_pin);
        h->obs_dev = h->rew_dev = nullptr; h->obs_pin = h->rew_pin = nullptr; h->obs_cap = 0;
        CK(cudaMalloc(&h->obs_dev, obs_elems * sizeof(double))); CK(cudaMalloc(&h->rew_dev, (size_t)h->n_inst * sizeof(double)));
        CK(cudaMallocHost(&h->obs_pin, obs_elems * sizeof(double))); CK(cudaMallocHost(&h->rew_pin, (size_t)h->n_inst * sizeof(double)));
        h->obs_cap = obs_elems;
    }
    if (actions_host) {
        memcpy(h->act_pin, actions_host, act_bytes);
        CK(cudaMemcpyAsync(h->act_dev, h->act_pin, act_bytes, cudaMemcpyHostToDevice, st));
    }
    rc = launch_sim(h, 2, actions_host ? h->act_dev : nullptr, until_time, max_steps, st);
    if (rc) return rc;
    int total = h->n_inst * (obs_dim > 0 ? obs_dim : 1);
    nbiot_obs_kernel<<<(total + 255) / 256, 256, 0, st>>>((const nbiot_record_t *)(h->state + h->L.off_rec), h->n_inst, h->items_dev, n_items, obs_dim,
                                                       obs_dim > 0 ? h->obs_dev : nullptr, h->rew_dev);
    CK(cudaGetLastError());
    h->launches += 1;
    if (obs_host && obs_dim > 0) CK(cudaMemcpyAsync(h->obs_pin, h->obs_dev, (size_t)h->n_inst * obs_dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (reward_host) CK(cudaMemcpyAsync(h->rew_pin, h->rew_dev, (size_t)h->n_inst * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (obs_host && obs_dim > 0) memcpy(obs_host, h->obs_pin, (size_t)h->n_inst * obs_dim * sizeof(double));
    if (reward_host) memcpy(reward_host, h->rew_pin, (size_t)h->n_inst * sizeof(double));
    return 0;
}

int nbiot_stat_ptrs(nbiot_handle_t *h, void **stat_dev, int *stat_cap)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_stat_ptrs: null handle");
    if (stat_dev) *stat_dev = h->L.stat_cap > 0 ? (void *)(h->state + h->L.off_stat) : nullptr;
    if (stat_cap) *stat_cap = h->L.stat_cap;
    return 0;
}

int nbiot_spill_ptr(nbiot_handle_t *h, void **spill_dev, int *list_cap, int *spill_cap)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_spill_ptr: null handle");
    if (spill_dev) *spill_dev = h->state + h->L.off_spill;
    if (list_cap) *list_cap = NBIOT_STEP_LIST;
    if (spill_cap) *spill_cap = nb_spill_cap(h->L.hist_cap);
    return 0;
}

int nbiot_trace_ptrs(nbiot_handle_t *h, void **trace_dev, int *trace_cap)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_trace_ptrs: null handle");
    if (trace_dev) *trace_dev = h->L.trace_cap > 0 ? (void *)(h->state + h->L.off_trace) : nullptr;
    if (trace_cap) *trace_cap = h->L.trace_cap;
    return 0;
}

int nbiot_trace_counts(nbiot_handle_t *h, int32_t *counts_dev, void *stream)
{
    if (!h || !counts_dev) return set_err(NBIOT_E_INVALID, "nbiot_trace_counts: null argument");
    CK(cudaSetDevice(h->device));
    nbiot_trace_count_kernel<<<(h->n_inst + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const NbOccHist *)(h->state + h->L.off_occ), h->n_inst, counts_dev);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int nbiot_stat_counts(nbiot_handle_t *h, int32_t *counts_dev, void *stream)
{
    if (!h || !counts_dev) return set_err(NBIOT_E_INVALID, "nbiot_stat_counts: null argument");
    CK(cudaSetDevice(h->device));
    nbiot_stat_count_kernel<<<(h->n_inst + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const NbOccHist *)(h->state + h->L.off_occ), h->n_inst, counts_dev);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int nbiot_stat_histogram(nbiot_handle_t *h, int which, int bin_width, int n_bins, int64_t *hist_dev, void *stream)
{
    if (!h || !hist_dev || which < 0 || which > 3 || bin_width < 1 || n_bins < 1) return set_err(NBIOT_E_INVALID, "nbiot_stat_histogram: bad argument");
    if (h->L.stat_cap <= 0) return set_err(NBIOT_E_INVALID, "nbiot_stat_histogram: the batch was created without statistics (stat_cap = 0)");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(hist_dev, 0, sizeof(int64_t) * (size_t)n_bins, st));
    int blocks = (h->n_inst + 3) / 4; if (blocks > 148 * 8) blocks = 148 * 8;
    nbiot_stat_hist_kernel<<<blocks, 128, 0, st>>>((const NbOccHist *)(h->state + h->L.off_occ), (const int32_t *)(h->state + h->L.off_stat), h->L.stat_cap,
                                                    h->n_inst, which, bin_width, n_bins, (unsigned long long *)hist_dev);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int nbiot_set_event_log(nbiot_handle_t *h, int32_t *evlog_dev, int cap)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_set_event_log: null handle");
    h->evlog = evlog_dev; h->evlog_cap = evlog_dev ? cap : 0;
    return 0;
}

int nbiot_instance_times(nbiot_handle_t *h, unsigned long long *out_host)
{
    if (!h || !out_host) return set_err(NBIOT_E_INVALID, "nbiot_instance_times: null argument");
    if (!h->prof_dev) return set_err(NBIOT_E_INVALID, "nbiot_instance_times: create the batch with NBIOT_PROF_INST=1 in the environment");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_host, h->prof_dev, sizeof(unsigned long long) * 4 * (size_t)h->n_inst, cudaMemcpyDeviceToHost));
    return 0;
}

long long nbiot_launch_count(nbiot_handle_t *h) { return h ? h->launches : 0; }

int nbiot_launch_geometry(nbiot_handle_t *h, int *blocks, int *warps_per_block, int *smem_bytes)
{
    if (!h) return set_err(NBIOT_E_INVALID, "nbiot_launch_geometry: null handle");
    if (blocks) *blocks = h->blocks;
    if (warps_per_block) *warps_per_block = NB_WPB;
    if (smem_bytes) *smem_bytes = h->smem_bytes;
    return 0;
}

const char *nbiot_kernel_name(nbiot_handle_t *h)
{
    if (!h) return "";
    int t = h->auto_kernel ? h->last_long_tpc : h->use_tpc;      /* automatic choice: what the last long launch ran on */
    return t == 2 ? "thread-per-cell, voted dispatch (long launches)" : t ? "thread-per-cell (long launches)" : "warp-per-cell";
}

double nbiot_rng_u01(uint64_t seed, uint64_t ctr, uint32_t stream) { return nb_rng_u01(seed, ctr, stream); }
void nbiot_rng_pair(uint64_t seed, uint64_t ctr, uint32_t stream, double *xy) { nb_rng_pair(seed, ctr, stream, &xy[0], &xy[1]); }
double nbiot_rng_normal(uint64_t seed, uint64_t ctr, uint32_t stream) { return nb_rng_normal(seed, ctr, stream); }
uint64_t nbiot_rng_bounded(uint64_t seed, uint64_t ctr, uint32_t stream, uint64_t n) { return nb_rng_bounded(seed, ctr, stream, n); }

}   /* extern "C" */

#include "nbiot_agent_impl.cuh"
