/*
 * nbiot_queue.cu -- the migrating-instance kernel: Controller.step / run_until for every instance of the batch
 * (the same contract as nbiot_sim_kernel in nbiot_kernels.cu) with the instances MOVING between SMs.
 *
 * Why. The warp-per-instance kernel is bound by instruction supply (profiles/r01_icache.txt): the event cycle of one
 * cell walks through ~22 KB of code per event and ~60 KB per cycle of event types, 28 warps of an SM are in 28
 * unrelated places of it, the 32 KB SM instruction cache misses 42 % of its requests and the GPC-level cache behind
 * it saturates. Nothing in the cells couples them, so nothing forces a cell to stay on one SM either:
 *
 *   * the control flow of a cell is cut into TASKS (nbiot_core.cuh "tasks"): one heavy handler + the glue behind it;
 *     all control state between two tasks lives in the 3.5 KB header (NbHdr.tk_*), the rest of a cell is in HBM anyway;
 *   * there is one global queue per task class (9 classes). A persistent CTA per SM works on ONE class at a time:
 *     its dispatcher warp claims tasks of that class for the idle worker warps, a worker loads the header into its
 *     shared-memory slice (3.5 KB, coalesced 16-byte copies, through L2), runs the task, stores the header and
 *     pushes the cell onto the queue of the class its glue named -- unless that is the class this SM is working on,
 *     in which case the worker just carries on with the same cell (no copy, no queue);
 *   * an SM therefore executes one handler + glue (10-25 KB of SASS) with all its warps, which fits the SM
 *     instruction cache; an SM changes class only when its class runs dry while another one has work.
 *
 * The look-ahead ring stays in HBM and is read in place (the mode the warp-per-instance kernel uses for short
 * launches; tests hold both to bit-identical states). Queue protocol: push = reserve a slot index with atomicAdd on
 * `tail`, release-store the cell index into the slot, atomicAdd `avail`; pop = the dispatcher takes up to n_idle units
 * from `avail` (atomicSub, excess handed back), reserves that many slot indices from `head` and hands (class, index)
 * to idle workers through shared-memory mailboxes; the worker acquire-loads its slot (spinning for the few hundred
 * nanoseconds a producer may still be between its two atomics), which orders its loads of the cell after the
 * producer's stores. Every atomic is an add: nothing retries, so the queues scale with the number of SMs.
 * Termination: `remaining` counts the cells that have not finished their launch; at zero every dispatcher retires
 * its workers.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NB_TASKS 1
#include "nbiot.h"
#include "nbiot_core.cuh"
#include "nbiot_launch.h"

#ifndef NB_Q_SPIN_LIMIT
#define NB_Q_SPIN_LIMIT (1u << 25)     /* polls (~50 ns each) a worker or dispatcher may wait with nothing happening: a scheduler bug, not a hang */
#endif

struct NbGqClass {                     /* one 128-byte line per counter: producers, dispatchers and pollers do not share lines */
    unsigned int tail; unsigned int pad0[31];
    unsigned int head; unsigned int pad1[31];
    int avail; int pad2[31];
};
struct NbGq {
    NbGqClass c[NB_TK_N];
    int remaining; int bail; int pad[30];
    unsigned long long tasks[NB_TK_N], cycles[NB_TK_N], hops, local, switches, idle_polls;
};

struct NbQueueRt {
    int device, n_inst, workers, blocks, smem_bytes, smem_per_warp;
    unsigned int cap_mask;
    NbGq *gq;
    uint32_t *slots;                   /* [NB_TK_N][cap]: cell index + 1, 0 = empty */
    int hop_min, idle_before_hop, stats, batch_min;
};

/* ------------------------------------------------------------------ device side */
__device__ __forceinline__ unsigned int nbq_ld_acquire(const unsigned int *p)
{ unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void nbq_st_release(unsigned int *p, unsigned int v)
{ asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int nbq_ld_volatile(const int *p) { return *(const volatile int *)p; }

__global__ void nbq_init_kernel(NbGq *gq, uint32_t *slots, unsigned int cap, int n_inst, const uint8_t *active)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst || (active && !active[i])) return;
    unsigned int t = atomicAdd(&gq->c[NB_TK_BEGIN].tail, 1u);
    slots[(size_t)NB_TK_BEGIN * cap + t] = (uint32_t)i + 1u;
    atomicAdd(&gq->c[NB_TK_BEGIN].avail, 1);
    atomicAdd(&gq->remaining, 1);
}

/* mailbox of a worker: 0 = idle, NBQ_EXIT = retire, else ((class + 1) << 32) | slot index */
#define NBQ_EXIT 0xFFFFFFFFFFFFFFFFull

template <int QW>
__global__ void __launch_bounds__((QW + 1) * 32, 1)
nbiot_simg_kernel(NbKernelArgs A, const uint8_t *ext_actions, int until_time, int max_steps, NbGq *gq, uint32_t *slots, unsigned int cap_mask,
                  int hop_min, int idle_before_hop, int stats, int batch_min)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ NbCtaConst s_ctrl;
    __shared__ unsigned long long s_mbox[32];
    __shared__ int s_cur;
    for (int k = threadIdx.x; k < (int)(sizeof(NbCtaConst) / 4); k += blockDim.x) ((uint32_t *)&s_ctrl)[k] = ((const uint32_t *)A.ctrl)[k];
    if (threadIdx.x < 32) s_mbox[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) s_cur = NB_TK_BEGIN;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t cap = (size_t)cap_mask + 1u;

    if (warp == QW) {
        /* ---------------- dispatcher: keeps the SM on one class, feeds idle workers ---------------- */
        int cur = NB_TK_BEGIN, idle_rounds = 0, batch_wait = 0;
        unsigned int spins = 0;
        unsigned long long n_switch = 0, n_idle = 0;
        for (;;) {
            unsigned long long mb = lane < QW ? *(volatile unsigned long long *)&s_mbox[lane] : 1ull;
            uint32_t idle_mask = __ballot_sync(0xFFFFFFFFu, mb == 0ull);
            int n_free = __popc(idle_mask);
            /* batch_min > 1: hand tasks out only to batch_min idle workers at once, so that they enter the handler together
             * and fetch the same instruction lines at the same time (measurement knob NBIOT_Q_BATCH) */
            if (n_free < batch_min && batch_wait < 4000) { batch_wait += 1; __nanosleep(40); continue; }
            batch_wait = 0;
            /* lane k looks at class k */
            int a = lane < NB_TK_N ? nbq_ld_volatile(&gq->c[lane].avail) : 0;
            int a_cur = __shfl_sync(0xFFFFFFFFu, a, cur);
            if (a_cur <= 0) {
                uint32_t score = a > 0 ? (((uint32_t)(a < 0xFFFFFF ? a : 0xFFFFFF) << 8) | (uint32_t)(31 - lane)) : 0u;
                uint32_t best = __reduce_max_sync(0xFFFFFFFFu, score);
                int rem = 0, bail = 0;
                if (lane == 0) { rem = nbq_ld_volatile(&gq->remaining); bail = nbq_ld_volatile(&gq->bail); }
                rem = __shfl_sync(0xFFFFFFFFu, rem, 0); bail = __shfl_sync(0xFFFFFFFFu, bail, 0);
                if (rem <= 0 || bail) break;
                int need = idle_rounds >= idle_before_hop ? 1 : hop_min;
                if ((int)(best >> 8) >= need) {
                    cur = 31 - (int)(best & 0xFFu);
                    if (lane == 0) *(volatile int *)&s_cur = cur;
                    idle_rounds = 0; n_switch += 1;
                    a_cur = __shfl_sync(0xFFFFFFFFu, a, cur);
                } else {
                    idle_rounds += 1; n_idle += 1;
                    __nanosleep(100);
                    if (++spins > NB_Q_SPIN_LIMIT) { if (lane == 0) atomicExch(&gq->bail, 1); break; }
                    continue;
                }
            }
            spins = 0; idle_rounds = 0;
            int want = a_cur < n_free ? a_cur : n_free, got = 0;
            unsigned int h0 = 0;
            if (lane == 0) {
                int old = atomicSub(&gq->c[cur].avail, want);
                got = old < 0 ? 0 : (old < want ? old : want);
                if (got < want) atomicAdd(&gq->c[cur].avail, want - got);
                if (got) h0 = atomicAdd(&gq->c[cur].head, (unsigned int)got);
            }
            got = __shfl_sync(0xFFFFFFFFu, got, 0); h0 = __shfl_sync(0xFFFFFFFFu, h0, 0);
            if (lane < got) {
                int w = __fns(idle_mask, 0, lane + 1);           /* the (lane+1)-th idle worker */
                *(volatile unsigned long long *)&s_mbox[w] = ((unsigned long long)(cur + 1) << 32) | (unsigned long long)(h0 + (unsigned int)lane);
            }
            __syncwarp();
        }
        /* retire the workers (all idle: `remaining` reached zero after the last task stored its cell) */
        if (lane < QW) *(volatile unsigned long long *)&s_mbox[lane] = NBQ_EXIT;
        if (lane == 0) { atomicAdd(&gq->switches, n_switch); atomicAdd(&gq->idle_polls, n_idle); }
        return;
    }

    /* ---------------- worker ---------------- */
    unsigned char *my = smem + (size_t)warp * A.smem_per_warp;
    NbCtx &c = *(NbCtx *)my;
    unsigned char *hdr_s = my + sizeof(NbCtx);
    NbHdr *hdr_g = (NbHdr *)(A.state + A.L.off_hdr);
    if (lane == 0) {
        c.lut2 = A.lut2; c.mcs = A.mcs; c.conf = *(const nbiot_conf_hot_t *)A.conf; c.ctrl = &s_ctrl.ctrl;
        c.W = A.L.grid_w; c.ue_cap = A.L.ue_cap; c.ra_cap = A.L.ra_cap; c.hist_cap = A.L.hist_cap; c.C = A.L.n_carriers;
        c.evlog = nullptr; c.evlog_cap = A.evlog_cap; c.tk_until = until_time; c.tk_max_steps = max_steps; c.tk_internal = ext_actions ? 0 : 1;
    }
    __syncwarp();
    unsigned long long st_hops = 0, st_local = 0;
    unsigned int spins = 0;
    for (;;) {
        unsigned long long mb = 0ull;
        if (lane == 0) mb = *(volatile unsigned long long *)&s_mbox[warp];
        mb = __shfl_sync(0xFFFFFFFFu, mb, 0);
        if (mb == 0ull) { __nanosleep(40); continue; }
        if (mb == NBQ_EXIT) break;
        int cls = (int)(mb >> 32) - 1;
        unsigned int ticket = (unsigned int)mb;
        unsigned int *sp = &slots[(size_t)cls * cap + (ticket & cap_mask)];
        unsigned int v = 0;
        for (;;) {                                              /* the producer may still be between `tail` and the slot store */
            if (lane == 0) v = nbq_ld_acquire(sp);
            v = __shfl_sync(0xFFFFFFFFu, v, 0);
            if (v) break;
            if (++spins > NB_Q_SPIN_LIMIT) { if (lane == 0) atomicExch(&gq->bail, 1); break; }
        }
        if (!v) break;
        spins = 0;
        if (lane == 0) *(volatile unsigned int *)sp = 0u;
        const int i = (int)v - 1;
        if (lane == 0) {
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
            c.grid = (uint16_t *)(s + L.off_grid) + (size_t)i * L.n_carriers * L.grid_w;
            c.evlog = i == 0 ? A.evlog : nullptr;
        }
        {   /* header: HBM / L2 -> shared memory (L1 bypassed: the last writer was another SM) */
            const uint4 *src = (const uint4 *)(hdr_g + i); uint4 *dst = (uint4 *)hdr_s;
            for (int k = lane; k < A.hdr_bytes16; k += 32) dst[k] = __ldcg(src + k);
        }
        __syncwarp();
        /* run the task; carry on with this cell while its next task is what this SM is working on (one call site: the handlers
         * are not duplicated in the kernel body) */
        int next = cls;
        NB_NOUNROLL
        for (int first = 1;; first = 0) {
            if (!first) {
                if (next == NB_TK_DONE) break;
                int cur = 0;
                if (lane == 0) cur = *(volatile int *)&s_cur;
                cur = __shfl_sync(0xFFFFFFFFu, cur, 0);
                if (next != cur) break;
                st_local += 1;
            }
            long long t0 = stats ? clock64() : 0;
            int k = next;
            next = k == NB_TK_BEGIN ? nbc_task_begin(c, ext_actions ? ext_actions + (size_t)i * s_ctrl.ctrl.ext_k : nullptr) : nbc_task_run(c, k);
            if (stats && lane == 0) { atomicAdd(&gq->tasks[k], 1ull); atomicAdd(&gq->cycles[k], (unsigned long long)(clock64() - t0)); }
        }
        __syncwarp();
        {
            const uint4 *src = (const uint4 *)hdr_s; uint4 *dst = (uint4 *)(hdr_g + i);
            for (int k = lane; k < A.hdr_bytes16; k += 32) dst[k] = src[k];
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            if (next == NB_TK_DONE) atomicSub(&gq->remaining, 1);
            else {
                unsigned int t = atomicAdd(&gq->c[next].tail, 1u);
                nbq_st_release(&slots[(size_t)next * cap + (t & cap_mask)], (unsigned int)i + 1u);
                atomicAdd(&gq->c[next].avail, 1);
            }
            atomicCAS(&s_mbox[warp], mb, 0ull);               /* idle again -- unless the dispatcher has already written NBQ_EXIT over it */
        }
        st_hops += 1;
        __syncwarp();
    }
    if (lane == 0) { atomicAdd(&gq->hops, st_hops); atomicAdd(&gq->local, st_local); }
}

/* ------------------------------------------------------------------ host side */
#define QCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, err_len, "%s: %s", #call, cudaGetErrorString(e_)); return NBIOT_E_CUDA; } } while (0)

#ifndef NB_Q_WORKERS
#define NB_Q_WORKERS 27
#endif

int nbq_create(NbQueueRt **out, int n_inst, int device, int /*unused*/, char *err, size_t err_len)
{
    QCK(cudaSetDevice(device));
    NbQueueRt *q = new NbQueueRt();
    memset(q, 0, sizeof *q);
    q->device = device; q->n_inst = n_inst;
    cudaDeviceProp prop;
    QCK(cudaGetDeviceProperties(&prop, device));
    q->workers = NB_Q_WORKERS;
    q->smem_per_warp = (int)(sizeof(NbCtx) + sizeof(NbHdr));
    q->smem_bytes = q->workers * q->smem_per_warp;
    q->blocks = prop.multiProcessorCount;
    if (const char *e = getenv("NBIOT_Q_BLOCKS")) { int v = atoi(e); if (v >= 1 && v < q->blocks) q->blocks = v; }   /* test knob */
    /* a slot index can run ahead of `tail` by at most the cells in flight, and lag behind by at most n_inst */
    unsigned int cap = 1024; while (cap < (unsigned int)n_inst + 8192u) cap <<= 1;
    q->cap_mask = cap - 1u;
    QCK(cudaMalloc(&q->gq, sizeof(NbGq)));
    QCK(cudaMalloc(&q->slots, sizeof(uint32_t) * (size_t)NB_TK_N * cap));
    QCK(cudaMemset(q->gq, 0, sizeof(NbGq)));
    QCK(cudaMemset(q->slots, 0, sizeof(uint32_t) * (size_t)NB_TK_N * cap));
    QCK(cudaFuncSetAttribute(nbiot_simg_kernel<NB_Q_WORKERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, q->smem_bytes));
    q->hop_min = getenv("NBIOT_Q_HOP_MIN") ? atoi(getenv("NBIOT_Q_HOP_MIN")) : 8;
    q->idle_before_hop = getenv("NBIOT_Q_IDLE") ? atoi(getenv("NBIOT_Q_IDLE")) : 20;
    q->stats = getenv("NBIOT_Q_STATS") ? atoi(getenv("NBIOT_Q_STATS")) : 0;
    q->batch_min = getenv("NBIOT_Q_BATCH") ? atoi(getenv("NBIOT_Q_BATCH")) : 1;
    if (q->batch_min < 1) q->batch_min = 1;
    if (q->batch_min > q->workers) q->batch_min = q->workers;
    *out = q;
    return 0;
}

void nbq_destroy(NbQueueRt *q)
{
    if (!q) return;
    cudaSetDevice(q->device);
    cudaFree(q->gq); cudaFree(q->slots);
    delete q;
}

int nbq_launch(NbQueueRt *q, const NbKernelArgs &A0, const uint8_t *ext_actions, const uint8_t *active, int until_time, int max_steps,
               cudaStream_t st, char *err, size_t err_len)
{
    NbKernelArgs A = A0;
    A.stage_grid = 0; A.smem_per_warp = q->smem_per_warp;
    /* counters restart every launch; the statistics at the end of NbGq survive */
    QCK(cudaMemsetAsync(q->gq, 0, offsetof(NbGq, tasks), st));
    nbq_init_kernel<<<(q->n_inst + 255) / 256, 256, 0, st>>>(q->gq, q->slots, q->cap_mask + 1u, q->n_inst, active);
    nbiot_simg_kernel<NB_Q_WORKERS><<<q->blocks, (NB_Q_WORKERS + 1) * 32, q->smem_bytes, st>>>(A, ext_actions, until_time, max_steps, q->gq, q->slots,
                                                                                                 q->cap_mask, q->hop_min, q->idle_before_hop, q->stats, q->batch_min);
    QCK(cudaGetLastError());
    return 0;
}

int nbq_stats(NbQueueRt *q, unsigned long long out[32])
{
    NbGq g;
    if (cudaSetDevice(q->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return NBIOT_E_CUDA;
    if (cudaMemcpy(&g, q->gq, sizeof g, cudaMemcpyDeviceToHost) != cudaSuccess) return NBIOT_E_CUDA;
    for (int k = 0; k < NB_TK_N; ++k) { out[k] = g.tasks[k]; out[NB_TK_N + k] = g.cycles[k]; }
    out[20] = g.hops; out[21] = g.local; out[22] = g.switches; out[23] = g.idle_polls; out[24] = (unsigned long long)g.bail;
    for (int k = 25; k < 32; ++k) out[k] = 0;
    return 0;
}

void nbq_geometry(NbQueueRt *q, int *blocks, int *warps_per_block, int *smem_bytes)
{
    if (blocks) *blocks = q->blocks;
    if (warps_per_block) *warps_per_block = q->workers + 1;
    if (smem_bytes) *smem_bytes = q->smem_bytes;
}
