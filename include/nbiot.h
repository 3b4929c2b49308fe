/*
 * nbiot.h -- C ABI of libnbiot_b200.so, the B200-native batched NB-IoT cell simulator.
 *
 * The reference has no native code and no vector environment: one
 * cell is a graph of Python objects built by create_system(rng, conf)
 * (system/system_creator.py:22-59) and driven through Controller.reset()/step()
 * (controller.py:132-143, :234-260) behind gym_system.Environment
 * (gym-system/gym_system/environment.py:22-37). Each entry point below names the reference
 * interface it replaces for a BATCH of n_inst independent cells; the Python binding a
 * maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * NBIOT_E_* code otherwise, with a message available from nbiot_last_error(); calls on one handle
 * are stream-ordered and asynchronous unless stated; a handle is single-owner (no internal
 * locking); there is no CPU fallback -- without a CUDA device every compute call fails.
 */
#ifndef NBIOT_H
#define NBIOT_H

#include <stddef.h>
#include <stdint.h>

#include "nbiot_conf.h"
#include "nbiot_record.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NBIOT_ABI_VERSION 1

#define NBIOT_E_INVALID (-1)   /* bad argument */
#define NBIOT_E_CUDA (-2)      /* CUDA runtime error (text in nbiot_last_error) */
#define NBIOT_E_NOGPU (-3)     /* no usable CUDA device */
#define NBIOT_E_SIZE (-4)      /* caller-provided buffer too small */

typedef struct nbiot_handle nbiot_handle_t;
struct nbiot_ctrl;             /* controller write tables, include/nbiot_ctrl.h */

/* one observation item of Controller.get_obs (controller.py:105-120): obs = min(1, value / norm) */
typedef struct nbiot_obs_item {
    int32_t offset;            /* byte offset of the first element inside nbiot_record_t          */
    int32_t kind;              /* 0: int32, 1: double, 2: carrier_state (1 - busy_u8 / 12)         */
    int32_t length;            /* number of elements                                              */
    uint32_t state_mask;       /* NodeB states whose info dict holds the key (others give zeros)  */
    double norm;               /* parameters.obs_dict normalisation constant                      */
} nbiot_obs_item_t;

#define NBIOT_N_STATS 24       /* length of the nbiot_reduce_stats vector, see nbiot_stat_index    */
enum nbiot_stat_index {
    NBIOT_STAT_ARRIVALS_CE0 = 0,   /* PerfMonitor.ue_arrivals[3]    (perf_monitor.py:52)  */
    NBIOT_STAT_DEPARTURES_CE0 = 3, /* PerfMonitor.ue_departures[3]  (:55)                 */
    NBIOT_STAT_BACKOFFS_CE0 = 6,   /* PerfMonitor.backoffs[3]       (:56)                 */
    NBIOT_STAT_CONNECTIONS = 9,    /* ue_connections (:53)                                */
    NBIOT_STAT_ERRORS = 10,        /* error_count (:43)                                   */
    NBIOT_STAT_ATTEMPTS = 11,      /* attempts_count (:44)                                */
    NBIOT_STAT_TOTAL_BITS = 12,    /* total_bits (:46)                                    */
    NBIOT_STAT_SUBFRAMES = 13,     /* simulated subframes advanced since reset (throughput unit) */
    NBIOT_STAT_EVENTS = 14,        /* events handled                                      */
    NBIOT_STAT_DRAWS = 15,         /* random counters consumed                            */
    NBIOT_STAT_FAULTED = 16,       /* instances stopped by a fatal fault                  */
    NBIOT_STAT_LIVE_UES = 17,      /* UEs currently in the cell (sum over instances)      */
    NBIOT_STAT_SUM_DELAY_X1000 = 18, /* sum of PerfMonitor.av_delay * departures * 1000 (fixed point) */
    NBIOT_STAT_TRUNCATED = 19,     /* instances with a truncated info list                */
    NBIOT_STAT_MAX_LOOKAHEAD = 20, /* MAX over instances: carrier columns kept ahead (<= grid_w)     */
    NBIOT_STAT_MAX_LIVE_UES = 21   /* MAX over instances: simultaneous live UEs (<= ue_cap)          */
};

int nbiot_abi_version(void);
const char *nbiot_last_error(void);

/* Bytes of device memory one batch needs. The CALLER allocates them (e.g. a torch uint8 tensor)
 * and keeps ownership: the buffer is the complete simulation state, so saving it is a checkpoint. */
size_t nbiot_state_bytes(const nbiot_conf_t *conf, int n_inst);
size_t nbiot_hdr_bytes(void);
/* trailing bytes of each header that hold the task-control words of the migrating-instance kernel (NBIOT_SCHED=queue); only that
 * kernel writes them, everything else in the state buffer is independent of the kernel that produced it */
size_t nbiot_hdr_task_bytes(void);

/* replaces create_system(rng, conf) + Controller(node, agents) for n_inst cells
 * (system/system_creator.py:22-59, controller.py:30-41). seeds/lut2/mcs are host pointers:
 * seeds[n_inst] Philox keys, lut2 = uint16[7*8*501] (LUT2.csv), mcs = uint8[500*30*3]
 * (loss_tbs_to_mcs.p), see tools/make_tables.py. */
int nbiot_create(const nbiot_conf_t *conf, int n_inst, int device, void *state_dev, size_t state_bytes,
                 const uint64_t *seeds, const uint16_t *lut2, const uint8_t *mcs, const struct nbiot_ctrl *ctrl,
                 void *stream, nbiot_handle_t **out);
int nbiot_destroy(nbiot_handle_t *h);

/* replaces Controller.set_ext_agent / replace_agent / DummyAgent.set_action (controller.py:71-75, :122-130) */
int nbiot_set_controller(nbiot_handle_t *h, const struct nbiot_ctrl *ctrl, void *stream);

/* replaces Controller.reset -> NodeB.reset (controller.py:132-143, system/node_b.py:163-189) */
int nbiot_reset(nbiot_handle_t *h, void *stream);

/* replaces Controller.step + to_next_step (controller.py:234-288), or Controller.run_until (:185-195)
 * when ext_actions is NULL. ext_actions: device uint8[n_inst][ext_k]. until_time < 0: no time bound. */
int nbiot_advance(nbiot_handle_t *h, const uint8_t *ext_actions_dev, int until_time, int max_steps, void *stream);
/* the same for a subset: instances with active_dev[i] == 0 are left untouched (their record, clock and random
 * stream do not move) -- what a per-instance wrapper does when it answers without calling env.step
 * (wrappers.py:94-103). active_dev: device uint8[n_inst], NULL = all. */
int nbiot_advance_masked(nbiot_handle_t *h, const uint8_t *ext_actions_dev, const uint8_t *active_dev, int until_time, int max_steps, void *stream);

/* device pointer to nbiot_record_t[n_inst] (reward + info dict of NodeB.get_node_info, node_b.py:278-418)
 * and to the per-period lists incoming (double) / delays / service_times (int32), hist_cap entries each */
int nbiot_records_ptr(nbiot_handle_t *h, void **records_dev);
int nbiot_hist_ptrs(nbiot_handle_t *h, void **incoming_dev, void **delays_dev, void **service_dev, int *hist_cap);

/* the per-cell sample ring of conf['traces'] / conf['statistics'] (PerfMonitor.nprach_sample, rar_sample, rar_window_sample,
 * register_*_resources, npusch_delivered; perf_monitor.py:61-79, :111-140, :161-202): trace_dev = nbiot_trace_rec_t[n_inst][trace_cap]
 * inside the state buffer (NULL when off), entry k of an instance at k % trace_cap; counts_dev = int32[n_inst] samples since reset */
int nbiot_trace_ptrs(nbiot_handle_t *h, void **trace_dev, int *trace_cap);
/* the per-epoch info lists (info['receptions'], ['errors'], ['unfit'], ['access']; perf_monitor.py:92-103) beyond the list_cap = 16 entries
 * the record holds: spill_dev = int32[n_inst][6][spill_cap] inside the state buffer, rows = rx id, rx delay, error id, unfit id, access id,
 * access delay; entry k >= list_cap of a list at column k - list_cap (the record's n_* fields are the true lengths) */
int nbiot_spill_ptr(nbiot_handle_t *h, void **spill_dev, int *list_cap, int *spill_cap);   /* spill_cap = max(NBIOT_STEP_SPILL, hist_cap) */
int nbiot_trace_counts(nbiot_handle_t *h, int32_t *counts_dev, void *stream);
/* conf.stat_cap > 0 (conf['statistics']): the per-departure histories of PerfMonitor (perf_monitor.py:61-84, :161-171; the arrays
 * experiments_RL.py:139-140 / experiments_MAMBRL.py:121-122 save) as rings in HBM.
 * stat_dev = int32[n_inst][4][stat_cap]: 0 delay_history, 1 connection_history, 2 access_history, 3 acked bits (th_history = bits / delay);
 * entry k of an instance is departure number (k + m * stat_cap) since reset -- nbiot_stat_counts gives the number recorded. */
int nbiot_stat_ptrs(nbiot_handle_t *h, void **stat_dev, int *stat_cap);
int nbiot_stat_counts(nbiot_handle_t *h, int32_t *counts_dev, void *stream);
/* histogram of one history over ALL instances, on the device: hist_dev = int64[n_bins], bin b counts values in
 * [b * bin_width, (b + 1) * bin_width), the last bin is open-ended (delay CDFs of a batch without a host round trip) */
int nbiot_stat_histogram(nbiot_handle_t *h, int which, int bin_width, int n_bins, int64_t *hist_dev, void *stream);

/* replaces Controller.get_obs (controller.py:105-120): obs_dev = double[n_inst][sum(length)],
 * reward_dev = double[n_inst] (either may be NULL) */
int nbiot_gather_obs(nbiot_handle_t *h, const nbiot_obs_item_t *items, int n_items, double *obs_dev, double *reward_dev, void *stream);

/* episode statistics over the instances of this handle: out_dev = int64[NBIOT_N_STATS]; entries below
 * NBIOT_STAT_MAX_LOOKAHEAD are sums (all-reduce with SUM across GPUs), the last two are maxima (MAX) */
int nbiot_reduce_stats(nbiot_handle_t *h, int64_t *out_dev, void *stream);

/* the reference-facing call with HOST buffers: copies actions host->device, advances, gathers the
 * observation, copies obs/reward device->host and synchronises (gym step on a batch).
 * actions_host: uint8[n_inst][ext_k] or NULL; obs_host: double[n_inst][obs_dim]; reward_host: double[n_inst] */
int nbiot_step_host(nbiot_handle_t *h, const uint8_t *actions_host, int until_time, int max_steps,
                    const nbiot_obs_item_t *items, int n_items, double *obs_host, double *reward_host, void *stream);

/* debugging: log (time, type) of every event of instance 0 into evlog_dev = int32[cap][2] */
int nbiot_set_event_log(nbiot_handle_t *h, int32_t *evlog_dev, int cap);
/* kernel launches issued through this handle so far */
long long nbiot_launch_count(nbiot_handle_t *h);
/* measurement aid (NBIOT_PROF_INST=1 at creation): { start ns, end ns, SM id, 0 } of every instance in the last simulation launch,
 * out_host = uint64[n_inst][4]; synchronises the device */
int nbiot_instance_times(nbiot_handle_t *h, unsigned long long *out_host);
/* launch geometry chosen at creation: blocks, warps per block, dynamic shared memory per block */
int nbiot_launch_geometry(nbiot_handle_t *h, int *blocks, int *warps_per_block, int *smem_bytes);
/* which simulation kernel advances this batch: "warp-per-cell", "thread-per-cell" or "thread-per-cell, voted dispatch" (chosen at
 * creation by batch size; NBIOT_KERNEL=warp|tpc|tpcv in the environment overrides; the results are bit-identical) */
const char *nbiot_kernel_name(nbiot_handle_t *h);

/* host twins of the counter-based random stream (include/nbiot_rng.h): the same compiled arithmetic
 * the kernels use, for the generator shim that drives the reference (rng argument of
 * system/system_creator.py:22) */
double nbiot_rng_u01(uint64_t seed, uint64_t ctr, uint32_t stream);
void nbiot_rng_pair(uint64_t seed, uint64_t ctr, uint32_t stream, double *xy);
double nbiot_rng_normal(uint64_t seed, uint64_t ctr, uint32_t stream);
uint64_t nbiot_rng_bounded(uint64_t seed, uint64_t ctr, uint32_t stream, uint64_t n);

#ifdef __cplusplus
}
#endif

#endif /* NBIOT_H */
