/*
 * nbiot_state.h -- per-instance state of the batched NB-IoT cell simulator as it lives
 * in HBM (and, for the hot header + carrier grid, in shared memory while a kernel runs).
 *
 * The reference keeps one cell in Python objects: a heap of events
 * (system/event_manager.py:45-86), ordered lists / dicts of UE objects
 * (system/population.py:64-70, system/node_b.py:147-152), a boolean 12 x W grid
 * (system/carrier.py:430-444).  Here one instance is
 *   - a fixed-size header (NbHdr): clock, FSM, counters, system-event slots, NPRACH
 *     resource records, per-epoch info scratch;
 *   - a ring of 12-bit column masks per carrier (the look-ahead grid);
 *   - three random-access lists of 16-byte entries kept in list order (order is
 *     semantic: it pairs preambles with UEs and picks the msg3 grantee);
 *   - the connected queue as 16-byte entries kept sorted by (sort key, insertion stamp);
 *   - the contention set (UEs between msg3 grant and msg4 / MAC timeout);
 *   - a pool of 64-byte UE records addressed by slot, with a free-slot stack;
 *   - the per-period info lists (incoming / delays / service_times) and the per-occasion
 *     NPRACH histories.
 * All arrays are instance-major so that the warp that owns an instance reads contiguous
 * memory. Capacities are fixed at creation; exceeding one raises a per-instance fault.
 */
#ifndef NBIOT_STATE_H
#define NBIOT_STATE_H

#include <stdint.h>
#include "nbiot_conf.h"
#include "nbiot_ctrl.h"
#include "nbiot_record.h"
#include "nbiot_params.h"

#define NB_EV_POOL 48        /* pending RAR_window_end / Msg3 / NPUSCH_end events            */
#define NB_OCC_RING 24       /* pending NPRACH occasions per CE level (look-ahead / 80 sf)   */
#define NB_RAR_WIN 8         /* open RAR windows per CE level (300 sf window / 80 sf period) */
#define NB_SCLOG_CAP 512     /* non-anchor NPRACH starts not yet passed by the clock (2-carrier cells; cold ring in HBM): a
                                4096-column look-ahead holds at most 3 x 4096 / 40 of them */
#define NB_RA_SLACK 64       /* tombstones tolerated in a random-access list                 */
#define NB_FIXED_KEYS 10      /* fixed event slots (9 used)                                   */
#define NB_BLK_MAX 128       /* 32-column block summaries per carrier: grid_w <= 4096        */

/* event types kept in the generic pool */
enum { NB_EV_NONE = 0, NB_EV_RAR_END = 1, NB_EV_MSG3 = 2, NB_EV_NPUSCH_END = 3 };

/* containers a UE can be in */
enum { NB_IN_FREE = 0, NB_IN_RA = 1, NB_IN_CONTENTION = 2, NB_IN_CONNECTED = 3, NB_IN_SCHEDULED = 4 };

/* 64-bit event key: (time << 32) | insertion count -- the heap order of the reference */
#define NB_KEY(t, s) (((uint64_t)(uint32_t)(t) << 32) | (uint64_t)(uint32_t)(s))
#define NB_KEY_NONE 0xFFFFFFFFFFFFFFFFULL

typedef struct {            /* 16 B: one pending system event */
    uint64_t key;
    int32_t aux;            /* RAR_END: rar_w_id ; MSG3 / NPUSCH_END: UE slot */
    uint8_t type, ce, pad0, pad1;
} NbEvent;

typedef struct {            /* 12 B: one NPRACH occasion generated with its look-ahead column */
    int32_t t_start;        /* NPRACH_start fires at t_start (seq), NPRACH_end at t_start + N_sf (seq + 1) */
    uint32_t seq;
    uint16_t N_sf;          /* snapshot of the configuration in force when the column was generated */
    uint8_t N_sc, pad;
} NbOccasion;

typedef struct {            /* 24 B: carrier.py NprachResource (:293-299, :339-343) */
    int32_t t_start, t_next, sf_to_go, t_offset;
    int16_t periodicity, N_sf;
    uint8_t N_sc, sc_offset;
    uint16_t mask;          /* subcarrier mask of the running NPRACH */
} NbResource;

typedef struct {            /* 16 B: entry of a random-access list (population.ra_lists) */
    uint32_t wake_t;        /* BACKOFF: (wake_t, wake_seq) = key of its Backoff_end ; RA: preamble scratch */
    uint32_t wake_seq;
    int32_t rar_w_id;
    uint16_t slot;
    uint8_t st_ce;          /* UE state (low nibble) | CE_level << 4 ; 0xFF = removed (tombstone) */
    uint8_t attempts;       /* ra_attempts, saturating */
} NbRaEntry;
#define NB_RA_TOMB 0xFFu

typedef struct {            /* 16 B: entry of the connected queue, kept sorted by (k1, order) */
    uint64_t k1;            /* t_connection, or the bits of loss when sort_criterium == 'loss' */
    uint32_t order;         /* dict insertion stamp */
    uint16_t slot;
    uint8_t ce, pad;
} NbConnEntry;

typedef struct {            /* 16 B: UE in contention resolution; key = its MAC_timer event */
    uint64_t key;
    int32_t ue_id;
    uint16_t slot, ord;     /* ord: position in the contention list of its msg3 grant (0 unless capture_effect put several UEs on one grant) */
} NbContEntry;

typedef struct {            /* 64 B: reference system/user.py:35-70 minus never-read fields */
    double loss, sinr;
    int32_t id, t_arrival, t_connection, timestamp, acked_data, rar_w_id;
    uint16_t buffer, retx_buffer, bits, tbs;
    uint8_t I_tbs, N_rep, N_ru, CE_level, state, attempts, beta, new_data;
    uint8_t where, pad[7];
} NbUe;

typedef struct {            /* 8 B: carrier.py CElevelSClog entry: subcarriers a non-anchor NPRACH start at time t gave its CE level */
    int32_t t;
    uint8_t sc[4];
} NbSclogEntry;

typedef struct {            /* cold per-instance block in HBM: per-occasion NPRACH histories of the current update period,
                               clustered-placement memory (population.py:57-63), statistics cursor */
    int16_t det[3][NBIOT_MAX_OCC];
    int16_t col[3][NBIOT_MAX_OCC];
    int32_t cluster_count;  /* UEs still to be placed around the current cluster centre */
    int32_t stat_n;         /* departures recorded in the statistics histories since reset */
    double cluster_cx, cluster_cy;
    int32_t trace_n;        /* samples recorded in the sample ring since reset */
    int32_t pad_t;
} NbOccHist;

typedef struct __attribute__((aligned(16))) NbHdr {
    /* ---- clock, FSM, random stream ---- */
    uint64_t seed, draws;
    uint64_t cur_key;                 /* key of the event being handled (orders lazily retired Backoff_end) */
    int32_t time, state, fault;
    uint32_t seq;                     /* event insertion counter (event_manager.py:48) */
    int32_t events;
    int32_t NPDCCH_sf_left, valid_CE_control, total_departures;
    int32_t action3[3];
    /* keys of the fixed event slots, NB_KEY_NONE when nothing is pending:
     * 0 NPDCCH_arrival, 1 NPRACH_update, 2+ce next NPRACH_start, 5+ce next NPRACH_end, 8 next MAC_timer */
    uint64_t keys[NB_FIXED_KEYS];
    int32_t cur_ce, cur_wid, cur_t;   /* payload of the RAR_window_end being decided */
    uint8_t ctrl_action[24];          /* the controller's shared action vector (controller.py:33) */
    /* ---- NPRACH configuration ---- */
    uint8_t nprach_conf[16];          /* NPRACH_items order */
    int32_t RAR_window_size[3];
    int32_t msg3_nrep[3];
    /* ---- RAR bookkeeping (node_b.py:134-144) ---- */
    int32_t rar_n[3], rar_tot[3];
    int32_t RAR_UEs[3][NB_RAR_WIN], RAR_w_ids[3][NB_RAR_WIN];
    int32_t RAR_in[3], RAR_ids[3], RAR_attmp[3], RAR_sent[3], RAR_detected[3], RAR_failed[3];
    int32_t m3_cnt[3], m3_sent_sum[3], m3_recv_sum[3];
    double m3_eff_sum[3];
    int32_t occ_n[3], det_sum[3], col_sum[3];
    double distribution[15];
    int32_t k_dist;
    /* ---- scheduling queue ---- */
    int32_t n_sel;
    int32_t sel_slot[4];
    int32_t conn_n, cont_n, ra_n[3], n_scheduled, free_top;
    uint32_t conn_order;
    uint64_t next_mac;                /* cached min key over the contention set, NB_KEY_NONE if unknown/empty */
    int32_t next_mac_idx, mac_valid;
    /* ---- carrier ---- */
    int32_t t_now, t_ahead;
#if defined(NB_GRID_MODE) && NB_GRID_MODE == 2
    uint16_t blk[2][NB_BLK_MAX];      /* OR of the column masks of each 32-column block of the ring (window tests) */
#endif
    int32_t max_lookahead, max_live;  /* high-water marks since reset: columns kept ahead, live UEs (capacity planning) */
    NbResource res[2][3];
    int32_t occ_head_start[3], occ_head_end[3], occ_tail[3];   /* ring cursors (monotone counters) */
    NbOccasion occ[3][NB_OCC_RING];
    int32_t sclog_n, sclog_lo;        /* cursors of the cold ring of non-anchor NPRACH starts (NbSclogEntry, monotone counters) */
    /* ---- population / access procedure ---- */
    int32_t ue_count, CE_thresholds[2], preamble_trans_max, MAC_timer;
    double probability_anchor;
    int32_t detected_pre[3], collided_pre[3], contenders[3], RAR_counter[3];
    /* ---- traffic generator ---- */
    int32_t M_uniform, M_beta, M_beta_max, gen_t, gen_counter, pad_g;
    double gen_p;
    /* ---- performance monitor ---- */
    int32_t perf_t, error_count, attempts_count, thr_count;
    int64_t total_bits, msg3_resources, NPUSCH_resources, NPRACH_resources;
    int32_t ue_arrivals[3], ue_departures[3], backoffs[3], ue_connections;
    double av_delay;
    int32_t n_rx, n_err, n_unfit, n_acc;
    int32_t rx_id[NBIOT_STEP_LIST], rx_delay[NBIOT_STEP_LIST], err_id[NBIOT_STEP_LIST];
    int32_t unfit_id[NBIOT_STEP_LIST], acc_id[NBIOT_STEP_LIST], acc_delay[NBIOT_STEP_LIST];
    double rx_sum, rx_inv_sum, rx_log_sum;          /* running reward terms over this epoch's receptions */
    int32_t n_incoming, n_departures;
    int64_t sum_delays;
    int64_t subframes_done;           /* simulated time advanced since reset (the throughput unit) */
    /* ---- preamble scratch (NPRACH pass) ---- */
    uint32_t pre_once[8], pre_twice[8];
    /* ---- pending system events ---- */
    uint64_t pool_free;               /* bit i set: pool[i] is free */
    NbEvent pool[NB_EV_POOL];
    /* ---- control state between two tasks of the migrating-instance kernel (nbiot_core.cuh "tasks"): the NodeB steps taken
     *      in this launch, the phase of the controller loop, and the arguments of the event handler that runs next
     *      (at the end: the offsets of everything the warp-per-instance kernel touches stay what they were) ---- */
    int32_t tk_steps, tk_phase, tk_a, tk_b;
    NbOccasion tk_occ;
    int32_t tk_pad;
} NbHdr;

/* device-side form of nbiot_ctrl_t (include/nbiot_ctrl.h): write tables as (bit mask, values) */
typedef struct { uint32_t mask; uint8_t val[24]; uint32_t pad; } NbWrites;
typedef struct {
    uint32_t ext_state_mask; int32_t ext_k;
    uint8_t ext_a_mask[24];
    NbWrites ext_post, state_writes[4];
    uint8_t init_action[24];
    NbWrites full_writes[4];              /* every agent of the state (advances without an external action) */
} NbCtrlDev;

/* configuration read on rare paths only (arrivals, departures): kept out of the per-warp context */
typedef struct {
    int32_t stat_cap, clustered, cluster_lo, cluster_hi;
    double cluster_prob, cluster_range;   /* cluster_range already divided by the cell range (population.py:62) */
    NbOccHist *occ_base;                  /* instance index = c.occ - occ_base */
    int32_t *stat_base;                   /* int32[n_inst][4][stat_cap]: delay, connection, access, acked bits */
    nbiot_trace_rec_t *trace_base;        /* [n_inst][trace_cap] */
    int32_t trace_cap, trace_mask;
    int32_t capture_effect, pad_cold;     /* Channel(capture_effect=True), channel.py:124-126 */
    NbSclogEntry *sclog_base;             /* [n_inst][NB_SCLOG_CAP], two-carrier batches only */
} NbColdConf;

/* per-CTA constants in shared memory: NbCtx.ctrl points at the first member */
typedef struct { NbCtrlDev ctrl; NbColdConf cold; } NbCtaConst;

static inline NbCtrlDev nb_make_ctrl_dev(const nbiot_ctrl_t *c)
{
    NbCtrlDev d;
    d.ext_state_mask = c->ext_state_mask; d.ext_k = c->ext_k;
    for (int i = 0; i < 24; ++i) { d.ext_a_mask[i] = c->ext_a_mask[i]; d.init_action[i] = c->init_action[i]; }
    for (int t = 0; t < 9; ++t) {
        NbWrites *w = t < 4 ? &d.state_writes[t] : t == 4 ? &d.ext_post : &d.full_writes[t - 5];
        const int8_t *src = t < 4 ? c->state_writes[t] : t == 4 ? c->ext_post : c->full_writes[t - 5];
        w->mask = 0; w->pad = 0;
        for (int i = 0; i < 24; ++i) { w->val[i] = 0; if (i < 23 && src[i] >= 0) { w->mask |= 1u << i; w->val[i] = (uint8_t)src[i]; } }
    }
    return d;
}

/* byte offsets of the per-instance arrays inside the caller-owned state buffer */
typedef struct {
    int32_t n_inst, ue_cap, grid_w, hist_cap, n_carriers, ra_cap, stat_cap, trace_cap;
    uint64_t off_hdr, off_grid, off_ra, off_conn, off_cont, off_ue, off_free, off_occ;
    uint64_t off_incoming, off_delays, off_service, off_rec, off_scratch, off_stat, off_trace, off_spill, off_sclog, total;
} NbLayout;

static inline uint64_t nb_align256(uint64_t x) { return (x + 255u) & ~(uint64_t)255u; }
#define nb_spill_cap(hist_cap) ((hist_cap) > NBIOT_STEP_SPILL ? (hist_cap) : NBIOT_STEP_SPILL)   /* entries per spilled info list (host and device) */

static inline NbColdConf nb_make_cold(const nbiot_conf_t *c, uint8_t *state, uint64_t off_occ, uint64_t off_stat, uint64_t off_trace, uint64_t off_sclog)
{
    NbColdConf k;
    double range = c->max_d > 0 ? c->max_d : NB_MAX_RANGE;      /* channel.py:128-131 */
    k.stat_cap = c->stat_cap > 0 ? c->stat_cap : 0; k.clustered = c->clustered; k.cluster_lo = c->cluster_lo; k.cluster_hi = c->cluster_hi;
    k.cluster_prob = c->cluster_prob; k.cluster_range = c->cluster_range / range;
    k.occ_base = (NbOccHist *)(state + off_occ); k.stat_base = (int32_t *)(state + off_stat);
    k.trace_base = (nbiot_trace_rec_t *)(state + off_trace);
    k.trace_cap = c->trace_cap > 0 ? c->trace_cap : 0; k.trace_mask = k.trace_cap ? c->trace_mask : 0;
    k.capture_effect = c->capture_effect ? 1 : 0; k.pad_cold = 0;
    k.sclog_base = (NbSclogEntry *)(state + off_sclog);
    return k;
}

static inline NbLayout nb_make_layout(const nbiot_conf_t *c, int n)
{
    NbLayout L;
    L.n_inst = n; L.ue_cap = c->ue_cap; L.grid_w = c->grid_w; L.hist_cap = c->hist_cap;
    L.n_carriers = c->n_carriers >= 2 ? 2 : 1; L.ra_cap = c->ue_cap + NB_RA_SLACK; L.stat_cap = c->stat_cap > 0 ? c->stat_cap : 0; L.trace_cap = c->trace_cap > 0 ? c->trace_cap : 0;
    uint64_t o = 0, N = (uint64_t)n;
    L.off_hdr = o;      o = nb_align256(o + N * sizeof(NbHdr));
    L.off_grid = o;     o = nb_align256(o + N * (uint64_t)L.n_carriers * (uint64_t)L.grid_w * 2u);
    L.off_ra = o;       o = nb_align256(o + N * 3u * (uint64_t)L.ra_cap * sizeof(NbRaEntry));
    L.off_conn = o;     o = nb_align256(o + N * (uint64_t)L.ue_cap * sizeof(NbConnEntry));
    L.off_cont = o;     o = nb_align256(o + N * (uint64_t)L.ue_cap * sizeof(NbContEntry));
    L.off_ue = o;       o = nb_align256(o + N * (uint64_t)L.ue_cap * sizeof(NbUe));
    L.off_free = o;     o = nb_align256(o + N * (uint64_t)L.ue_cap * 2u);
    L.off_occ = o;      o = nb_align256(o + N * sizeof(NbOccHist));
    L.off_incoming = o; o = nb_align256(o + N * (uint64_t)L.hist_cap * 8u);
    L.off_delays = o;   o = nb_align256(o + N * (uint64_t)L.hist_cap * 4u);
    L.off_service = o;  o = nb_align256(o + N * (uint64_t)L.hist_cap * 4u);
    L.off_rec = o;      o = nb_align256(o + N * sizeof(nbiot_record_t));
    L.off_scratch = o;  o = nb_align256(o + N * (uint64_t)L.ra_cap * 2u);
    L.off_stat = o;     o = nb_align256(o + N * 4u * (uint64_t)L.stat_cap * 4u);
    L.off_trace = o;    o = nb_align256(o + N * (uint64_t)L.trace_cap * sizeof(nbiot_trace_rec_t));
    L.off_spill = o;    o = nb_align256(o + N * 6u * (uint64_t)nb_spill_cap(L.hist_cap) * 4u);
    L.off_sclog = o;    o = nb_align256(o + (L.n_carriers == 2 ? N * (uint64_t)NB_SCLOG_CAP * sizeof(NbSclogEntry) : 0u));
    L.total = o;
    return L;
}

#endif /* NBIOT_STATE_H */
