/*
 * nbiot_oracle.c -- CPU ORACLE (test infrastructure, not a product path).
 *
 * A plain-C restatement of the reference's discrete-event NB-IoT cell simulator for
 * ONE cell instance, structured like the reference: a binary-heap future event list
 * keyed (time, insertion count), ordered Python-style containers, one handler per
 * event type. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (nb_iot_environment_b200) never does.
 *
 * Parity status: PINNED by golden traces generated in the build container by running
 * the unmodified Python reference (/root/reference) with the Philox shim injected as
 * its `rng` (tools/gen_golden.py -> tests/golden/; tests/test_oracle_golden.py).
 * The reference ships no tests or golden vectors of its own.
 *
 * Reference files followed (path:line relative to the reference root):
 *   system/event_manager.py:45-92      FEL heap, schedule_event
 *   system/node_b.py:131-830           NodeB FSM, step handlers, get_node_info
 *   system/carrier.py:115-252,278-404,414-650  compute_offsets, NprachResource, SFGenerator, Carrier
 *   system/population.py:64-320        Population
 *   system/access_procedure.py:28-122  AccessProcedure
 *   system/channel.py:76-256           Channel
 *   system/rx_procedure.py:29-61       RxProcedure
 *   system/ue_generator.py:38-104      UEGenerator
 *   system/perf_monitor.py:41-322      PerfMonitor
 *   system/action_reader.py:25-83      ActionReader
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>
#include <time.h>

#include "nbiot_conf.h"
#include "nbiot_ctrl.h"
#include "nbiot_params.h"
#include "nbiot_record.h"
#include "nbiot_rng.h"

/* ------------------------------------------------------------------ small vectors */
#define VEC(T) struct { T *d; int n, cap; }
#define VPUSH(v, x) do { if ((v).n == (v).cap) { (v).cap = (v).cap ? 2 * (v).cap : 8; \
    (v).d = realloc((v).d, sizeof(*(v).d) * (size_t)(v).cap); } (v).d[(v).n++] = (x); } while (0)
#define VCLEAR(v) ((v).n = 0)
#define VFREE(v) do { free((v).d); (v).d = NULL; (v).n = (v).cap = 0; } while (0)
#define VDEL(v, i) do { memmove(&(v).d[i], &(v).d[(i) + 1], sizeof(*(v).d) * (size_t)((v).n - (i) - 1)); (v).n--; } while (0)

typedef VEC(int) ivec;
typedef VEC(double) dvec;

/* ------------------------------------------------------------------ UE (system/user.py:35-70) */
typedef struct UE {
    int id, state, beta, CE_level, t_arrival, ra_attempts, preamble, rar_w_id;
    int t_contention, t_connection, t_disconnection, backoff_counter, timeout_counter, harq_rtx;
    int buffer, retx_buffer;
    double x, y, sinr, loss;
    int I_tbs, N_rep, N_ru, bits, tbs, new_data, acked_data;
    int timestamp;               /* perf_monitor.ue_timestamp[id] */
} UE;
typedef VEC(UE *) uvec;

/* ------------------------------------------------------------------ events (system/event_manager.py) */
enum { EV_NPDCCH_ARRIVAL, EV_NPRACH_UPDATE, EV_RAR_WINDOW_END, EV_NPRACH_START, EV_NPRACH_END,
       EV_MSG3, EV_NPUSCH_END, EV_MAC_TIMER, EV_BACKOFF_END };

typedef struct {
    int type, t, CE_level, carrier, N_pream, N_rep, N_sf, N_sc, rar_w_id;
    UE *ue;       /* Msg3 ue_list[0], NPUSCH_end, Backoff_end */
    int ue_id;    /* MAC_timer ue_id_list[0] */
    UE **ue_list; int *id_list; int n_list;   /* Msg3 ue_list / MAC_timer ue_id_list with several contenders (capture effect); NULL / 0 for one */
} Event;

typedef struct { int time; uint64_t count; Event ev; } HeapEntry;

/* ------------------------------------------------------------------ carrier */
typedef struct {
    int t_start, t_next, sf_to_go, t_offset, periodicity, N_sf, N_sc, sc_offset;
    uint16_t nprach_sf;          /* 12-bit subcarrier mask of the running NPRACH */
} NprachResource;

typedef struct { int t; int sc[3]; } ScLogEntry;

typedef struct Oracle {
    nbiot_conf_t conf;
    uint64_t seed, draws;
    const uint16_t *lut2;        /* [56][501] BLER * 2^15 */
    const uint8_t *mcs;          /* [500][30][3] (I_tbs, N_rep, N_ru) */
    double range;
    int n_carriers;

    /* FEL */
    HeapEntry *heap; int heap_n, heap_cap; uint64_t counter;
    long long pops;

    /* NodeB (node_b.py:87-161) */
    int NPDCCH_sf_left;
    int NPRACH_conf[16];
    int RAR_window_size[3];
    int valid_CE_control;
    int action[3], final_action[3];
    int total_departures;
    int state, time;
    Event event;
    ivec RAR_UEs[3], RAR_w_ids[3];
    int RAR_in[3], RAR_ids[3], RAR_attmp[3], RAR_sent[3], RAR_detected[3], RAR_failed[3];
    ivec msg3_sent[3], msg3_received[3]; dvec msg3_efficiency[3];
    uvec connected;              /* dict in insertion order */
    int selectable[4], n_selectable;
    int n_scheduled;
    dvec incoming; ivec departures, service_times;
    double distribution[15]; int k;
    int msg3_conf_nrep[3];       /* module-global msg3_default_conf (node_b.py:36): (2,1,N_rep) */

    /* carrier (carrier.py:430-444) */
    NprachResource res[2][3];
    uint16_t *cols[2]; int col_cap;  /* ring of 12-bit column masks, index t & (cap-1) */
    int t_now, t_ahead;
    VEC(ScLogEntry) sc_log;

    /* population (population.py:41-70) */
    uvec ra_lists[3];
    uvec contention;
    int ue_count;
    int CE_thresholds[2];
    int preamble_trans_max, MAC_timer;

    /* access procedure (access_procedure.py:28-35) */
    double probability_anchor;
    int detected_preambles_per_CE[3], collided_preambles_per_CE[3], contenders_per_CE[3], RAR_counter[3];
    ivec detection_history[3], collision_history[3];

    /* clustered placement (population.py:57-63): survives reset() */
    int cluster_count; double cluster_center[2], cluster_range;

    /* ue generator (ue_generator.py:23-42) */
    int M_uniform, M_beta, M_beta_max, gen_t, gen_counter; double gen_p;

    /* perf monitor (perf_monitor.py:41-59) */
    int perf_t, error_count, attempts_count, thr_count; long long total_bits;
    int ue_arrivals[3], ue_departures[3], backoffs[3], ue_connections;
    double av_delay;
    long long msg3_resources, NPUSCH_resources, NPRACH_resources;
    ivec rx_id, rx_delay, err_id, unfit_id, acc_id, acc_delay;
    ivec access_history, connection_history, delay_history, bits_history;   /* statistics (perf_monitor.py:61-84); th = bits / delay */
    VEC(nbiot_trace_rec_t) samples;   /* nprach_sample / rar_sample / rar_window_sample / register_*_resources / npusch_delivered (perf_monitor.py:61-79) */

    int fault;
    /* last NPRACH_update info lists (copied out on request) */
    dvec info_incoming; ivec info_delays, info_service;
    /* optional event log */
    int *evlog; int evlog_n, evlog_cap;
} Oracle;

/* ------------------------------------------------------------------ rng (law in nbiot_rng.h) */
static double rng_u01(Oracle *o) { return nb_rng_u01(o->seed, o->draws++, NBIOT_STREAM_ENV); }
static double rng_uniform(Oracle *o, double a, double b) { return a + (b - a) * rng_u01(o); }
static double rng_normal(Oracle *o, double mu, double sigma)
{ return mu + sigma * nb_rng_normal(o->seed, o->draws++, NBIOT_STREAM_ENV); }
static int rng_bernoulli(Oracle *o, double p) { return nb_rng_bernoulli(o->seed, o->draws++, NBIOT_STREAM_ENV, p); }
static int rng_binomial(Oracle *o, int n, double p) { int s = 0; for (int i = 0; i < n; ++i) s += rng_bernoulli(o, p); return s; }
static int rng_integers(Oracle *o, int low, int high)
{ return low + (int)nb_rng_bounded(o->seed, o->draws++, NBIOT_STREAM_ENV, (uint64_t)(high - low)); }

/* ------------------------------------------------------------------ FEL */
static int heap_less(const HeapEntry *a, const HeapEntry *b)
{ return a->time < b->time || (a->time == b->time && a->count < b->count); }

static void schedule_event(Oracle *o, int time, Event ev)
{
    ev.t = time;
    if (o->heap_n == o->heap_cap) { o->heap_cap = o->heap_cap ? 2 * o->heap_cap : 64;
        o->heap = realloc(o->heap, sizeof(HeapEntry) * (size_t)o->heap_cap); }
    HeapEntry e = { time, o->counter++, ev };
    int i = o->heap_n++;
    while (i > 0) { int p = (i - 1) / 2; if (!heap_less(&e, &o->heap[p])) break; o->heap[i] = o->heap[p]; i = p; }
    o->heap[i] = e;
}

static Event fel_pop(Oracle *o)
{
    HeapEntry top = o->heap[0];
    HeapEntry last = o->heap[--o->heap_n];
    int i = 0, n = o->heap_n;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && heap_less(&o->heap[c + 1], &o->heap[c])) c++;
        if (!heap_less(&o->heap[c], &last)) break;
        o->heap[i] = o->heap[c]; i = c;
    }
    if (n > 0) o->heap[i] = last;
    return top.ev;
}

static Event make_event(int type) { Event e; memset(&e, 0, sizeof e); e.type = type; return e; }

/* ------------------------------------------------------------------ perf monitor */
static void perf_reset(Oracle *o)
{
    o->perf_t = 0; o->error_count = o->attempts_count = o->thr_count = 0; o->total_bits = 0;
    VCLEAR(o->rx_id); VCLEAR(o->rx_delay); VCLEAR(o->err_id); VCLEAR(o->unfit_id); VCLEAR(o->acc_id); VCLEAR(o->acc_delay);
    for (int i = 0; i < 3; ++i) o->ue_arrivals[i] = o->ue_departures[i] = o->backoffs[i] = 0;
    o->ue_connections = 0; o->av_delay = 0; o->msg3_resources = o->NPUSCH_resources = o->NPRACH_resources = 0;
    VCLEAR(o->access_history); VCLEAR(o->connection_history); VCLEAR(o->delay_history); VCLEAR(o->bits_history);
    VCLEAR(o->samples);
}
static void perf_clear_info(Oracle *o)
{ VCLEAR(o->rx_id); VCLEAR(o->rx_delay); VCLEAR(o->err_id); VCLEAR(o->unfit_id); VCLEAR(o->acc_id); VCLEAR(o->acc_delay); }

static void dict_set(ivec *keys, ivec *vals, int key, int val)
{
    for (int i = 0; i < keys->n; ++i) if (keys->d[i] == key) { vals->d[i] = val; return; }
    VPUSH(*keys, key); VPUSH(*vals, val);
}

static double perf_reward(Oracle *o)
{
    double tt = 0.0; int c = o->rx_id.n;
    switch (o->conf.reward_criteria) {
    case NB_RW_THROUGHPUT: return (double)o->thr_count;                      /* perf_monitor.py:276-280 */
    case NB_RW_AVERAGE_DELAY:                                                /* :253-262 */
        for (int i = 0; i < c; ++i) tt += o->rx_delay.d[i];
        return -1.0 * tt / (double)(c > 1 ? c : 1);
    case NB_RW_ACCUMULATED_DELAY:                                            /* :282-289 */
        for (int i = 0; i < c; ++i) tt += o->rx_delay.d[i];
        return -1.0 * tt;
    case NB_RW_USERS: return (double)c;                                      /* :291-295 */
    case NB_RW_INVD_USERS:                                                   /* :306-313 */
        for (int i = 0; i < c; ++i) tt += 1.0 / ((double)o->rx_delay.d[i] / 1000.0);
        return tt;
    case NB_RW_LOG_INVD_USERS:                                               /* :315-322 */
        for (int i = 0; i < c; ++i) tt += nb_log((double)o->rx_delay.d[i] / 1000.0);
        return -1.0 * tt;
    default: return 0.0;
    }
}

/* ------------------------------------------------------------------ ue generator (ue_generator.py) */
static void gen_reset(Oracle *o)
{
    double r = o->conf.ratio < 1.0 ? o->conf.ratio : 1.0;
    o->M_uniform = (int)rint((double)o->conf.M * r);        /* Python round(): half-to-even */
    o->M_beta = o->conf.M - o->M_uniform;
    o->M_beta_max = o->M_beta;
    o->gen_t = 0;
}

static void gen_arrivals(Oracle *o, int t, int *arrivals_out, int *beta_out)
{
    if (o->conf.random_traffic) {                            /* update_counter :45-52 */
        int new_count = t / NB_T_UNIFORM;
        if (new_count > o->gen_counter) { o->gen_counter = new_count; o->gen_p = rng_uniform(o, 0.0, 0.8); }
    }
    int t_elapsed = t - o->gen_t;
    double arrivals = 0, beta_arrivals = 0;
    if (o->M_uniform > 0) {
        double intensity = (double)o->M_uniform * ((double)t_elapsed / (double)NB_T_UNIFORM);
        double ip; double p = modf(intensity, &ip);
        arrivals = ip + (double)rng_bernoulli(o, p);
        o->M_uniform -= (int)arrivals;
    }
    if (o->M_beta > 0) {
        double t_s = (double)(o->gen_t % NB_T_UNIFORM) / (double)NB_T_BETA;
        double t_e = t_s + (double)t_elapsed / (double)NB_T_BETA;
        double intensity = (double)o->M_beta * (nb_beta34_cdf(t_e) - nb_beta34_cdf(t_s));
        double ip; double p = modf(intensity, &ip);
        beta_arrivals = ip + (double)rng_bernoulli(o, p);
        o->M_beta -= (int)beta_arrivals;
    }
    o->gen_t = t;
    *arrivals_out = (int)arrivals; *beta_out = (int)beta_arrivals;
}

static void gen_departure(Oracle *o, UE *ue)
{
    if (o->conf.random_traffic) { if (rng_bernoulli(o, o->gen_p) > 0) o->M_beta++; else o->M_uniform++; }
    else if (ue->beta && o->M_beta < o->M_beta_max) o->M_beta++;
    else o->M_uniform++;
}

/* ------------------------------------------------------------------ channel (channel.py) */
static double find_y_value(double x1, double y1, double x2, double y2, double x)
{ double m = (y2 - y1) / (x2 - x1); double b = -m * x1 + y1; return m * x + b; }

#define S34 0.4330127018922193   /* np.sqrt(3)/4 */
#define S32 0.8660254037844386   /* np.sqrt(3)/2 */

static void channel_set_xy_loss_at(Oracle *o, UE *ue, double x, double y)
{
    if (x < 0) for (;;) {                                    /* generate_xy :108-114 (also when a clustered x falls left of the cell, :143-144) */
        nb_rng_pair(o->seed, o->draws++, NBIOT_STREAM_ENV, &x, &y);
        int in_cell = (y > find_y_value(0, S34, 0.25, 0, x)) && (y > find_y_value(0.75, 0, 1, S34, x)) &&
                      (y < find_y_value(0, S34, 0.25, S32, x)) && (y < find_y_value(0.75, S32, 1, S34, x)) && (y < S32);
        if (in_cell) break;
    }
    double x_t = x - 0.5 / 2;                                /* location :100-106 */
    double distance = NB_SQRT(x_t * x_t + y * y);
    double cos_theta = x_t / distance;
    double theta = nb_acos(cos_theta) * NB_RAD2DEG - 60;
    double R = distance * o->range; if (!(R > 0.1)) R = 0.1; /* max(d*range, 0.1) */
    double q = theta / 65; double ap = 12 * (q * q); if (!(ap < 20)) ap = 20;
    double G = NB_GMAX + -1 * ap;
    double L = 128.1 + 37.6 * nb_log10(R);
    if (!(L > 0)) L = 0;
    L = L - G + 0;
    ue->loss = L; ue->x = x; ue->y = y;
}
static void channel_set_xy_loss(Oracle *o, UE *ue) { channel_set_xy_loss_at(o, ue, -1, -1); }

/* Population.clustered_ue_generation (population.py:75-101) */
static void clustered_ue_generation(Oracle *o, UE *ue)
{
    double p = rng_u01(o);
    if (p < o->conf.cluster_prob) {
        if (!o->cluster_count) {
            channel_set_xy_loss(o, ue);
            o->cluster_center[0] = ue->x; o->cluster_center[1] = ue->y;
            o->cluster_count = rng_integers(o, o->conf.cluster_lo, o->conf.cluster_hi);
        } else {
            double ua = rng_u01(o);                          /* angle = uniform(0, 2 pi): cos / sin taken at 2 pi u */
            double radius = o->cluster_range * NB_SQRT(rng_uniform(o, 0, 1));
            double x = radius * nb_cos2pi(ua), y = radius * nb_sin2pi(ua);
            x = x + o->cluster_center[0]; y = y + o->cluster_center[1];
            channel_set_xy_loss_at(o, ue, x, y);
            o->cluster_count -= 1;
        }
    } else channel_set_xy_loss(o, ue);
}

static double get_bler(Oracle *o, double snr, int itbs, int nr)
{
    if (!(snr > -30)) snr = -30;                             /* min(max(snr,-30),20) */
    if (!(snr < 20)) snr = 20;
    int idx = (int)rint(snr * 10.0) + 300;                   /* numpy round(snr,1) = rint(snr*10)/10 */
    int row = nb_itbs_to_row[itbs], rep = 0;
    while ((1 << rep) < nr) rep++;
    return (double)o->lut2[(row * 8 + rep) * NB_LUT_SNR_PTS + idx] / NB_LUT_SCALE;
}

static int rep_index(int N_rep) { int r = 0; while ((1 << r) < N_rep) r++; return r; }

static int preamble_detection_single(Oracle *o, UE *ue, int N_rep)   /* channel.py:156-176 */
{
    double LogF = rng_normal(o, 0, NB_SIGMA_F);
    double rx_pw = NB_TX_PW - ue->loss - LogF;
    double SINR = rx_pw - NB_NOISE_DBM - NB_NOISE_FIG;
    int r = rep_index(N_rep);
    double p = 1 / (1 + nb_exp(-nb_nprach_k[r] * (SINR - nb_nprach_x0[r])));
    int detected = rng_bernoulli(o, p);
    if (detected) ue->state = NB_UE_RAR;
    return detected;
}

static double sinr_estimation(Oracle *o, UE **ues, int n, int *max_i)   /* channel.py:178-198 */
{
    double sum = 0, max_rx = -INFINITY;
    *max_i = 0;
    for (int i = 0; i < n; ++i) {
        double LogF = rng_normal(o, 0, NB_SIGMA_F);
        double rx_pw = NB_TX_PW - ues[i]->loss - LogF;
        rx_pw = nb_exp10(rx_pw / 10);
        sum = sum + rx_pw;                                     /* Python sum(pw): left to right from 0 */
        if (rx_pw > max_rx) { max_rx = rx_pw; *max_i = i; }
    }
    double ni = NB_NOISE_MW + sum - max_rx;
    return 10 * nb_log10(max_rx / ni);
}

/* Channel.multiple_preamble_detection (channel.py:200-219): the strongest of the colliding signals may still be detected */
static int multiple_preamble_detection(Oracle *o, UE **ues, int n, int N_rep)
{
    int mi;
    double SINR = sinr_estimation(o, ues, n, &mi);
    int r = rep_index(N_rep);
    double p = 1 / (1 + nb_exp(-nb_nprach_k[r] * (SINR - nb_nprach_x0[r])));
    int detected = rng_bernoulli(o, p);
    for (int i = 0; i < n; ++i) ues[i]->state = detected ? NB_UE_CAPTURE : NB_UE_COLLIDED;
    return detected;
}

static double sinr_estimation_single(Oracle *o, UE *ue)      /* channel.py:178-198, one contender */
{
    double LogF = rng_normal(o, 0, NB_SIGMA_F);
    double rx_pw = NB_TX_PW - ue->loss - LogF;
    rx_pw = nb_exp10(rx_pw / 10);
    double max_rx = rx_pw;
    double ni = NB_NOISE_MW + (0 + rx_pw) - max_rx;
    double sinr = max_rx / ni;
    return 10 * nb_log10(sinr);
}

/* ------------------------------------------------------------------ carrier (carrier.py) */
static void cols_grow(Oracle *o, int need)
{
    int cap = o->col_cap;
    while (cap < need) cap *= 2;
    if (cap == o->col_cap) return;
    for (int c = 0; c < o->n_carriers; ++c) {
        uint16_t *n = calloc((size_t)cap, sizeof(uint16_t));
        for (int t = o->t_now; t < o->t_ahead; ++t) n[t & (cap - 1)] = o->cols[c][t & (o->col_cap - 1)];
        free(o->cols[c]); o->cols[c] = n;
    }
    o->col_cap = cap;
}

static int nprach_sample(NprachResource *r, int t, uint16_t *mask)   /* carrier.py:301-332 */
{
    *mask = 0;
    if (!r->N_sc) return 0;
    t -= r->t_offset;
    if (r->sf_to_go > 0) { r->sf_to_go -= 1; *mask = r->nprach_sf; return 0; }
    if (t > r->t_next) {
        r->t_next = r->t_start + r->periodicity;
        while (r->t_next < t) r->t_next += r->periodicity;
    }
    if (t == r->t_next || !t) {
        r->nprach_sf = (uint16_t)(((1u << r->N_sc) - 1u) << r->sc_offset) & 0xFFFu;
        r->t_start = t;
        r->sf_to_go = r->N_sf - 1;
        r->t_next = r->t_start + r->periodicity;
        *mask = r->nprach_sf;
        return 1;
    }
    return 0;
}

static int nsf_to_nrep(int N_sf)
{ for (int i = 0; i < 8; ++i) if (nb_nprach_nsf_list[i] == N_sf) return 1 << i; return -1; }

static void sc_log_register(Oracle *o, int t, int CE, int sc)         /* carrier.py:262-265 */
{
    for (int i = o->sc_log.n - 1; i >= 0; --i) if (o->sc_log.d[i].t == t) return;
    ScLogEntry e = { t, { 0, 0, 0 } }; e.sc[CE] += sc;
    VPUSH(o->sc_log, e);
}
static int sc_log_check(Oracle *o, int t, int CE)                     /* carrier.py:267-271, :454-459 */
{
    if (o->n_carriers == 1) return 0;
    if (t == 0) return 0;                                    /* log starts as {0:[0,0,0]} */
    for (int i = o->sc_log.n - 1; i >= 0; --i) if (o->sc_log.d[i].t == t) return o->sc_log.d[i].sc[CE];
    return 0;
}

static uint16_t sfgen_next(Oracle *o, int carrier, int t)             /* carrier.py:382-404 */
{
    uint16_t sf = 0;
    for (int CE = 0; CE < 3; ++CE) {
        NprachResource *r = &o->res[carrier][CE];
        uint16_t m; int ev = nprach_sample(r, t, &m);
        sf |= m;
        if (ev) {
            int N_rep = nsf_to_nrep(r->N_sf), N_pream = 4 * r->N_sc;
            if (carrier == 0) {
                Event s = make_event(EV_NPRACH_START);
                s.CE_level = CE; s.carrier = 0; s.N_pream = N_pream; s.N_rep = N_rep; s.N_sf = r->N_sf; s.N_sc = r->N_sc;
                Event e = make_event(EV_NPRACH_END); e.CE_level = CE; e.carrier = 0;
                schedule_event(o, t, s);
                schedule_event(o, t + r->N_sf, e);
            } else sc_log_register(o, t, CE, N_pream);
        }
    }
    return sf;
}

static void extend_lookahead(Oracle *o, int new_t_ahead)              /* carrier.py:473-479 */
{
    if (new_t_ahead - o->t_now > o->col_cap) cols_grow(o, new_t_ahead - o->t_now);
    while (o->t_ahead < new_t_ahead) {
        for (int c = 0; c < o->n_carriers; ++c)
            o->cols[c][o->t_ahead & (o->col_cap - 1)] = sfgen_next(o, c, o->t_ahead);
        o->t_ahead += 1;
    }
}

static void erase_past_sf(Oracle *o, int t)                           /* carrier.py:481-489 */
{
    if (t > o->t_now) o->t_now = t;     /* columns before t_now are dropped by the ring */
    extend_lookahead(o, t + NB_HORIZON);
}

static void carrier_reset(Oracle *o)                                  /* carrier.py:430-444 */
{
    o->t_now = o->t_ahead = 0;
    VCLEAR(o->sc_log);
    for (int c = 0; c < 2; ++c)
        for (int ce = 0; ce < 3; ++ce) {
            NprachResource *r = &o->res[c][ce];
            memset(r, 0, sizeof *r);
            r->periodicity = 2; r->N_sf = ce; r->N_sc = 0; r->t_next = 2;   /* default_conf_0..2 :49-51 */
        }
}

static int carrier_allocate(Oracle *o, int carrier_id, int t, int delay, int N_sc, int N_sf)   /* carrier.py:556-605 */
{
    int carrier_list[2], ncl = 0;
    carrier_list[ncl++] = carrier_id;
    for (int i = 0; i < o->n_carriers; ++i) if (i != carrier_id) carrier_list[ncl++] = i;
    int t_ul_end = 0;
    erase_past_sf(o, t);
    int step = N_sc;                       /* offsets_per_sc[N_sc] = range(0, 12, N_sc) :29-34 */
    for (int d = delay; d <= 64; d *= 2) { /* possible_delays[delay] :36-41 */
        int new_t_ahead = t + d + N_sf;
        extend_lookahead(o, new_t_ahead);
        for (int ci = 0; ci < ncl && !t_ul_end; ++ci) {
            int c = carrier_list[ci];
            for (int off = 0; off < 12; off += step) {
                uint16_t m = (uint16_t)(((1u << N_sc) - 1u) << off);
                int busy = 0;
                for (int k = 0; k < N_sf && !busy; ++k) busy = o->cols[c][(o->t_now + d + k) & (o->col_cap - 1)] & m;
                if (!busy) {
                    t_ul_end = new_t_ahead;
                    for (int k = 0; k < N_sf; ++k) o->cols[c][(o->t_now + d + k) & (o->col_cap - 1)] |= m;
                    break;
                }
            }
        }
        if (t_ul_end) break;
    }
    return t_ul_end;
}

/* compute_offsets (carrier.py:115-252); returns validity flag */
static int compute_offsets(const int periods[3], const int N_sfs[3], const int N_scs[3], int sc_offsets[3], int sf_offsets[3])
{
    for (int i = 0; i < 3; ++i) sc_offsets[i] = sf_offsets[i] = 0;
    int min_length = N_sfs[0] + N_sfs[1] + N_sfs[2];
    int crit[3], ncrit = 0, norm[3], nnorm = 0;
    for (int i = 0; i < 3; ++i) { if (periods[i] == 240) crit[ncrit++] = i; else norm[nnorm++] = i; }
    int max_sf = N_sfs[0]; if (N_sfs[1] > max_sf) max_sf = N_sfs[1]; if (N_sfs[2] > max_sf) max_sf = N_sfs[2];
    int sum_sc = N_scs[0] + N_scs[1] + N_scs[2];
#define MAX2(a, b) ((a) > (b) ? (a) : (b))
    if (ncrit == 1 && max_sf > 80) {
        int i = crit[0], j = norm[0], k = norm[1];
        if (N_scs[i] + MAX2(N_scs[j], N_scs[k]) > 12) return 0;
        if (sum_sc > 12) {
            min_length = MAX2(N_sfs[j] + N_sfs[k], N_sfs[i]);
            sc_offsets[i] = 0; sc_offsets[j] = N_scs[i]; sc_offsets[k] = N_scs[i];
            sf_offsets[i] = 0; sf_offsets[j] = 0; sf_offsets[k] = N_sfs[j];
        } else {
            min_length = N_sfs[2];
            sc_offsets[i] = 0; sc_offsets[j] = N_scs[i]; sc_offsets[k] = N_scs[i] + N_scs[j];
        }
    } else if (ncrit == 2 && max_sf > 80) {
        int i = crit[0], j = crit[1], k = norm[0];
        if (N_scs[k] + MAX2(N_scs[i], N_scs[j]) > 12) return 0;
        if (sum_sc > 12) {
            min_length = MAX2(N_sfs[i] + N_sfs[j], N_sfs[k]);
            sc_offsets[i] = 0; sc_offsets[j] = 0; sc_offsets[k] = MAX2(N_scs[i], N_scs[j]);
            sf_offsets[i] = 0; sf_offsets[j] = N_sfs[i]; sf_offsets[k] = 0;
        } else {
            min_length = N_sfs[2];
            sc_offsets[i] = 0; sc_offsets[j] = N_scs[i]; sc_offsets[k] = N_scs[i] + N_scs[j];
        }
    } else if (sum_sc > 12) {
        if (N_scs[2] + MAX2(N_scs[1], N_scs[0]) <= 12) {
            min_length = N_sfs[2] + N_sfs[0];
            sc_offsets[0] = N_scs[2]; sc_offsets[1] = N_scs[2]; sc_offsets[2] = 0;
            sf_offsets[0] = N_sfs[1]; sf_offsets[1] = 0; sf_offsets[2] = 0;
        } else if (N_scs[2] + N_scs[1] <= 12) {
            min_length = N_sfs[2] + N_sfs[0];
            sc_offsets[0] = 0; sc_offsets[1] = N_scs[2]; sc_offsets[2] = 0;
            sf_offsets[0] = N_sfs[2]; sf_offsets[1] = 0; sf_offsets[2] = 0;
        } else if (N_scs[1] + N_scs[0] <= 12) {
            min_length = N_sfs[2] + N_sfs[1];
            sc_offsets[0] = N_scs[1]; sc_offsets[1] = 0; sc_offsets[2] = 0;
            sf_offsets[0] = N_sfs[2]; sf_offsets[1] = N_sfs[2]; sf_offsets[2] = 0;
        } else if (N_scs[2] + N_scs[0] <= 12) {
            min_length = N_sfs[2] + N_sfs[1];
            sc_offsets[0] = N_scs[2]; sc_offsets[1] = 0; sc_offsets[2] = 0;
            sf_offsets[0] = 0; sf_offsets[1] = N_sfs[2]; sf_offsets[2] = 0;
        } else {
            /* sort_indexes: stable sort of 0,1,2 by (period, N_sc, N_sf) :90-113 */
            int a[3] = { 0, 1, 2 };
            for (int x = 1; x < 3; ++x)
                for (int y = x; y > 0; --y) {
                    int p = a[y - 1], q = a[y];
                    int gt = periods[p] != periods[q] ? periods[p] > periods[q]
                           : N_scs[p] != N_scs[q] ? N_scs[p] > N_scs[q] : N_sfs[p] > N_sfs[q];
                    if (gt) { a[y - 1] = q; a[y] = p; } else break;
                }
            int i = a[0], j = a[1], k = a[2];
            sf_offsets[i] = 0; sf_offsets[j] = N_sfs[i]; sf_offsets[k] = N_sfs[i] + N_sfs[j];
        }
    } else {
        min_length = MAX2(N_sfs[2], N_sfs[1] + N_sfs[0]);
        sc_offsets[0] = N_scs[2]; sc_offsets[1] = N_scs[2]; sc_offsets[2] = 0;
        if (N_scs[0] <= N_scs[1]) { sf_offsets[0] = N_sfs[1]; sf_offsets[1] = 0; sf_offsets[2] = 0; }
        else { sf_offsets[0] = 0; sf_offsets[1] = N_sfs[0]; sf_offsets[2] = 0; }
    }
#undef MAX2
    int min_periodicity = -1;                                  /* minpass :59-65 */
    for (int i = 0; i < 7; ++i)
        if (nb_nprach_period_list[i] >= min_length && (min_periodicity < 0 || nb_nprach_period_list[i] < min_periodicity))
            min_periodicity = nb_nprach_period_list[i];
    int minp = periods[0]; if (periods[1] < minp) minp = periods[1]; if (periods[2] < minp) minp = periods[2];
    if (min_periodicity < 0 || minp < min_periodicity) return 0;
    return 1;
}

static void carrier_update_ra_parameters(Oracle *o, int args[3][3], int th_C1, int th_C0)   /* carrier.py:610-650 */
{
    int n_c = o->n_carriers;
    int periods[3], N_sfs[3], N_scs[3];
    static const int nsc_tab[2][8] = { { 12, 24, 36, 48, 0, 0, 0, 0 }, { 12, 24, 36, 48, 60, 72, 84, 96 } };
    for (int i = 0; i < 3; ++i) {
        periods[i] = nb_nprach_period_list[args[i][0]];
        N_sfs[i] = nb_nprach_nsf_list[args[i][1]];
        N_scs[i] = nsc_tab[n_c - 1][args[i][2]] / 4;
    }
    if (th_C0 == 0) { periods[0] = nb_nprach_period_list[0]; N_sfs[0] = 0; N_scs[0] = 0; }
    if (th_C1 == 0) { periods[1] = nb_nprach_period_list[0]; N_sfs[1] = 0; N_scs[1] = 0; }
    int N_scs_list[2][3];
    for (int c = 0; c < n_c; ++c)
        for (int i = 0; i < 3; ++i) {
            int take = N_scs[i] < 12 ? N_scs[i] : 12;
            N_scs_list[c][i] = take;
            N_scs[i] = N_scs[i] - take > 0 ? N_scs[i] - take : 0;
        }
    int sc_off[3], sf_off[3];
    (void)compute_offsets(periods, N_sfs, N_scs_list[0], sc_off, sf_off);
    for (int c = 0; c < n_c; ++c)
        for (int ce = 0; ce < 3; ++ce) {
            /* SFGenerator.update :377-380 and NprachResource.update :345-355 */
            int rt = o->res[c][0].t_start;
            if (o->res[c][1].t_start > rt) rt = o->res[c][1].t_start;
            if (o->res[c][2].t_start > rt) rt = o->res[c][2].t_start;
            NprachResource *r = &o->res[c][ce];
            r->periodicity = periods[ce]; r->N_sf = N_sfs[ce]; r->N_sc = N_scs_list[c][ce]; r->sc_offset = sc_off[ce];
            r->t_next = r->t_start + periods[ce];
            if (r->t_next <= rt) r->t_next = rt + periods[ce];
            r->t_offset = sf_off[ce];
        }
}

/* ------------------------------------------------------------------ NodeB helpers */
static int ivec_sum(const ivec *v) { int s = 0; for (int i = 0; i < v->n; ++i) s += v->d[i]; return s; }

static void update_state(Oracle *o)                                   /* node_b.py:264-276 */
{
    int s = ivec_sum(&o->RAR_UEs[0]) + ivec_sum(&o->RAR_UEs[1]) + ivec_sum(&o->RAR_UEs[2]);
    if (s) o->state = NB_ST_RAR_WINDOW;
    else if (o->connected.n > 0) o->state = NB_ST_SCHEDULING;
    else o->state = NB_ST_IDLE;
}

static void schedule_NPDCCH_arrival(Oracle *o, int elapsed_sfs)       /* node_b.py:420-432 */
{
    o->NPDCCH_sf_left = o->NPDCCH_sf_left - elapsed_sfs > 0 ? o->NPDCCH_sf_left - elapsed_sfs : 0;
    int next_time = o->time + elapsed_sfs;
    if (!o->NPDCCH_sf_left) { next_time += NB_NPDCCH_PERIOD; o->NPDCCH_sf_left = NB_NPDCCH_SF; }
    schedule_event(o, next_time, o->event);
}

static int sf_per_ru(int N_sc) { return N_sc == 12 ? 1 : N_sc == 6 ? 2 : N_sc == 3 ? 4 : N_sc == 1 ? 8 : -1; }

/* NodeB.allocate_resources (node_b.py:435-450) */
static int node_allocate(Oracle *o, int c_id, int ref_t, int delay, int N_sc, int N_ru, int N_rep, int *out_sc, int *out_sf)
{
    int spr = sf_per_ru(N_sc);
    if (spr < 0) { o->fault |= NBIOT_FAULT_ACTION; *out_sc = N_sc; *out_sf = 0; return 0; }
    int n_sc = N_sc, N_sf = spr * N_ru * N_rep;
    int t_ul_end = carrier_allocate(o, c_id, ref_t, delay, N_sc, N_sf);
    if (!t_ul_end && o->conf.sc_adjustment) {
        int row = N_sc == 1 ? 0 : N_sc == 3 ? 1 : N_sc == 6 ? 2 : 3;
        for (int i = 0; i < 3; ++i) {
            n_sc = nb_sc_fallback[row][i];
            N_sf = sf_per_ru(n_sc) * N_ru * N_rep;
            t_ul_end = carrier_allocate(o, c_id, ref_t, delay, n_sc, N_sf);
            if (t_ul_end) break;
        }
    }
    *out_sc = n_sc; *out_sf = N_sf;
    return t_ul_end;
}

/* ------------------------------------------------------------------ population */
static void uvec_remove(uvec *v, UE *ue) { for (int i = 0; i < v->n; ++i) if (v->d[i] == ue) { VDEL(*v, i); return; } }

static void get_new_arrivals(Oracle *o, int t)                        /* population.py:103-132 */
{
    int arrivals, beta_arrivals;
    gen_arrivals(o, t, &arrivals, &beta_arrivals);
    arrivals += beta_arrivals;
    for (int n = 0; n < arrivals; ++n) {
        o->ue_count += 1;
        int buffer = o->conf.buffer_hi > o->conf.buffer_lo ? rng_integers(o, o->conf.buffer_lo, o->conf.buffer_hi + 1)
                                                           : o->conf.buffer_lo;     /* generate_buffer */
        UE *ue = calloc(1, sizeof(UE));
        ue->id = o->ue_count; ue->t_arrival = t; ue->buffer = buffer; ue->new_data = 1; ue->loss = 70;
        if (beta_arrivals > 0) { ue->beta = 1; beta_arrivals -= 1; }
        if (o->conf.clustered) clustered_ue_generation(o, ue);
        else channel_set_xy_loss(o, ue);
        double P_RSRP = NB_SRP - ue->loss;
        int CE_level;
        if (P_RSRP < o->CE_thresholds[0]) CE_level = 2;
        else if (P_RSRP < o->CE_thresholds[1]) CE_level = 1;
        else CE_level = 0;
        ue->CE_level = CE_level;
        VPUSH(o->ra_lists[CE_level], ue);
        o->ue_arrivals[CE_level] += 1;
    }
}

static void population_msg3_grant(Oracle *o, int CE_level, int t, int t_msg3_end, int I_tbs, int N_rep)   /* population.py:201-256 */
{
    uvec *l = &o->ra_lists[CE_level];
    uvec cl = { 0 };                                           /* contention_list */
    int preamble = -1;
    for (int i = 0; i < l->n; ++i) {
        UE *ue = l->d[i];
        if (!(ue->state == NB_UE_CAPTURE || ue->state == NB_UE_RAR)) continue;
        if (cl.n == 0) {
            preamble = ue->preamble;
            ue->I_tbs = I_tbs; ue->N_rep = N_rep; ue->t_contention = t;
            VPUSH(cl, ue);
            if (ue->state == NB_UE_RAR) { ue->state = NB_UE_CONREQ; break; }      /* did not collide: the only contender */
        } else if (ue->preamble == preamble) {                 /* captured preamble: everybody who sent it answers the grant */
            ue->I_tbs = I_tbs; ue->N_rep = N_rep; ue->t_contention = t;
            VPUSH(cl, ue);
        }
    }
    if (!cl.n) return;
    for (int i = 0; i < cl.n; ++i) { uvec_remove(l, cl.d[i]); VPUSH(o->contention, cl.d[i]); }
    Event m = make_event(EV_MSG3); m.ue = cl.d[0]; m.ue_id = cl.d[0]->id;
    Event tm = make_event(EV_MAC_TIMER); tm.ue_id = cl.d[0]->id;
    if (cl.n > 1) {
        m.ue_list = malloc(sizeof(UE *) * (size_t)cl.n); tm.id_list = malloc(sizeof(int) * (size_t)cl.n); m.n_list = tm.n_list = cl.n;
        for (int i = 0; i < cl.n; ++i) { m.ue_list[i] = cl.d[i]; tm.id_list[i] = cl.d[i]->id; }
    }
    schedule_event(o, t_msg3_end, m);
    int t_exp = t + o->MAC_timer > t_msg3_end + 1 ? t + o->MAC_timer : t_msg3_end + 1;
    schedule_event(o, t_exp, tm);
    VFREE(cl);
}

static void backoff_end(UE *ue) { ue->state = NB_UE_RAO; ue->backoff_counter += 1; }   /* population.py:143-149 */

static void population_RAR_window_end(Oracle *o, int t, int CE_level, int backoff, int rar_w_id)   /* population.py:258-306 */
{
    uvec *l = &o->ra_lists[CE_level];
    for (int i = 0; i < l->n; ++i) {
        UE *ue = l->d[i];
        if (ue->state == NB_UE_RAO || ue->state == NB_UE_BACKOFF || ue->rar_w_id != rar_w_id) continue;
        ue->state = NB_UE_BACKOFF;
        ue->ra_attempts += 1;
        if (ue->ra_attempts > o->preamble_trans_max && ue->CE_level < 2) { ue->ra_attempts = 0; ue->CE_level += 1; }
        o->backoffs[CE_level] += 1;
        if (backoff == 0) backoff_end(ue);
        else {
            int backoff_time = rng_integers(o, 0, backoff);
            if (backoff_time == 0) backoff_end(ue);
            else { Event b = make_event(EV_BACKOFF_END); b.ue = ue; schedule_event(o, t + backoff_time, b); }
        }
    }
}

static void population_MAC_timeout(Oracle *o, Event *ev)              /* population.py:151-163 */
{
    int n = ev->n_list > 1 ? ev->n_list : 1;
    for (int k = 0; k < n; ++k) {                              /* ue_id_list order; ids that already connected are skipped */
        int id = ev->n_list > 1 ? ev->id_list[k] : ev->ue_id;
        for (int i = 0; i < o->contention.n; ++i) {
            UE *ue = o->contention.d[i];
            if (ue->id == id) {
                VDEL(o->contention, i);
                ue->state = NB_UE_RAO; ue->timeout_counter += 1;
                VPUSH(o->ra_lists[ue->CE_level], ue);
                break;
            }
        }
    }
    free(ev->id_list); ev->id_list = NULL;
}

/* PerfMonitor's sample lists, in append order (perf_monitor.py:61-79, :111-140, :185-202); which lists are kept: conf.trace_mask */
static void perf_sample(Oracle *o, int type, int ce, int t, int a, int b)
{
    if (!(o->conf.trace_cap > 0 && (o->conf.trace_mask & type))) return;
    nbiot_trace_rec_t r; r.t = t; r.a = a; r.b = b; r.type = (uint8_t)type; r.ce = (uint8_t)ce; r.pad = 0;
    VPUSH(o->samples, r);
}

/* ------------------------------------------------------------------ access procedure */
static void access_NPRACH_start(Oracle *o, Event *ev)                 /* access_procedure.py:37-91 */
{
    int t = ev->t, CE_level = ev->CE_level, N_pream = ev->N_pream, N_sc = ev->N_sc, N_sf = ev->N_sf, N_rep = ev->N_rep;
    int counter = o->RAR_counter[CE_level];
    double p = o->probability_anchor;
    o->NPRACH_resources += (long long)N_sc * N_sf;
    perf_sample(o, NBIOT_TR_RES_NPRACH, CE_level, t, N_sc * N_sf, 0);
    /* population.NPRACH_start :168-188 */
    if (o->gen_t < t) get_new_arrivals(o, t);
    uvec *l = &o->ra_lists[CE_level];
    uvec ra = { 0 };
    for (int i = 0; i < l->n; ++i) {
        UE *ue = l->d[i];
        if (ue->state == NB_UE_RAO) { ue->rar_w_id = counter; ue->state = NB_UE_RA; VPUSH(ra, ue); }
    }
    int n_ues = ra.n;
    o->contenders_per_CE[CE_level] = n_ues;
    if (n_ues > 0) {
        int na_ues = 0;
        int Na_sc = sc_log_check(o, t, CE_level);
        int k = 0;
        if (Na_sc) {
            na_ues = rng_binomial(o, n_ues, p);
            for (int i = 0; i < na_ues; ++i) ra.d[k++]->preamble = rng_integers(o, N_pream, N_pream + 4 * Na_sc);
            o->NPRACH_resources += (long long)Na_sc * N_sf;
            perf_sample(o, NBIOT_TR_RES_NPRACH, CE_level, t, Na_sc * N_sf, 0);
        }
        int a_ues = n_ues - na_ues;
        for (int i = 0; i < a_ues; ++i) ra.d[k++]->preamble = rng_integers(o, 0, N_pream);
        /* group by preamble in first-occurrence order; singletons are tested, collisions are certain failures */
        int detected_preambles = 0, detected_signals = 0;
        for (int i = 0; i < n_ues; ++i) {
            int pre = ra.d[i]->preamble, first = 1, cnt = 0;
            for (int j = 0; j < i; ++j) if (ra.d[j]->preamble == pre) { first = 0; break; }
            if (!first) continue;
            for (int j = i; j < n_ues; ++j) if (ra.d[j]->preamble == pre) cnt++;
            if (cnt > 1 && o->conf.capture_effect) {           /* channel.py:158-159 */
                UE **grp = malloc(sizeof(UE *) * (size_t)cnt); int g = 0;
                for (int j = i; j < n_ues; ++j) if (ra.d[j]->preamble == pre) grp[g++] = ra.d[j];
                detected_preambles += multiple_preamble_detection(o, grp, cnt, N_rep); detected_signals += 1;
                free(grp);
            }
            else if (cnt > 1) { detected_signals += 1; }
            else { int d = preamble_detection_single(o, ra.d[i], N_rep); detected_preambles += d; detected_signals += d; }
        }
        int collided = detected_signals - detected_preambles;
        o->detected_preambles_per_CE[CE_level] = detected_preambles;
        o->collided_preambles_per_CE[CE_level] = collided;
        VPUSH(o->detection_history[CE_level], detected_preambles);
        VPUSH(o->collision_history[CE_level], collided);
    }
    VFREE(ra);
}

static void node_start_RAR_window(Oracle *o, int detections, int CE_level, int time, int rar_w_id)   /* node_b.py:524-554 */
{
    VPUSH(o->RAR_UEs[CE_level], detections);
    VPUSH(o->RAR_w_ids[CE_level], rar_w_id);
    o->state = NB_ST_RAR_WINDOW;
    int RAR_in = o->RAR_in[CE_level], RAR_detections = o->RAR_ids[CE_level], RAR_sent = o->RAR_sent[CE_level];
    VPUSH(o->msg3_sent[CE_level], RAR_sent);
    VPUSH(o->msg3_received[CE_level], RAR_detections);
    VPUSH(o->msg3_efficiency[CE_level], (double)RAR_detections / (double)(RAR_in > 1 ? RAR_in : 1));
    perf_sample(o, NBIOT_TR_RARWIN, CE_level, time, RAR_in | (o->RAR_attmp[CE_level] << 16), RAR_sent | (RAR_detections << 16));   /* rar_window_sample :543 */
    o->RAR_in[CE_level] = detections; o->RAR_ids[CE_level] = 0; o->RAR_attmp[CE_level] = 0; o->RAR_sent[CE_level] = 0;
    int e_time = time + o->RAR_window_size[CE_level];
    Event e = make_event(EV_RAR_WINDOW_END); e.CE_level = CE_level; e.rar_w_id = rar_w_id;
    schedule_event(o, e_time, e);
}

static void access_NPRACH_end(Oracle *o, Event *ev)                   /* access_procedure.py:93-116 */
{
    int CE_level = ev->CE_level, t = ev->t, counter = o->RAR_counter[CE_level];
    int n_ues = o->contenders_per_CE[CE_level];
    int detected = o->detected_preambles_per_CE[CE_level];
    perf_sample(o, NBIOT_TR_NPRACH, CE_level, t, n_ues, detected);       /* perf_monitor.nprach_sample :105 */
    if (n_ues > 0) {
        node_start_RAR_window(o, detected, CE_level, t, counter);
        o->detected_preambles_per_CE[CE_level] = 0; o->collided_preambles_per_CE[CE_level] = 0; o->contenders_per_CE[CE_level] = 0;
    }
    o->RAR_counter[CE_level] += 1;
}

/* ------------------------------------------------------------------ NodeB reception callbacks */
static void distribution_update(Oracle *o, double sample)             /* node_b.py:207-225 */
{
    if (o->k > 1000) return;
    o->k += 1;
    int idx = 14;
    for (int i = 0; i < 14; ++i) if (sample + 35 < nb_threshold_list[i]) { idx = i; break; }
    for (int i = 0; i < 15; ++i) {
        double prob = o->distribution[i];
        if (i == idx) o->distribution[i] += (1 - prob) / o->k;
        else o->distribution[i] += (0 - prob) / o->k;
    }
}

static void node_msg3_outcome(Oracle *o, UE *ue, int t, int connected)   /* node_b.py:770-799 */
{
    int w_id = ue->rar_w_id;
    ivec *wl = &o->RAR_w_ids[ue->CE_level];
    int i_ = 0, window_active = 0;
    for (int i = 0; i < wl->n; ++i) if (wl->d[i] == w_id) { window_active = 1; i_ = i; break; }
    if (connected && ue->state != NB_UE_RAO) {
        ue->state = NB_UE_CONNECTED;
        ue->t_connection = t;
        uvec_remove(&o->contention, ue);                      /* population.msg4 */
        VPUSH(o->connected, ue);
        ue->timestamp = ue->t_connection;                     /* perf.register_ue :224-229 */
        dict_set(&o->acc_id, &o->acc_delay, ue->id, ue->t_connection - ue->t_arrival);
        o->RAR_ids[ue->CE_level] += 1;
        o->RAR_detected[ue->CE_level] += 1;
        double sample = -1 * ue->loss;
        VPUSH(o->incoming, sample);
        distribution_update(o, sample);
    } else if (window_active) {
        ue->state = NB_UE_RAR;
        o->RAR_UEs[ue->CE_level].d[i_] += 1;
    }
    update_state(o);
}

static void rx_msg3_event(Oracle *o, Event *ev)                       /* rx_procedure.py:29-51, channel.py:221-234 */
{
    UE *ue = ev->ue;
    int contenders = ev->n_list > 1 ? ev->n_list : 1;
    double SINR;
    int I_tbs = ue->I_tbs, N_rep = ue->N_rep;                  /* of contention_list[0] (channel.py:226-227) */
    if (contenders > 1) { int mi; SINR = sinr_estimation(o, ev->ue_list, contenders, &mi); ue = ev->ue_list[mi]; }   /* the strongest signal */
    else SINR = sinr_estimation_single(o, ue);
    double bler_256 = get_bler(o, SINR, I_tbs, N_rep);
    double p = nb_pow01(1.0 - bler_256, (double)NB_MSG3_TBS / 256);
    int connected = rng_bernoulli(o, p);
    int ce = ue->CE_level;
    node_msg3_outcome(o, ue, ev->t, connected);
    if (connected) o->ue_connections += 1;                    /* perf.rar_sample(1,0,0,...) */
    /* a failure with one contender is an error, with several a collision (rx_procedure.py:41-51) */
    perf_sample(o, NBIOT_TR_MSG3, ce, ev->t, connected, connected ? 0 : (contenders > 1 ? 2 : 1));
    free(ev->ue_list); ev->ue_list = NULL;
}

static void rx_NPUSCH_end(Oracle *o, Event *ev)                       /* rx_procedure.py:54-61, channel.py:236-256, node_b.py:801-830 */
{
    UE *ue = ev->ue; int t = ev->t;
    double LogF = rng_normal(o, 0, NB_SIGMA_F);
    double rx_pw = NB_TX_PW - ue->loss - LogF;
    double SINR = rx_pw - NB_NOISE_DBM - NB_NOISE_FIG;
    if (!ue->new_data) SINR = 10 * nb_log10(nb_exp10(SINR / 10) + nb_exp10(ue->sinr / 10));
    double p;
    if (ue->tbs != 256) p = nb_pow01(1.0 - get_bler(o, SINR, ue->I_tbs, ue->N_rep), (double)ue->tbs / 256);
    else p = 1.0 - get_bler(o, SINR, ue->I_tbs, ue->N_rep);
    int received = rng_bernoulli(o, p);
    /* NodeB.NPUSCH_end */
    o->n_scheduled -= 1;
    ue->sinr = SINR;
    if (received) {
        int bits = ue->bits;
        ue->acked_data += bits; ue->retx_buffer = 0; ue->bits = 0; ue->new_data = 1;
        dict_set(&o->rx_id, &o->rx_delay, ue->id, t - ue->timestamp);   /* perf.account_rx :231-237 */
        ue->timestamp = t; o->attempts_count += 1;
        if (!ue->buffer) {
            ue->state = NB_UE_DONE; ue->t_disconnection = t;
            int delay = t - ue->t_connection, service_time = t - ue->t_arrival;   /* perf.unregister_ue :142-171 */
            o->ue_departures[ue->CE_level] += 1; o->thr_count += 1; o->total_bits += ue->acked_data;
            o->av_delay += ((double)delay - o->av_delay) / (double)(o->ue_departures[0] + o->ue_departures[1] + o->ue_departures[2]);
            if (o->conf.stat_cap > 0) {                                          /* statistics (perf_monitor.py:161-171) */
                VPUSH(o->access_history, ue->t_connection - ue->t_arrival); VPUSH(o->connection_history, service_time);
                VPUSH(o->delay_history, delay); VPUSH(o->bits_history, ue->acked_data);
            }
            perf_sample(o, NBIOT_TR_DEPART, ue->CE_level, t, 0, 0);            /* npusch_delivered (perf_monitor.py:164-168) */
            gen_departure(o, ue);
            VPUSH(o->departures, delay);
            o->total_departures += 1;
            VPUSH(o->service_times, service_time);
            update_state(o);
            free(ue);
            return;
        }
        VPUSH(o->connected, ue);
    } else {
        ue->new_data = 0; ue->harq_rtx += 1; ue->retx_buffer = ue->bits;
        VPUSH(o->err_id, ue->id); o->error_count += 1; o->attempts_count += 1;   /* perf.account_error */
        VPUSH(o->connected, ue);
    }
    update_state(o);
}

/* ------------------------------------------------------------------ step handlers */
typedef struct { int carrier, i, I_tbs, delay, N_sc, N_rep, N_ru; } SchedParams;

static int scheduling_action(Oracle *o, const int32_t *a, int RAR_action, SchedParams *p)   /* action_reader.py:25-66 */
{
    if (a[NB_A_CARRIER] >= o->n_carriers || a[NB_A_CARRIER] < 0) { o->fault |= NBIOT_FAULT_ACTION; return 0; }   /* IndexError in the reference */
    p->carrier = a[NB_A_CARRIER]; p->i = a[NB_A_ID]; p->N_ru = a[NB_A_NRU];
    p->N_rep = nb_n_rep_list[a[NB_A_NREP]];
    int I_mcs = a[NB_A_IMCS];
    if (RAR_action) {
        if (I_mcs > 2) { o->fault |= NBIOT_FAULT_ACTION; return 0; }     /* IndexError in the reference */
        p->I_tbs = I_mcs < 2 ? I_mcs : 2; p->N_ru = nb_rar_imcs_to_nru[I_mcs];
    } else p->I_tbs = nb_imcs_to_itbs[I_mcs];
    p->delay = nb_sf_delay_list[a[NB_A_DELAY]];
    p->N_sc = nb_n_sc_list[a[NB_A_SC]];
    return 1;
}

static int ilist_cmp(const ivec *a, const ivec *b)   /* Python list comparison */
{
    int n = a->n < b->n ? a->n : b->n;
    for (int i = 0; i < n; ++i) if (a->d[i] != b->d[i]) return a->d[i] < b->d[i] ? -1 : 1;
    return a->n == b->n ? 0 : (a->n < b->n ? -1 : 1);
}

static void step_RAR_window(Oracle *o, const int32_t *a)              /* node_b.py:453-521 */
{
    SchedParams P;
    if (!scheduling_action(o, a, 1, &P)) return;
    int CE_level = P.i, I_tbs = P.I_tbs, N_ru = P.N_ru, N_rep = P.N_rep;
    o->action[0] = CE_level; o->action[1] = I_tbs; o->action[2] = N_rep;
    int RAR_ues[3] = { ivec_sum(&o->RAR_UEs[0]), ivec_sum(&o->RAR_UEs[1]), ivec_sum(&o->RAR_UEs[2]) };
    o->valid_CE_control = 1;
    int valid_control = 1;
    int NPDCCH_sf = nb_npdcch_sf_per_ce[CE_level];
    if (!(RAR_ues[CE_level] > 0) || NPDCCH_sf > o->NPDCCH_sf_left || o->conf.ce_mcs_automatic) {
        /* sorted(enumerate(RAR_UEs), key=list, reverse=True): stable descending */
        int order[3] = { 0, 1, 2 };
        for (int x = 1; x < 3; ++x)
            for (int y = x; y > 0; --y) {
                if (ilist_cmp(&o->RAR_UEs[order[y - 1]], &o->RAR_UEs[order[y]]) < 0) { int tmp = order[y - 1]; order[y - 1] = order[y]; order[y] = tmp; }
                else break;
            }
        valid_control = 0;
        for (int x = 0; x < 3; ++x) {
            CE_level = order[x];
            NPDCCH_sf = nb_npdcch_sf_per_ce[CE_level];
            valid_control = (RAR_ues[CE_level] > 0) && (o->NPDCCH_sf_left >= NPDCCH_sf);
            if (valid_control) { o->valid_CE_control = 0; break; }
        }
    }
    if (!valid_control) { schedule_NPDCCH_arrival(o, o->NPDCCH_sf_left); return; }
    int carrier_id = P.carrier, delay = P.delay, N_sc = P.N_sc;
    if (o->conf.ce_mcs_automatic) { I_tbs = 2; N_ru = 1; N_rep = o->msg3_conf_nrep[CE_level]; }
    o->final_action[0] = CE_level; o->final_action[1] = I_tbs; o->final_action[2] = N_rep;
    int reference_time = o->time + NPDCCH_sf;
    int n_sf, t_ul_end = node_allocate(o, carrier_id, reference_time, delay, N_sc, N_ru, N_rep, &N_sc, &n_sf);
    o->RAR_attmp[CE_level] += 1;
    if (t_ul_end) {
        population_msg3_grant(o, CE_level, reference_time, t_ul_end, I_tbs, N_rep);
        o->RAR_sent[CE_level] += 1;
        ivec *r = &o->RAR_UEs[CE_level];                      /* subtract_one :59-68 */
        for (int i = 0; i < r->n; ++i) if (r->d[i] > 0) { r->d[i] -= 1; break; }
        o->msg3_resources += (long long)N_sc * n_sf;
        perf_sample(o, NBIOT_TR_RES_MSG3, CE_level, o->time, N_sc * n_sf, 0);
    } else o->valid_CE_control = 0;
    schedule_NPDCCH_arrival(o, NPDCCH_sf);
    update_state(o);
}

static void step_RAR_end(Oracle *o, const int32_t *a)                 /* node_b.py:558-581 */
{
    int CE_level = o->event.CE_level, rar_w_id = o->event.rar_w_id, t = o->event.t;
    int backoff = nb_backoff_list[a[NB_A_BACKOFF]];
    int period_i = o->NPRACH_conf[7 + 3 * CE_level];          /* period_C{CE} in NPRACH_items order */
    if (nb_backoff_list[period_i + 1] > backoff) backoff = nb_backoff_list[period_i + 1];
    population_RAR_window_end(o, t, CE_level, backoff, rar_w_id);
    if (o->RAR_UEs[CE_level].n) { o->RAR_failed[CE_level] = o->RAR_UEs[CE_level].d[0]; VDEL(o->RAR_UEs[CE_level], 0); }
    if (o->RAR_w_ids[CE_level].n) VDEL(o->RAR_w_ids[CE_level], 0);
    update_state(o);
}

static int next_tbs(int buffer)   /* utils.find_next_integer(TBS_list, n) :35-43 */
{ for (int i = 0; i < NB_N_TBS; ++i) if (nb_tbs_list[i] >= buffer) return i; return NB_N_TBS - 1; }

static void step_data(Oracle *o, const int32_t *a)                    /* node_b.py:584-693 */
{
    perf_clear_info(o);
    if (o->n_selectable == 0) { schedule_NPDCCH_arrival(o, o->NPDCCH_sf_left); update_state(o); return; }
    SchedParams P;
    if (!scheduling_action(o, a, 0, &P)) return;
    int ue_i = P.i; if (ue_i >= o->n_selectable) ue_i = 0;
    int ue_id = o->selectable[ue_i];
    UE *ue = NULL; int pos = -1;
    for (int i = 0; i < o->connected.n; ++i) if (o->connected.d[i]->id == ue_id) { ue = o->connected.d[i]; pos = i; break; }
    int CE_level = ue->CE_level, new_data = ue->new_data;
    int I_tbs, N_rep, N_ru, tbs;
    if (!new_data) { I_tbs = ue->I_tbs; N_rep = ue->N_rep; N_ru = ue->N_ru; tbs = ue->tbs; }
    else if (o->conf.mcs_automatic) {
        int li = (int)rint(ue->loss * 10.0);                  /* round(loss,1) clipped to [121.4,171.4] */
        if (li < 1214) li = 1214; if (li > 1714) li = 1714;
        if (li > 1713) { o->fault |= NBIOT_FAULT_MCS_KEY; return; }   /* KeyError in the reference (App. C #9) */
        int ti = next_tbs(ue->buffer);
        tbs = nb_tbs_list[ti];
        const uint8_t *e = &o->mcs[((li - 1214) * NB_N_TBS + ti) * 3];
        I_tbs = e[0]; N_rep = e[1]; N_ru = e[2];
    } else {
        N_rep = P.N_rep; I_tbs = P.I_tbs;
        const uint16_t *lst = nb_tbs_table[nb_itbs_to_row[I_tbs]];
        /* find_next_integer_index(lst, buffer): min (value, position) with value >= buffer, else (max, 7) */
        int p = 7; for (int i = 0; i < 8; ++i) if (lst[i] >= ue->buffer) { p = i; break; }
        if (o->conf.tx_all_buffer) { tbs = lst[p]; N_ru = nb_n_ru_list[p]; }     /* itbs_bl never upgrades (App. C #1) */
        else {
            int I_N_ru = P.N_ru;
            if (p < I_N_ru) { N_ru = nb_n_ru_list[p]; tbs = lst[p]; }
            else { N_ru = nb_n_ru_list[I_N_ru]; tbs = lst[I_N_ru]; }
        }
    }
    int carrier_id = P.carrier, delay = P.delay, N_sc = P.N_sc;
    int NPDCCH_sf = nb_npdcch_sf_per_ce[CE_level];
    int reference_time = o->time + NPDCCH_sf;
    int n_sf, t_ul_end = node_allocate(o, carrier_id, reference_time, delay, N_sc, N_ru, N_rep, &N_sc, &n_sf);
    if (t_ul_end) {
        /* data_grant :667-693 */
        VDEL(o->connected, pos);
        ue->I_tbs = I_tbs; ue->N_rep = N_rep; ue->N_ru = N_ru; ue->tbs = tbs;
        if (ue->retx_buffer > 0) ue->bits = ue->retx_buffer;
        else if (tbs >= ue->buffer) { ue->bits = ue->buffer; ue->buffer = 0; }
        else { ue->bits = ue->buffer < tbs ? ue->buffer : tbs; ue->buffer -= ue->bits; }
        o->n_scheduled += 1;
        Event e = make_event(EV_NPUSCH_END); e.ue = ue; e.ue_id = ue->id;
        schedule_event(o, t_ul_end, e);
        o->NPUSCH_resources += (long long)N_sc * n_sf;
        perf_sample(o, NBIOT_TR_RES_NPUSCH, CE_level, o->time, N_sc * n_sf, 0);
    } else if (new_data) VPUSH(o->unfit_id, ue->id);
    schedule_NPDCCH_arrival(o, NPDCCH_sf);
    update_state(o);
}

static void step_update(Oracle *o, const int32_t *a)                  /* node_b.py:695-768 */
{
    static const int item_idx[16] = { NB_A_BACKOFF, NB_A_RAR_WINDOW, NB_A_MAC_TIMER, NB_A_TRANSMAX, NB_A_PANCHOR, NB_A_TH_C1, NB_A_TH_C0,
        NB_A_PERIOD_C0, NB_A_REP_C0, NB_A_SC_C0, NB_A_PERIOD_C1, NB_A_REP_C1, NB_A_SC_C1, NB_A_PERIOD_C2, NB_A_REP_C2, NB_A_SC_C2 };
    o->thr_count = 0;                                          /* clear_nprach_metrics */
    int conf[16];
    for (int i = 0; i < 16; ++i) conf[i] = a[item_idx[i]];
    int th_C1 = nb_threshold_list[conf[5]], th_C0 = nb_threshold_list[conf[6]];
    if (th_C1 >= th_C0) th_C0 = 0;
    int rep_C2 = nb_thr_to_n_rep[NB_MIN_TH + 116], rep_C1, rep_C0;
    if (th_C1 < NB_NO_REP_TH) { if (th_C1 < NB_MIN_TH) th_C1 = NB_MIN_TH; rep_C1 = nb_thr_to_n_rep[th_C1 + 116]; } else rep_C1 = 0;
    if (th_C0 < NB_NO_REP_TH) { if (th_C0 < NB_MIN_TH) th_C0 = NB_MIN_TH; rep_C0 = nb_thr_to_n_rep[th_C0 + 116]; } else rep_C0 = 0;
    o->msg3_conf_nrep[1] = th_C1 < NB_NO_REP_MSG3_TH ? nb_msg3_conf_nrep[th_C1 + 116] : 1;
    o->msg3_conf_nrep[0] = th_C0 < NB_NO_REP_MSG3_TH ? nb_msg3_conf_nrep[th_C0 + 116] : 1;
    int RAR_window_size = nb_rar_window_list[conf[1]] * NB_NPDCCH_PERIOD;
    o->CE_thresholds[0] = th_C1; o->CE_thresholds[1] = th_C0;
    o->MAC_timer = nb_mac_timer_list[conf[2]] * NB_NPDCCH_PERIOD;
    o->preamble_trans_max = nb_transmax_list[conf[3]];
    o->probability_anchor = nb_panchor_den[conf[4]] ? 1.0 / nb_panchor_den[conf[4]] : 0.0;
    int args[3][3] = { { conf[7], rep_C0, conf[9] }, { conf[10], rep_C1, conf[12] }, { conf[13], rep_C2, conf[15] } };
    carrier_update_ra_parameters(o, args, th_C1, th_C0);
    for (int ce = 0; ce < 3; ++ce) o->RAR_window_size[ce] = RAR_window_size;
    memcpy(o->NPRACH_conf, conf, sizeof conf);
    schedule_event(o, o->event.t + NB_NPRACH_UPDATE_PERIOD, o->event);
    update_state(o);
}

/* ------------------------------------------------------------------ advance_fel / check_event */
static int check_event(Oracle *o, int type)                           /* node_b.py:248-261 */
{
    if (type == EV_NPDCCH_ARRIVAL) {
        if (o->state == NB_ST_RAR_WINDOW || o->state == NB_ST_SCHEDULING) return 1;
        schedule_NPDCCH_arrival(o, o->NPDCCH_sf_left);
        return 0;
    }
    o->state = type == EV_NPRACH_UPDATE ? NB_ST_NPRACH_UPDATE : NB_ST_RAR_WINDOW_END;
    return 1;
}

static void advance_fel(Oracle *o)                                    /* node_b.py:191-204 */
{
    for (;;) {
        Event ev = fel_pop(o);
        o->pops++;
        if (o->evlog && o->evlog_n < o->evlog_cap) { o->evlog[2 * o->evlog_n] = ev.t; o->evlog[2 * o->evlog_n + 1] = ev.type; o->evlog_n++; }
        o->time = ev.t;
        erase_past_sf(o, ev.t);                                /* carrier_set.step(t) */
        switch (ev.type) {
        case EV_NPDCCH_ARRIVAL: case EV_NPRACH_UPDATE: case EV_RAR_WINDOW_END:
            o->event = ev;
            if (check_event(o, ev.type)) return;
            break;
        case EV_NPRACH_START: access_NPRACH_start(o, &ev); break;
        case EV_NPRACH_END: access_NPRACH_end(o, &ev); break;
        case EV_MSG3: rx_msg3_event(o, &ev); break;
        case EV_NPUSCH_END: rx_NPUSCH_end(o, &ev); break;
        case EV_MAC_TIMER: population_MAC_timeout(o, &ev); break;
        case EV_BACKOFF_END: backoff_end(ev.ue); break;
        }
        if (o->fault & NBIOT_FAULT_FATAL_MASK) return;
    }
}

/* ------------------------------------------------------------------ get_node_info (node_b.py:278-418) */
static void fill_list(int32_t *dst, const ivec *src, Oracle *o)
{
    int n = src->n;
    int spill = o->conf.hist_cap > NBIOT_STEP_SPILL ? o->conf.hist_cap : NBIOT_STEP_SPILL;
    if (n > NBIOT_STEP_LIST + spill) o->fault |= NBIOT_FAULT_LIST_TRUNC;      /* the CUDA path keeps 16 + max(48, hist_cap) entries */
    if (n > NBIOT_STEP_LIST) n = NBIOT_STEP_LIST;
    for (int i = 0; i < n; ++i) dst[i] = src->d[i];
}

static void get_node_info(Oracle *o, nbiot_record_t *r)
{
    memset(r, 0, sizeof *r);
    r->time = o->time; r->state = o->state; r->total_ues = o->connected.n;
    for (int j = 0; j < NB_HORIZON; ++j)
        r->carrier_busy[j] = (uint8_t)__builtin_popcount(o->cols[0][(o->t_now + j) & (o->col_cap - 1)] & 0xFFFu);
    if (o->state == NB_ST_SCHEDULING || o->state == NB_ST_RAR_WINDOW_END) {
        int max_CE = nb_ce_for_sf_left[o->NPDCCH_sf_left];
        /* stable sort of connected UEs by sort_criterium, then the first N_users with CE_level <= max_CE */
        int n = o->connected.n;
        UE **s = malloc(sizeof(UE *) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < n; ++i) {
            UE *u = o->connected.d[i]; int j = i;
            while (j > 0 && (o->conf.sort_by_loss ? s[j - 1]->loss > u->loss : s[j - 1]->t_connection > u->t_connection)) { s[j] = s[j - 1]; --j; }
            s[j] = u;
        }
        int n_ue = 0;
        for (int i = 0; i < n && n_ue < NB_N_USERS; ++i) {
            UE *ue = s[i];
            if (ue->CE_level <= max_CE) {
                o->selectable[n_ue] = ue->id;
                r->ues[n_ue] = ue->id;
                r->connection_time[n_ue] = o->time - ue->t_connection;
                r->loss[n_ue] = ue->loss;
                r->buffer[n_ue] = ue->buffer;
                if (!ue->new_data) { double v = ue->sinr; if (!(v > -40)) v = -40; if (!(v < 30)) v = 30; r->sinr[n_ue] = v + 40; }
                n_ue++;
            }
        }
        free(s);
        o->n_selectable = n_ue; r->n_sel = n_ue;
        r->n_rx = o->rx_id.n; r->n_err = o->err_id.n; r->n_unfit = o->unfit_id.n; r->n_acc = o->acc_id.n;
        fill_list(r->rx_id, &o->rx_id, o); fill_list(r->rx_delay, &o->rx_delay, o);
        fill_list(r->err_id, &o->err_id, o); fill_list(r->unfit_id, &o->unfit_id, o);
        fill_list(r->acc_id, &o->acc_id, o); fill_list(r->acc_delay, &o->acc_delay, o);
    } else if (o->state == NB_ST_RAR_WINDOW) {
        int max_CE = nb_ce_for_sf_left[o->NPDCCH_sf_left];
        for (int i = 0; i < 3; ++i) {
            int s = ivec_sum(&o->RAR_UEs[i]);
            r->ues_per_CE[i] = i <= max_CE ? s : 0;
            r->RAR_in[i] = o->RAR_in[i]; r->RAR_sent[i] = o->RAR_sent[i]; r->RAR_ids[i] = o->RAR_ids[i];
            r->RAR_detected[i] = o->RAR_detected[i]; r->RAR_failed[i] = o->RAR_failed[i];
            r->action[i] = o->action[i];
        }
        r->valid_CE_control = o->valid_CE_control; r->NPDCCH_sf_left = o->NPDCCH_sf_left; r->total_departures = o->total_departures;
        memcpy(r->nprach_conf, o->NPRACH_conf, sizeof o->NPRACH_conf);
        o->RAR_failed[0] = o->RAR_failed[1] = o->RAR_failed[2] = 0;
    } else {   /* NPRACH update */
        int num_departures = o->departures.n;
        /* perf.estimate_carrier_resources :204-222 */
        if (o->time != 0) {
            int period = o->time - o->perf_t;
            long long total = 12LL * o->n_carriers * period;
            long long signalling = o->NPRACH_resources + o->msg3_resources;
            r->NPRACH_occupation = (double)o->NPRACH_resources / (double)total;
            r->RA_occupation = (double)signalling / (double)total;
            r->NPUSCH_occupation = (double)o->NPUSCH_resources / (double)(total - signalling);
            o->msg3_resources = o->NPRACH_resources = o->NPUSCH_resources = 0;
            o->perf_t = o->time;
        }
        long long sd = 0; for (int i = 0; i < num_departures; ++i) sd += o->departures.d[i];
        r->av_delay = num_departures ? (double)sd / (double)num_departures : 0;
        r->preambles[0] = (o->NPRACH_conf[9] + 1) * 12; r->preambles[1] = (o->NPRACH_conf[12] + 1) * 12; r->preambles[2] = (o->NPRACH_conf[15] + 1) * 12;
        for (int ce = 0; ce < 3; ++ce) {
            double sdr = 0; for (int i = 0; i < o->msg3_efficiency[ce].n; ++i) sdr += o->msg3_efficiency[ce].d[i];
            int n = o->msg3_efficiency[ce].n;
            r->msg3_detection[ce] = sdr / (double)(n > 1 ? n : 1);
            n = o->msg3_sent[ce].n; r->msg3_sent[ce] = (double)ivec_sum(&o->msg3_sent[ce]) / (double)(n > 1 ? n : 1);
            n = o->msg3_received[ce].n; r->msg3_received[ce] = (double)ivec_sum(&o->msg3_received[ce]) / (double)(n > 1 ? n : 1);
            n = o->detection_history[ce].n;
            r->detection_ratios[ce] = (double)ivec_sum(&o->detection_history[ce]) / (double)(n > 1 ? n : 1) / (double)r->preambles[ce];
            r->colision_ratios[ce] = (double)ivec_sum(&o->collision_history[ce]) / (double)(n > 1 ? n : 1) / (double)r->preambles[ce];
            r->n_occ[ce] = n;
            if (n > NBIOT_MAX_OCC) { n = NBIOT_MAX_OCC; o->fault |= NBIOT_FAULT_LIST_TRUNC; }
            for (int i = 0; i < n; ++i) { r->det_hist[ce][i] = (int16_t)o->detection_history[ce].d[i]; r->col_hist[ce][i] = (int16_t)o->collision_history[ce].d[i]; }
        }
        memcpy(r->distribution, o->distribution, sizeof o->distribution);
        r->departures = num_departures; r->n_incoming = o->incoming.n;
        memcpy(r->nprach_conf, o->NPRACH_conf, sizeof o->NPRACH_conf);
        /* hand the period lists over and reset */
        VCLEAR(o->info_incoming); VCLEAR(o->info_delays); VCLEAR(o->info_service);
        for (int i = 0; i < o->incoming.n; ++i) VPUSH(o->info_incoming, o->incoming.d[i]);
        for (int i = 0; i < num_departures; ++i) { VPUSH(o->info_delays, o->departures.d[i]); VPUSH(o->info_service, o->service_times.d[i]); }
        for (int ce = 0; ce < 3; ++ce) { VCLEAR(o->msg3_efficiency[ce]); VCLEAR(o->msg3_sent[ce]); VCLEAR(o->msg3_received[ce]);
                                         VCLEAR(o->detection_history[ce]); VCLEAR(o->collision_history[ce]); }
        VCLEAR(o->departures); VCLEAR(o->service_times); VCLEAR(o->incoming);
    }
    r->fault = o->fault;
    r->draws_lo = (int32_t)(uint32_t)o->draws; r->draws_hi = (int32_t)(uint32_t)(o->draws >> 32);
    r->events = (int32_t)o->pops;
}

/* ------------------------------------------------------------------ reset / step / API */
static void free_all_ues(Oracle *o)
{
    /* every live UE is in exactly one container, or is referenced by a pending NPUSCH_end event */
    for (int k = 0; k < 3; ++k) { for (int i = 0; i < o->ra_lists[k].n; ++i) free(o->ra_lists[k].d[i]); VCLEAR(o->ra_lists[k]); }
    for (int i = 0; i < o->contention.n; ++i) free(o->contention.d[i]); VCLEAR(o->contention);
    for (int i = 0; i < o->connected.n; ++i) free(o->connected.d[i]); VCLEAR(o->connected);
    for (int i = 0; i < o->heap_n; ++i) if (o->heap[i].ev.type == EV_NPUSCH_END) free(o->heap[i].ev.ue);
}

static void reset_storage(Oracle *o)                                  /* node_b.py:131-161 */
{
    for (int i = 0; i < 3; ++i) {
        VCLEAR(o->RAR_UEs[i]); VCLEAR(o->RAR_w_ids[i]);
        o->RAR_in[i] = o->RAR_ids[i] = o->RAR_attmp[i] = o->RAR_sent[i] = o->RAR_detected[i] = o->RAR_failed[i] = 0;
        VCLEAR(o->msg3_sent[i]); VCLEAR(o->msg3_received[i]); VCLEAR(o->msg3_efficiency[i]);
    }
    o->n_selectable = 0; o->n_scheduled = 0;
    VCLEAR(o->incoming); VCLEAR(o->departures); VCLEAR(o->service_times);
    for (int i = 0; i < 15; ++i) o->distribution[i] = 0.0;
    o->k = 0;
}

void orc_reset(Oracle *o, nbiot_record_t *out)                        /* node_b.py:163-189 */
{
    free_all_ues(o);
    o->heap_n = 0; o->counter = 0;                                    /* fel.reset() */
    /* m.reset(): population, carrier_set, perf_monitor, ue_generator, access_procedure */
    o->ue_count = 0;
    carrier_reset(o);
    perf_reset(o);
    gen_reset(o);
    o->probability_anchor = 0;
    for (int i = 0; i < 3; ++i) {
        o->detected_preambles_per_CE[i] = o->collided_preambles_per_CE[i] = o->contenders_per_CE[i] = o->RAR_counter[i] = 0;
        VCLEAR(o->detection_history[i]); VCLEAR(o->collision_history[i]);
    }
    reset_storage(o);
    o->state = NB_ST_NPRACH_UPDATE; o->time = 0;
    o->event = make_event(EV_NPRACH_UPDATE);
    schedule_event(o, 0, o->event);
    schedule_event(o, 0, make_event(EV_NPDCCH_ARRIVAL));
    advance_fel(o);
    get_node_info(o, out);
    out->reward = 0.0;
}

void orc_step(Oracle *o, const int32_t *action, nbiot_record_t *out)  /* node_b.py:228-245 */
{
    if (!(o->fault & NBIOT_FAULT_FATAL_MASK)) {
        switch (o->state) {
        case NB_ST_SCHEDULING: step_data(o, action); break;
        case NB_ST_RAR_WINDOW: step_RAR_window(o, action); break;
        case NB_ST_RAR_WINDOW_END: step_RAR_end(o, action); break;
        case NB_ST_NPRACH_UPDATE: step_update(o, action); break;
        default: break;
        }
        if (!(o->fault & NBIOT_FAULT_FATAL_MASK)) advance_fel(o);
    }
    get_node_info(o, out);
    out->reward = perf_reward(o);
}

Oracle *orc_create(const nbiot_conf_t *conf, uint64_t seed, const uint16_t *lut2, const uint8_t *mcs)
{
    Oracle *o = calloc(1, sizeof(Oracle));
    o->conf = *conf; o->seed = seed; o->lut2 = lut2; o->mcs = mcs;
    o->range = conf->max_d > 0 ? conf->max_d : NB_MAX_RANGE;
    o->n_carriers = conf->n_carriers >= 2 ? 2 : 1;
    o->col_cap = 1024;
    for (int c = 0; c < o->n_carriers; ++c) o->cols[c] = calloc((size_t)o->col_cap, sizeof(uint16_t));
    /* NodeB.__init__ (node_b.py:87-104) */
    o->NPDCCH_sf_left = NB_NPDCCH_SF;
    static const int default_conf[16] = { 3, 6, 2, 4, 12, 8, 11, 3, 0, 1, 2, 0, 1, 3, 2, 1 };   /* control_default_values in NPRACH_items order */
    memcpy(o->NPRACH_conf, default_conf, sizeof default_conf);
    for (int i = 0; i < 3; ++i) o->RAR_window_size[i] = 4 * NB_NPDCCH_PERIOD;
    o->valid_CE_control = 1;
    o->msg3_conf_nrep[0] = 1; o->msg3_conf_nrep[1] = 4; o->msg3_conf_nrep[2] = 16;   /* node_b.py:36 */
    /* Population.__init__ (population.py:41-49) */
    o->CE_thresholds[0] = -102; o->CE_thresholds[1] = -85;   /* [-102.5,-85.0]: overwritten by the t=0 update before any arrival */
    o->preamble_trans_max = 2; o->MAC_timer = 64;
    o->gen_p = 0.5; o->gen_counter = 0;
    o->cluster_count = 0; o->cluster_center[0] = o->cluster_center[1] = 0;
    o->cluster_range = conf->cluster_range / o->range;       /* population.py:62 */
    return o;
}

void orc_destroy(Oracle *o)
{
    if (!o) return;
    free_all_ues(o);
    free(o->heap);
    for (int c = 0; c < 2; ++c) free(o->cols[c]);
    for (int i = 0; i < 3; ++i) {
        VFREE(o->RAR_UEs[i]); VFREE(o->RAR_w_ids[i]); VFREE(o->msg3_sent[i]); VFREE(o->msg3_received[i]); VFREE(o->msg3_efficiency[i]);
        VFREE(o->ra_lists[i]); VFREE(o->detection_history[i]); VFREE(o->collision_history[i]);
    }
    VFREE(o->connected); VFREE(o->contention); VFREE(o->incoming); VFREE(o->departures); VFREE(o->service_times);
    VFREE(o->sc_log); VFREE(o->rx_id); VFREE(o->rx_delay); VFREE(o->err_id); VFREE(o->unfit_id); VFREE(o->acc_id); VFREE(o->acc_delay);
    VFREE(o->info_incoming); VFREE(o->info_delays); VFREE(o->info_service);
    VFREE(o->access_history); VFREE(o->connection_history); VFREE(o->delay_history); VFREE(o->bits_history);
    VFREE(o->samples);
    free(o);
}

/* statistics histories since the last reset: which = 0 delay, 1 connection (service time), 2 access, 3 acked bits;
 * returns the true length, copies at most cap entries */
int orc_get_stats(Oracle *o, int which, int32_t *out, int cap)
{
    ivec *v = which == 0 ? &o->delay_history : which == 1 ? &o->connection_history : which == 2 ? &o->access_history : &o->bits_history;
    for (int i = 0; i < v->n && i < cap; ++i) out[i] = v->d[i];
    return v->n;
}

/* the per-epoch info lists in full (perf_monitor.py:92-103): which = 0 rx id, 1 rx delay, 2 errors, 3 unfit, 4 access id, 5 access delay;
 * returns the true length, copies at most cap entries */
int orc_get_step_list(Oracle *o, int which, int32_t *out, int cap)
{
    ivec *v = which == 0 ? &o->rx_id : which == 1 ? &o->rx_delay : which == 2 ? &o->err_id : which == 3 ? &o->unfit_id : which == 4 ? &o->acc_id : &o->acc_delay;
    for (int i = 0; i < v->n && i < cap; ++i) out[i] = v->d[i];
    return v->n;
}

/* the sample lists since the last reset, in append order; returns the true length, copies at most cap samples */
int orc_get_samples(Oracle *o, nbiot_trace_rec_t *out, int cap)
{
    for (int i = 0; i < o->samples.n && i < cap; ++i) out[i] = o->samples.d[i];
    return o->samples.n;
}

/* lists of the last NPRACH_update info: returns the true lengths, copies at most cap entries */
void orc_get_hist(Oracle *o, double *incoming, int32_t *delays, int32_t *service, int cap, int32_t *n_incoming, int32_t *n_departures)
{
    *n_incoming = o->info_incoming.n; *n_departures = o->info_delays.n;
    for (int i = 0; i < o->info_incoming.n && i < cap; ++i) incoming[i] = o->info_incoming.d[i];
    for (int i = 0; i < o->info_delays.n && i < cap; ++i) { delays[i] = o->info_delays.d[i]; service[i] = o->info_service.d[i]; }
}

/* PerfMonitor cumulative counters (perf_monitor.py:41-59, :303-304, :353) */
void orc_get_counters(Oracle *o, int64_t out[16], double *av_delay)
{
    for (int i = 0; i < 3; ++i) { out[i] = o->ue_arrivals[i]; out[3 + i] = o->ue_departures[i]; out[6 + i] = o->backoffs[i]; }
    out[9] = o->ue_connections; out[10] = o->error_count; out[11] = o->attempts_count; out[12] = o->total_bits;
    out[13] = (int64_t)o->draws; out[14] = o->pops; out[15] = o->ue_count;
    *av_delay = o->av_delay;
}

void orc_set_event_log(Oracle *o, int *buf, int cap) { o->evlog = buf; o->evlog_cap = cap; o->evlog_n = 0; }
int orc_event_log_len(Oracle *o) { return o->evlog_n; }

/* the notebook-cell-8 loop: `while node.time < T: node.step(action)` with one fixed action;
 * returns the number of decisions taken (used by the CPU-baseline timing) */
long long orc_run_fixed(Oracle *o, const int32_t *action, int until_time, nbiot_record_t *last)
{
    long long n = 0;
    while (o->time < until_time && !(o->fault & NBIOT_FAULT_FATAL_MASK)) { orc_step(o, action, last); n++; }
    return n;
}

/* ------------------------------------------------------------------ controller + multi-cell timing driver */
/* Controller.step + to_next_step / run_until (controller.py:185-195, :234-288) with fixed-action internal
 * agents, over the same write tables the CUDA path uses (include/nbiot_ctrl.h). ctrl_action is the shared
 * 23-int vector of this cell (controller.py:33). Returns the number of NodeB steps. */
static void ctrl_apply(int32_t *act, const int8_t *w) { for (int i = 0; i < NB_N_ACTIONS; ++i) if (w[i] >= 0) act[i] = w[i]; }

int orc_ctrl_advance(Oracle *o, const nbiot_ctrl_t *ctrl, int32_t *ctrl_action, const uint8_t *ext_action, int until_time, int max_steps, nbiot_record_t *out)
{
    int steps = 0;
    if (o->fault & NBIOT_FAULT_FATAL_MASK) return 0;
    /* without an external action (run_until / single_step, controller.py:156-195) every agent of a state acts, the one
     * registered as external included, and no state hands control back */
    const int8_t (*sw)[24] = ext_action ? ctrl->state_writes : ctrl->full_writes;
    const uint32_t ext_mask = ext_action ? ctrl->ext_state_mask : 0u;
    if (ext_action) {
        for (int i = 0; i < ctrl->ext_k; ++i) ctrl_action[ctrl->ext_a_mask[i]] = ext_action[i];
        ctrl_apply(ctrl_action, ctrl->ext_post);
    } else ctrl_apply(ctrl_action, sw[o->state]);
    while (steps < max_steps && !(until_time >= 0 && o->time >= until_time)) {
        if (steps > 0) ctrl_apply(ctrl_action, sw[o->state]);
        orc_step(o, ctrl_action, out);
        steps += 1;
        if (o->fault & NBIOT_FAULT_FATAL_MASK) break;
        if ((ext_mask >> o->state) & 1u) { ctrl_apply(ctrl_action, sw[o->state]); break; }
    }
    return steps;
}

typedef struct {
    const nbiot_conf_t *conf; const nbiot_ctrl_t *ctrl; const uint16_t *lut2; const uint8_t *mcs;
    const uint64_t *seeds; const uint8_t *actions;   /* [steps][n_inst][ext_k] or NULL */
    int n_inst, steps, until_time, first, stride;
    long long subframes, node_steps; double seconds;
    nbiot_record_t *records;      /* optional [n_inst]: the last decision record of every cell */
} BenchJob;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

static void *bench_worker(void *arg)
{
    BenchJob *j = (BenchJob *)arg;
    nbiot_record_t rec;
    /* construction and reset are outside the timed region, like the import/creation cost of the reference */
    int mine = 0;
    for (int i = j->first; i < j->n_inst; i += j->stride) mine++;
    Oracle **cells = malloc(sizeof(Oracle *) * (size_t)(mine > 0 ? mine : 1));
    int32_t (*acts)[24] = malloc(sizeof(int32_t[24]) * (size_t)(mine > 0 ? mine : 1));
    int k = 0;
    for (int i = j->first; i < j->n_inst; i += j->stride, ++k) {
        cells[k] = orc_create(j->conf, j->seeds[i], j->lut2, j->mcs);
        for (int a = 0; a < 24; ++a) acts[k][a] = j->ctrl->init_action[a];
        orc_reset(cells[k], &rec);
    }
    double t0 = now_s();
    long long sf = 0, ns = 0;
    k = 0;
    for (int i = j->first; i < j->n_inst; i += j->stride, ++k) {
        int t_before = cells[k]->time;
        if (j->actions)
            for (int s = 0; s < j->steps; ++s)
                ns += orc_ctrl_advance(cells[k], j->ctrl, acts[k], j->actions + ((size_t)s * j->n_inst + i) * j->ctrl->ext_k, -1, 1 << 30, &rec);
        else ns += orc_ctrl_advance(cells[k], j->ctrl, acts[k], NULL, j->until_time, 1 << 30, &rec);
        sf += cells[k]->time - t_before;
        if (j->records) j->records[i] = rec;
    }
    j->seconds = now_s() - t0;
    j->subframes = sf; j->node_steps = ns;
    for (k = 0; k < mine; ++k) orc_destroy(cells[k]);
    free(cells); free(acts);
    return NULL;
}

/* simulate n_inst cells on n_threads host threads (instances are dealt round-robin); returns the slowest
 * thread's wall-clock seconds inside the step loop and the total simulated subframes */
double orc_bench_records(const nbiot_conf_t *conf, const nbiot_ctrl_t *ctrl, const uint16_t *lut2, const uint8_t *mcs, const uint64_t *seeds,
                         int n_inst, const uint8_t *actions, int steps, int until_time, int n_threads, long long *subframes, long long *node_steps,
                         nbiot_record_t *records)
{
    if (n_threads < 1) n_threads = 1;
    BenchJob *jobs = calloc((size_t)n_threads, sizeof(BenchJob));
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        BenchJob j = { conf, ctrl, lut2, mcs, seeds, actions, n_inst, steps, until_time, t, n_threads, 0, 0, 0.0, records };
        jobs[t] = j;
        pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
    }
    double worst = 0; long long sf = 0, ns = 0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].seconds > worst) worst = jobs[t].seconds;
        sf += jobs[t].subframes; ns += jobs[t].node_steps;
    }
    *subframes = sf; *node_steps = ns;
    free(jobs); free(th);
    return worst;
}

double orc_bench(const nbiot_conf_t *conf, const nbiot_ctrl_t *ctrl, const uint16_t *lut2, const uint8_t *mcs, const uint64_t *seeds,
                 int n_inst, const uint8_t *actions, int steps, int until_time, int n_threads, long long *subframes, long long *node_steps)
{ return orc_bench_records(conf, ctrl, lut2, mcs, seeds, n_inst, actions, steps, until_time, n_threads, subframes, node_steps, NULL); }

/* host twins of the random stream, for the Python shim that drives the reference */
double orc_rng_u01(uint64_t seed, uint64_t ctr, uint32_t stream) { return nb_rng_u01(seed, ctr, stream); }
void orc_rng_pair(uint64_t seed, uint64_t ctr, uint32_t stream, double *xy) { nb_rng_pair(seed, ctr, stream, &xy[0], &xy[1]); }
double orc_rng_normal(uint64_t seed, uint64_t ctr, uint32_t stream) { return nb_rng_normal(seed, ctr, stream); }
uint64_t orc_rng_bounded(uint64_t seed, uint64_t ctr, uint32_t stream, uint64_t n) { return nb_rng_bounded(seed, ctr, stream, n); }
int orc_record_size(void) { return (int)sizeof(nbiot_record_t); }
int orc_conf_size(void) { return (int)sizeof(nbiot_conf_t); }
