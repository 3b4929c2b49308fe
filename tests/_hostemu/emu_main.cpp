// CPU build of the device simulation core (nb_iot_environment_b200/csrc/nbiot_core.cuh) for
// debugging without a GPU: same source, same state layout, lanes emulated (see nbiot_warp.h).
// TEST AID ONLY: built and loaded by tests/, never by the package.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "nbiot_core.cuh"

struct Emu {
    nbiot_conf_t conf;
    NbCtaConst cc;
    NbLayout L;
    uint8_t *state;
    const uint16_t *lut2;
    const uint8_t *mcs;
    int32_t *evlog; int evlog_cap;
};

// mirrors the kernel's staging: [NbCtx | NbHdr | grid] in one scratch slice, copied in and out
struct Staged {
    Emu *e; int i; unsigned char *buf; NbCtx *c; uint16_t *blkbuf = NULL;
    Staged(Emu *e_, int i_, bool load) : e(e_), i(i_)
    {
        const NbLayout &L = e->L;
        size_t gbytes = (size_t)L.n_carriers * L.grid_w * 2;
        buf = (unsigned char *)aligned_alloc(16, sizeof(NbCtx) + sizeof(NbHdr) + gbytes);
        c = (NbCtx *)buf;
        uint8_t *s = e->state;
        for (int k = 0; k < 3; ++k) c->ra[k] = (NbRaEntry *)(s + L.off_ra) + ((size_t)i * 3 + k) * L.ra_cap;
        c->conn = (NbConnEntry *)(s + L.off_conn) + (size_t)i * L.ue_cap;
        c->cont = (NbContEntry *)(s + L.off_cont) + (size_t)i * L.ue_cap;
        c->ue = (NbUe *)(s + L.off_ue) + (size_t)i * L.ue_cap;
        c->free_stack = (uint16_t *)(s + L.off_free) + (size_t)i * L.ue_cap;
        c->occ = (NbOccHist *)(s + L.off_occ) + i;
        c->incoming = (double *)(s + L.off_incoming) + (size_t)i * L.hist_cap;
        c->delays = (int32_t *)(s + L.off_delays) + (size_t)i * L.hist_cap;
        c->service = (int32_t *)(s + L.off_service) + (size_t)i * L.hist_cap;
        c->scratch = (uint16_t *)(s + L.off_scratch) + (size_t)i * L.ra_cap;
        c->rec = (nbiot_record_t *)(s + L.off_rec) + i;
        c->spill = (int32_t *)(s + L.off_spill) + (size_t)i * 6u * nb_spill_cap(L.hist_cap);
        c->lut2 = e->lut2; c->mcs = e->mcs; memcpy(&c->conf, &e->conf, sizeof(nbiot_conf_hot_t)); c->ctrl = &e->cc.ctrl;
        c->W = L.grid_w; c->ue_cap = L.ue_cap; c->ra_cap = L.ra_cap; c->hist_cap = L.hist_cap; c->C = L.n_carriers;
        c->evlog = (e->evlog && i == 0) ? e->evlog : NULL; c->evlog_cap = e->evlog_cap;
        c->grid = (uint16_t *)(buf + sizeof(NbCtx) + sizeof(NbHdr));
#if NB_LANES == 1
        c->blk = NULL;
#endif
        if (load) {
            memcpy(buf + sizeof(NbCtx), (NbHdr *)(s + L.off_hdr) + i, sizeof(NbHdr));
            memcpy(buf + sizeof(NbCtx) + sizeof(NbHdr), (uint16_t *)(s + L.off_grid) + (size_t)i * L.n_carriers * L.grid_w, gbytes);
#if defined(NBIOT_EMU_BLK) && NB_LANES == 1
            // the thread-per-cell kernel's block summaries: derived data, filled with garbage and rebuilt whenever a cell is loaded
            blkbuf = (uint16_t *)malloc(sizeof(uint16_t) * L.n_carriers * (L.grid_w >> 5));
            memset(blkbuf, 0xA5, sizeof(uint16_t) * L.n_carriers * (L.grid_w >> 5));
            c->blk = blkbuf;
            nbc_blk_rebuild(*c);
#endif
        }
    }
    ~Staged()
    {
        const NbLayout &L = e->L;
        size_t gbytes = (size_t)L.n_carriers * L.grid_w * 2;
        memcpy((NbHdr *)(e->state + L.off_hdr) + i, buf + sizeof(NbCtx), sizeof(NbHdr));
        memcpy((uint16_t *)(e->state + L.off_grid) + (size_t)i * L.n_carriers * L.grid_w, buf + sizeof(NbCtx) + sizeof(NbHdr), gbytes);
        free(buf); free(blkbuf);
    }
};

#ifdef NBIOT_EMU_TASKS
// optional (instance, task class) log of the task chain: measurement aid for scheduling studies (tools/vote_sim.py)
int32_t *emu_tasklog = NULL; int emu_tasklog_cap = 0, emu_tasklog_n = 0;
extern "C" void emu_set_tasklog(int32_t *buf, int cap) { emu_tasklog = buf; emu_tasklog_cap = cap; emu_tasklog_n = 0; }
extern "C" int emu_tasklog_len(void) { return emu_tasklog_n; }
#endif

extern "C" {

int emu_lanes(void) { return NB_LANES; }
#ifdef NBIOT_EMU_TASKS
int emu_tasks(void) { return 1; }
#else
int emu_tasks(void) { return 0; }
#endif
#ifdef NBIOT_EMU_COUNTERS
void emu_debug_counts(long long out[16]) { for (int i = 0; i < 16; ++i) out[i] = nb_dbg_count[i]; }
#endif
int emu_hdr_size(void) { return (int)sizeof(NbHdr); }

void *emu_create(const nbiot_conf_t *conf, int n_inst, const uint64_t *seeds, const uint16_t *lut2, const uint8_t *mcs, const nbiot_ctrl_t *ctrl)
{
    Emu *e = (Emu *)calloc(1, sizeof(Emu));
    e->conf = *conf; e->cc.ctrl = nb_make_ctrl_dev(ctrl); e->lut2 = lut2; e->mcs = mcs;
    e->L = nb_make_layout(conf, n_inst);
    e->state = (uint8_t *)calloc(1, e->L.total);
    e->cc.cold = nb_make_cold(conf, e->state, e->L.off_occ, e->L.off_stat, e->L.off_trace, e->L.off_sclog);
    for (int i = 0; i < n_inst; ++i) { Staged st(e, i, false); nbc_construct(*st.c, seeds[i], &e->cc.ctrl); }
    return e;
}

void emu_destroy(void *p) { Emu *e = (Emu *)p; free(e->state); free(e); }
void emu_set_ctrl(void *p, const nbiot_ctrl_t *ctrl) { ((Emu *)p)->cc.ctrl = nb_make_ctrl_dev(ctrl); }
void emu_set_evlog(void *p, int32_t *buf, int cap) { Emu *e = (Emu *)p; e->evlog = buf; e->evlog_cap = cap; }

void emu_reset(void *p)
{
    Emu *e = (Emu *)p;
    for (int i = 0; i < e->L.n_inst; ++i) { Staged st(e, i, true); nbc_reset(*st.c); }
}

// ext_actions: [n_inst][ext_k] or NULL; steps_out: [n_inst] or NULL
void emu_advance(void *p, const uint8_t *ext_actions, int until_time, int max_steps, int32_t *steps_out)
{
    Emu *e = (Emu *)p;
#ifdef NBIOT_EMU_TASKS
    extern int32_t *emu_tasklog; extern int emu_tasklog_cap, emu_tasklog_n;
    // the task chain of the migrating-instance kernel, with the instance stored to and reloaded from the state buffer between
    // two tasks into a fresh context -- what a hop to another SM does -- so that nothing can survive outside NbHdr
    for (int i = 0; i < e->L.n_inst; ++i) {
        const uint8_t *a = ext_actions ? ext_actions + (size_t)i * e->cc.ctrl.ext_k : NULL;
        int k;
        { Staged st(e, i, true); st.c->tk_until = until_time; st.c->tk_max_steps = max_steps; st.c->tk_internal = a ? 0 : 1; k = nbc_task_begin(*st.c, a); }
        while (k != NB_TK_DONE) {
            if (emu_tasklog && emu_tasklog_n < emu_tasklog_cap) { emu_tasklog[2 * emu_tasklog_n] = i; emu_tasklog[2 * emu_tasklog_n + 1] = k; emu_tasklog_n += 1; }
            Staged st(e, i, true); st.c->tk_until = until_time; st.c->tk_max_steps = max_steps; st.c->tk_internal = a ? 0 : 1; k = nbc_task_run(*st.c, k);
        }
        if (steps_out) steps_out[i] = ((NbHdr *)(e->state + e->L.off_hdr) + i)->tk_steps;
    }
#else
    for (int i = 0; i < e->L.n_inst; ++i) {
        Staged st(e, i, true);
        int s = nbc_controller_advance(*st.c, &e->cc.ctrl, ext_actions ? ext_actions + (size_t)i * e->cc.ctrl.ext_k : NULL, until_time, max_steps);
        if (steps_out) steps_out[i] = s;
    }
#endif
}

// statistics histories of instance i: which = 0 delay, 1 connection, 2 access, 3 bits; returns the number recorded
int emu_stats(void *p, int i, int which, int32_t *out, int cap)
{
    Emu *e = (Emu *)p;
    int n = ((NbOccHist *)(e->state + e->L.off_occ))[i].stat_n, sc = e->L.stat_cap;
    const int32_t *v = (const int32_t *)(e->state + e->L.off_stat) + ((size_t)i * 4 + which) * sc;
    for (int k = 0; k < n && k < cap && k < sc; ++k) out[k] = v[k];
    return n;
}

// sample ring of instance i (conf['traces'] / conf['statistics']): copies up to cap samples in order of occurrence, returns the number recorded
int emu_trace(void *p, int i, nbiot_trace_rec_t *out, int cap)
{
    Emu *e = (Emu *)p;
    int n = ((NbOccHist *)(e->state + e->L.off_occ))[i].trace_n, tc = e->L.trace_cap;
    const nbiot_trace_rec_t *v = (const nbiot_trace_rec_t *)(e->state + e->L.off_trace) + (size_t)i * tc;
    int first = n > tc ? n - tc : 0;
    for (int k = first, j = 0; k < n && j < cap; ++k, ++j) out[j] = v[k % tc];
    return n;
}

const void *emu_records(void *p) { Emu *e = (Emu *)p; return e->state + e->L.off_rec; }
const void *emu_state(void *p) { return ((Emu *)p)->state; }
void emu_layout(void *p, NbLayout *out) { *out = ((Emu *)p)->L; }

// PerfMonitor counters of instance i (same order as orc_get_counters)
void emu_counters(void *p, int i, int64_t out[16], double *av_delay)
{
    Emu *e = (Emu *)p; NbHdr *h = (NbHdr *)(e->state + e->L.off_hdr) + i;
    for (int k = 0; k < 3; ++k) { out[k] = h->ue_arrivals[k]; out[3 + k] = h->ue_departures[k]; out[6 + k] = h->backoffs[k]; }
    out[9] = h->ue_connections; out[10] = h->error_count; out[11] = h->attempts_count; out[12] = h->total_bits;
    out[13] = (int64_t)h->draws; out[14] = h->events; out[15] = h->ue_count;
    *av_delay = h->av_delay;
}

}
