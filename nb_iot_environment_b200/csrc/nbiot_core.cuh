/*
 * nbiot_core.cuh -- the simulation core: one warp advances one NB-IoT cell instance.
 *
 * The reference is a discrete-event simulator (heap of (time, seq, event), one Python
 * handler per event; system/node_b.py:191-204). This core keeps the reference's event
 * order -- every event carries the same (time, insertion count) key the reference's heap
 * would give it -- but has no heap and no per-UE objects:
 *   * the clock jumps to the minimum key over a handful of fixed event slots
 *     (nb_next_event), i.e. a time-stepped loop that skips the subframes without work;
 *   * Backoff_end events (the bulk of the reference's heap under congestion,
 *     population.py:143-149, :291-296) are never popped: a UE in BACKOFF whose wake key is
 *     older than the key of the NPRACH occasion being processed counts as RAO;
 *   * a MAC_timer whose UE already connected is dropped at msg4 (the reference pops it as
 *     a no-op, population.py:151-163);
 *   * the look-ahead grid is a ring of 12-bit column masks filled run by run, not column
 *     by column (carrier.py:382-404, :473-489), with the same event insertion counts;
 *   * containers are flat arrays in list order, scanned 32 entries at a time with
 *     ballot / popc ranks standing in for Python list order.
 * Scalar, order-dependent logic is executed uniformly by all lanes; lane-parallel
 * sections are wrapped in nbw_run(...) (see nbiot_warp.h).
 *
 * Reference functions restated here are cited at each function (path:line under the
 * reference root).
 */
#ifndef NBIOT_CORE_CUH
#define NBIOT_CORE_CUH

#ifndef NB_GRID_MODE
#define NB_GRID_MODE 0     /* 0: one column per lane per step; 1: 8-column chunks; 2: chunks + 32-column block summaries */
#endif
#include "nbiot_params.h"
#include "nbiot_rng.h"
#include "nbiot_state.h"
#include "nbiot_warp.h"

/* Per-warp working context. It sits in shared memory directly in front of the staged header and grid
 *   [ NbCtx | NbHdr | grid[C][W] (when staged) ]
 * so the hot state is reached from the one register holding &ctx with constant offsets (no pointer chase),
 * and the compiler is told it is shared memory. The HBM-resident arrays are reached through the pointers. */
struct __attribute__((aligned(16))) NbCtx {
    NbRaEntry *ra[3];
    NbConnEntry *conn;
    NbContEntry *cont;
    NbUe *ue;
    uint16_t *free_stack;
    NbOccHist *occ;
    double *incoming;
    int32_t *delays, *service;
    uint16_t *scratch;
    nbiot_record_t *rec;
    int32_t *spill;           /* [6][spill cap]: entries 16.. of the per-epoch info lists (rx id / delay, errors, unfit, access id / delay) */
    const uint16_t *lut2;
    const uint8_t *mcs;
    nbiot_conf_hot_t conf;    /* by value: one shared-memory load per use instead of a pointer chase into HBM */
    const NbCtrlDev *ctrl;    /* controller tables: the first member of an NbCtaConst (shared memory in the kernels) */
    int32_t *evlog;           /* optional (t, type) pop log for debugging */
    int W, ue_cap, ra_cap, hist_cap, C, evlog_cap;
    uint16_t *grid;           /* the look-ahead ring: staged in shared memory behind the header, or in place in HBM (short launches) */
#if NB_LANES == 1
    uint16_t *blk;            /* [C][W / 32] block summaries of the ring (thread-per-cell kernel; NULL: plain column scans) */
#endif
#ifdef NB_TPC_GLOBAL_HDR
    struct NbHdr *hdr_g;
#endif
#ifdef NB_TASKS
    int tk_until, tk_max_steps;   /* bounds of this launch (see "tasks" at the end of the file) */
    int tk_internal;              /* the launch has no external action: every agent of a state acts, no state hands control back */
#endif
};

/* specialised builds: NB_FIXED_C = 1 compiles the carrier loops for single-carrier cells, NB_NO_EVLOG drops the debug event log */
#ifdef NB_FIXED_C
#define NB_NC(c) NB_FIXED_C
#else
#define NB_NC(c) ((c).C)
#endif
#ifdef NB_NO_EVLOG
#define NB_EVLOG(c) ((int32_t *)0)
#else
#define NB_EVLOG(c) ((c).evlog)
#endif

/* NBIOT_EMU_COUNTERS (tests/_hostemu only): how often the first-fit search does what */
#if defined(NBIOT_EMU_COUNTERS) && !defined(__CUDACC__)
static long long nb_dbg_count[16];
#define NB_COUNT(i, n) (nb_dbg_count[i] += (n))
#else
#define NB_COUNT(i, n) ((void)0)
#endif

#if defined(__CUDA_ARCH__) && !defined(NB_TPC)
#define NB_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define NB_ASSUME_SHARED(p) ((void)0)
#endif
#ifdef NB_TPC_GLOBAL_HDR   /* experiment: the thread-per-cell kernel reaches the header in place in HBM (no local copy) */
NB_DEV NbHdr *nbc_hdr(NbCtx &c) { return c.hdr_g; }
#else
NB_DEV NbHdr *nbc_hdr(NbCtx &c) { NbHdr *h = (NbHdr *)((char *)&c + sizeof(NbCtx)); NB_ASSUME_SHARED(h); return h; }
#endif
NB_DEV const NbColdConf *nbc_cold(NbCtx &c) { return &((const NbCtaConst *)c.ctrl)->cold; }
/* NB_GRID_STAGED: a build that only runs with the look-ahead ring staged behind the header (long launches): the ring is reached from the
 * context pointer with constant offsets and shared-memory loads, not through a generic pointer that may also point into HBM */
#ifdef NB_GRID_STAGED
NB_DEV uint16_t *nbc_grid(NbCtx &c)
{
    /* through an integer: the ring lies beyond the NbHdr object behind the context, and pointer arithmetic past an object is
     * not something the optimiser has to honour (it folded these loads to zero) */
    uint16_t *g = (uint16_t *)((uintptr_t)&c + sizeof(NbCtx) + sizeof(NbHdr));
    NB_ASSUME_SHARED(g);
    return g;
}
#else
NB_DEV uint16_t *nbc_grid(NbCtx &c) { return c.grid; }
#endif
#define NB_WARP_BYTES(C, W) (sizeof(NbCtx) + sizeof(NbHdr) + (size_t)(C) * (size_t)(W) * 2u)

#define NB_FAULT(c, bit) (nbc_hdr(c)->fault |= (int32_t)(bit))
#define NB_FATAL(c) (nbc_hdr(c)->fault & (int32_t)NBIOT_FAULT_FATAL_MASK)
#define NB_STATE(e) ((e).st_ce & 15u)
#define NB_CE(e) (((e).st_ce >> 4) & 3u)

/* PerfMonitor's sample lists (perf_monitor.py:61-79: nprach_sample, rar_sample, rar_window_sample, register_*_resources, npusch_delivered)
 * as one ring of 16-byte samples per cell in HBM, in the order the reference appends; out of line and off unless conf['traces'] /
 * conf['statistics'] (NbColdConf.trace_mask) */
NB_FN void nbc_trace(NbCtx &c, int type, int ce, int t, int a, int b)
{
    const NbColdConf *cold = (const NbColdConf *)&((const NbCtaConst *)c.ctrl)->cold;
    nbiot_trace_rec_t *tr = cold->trace_base + (size_t)(c.occ - cold->occ_base) * (size_t)cold->trace_cap;
    nbiot_trace_rec_t r; r.t = t; r.a = a; r.b = b; r.type = (uint8_t)type; r.ce = (uint8_t)ce; r.pad = 0;
    tr[c.occ->trace_n % cold->trace_cap] = r;
    c.occ->trace_n += 1;
}
#define NB_TRACE(c, type, ce, t, a, b) do { if (((const NbCtaConst *)(c).ctrl)->cold.trace_mask & (type)) nbc_trace(c, type, ce, t, a, b); } while (0)

/* entry k >= NBIOT_STEP_LIST of a per-epoch info list (0 rx id, 1 rx delay, 2 errors, 3 unfit, 4 access id, 5 access delay): the header
 * and the record hold the first 16, the rest goes to the side buffer in HBM (rare: only loaded cells between two scheduling decisions) */
NB_FN void nbc_step_spill(NbCtx &c, int list, int k, int v)
{
    const int cap = nb_spill_cap(c.hist_cap);
    if (k - NBIOT_STEP_LIST < cap) c.spill[list * cap + (k - NBIOT_STEP_LIST)] = v;
}

/* ------------------------------------------------------------------ scalar draws */
NB_DEV double nbc_u01(NbCtx &c) { return nb_rng_u01(nbc_hdr(c)->seed, nbc_hdr(c)->draws++, NBIOT_STREAM_ENV); }
NB_DEV double nbc_normal(NbCtx &c, double mu, double sigma)
{ return NB_ADD(mu, NB_MUL(sigma, nb_rng_normal(nbc_hdr(c)->seed, nbc_hdr(c)->draws++, NBIOT_STREAM_ENV))); }
NB_DEV int nbc_bernoulli(NbCtx &c, double p) { return nbc_u01(c) < p ? 1 : 0; }
NB_DEV int nbc_integers(NbCtx &c, int low, int high)
{ return low + (int)nb_rng_bounded(nbc_hdr(c)->seed, nbc_hdr(c)->draws++, NBIOT_STREAM_ENV, (uint64_t)(high - low)); }

NB_DEV int nbc_rep_index(int N_rep) { int r = 0; NB_NOUNROLL while ((1 << r) < N_rep) ++r; return r; }
NB_DEV int nbc_sf_per_ru(int N_sc) { return N_sc == 12 ? 1 : N_sc == 6 ? 2 : N_sc == 3 ? 4 : N_sc == 1 ? 8 : -1; }

/* ------------------------------------------------------------------ event pool */
NB_FNM2 void nbc_pool_push(NbCtx &c, uint64_t key, int type, int ce, int aux)
{
    NbHdr *h = nbc_hdr(c);
    if (!h->pool_free) { NB_FAULT(c, NBIOT_FAULT_EVENT_SLOTS); return; }
#ifdef NB_TEST_POOL_HIGH
    int i = 63 - __builtin_clzll(h->pool_free);              /* test builds (tests/_hostemu): highest free entry first, so that the second
                                                              * pass of the event selection is exercised by ordinary traces */
#else
    int i = nbw_ffsll(h->pool_free) - 1;
#endif
    h->pool_free &= ~(1ULL << i);
    NbEvent e; e.key = key; e.aux = aux; e.type = (uint8_t)type; e.ce = (uint8_t)ce; e.pad0 = e.pad1 = 0;
    h->pool[i] = e;
}

NB_DEV void nbc_pool_release(NbCtx &c, int i) { nbc_hdr(c)->pool[i].type = NB_EV_NONE; nbc_hdr(c)->pool[i].key = NB_KEY_NONE; nbc_hdr(c)->pool_free |= 1ULL << i; }

/* ------------------------------------------------------------------ carrier grid (carrier.py:414-605) */
/* The look-ahead grid is a ring of W 12-bit column masks per carrier (uint16 each). Window operations
 * work on aligned 16-byte chunks of 8 columns: one chunk per lane per step, 256 columns per warp step. */
struct __attribute__((aligned(16))) NbChunk { uint32_t w[4]; };

NB_DEV NbChunk *nbc_chunk(NbCtx &c, int carrier, int q) { return (NbChunk *)&nbc_grid(c)[carrier * c.W + ((q << 3) & (c.W - 1))]; }
#if defined(NB_FIXED_C) && NB_FIXED_C == 1
NB_DEV uint16_t *nbc_col(NbCtx &c, int, int t) { return &nbc_grid(c)[t & (c.W - 1)]; }      /* one carrier: its index is 0 */
#else
NB_DEV uint16_t *nbc_col(NbCtx &c, int carrier, int t) { return &nbc_grid(c)[carrier * c.W + (t & (c.W - 1))]; }
#endif

/* 32-bit lane mask of word j (columns lo+2j, lo+2j+1) restricted to columns [first, last) of the chunk */
NB_DEV uint32_t nbc_word_mask(int j, int first, int last)
{ return ((2 * j >= first && 2 * j < last) ? 0xFFFFu : 0u) | ((2 * j + 1 >= first && 2 * j + 1 < last) ? 0xFFFF0000u : 0u); }

/* this lane's share of the OR of columns [a, a+n) read chunk-wise (no reduction) */
NB_DEV uint32_t nbc_scan_chunks(NbCtx &c, int carrier, int a, int n)
{
    uint32_t acc = 0;
    if (n <= 0) return 0u;
    int q0 = a >> 3, q1 = (a + n - 1) >> 3;
    NB_NOUNROLL
    for (int q = q0 + nbw_lane(); q <= q1; q += NB_LANES) {
        NbChunk v = *nbc_chunk(c, carrier, q);
        int lo = q << 3, first = a - lo, last = a + n - lo;
        if (first <= 0 && last >= 8) acc |= v.w[0] | v.w[1] | v.w[2] | v.w[3];
        else {
            NB_NOUNROLL
            for (int w = 0; w < 4; ++w) acc |= v.w[w] & nbc_word_mask(w, first, last);
        }
    }
    return acc;
}

#if NB_LANES == 1
/* One lane owns the cell (thread-per-cell kernel, 1-lane CPU build): four columns per 8-byte word, ragged ends one by one. The ring
 * is 8-byte aligned and W is a multiple of 4, so a word whose first column index is a multiple of 4 never wraps.
 * Long windows -- a high-CE transmission is hundreds to thousands of subframes -- would keep ONE lane of a warp scanning while the
 * other 31 cells wait (19 % of the executed instructions at 1.4 active lanes, profiles/r02d_*), so the thread-per-cell kernel keeps
 * BLOCK SUMMARIES next to the ring: c.blk[carrier][b] = OR of the 32 columns of ring block b. They are derived data: rebuilt from the
 * ring at the start of a launch (nbc_blk_rebuild), maintained by every fill, never part of the state (the warp-per-cell kernel, which
 * scans 32 columns per step, does not know them). A window then costs its two ragged ends plus one summary per whole block. */
typedef uint64_t __attribute__((may_alias)) nb_u64a;
typedef struct __attribute__((aligned(16), may_alias)) { uint64_t lo, hi; } nb_u128a;
NB_DEV uint32_t nbc_cols_or(const uint16_t *g, int wm, int k, int e)
{
    uint32_t v = 0; uint64_t v8 = 0;
    NB_NOUNROLL
    while (k < e && (k & 3)) { v |= g[k & wm]; ++k; }
#ifndef NB_SCAN_8B
    /* eight columns per 16-byte load once the index is a multiple of 8 (the ring is 16-byte aligned, W a multiple of 8) */
    if ((k & 4) && k + 4 <= e) { v8 = *(const nb_u64a *)&g[k & wm]; k += 4; }
    NB_NOUNROLL
    for (; k + 8 <= e; k += 8) { const nb_u128a q = *(const nb_u128a *)&g[k & wm]; v8 |= q.lo | q.hi; }
#endif
    NB_NOUNROLL
    for (; k + 4 <= e; k += 4) v8 |= *(const nb_u64a *)&g[k & wm];
    NB_NOUNROLL
    while (k < e) { v |= g[k & wm]; ++k; }
    v8 |= v8 >> 32;
    v |= (uint32_t)v8 | ((uint32_t)v8 >> 16);
    return v & 0xFFFFu;
}
NB_FN1S uint32_t nbc_grid_busy(NbCtx &c, int carrier, int a, int n)
{
    const uint16_t *g = nbc_col(c, carrier, 0);
    const int wm = c.W - 1, e = a + n;
    if (!c.blk || n < 64) return nbc_cols_or(g, wm, a, e);
    const uint16_t *B = c.blk + carrier * (c.W >> 5);
    const int bm = (c.W >> 5) - 1, hb = (a + 31) & ~31, te = e & ~31;      /* n >= 64: at least one whole block, hb < te */
    uint32_t v = nbc_cols_or(g, wm, a, hb) | nbc_cols_or(g, wm, te, e);
    NB_NOUNROLL
    for (int b = hb >> 5; b < (te >> 5); ++b) v |= B[b & bm];
    return v & 0xFFFFu;
}
NB_FNM3 void nbc_grid_update(NbCtx &c, int carrier, int a, int n, uint32_t mask, int set)
{
    uint16_t *g = nbc_col(c, carrier, 0);
    const int wm = c.W - 1;
    int k = a, e = a + n;
    const uint64_t m4 = (uint64_t)(mask & 0xFFFFu) * 0x0001000100010001ULL;
    NB_NOUNROLL
    while (k < e && (k & 3)) { g[k & wm] = set ? (uint16_t)(g[k & wm] | mask) : (uint16_t)0; ++k; }
    if (set) {
        NB_NOUNROLL
        for (; k + 4 <= e; k += 4) { nb_u64a *p = (nb_u64a *)&g[k & wm]; *p = *p | m4; }
    } else {
        NB_NOUNROLL
        for (; k + 4 <= e; k += 4) *(nb_u64a *)&g[k & wm] = 0ULL;
    }
    NB_NOUNROLL
    while (k < e) { g[k & wm] = set ? (uint16_t)(g[k & wm] | mask) : (uint16_t)0; ++k; }
    if (c.blk && n > 0) {
        uint16_t *B = c.blk + carrier * (c.W >> 5);
        const int bm = (c.W >> 5) - 1;
        if (set) {                                             /* every block the fill touches gains the mask */
            NB_NOUNROLL
            for (int b = a >> 5; b <= ((e - 1) >> 5); ++b) B[b & bm] = (uint16_t)(B[b & bm] | mask);
        } else {                                               /* freshly generated columns: a block restarts when its first column does */
            NB_NOUNROLL
            for (int b = (a + 31) >> 5; b <= ((e - 1) >> 5); ++b) B[b & bm] = (uint16_t)0;
        }
    }
}
/* block summaries of the live part of the ring, [t_now, t_ahead): oldest block first, so that when the live range wraps onto the slot
 * of its own first (partial, never whole inside a window) block the newer one wins */
NB_FN void nbc_blk_rebuild(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    if (!c.blk || h->t_ahead <= h->t_now) return;
    const int wm = c.W - 1, bm = (c.W >> 5) - 1;
    NB_NOUNROLL
    for (int cc = 0; cc < NB_NC(c); ++cc) {
        const uint16_t *g = nbc_col(c, cc, 0);
        uint16_t *B = c.blk + cc * (c.W >> 5);
        NB_NOUNROLL
        for (int b = h->t_now >> 5; b <= ((h->t_ahead - 1) >> 5); ++b) {
            int lo = b << 5, hi = lo + 32;
            if (lo < h->t_now) lo = h->t_now;
            if (hi > h->t_ahead) hi = h->t_ahead;
            B[b & bm] = (uint16_t)nbc_cols_or(g, wm, lo, hi);
        }
    }
}
#elif NB_GRID_MODE == 0
NB_FN1S uint32_t nbc_grid_busy(NbCtx &c, int carrier, int a, int n)
{
    return (uint32_t)nbw_run([&]() -> int {
        uint32_t v = 0;
        const uint16_t *g = nbc_col(c, carrier, 0);            /* column 0 of the carrier; the ring mask is loop-invariant */
        const int wm = c.W - 1;
        NB_NOUNROLL
        for (int k = a + nbw_lane(); k < a + n; k += NB_LANES) v |= g[k & wm];
        return (int)nbw_or_u32(v);
    });
}
NB_FNM3 void nbc_grid_update(NbCtx &c, int carrier, int a, int n, uint32_t mask, int set)
{
    nbw_run([&]() -> int {
        uint16_t *g = nbc_col(c, carrier, 0);
        const int wm = c.W - 1;
        if (set) {
            NB_NOUNROLL
            for (int k = a + nbw_lane(); k < a + n; k += NB_LANES) g[k & wm] = (uint16_t)(g[k & wm] | mask);
        } else {
            NB_NOUNROLL
            for (int k = a + nbw_lane(); k < a + n; k += NB_LANES) g[k & wm] = (uint16_t)0;
        }
        nbw_sync();
        return 0;
    });
}
#else
/* OR of the column masks of [a, a+n), n >= 1: whole 32-column blocks come from the block summaries,
 * the ragged head and tail are read chunk-wise, so the cost does not grow with the window length */
NB_FN1S uint32_t nbc_grid_busy(NbCtx &c, int carrier, int a, int n)
{
    NbHdr *h = nbc_hdr(c);
    int nblk = c.W >> 5;
    return (uint32_t)nbw_run([&]() -> int {
        int e = a + n, hb = (a + 31) & ~31, te = e & ~31;
        uint32_t acc;
        if (hb >= te) acc = nbc_scan_chunks(c, carrier, a, n);
        else {
            acc = nbc_scan_chunks(c, carrier, a, hb - a) | nbc_scan_chunks(c, carrier, te, e - te);
#if NB_GRID_MODE == 2
            NB_NOUNROLL
            for (int b = (hb >> 5) + nbw_lane(); b < (te >> 5); b += NB_LANES) acc |= h->blk[carrier][b & (nblk - 1)];
#else
            acc |= nbc_scan_chunks(c, carrier, hb, te - hb);
#endif
        }
        acc = (acc | (acc >> 16)) & 0xFFFu;
        return (int)nbw_or_u32(acc);
    });
}

/* columns [a, a+n): OR `mask` in (set = 1) or start them empty (set = 0, only used for freshly generated columns) */
NB_FNM3 void nbc_grid_update(NbCtx &c, int carrier, int a, int n, uint32_t mask, int set)
{
    if (n <= 0) return;
    NbHdr *h = nbc_hdr(c);
    int nblk = c.W >> 5;
    nbw_run([&]() -> int {
        uint32_t m2 = mask | (mask << 16);
        int q0 = a >> 3, q1 = (a + n - 1) >> 3;
        NB_NOUNROLL
        for (int q = q0 + nbw_lane(); q <= q1; q += NB_LANES) {
            NbChunk *p = nbc_chunk(c, carrier, q);
            int lo = q << 3, first = a - lo, last = a + n - lo;
            NbChunk v = *p;
            NB_NOUNROLL
            for (int w = 0; w < 4; ++w) {
                uint32_t wm = (first <= 0 && last >= 8) ? 0xFFFFFFFFu : nbc_word_mask(w, first, last);
                v.w[w] = set ? (v.w[w] | (m2 & wm)) : (v.w[w] & ~wm);
            }
            *p = v;
        }
#if NB_GRID_MODE == 2
        /* block summaries: a block restarts empty when its first column is generated; marking ORs the mask in */
        int b0 = a >> 5, b1 = (a + n - 1) >> 5;
        NB_NOUNROLL
        for (int b = b0 + nbw_lane(); b <= b1; b += NB_LANES) {
            uint16_t *s = &h->blk[carrier][b & (nblk - 1)];
            if (set) *s = (uint16_t)(*s | mask);
            else if ((b << 5) >= a) *s = 0;
        }
#endif
        nbw_sync();
        return 0;
    });
}

#endif

NB_DEV void nbc_grid_or(NbCtx &c, int carrier, int from, int count, uint32_t mask) { nbc_grid_update(c, carrier, from, count, mask, 1); }

/* refresh the fixed event slots of one CE level from its occasion ring */
NB_DEV void nbc_occ_keys(NbHdr *h, int ce)
{
    uint64_t ks = NB_KEY_NONE, ke = NB_KEY_NONE;
    if (h->occ_head_start[ce] < h->occ_tail[ce]) { NbOccasion o = h->occ[ce][h->occ_head_start[ce] % NB_OCC_RING]; ks = NB_KEY(o.t_start, o.seq); }
    if (h->occ_head_end[ce] < h->occ_tail[ce]) { NbOccasion o = h->occ[ce][h->occ_head_end[ce] % NB_OCC_RING]; ke = NB_KEY(o.t_start + o.N_sf, o.seq + 1u); }
    h->keys[2 + ce] = ks; h->keys[5 + ce] = ke;
}

/* NprachResource next-start search + mutation at an idle column (carrier.py:301-332) */
NB_DEV int nbc_res_next_start(NbResource &r, int col)
{
    int tau0 = col - r.t_offset;
    if (tau0 > r.t_next) {                       /* t_next was reduced or the offset moved: re-sync */
        r.t_next = r.t_start + r.periodicity;
        NB_NOUNROLL
        while (r.t_next < tau0) r.t_next += r.periodicity;
    }
    int cand = r.t_next;
    if (tau0 <= 0 && 0 < cand) cand = 0;         /* `not t`: a resource also starts at local time 0 */
    return cand + r.t_offset;
}

/* CElevelSClog (carrier.py:254-271): the reference keeps {t: subcarriers per CE level} for every non-anchor NPRACH start it ever
 * generated and asks for it once, when the anchor NPRACH of time t starts. Entries whose time the clock has passed are never read
 * again, so a ring in the cell's cold block holds the live ones: up to NB_SCLOG_CAP (a long look-ahead over short periods
 * holds hundreds), faulting loudly beyond. Only the FIRST registration of a time counts, as in the reference (:262-265). */
NB_DEV NbSclogEntry *nbc_sclog_ring(NbCtx &c)
{
    const NbColdConf *cold = (const NbColdConf *)&((const NbCtaConst *)c.ctrl)->cold;
    return cold->sclog_base + (size_t)(c.occ - cold->occ_base) * (size_t)NB_SCLOG_CAP;
}

NB_FN void nbc_sclog_register(NbCtx &c, int t, int ce, int sc)   /* carrier.py:262-265 */
{
    NbHdr *h = nbc_hdr(c);
    NbSclogEntry *ring = nbc_sclog_ring(c);
    NB_NOUNROLL
    while (h->sclog_lo < h->sclog_n && ring[h->sclog_lo % NB_SCLOG_CAP].t < h->time) h->sclog_lo += 1;      /* passed by the clock */
    NB_NOUNROLL
    for (int i = h->sclog_lo; i < h->sclog_n; ++i) if (ring[i % NB_SCLOG_CAP].t == t) return;
    if (h->sclog_n - h->sclog_lo >= NB_SCLOG_CAP) { NB_FAULT(c, NBIOT_FAULT_LOOKAHEAD); return; }
    NbSclogEntry e; e.t = t; e.sc[0] = e.sc[1] = e.sc[2] = e.sc[3] = 0; e.sc[ce] = (uint8_t)sc;
    ring[h->sclog_n % NB_SCLOG_CAP] = e;
    h->sclog_n += 1;
}

NB_FN int nbc_sclog_check(NbCtx &c, int t, int ce)               /* carrier.py:267-271, :454-459 */
{
    NbHdr *h = nbc_hdr(c);
    if (NB_NC(c) == 1 || t == 0) return 0;
    const NbSclogEntry *ring = nbc_sclog_ring(c);
    NB_NOUNROLL
    for (int i = h->sclog_lo; i < h->sclog_n; ++i) if (ring[i % NB_SCLOG_CAP].t == t) return ring[i % NB_SCLOG_CAP].sc[ce];
    return 0;
}

/* Carrier.extend_lookahead (carrier.py:473-479) with SFGenerator.next (:382-404) run by run */
NB_FN void nbc_extend_lookahead(NbCtx &c, int new_t_ahead)
{
    NbHdr *h = nbc_hdr(c);
    int a = h->t_ahead, b = new_t_ahead;
    if (b <= a) return;
    if (b - h->t_now > h->max_lookahead) h->max_lookahead = b - h->t_now;
    if (b - h->t_now > c.W) { NB_FAULT(c, NBIOT_FAULT_LOOKAHEAD); return; }
    /* fresh columns start empty */
    NB_NOUNROLL
    for (int cc = 0; cc < NB_NC(c); ++cc) nbc_grid_update(c, cc, a, b - a, 0u, 0);
    int cursor[2][3];
    NB_NOUNROLL
    for (int cc = 0; cc < NB_NC(c); ++cc)
        NB_SMALL_LOOP
        for (int ce = 0; ce < 3; ++ce) {
            NbResource &r = h->res[cc][ce];
            cursor[cc][ce] = a;
            if (r.N_sc && r.sf_to_go > 0) {      /* an NPRACH that is still running */
                int len = r.sf_to_go < b - a ? r.sf_to_go : b - a;
                nbc_grid_or(c, cc, a, len, r.mask);
                r.sf_to_go -= len;
                cursor[cc][ce] = a + len;
            }
        }
    for (;;) {
        /* earliest start in (column, carrier, CE) order */
        int best = b, bc = -1, bce = -1;
        NB_NOUNROLL
        for (int cc = 0; cc < NB_NC(c); ++cc)
            NB_SMALL_LOOP
            for (int ce = 0; ce < 3; ++ce) {
                NbResource &r = h->res[cc][ce];
                if (!r.N_sc || cursor[cc][ce] >= b) continue;
                int s = nbc_res_next_start(r, cursor[cc][ce]);
                if (s < best) { best = s; bc = cc; bce = ce; }
            }
        if (bc < 0) break;
        NbResource &r = h->res[bc][bce];
        int tau = best - r.t_offset;
        r.mask = (uint16_t)((((1u << r.N_sc) - 1u) << r.sc_offset) & 0xFFFu);
        r.t_start = tau;
        r.t_next = tau + r.periodicity;
        int len = r.N_sf < b - best ? r.N_sf : b - best;
        nbc_grid_or(c, bc, best, len, r.mask);
        r.sf_to_go = r.N_sf - len;
        cursor[bc][bce] = best + len;
        if (bc == 0) {                           /* NPRACH_start @ best (seq), NPRACH_end @ best + N_sf (seq+1) */
            if (h->occ_tail[bce] - h->occ_head_end[bce] >= NB_OCC_RING) { NB_FAULT(c, NBIOT_FAULT_LOOKAHEAD); return; }
            NbOccasion o; o.t_start = best; o.seq = h->seq; o.N_sf = (uint16_t)r.N_sf; o.N_sc = r.N_sc; o.pad = 0;
            h->occ[bce][h->occ_tail[bce] % NB_OCC_RING] = o;
            h->occ_tail[bce] += 1;
            h->seq += 2;
            nbc_occ_keys(h, bce);
        } else nbc_sclog_register(c, best, bce, 4 * r.N_sc);
    }
    h->t_ahead = b;
}

NB_DEV void nbc_erase_past_sf(NbCtx &c, int t)                   /* carrier.py:481-489 */
{
    NbHdr *h = nbc_hdr(c);
    if (t > h->t_now) h->t_now = t;
    if (t + NB_HORIZON > h->t_ahead) nbc_extend_lookahead(c, t + NB_HORIZON);   /* the usual case needs no call */
}

/* Carrier.allocate_resources (carrier.py:556-605): first fit over delays x carriers x offsets */
NB_FNM4 int nbc_carrier_allocate(NbCtx &c, int carrier_id, int t, int delay, int N_sc, int N_sf)
{
    NbHdr *h = nbc_hdr(c);
    nbc_erase_past_sf(c, t);
    if (NB_FATAL(c)) return 0;
    const int row = N_sc == 1 ? 0 : N_sc == 3 ? 1 : N_sc == 6 ? 2 : 3;      /* once per call: the loops below only index with it */
    NB_COUNT(0, 1); NB_COUNT(5, t + 64 + N_sf <= h->t_ahead);
    NB_NOUNROLL
    for (int d = delay; d <= 64; d *= 2) {
        int new_t_ahead = t + d + N_sf;
        NB_COUNT(1, 1); NB_COUNT(2, N_sf); NB_COUNT(6, d == delay);
        if (new_t_ahead > h->t_ahead) {
            NB_COUNT(3, 1);
            nbc_extend_lookahead(c, new_t_ahead);
            if (NB_FATAL(c)) return 0;
        }
        int from = h->t_now + d;
        NB_NOUNROLL
        for (int ci = 0; ci < NB_NC(c); ++ci) {
            int cc = ci == 0 ? carrier_id : (carrier_id == 0 ? 1 : 0);
#ifndef NB_NO_EDGE_REJECT
            /* the subcarriers busy in the first or the last column of the window are busy in the window: when they already leave
             * no aligned group free -- the usual case in a loaded cell, where transmissions longer than the delays cover the window
             * starts -- the window is lost without looking at the columns in between */
            {
                uint32_t eb = (uint32_t)*nbc_col(c, cc, from) | (uint32_t)*nbc_col(c, cc, from + N_sf - 1);
                uint32_t ef = eb;
                if (row >= 1) ef |= (eb >> 1) | (eb >> 2);
                if (row >= 2) ef |= ef >> 3;
                if (row == 3) ef |= ef >> 6;
                if (!(~ef & nb_sc_heads[row])) { NB_COUNT(11, 1); continue; }
            }
#endif
            uint32_t busy = nbc_grid_busy(c, cc, from, N_sf);
            /* first free aligned group of N_sc subcarriers (offsets_per_sc, carrier.py:29-34): fold the
             * group's bits onto its lowest one, then take the lowest clear group bit */
            uint32_t fold = busy;
            if (row >= 1) fold |= (busy >> 1) | (busy >> 2);
            if (row >= 2) fold |= fold >> 3;
            if (row == 3) fold |= fold >> 6;
            uint32_t pick = ~fold & nb_sc_heads[row];
            if (pick) {
                NB_COUNT(4, 1); NB_COUNT(7, d == delay);
                nbc_grid_or(c, cc, from, N_sf, ((1u << N_sc) - 1u) << (nbw_ffs(pick) - 1));
                return new_t_ahead;
            }
        }
    }
    return 0;
}

/* NodeB.allocate_resources (node_b.py:435-450): retry with other subcarrier counts */
NB_FNM3 int nbc_node_allocate(NbCtx &c, int c_id, int ref_t, int delay, int N_sc, int N_ru, int N_rep, int *out_sc, int *out_sf)
{
    int spr = nbc_sf_per_ru(N_sc);
    if (spr < 0) { NB_FAULT(c, NBIOT_FAULT_ACTION); *out_sc = N_sc; *out_sf = 0; return 0; }
    int n_sc = N_sc, N_sf = spr * N_ru * N_rep;
    NB_COUNT(8, 1);
#if NB_LANES == 32 && !defined(NB_NO_EDGE_REJECT) && defined(NB_FAST_FIT)
    /* All (subcarrier count, delay, carrier) combinations at once. The search below tries up to 4 x 4 x C windows one after the
     * other and, in a loaded cell, loses most of them on the edge test alone (the subcarriers busy in the first or last column of
     * a window leave no aligned group free). When every window of the search already lies inside the generated look-ahead, no
     * combination can extend it, so the outcome of each is a function of the ring alone: one lane per combination, in the order
     * of the search, takes the edge test; the ballot names the windows that still need their columns scanned, and they are
     * scanned in that order. An attempt that fails everywhere -- two thirds of them on the NPRACH-control workload -- costs one
     * pass instead of sixteen. */
    {
        NbHdr *h = nbc_hdr(c);
        nbc_erase_past_sf(c, ref_t);
        if (NB_FATAL(c)) { *out_sc = N_sc; *out_sf = N_sf; return 0; }
        const int row0 = N_sc == 1 ? 0 : N_sc == 3 ? 1 : N_sc == 6 ? 2 : 3;
        const int n_try = c.conf.sc_adjustment ? 4 : 1;
        const int sc_min = n_try == 4 ? (nb_sc_fallback[row0][2] < N_sc ? nb_sc_fallback[row0][2] : N_sc) : N_sc;
        const int t_now = h->t_now;
        if (delay >= 1 && t_now + 64 + nbc_sf_per_ru(sc_min) * N_ru * N_rep <= h->t_ahead) {
            const int unit = N_ru * N_rep;
            NB_COUNT(12, 1);
            uint32_t live = (uint32_t)nbw_run([&]() -> int {
                const int l = nbw_lane(), s = l >> 3, k = (l >> 1) & 3, ci = l & 1;
                const int nsc = s == 0 ? N_sc : nb_sc_fallback[row0][s > 0 ? s - 1 : 0];
                const int d = delay << k;
                int ok = s < n_try && d <= 64 && ci < NB_NC(c);
                if (ok) {
                    const int row = nsc == 1 ? 0 : nsc == 3 ? 1 : nsc == 6 ? 2 : 3;
                    const int cc = ci == 0 ? c_id : (c_id == 0 ? 1 : 0), from = t_now + d, nsf = nbc_sf_per_ru(nsc) * unit;
                    uint32_t ef = (uint32_t)*nbc_col(c, cc, from) | (uint32_t)*nbc_col(c, cc, from + nsf - 1);
                    if (row >= 1) ef |= (ef >> 1) | (ef >> 2);
                    if (row >= 2) ef |= ef >> 3;
                    if (row == 3) ef |= ef >> 6;
                    ok = (~ef & nb_sc_heads[row]) != 0u;
                }
                return (int)nbw_ballot(ok);
            });
            NB_COUNT(14, live == 0u);
            NB_NOUNROLL
            while (live) {
                const int l = nbw_ffs(live) - 1; live &= live - 1u;
                NB_COUNT(15, 1);
                const int s = l >> 3, k = (l >> 1) & 3, ci = l & 1;
                const int nsc = s == 0 ? N_sc : nb_sc_fallback[row0][s - 1];
                const int row = nsc == 1 ? 0 : nsc == 3 ? 1 : nsc == 6 ? 2 : 3;
                const int cc = ci == 0 ? c_id : (c_id == 0 ? 1 : 0), d = delay << k, from = t_now + d, nsf = nbc_sf_per_ru(nsc) * unit;
                uint32_t busy = nbc_grid_busy(c, cc, from, nsf), fold = busy;
                if (row >= 1) fold |= (busy >> 1) | (busy >> 2);
                if (row >= 2) fold |= fold >> 3;
                if (row == 3) fold |= fold >> 6;
                uint32_t pick = ~fold & nb_sc_heads[row];
                if (pick) {
                    nbc_grid_or(c, cc, from, nsf, ((1u << nsc) - 1u) << (nbw_ffs(pick) - 1));
                    *out_sc = nsc; *out_sf = nsf;
                    NB_COUNT(13, 1);
                    return ref_t + d + nsf;
                }
            }
            n_sc = n_try == 4 ? nb_sc_fallback[row0][2] : N_sc;
            *out_sc = n_sc; *out_sf = nbc_sf_per_ru(n_sc) * unit;
            return 0;
        }
    }
#endif
    int t_ul_end = nbc_carrier_allocate(c, c_id, ref_t, delay, N_sc, N_sf);
    NB_COUNT(9, t_ul_end != 0);
    if (!t_ul_end && c.conf.sc_adjustment && !NB_FATAL(c)) {
        int row = N_sc == 1 ? 0 : N_sc == 3 ? 1 : N_sc == 6 ? 2 : 3;
        NB_NOUNROLL
        for (int i = 0; i < 3; ++i) {
            n_sc = nb_sc_fallback[row][i];
            N_sf = nbc_sf_per_ru(n_sc) * N_ru * N_rep;
            t_ul_end = nbc_carrier_allocate(c, c_id, ref_t, delay, n_sc, N_sf);
            if (t_ul_end || NB_FATAL(c)) break;
        }
    }
    *out_sc = n_sc; *out_sf = N_sf;
    return t_ul_end;
}

/* compute_offsets (carrier.py:115-252): time/frequency packing of the three NPRACH resources */
NB_FN1 int nbc_compute_offsets(const int periods[3], const int N_sfs[3], const int N_scs[3], int sc_off[3], int sf_off[3])
{
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) sc_off[i] = sf_off[i] = 0;
    int min_length = N_sfs[0] + N_sfs[1] + N_sfs[2];
    int crit[3], ncrit = 0, norm[3], nnorm = 0;
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) { if (periods[i] == 240) crit[ncrit++] = i; else norm[nnorm++] = i; }
    int max_sf = N_sfs[0] > N_sfs[1] ? N_sfs[0] : N_sfs[1]; if (N_sfs[2] > max_sf) max_sf = N_sfs[2];
    int sum_sc = N_scs[0] + N_scs[1] + N_scs[2];
    if (ncrit == 1 && max_sf > 80) {
        int i = crit[0], j = norm[0], k = norm[1];
        int mjk = N_scs[j] > N_scs[k] ? N_scs[j] : N_scs[k];
        if (N_scs[i] + mjk > 12) return 0;
        if (sum_sc > 12) {
            int s = N_sfs[j] + N_sfs[k]; min_length = s > N_sfs[i] ? s : N_sfs[i];
            sc_off[j] = N_scs[i]; sc_off[k] = N_scs[i]; sf_off[k] = N_sfs[j];
        } else { min_length = N_sfs[2]; sc_off[j] = N_scs[i]; sc_off[k] = N_scs[i] + N_scs[j]; }
    } else if (ncrit == 2 && max_sf > 80) {
        int i = crit[0], j = crit[1], k = norm[0];
        int mij = N_scs[i] > N_scs[j] ? N_scs[i] : N_scs[j];
        if (N_scs[k] + mij > 12) return 0;
        if (sum_sc > 12) {
            int s = N_sfs[i] + N_sfs[j]; min_length = s > N_sfs[k] ? s : N_sfs[k];
            sc_off[k] = mij; sf_off[j] = N_sfs[i];
        } else { min_length = N_sfs[2]; sc_off[j] = N_scs[i]; sc_off[k] = N_scs[i] + N_scs[j]; }
    } else if (sum_sc > 12) {
        int m10 = N_scs[1] > N_scs[0] ? N_scs[1] : N_scs[0];
        if (N_scs[2] + m10 <= 12) {
            min_length = N_sfs[2] + N_sfs[0]; sc_off[0] = N_scs[2]; sc_off[1] = N_scs[2]; sf_off[0] = N_sfs[1];
        } else if (N_scs[2] + N_scs[1] <= 12) {
            min_length = N_sfs[2] + N_sfs[0]; sc_off[1] = N_scs[2]; sf_off[0] = N_sfs[2];
        } else if (N_scs[1] + N_scs[0] <= 12) {
            min_length = N_sfs[2] + N_sfs[1]; sc_off[0] = N_scs[1]; sf_off[0] = N_sfs[2]; sf_off[1] = N_sfs[2];
        } else if (N_scs[2] + N_scs[0] <= 12) {
            min_length = N_sfs[2] + N_sfs[1]; sc_off[0] = N_scs[2]; sf_off[1] = N_sfs[2];
        } else {
            int a[3] = { 0, 1, 2 };              /* stable sort by (period, N_sc, N_sf) */
            NB_NOUNROLL
            for (int x = 1; x < 3; ++x)
                NB_NOUNROLL
                for (int y = x; y > 0; --y) {
                    int p = a[y - 1], q = a[y];
                    int gt = periods[p] != periods[q] ? periods[p] > periods[q]
                           : N_scs[p] != N_scs[q] ? N_scs[p] > N_scs[q] : N_sfs[p] > N_sfs[q];
                    if (gt) { a[y - 1] = q; a[y] = p; } else break;
                }
            sf_off[a[1]] = N_sfs[a[0]]; sf_off[a[2]] = N_sfs[a[0]] + N_sfs[a[1]];
        }
    } else {
        int s = N_sfs[1] + N_sfs[0]; min_length = N_sfs[2] > s ? N_sfs[2] : s;
        sc_off[0] = N_scs[2]; sc_off[1] = N_scs[2];
        if (N_scs[0] <= N_scs[1]) sf_off[0] = N_sfs[1]; else sf_off[1] = N_sfs[0];
    }
    int min_periodicity = -1;
    NB_NOUNROLL
    for (int i = 0; i < 7; ++i)
        if (nb_nprach_period_list[i] >= min_length && (min_periodicity < 0 || nb_nprach_period_list[i] < min_periodicity))
            min_periodicity = nb_nprach_period_list[i];
    int minp = periods[0] < periods[1] ? periods[0] : periods[1]; if (periods[2] < minp) minp = periods[2];
    return !(min_periodicity < 0 || minp < min_periodicity);
}

/* Carrier.update_ra_parameters (carrier.py:610-650) + SFGenerator.update (:377-380) + NprachResource.update (:345-355) */
NB_FN1 void nbc_carrier_update_ra(NbCtx &c, const int args[3][3], int th_C1, int th_C0)
{
    NbHdr *h = nbc_hdr(c);
    int periods[3], N_sfs[3], N_scs[3];
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) {
        periods[i] = nb_nprach_period_list[args[i][0]];
        N_sfs[i] = nb_nprach_nsf_list[args[i][1]];
        N_scs[i] = (12 * (args[i][2] + 1)) / 4;      /* NPRACH_N_sc_list[n_c-1][idx] // 4 (parameters.py:327-330) */
    }
    if (th_C0 == 0) { periods[0] = nb_nprach_period_list[0]; N_sfs[0] = 0; N_scs[0] = 0; }
    if (th_C1 == 0) { periods[1] = nb_nprach_period_list[0]; N_sfs[1] = 0; N_scs[1] = 0; }
    int per_carrier[2][3];
    NB_NOUNROLL
    for (int cc = 0; cc < NB_NC(c); ++cc)
        NB_NOUNROLL
        for (int i = 0; i < 3; ++i) {
            int take = N_scs[i] < 12 ? N_scs[i] : 12;
            per_carrier[cc][i] = take;
            N_scs[i] = N_scs[i] - take;
        }
    int sc_off[3], sf_off[3];
    (void)nbc_compute_offsets(periods, N_sfs, per_carrier[0], sc_off, sf_off);   /* validity ignored (carrier.py:638) */
    NB_NOUNROLL
    for (int cc = 0; cc < NB_NC(c); ++cc)
        NB_NOUNROLL
        for (int ce = 0; ce < 3; ++ce) {
            int rt = h->res[cc][0].t_start;
            if (h->res[cc][1].t_start > rt) rt = h->res[cc][1].t_start;
            if (h->res[cc][2].t_start > rt) rt = h->res[cc][2].t_start;
            NbResource &r = h->res[cc][ce];
            r.periodicity = (int16_t)periods[ce]; r.N_sf = (int16_t)N_sfs[ce];
            r.N_sc = (uint8_t)per_carrier[cc][ce]; r.sc_offset = (uint8_t)sc_off[ce];
            r.t_next = r.t_start + periods[ce];
            if (r.t_next <= rt) r.t_next = rt + periods[ce];
            r.t_offset = sf_off[ce];
        }
}

/* ------------------------------------------------------------------ NodeB helpers */
NB_DEV int nbc_rar_sum(NbCtx &c, int ce) { return nbc_hdr(c)->rar_tot[ce]; }   /* sum(RAR_UEs[ce]), kept incrementally */

NB_FNM void nbc_update_state(NbCtx &c)                           /* node_b.py:264-276 */
{
    NbHdr *h = nbc_hdr(c);
    if (nbc_rar_sum(c, 0) + nbc_rar_sum(c, 1) + nbc_rar_sum(c, 2)) h->state = NB_ST_RAR_WINDOW;
    else if (h->conn_n > 0) h->state = NB_ST_SCHEDULING;
    else h->state = NB_ST_IDLE;
}

NB_FNM void nbc_schedule_NPDCCH(NbCtx &c, int elapsed_sfs)       /* node_b.py:420-432 */
{
    NbHdr *h = nbc_hdr(c);
    int left = h->NPDCCH_sf_left - elapsed_sfs; if (left < 0) left = 0;
    int next_time = h->time + elapsed_sfs;
    if (!left) { next_time += NB_NPDCCH_PERIOD; left = NB_NPDCCH_SF; }
    h->NPDCCH_sf_left = left;
    h->keys[0] = NB_KEY(next_time, h->seq);
    h->seq += 1;
}

/* ------------------------------------------------------------------ random-access lists */
/* order-preserving removal of tombstones */
NB_FN void nbc_ra_compact(NbCtx &c, int ce)
{
    NbHdr *h = nbc_hdr(c);
    int n = h->ra_n[ce];
    NbRaEntry *L = c.ra[ce];
    int out = nbw_run([&]() -> int {
        int o = 0;
        NB_NOUNROLL
        for (int base = 0; base < n; base += NB_LANES) {
            int i = base + nbw_lane();
            NbRaEntry e; e.st_ce = NB_RA_TOMB;
            if (i < n) e = L[i];
            int keep = i < n && e.st_ce != NB_RA_TOMB;
            uint32_t mk = nbw_ballot(keep);
            nbw_sync();
            if (keep) L[o + nbw_popc(mk & nbw_lt_mask())] = e;
            o += nbw_popc(mk);
            nbw_sync();
        }
        return o;
    });
    h->ra_n[ce] = out;
}

NB_FNM3 void nbc_ra_append(NbCtx &c, int ce, int slot, int state, int ue_ce, int attempts, int rar_w_id)
{
    NbHdr *h = nbc_hdr(c);
    if (h->ra_n[ce] >= c.ra_cap) {
        nbc_ra_compact(c, ce);
        if (h->ra_n[ce] >= c.ra_cap) { NB_FAULT(c, NBIOT_FAULT_UE_SLOTS); return; }
    }
    NbRaEntry e; e.wake_t = 0; e.wake_seq = 0; e.rar_w_id = rar_w_id; e.slot = (uint16_t)slot;
    e.st_ce = (uint8_t)(state | (ue_ce << 4)); e.attempts = (uint8_t)attempts;
    c.ra[ce][h->ra_n[ce]] = e;
    h->ra_n[ce] += 1;
}

/* ------------------------------------------------------------------ traffic + placement */
NB_DEV double nbc_find_y(double x1, double y1, double x2, double y2, double x)   /* channel.py:82-86 */
{
    double m = NB_DIV(NB_SUB(y2, y1), NB_SUB(x2, x1));
    double b = NB_ADD(NB_MUL(-m, x1), y1);
    return NB_ADD(NB_MUL(m, x), b);
}

#define NB_S34 0.4330127018922193
#define NB_S32 0.8660254037844386

/* Channel.set_xy_loss / generate_xy / location / antenna_pattern (channel.py:100-118, :134-151) */
NB_DEV double nbc_place_ue(NbCtx &c, double x, double y, double *xy_out)
{
    if (x < 0) for (;;) {                        /* generate_xy; also when a clustered x falls left of the cell (channel.py:143-144) */
        nb_rng_pair(nbc_hdr(c)->seed, nbc_hdr(c)->draws++, NBIOT_STREAM_ENV, &x, &y);
        int in_cell = (y > nbc_find_y(0, NB_S34, 0.25, 0, x)) && (y > nbc_find_y(0.75, 0, 1, NB_S34, x)) &&
                      (y < nbc_find_y(0, NB_S34, 0.25, NB_S32, x)) && (y < nbc_find_y(0.75, NB_S32, 1, NB_S34, x)) && (y < NB_S32);
        if (in_cell) break;
    }
    if (xy_out) { xy_out[0] = x; xy_out[1] = y; }
    double x_t = NB_SUB(x, 0.25);
    double distance = NB_SQRT(NB_ADD(NB_MUL(x_t, x_t), NB_MUL(y, y)));
    double cos_theta = NB_DIV(x_t, distance);
    double theta = NB_SUB(NB_MUL(nb_acos(cos_theta), NB_RAD2DEG), 60.0);
    double range = c.conf.max_d > 0 ? c.conf.max_d : NB_MAX_RANGE;
    double R = NB_MUL(distance, range); if (!(R > 0.1)) R = 0.1;
    double q = NB_DIV(theta, 65.0);
    double ap = NB_MUL(12.0, NB_MUL(q, q)); if (!(ap < 20.0)) ap = 20.0;
    double G = NB_ADD(NB_GMAX, NB_MUL(-1.0, ap));
    double L = NB_ADD(128.1, NB_MUL(37.6, nb_log10(R)));
    if (!(L > 0.0)) L = 0.0;
    return NB_ADD(NB_SUB(L, G), 0.0);
}

NB_FN double nbc_place_ue_cold(NbCtx &c, double x, double y, double *xy_out) { return nbc_place_ue(c, x, y, xy_out); }

/* Population.clustered_ue_generation (population.py:75-101): out of line, only clustered cells come here */
NB_FN double nbc_place_clustered(NbCtx &c)
{
    const NbColdConf *cold = nbc_cold(c);
    double p = nbc_u01(c), xy[2];
    if (!(p < cold->cluster_prob)) return nbc_place_ue_cold(c, -1.0, -1.0, nullptr);
    if (!c.occ->cluster_count) {                 /* a new cluster: its first UE is placed anywhere and becomes the centre */
        double loss = nbc_place_ue_cold(c, -1.0, -1.0, xy);
        c.occ->cluster_cx = xy[0]; c.occ->cluster_cy = xy[1];
        c.occ->cluster_count = nbc_integers(c, cold->cluster_lo, cold->cluster_hi);
        return loss;
    }
    double ua = nbc_u01(c);                      /* angle = uniform(0, 2 pi); cos / sin taken at 2 pi u */
    double radius = NB_MUL(cold->cluster_range, NB_SQRT(NB_ADD(0.0, NB_MUL(1.0, nbc_u01(c)))));
    double x = NB_ADD(NB_MUL(radius, nb_cos2pi(ua)), c.occ->cluster_cx), y = NB_ADD(NB_MUL(radius, nb_sin2pi(ua)), c.occ->cluster_cy);
    c.occ->cluster_count -= 1;
    return nbc_place_ue_cold(c, x, y, nullptr);
}

/* UEGenerator.generate_arrivals (ue_generator.py:54-81) + Population.get_new_arrivals (population.py:103-132) */
NB_FN1 void nbc_new_arrivals(NbCtx &c, int t)
{
    NbHdr *h = nbc_hdr(c);
    if (c.conf.random_traffic) {                 /* update_counter (ue_generator.py:45-52) */
        int new_count = t / NB_T_UNIFORM;
        if (new_count > h->gen_counter) { h->gen_counter = new_count; h->gen_p = NB_ADD(0.0, NB_MUL(0.8, nbc_u01(c))); }
    }
    int t_elapsed = t - h->gen_t;
    int arrivals = 0, beta_arrivals = 0;
    if (h->M_uniform > 0) {
        double intensity = NB_MUL((double)h->M_uniform, NB_DIV((double)t_elapsed, (double)NB_T_UNIFORM));
        double ip = NB_FLOOR(intensity), p = NB_SUB(intensity, ip);      /* modf, intensity >= 0 */
        arrivals = (int)ip + nbc_bernoulli(c, p);
        h->M_uniform -= arrivals;
    }
    if (h->M_beta > 0) {
        double t_s = NB_DIV((double)(h->gen_t % NB_T_UNIFORM), (double)NB_T_BETA);
        double t_e = NB_ADD(t_s, NB_DIV((double)t_elapsed, (double)NB_T_BETA));
        double intensity = NB_MUL((double)h->M_beta, NB_SUB(nb_beta34_cdf(t_e), nb_beta34_cdf(t_s)));
        double ip = NB_FLOOR(intensity), p = NB_SUB(intensity, ip);
        beta_arrivals = (int)ip + nbc_bernoulli(c, p);
        h->M_beta -= beta_arrivals;
    }
    h->gen_t = t;
    int total = arrivals + beta_arrivals;
    NB_NOUNROLL
    for (int n = 0; n < total; ++n) {
        h->ue_count += 1;
        int buffer = c.conf.buffer_hi > c.conf.buffer_lo ? nbc_integers(c, c.conf.buffer_lo, c.conf.buffer_hi + 1) : c.conf.buffer_lo;
        double loss = nbc_cold(c)->clustered ? nbc_place_clustered(c) : nbc_place_ue(c, -1.0, -1.0, nullptr);
        int beta = 0;
        if (beta_arrivals > 0) { beta = 1; beta_arrivals -= 1; }
        double P_RSRP = NB_SUB(NB_SRP, loss);
        int CE_level = P_RSRP < (double)h->CE_thresholds[0] ? 2 : (P_RSRP < (double)h->CE_thresholds[1] ? 1 : 0);
        if (h->free_top <= 0) { NB_FAULT(c, NBIOT_FAULT_UE_SLOTS); return; }
        h->free_top -= 1;
        if (c.ue_cap - h->free_top > h->max_live) h->max_live = c.ue_cap - h->free_top;
        int slot = c.free_stack[h->free_top];
        NbUe u;
        u.loss = loss; u.sinr = 0.0; u.id = h->ue_count; u.t_arrival = t; u.t_connection = 0; u.timestamp = 0; u.acked_data = 0;
        u.rar_w_id = 0; u.buffer = (uint16_t)buffer; u.retx_buffer = 0; u.bits = 0; u.tbs = 0; u.I_tbs = 0; u.N_rep = 0; u.N_ru = 0;
        u.CE_level = (uint8_t)CE_level; u.state = NB_UE_RAO; u.attempts = 0; u.beta = (uint8_t)beta; u.new_data = 1; u.where = NB_IN_RA;
        NB_NOUNROLL
        for (int k = 0; k < 7; ++k) u.pad[k] = 0;
        c.ue[slot] = u;
        nbc_ra_append(c, CE_level, slot, NB_UE_RAO, CE_level, 0, 0);
        if (NB_FATAL(c)) return;
        h->ue_arrivals[CE_level] += 1;
    }
}

#ifndef NB_NO_CAPTURE
/* ------------------------------------------------------------------ capture effect (Channel(capture_effect=True); generic builds only)
 * Plain scalar code, out of line and off the hot path: every lane of a warp-per-cell build walks it with the same values. */
/* Channel.sinr_estimation (channel.py:178-198) over the UEs in slots[0..n): one shadow-fading draw each, the strongest signal against
 * noise + the others; returns the SINR in dB and the index of the strongest */
NB_FN double nbc_sinr_estimation(NbCtx &c, const uint16_t *slots, int n, int *max_i)
{
    double sum = 0.0, max_rx = -1.0;                           /* powers are positive */
    *max_i = 0;
    NB_NOUNROLL
    for (int i = 0; i < n; ++i) {
        double LogF = nbc_normal(c, 0.0, NB_SIGMA_F);
        double rx_pw = nb_exp10(NB_DIV(NB_SUB(NB_SUB(NB_TX_PW, c.ue[slots[i]].loss), LogF), 10.0));
        sum = NB_ADD(sum, rx_pw);
        if (rx_pw > max_rx) { max_rx = rx_pw; *max_i = i; }
    }
    double ni = NB_SUB(NB_ADD(NB_NOISE_MW, sum), max_rx);
    return NB_MUL(10.0, nb_log10(NB_DIV(max_rx, ni)));
}

/* the detection pass of AccessProcedure.NPRACH_start (access_procedure.py:76-91) with Channel.preamble_detection /
 * multiple_preamble_detection (channel.py:156-176, :200-219): preamble groups in first-occurrence order; a shared preamble is detected
 * when its strongest signal is (state CAPTURE for all its UEs, else COLLIDED). c.scratch[r] = list position of contender r, its
 * preamble in the entry's wake_t. */
NB_FN void nbc_nprach_detect_capture(NbCtx &c, int ce, int n_ues, int rep)
{
    NbHdr *h = nbc_hdr(c);
    NbRaEntry *L = c.ra[ce];
    const double x0 = nb_nprach_x0[rep], kk = nb_nprach_k[rep];
    uint32_t seen[8] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
    int detected = 0, signals = 0;
    NB_NOUNROLL
    for (int r = 0; r < n_ues; ++r) {
        const int pre = (int)L[c.scratch[r]].wake_t;
        const uint32_t bit = 1u << (pre & 31);
        if (seen[(pre >> 5) & 7] & bit) continue;
        seen[(pre >> 5) & 7] |= bit;
        if (!((h->pre_twice[(pre >> 5) & 7] >> (pre & 31)) & 1u)) {                    /* alone on its preamble */
            NbRaEntry *e = &L[c.scratch[r]];
            double LogF = nbc_normal(c, 0.0, NB_SIGMA_F);
            double SINR = NB_SUB(NB_SUB(NB_SUB(NB_SUB(NB_TX_PW, c.ue[e->slot].loss), LogF), NB_NOISE_DBM), NB_NOISE_FIG);
            double p = NB_DIV(1.0, NB_ADD(1.0, nb_exp(NB_MUL(-kk, NB_SUB(SINR, x0)))));
            int det = nbc_bernoulli(c, p);
            if (det) e->st_ce = (uint8_t)((e->st_ce & 0xF0u) | NB_UE_RAR);
            detected += det; signals += det;
            continue;
        }
        /* several UEs on this preamble: one draw each in list order, then one detection draw for the strongest */
        double sum = 0.0, max_rx = -1.0;
        NB_NOUNROLL
        for (int q = r; q < n_ues; ++q) {
            const NbRaEntry *e = &L[c.scratch[q]];
            if ((int)e->wake_t != pre) continue;
            double LogF = nbc_normal(c, 0.0, NB_SIGMA_F);
            double rx_pw = nb_exp10(NB_DIV(NB_SUB(NB_SUB(NB_TX_PW, c.ue[e->slot].loss), LogF), 10.0));
            sum = NB_ADD(sum, rx_pw);
            if (rx_pw > max_rx) max_rx = rx_pw;
        }
        double SINR = NB_MUL(10.0, nb_log10(NB_DIV(max_rx, NB_SUB(NB_ADD(NB_NOISE_MW, sum), max_rx))));
        double p = NB_DIV(1.0, NB_ADD(1.0, nb_exp(NB_MUL(-kk, NB_SUB(SINR, x0)))));
        int det = nbc_bernoulli(c, p);
        NB_NOUNROLL
        for (int q = r; q < n_ues; ++q) {
            NbRaEntry *e = &L[c.scratch[q]];
            if ((int)e->wake_t == pre) e->st_ce = (uint8_t)((e->st_ce & 0xF0u) | (det ? NB_UE_CAPTURE : NB_UE_COLLIDED));
        }
        detected += det; signals += 1;
    }
    int collided = signals - detected;
    h->detected_pre[ce] = detected;
    h->collided_pre[ce] = collided;
    if (h->occ_n[ce] < NBIOT_MAX_OCC) { c.occ->det[ce][h->occ_n[ce]] = (int16_t)detected; c.occ->col[ce][h->occ_n[ce]] = (int16_t)collided; }
    else NB_FAULT(c, NBIOT_FAULT_LIST_TRUNC);
    h->occ_n[ce] += 1; h->det_sum[ce] += detected; h->col_sum[ce] += collided;
}

/* Population.msg3_grant (population.py:201-256) in full: the first UE of the list in RAR or CAPTURE state; when it was captured,
 * every later RAR / CAPTURE UE of the list on the same preamble answers the grant too (one Msg3 event, one MAC timer for all) */
NB_FN void nbc_msg3_grant_capture(NbCtx &c, int ce, int t, int t_msg3_end, int I_tbs, int N_rep)
{
    NbHdr *h = nbc_hdr(c);
    NbRaEntry *L = c.ra[ce];
    const int n = h->ra_n[ce];
    int first = -1, preamble = -1, members = 0, alone = 0;
    const uint32_t seq_msg3 = h->seq, seq_timer = h->seq + 1u;
    const int t_exp = t + h->MAC_timer > t_msg3_end + 1 ? t + h->MAC_timer : t_msg3_end + 1;
    NB_NOUNROLL
    for (int i = 0; i < n && !alone; ++i) {
        NbRaEntry e = L[i];
        if (e.st_ce == NB_RA_TOMB) continue;
        const uint32_t st = NB_STATE(e);
        if (st != NB_UE_RAR && st != NB_UE_CAPTURE) continue;
        if (first < 0) { first = i; preamble = (int)e.wake_t; alone = st == NB_UE_RAR; }
        else if ((int)e.wake_t != preamble) continue;
        L[i].st_ce = NB_RA_TOMB;
        NbUe *u = &c.ue[e.slot];
        u->state = (uint8_t)(st == NB_UE_RAR && members == 0 ? NB_UE_CONREQ : st);    /* only an uncollided first UE becomes CONREQ (:232-234) */
        u->I_tbs = (uint8_t)I_tbs; u->N_rep = (uint8_t)N_rep; u->rar_w_id = e.rar_w_id;
        u->attempts = e.attempts; u->CE_level = (uint8_t)NB_CE(e); u->where = NB_IN_CONTENTION;
        if (members == 0) nbc_pool_push(c, NB_KEY(t_msg3_end, seq_msg3), NB_EV_MSG3, ce, e.slot);
        if (h->cont_n >= c.ue_cap) { NB_FAULT(c, NBIOT_FAULT_UE_SLOTS); return; }
        NbContEntry ce_; ce_.key = NB_KEY(t_exp, seq_timer); ce_.ue_id = u->id; ce_.slot = e.slot; ce_.ord = (uint16_t)members;
        c.cont[h->cont_n] = ce_;
        if (h->mac_valid && ce_.key < h->next_mac) { h->next_mac = ce_.key; h->next_mac_idx = h->cont_n; }
        h->cont_n += 1;
        members += 1;
    }
    if (members) h->seq += 2;
}

/* the contention list of the Msg3 event whose first UE sits in `slot`, in grant order: the entries of the contention set that share its
 * MAC-timer key. Returns their number (written to c.scratch as UE slots). */
NB_FN int nbc_msg3_contenders(NbCtx &c, int slot)
{
    NbHdr *h = nbc_hdr(c);
    uint64_t key = NB_KEY_NONE;
    NB_NOUNROLL
    for (int i = 0; i < h->cont_n; ++i) if (c.cont[i].slot == (uint16_t)slot) { key = c.cont[i].key; break; }
    int n = 0;
    NB_NOUNROLL
    for (int i = 0; i < h->cont_n; ++i) if (c.cont[i].key == key) n += 1;
    NB_NOUNROLL
    for (int i = 0; i < h->cont_n; ++i) if (c.cont[i].key == key) c.scratch[c.cont[i].ord] = c.cont[i].slot;
    return n;
}

/* Population.MAC_timeout (population.py:151-163) for a timer that covers several UEs: those still in contention time out in the
 * order of the grant's list */
NB_FN void nbc_MAC_timeout_group(NbCtx &c, uint64_t key)
{
    NbHdr *h = nbc_hdr(c);
    for (;;) {
        int best = -1;
        NB_NOUNROLL
        for (int i = 0; i < h->cont_n; ++i) if (c.cont[i].key == key && (best < 0 || c.cont[i].ord < c.cont[best].ord)) best = i;
        if (best < 0) break;
        NbContEntry e = c.cont[best];
        h->cont_n -= 1;
        if (best != h->cont_n) c.cont[best] = c.cont[h->cont_n];
        NbUe *u = &c.ue[e.slot];
        u->state = NB_UE_RAO; u->where = NB_IN_RA;
        nbc_ra_append(c, u->CE_level, e.slot, NB_UE_RAO, u->CE_level, u->attempts, u->rar_w_id);
        if (NB_FATAL(c)) break;
    }
    h->mac_valid = 0;
}
#endif /* NB_NO_CAPTURE */

/* ------------------------------------------------------------------ NPRACH pass */
/* AccessProcedure.NPRACH_start (access_procedure.py:37-91), Population.NPRACH_start (population.py:168-188),
 * Channel.preamble_detection (channel.py:156-176) */
NB_FN1 void nbc_NPRACH_start(NbCtx &c, int ce, NbOccasion occ)
{
    NbHdr *h = nbc_hdr(c);
    int t = occ.t_start, N_sc = occ.N_sc, N_sf = occ.N_sf, N_pream = 4 * occ.N_sc;
    int rep = 0; while (rep < 7 && nb_nprach_nsf_list[rep] != N_sf) ++rep;      /* N_sf_to_N_rep (carrier.py:25-26) */
    int counter = h->RAR_counter[ce];
    h->NPRACH_resources += (int64_t)N_sc * N_sf;
    NB_TRACE(c, NBIOT_TR_RES_NPRACH, ce, t, N_sc * N_sf, 0);
    if (h->gen_t < t) { nbc_new_arrivals(c, t); if (NB_FATAL(c)) return; }
    NbRaEntry *L = c.ra[ce];
    int n = h->ra_n[ce];
    uint64_t cur_key = h->cur_key;
    int out_n = 0;
    /* pass 1: compact the list, pick the UEs waiting for an opportunity, in list order */
    int n_ues = nbw_run([&]() -> int {
        int o = 0, ns = 0;
        NB_NOUNROLL
        for (int base = 0; base < n; base += NB_LANES) {
            int i = base + nbw_lane();
            NbRaEntry e; e.st_ce = NB_RA_TOMB; e.wake_t = e.wake_seq = 0; e.rar_w_id = 0; e.slot = 0; e.attempts = 0;
            if (i < n) e = L[i];
            int keep = i < n && e.st_ce != NB_RA_TOMB;
            uint32_t st = NB_STATE(e);
            int sel = keep && (st == NB_UE_RAO || (st == NB_UE_BACKOFF && NB_KEY(e.wake_t, e.wake_seq) < cur_key));
            if (sel) { e.st_ce = (uint8_t)((e.st_ce & 0xF0u) | NB_UE_RA); e.rar_w_id = counter; }
            uint32_t mk = nbw_ballot(keep), ms = nbw_ballot(sel);
            int pos = o + nbw_popc(mk & nbw_lt_mask());
            nbw_sync();
            if (keep) L[pos] = e;
            if (sel) c.scratch[ns + nbw_popc(ms & nbw_lt_mask())] = (uint16_t)pos;
            o += nbw_popc(mk); ns += nbw_popc(ms);
            nbw_sync();
        }
        out_n = o;
        return ns;
    });
    h->ra_n[ce] = out_n;
    h->contenders[ce] = n_ues;
    if (n_ues <= 0) return;

    uint64_t seed = h->seed, ctr = h->draws;
    int Na_sc = nbc_sclog_check(c, t, ce);
    int na_ues = 0;
    if (Na_sc) {                                  /* binomial(n_ues, p_anchor): n_ues Bernoulli counters */
        double p = h->probability_anchor;
        na_ues = nbw_run([&]() -> int {
            int s = 0;
            NB_NOUNROLL
            for (int r = nbw_lane(); r < n_ues; r += NB_LANES) s += nb_rng_bernoulli(seed, ctr + (uint64_t)r, NBIOT_STREAM_ENV, p);
            return nbw_add_i32(s);
        });
        ctr += (uint64_t)n_ues;
        h->NPRACH_resources += (int64_t)Na_sc * N_sf;
        NB_TRACE(c, NBIOT_TR_RES_NPRACH, ce, t, Na_sc * N_sf, 0);
    }
    NB_NOUNROLL
    for (int w = 0; w < 8; ++w) { h->pre_once[w] = 0; h->pre_twice[w] = 0; }
    /* pass 2: one preamble per contender (counter ctr + rank), histogram into once/twice bit sets */
    nbw_run([&]() -> int {
        nbw_sync();
        NB_NOUNROLL
        for (int r = nbw_lane(); r < n_ues; r += NB_LANES) {
            int pre = r < na_ues ? N_pream + (int)nb_rng_bounded(seed, ctr + (uint64_t)r, NBIOT_STREAM_ENV, (uint64_t)(4 * Na_sc))
                                 : (int)nb_rng_bounded(seed, ctr + (uint64_t)r, NBIOT_STREAM_ENV, (uint64_t)N_pream);
            L[c.scratch[r]].wake_t = (uint32_t)pre;
            uint32_t bit = 1u << (pre & 31);
            uint32_t old = nbw_atomic_or(&h->pre_once[(pre >> 5) & 7], bit);
            if (old & bit) nbw_atomic_or(&h->pre_twice[(pre >> 5) & 7], bit);
        }
        nbw_sync();
        return 0;
    });
    ctr += (uint64_t)n_ues;
#ifndef NB_NO_CAPTURE
    if (nbc_cold(c)->capture_effect) {                         /* every preamble group in first-occurrence order, the shared ones included */
        h->draws = ctr;
        nbc_nprach_detect_capture(c, ce, n_ues, rep);
        return;
    }
#endif
    /* pass 3: singleton preambles are detected with a sigmoid(SINR) Bernoulli draw, in list order;
     * a preamble chosen twice or more is a certain failure and consumes no draw */
    double x0 = nb_nprach_x0[rep], kk = nb_nprach_k[rep];
    int n_single = 0;
    int detected = nbw_run([&]() -> int {
        int det_total = 0, q0 = 0;
        NB_NOUNROLL
        for (int base = 0; base < n_ues; base += NB_LANES) {
            int r = base + nbw_lane();
            int idx = 0, pre = 0;
            if (r < n_ues) { idx = c.scratch[r]; pre = (int)L[idx].wake_t; }
            int single = r < n_ues && !((h->pre_twice[(pre >> 5) & 7] >> (pre & 31)) & 1u);
            uint32_t ms = nbw_ballot(single);
            int det = 0;
            if (single) {
                int q = q0 + nbw_popc(ms & nbw_lt_mask());
                double loss = c.ue[L[idx].slot].loss;
                double LogF = NB_ADD(0.0, NB_MUL(NB_SIGMA_F, nb_rng_normal(seed, ctr + 2u * (uint64_t)q, NBIOT_STREAM_ENV)));
                double rx_pw = NB_SUB(NB_SUB(NB_TX_PW, loss), LogF);
                double SINR = NB_SUB(NB_SUB(rx_pw, NB_NOISE_DBM), NB_NOISE_FIG);
                double p = NB_DIV(1.0, NB_ADD(1.0, nb_exp(NB_MUL(-kk, NB_SUB(SINR, x0)))));
                det = nb_rng_u01(seed, ctr + 2u * (uint64_t)q + 1u, NBIOT_STREAM_ENV) < p;
                if (det) L[idx].st_ce = (uint8_t)((L[idx].st_ce & 0xF0u) | NB_UE_RAR);
            }
            det_total += nbw_popc(nbw_ballot(det));
            q0 += nbw_popc(ms);
        }
        nbw_sync();
        n_single = q0;
        return det_total;
    });
    h->draws = ctr + 2u * (uint64_t)n_single;
    int collided = 0;
    NB_NOUNROLL
    for (int w = 0; w < 8; ++w) collided += nbw_popc(h->pre_twice[w]);
    h->detected_pre[ce] = detected;
    h->collided_pre[ce] = collided;
    if (h->occ_n[ce] < NBIOT_MAX_OCC) { c.occ->det[ce][h->occ_n[ce]] = (int16_t)detected; c.occ->col[ce][h->occ_n[ce]] = (int16_t)collided; }
    else NB_FAULT(c, NBIOT_FAULT_LIST_TRUNC);
    h->occ_n[ce] += 1; h->det_sum[ce] += detected; h->col_sum[ce] += collided;
}

/* AccessProcedure.NPRACH_end (access_procedure.py:93-116) + NodeB.start_RAR_window (node_b.py:524-554) */
NB_FN1S void nbc_NPRACH_end(NbCtx &c, int ce, int t)
{
    NbHdr *h = nbc_hdr(c);
    int counter = h->RAR_counter[ce];
    int n_ues = h->contenders[ce];
    int detections = h->detected_pre[ce];
    NB_TRACE(c, NBIOT_TR_NPRACH, ce, t, n_ues, detections);          /* perf_monitor.nprach_sample (access_procedure.py:105) */
    if (n_ues > 0) {
        if (h->rar_n[ce] >= NB_RAR_WIN) { NB_FAULT(c, NBIOT_FAULT_EVENT_SLOTS); return; }
        h->RAR_UEs[ce][h->rar_n[ce]] = detections; h->rar_tot[ce] += detections;
        h->RAR_w_ids[ce][h->rar_n[ce]] = counter;
        h->rar_n[ce] += 1;
        h->state = NB_ST_RAR_WINDOW;
        int RAR_in = h->RAR_in[ce], RAR_det = h->RAR_ids[ce];
        h->m3_cnt[ce] += 1; h->m3_sent_sum[ce] += h->RAR_sent[ce]; h->m3_recv_sum[ce] += RAR_det;
        h->m3_eff_sum[ce] = NB_ADD(h->m3_eff_sum[ce], NB_DIV((double)RAR_det, (double)(RAR_in > 1 ? RAR_in : 1)));
        NB_TRACE(c, NBIOT_TR_RARWIN, ce, t, RAR_in | (h->RAR_attmp[ce] << 16), h->RAR_sent[ce] | (RAR_det << 16));   /* rar_window_sample (node_b.py:543) */
        h->RAR_in[ce] = detections; h->RAR_ids[ce] = 0; h->RAR_attmp[ce] = 0; h->RAR_sent[ce] = 0;
        nbc_pool_push(c, NB_KEY(t + h->RAR_window_size[ce], h->seq), NB_EV_RAR_END, ce, counter);
        h->seq += 1;
        h->detected_pre[ce] = 0; h->collided_pre[ce] = 0; h->contenders[ce] = 0;
    }
    h->RAR_counter[ce] += 1;
}

/* ------------------------------------------------------------------ contention set / MAC timers */
NB_FN1 void nbc_mac_refresh(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    int n = h->cont_n;
    uint64_t best = NB_KEY_NONE; int bidx = -1;
    nbw_run([&]() -> int {
        uint64_t k = NB_KEY_NONE; int idx = -1;
        NB_NOUNROLL
        for (int i = nbw_lane(); i < n; i += NB_LANES) { uint64_t ki = c.cont[i].key; if (ki < k) { k = ki; idx = i; } }
        uint64_t m = nbw_min_u64(k);
        uint32_t who = nbw_ballot(k == m && idx >= 0);
        int src = who ? nbw_ffs(who) - 1 : 0;
        best = m; bidx = (int)nbw_shfl((uint32_t)idx, src);
        return 0;
    });
    h->next_mac = best; h->next_mac_idx = n > 0 ? bidx : -1; h->mac_valid = 1;
}

NB_DEV void nbc_cont_remove(NbCtx &c, int idx)
{
    NbHdr *h = nbc_hdr(c);
    h->cont_n -= 1;
    if (idx != h->cont_n) c.cont[idx] = c.cont[h->cont_n];
    h->mac_valid = 0;
}

/* Population.msg3_grant (population.py:201-256): the first UE of the list that is in RAR state */
NB_FN1S void nbc_msg3_grant(NbCtx &c, int ce, int t, int t_msg3_end, int I_tbs, int N_rep)
{
    NbHdr *h = nbc_hdr(c);
#ifndef NB_NO_CAPTURE
    if (nbc_cold(c)->capture_effect) { nbc_msg3_grant_capture(c, ce, t, t_msg3_end, I_tbs, N_rep); return; }
#endif
    NbRaEntry *L = c.ra[ce];
    int n = h->ra_n[ce];
    int idx = nbw_run([&]() -> int {
        NB_NOUNROLL
        for (int base = 0; base < n; base += NB_LANES) {
            int i = base + nbw_lane();
            int hit = i < n && L[i].st_ce != NB_RA_TOMB && NB_STATE(L[i]) == NB_UE_RAR;
            uint32_t m = nbw_ballot(hit);
            if (m) return base + nbw_ffs(m) - 1;
        }
        return -1;
    });
    if (idx < 0) return;
    NbRaEntry e = L[idx];
    L[idx].st_ce = NB_RA_TOMB;
    NbUe *u = &c.ue[e.slot];
    u->state = NB_UE_CONREQ; u->I_tbs = (uint8_t)I_tbs; u->N_rep = (uint8_t)N_rep; u->rar_w_id = e.rar_w_id;
    u->attempts = e.attempts; u->CE_level = (uint8_t)NB_CE(e); u->where = NB_IN_CONTENTION;
    nbc_pool_push(c, NB_KEY(t_msg3_end, h->seq), NB_EV_MSG3, ce, e.slot);
    h->seq += 1;
    int t_exp = t + h->MAC_timer > t_msg3_end + 1 ? t + h->MAC_timer : t_msg3_end + 1;
    NbContEntry ce_; ce_.key = NB_KEY(t_exp, h->seq); ce_.ue_id = u->id; ce_.slot = e.slot; ce_.ord = 0;
    h->seq += 1;
    c.cont[h->cont_n] = ce_;
    if (h->mac_valid && ce_.key < h->next_mac) { h->next_mac = ce_.key; h->next_mac_idx = h->cont_n; }
    h->cont_n += 1;
}

/* Population.MAC_timeout (population.py:151-163) */
NB_FN1S void nbc_MAC_timeout(NbCtx &c, int idx)
{
#ifndef NB_NO_CAPTURE
    if (nbc_cold(c)->capture_effect) { nbc_MAC_timeout_group(c, c.cont[idx].key); return; }
#endif
    NbContEntry e = c.cont[idx];
    nbc_cont_remove(c, idx);
    NbUe *u = &c.ue[e.slot];
    u->state = NB_UE_RAO; u->where = NB_IN_RA;
    nbc_ra_append(c, u->CE_level, e.slot, NB_UE_RAO, u->CE_level, u->attempts, u->rar_w_id);
}

/* Population.RAR_window_end (population.py:258-306): unserved UEs of the window back off */
NB_FN1 void nbc_population_RAR_window_end(NbCtx &c, int t, int ce, int backoff, int rar_w_id)
{
    NbHdr *h = nbc_hdr(c);
    NbRaEntry *L = c.ra[ce];
    int n = h->ra_n[ce];
    uint64_t seed = h->seed, ctr = h->draws;
    uint32_t seq0 = h->seq;
    int trans_max = h->preamble_trans_max;
    int n_nz = 0;
    int n_sel = nbw_run([&]() -> int {
        int ns = 0, nz = 0;
        NB_NOUNROLL
        for (int base = 0; base < n; base += NB_LANES) {
            int i = base + nbw_lane();
            NbRaEntry e; e.st_ce = NB_RA_TOMB; e.wake_t = e.wake_seq = 0; e.rar_w_id = 0; e.slot = 0; e.attempts = 0;
            if (i < n) e = L[i];
            uint32_t st = NB_STATE(e);
            int sel = i < n && e.st_ce != NB_RA_TOMB && st != NB_UE_RAO && st != NB_UE_BACKOFF && e.rar_w_id == rar_w_id;
            uint32_t ms = nbw_ballot(sel);
            int nonzero = 0;
            if (sel) {
                int r = ns + nbw_popc(ms & nbw_lt_mask());
                int attempts = e.attempts < 255 ? e.attempts + 1 : 255;
                int uce = (int)NB_CE(e);
                if (attempts > trans_max && uce < 2) { attempts = 0; uce += 1; }
                int bt = backoff == 0 ? 0 : (int)nb_rng_bounded(seed, ctr + (uint64_t)r, NBIOT_STREAM_ENV, (uint64_t)backoff);
                nonzero = bt != 0;
                e.attempts = (uint8_t)attempts;
                e.st_ce = (uint8_t)((nonzero ? NB_UE_BACKOFF : NB_UE_RAO) | (uce << 4));
                e.wake_t = (uint32_t)(t + bt);
            }
            uint32_t mz = nbw_ballot(nonzero);
            if (sel) {
                e.wake_seq = nonzero ? seq0 + (uint32_t)(nz + nbw_popc(mz & nbw_lt_mask())) : 0u;
                L[i] = e;
            }
            ns += nbw_popc(ms); nz += nbw_popc(mz);
        }
        nbw_sync();
        n_nz = nz;
        return ns;
    });
    if (backoff != 0) h->draws = ctr + (uint64_t)n_sel;
    h->seq = seq0 + (uint32_t)n_nz;
    h->backoffs[ce] += n_sel;
}

/* ------------------------------------------------------------------ connected queue */
NB_FNM4 void nbc_conn_insert(NbCtx &c, int slot)
{
    NbHdr *h = nbc_hdr(c);
    NbUe *u = &c.ue[slot];
    uint64_t k1;
    if (c.conf.sort_by_loss) k1 = nb_d2u(u->loss);           /* loss > 0: the bit pattern orders like the value */
    else k1 = (uint64_t)(uint32_t)u->t_connection;
    int n = h->conn_n;
    NbConnEntry *Q = c.conn;
    if (n >= c.ue_cap) { NB_FAULT(c, NBIOT_FAULT_UE_SLOTS); return; }
    NbConnEntry e; e.k1 = k1; e.order = h->conn_order; e.slot = (uint16_t)slot; e.ce = u->CE_level; e.pad = 0;
    h->conn_order += 1;
    nbw_run([&]() -> int {
        /* stable sort position: after every entry whose key is <= k1 (the new stamp is the largest) */
        int pos = n;
        if (n > 0 && Q[n - 1].k1 > k1) {                       /* not the usual append: count the keys <= k1 */
            int cnt = 0;
            NB_NOUNROLL
            for (int i = nbw_lane(); i < n; i += NB_LANES) cnt += Q[i].k1 <= k1;
            pos = nbw_add_i32(cnt);
        }
        NB_NOUNROLL
        for (int hi = n; hi > pos; hi -= NB_LANES) {
            int i = hi - 1 - nbw_lane();
            NbConnEntry m; m.k1 = 0; m.order = 0; m.slot = 0; m.ce = 0; m.pad = 0;
            if (i >= pos) m = Q[i];
            nbw_sync();
            if (i >= pos) Q[i + 1] = m;
            nbw_sync();
        }
        if (nbw_lane() == 0) Q[pos] = e;
        nbw_sync();
        return 0;
    });
    h->conn_n = n + 1;
    u->where = NB_IN_CONNECTED;
}

NB_FN1S void nbc_conn_remove(NbCtx &c, int slot)
{
    NbHdr *h = nbc_hdr(c);
    int n = h->conn_n;
    NbConnEntry *Q = c.conn;
    int found = nbw_run([&]() -> int {
        int idx = -1;
        NB_NOUNROLL
        for (int base = 0; base < n; base += NB_LANES) {
            int i = base + nbw_lane();
            uint32_t m = nbw_ballot(i < n && Q[i].slot == (uint16_t)slot);
            if (m) { idx = base + nbw_ffs(m) - 1; break; }
        }
        if (idx < 0) return -1;
#if NB_LANES == 1 && !defined(NB_CONN_SHIFT_1)
        /* one lane owns the cell and shifts a sorted queue of up to hundreds of 16-byte entries: two loads in flight per step. +19 % on cells
         * with 150-300 live UEs, -0.4 % on lightly loaded ones; four in flight, or the same in nbc_conn_insert, cost the lightly loaded cells
         * 1.3-3.5 % in code size for no more (profiles/r02_tpc_variants.txt section 18) */
        int lo = idx;
        NB_NOUNROLL
        for (; lo + 2 <= n - 1; lo += 2) {
            NbConnEntry a = Q[lo + 1], b = Q[lo + 2];
            Q[lo] = a; Q[lo + 1] = b;
        }
        NB_NOUNROLL
        for (; lo < n - 1; ++lo) Q[lo] = Q[lo + 1];
#else
        NB_NOUNROLL
        for (int base = idx; base < n - 1; base += NB_LANES) {
            int i = base + nbw_lane();
            NbConnEntry m; m.k1 = 0; m.order = 0; m.slot = 0; m.ce = 0; m.pad = 0;
            if (i < n - 1) m = Q[i + 1];
            nbw_sync();
            if (i < n - 1) Q[i] = m;
            nbw_sync();
        }
#endif
        return idx;
    });
    if (found >= 0) h->conn_n = n - 1;
}

/* the first N_users UEs of the sorted queue whose CE level the remaining NPDCCH subframes allow
 * (NodeB.get_node_info, node_b.py:301-312) */
NB_FN1S void nbc_conn_select(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    int max_CE = nb_ce_for_sf_left[h->NPDCCH_sf_left];
    int n = h->conn_n;
    NbConnEntry *Q = c.conn;
    int sel[4] = { -1, -1, -1, -1 };
    int cnt = nbw_run([&]() -> int {
        int ns = 0;
        NB_NOUNROLL
        for (int base = 0; base < n && ns < NB_N_USERS; base += NB_LANES) {
            int i = base + nbw_lane();
            int ok = i < n && Q[i].ce <= max_CE;
            uint32_t slot = i < n ? Q[i].slot : 0u;
            uint32_t m = nbw_ballot(ok);
            NB_NOUNROLL
            while (m && ns < NB_N_USERS) {
                int src = nbw_ffs(m) - 1;
                sel[ns++] = (int)nbw_shfl(slot, src);
                m &= m - 1;
            }
        }
        return ns;
    });
    h->n_sel = cnt;
    NB_NOUNROLL
    for (int i = 0; i < 4; ++i) h->sel_slot[i] = sel[i];
}

/* ------------------------------------------------------------------ receptions */
NB_FNM2 double nbc_get_bler(NbCtx &c, double snr, int itbs, int nr)   /* channel.py:76-79 */
{
    if (!(snr > -30.0)) snr = -30.0;
    if (!(snr < 20.0)) snr = 20.0;
    int idx = (int)NB_RINT(NB_MUL(snr, 10.0)) + 300;
    int row = nb_itbs_to_row[itbs], rep = nbc_rep_index(nr);
    return NB_DIV((double)NB_LDG(&c.lut2[(row * 8 + rep) * NB_LUT_SNR_PTS + idx]), NB_LUT_SCALE);
}

NB_FN1S void nbc_distribution_update(NbCtx &c, double sample)       /* node_b.py:207-225 */
{
    NbHdr *h = nbc_hdr(c);
    if (h->k_dist > 1000) return;
    h->k_dist += 1;
    int idx = 14;
    NB_NOUNROLL
    for (int i = 0; i < 14; ++i) if (NB_ADD(sample, 35.0) < (double)nb_threshold_list[i]) { idx = i; break; }
    double k = (double)h->k_dist;
    nbw_run([&]() -> int {                                     /* one interval per lane */
        nbw_sync();
        NB_NOUNROLL
        for (int i = nbw_lane(); i < 15; i += NB_LANES) {
            double prob = h->distribution[i];
            h->distribution[i] = NB_ADD(prob, NB_DIV(NB_SUB(i == idx ? 1.0 : 0.0, prob), k));
        }
        nbw_sync();
        return 0;
    });
}

/* RxProcedure.msg3_event (rx_procedure.py:29-51), Channel.msg3_detection / sinr_estimation
 * (channel.py:178-198, :221-234), NodeB.msg3_outcome (node_b.py:770-799) */
NB_FN1 void nbc_msg3_event(NbCtx &c, int slot, int t)
{
    NbHdr *h = nbc_hdr(c);
    NbUe *u = &c.ue[slot];
    const int I_tbs = u->I_tbs, N_rep = u->N_rep;              /* of the first UE of the contention list (channel.py:226-227) */
    int contenders = 1;
    double SINR;
#ifndef NB_NO_CAPTURE
    if (nbc_cold(c)->capture_effect && (contenders = nbc_msg3_contenders(c, slot)) > 1) {
        int mi;                                                /* several UEs answered the grant: the strongest one is the candidate (channel.py:221-234) */
        SINR = nbc_sinr_estimation(c, c.scratch, contenders, &mi);
        slot = c.scratch[mi]; u = &c.ue[slot];
    } else
#endif
    {
        double LogF = nbc_normal(c, 0.0, NB_SIGMA_F);
        double rx_pw = NB_SUB(NB_SUB(NB_TX_PW, u->loss), LogF);
        rx_pw = nb_exp10(NB_DIV(rx_pw, 10.0));
        double ni = NB_SUB(NB_ADD(NB_NOISE_MW, NB_ADD(0.0, rx_pw)), rx_pw);
        SINR = NB_MUL(10.0, nb_log10(NB_DIV(rx_pw, ni)));
    }
    double bler = nbc_get_bler(c, SINR, I_tbs, N_rep);
    double p = nb_pow01(NB_SUB(1.0, bler), (double)NB_MSG3_TBS / 256.0);
    int connected = nbc_bernoulli(c, p);
    int uce = u->CE_level, i_ = 0, window_active = 0;
    NB_NOUNROLL
    for (int i = 0; i < h->rar_n[uce]; ++i) if (h->RAR_w_ids[uce][i] == u->rar_w_id) { window_active = 1; i_ = i; break; }
    if (connected) {
        u->state = NB_UE_CONNECTED; u->t_connection = t;
        int n = h->cont_n;
        int idx = nbw_run([&]() -> int {                       /* population.msg4: leave the contention set */
            NB_NOUNROLL
            for (int base = 0; base < n; base += NB_LANES) {
                int i = base + nbw_lane();
                uint32_t m = nbw_ballot(i < n && c.cont[i].slot == (uint16_t)slot);
                if (m) return base + nbw_ffs(m) - 1;
            }
            return -1;
        });
        if (idx >= 0) nbc_cont_remove(c, idx);
        nbc_conn_insert(c, slot);
        u->timestamp = t;                                      /* perf.register_ue (perf_monitor.py:224-229) */
        if (h->n_acc < NBIOT_STEP_LIST) { h->acc_id[h->n_acc] = u->id; h->acc_delay[h->n_acc] = t - u->t_arrival; }
        else { nbc_step_spill(c, 4, h->n_acc, u->id); nbc_step_spill(c, 5, h->n_acc, t - u->t_arrival); }
        h->n_acc += 1;
        h->RAR_ids[uce] += 1; h->RAR_detected[uce] += 1;
        double sample = NB_MUL(-1.0, u->loss);
        if (h->n_incoming < c.hist_cap) c.incoming[h->n_incoming] = sample; else NB_FAULT(c, NBIOT_FAULT_LIST_TRUNC);
        h->n_incoming += 1;
        nbc_distribution_update(c, sample);
        h->ue_connections += 1;
    } else if (window_active) {
        u->state = NB_UE_RAR;
        h->RAR_UEs[uce][i_] += 1; h->rar_tot[uce] += 1;
    }
    NB_TRACE(c, NBIOT_TR_MSG3, uce, t, connected, connected ? 0 : (contenders > 1 ? 2 : 1));   /* perf_monitor.rar_sample (rx_procedure.py:41-51): a failure is an error with one contender, a collision with several */
    nbc_update_state(c);
}

/* statistics histories (perf_monitor.py:161-171): rings of stat_cap entries per cell, out of line (off unless conf['statistics']) */
NB_FN void nbc_stat_append(NbCtx &c, int delay, int service_time, int access, int bits)
{
    const NbColdConf *cold = nbc_cold(c);
    int32_t *st = cold->stat_base + (size_t)(c.occ - cold->occ_base) * 4u * (size_t)cold->stat_cap;
    int j = c.occ->stat_n % cold->stat_cap;
    st[j] = delay; st[cold->stat_cap + j] = service_time; st[2 * cold->stat_cap + j] = access; st[3 * cold->stat_cap + j] = bits;
    c.occ->stat_n += 1;
}

/* RxProcedure.NPUSCH_end (rx_procedure.py:54-61), Channel.NPUSCH_detection (channel.py:236-256),
 * NodeB.NPUSCH_end (node_b.py:801-830), PerfMonitor.account_rx/account_error/unregister_ue */
NB_FN1 void nbc_NPUSCH_end(NbCtx &c, int slot, int t)
{
    NbHdr *h = nbc_hdr(c);
    NbUe *u = &c.ue[slot];
    double LogF = nbc_normal(c, 0.0, NB_SIGMA_F);
    double rx_pw = NB_SUB(NB_SUB(NB_TX_PW, u->loss), LogF);
    double SINR = NB_SUB(NB_SUB(rx_pw, NB_NOISE_DBM), NB_NOISE_FIG);
    if (!u->new_data) SINR = NB_MUL(10.0, nb_log10(NB_ADD(nb_exp10(NB_DIV(SINR, 10.0)), nb_exp10(NB_DIV(u->sinr, 10.0)))));
    double bler = nbc_get_bler(c, SINR, u->I_tbs, u->N_rep);
    double p = u->tbs != 256 ? nb_pow01(NB_SUB(1.0, bler), NB_DIV((double)u->tbs, 256.0)) : NB_SUB(1.0, bler);
    int received = nbc_bernoulli(c, p);
    h->n_scheduled -= 1;
    u->sinr = SINR;
    if (received) {
        u->acked_data += u->bits; u->retx_buffer = 0; u->bits = 0; u->new_data = 1;
        int d = t - u->timestamp;
        if (h->n_rx < NBIOT_STEP_LIST) { h->rx_id[h->n_rx] = u->id; h->rx_delay[h->n_rx] = d; }
        else { nbc_step_spill(c, 0, h->n_rx, u->id); nbc_step_spill(c, 1, h->n_rx, d); }
        h->n_rx += 1;
        h->rx_sum = NB_ADD(h->rx_sum, (double)d);
        if (c.conf.reward_criteria == NB_RW_INVD_USERS) h->rx_inv_sum = NB_ADD(h->rx_inv_sum, NB_DIV(1.0, NB_DIV((double)d, 1000.0)));
        if (c.conf.reward_criteria == NB_RW_LOG_INVD_USERS) h->rx_log_sum = NB_ADD(h->rx_log_sum, nb_log(NB_DIV((double)d, 1000.0)));
        u->timestamp = t; h->attempts_count += 1;
        if (!u->buffer) {
            u->state = NB_UE_DONE;
            int delay = t - u->t_connection, service_time = t - u->t_arrival;
            h->ue_departures[u->CE_level] += 1; h->thr_count += 1; h->total_bits += u->acked_data;
            int deps = h->ue_departures[0] + h->ue_departures[1] + h->ue_departures[2];
            h->av_delay = NB_ADD(h->av_delay, NB_DIV(NB_SUB((double)delay, h->av_delay), (double)deps));
            if (nbc_cold(c)->stat_cap > 0) nbc_stat_append(c, delay, service_time, u->t_connection - u->t_arrival, u->acked_data);
            NB_TRACE(c, NBIOT_TR_DEPART, u->CE_level, t, 0, 0);        /* npusch_delivered (perf_monitor.py:164-168) */
            if (c.conf.random_traffic) { if (nbc_bernoulli(c, h->gen_p) > 0) h->M_beta += 1; else h->M_uniform += 1; }   /* ue_generator.departure */
            else if (u->beta && h->M_beta < h->M_beta_max) h->M_beta += 1;
            else h->M_uniform += 1;
            if (h->n_departures < c.hist_cap) { c.delays[h->n_departures] = delay; c.service[h->n_departures] = service_time; }
            else NB_FAULT(c, NBIOT_FAULT_LIST_TRUNC);
            h->n_departures += 1; h->sum_delays += delay;
            h->total_departures += 1;
            u->where = NB_IN_FREE;
            c.free_stack[h->free_top] = (uint16_t)slot;
            h->free_top += 1;
        } else nbc_conn_insert(c, slot);
    } else {
        u->new_data = 0; u->retx_buffer = u->bits;
        if (h->n_err < NBIOT_STEP_LIST) h->err_id[h->n_err] = u->id; else nbc_step_spill(c, 2, h->n_err, u->id);
        h->n_err += 1; h->error_count += 1; h->attempts_count += 1;
        nbc_conn_insert(c, slot);
    }
    nbc_update_state(c);
}

/* ------------------------------------------------------------------ step handlers */
struct NbSched { int carrier, i, I_tbs, delay, N_sc, N_rep, N_ru; };

NB_DEV int nbc_scheduling_action(NbCtx &c, const uint8_t *a, int RAR_action, NbSched *p)   /* action_reader.py:25-66 */
{
    if (a[NB_A_CARRIER] >= NB_NC(c) || a[NB_A_NREP] > 7 || a[NB_A_DELAY] > 3 || a[NB_A_SC] > 3) { NB_FAULT(c, NBIOT_FAULT_ACTION); return 0; }
    p->carrier = a[NB_A_CARRIER]; p->i = a[NB_A_ID]; p->N_ru = a[NB_A_NRU];
    p->N_rep = nb_n_rep_list[a[NB_A_NREP] & 7];
    int I_mcs = a[NB_A_IMCS];
    if (RAR_action) {
        if (I_mcs > 2) { NB_FAULT(c, NBIOT_FAULT_ACTION); return 0; }
        p->I_tbs = I_mcs; p->N_ru = nb_rar_imcs_to_nru[I_mcs];
    } else {
        if (I_mcs > 6) { NB_FAULT(c, NBIOT_FAULT_ACTION); return 0; }
        p->I_tbs = nb_imcs_to_itbs[I_mcs];
    }
    p->delay = nb_sf_delay_list[a[NB_A_DELAY] & 3];
    p->N_sc = nb_n_sc_list[a[NB_A_SC] & 3];
    return 1;
}

NB_DEV int nbc_rar_list_cmp(NbCtx &c, int a, int b)              /* Python list comparison of RAR_UEs[a], RAR_UEs[b] */
{
    NbHdr *h = nbc_hdr(c);
    int na = h->rar_n[a], nb = h->rar_n[b], n = na < nb ? na : nb;
    NB_NOUNROLL
    for (int i = 0; i < n; ++i) if (h->RAR_UEs[a][i] != h->RAR_UEs[b][i]) return h->RAR_UEs[a][i] < h->RAR_UEs[b][i] ? -1 : 1;
    return na == nb ? 0 : (na < nb ? -1 : 1);
}

NB_FN1 void nbc_step_RAR_window(NbCtx &c, const uint8_t *a)      /* node_b.py:453-521 */
{
    NbHdr *h = nbc_hdr(c);
    NbSched P;
    if (!nbc_scheduling_action(c, a, 1, &P)) return;
    int CE_level = P.i, I_tbs = P.I_tbs, N_ru = P.N_ru, N_rep = P.N_rep;
    if (CE_level > 2) { NB_FAULT(c, NBIOT_FAULT_ACTION); return; }
    h->action3[0] = CE_level; h->action3[1] = I_tbs; h->action3[2] = N_rep;
    int RAR_ues[3] = { nbc_rar_sum(c, 0), nbc_rar_sum(c, 1), nbc_rar_sum(c, 2) };
    h->valid_CE_control = 1;
    int valid_control = 1;
    int NPDCCH_sf = nb_npdcch_sf_per_ce[CE_level];
    if (!(RAR_ues[CE_level] > 0) || NPDCCH_sf > h->NPDCCH_sf_left || c.conf.ce_mcs_automatic) {
        int order[3] = { 0, 1, 2 };                            /* stable descending sort by list value */
        NB_NOUNROLL
        for (int x = 1; x < 3; ++x)
            NB_NOUNROLL
            for (int y = x; y > 0; --y) {
                if (nbc_rar_list_cmp(c, order[y - 1], order[y]) < 0) { int tmp = order[y - 1]; order[y - 1] = order[y]; order[y] = tmp; }
                else break;
            }
        valid_control = 0;
        NB_NOUNROLL
        for (int x = 0; x < 3; ++x) {
            CE_level = order[x];
            NPDCCH_sf = nb_npdcch_sf_per_ce[CE_level];
            valid_control = (RAR_ues[CE_level] > 0) && (h->NPDCCH_sf_left >= NPDCCH_sf);
            if (valid_control) { h->valid_CE_control = 0; break; }
        }
    }
    if (!valid_control) { nbc_schedule_NPDCCH(c, h->NPDCCH_sf_left); return; }
    int N_sc = P.N_sc;
    if (c.conf.ce_mcs_automatic) { I_tbs = 2; N_ru = 1; N_rep = h->msg3_nrep[CE_level]; }
    int reference_time = h->time + NPDCCH_sf;
    int n_sf = 0;
    int t_ul_end = nbc_node_allocate(c, P.carrier, reference_time, P.delay, N_sc, N_ru, N_rep, &N_sc, &n_sf);
    if (NB_FATAL(c)) return;
    h->RAR_attmp[CE_level] += 1;
    if (t_ul_end) {
        nbc_msg3_grant(c, CE_level, reference_time, t_ul_end, I_tbs, N_rep);
        h->RAR_sent[CE_level] += 1;
        NB_NOUNROLL
        for (int i = 0; i < h->rar_n[CE_level]; ++i) if (h->RAR_UEs[CE_level][i] > 0) { h->RAR_UEs[CE_level][i] -= 1; h->rar_tot[CE_level] -= 1; break; }
        h->msg3_resources += (int64_t)N_sc * n_sf;
        NB_TRACE(c, NBIOT_TR_RES_MSG3, CE_level, h->time, N_sc * n_sf, 0);
    } else h->valid_CE_control = 0;
    nbc_schedule_NPDCCH(c, NPDCCH_sf);
    nbc_update_state(c);
}

NB_FN1S void nbc_step_RAR_end(NbCtx &c, const uint8_t *a)         /* node_b.py:558-581 */
{
    NbHdr *h = nbc_hdr(c);
    int ce = h->cur_ce;
    if (a[NB_A_BACKOFF] > 11) { NB_FAULT(c, NBIOT_FAULT_ACTION); return; }
    int backoff = nb_backoff_list[a[NB_A_BACKOFF]];
    int period_i = h->nprach_conf[7 + 3 * ce];
    if (nb_backoff_list[period_i + 1] > backoff) backoff = nb_backoff_list[period_i + 1];
    nbc_population_RAR_window_end(c, h->cur_t, ce, backoff, h->cur_wid);
    if (h->rar_n[ce] > 0) {
        h->RAR_failed[ce] = h->RAR_UEs[ce][0]; h->rar_tot[ce] -= h->RAR_UEs[ce][0];
        NB_NOUNROLL
        for (int i = 1; i < h->rar_n[ce]; ++i) { h->RAR_UEs[ce][i - 1] = h->RAR_UEs[ce][i]; h->RAR_w_ids[ce][i - 1] = h->RAR_w_ids[ce][i]; }
        h->rar_n[ce] -= 1;
    }
    nbc_update_state(c);
}

NB_DEV void nbc_clear_info(NbCtx &c)                             /* perf_monitor.py:96-103 */
{
    NbHdr *h = nbc_hdr(c);
    h->n_rx = h->n_err = h->n_unfit = h->n_acc = 0;
    h->rx_sum = 0.0; h->rx_inv_sum = 0.0; h->rx_log_sum = 0.0;
}

NB_FN1 void nbc_step_data(NbCtx &c, const uint8_t *a)            /* node_b.py:584-693 */
{
    NbHdr *h = nbc_hdr(c);
    nbc_clear_info(c);
    if (h->n_sel == 0) { nbc_schedule_NPDCCH(c, h->NPDCCH_sf_left); nbc_update_state(c); return; }
    NbSched P;
    if (!nbc_scheduling_action(c, a, 0, &P)) return;
    int ue_i = P.i; if (ue_i >= h->n_sel) ue_i = 0;
    int slot = h->sel_slot[ue_i];
    NbUe *u = &c.ue[slot];
    int CE_level = u->CE_level, new_data = u->new_data;
    int I_tbs, N_rep, N_ru, tbs;
    if (!new_data) { I_tbs = u->I_tbs; N_rep = u->N_rep; N_ru = u->N_ru; tbs = u->tbs; }
    else if (c.conf.mcs_automatic) {
        int li = (int)NB_RINT(NB_MUL(u->loss, 10.0));
        if (li < 1214) li = 1214;
        if (li > 1713) { NB_FAULT(c, NBIOT_FAULT_MCS_KEY); return; }
        /* utils.find_next_integer_index: the first transport block size that holds the buffer, one candidate per lane */
        const int buf = u->buffer;
        int ti = nbw_run([&]() -> int {
            NB_NOUNROLL
            for (int base = 0; base < NB_N_TBS; base += NB_LANES) {
                int i = base + nbw_lane();
                uint32_t m = nbw_ballot(i < NB_N_TBS && nb_tbs_list[i < NB_N_TBS ? i : 0] >= buf);
                if (m) return base + nbw_ffs(m) - 1;
            }
            return NB_N_TBS - 1;
        });
        tbs = nb_tbs_list[ti];
        const uint8_t *e = &c.mcs[((li - 1214) * NB_N_TBS + ti) * 3];
        I_tbs = NB_LDG(&e[0]); N_rep = NB_LDG(&e[1]); N_ru = NB_LDG(&e[2]);
    } else {
        N_rep = P.N_rep; I_tbs = P.I_tbs;
        int row = nb_itbs_to_row[I_tbs];
        int p = 7; for (int i = 0; i < 8; ++i) if (nb_tbs_table[row][i] >= u->buffer) { p = i; break; }
        if (c.conf.tx_all_buffer) { tbs = nb_tbs_table[row][p]; N_ru = nb_n_ru_list[p]; }
        else {
            int I_N_ru = P.N_ru & 7;
            if (p < I_N_ru) { N_ru = nb_n_ru_list[p]; tbs = nb_tbs_table[row][p]; }
            else { N_ru = nb_n_ru_list[I_N_ru]; tbs = nb_tbs_table[row][I_N_ru]; }
        }
    }
    int N_sc = P.N_sc;
    int NPDCCH_sf = nb_npdcch_sf_per_ce[CE_level];
    int reference_time = h->time + NPDCCH_sf;
    int n_sf = 0;
    int t_ul_end = nbc_node_allocate(c, P.carrier, reference_time, P.delay, N_sc, N_ru, N_rep, &N_sc, &n_sf);
    if (NB_FATAL(c)) return;
    if (t_ul_end) {                                            /* data_grant (node_b.py:667-693) */
        nbc_conn_remove(c, slot);
        u->I_tbs = (uint8_t)I_tbs; u->N_rep = (uint8_t)N_rep; u->N_ru = (uint8_t)N_ru; u->tbs = (uint16_t)tbs;
        if (u->retx_buffer > 0) u->bits = u->retx_buffer;
        else if (tbs >= u->buffer) { u->bits = u->buffer; u->buffer = 0; }
        else { u->bits = (uint16_t)(u->buffer < tbs ? u->buffer : tbs); u->buffer = (uint16_t)(u->buffer - u->bits); }
        u->where = NB_IN_SCHEDULED;
        h->n_scheduled += 1;
        nbc_pool_push(c, NB_KEY(t_ul_end, h->seq), NB_EV_NPUSCH_END, CE_level, slot);
        h->seq += 1;
        h->NPUSCH_resources += (int64_t)N_sc * n_sf;
        NB_TRACE(c, NBIOT_TR_RES_NPUSCH, CE_level, h->time, N_sc * n_sf, 0);
    } else if (new_data) {
        if (h->n_unfit < NBIOT_STEP_LIST) h->unfit_id[h->n_unfit] = u->id; else nbc_step_spill(c, 3, h->n_unfit, u->id);
        h->n_unfit += 1;
    }
    nbc_schedule_NPDCCH(c, NPDCCH_sf);
    nbc_update_state(c);
}

NB_FN1 void nbc_step_update(NbCtx &c, const uint8_t *a)          /* node_b.py:695-768 */
{
    NbHdr *h = nbc_hdr(c);
    h->thr_count = 0;
    int conf[16];
    NB_NOUNROLL
    for (int i = 0; i < 16; ++i) conf[i] = a[nb_nprach_item_idx[i]];
    int sc_max = NB_NC(c) == 1 ? 3 : 7;
    if (conf[1] > 7 || conf[2] > 7 || conf[3] > 10 || conf[4] > 15 || conf[5] > 14 || conf[6] > 14 || conf[7] > 6 || conf[10] > 6 ||
        conf[13] > 6 || conf[9] > sc_max || conf[12] > sc_max || conf[15] > sc_max) { NB_FAULT(c, NBIOT_FAULT_ACTION); return; }
    int th_C1 = nb_threshold_list[conf[5]], th_C0 = nb_threshold_list[conf[6]];
    if (th_C1 >= th_C0) th_C0 = 0;
    int rep_C2 = nb_thr_to_n_rep[NB_MIN_TH + 116], rep_C1 = 0, rep_C0 = 0;
    if (th_C1 < NB_NO_REP_TH) rep_C1 = nb_thr_to_n_rep[th_C1 + 116];
    if (th_C0 < NB_NO_REP_TH) rep_C0 = nb_thr_to_n_rep[th_C0 + 116];
    h->msg3_nrep[1] = th_C1 < NB_NO_REP_MSG3_TH ? nb_msg3_conf_nrep[th_C1 + 116] : 1;
    h->msg3_nrep[0] = th_C0 < NB_NO_REP_MSG3_TH ? nb_msg3_conf_nrep[th_C0 + 116] : 1;
    int RAR_window_size = nb_rar_window_list[conf[1]] * NB_NPDCCH_PERIOD;
    h->CE_thresholds[0] = th_C1; h->CE_thresholds[1] = th_C0;
    h->MAC_timer = nb_mac_timer_list[conf[2]] * NB_NPDCCH_PERIOD;
    h->preamble_trans_max = nb_transmax_list[conf[3]];
    h->probability_anchor = nb_panchor_den[conf[4]] ? NB_DIV(1.0, (double)nb_panchor_den[conf[4]]) : 0.0;
    const int args[3][3] = { { conf[7], rep_C0, conf[9] }, { conf[10], rep_C1, conf[12] }, { conf[13], rep_C2, conf[15] } };
    nbc_carrier_update_ra(c, args, th_C1, th_C0);
    NB_NOUNROLL
    for (int ce = 0; ce < 3; ++ce) h->RAR_window_size[ce] = RAR_window_size;
    NB_NOUNROLL
    for (int i = 0; i < 16; ++i) h->nprach_conf[i] = (uint8_t)conf[i];
    h->keys[1] = NB_KEY(h->cur_t + NB_NPRACH_UPDATE_PERIOD, h->seq);
    h->seq += 1;
    nbc_update_state(c);
}

/* ------------------------------------------------------------------ event selection */
/* candidate codes: 0 NPDCCH_arrival, 1 NPRACH_update, 2+ce NPRACH_start, 5+ce NPRACH_end, 8 MAC_timer, 16+i pool[i].
 * The minimum (time, insertion count) key over the fixed slots and the pool = the reference's heap pop. */
NB_FN1S uint64_t nbc_next_event(NbCtx &c, int *code)
{
    NbHdr *h = nbc_hdr(c);
    if (!h->mac_valid) nbc_mac_refresh(c);
    h->keys[8] = h->cont_n > 0 ? h->next_mac : NB_KEY_NONE;
    uint64_t best = NB_KEY_NONE; int bcode = -1;
#if NB_LANES == 1
    /* one lane: the fixed slots, then only the pool entries that hold an event (keys are unique, so no tie to break) */
    {
        NB_SMALL_LOOP
        for (int j = 0; j < 9; ++j) { uint64_t kj = h->keys[j]; if (kj < best) { best = kj; bcode = j; } }
        uint64_t used = ~h->pool_free & ((1ULL << NB_EV_POOL) - 1ULL);
        NB_NOUNROLL
        while (used) {
            int i = nbw_ffsll(used) - 1; used &= used - 1ULL;
            uint64_t kj = h->pool[i].key;
            if (kj < best) { best = kj; bcode = 16 + i; }
        }
    }
#else
    nbw_run([&]() -> int {
        nbw_sync();
        uint64_t k = NB_KEY_NONE; int cd = -1;
#if NB_LANES == 32
        /* one key per lane covers the fixed slots and pool[0 .. 32 - NB_FIXED_KEYS); the pool hands out its lowest free entry, so
         * the entries above that are in use only when more than 22 events are pending: a second pass then, and only then */
        {
            int j = nbw_lane();
            uint64_t kj = j < NB_FIXED_KEYS ? h->keys[j] : h->pool[j - NB_FIXED_KEYS].key;
            if (kj < k) { k = kj; cd = j < NB_FIXED_KEYS ? j : j + (16 - NB_FIXED_KEYS); }
            if ((~h->pool_free & ((1ULL << NB_EV_POOL) - 1ULL)) >> (32 - NB_FIXED_KEYS)) {
                j += 32;
                if (j < NB_FIXED_KEYS + NB_EV_POOL) {
                    kj = h->pool[j - NB_FIXED_KEYS].key;
                    if (kj < k) { k = kj; cd = j + (16 - NB_FIXED_KEYS); }
                }
            }
        }
#else
        NB_NOUNROLL
        for (int j = nbw_lane(); j < NB_FIXED_KEYS + NB_EV_POOL; j += NB_LANES) {
            uint64_t kj = j < NB_FIXED_KEYS ? h->keys[j] : h->pool[j - NB_FIXED_KEYS].key;
            if (kj < k) { k = kj; cd = j < NB_FIXED_KEYS ? j : j + (16 - NB_FIXED_KEYS); }
        }
#endif
        uint64_t m = nbw_min_u64(k);
        uint32_t who = nbw_ballot(k == m && cd >= 0);
        int src = who ? nbw_ffs(who) - 1 : 0;
        best = m; bcode = (int)nbw_shfl((uint32_t)cd, src);
        return 0;
    });
#endif
    *code = best == NB_KEY_NONE ? -1 : bcode;
    return best;
}

NB_FN void nbc_log_event(NbCtx &c, int t, int code)
{
    NbHdr *h = nbc_hdr(c);
    /* oracle numbering: 0 NPDCCH_arrival 1 NPRACH_update 2 RAR_window_end 3 NPRACH_start 4 NPRACH_end 5 Msg3 6 NPUSCH_end 7 MAC_timer */
    int type = code < 2 ? code : code < 5 ? 3 : code < 8 ? 4 : code == 8 ? 7 : (h->pool[code - 16].type == NB_EV_RAR_END ? 2 : h->pool[code - 16].type == NB_EV_MSG3 ? 5 : 6);
    if (h->events < c.evlog_cap) { c.evlog[2 * h->events] = t; c.evlog[2 * h->events + 1] = type; }
}

/* NodeB.advance_fel (node_b.py:191-204) + check_event (:248-261): run until a decision is needed */
NB_FN void nbc_advance_fel(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    for (;;) {
        int code; uint64_t key = nbc_next_event(c, &code);
        if (code < 0) { NB_FAULT(c, NBIOT_FAULT_EVENT_SLOTS); return; }
        int t = (int)(key >> 32);
        h->subframes_done += (int64_t)(t - h->time);
        h->time = t; h->cur_key = key;
        nbc_erase_past_sf(c, t);                               /* carrier_set.step(t) */
        if (NB_FATAL(c)) return;
        if (NB_EVLOG(c)) nbc_log_event(c, t, code);
        h->events += 1;
        if (code == 0) {
            h->keys[0] = NB_KEY_NONE;
            if (h->state == NB_ST_RAR_WINDOW || h->state == NB_ST_SCHEDULING) return;
            nbc_schedule_NPDCCH(c, h->NPDCCH_sf_left);
        } else if (code == 1) {
            h->keys[1] = NB_KEY_NONE; h->cur_t = t; h->state = NB_ST_NPRACH_UPDATE;
            return;
        } else if (code < 5) {
            int ce = code - 2;
            NbOccasion o = h->occ[ce][h->occ_head_start[ce] % NB_OCC_RING];
            h->occ_head_start[ce] += 1;
            nbc_occ_keys(h, ce);
            nbc_NPRACH_start(c, ce, o);
        } else if (code < 8) {
            int ce = code - 5;
            h->occ_head_end[ce] += 1;
            nbc_occ_keys(h, ce);
            nbc_NPRACH_end(c, ce, t);
        } else if (code == 8) {
            nbc_MAC_timeout(c, h->next_mac_idx);
        } else {
            int i = code - 16;
            NbEvent e = h->pool[i];
            nbc_pool_release(c, i);
            if (e.type == NB_EV_RAR_END) {
                h->cur_ce = e.ce; h->cur_wid = e.aux; h->cur_t = t; h->state = NB_ST_RAR_WINDOW_END;
                return;
            } else if (e.type == NB_EV_MSG3) nbc_msg3_event(c, e.aux, t);
            else { nbc_NPUSCH_end(c, e.aux, t); }
        }
        if (NB_FATAL(c)) return;
    }
}

/* ------------------------------------------------------------------ decision record */
NB_DEV double nbc_reward(NbCtx &c)                               /* perf_monitor.py:253-322 */
{
    NbHdr *h = nbc_hdr(c);
    int n = h->n_rx;
    switch (c.conf.reward_criteria) {
    case NB_RW_THROUGHPUT: return (double)h->thr_count;
    case NB_RW_AVERAGE_DELAY: return NB_DIV(NB_MUL(-1.0, h->rx_sum), (double)(n > 1 ? n : 1));
    case NB_RW_ACCUMULATED_DELAY: return NB_MUL(-1.0, h->rx_sum);
    case NB_RW_USERS: return (double)n;
    case NB_RW_INVD_USERS: return h->rx_inv_sum;
    case NB_RW_LOG_INVD_USERS: return NB_MUL(-1.0, h->rx_log_sum);
    default: return 0.0;
    }
}

/* materialise the decision record (reward + info dict of NodeB.get_node_info, node_b.py:278-418) in HBM;
 * only called when an external agent or the host will look at it (cold path) */
NB_FN1 void nbc_write_record(NbCtx &c, double reward, double occ_nprach, double occ_ra, double occ_npusch)
{
    NbHdr *h = nbc_hdr(c);
    nbiot_record_t *r = c.rec;
    nbw_run([&]() -> int {
        uint32_t *w = (uint32_t *)r;
        NB_NOUNROLL
        for (int i = nbw_lane(); i < (int)(sizeof(nbiot_record_t) / 4); i += NB_LANES) w[i] = 0u;
        nbw_sync();
        NB_NOUNROLL
        for (int j = nbw_lane(); j < NB_HORIZON; j += NB_LANES) r->carrier_busy[j] = (uint8_t)nbw_popc(*nbc_col(c, 0, h->t_now + j) & 0xFFFu);
        nbw_sync();
        return 0;
    });
    r->reward = reward; r->time = h->time; r->state = h->state; r->total_ues = h->conn_n;
    if (h->state == NB_ST_SCHEDULING || h->state == NB_ST_RAR_WINDOW_END) {
        r->n_sel = h->n_sel;
        NB_NOUNROLL
        for (int i = 0; i < h->n_sel; ++i) {
            NbUe *u = &c.ue[h->sel_slot[i]];
            r->ues[i] = u->id; r->connection_time[i] = h->time - u->t_connection; r->loss[i] = u->loss; r->buffer[i] = u->buffer;
            if (!u->new_data) { double v = u->sinr; if (!(v > -40.0)) v = -40.0; if (!(v < 30.0)) v = 30.0; r->sinr[i] = NB_ADD(v, 40.0); }
        }
        r->n_rx = h->n_rx; r->n_err = h->n_err; r->n_unfit = h->n_unfit; r->n_acc = h->n_acc;
        const int lim = NBIOT_STEP_LIST + nb_spill_cap(c.hist_cap);         /* 16 in the record, the rest in the side buffer */
        if (h->n_rx > lim || h->n_err > lim || h->n_unfit > lim || h->n_acc > lim)
            NB_FAULT(c, NBIOT_FAULT_LIST_TRUNC);
        nbw_run([&]() -> int {                                  /* one list position per lane */
            NB_NOUNROLL
            for (int i = nbw_lane(); i < NBIOT_STEP_LIST; i += NB_LANES) {
                if (i < h->n_rx) { r->rx_id[i] = h->rx_id[i]; r->rx_delay[i] = h->rx_delay[i]; }
                if (i < h->n_err) r->err_id[i] = h->err_id[i];
                if (i < h->n_unfit) r->unfit_id[i] = h->unfit_id[i];
                if (i < h->n_acc) { r->acc_id[i] = h->acc_id[i]; r->acc_delay[i] = h->acc_delay[i]; }
            }
            nbw_sync();
            return 0;
        });
    } else if (h->state == NB_ST_RAR_WINDOW) {
        int max_CE = nb_ce_for_sf_left[h->NPDCCH_sf_left];
        NB_NOUNROLL
        for (int i = 0; i < 3; ++i) {
            r->ues_per_CE[i] = i <= max_CE ? nbc_rar_sum(c, i) : 0;
            r->RAR_in[i] = h->RAR_in[i]; r->RAR_sent[i] = h->RAR_sent[i]; r->RAR_ids[i] = h->RAR_ids[i];
            r->RAR_detected[i] = h->RAR_detected[i]; r->RAR_failed[i] = h->RAR_failed[i]; r->action[i] = h->action3[i];
        }
        r->valid_CE_control = h->valid_CE_control; r->NPDCCH_sf_left = h->NPDCCH_sf_left; r->total_departures = h->total_departures;
        NB_NOUNROLL
        for (int i = 0; i < 16; ++i) r->nprach_conf[i] = h->nprach_conf[i];
    } else {
        int nd = h->n_departures;
        r->NPRACH_occupation = occ_nprach; r->RA_occupation = occ_ra; r->NPUSCH_occupation = occ_npusch;
        r->av_delay = nd ? NB_DIV((double)h->sum_delays, (double)nd) : 0.0;
        NB_NOUNROLL
        for (int ce = 0; ce < 3; ++ce) {
            r->preambles[ce] = (h->nprach_conf[9 + 3 * ce] + 1) * 12;
            int n = h->m3_cnt[ce]; double dn = (double)(n > 1 ? n : 1);
            r->msg3_detection[ce] = NB_DIV(h->m3_eff_sum[ce], dn);
            r->msg3_sent[ce] = NB_DIV((double)h->m3_sent_sum[ce], dn);
            r->msg3_received[ce] = NB_DIV((double)h->m3_recv_sum[ce], dn);
            int no = h->occ_n[ce]; double dno = (double)(no > 1 ? no : 1);
            r->detection_ratios[ce] = NB_DIV(NB_DIV((double)h->det_sum[ce], dno), (double)r->preambles[ce]);
            r->colision_ratios[ce] = NB_DIV(NB_DIV((double)h->col_sum[ce], dno), (double)r->preambles[ce]);
            r->n_occ[ce] = no;
            int lim = no < NBIOT_MAX_OCC ? no : NBIOT_MAX_OCC;
            NB_NOUNROLL
            for (int i = 0; i < lim; ++i) { r->det_hist[ce][i] = c.occ->det[ce][i]; r->col_hist[ce][i] = c.occ->col[ce][i]; }
        }
        NB_NOUNROLL
        for (int i = 0; i < 15; ++i) r->distribution[i] = h->distribution[i];
        r->departures = nd; r->n_incoming = h->n_incoming;
        NB_NOUNROLL
        for (int i = 0; i < 16; ++i) r->nprach_conf[i] = h->nprach_conf[i];
    }
    r->fault = h->fault;
    r->draws_lo = (int32_t)(uint32_t)h->draws; r->draws_hi = (int32_t)(uint32_t)(h->draws >> 32);
    r->events = h->events;
}

/* NodeB.get_node_info (node_b.py:278-418): its side effects always happen, the record is only materialised
 * when somebody will read it */
NB_FNM4 void nbc_node_info(NbCtx &c, int write_record, double reward)
{
    NbHdr *h = nbc_hdr(c);
    double occ_nprach = 0.0, occ_ra = 0.0, occ_npusch = 0.0;
    if (h->state == NB_ST_SCHEDULING || h->state == NB_ST_RAR_WINDOW_END) nbc_conn_select(c);
    else if (h->state == NB_ST_NPRACH_UPDATE && h->time != 0) {
        /* PerfMonitor.estimate_carrier_resources (perf_monitor.py:204-222) */
        int period = h->time - h->perf_t;
        int64_t total = 12LL * NB_NC(c) * period;
        int64_t signalling = h->NPRACH_resources + h->msg3_resources;
        if (write_record) {
            occ_nprach = NB_DIV((double)h->NPRACH_resources, (double)total);
            occ_ra = NB_DIV((double)signalling, (double)total);
            occ_npusch = NB_DIV((double)h->NPUSCH_resources, (double)(total - signalling));
        }
        h->msg3_resources = 0; h->NPRACH_resources = 0; h->NPUSCH_resources = 0;
        h->perf_t = h->time;
    }
    if (write_record) nbc_write_record(c, reward, occ_nprach, occ_ra, occ_npusch);
    if (h->state == NB_ST_RAR_WINDOW) { h->RAR_failed[0] = 0; h->RAR_failed[1] = 0; h->RAR_failed[2] = 0; }
    else if (h->state == NB_ST_NPRACH_UPDATE) {
        NB_NOUNROLL
        for (int ce = 0; ce < 3; ++ce) {
            h->m3_cnt[ce] = 0; h->m3_sent_sum[ce] = 0; h->m3_recv_sum[ce] = 0; h->m3_eff_sum[ce] = 0.0;
            h->occ_n[ce] = 0; h->det_sum[ce] = 0; h->col_sum[ce] = 0;
        }
        /* the period lists stay readable in the side buffers until the next departure / connection */
        h->n_departures = 0; h->n_incoming = 0; h->sum_delays = 0;
    }
}

/* ------------------------------------------------------------------ instance life cycle */
/* NodeB.__init__ / create_system values that survive reset() (node_b.py:87-104, population.py:41-49,
 * ue_generator.py:23-37) */
NB_FN void nbc_construct(NbCtx &c, uint64_t seed, const NbCtrlDev *ctrl)
{
    NbHdr *h = nbc_hdr(c);
    nbw_run([&]() -> int {
        uint32_t *w = (uint32_t *)h;
        NB_NOUNROLL
        for (int i = nbw_lane(); i < (int)(sizeof(NbHdr) / 4); i += NB_LANES) w[i] = 0u;
        nbw_sync();
        return 0;
    });
    h->seed = seed;
    c.occ->cluster_count = 0; c.occ->stat_n = 0; c.occ->cluster_cx = 0.0; c.occ->cluster_cy = 0.0;   /* Population.__init__ (:57-63): not reset() */
    c.occ->trace_n = 0; c.occ->pad_t = 0;
    h->NPDCCH_sf_left = NB_NPDCCH_SF;
    NB_NOUNROLL
    for (int i = 0; i < 16; ++i) h->nprach_conf[i] = nb_nprach_default_conf[i];
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) h->RAR_window_size[i] = 4 * NB_NPDCCH_PERIOD;
    h->valid_CE_control = 1;
    h->msg3_nrep[0] = 1; h->msg3_nrep[1] = 4; h->msg3_nrep[2] = 16;
    h->CE_thresholds[0] = -102; h->CE_thresholds[1] = -85;
    h->preamble_trans_max = 2; h->MAC_timer = 64;
    h->gen_p = 0.5; h->gen_counter = 0;
    NB_NOUNROLL
    for (int i = 0; i < 24; ++i) h->ctrl_action[i] = ctrl->init_action[i];
}

/* NodeB.reset (node_b.py:163-189) with MessageSwitch.reset (message_switch.py:48-54) */
NB_FN void nbc_reset(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    /* the random generator is an object outside the cell (system_creator.py:22): a reset does not rewind it */
    h->fault = 0; h->events = 0; h->subframes_done = 0;
    h->seq = 0; h->pool_free = (1ULL << NB_EV_POOL) - 1ULL;
    NB_NOUNROLL
    for (int i = 0; i < NB_FIXED_KEYS; ++i) h->keys[i] = NB_KEY_NONE;
    NB_NOUNROLL
    for (int i = 0; i < NB_EV_POOL; ++i) { h->pool[i].type = NB_EV_NONE; h->pool[i].key = NB_KEY_NONE; }
    /* population.reset */
    h->ra_n[0] = h->ra_n[1] = h->ra_n[2] = 0; h->cont_n = 0; h->ue_count = 0; h->mac_valid = 0; h->next_mac = NB_KEY_NONE; h->next_mac_idx = -1;
    int cap = c.ue_cap;
    nbw_run([&]() -> int {
        NB_NOUNROLL
        for (int i = nbw_lane(); i < cap; i += NB_LANES) c.free_stack[i] = (uint16_t)(cap - 1 - i);
        nbw_sync();
        return 0;
    });
    h->free_top = cap;
    /* carrier_set.reset (carrier.py:430-444) */
    h->t_now = 0; h->t_ahead = 0; h->sclog_n = 0; h->sclog_lo = 0; h->max_lookahead = 0; h->max_live = 0;
    NB_NOUNROLL
    for (int cc = 0; cc < 2; ++cc)
        NB_NOUNROLL
        for (int ce = 0; ce < 3; ++ce) {
            NbResource r; r.t_start = 0; r.t_next = 2; r.sf_to_go = 0; r.t_offset = 0; r.periodicity = 2; r.N_sf = (int16_t)ce; r.N_sc = 0; r.sc_offset = 0; r.mask = 0;
            h->res[cc][ce] = r;
        }
    NB_NOUNROLL
    for (int ce = 0; ce < 3; ++ce) { h->occ_head_start[ce] = h->occ_head_end[ce] = h->occ_tail[ce] = 0; }
    /* perf_monitor.reset (perf_monitor.py:41-59) */
    h->perf_t = 0; h->error_count = 0; h->attempts_count = 0; h->thr_count = 0; h->total_bits = 0;
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) { h->ue_arrivals[i] = 0; h->ue_departures[i] = 0; h->backoffs[i] = 0; }
    h->ue_connections = 0; h->av_delay = 0.0; h->msg3_resources = 0; h->NPUSCH_resources = 0; h->NPRACH_resources = 0;
    c.occ->stat_n = 0; c.occ->trace_n = 0;                     /* the histories restart (perf_monitor.py:61-84) */
    nbc_clear_info(c);
    /* ue_generator.reset (ue_generator.py:38-42) */
    double ratio = c.conf.ratio < 1.0 ? c.conf.ratio : 1.0;
    h->M_uniform = (int)NB_RINT(NB_MUL((double)c.conf.M, ratio));
    h->M_beta = c.conf.M - h->M_uniform; h->M_beta_max = h->M_beta; h->gen_t = 0;
    /* access_procedure.reset (access_procedure.py:28-35) */
    h->probability_anchor = 0.0;
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) { h->detected_pre[i] = 0; h->collided_pre[i] = 0; h->contenders[i] = 0; h->RAR_counter[i] = 0; h->occ_n[i] = 0; h->det_sum[i] = 0; h->col_sum[i] = 0; }
    /* reset_storage (node_b.py:131-161) */
    NB_NOUNROLL
    for (int i = 0; i < 3; ++i) {
        h->rar_n[i] = 0; h->rar_tot[i] = 0; h->RAR_in[i] = 0; h->RAR_ids[i] = 0; h->RAR_attmp[i] = 0; h->RAR_sent[i] = 0; h->RAR_detected[i] = 0; h->RAR_failed[i] = 0;
        h->m3_cnt[i] = 0; h->m3_sent_sum[i] = 0; h->m3_recv_sum[i] = 0; h->m3_eff_sum[i] = 0.0;
    }
    h->conn_n = 0; h->conn_order = 0; h->n_sel = 0; h->n_scheduled = 0;
    h->n_incoming = 0; h->n_departures = 0; h->sum_delays = 0;
    NB_NOUNROLL
    for (int i = 0; i < 15; ++i) h->distribution[i] = 0.0;
    h->k_dist = 0;
    h->state = NB_ST_NPRACH_UPDATE; h->time = 0;
    h->keys[1] = NB_KEY(0, 0); h->keys[0] = NB_KEY(0, 1); h->seq = 2;
    nbc_advance_fel(c);
    nbc_node_info(c, 1, 0.0);
}

/* NodeB.step (node_b.py:228-245): apply the action for the current state, run to the next decision */
NB_DEV void nbc_node_step(NbCtx &c, const uint8_t *a)
{
    NbHdr *h = nbc_hdr(c);
    switch (h->state) {
    case NB_ST_SCHEDULING: nbc_step_data(c, a); break;
    case NB_ST_RAR_WINDOW: nbc_step_RAR_window(c, a); break;
    case NB_ST_RAR_WINDOW_END: nbc_step_RAR_end(c, a); break;
    case NB_ST_NPRACH_UPDATE: nbc_step_update(c, a); break;
    default: break;
    }
    if (!NB_FATAL(c)) nbc_advance_fel(c);
}

/* fixed actions of internal agents: one value per set bit of the table's mask */
NB_FNM void nbc_apply_writes(NbCtx &c, const NbWrites *w)
{
    uint32_t m = w->mask;                                      /* tables live in shared memory: plain loads */
    NbHdr *h = nbc_hdr(c);
    NB_NOUNROLL
    while (m) { int i = nbw_ffs(m) - 1; m &= m - 1; h->ctrl_action[i] = w->val[i]; }
}

/* Controller.step + to_next_step (controller.py:234-288), or Controller.run_until (:185-195) when
 * there is no external agent, for one instance.
 *   ext_action : the external agent's k items, scattered into the shared 23-vector followed by its
 *                `next` chain (NULL: the agents of the current state act, as in single_step :156-165)
 *   until_time : stop once time >= until_time (tested before every NodeB step, like run_until);
 *                negative: no time bound
 *   max_steps  : bound on NodeB steps in this call
 * ctrl->state_writes[s] holds the fixed actions of the internal agents of state s; for a state of
 * the external agent it holds only the agents listed before it (controller.py:283-288).
 * Without an external action (run_until / single_step, controller.py:156-195) EVERY agent of a state acts, the one
 * registered as external with its fixed action (ctrl->full_writes), and no state hands control back.
 * Returns the number of NodeB steps taken; the decision record is rewritten iff it is > 0. */
NB_FN int nbc_controller_advance(NbCtx &c, const NbCtrlDev *ctrl, const uint8_t *ext_action, int until_time, int max_steps)
{
    NbHdr *h = nbc_hdr(c);
    int steps = 0;
    if (NB_FATAL(c)) { c.rec->fault = h->fault; return 0; }
    const NbWrites *sw = ext_action ? ctrl->state_writes : ctrl->full_writes;
    const uint32_t ext_mask = ext_action ? ctrl->ext_state_mask : 0u;
    if (ext_action) {
        NB_NOUNROLL
        for (int i = 0; i < ctrl->ext_k; ++i) h->ctrl_action[ctrl->ext_a_mask[i]] = ext_action[i];
        nbc_apply_writes(c, &ctrl->ext_post);
    } else nbc_apply_writes(c, &sw[h->state]);
    NB_NOUNROLL
    while (steps < max_steps && !(until_time >= 0 && h->time >= until_time)) {
        if (steps > 0) {
            nbc_node_info(c, 0, 0.0);                          /* an internal agent looks at nothing */
            nbc_apply_writes(c, &sw[h->state]);
        }
        nbc_node_step(c, h->ctrl_action);
        steps += 1;
        if (NB_FATAL(c)) break;
        if ((ext_mask >> h->state) & 1u) { nbc_apply_writes(c, &sw[h->state]); break; }
        if (ext_action && until_time < 0 && steps >= NBIOT_STALL_STEPS) { NB_FAULT(c, NBIOT_FAULT_STALL); break; }   /* never hang the GPU */
    }
    if (steps > 0) nbc_node_info(c, 1, nbc_reward(c));
    else c.rec->fault = h->fault;
    return steps;
}

#ifdef NB_TASKS
/* ------------------------------------------------------------------ tasks
 * The control flow of one instance -- Controller.step / run_until (controller.py:185-195, :234-288) around
 * NodeB.step (node_b.py:228-245) around NodeB.advance_fel (:191-204) -- cut into TASKS: one heavy handler each
 * (a NodeB step of one decision state, an NPRACH occasion, a msg3 or NPUSCH reception, the final record).
 * Everything between two handlers (event selection, the light events, the controller's loop tests and the
 * internal agents) is the glue that runs at the end of every task and names the next one. Nothing lives in
 * registers or in the per-warp context between tasks -- the control state is in the header (NbHdr.tk_*) -- so any
 * warp of any SM can run the next task of any instance: the migrating-instance kernel (nbiot_queue.cu) moves an
 * instance to an SM whose warps are all in that task's handler, which keeps an SM's instruction working set to one
 * handler + glue instead of the whole event cycle. nbc_task_chain runs the same tasks in turn on one owner
 * (CPU builds) and is step for step the monolithic nbc_controller_advance above. */
enum { NB_TK_DONE = -1, NB_TK_BEGIN = 0, NB_TK_STEP_DATA = 1, NB_TK_STEP_RARW = 2, NB_TK_STEP_RAREND = 3, NB_TK_STEP_UPDATE = 4,
       NB_TK_NPRACH_START = 5, NB_TK_MSG3 = 6, NB_TK_NPUSCH_END = 7, NB_TK_FINISH = 8, NB_TK_GLUE = 9, NB_TK_N = 10 };
/* NB_TASK_SPLIT_GLUE = 1: the glue is a task class of its own (a handler names NB_TK_GLUE, the glue names the next handler), which
 * halves the code an SM of the migrating-instance kernel walks through per class at the price of twice the hops */
#ifndef NB_TASK_SPLIT_GLUE
#define NB_TASK_SPLIT_GLUE 0
#endif
enum { NB_PH_FEL = 0, NB_PH_STEPPED = 1 };

/* NodeB.advance_fel (node_b.py:191-204) + check_event (:248-261): pop events, handling the light ones in place,
 * until a heavy one is due (returns its task class, arguments in NbHdr.tk_*) or a decision is needed (returns 0) */
NB_FN int nbc_fel_next(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    for (;;) {
        if (NB_FATAL(c)) return 0;
        int code; uint64_t key = nbc_next_event(c, &code);
        if (code < 0) { NB_FAULT(c, NBIOT_FAULT_EVENT_SLOTS); return 0; }
        int t = (int)(key >> 32);
        h->subframes_done += (int64_t)(t - h->time);
        h->time = t; h->cur_key = key;
        nbc_erase_past_sf(c, t);                               /* carrier_set.step(t) */
        if (NB_FATAL(c)) return 0;
        if (NB_EVLOG(c)) nbc_log_event(c, t, code);
        h->events += 1;
        if (code == 0) {
            h->keys[0] = NB_KEY_NONE;
            if (h->state == NB_ST_RAR_WINDOW || h->state == NB_ST_SCHEDULING) return 0;
            nbc_schedule_NPDCCH(c, h->NPDCCH_sf_left);
        } else if (code == 1) {
            h->keys[1] = NB_KEY_NONE; h->cur_t = t; h->state = NB_ST_NPRACH_UPDATE;
            return 0;
        } else if (code < 5) {
            int ce = code - 2;
            h->tk_occ = h->occ[ce][h->occ_head_start[ce] % NB_OCC_RING];
            h->occ_head_start[ce] += 1;
            nbc_occ_keys(h, ce);
            h->tk_a = ce;
            return NB_TK_NPRACH_START;
        } else if (code < 8) {
            int ce = code - 5;
            h->occ_head_end[ce] += 1;
            nbc_occ_keys(h, ce);
            nbc_NPRACH_end(c, ce, t);
        } else if (code == 8) {
            nbc_MAC_timeout(c, h->next_mac_idx);
        } else {
            int i = code - 16;
            NbEvent e = h->pool[i];
            nbc_pool_release(c, i);
            if (e.type == NB_EV_RAR_END) {
                h->cur_ce = e.ce; h->cur_wid = e.aux; h->cur_t = t; h->state = NB_ST_RAR_WINDOW_END;
                return 0;
            }
            h->tk_a = e.aux; h->tk_b = t;
            return e.type == NB_EV_MSG3 ? NB_TK_MSG3 : NB_TK_NPUSCH_END;
        }
#ifdef NB_GLUE_ONE_POP
        return NB_TK_GLUE;                                     /* one event per glue task: the lanes of a warp pop in lock step */
#endif
    }
}

NB_DEV int nbc_step_class(int state)                             /* NodeB.step (node_b.py:228-245): the handler of a decision state */
{
    return state == NB_ST_SCHEDULING ? NB_TK_STEP_DATA : state == NB_ST_RAR_WINDOW ? NB_TK_STEP_RARW :
           state == NB_ST_RAR_WINDOW_END ? NB_TK_STEP_RAREND : state == NB_ST_NPRACH_UPDATE ? NB_TK_STEP_UPDATE : 0;
}

/* between two handlers: the rest of advance_fel, then the controller's loop (controller.py:234-288) */
NB_FN int nbc_task_glue(NbCtx &c)
{
    NbHdr *h = nbc_hdr(c);
    const NbCtrlDev *ctrl = c.ctrl;
    for (;;) {
        if (h->tk_phase == NB_PH_FEL) {
            int k = nbc_fel_next(c);
            if (k) return k;
            h->tk_phase = NB_PH_STEPPED;
        }
        /* a NodeB step has run to the next decision */
        h->tk_steps += 1;
        if (NB_FATAL(c)) return NB_TK_FINISH;
        const NbWrites *sw = c.tk_internal ? ctrl->full_writes : ctrl->state_writes;
        if (!c.tk_internal && ((ctrl->ext_state_mask >> h->state) & 1u)) { nbc_apply_writes(c, &sw[h->state]); return NB_TK_FINISH; }
        if (!(h->tk_steps < c.tk_max_steps && !(c.tk_until >= 0 && h->time >= c.tk_until))) return NB_TK_FINISH;
        if (!c.tk_internal && c.tk_until < 0 && h->tk_steps >= NBIOT_STALL_STEPS) { NB_FAULT(c, NBIOT_FAULT_STALL); return NB_TK_FINISH; }
        nbc_node_info(c, 0, 0.0);                              /* an internal agent looks at nothing */
        nbc_apply_writes(c, &sw[h->state]);
        h->tk_phase = NB_PH_FEL;
        int k = nbc_step_class(h->state);
        if (k) return k;
    }
}

/* first task of a launch: the external action (or the internal agents of the current state), then the first handler.
 * c.tk_until / c.tk_max_steps hold the launch's bounds (set by whoever fills the context). */
NB_FN int nbc_task_begin(NbCtx &c, const uint8_t *ext_action)
{
    NbHdr *h = nbc_hdr(c);
    const NbCtrlDev *ctrl = c.ctrl;
    h->tk_steps = 0; h->tk_phase = NB_PH_FEL;
    if (NB_FATAL(c)) return NB_TK_FINISH;
    if (ext_action) {
        NB_NOUNROLL
        for (int i = 0; i < ctrl->ext_k; ++i) h->ctrl_action[ctrl->ext_a_mask[i]] = ext_action[i];
        nbc_apply_writes(c, &ctrl->ext_post);
    } else nbc_apply_writes(c, &ctrl->full_writes[h->state]);
    if (!(0 < c.tk_max_steps && !(c.tk_until >= 0 && h->time >= c.tk_until))) return NB_TK_FINISH;
    int k = nbc_step_class(h->state);
    if (k) return k;
    h->tk_phase = NB_PH_FEL;                                   /* a state without a handler (node_b.py:228-245 `default`): advance_fel only */
    return nbc_task_glue(c);
}

/* run task k (k != NB_TK_BEGIN) of the instance and the glue behind it; returns the next task (NB_TK_DONE after NB_TK_FINISH) */
NB_DEV int nbc_task_run(NbCtx &c, int k)
{
    NbHdr *h = nbc_hdr(c);
    switch (k) {
    case NB_TK_STEP_DATA: nbc_step_data(c, h->ctrl_action); break;
    case NB_TK_STEP_RARW: nbc_step_RAR_window(c, h->ctrl_action); break;
    case NB_TK_STEP_RAREND: nbc_step_RAR_end(c, h->ctrl_action); break;
    case NB_TK_STEP_UPDATE: nbc_step_update(c, h->ctrl_action); break;
    case NB_TK_NPRACH_START: nbc_NPRACH_start(c, h->tk_a, h->tk_occ); break;
    case NB_TK_MSG3: nbc_msg3_event(c, h->tk_a, h->tk_b); break;
    case NB_TK_NPUSCH_END: nbc_NPUSCH_end(c, h->tk_a, h->tk_b); break;
    case NB_TK_GLUE: return nbc_task_glue(c);
    default:                                                   /* NB_TK_FINISH */
        if (h->tk_steps > 0) nbc_node_info(c, 1, nbc_reward(c));
        else c.rec->fault = h->fault;
        return NB_TK_DONE;
    }
#if NB_TASK_SPLIT_GLUE
    return NB_TK_GLUE;
#else
    return nbc_task_glue(c);
#endif
}

/* the chain run by one owner; returns the NodeB steps taken */
NB_FN int nbc_task_chain(NbCtx &c, const uint8_t *ext_action, int until_time, int max_steps)
{
    c.tk_until = until_time; c.tk_max_steps = max_steps; c.tk_internal = ext_action ? 0 : 1;
    int k = nbc_task_begin(c, ext_action);
    NB_NOUNROLL
    while (k != NB_TK_DONE) k = nbc_task_run(c, k);
    return nbc_hdr(c)->tk_steps;
}
#endif /* NB_TASKS */

#endif /* NBIOT_CORE_CUH */
