// 32-lane warp emulation on CPU fibers (tests/_hostemu only; a debugging aid, never a product path).
// nbw_emu_run() runs one lane-parallel section of nbiot_core.cuh on 32 cooperatively scheduled
// fibers; the collectives are rendezvous points with the semantics of the CUDA *_sync intrinsics.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

static ucontext_t g_main, g_lane[32];
static char *g_stack[32];
static int g_cur = -1;
static bool g_done[32];
static void (*g_tramp)(void *, int);
static void *g_closure;
static uint32_t g_xbuf[2][32];
static int g_cnt[2];
static uint64_t g_tag[2] = { ~0ULL, ~0ULL }, g_gen[32];
static const size_t STACK = 256 * 1024;

static void lane_entry(int lane)
{
    g_tramp(g_closure, lane);
    g_done[lane] = true;
    swapcontext(&g_lane[lane], &g_main);
}

void nbw_emu_mismatch(const char *what) { fprintf(stderr, "hostemu: %s\n", what); abort(); }

void nbw_emu_run(void (*tramp)(void *, int), void *closure)
{
    if (g_cur >= 0) nbw_emu_mismatch("nested lane-parallel section");
    g_tramp = tramp; g_closure = closure;
    uint64_t g0 = g_gen[0];
    for (int l = 0; l < 32; ++l) {
        if (!g_stack[l]) g_stack[l] = (char *)malloc(STACK);
        getcontext(&g_lane[l]);
        g_lane[l].uc_stack.ss_sp = g_stack[l];
        g_lane[l].uc_stack.ss_size = STACK;
        g_lane[l].uc_link = &g_main;
        makecontext(&g_lane[l], (void (*)())lane_entry, 1, l);
        g_done[l] = false;
        if (g_gen[l] != g0) nbw_emu_mismatch("lanes left the previous section at different collectives");
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < 32; ++l)
            if (!g_done[l]) { any = true; g_cur = l; swapcontext(&g_main, &g_lane[l]); }
        if (!any) break;
    }
    g_cur = -1;
    for (int l = 1; l < 32; ++l) if (g_gen[l] != g_gen[0]) nbw_emu_mismatch("lanes executed a different number of collectives");
}

static const uint32_t *exchange(uint32_t v)
{
    int l = g_cur;
    if (l < 0) nbw_emu_mismatch("collective outside a lane-parallel section");
    uint64_t G = g_gen[l]++;
    int b = (int)(G & 1);
    if (g_tag[b] != G) { g_tag[b] = G; g_cnt[b] = 0; }
    g_xbuf[b][l] = v;
    g_cnt[b]++;
    while (g_cnt[b] < 32) {
        // a finished lane can never arrive: that is a divergence bug in the section
        for (int k = 0; k < 32; ++k) if (g_done[k]) nbw_emu_mismatch("a lane returned while others wait in a collective");
        swapcontext(&g_lane[l], &g_main);
        g_cur = l;
    }
    return g_xbuf[b];
}

int nbw_lane() { return g_cur < 0 ? 0 : g_cur; }
void nbw_sync() { exchange(0); }
uint32_t nbw_ballot(int p) { const uint32_t *x = exchange(p ? 1u : 0u); uint32_t m = 0; for (int i = 0; i < 32; ++i) m |= (x[i] & 1u) << i; return m; }
uint32_t nbw_shfl(uint32_t v, int src) { const uint32_t *x = exchange(v); return x[src & 31]; }
uint32_t nbw_shfl_xor(uint32_t v, int m) { int l = g_cur; const uint32_t *x = exchange(v); return x[(l ^ m) & 31]; }
