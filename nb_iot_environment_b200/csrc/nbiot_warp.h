/*
 * nbiot_warp.h -- the warp-level vocabulary of the simulation core.
 *
 * One warp owns one cell instance. Order-dependent control logic is executed uniformly
 * by all 32 lanes (same values, same addresses); data-parallel pieces (list scans,
 * ordered compaction, histogramming, grid fills, min-key search) spread over the lanes
 * and use the collectives below. Every lane-divergent region ends with nbw_sync().
 *
 * Three builds of the same source:
 *   __CUDA_ARCH__            : the product -- real shuffles / ballots on sm_100a.
 *   NBIOT_EMU_LANES == 32    : tests/_hostemu -- lanes are CPU fibers, collectives are
 *                              rendezvous points (a debugging aid to exercise the exact
 *                              device control flow without a GPU; never used by the package).
 *   NBIOT_EMU_LANES == 1     : tests/_hostemu fast mode -- a single lane, collectives are
 *                              identities.
 */
#ifndef NBIOT_WARP_H
#define NBIOT_WARP_H

#include <stdint.h>

#if defined(__CUDACC__)

#define NB_LANES 32
#define NB_DEV __device__ __forceinline__
#define NB_FN static __device__ __noinline__
/* functions with a single call site: NB_INLINE_SINGLE = 0 keeps them out of line, 1 inlines the small ones, 2 all */
#ifndef NB_INLINE_SINGLE
#define NB_INLINE_SINGLE 1
#endif
/* small functions with several call sites: NB_INLINE_MULTI = 1 inlines the tiny ones, 2 also the mid-size ones */
#ifndef NB_INLINE_MULTI
#define NB_INLINE_MULTI 3
#endif
#if NB_INLINE_MULTI >= 3
#define NB_FNM3 __device__ __forceinline__
#else
#define NB_FNM3 static __device__ __noinline__
#endif
#if NB_INLINE_MULTI >= 4
#define NB_FNM4 __device__ __forceinline__
#else
#define NB_FNM4 static __device__ __noinline__
#endif
#if NB_INLINE_MULTI >= 1
#define NB_FNM __device__ __forceinline__
#else
#define NB_FNM static __device__ __noinline__
#endif
#if NB_INLINE_MULTI >= 2
#define NB_FNM2 __device__ __forceinline__
#else
#define NB_FNM2 static __device__ __noinline__
#endif
#if NB_INLINE_SINGLE >= 1
#define NB_FN1S __device__ __forceinline__
#else
#define NB_FN1S static __device__ __noinline__
#endif
#if NB_INLINE_SINGLE >= 2
#define NB_FN1 __device__ __forceinline__
#else
#define NB_FN1 static __device__ __noinline__
#endif
#define NB_FULL 0xFFFFFFFFu
#if defined(NB_TPC) && defined(NB_TPC_UNROLL_ALL)      /* measurement build: the compiler's own unrolling in the thread-per-cell kernel */
#define NB_NOUNROLL
#else
#define NB_NOUNROLL _Pragma("unroll 1")
#endif
/* short fixed-trip loops over header arrays (event keys, NPRACH resources): rolled in the warp-per-cell kernels, which are short of instruction
 * fetch; unrolled in the thread-per-cell kernel, which waits on memory and wants the independent loads in flight together */
#if defined(NB_TPC) && defined(NB_TPC_MLP)
#define NB_SMALL_LOOP _Pragma("unroll")
#else
#define NB_SMALL_LOOP NB_NOUNROLL
#endif
#ifdef NB_TPC
/* thread-per-cell build (nbiot_simt.cu): one THREAD owns a cell, a warp carries 32 unrelated cells. The lane-parallel sections
 * degenerate to scalar loops (a "warp" of one lane, exactly the NBIOT_EMU_LANES == 1 build the CPU tests run), nothing is
 * exchanged between the lanes, and divergence between the 32 cells of a warp is left to the hardware. */
#undef NB_LANES
#define NB_LANES 1
NB_DEV int nbw_lane() { return 0; }
NB_DEV void nbw_sync() {}
NB_DEV uint32_t nbw_ballot(int p) { return p ? 1u : 0u; }
NB_DEV uint32_t nbw_shfl(uint32_t v, int) { return v; }
NB_DEV uint32_t nbw_shfl_xor(uint32_t v, int) { return v; }
NB_DEV uint32_t nbw_atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
#else
NB_DEV int nbw_lane() { return (int)(threadIdx.x & 31u); }
NB_DEV void nbw_sync() { __syncwarp(NB_FULL); }
NB_DEV uint32_t nbw_ballot(int p) { return __ballot_sync(NB_FULL, p); }
NB_DEV uint32_t nbw_shfl(uint32_t v, int src) { return __shfl_sync(NB_FULL, v, src); }
NB_DEV uint32_t nbw_shfl_xor(uint32_t v, int m) { return __shfl_xor_sync(NB_FULL, v, m); }
NB_DEV uint32_t nbw_atomic_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
#endif
NB_DEV int nbw_popc(uint32_t v) { return __popc(v); }
NB_DEV int nbw_ffs(uint32_t v) { return __ffs((int)v); }
NB_DEV int nbw_popcll(uint64_t v) { return __popcll(v); }
NB_DEV int nbw_ffsll(uint64_t v) { return __ffsll((long long)v); }
#define NB_LDG(p) __ldg(p)

#else /* host emulation */

#ifndef NBIOT_EMU_LANES
#error "host build of the simulation core needs NBIOT_EMU_LANES (tests/_hostemu only)"
#endif
#define NB_LANES NBIOT_EMU_LANES
#define NB_DEV static inline
#define NB_FN static
#define NB_FN1 static
#define NB_FN1S static
#define NB_FNM static
#define NB_FNM2 static
#define NB_FNM3 static
#define NB_FNM4 static
#define NB_LDG(p) (*(p))
#define NB_NOUNROLL
#define NB_SMALL_LOOP
static inline int nbw_popc(uint32_t v) { return __builtin_popcount(v); }
static inline int nbw_ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int nbw_popcll(uint64_t v) { return __builtin_popcountll(v); }
static inline int nbw_ffsll(uint64_t v) { return __builtin_ffsll((long long)v); }
#if NBIOT_EMU_LANES == 1
static inline int nbw_lane() { return 0; }
static inline void nbw_sync() {}
static inline uint32_t nbw_ballot(int p) { return p ? 1u : 0u; }
static inline uint32_t nbw_shfl(uint32_t v, int src) { (void)src; return v; }
static inline uint32_t nbw_shfl_xor(uint32_t v, int m) { (void)m; return v; }
static inline uint32_t nbw_atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
#else
/* provided by the fiber scheduler in tests/_hostemu/emu_fibers.cpp */
int nbw_lane();
void nbw_sync();
uint32_t nbw_ballot(int p);
uint32_t nbw_shfl(uint32_t v, int src);
uint32_t nbw_shfl_xor(uint32_t v, int m);
static inline uint32_t nbw_atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
#endif

#endif

/* nbw_run(f): execute the lane-parallel section f. Product build: plain call (all 32 lanes are
 * already here). 32-lane emulation: the scalar code around it runs once, f runs on 32 fibers and
 * must return the same value on every lane. */
#if defined(__CUDACC__) || NBIOT_EMU_LANES == 1
template <class F> NB_DEV auto nbw_run(F f) -> decltype(f()) { return f(); }
#else
void nbw_emu_run(void (*tramp)(void *, int), void *closure);   /* runs tramp(closure, lane) on 32 fibers */
template <class F> struct NbwEmuBox { F *f; decltype((*(F *)0)()) out[32]; };
template <class F> static void nbw_emu_tramp(void *p, int lane) { auto *b = (NbwEmuBox<F> *)p; b->out[lane] = (*b->f)(); }
void nbw_emu_mismatch(const char *what);
template <class F> static inline auto nbw_run(F f) -> decltype(f())
{
    NbwEmuBox<F> box; box.f = &f;
    nbw_emu_run(&nbw_emu_tramp<F>, &box);
    for (int l = 1; l < 32; ++l) if (!(box.out[l] == box.out[0])) nbw_emu_mismatch("non-uniform result of a lane-parallel section");
    return box.out[0];
}
#endif

/* ---- derived collectives (same code in every build) ---- */
NB_DEV uint32_t nbw_lt_mask() { return (NB_LANES == 32) ? ((1u << (nbw_lane() & 31)) - 1u) : 0u; }

NB_DEV uint64_t nbw_shfl64(uint64_t v, int src)
{
    uint32_t lo = nbw_shfl((uint32_t)v, src), hi = nbw_shfl((uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

NB_DEV uint64_t nbw_min_u64(uint64_t v)
{
#if NB_LANES == 1
    return v;
#elif defined(__CUDA_ARCH__)
    /* two redux.sync steps: the smallest high word, then the smallest low word among its holders */
    uint32_t hi = (uint32_t)(v >> 32), mh = __reduce_min_sync(NB_FULL, hi);
    uint32_t ml = __reduce_min_sync(NB_FULL, hi == mh ? (uint32_t)v : 0xFFFFFFFFu);
    return ((uint64_t)mh << 32) | ml;
#else
    for (int m = NB_LANES / 2; m > 0; m >>= 1) {
        uint32_t lo = nbw_shfl_xor((uint32_t)v, m), hi = nbw_shfl_xor((uint32_t)(v >> 32), m);
        uint64_t o = ((uint64_t)hi << 32) | lo;
        v = o < v ? o : v;
    }
    return v;
#endif
}

NB_DEV uint32_t nbw_or_u32(uint32_t v)
{
#if NB_LANES == 1
    return v;
#elif defined(__CUDA_ARCH__)
    return __reduce_or_sync(NB_FULL, v);
#else
    for (int m = NB_LANES / 2; m > 0; m >>= 1) v |= nbw_shfl_xor(v, m);
    return v;
#endif
}

NB_DEV int32_t nbw_add_i32(int32_t v)
{
#if NB_LANES == 1
    return v;
#elif defined(__CUDA_ARCH__)
    return __reduce_add_sync(NB_FULL, v);
#else
    for (int m = NB_LANES / 2; m > 0; m >>= 1) v += (int32_t)nbw_shfl_xor((uint32_t)v, m);
    return v;
#endif
}

#endif /* NBIOT_WARP_H */
