/*
 * nbiot_math.h -- deterministic double-precision elementary functions.
 *
 * The simulated cell thresholds and rounds floating-point values (RSRP < th,
 * rint(snr*10), u < p; reference system/channel.py:76-79,156-176,
 * system/population.py:122-128), so a 1-ulp difference between the host libm and
 * the CUDA libm would eventually become an integer divergence between the GPU path
 * and its CPU checker. Everything here is built from IEEE-754 correctly rounded
 * primitives only (+ - * / sqrt fma rint) in a fixed order, so the same source gives
 * the same bits under gcc and under nvcc (device code uses the *_rn intrinsics, which
 * are never contracted; host code must be compiled with -ffp-contract=off).
 * Accuracy is ~1-3 ulp, i.e. ~1e-15 relative against numpy/libm.
 */
#ifndef NBIOT_MATH_H
#define NBIOT_MATH_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
/* the transcendental kernels and the Philox block stay out of line on the device: one copy each keeps
 * the simulation kernel inside the instruction cache */
#define NB_HD_CALL static __host__ __device__ __noinline__
#else
#define NB_HD static inline
#define NB_HD_CALL static inline
#endif

#if defined(__FP_FAST_FMA) && !defined(__CUDACC__) && !defined(NBIOT_NO_CONTRACT)
#error "compile host code with -ffp-contract=off -DNBIOT_NO_CONTRACT (bit-exact host/device arithmetic)"
#endif

#ifdef __CUDA_ARCH__
#define NB_MUL(a, b) __dmul_rn((a), (b))
#define NB_ADD(a, b) __dadd_rn((a), (b))
#define NB_SUB(a, b) __dsub_rn((a), (b))
#define NB_DIV(a, b) nb_ddiv((a), (b))
#define NB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define NB_SQRT(a) nb_dsqrt((a))
#define NB_RINT(a) rint((a))
#define NB_FLOOR(a) floor((a))
#else
#define NB_MUL(a, b) ((a) * (b))
#define NB_ADD(a, b) ((a) + (b))
#define NB_SUB(a, b) ((a) - (b))
#define NB_DIV(a, b) ((a) / (b))
#define NB_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define NB_SQRT(a) __builtin_sqrt((a))
#define NB_RINT(a) __builtin_rint((a))
#define NB_FLOOR(a) __builtin_floor((a))
#endif

#define NB_LN2_HI 0x1.62e42fee00000p-1   /* 32 significant bits of ln 2            */
#define NB_LN2_LO 0x1.a39ef35793c76p-33  /* ln 2 - NB_LN2_HI                       */
#define NB_INV_LN2 0x1.71547652b82fep+0
#define NB_LN10 0x1.26bb1bbb55516p+1
#define NB_LN10_LO (-0x1.f48ad494ea3e9p-53)
#define NB_INV_LN10 0x1.bcb7b1526e50ep-2
#define NB_PI_2 0x1.921fb54442d18p+0
#define NB_PI 0x1.921fb54442d18p+1
#define NB_RAD2DEG 0x1.ca5dc1a63c1f8p+5  /* 180.0 / pi as numpy's degrees() uses it */
#define NB_SQRT2 0x1.6a09e667f3bcdp+0

#ifdef __CUDACC__
/* IEEE division and square root expand to ~40 instructions each: keep one copy */
static __device__ __noinline__ double nb_ddiv_dev(double a, double b) { return __ddiv_rn(a, b); }
static __device__ __noinline__ double nb_dsqrt_dev(double a) { return __dsqrt_rn(a); }
#endif
#ifdef __CUDA_ARCH__
#define nb_ddiv(a, b) nb_ddiv_dev((a), (b))
#define nb_dsqrt(a) nb_dsqrt_dev((a))
#endif

NB_HD uint64_t nb_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
NB_HD double nb_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
NB_HD double nb_inf(void) { return nb_u2d(0x7ff0000000000000ULL); }

/* natural logarithm, x > 0 and normal (x == 0 gives -inf) */
NB_HD_CALL double nb_log(double x)
{
    if (x == 0.0) return -nb_inf();
    uint64_t bits = nb_d2u(x);
    int k = (int)((bits >> 52) & 0x7ffu) - 1023;
    double m = nb_u2d((bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL); /* [1,2) */
    if (m > NB_SQRT2) { m = NB_MUL(m, 0.5); k += 1; }                           /* [.707,1.414] */
    double f = NB_SUB(m, 1.0);                 /* exact */
    double s = NB_DIV(f, NB_ADD(2.0, f));
    double z = NB_MUL(s, s);
    /* ln m = 2 atanh(s) = 2s (1 + z/3 + z^2/5 + ...), z <= 0.0295 */
    double p = 1.0 / 23.0;
    p = NB_FMA(p, z, 1.0 / 21.0);
    p = NB_FMA(p, z, 1.0 / 19.0);
    p = NB_FMA(p, z, 1.0 / 17.0);
    p = NB_FMA(p, z, 1.0 / 15.0);
    p = NB_FMA(p, z, 1.0 / 13.0);
    p = NB_FMA(p, z, 1.0 / 11.0);
    p = NB_FMA(p, z, 1.0 / 9.0);
    p = NB_FMA(p, z, 1.0 / 7.0);
    p = NB_FMA(p, z, 1.0 / 5.0);
    p = NB_FMA(p, z, 1.0 / 3.0);
    double s2 = NB_ADD(s, s);
    double t = NB_MUL(NB_MUL(s2, z), p);
    double dk = (double)k;
    /* k*ln2_hi is exact (32-bit constant, |k| < 2^11) */
    double lo = NB_FMA(dk, NB_LN2_LO, t);
    return NB_ADD(NB_MUL(dk, NB_LN2_HI), NB_ADD(lo, s2));
}

NB_HD double nb_log10(double x) { return NB_MUL(nb_log(x), NB_INV_LN10); }

/* e^x */
NB_HD_CALL double nb_exp(double x)
{
    if (x > 709.0) return nb_inf();
    if (x < -708.0) return 0.0;
    double dk = NB_RINT(NB_MUL(x, NB_INV_LN2));
    double r = NB_FMA(-dk, NB_LN2_HI, x);
    r = NB_FMA(-dk, NB_LN2_LO, r);             /* |r| <= 0.3466 */
    double p = 1.0 / 87178291200.0;            /* 1/14! */
    p = NB_FMA(p, r, 1.0 / 6227020800.0);
    p = NB_FMA(p, r, 1.0 / 479001600.0);
    p = NB_FMA(p, r, 1.0 / 39916800.0);
    p = NB_FMA(p, r, 1.0 / 3628800.0);
    p = NB_FMA(p, r, 1.0 / 362880.0);
    p = NB_FMA(p, r, 1.0 / 40320.0);
    p = NB_FMA(p, r, 1.0 / 5040.0);
    p = NB_FMA(p, r, 1.0 / 720.0);
    p = NB_FMA(p, r, 1.0 / 120.0);
    p = NB_FMA(p, r, 1.0 / 24.0);
    p = NB_FMA(p, r, 1.0 / 6.0);
    p = NB_FMA(p, r, 0.5);
    p = NB_FMA(p, r, 1.0);
    p = NB_FMA(p, r, 1.0);
    int k = (int)dk;                            /* |k| <= 1023 here */
    return NB_MUL(p, nb_u2d((uint64_t)(k + 1023) << 52));
}

/* 10^x with a compensated product x*ln10 */
NB_HD_CALL double nb_exp10(double x)
{
    double hi = NB_MUL(x, NB_LN10);
    double lo = NB_FMA(x, NB_LN10, -hi);
    lo = NB_FMA(x, NB_LN10_LO, lo);
    double e = nb_exp(hi);
    return NB_FMA(e, lo, e);
}

/* b^e for 0 <= b <= 1, e > 0 (success probability of a block of e*256 bits) */
NB_HD_CALL double nb_pow01(double b, double e)
{
    if (b <= 0.0) return 0.0;
    if (b >= 1.0) return 1.0;
    return nb_exp(NB_MUL(e, nb_log(b)));
}

/* sin / cos Taylor kernels on [0, pi/4] */
NB_HD double nb_sin_k(double y)
{
    double w = NB_MUL(y, y);
    double p = 1.0 / 355687428096000.0;         /* 1/17! */
    p = NB_FMA(p, w, -1.0 / 1307674368000.0);
    p = NB_FMA(p, w, 1.0 / 6227020800.0);
    p = NB_FMA(p, w, -1.0 / 39916800.0);
    p = NB_FMA(p, w, 1.0 / 362880.0);
    p = NB_FMA(p, w, -1.0 / 5040.0);
    p = NB_FMA(p, w, 1.0 / 120.0);
    p = NB_FMA(p, w, -1.0 / 6.0);
    return NB_FMA(NB_MUL(y, w), p, y);
}

NB_HD double nb_cos_k(double y)
{
    double w = NB_MUL(y, y);
    double p = -1.0 / 6402373705728000.0;       /* -1/18! */
    p = NB_FMA(p, w, 1.0 / 20922789888000.0);
    p = NB_FMA(p, w, -1.0 / 87178291200.0);
    p = NB_FMA(p, w, 1.0 / 479001600.0);
    p = NB_FMA(p, w, -1.0 / 3628800.0);
    p = NB_FMA(p, w, 1.0 / 40320.0);
    p = NB_FMA(p, w, -1.0 / 720.0);
    p = NB_FMA(p, w, 1.0 / 24.0);
    p = NB_FMA(p, w, -0.5);
    return NB_FMA(p, w, 1.0);
}

/* cos(2 pi u), u in [0,1): exact octant reduction, then a Taylor kernel */
NB_HD_CALL double nb_cos2pi(double u)
{
    double t = NB_MUL(u, 4.0);                  /* exact */
    double qf = NB_FLOOR(t);
    int q = (int)qf;                            /* 0..3 */
    double f = NB_SUB(t, qf);                   /* exact, [0,1) */
    int swap = f > 0.5;
    double g = swap ? NB_SUB(1.0, f) : f;       /* exact, [0,0.5] */
    double y = NB_MUL(g, NB_PI_2);
    double sn = nb_sin_k(y), cs = nb_cos_k(y);
    double c = swap ? sn : cs;                  /* cos(pi/2 f) */
    double s = swap ? cs : sn;                  /* sin(pi/2 f) */
    switch (q & 3) {
    case 0: return c;
    case 1: return -s;
    case 2: return -c;
    default: return s;
    }
}

/* sin(2 pi u), u in [0,1): the same reduction (clustered placement, population.py:90-96) */
NB_HD_CALL double nb_sin2pi(double u)
{
    double t = NB_MUL(u, 4.0);
    double qf = NB_FLOOR(t);
    int q = (int)qf;
    double f = NB_SUB(t, qf);
    int swap = f > 0.5;
    double g = swap ? NB_SUB(1.0, f) : f;
    double y = NB_MUL(g, NB_PI_2);
    double sn = nb_sin_k(y), cs = nb_cos_k(y);
    double c = swap ? sn : cs;
    double s = swap ? cs : sn;
    switch (q & 3) {
    case 0: return s;
    case 1: return c;
    case 2: return -s;
    default: return -c;
    }
}

/* atan(t), t >= 0 */
NB_HD_CALL double nb_atan_pos(double t)
{
    int inv = t > 1.0;
    if (inv) t = NB_DIV(1.0, t);
    /* two half-angle steps: atan t = 4 atan t2, t2 <= tan(pi/16) */
    double t1 = NB_DIV(t, NB_ADD(1.0, NB_SQRT(NB_FMA(t, t, 1.0))));
    double v = NB_DIV(t1, NB_ADD(1.0, NB_SQRT(NB_FMA(t1, t1, 1.0))));
    double w = NB_MUL(v, v);
    double p = -1.0 / 23.0;
    p = NB_FMA(p, w, 1.0 / 21.0);
    p = NB_FMA(p, w, -1.0 / 19.0);
    p = NB_FMA(p, w, 1.0 / 17.0);
    p = NB_FMA(p, w, -1.0 / 15.0);
    p = NB_FMA(p, w, 1.0 / 13.0);
    p = NB_FMA(p, w, -1.0 / 11.0);
    p = NB_FMA(p, w, 1.0 / 9.0);
    p = NB_FMA(p, w, -1.0 / 7.0);
    p = NB_FMA(p, w, 1.0 / 5.0);
    p = NB_FMA(p, w, -1.0 / 3.0);
    double a = NB_MUL(4.0, NB_FMA(NB_MUL(v, w), p, v));
    return inv ? NB_SUB(NB_PI_2, a) : a;
}

/* acos(x), -1 <= x <= 1 */
NB_HD_CALL double nb_acos(double x)
{
    if (x <= -1.0) return NB_PI;
    if (x >= 1.0) return 0.0;
    double t = NB_SQRT(NB_DIV(NB_SUB(1.0, x), NB_ADD(1.0, x)));
    return NB_MUL(2.0, nb_atan_pos(t));
}

/* regularised incomplete beta I_x(3,4): the CDF scipy.stats.beta(3,4) evaluates in
 * the reference traffic model (system/ue_generator.py:30,74) */
NB_HD_CALL double nb_beta34_cdf(double x)
{
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    double y = NB_SUB(1.0, x);
    double y2 = NB_MUL(y, y);
    /* x^3 (20 y^3 + x (15 y^2 + x (6 y + x))) */
    double in = NB_FMA(6.0, y, x);
    in = NB_FMA(x, in, NB_MUL(15.0, y2));
    in = NB_FMA(x, in, NB_MUL(20.0, NB_MUL(y2, y)));
    return NB_MUL(NB_MUL(NB_MUL(x, x), x), in);
}

#endif /* NBIOT_MATH_H */
