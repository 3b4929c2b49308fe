/*
 * nbiot_rng.h -- the counter-based random stream shared by the CUDA kernels, the
 * host twins exported through the C-ABI, the CPU oracle and the Python shim that is
 * injected into the unmodified reference (rng argument of
 * system/system_creator.py:22, draw sites listed in SURVEY.md App. B).
 *
 * Law (this header DEFINES it; numpy's Generator algorithms are not reproduced):
 *   block(seed, ctr)     = Philox4x32-10, key = (seed.lo, seed.hi),
 *                          counter = (ctr.lo, ctr.hi, stream, 0)
 *   u01                  = 53-bit fraction from words 0,1           -- [0,1)
 *   pair (random(2))     = (u01 from words 0,1 ; u01 from words 2,3) -- ONE counter
 *   bernoulli(p)         = u01 < p                                   -- ONE counter
 *   bounded(n)           = mulhi64(words0..1 as u64, n)              -- ONE counter
 *   normal z             = sqrt(-2 ln u1) * cos(2 pi u2), u1 in (0,1] from words 0,1,
 *                          u2 in [0,1) from words 2,3                -- ONE counter
 * Every draw consumes exactly one counter value of its instance's stream, vector
 * draws (integers(size=k)) consume k consecutive counters, binomial(n>1,p) consumes
 * n counters. ln/cos/sqrt come from nbiot_math.h so host and device agree bit for bit.
 */
#ifndef NBIOT_RNG_H
#define NBIOT_RNG_H

#include <stdint.h>
#include "nbiot_math.h"

#define NBIOT_STREAM_ENV    0u   /* draws made by the simulated cell            */
#define NBIOT_STREAM_POLICY 1u   /* draws made by a benchmark / test policy     */

typedef struct { uint32_t w[4]; } nb_block_t;

NB_HD uint32_t nb_mulhi32(uint32_t a, uint32_t b, uint32_t *lo)
{
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p;
    return (uint32_t)(p >> 32);
}

NB_HD_CALL nb_block_t nb_philox(uint64_t seed, uint64_t ctr, uint32_t stream)
{
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#if defined(__CUDA_ARCH__) && !defined(NB_PHILOX_UNROLL)
#pragma unroll 1   /* rolled: 10x fewer instruction-cache lines for the same work */
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, lo1;
        uint32_t hi0 = nb_mulhi32(0xD2511F53u, c0, &lo0);
        uint32_t hi1 = nb_mulhi32(0xCD9E8D57u, c2, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    nb_block_t b;
    b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
    return b;
}

/* raw Philox with a full 4-word counter / 2-word key: only for known-answer tests */
NB_HD nb_block_t nb_philox_raw(const uint32_t c[4], const uint32_t k[2])
{
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, lo1;
        uint32_t hi0 = nb_mulhi32(0xD2511F53u, c0, &lo0);
        uint32_t hi1 = nb_mulhi32(0xCD9E8D57u, c2, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    nb_block_t b;
    b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
    return b;
}

/* 53-bit fraction in [0,1): exact in double, same construction for both halves */
NB_HD double nb_u53(uint32_t a, uint32_t b)
{
    uint64_t m = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return (double)m * (1.0 / 9007199254740992.0);
}

NB_HD double nb_rng_u01(uint64_t seed, uint64_t ctr, uint32_t stream)
{
    nb_block_t b = nb_philox(seed, ctr, stream);
    return nb_u53(b.w[0], b.w[1]);
}

NB_HD void nb_rng_pair(uint64_t seed, uint64_t ctr, uint32_t stream, double *x, double *y)
{
    nb_block_t b = nb_philox(seed, ctr, stream);
    *x = nb_u53(b.w[0], b.w[1]);
    *y = nb_u53(b.w[2], b.w[3]);
}

NB_HD int nb_rng_bernoulli(uint64_t seed, uint64_t ctr, uint32_t stream, double p)
{
    return nb_rng_u01(seed, ctr, stream) < p ? 1 : 0;
}

/* uniform integer in [0, n), n >= 1 */
NB_HD uint64_t nb_rng_bounded(uint64_t seed, uint64_t ctr, uint32_t stream, uint64_t n)
{
    nb_block_t b = nb_philox(seed, ctr, stream);
    uint64_t x = ((uint64_t)b.w[1] << 32) | (uint64_t)b.w[0];
#ifdef __CUDA_ARCH__
    return __umul64hi(x, n);
#else
    return (uint64_t)(((unsigned __int128)x * (unsigned __int128)n) >> 64);
#endif
}

/* standard normal deviate */
NB_HD_CALL double nb_rng_normal(uint64_t seed, uint64_t ctr, uint32_t stream)
{
    nb_block_t b = nb_philox(seed, ctr, stream);
    /* u1 in (0,1]: (m+1)/2^53 keeps ln finite; u2 in [0,1) */
    uint64_t m = ((uint64_t)(b.w[0] >> 5) << 26) | (uint64_t)(b.w[1] >> 6);
    double u1 = (double)(m + 1u) * (1.0 / 9007199254740992.0);
    double u2 = nb_u53(b.w[2], b.w[3]);
    double r = NB_SQRT(NB_MUL(-2.0, nb_log(u1)));
    return NB_MUL(r, nb_cos2pi(u2));
}

#endif /* NBIOT_RNG_H */
