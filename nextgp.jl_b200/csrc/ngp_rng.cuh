// ngp_rng.cuh — counter-based variate stream of libngp.
//
// Replaces the reference's use of Julia's default RNG + Distributions.jl
// (functions.jl:494 Normal, :174/:216 rand(), :510/:524 Chisq, :532 Beta).
// Philox4x32-10; every variate is addressed by
//   key  = seed
//   ctr0 = index (marker / region), ctr1 = purpose | attempt<<8 | set<<20 | comp<<26,
//   ctr2 = iteration (1-based),     ctr3 = chain id
// so the order in which draws are consumed is irrelevant by construction.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define NGP_HD __host__ __device__ __forceinline__
#else
#define NGP_HD inline
#endif

namespace ngp {

enum Purpose : uint32_t {
    P_CHI2_E = 0, P_Z_MU = 1, P_U = 2, P_Z = 3, P_CHI2_B = 4, P_PI_A = 5, P_PI_B = 6, P_IW = 7,
    P_U_ANNOT = 8, P_G_ANNOT = 9       // BayesRCpi: the Categorical draw of the annotation, the gammas of sampleProb
};

struct Stream {
    uint32_t key0, key1;  // seed
    uint32_t chain;
    uint32_t iter;
    uint32_t set_id;
};

NGP_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

NGP_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

NGP_HD void stream_words(const Stream& s, uint32_t purpose, uint32_t idx, uint32_t attempt, uint32_t comp, uint32_t w[4])
{
    philox4x32_10(idx, purpose | (attempt << 8) | (s.set_id << 20) | (comp << 26), s.iter, s.chain, s.key0, s.key1, w);
}

// 52-bit uniform strictly inside (0,1): (k + 1/2) * 2^-52, k in [0, 2^52) — every value exact in fp64
NGP_HD double u53(uint32_t hi, uint32_t lo)
{
    return (((double)(hi >> 6)) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 4503599627370496.0);
}

NGP_HD double stream_uniform(const Stream& s, uint32_t purpose, uint32_t idx, uint32_t attempt = 0, uint32_t comp = 0)
{
    uint32_t w[4];
    stream_words(s, purpose, idx, attempt, comp, w);
    return u53(w[0], w[1]);
}

// Box-Muller
NGP_HD double stream_normal(const Stream& s, uint32_t purpose, uint32_t idx, uint32_t attempt = 0, uint32_t comp = 0)
{
    uint32_t w[4];
    stream_words(s, purpose, idx, attempt, comp, w);
    const double u1 = u53(w[0], w[1]);
    const double u2 = u53(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

// Marsaglia & Tsang (2000), shape >= 1, scale 1.  attempt 2t -> normal, 2t+1 -> uniform.
NGP_HD double stream_gamma(const Stream& s, uint32_t purpose, uint32_t idx, uint32_t comp, double shape)
{
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t < 2048; ++t) {
        const double x = stream_normal(s, purpose, idx, 2 * t, comp);
        const double u = stream_uniform(s, purpose, idx, 2 * t + 1, comp);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
    return d;
}

NGP_HD double stream_chisq(const Stream& s, uint32_t purpose, uint32_t idx, uint32_t comp, double df)
{
    return 2.0 * stream_gamma(s, purpose, idx, comp, 0.5 * df);
}

NGP_HD double stream_beta(const Stream& s, double a, double b)
{
    const double ga = stream_gamma(s, P_PI_A, 0, 0, a);
    const double gb = stream_gamma(s, P_PI_B, 0, 0, b);
    return ga / (ga + gb);
}

}  // namespace ngp
