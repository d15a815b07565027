/*
 * ngp_oracle.c — CPU ORACLE for the NextGP.jl marker-effect Gibbs sweep.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: it
 * may be imported / linked / executed only from tests/, from
 * __graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under nextgp.jl_b200/ links to or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 * (/root/reference/test/runtests.jl:4-7 is an empty @testset) and Julia is not
 * installed, so this restatement cannot be checked against reference outputs.
 * It is pinned only by (i) the hyper-parameter transcripts in the reference
 * docs (docs/src/BWGR/BWGR.md:52-55 etc., see tests/test_oracle.py), (ii) an
 * independent numpy restatement (oracle/restate_numpy.py) and (iii) the
 * Philox4x32-10 known-answer vectors of Random123.
 *
 * What it restates (all file:line under /root/reference/src):
 *   - iteration order                    samplers.jl:32-53
 *   - residual variance draw             functions.jl:523-525
 *   - intercept (single-column X) draw   functions.jl:39-47
 *   - BayesPR sweep                      functions.jl:118-137
 *   - BayesB  sweep                      functions.jl:157-195
 *   - BayesC  sweep                      functions.jl:197-236
 *   - multi-breed (Tuple) BayesPR        functions.jl:140-154, 513-516
 *   - sampleBeta / sampleVarBetaPR / samplePi   functions.jl:493-495, 509-511, 531-533
 *   - column norms mpm                   mme.jl:305-307
 *   - centring                           prepMatVec.jl:129
 * The memory behaviour deliberately mimics the reference: dense fp64
 * column-major centred genotypes, add-back axpy, dot, subtract axpy, and for
 * BayesB/BayesC a SECOND dot on the (optional) Mp copy for included loci.
 *
 * Third-party arithmetic that is not in /root/reference: Distributions.jl
 * (compat 0.25.58, Project.toml:29) Normal / Chisq / Beta / MvNormal /
 * InverseWishart samplers and Julia's Xoshiro256++ default RNG.  Julia's
 * (unseeded) bit-stream cannot be reproduced; parity is therefore defined at
 * the VARIATE level: every draw is an explicit input ("replay"), or comes from
 * the counter-based Philox4x32-10 stream specified in DESIGN.md §RNG, whose
 * transforms are restated here independently of the CUDA implementation:
 *   Normal   : Box-Muller on two 52-bit uniforms
 *   Chisq(v) : 2*Gamma(v/2), Gamma by Marsaglia & Tsang (2000)
 *   Beta(a,b): Ga/(Ga+Gb)
 *   MvNormal : mean + chol(C)_lower * z
 *   InvWishart(df,S): Bartlett factor of Wishart(df, S^-1), inverted
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011; Random123 reference).      */
/* ------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n1 = lo1;
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    const uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void ngo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* Stream layout:
 *   key  = (seed lo32, seed hi32)
 *   ctr0 = index (marker, region, ...)
 *   ctr1 = purpose | attempt<<8 | set_id<<20 | component<<26
 *   ctr2 = iteration (1-based)
 *   ctr3 = chain id                                                      */
enum {
    NGO_P_CHI2_E = 0, NGO_P_Z_MU = 1, NGO_P_U = 2, NGO_P_Z = 3,
    NGO_P_CHI2_B = 4, NGO_P_PI_A = 5, NGO_P_PI_B = 6, NGO_P_IW = 7,
    NGO_P_U_ANNOT = 8, NGO_P_G_ANNOT = 9      /* BayesRCpi: the Categorical draw of the annotation, the gammas of sampleProb */
};

typedef struct { uint64_t seed; uint32_t chain; uint32_t iter; uint32_t set_id; } ngo_stream;

static inline void stream_words(const ngo_stream* s, uint32_t purpose, uint32_t idx,
                                uint32_t attempt, uint32_t comp, uint32_t w[4])
{
    uint32_t ctr[4], key[2];
    ctr[0] = idx;
    ctr[1] = purpose | (attempt << 8) | (s->set_id << 20) | (comp << 26);
    ctr[2] = s->iter;
    ctr[3] = s->chain;
    key[0] = (uint32_t)(s->seed & 0xffffffffu);
    key[1] = (uint32_t)(s->seed >> 32);
    ngo_philox4x32_10(ctr, key, w);
}

/* 52-bit uniform strictly inside (0,1): (k + 1/2) * 2^-52, k in [0, 2^52) — every value exact in fp64 */
static inline double u53(uint32_t hi, uint32_t lo)
{
    const double two26 = 67108864.0;
    return (((double)(hi >> 6)) * two26 + (double)(lo >> 6) + 0.5) * (1.0 / 4503599627370496.0);
}

static double stream_uniform(const ngo_stream* s, uint32_t purpose, uint32_t idx, uint32_t attempt, uint32_t comp)
{
    uint32_t w[4];
    stream_words(s, purpose, idx, attempt, comp, w);
    return u53(w[0], w[1]);
}

static double stream_normal(const ngo_stream* s, uint32_t purpose, uint32_t idx, uint32_t attempt, uint32_t comp)
{
    uint32_t w[4];
    stream_words(s, purpose, idx, attempt, comp, w);
    const double u1 = u53(w[0], w[1]);
    const double u2 = u53(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

/* Marsaglia-Tsang Gamma(shape>=1, scale 1); attempt 2t -> normal, 2t+1 -> uniform */
static double stream_gamma(const ngo_stream* s, uint32_t purpose, uint32_t idx, uint32_t comp, double shape)
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
    return d; /* unreachable in practice */
}

static double stream_chisq(const ngo_stream* s, uint32_t purpose, uint32_t idx, uint32_t comp, double df)
{
    return 2.0 * stream_gamma(s, purpose, idx, comp, 0.5 * df);
}

/* exported scalar generators (tests compare the CUDA stream against these) */
double ngo_stream_uniform(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t set_id, uint32_t purpose, uint32_t idx)
{ ngo_stream s = {seed, chain, iter, set_id}; return stream_uniform(&s, purpose, idx, 0, 0); }
double ngo_stream_normal(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t set_id, uint32_t purpose, uint32_t idx, uint32_t comp)
{ ngo_stream s = {seed, chain, iter, set_id}; return stream_normal(&s, purpose, idx, 0, comp); }
double ngo_stream_chisq(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t set_id, uint32_t purpose, uint32_t idx, uint32_t comp, double df)
{ ngo_stream s = {seed, chain, iter, set_id}; return stream_chisq(&s, purpose, idx, comp, df); }
double ngo_stream_beta(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t set_id, double a, double b)
{
    ngo_stream s = {seed, chain, iter, set_id};
    const double ga = stream_gamma(&s, NGO_P_PI_A, 0, 0, a);
    const double gb = stream_gamma(&s, NGO_P_PI_B, 0, 0, b);
    return ga / (ga + gb);
}

/* ------------------------------------------------------------------------- */
/* level-1 kernels with the reference's memory behaviour (OpenBLAS daxpy/ddot) */
/* ------------------------------------------------------------------------- */
static int g_threads = 1;
void ngo_set_threads(int t)
{
    g_threads = t < 1 ? 1 : t;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}
int ngo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double ddot(int64_t n, const double* x, const double* y)
{
    double s = 0.0;
    if (g_threads > 1 && n >= 8192) {
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    } else {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int64_t i = 0;
        for (; i + 3 < n; i += 4) {
            s0 += x[i] * y[i]; s1 += x[i + 1] * y[i + 1];
            s2 += x[i + 2] * y[i + 2]; s3 += x[i + 3] * y[i + 3];
        }
        for (; i < n; ++i) s0 += x[i] * y[i];
        s = (s0 + s1) + (s2 + s3);
    }
    return s;
}

static void daxpy(int64_t n, double a, const double* x, double* y)
{
    if (a == 0.0) return; /* OpenBLAS daxpy returns early when alpha == 0 */
    if (g_threads > 1 && n >= 8192) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
    } else {
        for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
    }
}

/* ------------------------------------------------------------------------- */
/* data preparation: prepMatVec.jl:129 (centre), mme.jl:305-307 (mpm)          */
/* ------------------------------------------------------------------------- */
/* codes: int8 column-major n x p (values 0/1/2) -> centred fp64 X, means, mpm */
void ngo_center_codes(int64_t n, int64_t p, const int8_t* codes, double* X, double* mean, double* mpm)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < p; ++j) {
        const int8_t* g = codes + j * n;
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += (double)g[i];
        const double m = s / (double)n; /* mean(thisM,dims=1) */
        double* x = X + j * n;
        for (int64_t i = 0; i < n; ++i) x[i] = (double)g[i] - m;
        if (mean) mean[j] = m;
        if (mpm) {
            double d = 0.0;
            for (int64_t i = 0; i < n; ++i) d += x[i] * x[i]; /* dot(c,c) */
            mpm[j] = d;
        }
    }
}

void ngo_center_f64(int64_t n, int64_t p, double* X, double* mean, double* mpm)
{
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < p; ++j) {
        double* x = X + j * n;
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += x[i];
        const double m = s / (double)n;
        for (int64_t i = 0; i < n; ++i) x[i] -= m;
        if (mean) mean[j] = m;
        if (mpm) {
            double d = 0.0;
            for (int64_t i = 0; i < n; ++i) d += x[i] * x[i];
            mpm[j] = d;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* model structs (mirrored by ctypes in oracle/oracle.py)                      */
/* ------------------------------------------------------------------------- */
enum { NGO_BAYESPR = 0, NGO_BAYESB = 1, NGO_BAYESC = 2 };

typedef struct {
    int64_t n, p;
    const double* X;          /* n x p column-major, centred                      */
    const double* Mp;         /* second copy (mme.jl:308) or NULL -> reuse X      */
    const double* mpm;        /* p                                                */
    const double* lhs0;       /* p or NULL (mme.jl:314-322)                       */
    const double* rhs0;       /* p or NULL                                        */
    int32_t method;           /* NGO_BAYES*                                       */
    int32_t est_pi;
    int64_t n_regions;        /* BayesPR only                                     */
    const int64_t* region_off;/* n_regions+1, 0-based half-open                   */
    double df, scale;         /* mme.jl:492-506                                   */
    int32_t set_id;
    int32_t pad_;
} ngo_set;

typedef struct {
    double* beta;             /* p                                                */
    int64_t* delta;           /* p (Int64, mme.jl:444)                            */
    double* varBeta;          /* PR: n_regions ; B: p ; C: 1                      */
    double piHat[2];          /* [not fitted, fitted] mme.jl:359,371              */
    double logPi[2];
} ngo_set_state;

/* variates of ONE iteration for ONE marker set.  u,z (and chi2_b for PR/B) are
 * always filled before the sweep (by ngo_fill_marker_variates or by the caller
 * when replaying); chi2_b[0] for BayesC and beta_pi depend on the chain state
 * and are generated lazily unless replay != 0.  Generated values are stored,
 * so after the call the struct is the replay log of the iteration.           */
typedef struct {
    int32_t replay;
    int32_t pad_;
    uint64_t seed;
    uint32_t chain, iter;
    double* u;                /* p  (B, C)                                        */
    double* z;                /* p                                                */
    double* chi2_b;           /* PR: n_regions ; B: p ; C: 1                      */
    double* beta_pi;          /* 1                                                */
} ngo_variates;

void ngo_fill_marker_variates(const ngo_set* S, ngo_variates* V)
{
    ngo_stream s = {V->seed, V->chain, V->iter, (uint32_t)S->set_id};
    const int64_t p = S->p;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < p; ++j) {
        if (S->method != NGO_BAYESPR && V->u) V->u[j] = stream_uniform(&s, NGO_P_U, (uint32_t)j, 0, 0);
        V->z[j] = stream_normal(&s, NGO_P_Z, (uint32_t)j, 0, 0);
        if (S->method == NGO_BAYESB) V->chi2_b[j] = stream_chisq(&s, NGO_P_CHI2_B, (uint32_t)j, 0, S->df + 1.0);
    }
    if (S->method == NGO_BAYESPR) {
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < S->n_regions; ++r) {
            const double k = (double)(S->region_off[r + 1] - S->region_off[r]);
            V->chi2_b[r] = stream_chisq(&s, NGO_P_CHI2_B, (uint32_t)r, 0, S->df + k);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* residual variance: functions.jl:523-525, called at samplers.jl:32-35        */
/* ------------------------------------------------------------------------- */
double ngo_sample_varE(int64_t n, const double* e, double df_e, double scale_e,
                       int replay, uint64_t seed, uint32_t chain, uint32_t iter, double* chi2_e)
{
    if (!replay) {
        ngo_stream s = {seed, chain, iter, 0};
        *chi2_e = stream_chisq(&s, NGO_P_CHI2_E, 0, 0, df_e + (double)n);
    }
    return (df_e * scale_e + ddot(n, e, e)) / (*chi2_e);
}

/* ------------------------------------------------------------------------- */
/* intercept = single-column fixed effect of ones: functions.jl:39-47          */
/*   ycorr += 1*b ; rhs = (1'ycorr)*iVarE + rhs0 ; lhs = n*iVarE + lhs0 ;      */
/*   b ~ N(lhs\rhs, sqrt(inv(lhs))) ; ycorr -= 1*b                             */
/* ------------------------------------------------------------------------- */
double ngo_sample_intercept(int64_t n, double* e, double mu_old, double varE, double lhs0, double rhs0,
                            int replay, uint64_t seed, uint32_t chain, uint32_t iter, double* z_mu)
{
    const double iVarE = 1.0 / varE;
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) { e[i] += mu_old; sum += e[i]; }
    const double rhs = sum * iVarE + rhs0;
    const double lhs = (double)n * iVarE + lhs0;
    const double meanMu = rhs / lhs;
    if (!replay) {
        ngo_stream s = {seed, chain, iter, 0};
        *z_mu = stream_normal(&s, NGO_P_Z_MU, 0, 0, 0);
    }
    const double mu = meanMu + sqrt(1.0 / lhs) * (*z_mu);
    for (int64_t i = 0; i < n; ++i) e[i] -= mu;
    return mu;
}

/* ------------------------------------------------------------------------- */
/* weighted residuals, E.str == "D" (mme.jl:70-73: iVarStr = inv.(D), w_i = 1/d_ii) */
/*   residual variance  functions.jl:526-528 (called at samplers.jl:32-33):     */
/*     (df*scale + sum(w .* e.^2)) / chi2(df + n)                               */
/*   intercept          functions.jl:39-47 with xpx = 1'W1, Xp = (1 .* w)'       */
/*     (mme.jl:133-136)                                                         */
/*   the marker sweeps take Mp = (x_j .* w)' and mpm = sum(x_j .* w .* x_j)     */
/*   (mme.jl:299-303) through ngo_set.Mp / ngo_set.mpm; the add-back and the    */
/*   BayesB/C inclusion dot (functions.jl:168, :208) stay unweighted.           */
/* ------------------------------------------------------------------------- */
double ngo_sample_varE_w(int64_t n, const double* e, const double* w, double df_e, double scale_e,
                         int replay, uint64_t seed, uint32_t chain, uint32_t iter, double* chi2_e)
{
    if (!replay) {
        ngo_stream s = {seed, chain, iter, 0};
        *chi2_e = stream_chisq(&s, NGO_P_CHI2_E, 0, 0, df_e + (double)n);
    }
    double ss = 0.0;
    for (int64_t i = 0; i < n; ++i) ss += w[i] * (e[i] * e[i]);
    return (df_e * scale_e + ss) / (*chi2_e);
}

double ngo_sample_intercept_w(int64_t n, double* e, const double* w, double mu_old, double varE, double lhs0, double rhs0,
                              int replay, uint64_t seed, uint32_t chain, uint32_t iter, double* z_mu)
{
    const double iVarE = 1.0 / varE;
    double sum = 0.0, xpx = 0.0;
    for (int64_t i = 0; i < n; ++i) { e[i] += mu_old; sum += w[i] * e[i]; xpx += w[i]; }
    const double rhs = sum * iVarE + rhs0;
    const double lhs = xpx * iVarE + lhs0;
    const double meanMu = rhs / lhs;
    if (!replay) {
        ngo_stream s = {seed, chain, iter, 0};
        *z_mu = stream_normal(&s, NGO_P_Z_MU, 0, 0, 0);
    }
    const double mu = meanMu + sqrt(1.0 / lhs) * (*z_mu);
    for (int64_t i = 0; i < n; ++i) e[i] -= mu;
    return mu;
}

/* sampleVarBetaPR: functions.jl:509-511 */
static double sample_var_beta(double scale, double df, const double* b, int64_t k, double regionSizeOrNLoci, double chi2)
{
    (void)regionSizeOrNLoci; /* enters only through the df of chi2 */
    return (scale * df + ddot(k, b, b)) / chi2;
}

/* ------------------------------------------------------------------------- */
/* BayesPR sweep: functions.jl:118-137                                         */
/* ------------------------------------------------------------------------- */
static void sweep_PR(const ngo_set* S, ngo_set_state* T, double* e, double varE, ngo_variates* V)
{
    const int64_t n = S->n;
    const double* Mp = S->Mp ? S->Mp : S->X;
    const double iVarE = 1.0 / varE;
    for (int64_t r = 0; r < S->n_regions; ++r) {
        const int64_t j0 = S->region_off[r], j1 = S->region_off[r + 1];
        const double iVarBeta = 1.0 / T->varBeta[r];
        for (int64_t j = j0; j < j1; ++j) {
            const double* x = S->X + j * n;
            daxpy(n, T->beta[j], x, e);                                   /* :128 */
            const double rhs = ddot(n, Mp + j * n, e) * iVarE + (S->rhs0 ? S->rhs0[j] : 0.0);      /* :129 */
            const double lhs = S->mpm[j] * iVarE + (S->lhs0 ? S->lhs0[j] : 0.0) + iVarBeta;        /* :130 */
            const double meanBeta = rhs / lhs;                            /* :131 */
            T->beta[j] = meanBeta + sqrt(1.0 / lhs) * V->z[j];            /* :132, :493-495 */
            daxpy(n, -1.0 * T->beta[j], x, e);                            /* :133 */
        }
        T->varBeta[r] = sample_var_beta(S->scale, S->df, T->beta + j0, j1 - j0, (double)(j1 - j0), V->chi2_b[r]); /* :135 */
    }
}

/* ------------------------------------------------------------------------- */
/* BayesB sweep: functions.jl:157-195                                          */
/* ------------------------------------------------------------------------- */
static void sweep_B(const ngo_set* S, ngo_set_state* T, double* e, double varE, ngo_variates* V)
{
    const int64_t n = S->n, p = S->p;
    const double* Mp = S->Mp ? S->Mp : S->X;
    int64_t nLoci = 0;
    for (int64_t j = 0; j < p; ++j) {
        const double iVarE = 1.0 / varE;
        const double iVarBeta = 1.0 / T->varBeta[j];                      /* Inf when varBeta == 0.0 */
        const double* x = S->X + j * n;
        daxpy(n, T->beta[j], x, e);                                       /* :167 */
        const double rrr = ddot(n, x, e);                                 /* :168 */
        const double v0 = S->mpm[j] * varE;                               /* :169 */
        const double v1 = (S->mpm[j] * S->mpm[j]) * T->varBeta[j] + v0;   /* :170 */
        const double logDelta0 = -0.5 * (log(v0) + (rrr * rrr) / v0) + T->logPi[0];
        const double logDelta1 = -0.5 * (log(v1) + (rrr * rrr) / v1) + T->logPi[1];
        const double probDelta1 = 1.0 / (1.0 + exp(logDelta0 - logDelta1));
        if (V->u[j] < probDelta1) {                                       /* :174 strict < */
            T->delta[j] = 1;
            nLoci += 1;
            const double rhs = ddot(n, Mp + j * n, e) * iVarE + (S->rhs0 ? S->rhs0[j] : 0.0);      /* :177 */
            const double lhs = S->mpm[j] * iVarE + (S->lhs0 ? S->lhs0[j] : 0.0) + iVarBeta;        /* :178 */
            const double meanBeta = rhs / lhs;                            /* lhs\rhs ; rhs/Inf == 0 */
            T->beta[j] = meanBeta + sqrt(1.0 / lhs) * V->z[j];            /* one normal consumed */
            daxpy(n, -1.0 * T->beta[j], x, e);                            /* :181 */
            T->varBeta[j] = sample_var_beta(S->scale, S->df, T->beta + j, 1, 1.0, V->chi2_b[j]);   /* :182 */
        } else {
            T->beta[j] = 0.0;
            T->delta[j] = 0;
            T->varBeta[j] = 0.0;                                          /* :186 */
        }
    }
    if (S->est_pi) {                                                      /* :189-193 */
        if (!V->replay) {
            *V->beta_pi = ngo_stream_beta(V->seed, V->chain, V->iter, (uint32_t)S->set_id,
                                          (double)nLoci + 1.0, (double)(p - nLoci) + 1.0);
        }
        const double piIn = *V->beta_pi;
        T->piHat[0] = 1.0 - piIn; T->piHat[1] = piIn;
        T->logPi[0] = log(1.0 - piIn); T->logPi[1] = log(piIn);
    }
}

/* ------------------------------------------------------------------------- */
/* BayesC sweep: functions.jl:197-236                                          */
/* ------------------------------------------------------------------------- */
static void sweep_C(const ngo_set* S, ngo_set_state* T, double* e, double varE, ngo_variates* V)
{
    const int64_t n = S->n, p = S->p;
    const double* Mp = S->Mp ? S->Mp : S->X;
    int64_t nLoci = 0;
    const double iVarE = 1.0 / varE;
    const double iVarBeta = 1.0 / T->varBeta[0];
    for (int64_t j = 0; j < p; ++j) {
        const double* x = S->X + j * n;
        daxpy(n, T->beta[j], x, e);                                       /* :207 */
        const double rrr = ddot(n, x, e);                                 /* :208 */
        const double v0 = S->mpm[j] * varE;
        const double v1 = (S->mpm[j] * S->mpm[j]) * T->varBeta[0] + v0;
        const double logDelta0 = -0.5 * (log(v0) + (rrr * rrr) / v0) + T->logPi[0];
        const double logDelta1 = -0.5 * (log(v1) + (rrr * rrr) / v1) + T->logPi[1];
        const double probDelta1 = 1.0 / (1.0 + exp(logDelta0 - logDelta1));
        if (V->u[j] < probDelta1) {                                       /* :216 */
            T->delta[j] = 1;
            nLoci += 1;
            const double rhs = ddot(n, Mp + j * n, e) * iVarE;            /* :219, rhs0 NOT added */
            const double lhs = S->mpm[j] * iVarE + (S->lhs0 ? S->lhs0[j] : 0.0) + iVarBeta;        /* :220 */
            const double meanBeta = rhs / lhs;
            T->beta[j] = meanBeta + sqrt(1.0 / lhs) * V->z[j];
            daxpy(n, -1.0 * T->beta[j], x, e);                            /* :223 */
        } else {
            T->beta[j] = 0.0;
            T->delta[j] = 0;
        }
    }
    if (!V->replay) {
        ngo_stream s = {V->seed, V->chain, V->iter, (uint32_t)S->set_id};
        V->chi2_b[0] = stream_chisq(&s, NGO_P_CHI2_B, 0, 0, S->df + (double)nLoci);
    }
    T->varBeta[0] = sample_var_beta(S->scale, S->df, T->beta, p, (double)nLoci, V->chi2_b[0]);     /* :230 */
    if (S->est_pi) {                                                      /* :231-235 */
        if (!V->replay) {
            *V->beta_pi = ngo_stream_beta(V->seed, V->chain, V->iter, (uint32_t)S->set_id,
                                          (double)nLoci + 1.0, (double)(p - nLoci) + 1.0);
        }
        const double piIn = *V->beta_pi;
        T->piHat[0] = 1.0 - piIn; T->piHat[1] = piIn;
        T->logPi[0] = log(1.0 - piIn); T->logPi[1] = log(piIn);
    }
}

/* M[mSet].funct(mSet,M,beta,delta,ycorr,varE,varBeta) — samplers.jl:52 */
int ngo_sweep(const ngo_set* S, ngo_set_state* T, double* e, double varE, ngo_variates* V)
{
    switch (S->method) {
    case NGO_BAYESPR: sweep_PR(S, T, e, varE, V); return 0;
    case NGO_BAYESB:  sweep_B(S, T, e, varE, V);  return 0;
    case NGO_BAYESC:  sweep_C(S, T, e, varE, V);  return 0;
    default: return -1;
    }
}

/* ------------------------------------------------------------------------- */
/* whole-chain driver for ONE marker set + intercept (samplers.jl:29-53),      */
/* used for CPU-baseline timing and for native-stream parity runs.            */
/* trace_* (optional, n_iter rows) record the state after every iteration.    */
/* ------------------------------------------------------------------------- */
typedef struct {
    double df_e, scale_e;     /* mme.jl:87-94 */
    int32_t has_intercept, pad_;
    double mu_lhs0, mu_rhs0;
} ngo_fixed;

int ngo_run(const ngo_set* S, ngo_set_state* T, const ngo_fixed* F, double* e, double* mu, double* varE_io,
            uint64_t seed, uint32_t chain, uint32_t iter0, int32_t n_iter,
            double* scratch_u, double* scratch_z, double* scratch_chi2, /* sized like ngo_variates arrays */
            double* trace_beta, double* trace_varE, double* trace_mu, double* trace_varBeta, double* trace_pi,
            int64_t* trace_delta)
{
    const int64_t nvar = S->method == NGO_BAYESPR ? S->n_regions : (S->method == NGO_BAYESB ? S->p : 1);
    for (int32_t it = 0; it < n_iter; ++it) {
        const uint32_t iter = iter0 + (uint32_t)it;
        double chi2_e = 0.0, z_mu = 0.0, beta_pi = 0.0;
        const double varE = ngo_sample_varE(S->n, e, F->df_e, F->scale_e, 0, seed, chain, iter, &chi2_e);
        if (F->has_intercept) *mu = ngo_sample_intercept(S->n, e, *mu, varE, F->mu_lhs0, F->mu_rhs0, 0, seed, chain, iter, &z_mu);
        ngo_variates V;
        V.replay = 0; V.pad_ = 0; V.seed = seed; V.chain = chain; V.iter = iter;
        V.u = scratch_u; V.z = scratch_z; V.chi2_b = scratch_chi2; V.beta_pi = &beta_pi;
        ngo_fill_marker_variates(S, &V);
        if (ngo_sweep(S, T, e, varE, &V)) return -1;
        *varE_io = varE;
        if (trace_beta) memcpy(trace_beta + (int64_t)it * S->p, T->beta, sizeof(double) * (size_t)S->p);
        if (trace_delta) memcpy(trace_delta + (int64_t)it * S->p, T->delta, sizeof(int64_t) * (size_t)S->p);
        if (trace_varE) trace_varE[it] = varE;
        if (trace_mu) trace_mu[it] = *mu;
        if (trace_varBeta) memcpy(trace_varBeta + (int64_t)it * nvar, T->varBeta, sizeof(double) * (size_t)nvar);
        if (trace_pi) { trace_pi[2 * it] = T->piHat[0]; trace_pi[2 * it + 1] = T->piHat[1]; }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* multi-breed (Tuple) BayesPR: functions.jl:140-154, 513-516; mme.jl:448-467  */
/* k marker sets share loci; per locus the k effects are drawn jointly.       */
/*   Xk[b] : n x p centred matrix of breed b ; beta: k x p (row b = breed b)   */
/*   varBeta: n_regions x k x k (row-major per region)                         */
/*   z: p x k normals ; iw_z / iw_chi2: Bartlett variates per region           */
/* ------------------------------------------------------------------------- */
static int chol_lower(int k, const double* A, double* L)
{
    memset(L, 0, sizeof(double) * (size_t)(k * k));
    for (int i = 0; i < k; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[i * k + j];
            for (int t = 0; t < j; ++t) s -= L[i * k + t] * L[j * k + t];
            if (i == j) { if (s <= 0.0) return -1; L[i * k + i] = sqrt(s); }
            else L[i * k + j] = s / L[j * k + j];
        }
    return 0;
}

static int inv_spd(int k, const double* A, double* Ainv)
{
    double L[64], Li[64];
    if (k > 8 || chol_lower(k, A, L)) return -1;
    memset(Li, 0, sizeof(Li));
    for (int c = 0; c < k; ++c) {              /* Li = L^-1 (lower) */
        Li[c * k + c] = 1.0 / L[c * k + c];
        for (int i = c + 1; i < k; ++i) {
            double s = 0.0;
            for (int t = c; t < i; ++t) s -= L[i * k + t] * Li[t * k + c];
            Li[i * k + c] = s / L[i * k + i];
        }
    }
    for (int i = 0; i < k; ++i)                /* Ainv = Li' Li */
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int t = (i > j ? i : j); t < k; ++t) s += Li[t * k + i] * Li[t * k + j];
            Ainv[i * k + j] = s;
        }
    return 0;
}

typedef struct {
    int64_t n, p;
    int32_t k, set_id;
    const double* const* Xk;  /* k pointers, each n x p centred column-major      */
    int64_t n_regions;
    const int64_t* region_off;
    double df;                /* 3 + k, mme.jl:493                                */
    const double* scale;      /* k x k, v*(df-k-1), mme.jl:501                    */
} ngo_mb_set;

typedef struct {
    int32_t replay, pad_;
    uint64_t seed;
    uint32_t chain, iter;
    double* z;                /* p x k                                            */
    double* iw_chi2;          /* n_regions x k   (Bartlett diagonal, df-i)        */
    double* iw_z;             /* n_regions x k x k (strict lower used)            */
} ngo_mb_variates;

void ngo_mb_fill_variates(const ngo_mb_set* S, ngo_mb_variates* V)
{
    ngo_stream s = {V->seed, V->chain, V->iter, (uint32_t)S->set_id};
    const int k = S->k;
    for (int64_t j = 0; j < S->p; ++j)
        for (int c = 0; c < k; ++c) V->z[j * k + c] = stream_normal(&s, NGO_P_Z, (uint32_t)j, 0, (uint32_t)c);
    for (int64_t r = 0; r < S->n_regions; ++r) {
        const double dfr = S->df + (double)(S->region_off[r + 1] - S->region_off[r]);
        for (int i = 0; i < k; ++i) {
            V->iw_chi2[r * k + i] = stream_chisq(&s, NGO_P_IW, (uint32_t)r, (uint32_t)(i * k + i), dfr - (double)i);
            for (int j = 0; j < k; ++j)
                V->iw_z[(r * k + i) * k + j] = (j < i) ? stream_normal(&s, NGO_P_IW, (uint32_t)r, 0, (uint32_t)(i * k + j)) : 0.0;
        }
    }
}

/* Sigma ~ InvWishart(df, Psi): W = (L A)(L A)' ~ Wishart(df, Psi^-1) with L = chol(Psi^-1),
 * A lower Bartlett factor (A_ii = sqrt(chi2(df-i)), A_ij = N(0,1), i>j); Sigma = W^-1.    */
static int inv_wishart_bartlett(int k, const double* Psi, const double* chi2, const double* zl, double* Sigma)
{
    double Pinv[64], L[64], A[64], LA[64], W[64];
    if (inv_spd(k, Psi, Pinv) || chol_lower(k, Pinv, L)) return -1;
    memset(A, 0, sizeof(A));
    for (int i = 0; i < k; ++i) {
        A[i * k + i] = sqrt(chi2[i]);
        for (int j = 0; j < i; ++j) A[i * k + j] = zl[i * k + j];
    }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int t = 0; t < k; ++t) s += L[i * k + t] * A[t * k + j];
            LA[i * k + j] = s;
        }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int t = 0; t < k; ++t) s += LA[i * k + t] * LA[j * k + t];
            W[i * k + j] = s;
        }
    return inv_spd(k, W, Sigma);
}

int ngo_mb_sweep(const ngo_mb_set* S, double* beta /* k x p */, double* varBeta /* R x k x k */,
                 double* e, double varE, ngo_mb_variates* V)
{
    const int64_t n = S->n, p = S->p;
    const int k = S->k;
    if (k > 8) return -1;
    for (int64_t r = 0; r < S->n_regions; ++r) {
        const int64_t j0 = S->region_off[r], j1 = S->region_off[r + 1];
        double invB[64];
        if (inv_spd(k, varBeta + r * k * k, invB)) return -2;             /* :143 */
        for (int64_t j = j0; j < j1; ++j) {
            double RHS[8], MtM[64], LHS[64], C[64], Lc[64], mean[8];
            for (int b = 0; b < k; ++b) daxpy(n, beta[b * p + j], S->Xk[b] + j * n, e);           /* :145 */
            for (int b = 0; b < k; ++b) RHS[b] = ddot(n, S->Xk[b] + j * n, e) / varE;            /* :146 */
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) {
                    MtM[a * k + b] = ddot(n, S->Xk[a] + j * n, S->Xk[b] + j * n);                /* mpm: mme.jl:464 */
                    LHS[a * k + b] = MtM[a * k + b] / varE + invB[a * k + b];
                }
            if (inv_spd(k, LHS, C)) return -3;                            /* :147 */
            for (int a = 0; a < k; ++a) { double s = 0; for (int b = 0; b < k; ++b) s += C[a * k + b] * RHS[b]; mean[a] = s; }
            if (chol_lower(k, C, Lc)) return -4;
            for (int a = 0; a < k; ++a) {                                 /* :149 MvNormal(mean, C) */
                double s = mean[a];
                for (int b = 0; b <= a; ++b) s += Lc[a * k + b] * V->z[j * k + b];
                beta[a * p + j] = s;
            }
            for (int b = 0; b < k; ++b) daxpy(n, -beta[b * p + j], S->Xk[b] + j * n, e);          /* :150 */
        }
        double Psi[64];                                                   /* :152, :513-516 */
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) {
                double s = S->scale[a * k + b];
                for (int64_t j = j0; j < j1; ++j) s += beta[a * p + j] * beta[b * p + j];
                Psi[a * k + b] = s;
            }
        if (inv_wishart_bartlett(k, Psi, V->iw_chi2 + r * k, V->iw_z + r * k * k, varBeta + r * k * k)) return -5;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* fixed effects with one or several columns: functions.jl:22-36 (sampleb!,    */
/* "Wang's trick") and :39-54 (sampleX!); X[xSet].xpx = X'X, Xp = X' (mme.jl).  */
/* z: one standard normal per column (stream purpose Z_MU, index col0 + c).    */
/* ------------------------------------------------------------------------- */
typedef struct {
    int64_t n;
    int32_t n_cols, col0;     /* col0: index of the set's first column among all fixed columns (0 is the intercept) */
    const double* data;       /* n x n_cols column-major                          */
    const double* xpx;        /* n_cols x n_cols                                  */
    double lhs0, rhs0;        /* single-column sets only (X[xSet].lhs / .rhs)     */
    const double* Xp;         /* E.str == "D": (X .* w) columns (mme.jl:136), xpx then X'(w .* X) (mme.jl:135); NULL -> data */
} ngo_fx_set;

void ngo_sample_fixed(const ngo_fx_set* F, double* b, double* e, double varE,
                      int replay, uint64_t seed, uint32_t chain, uint32_t iter, double* z)
{
    const int64_t n = F->n;
    const int c = F->n_cols;
    const double iVarE = 1.0 / varE;
    ngo_stream s = {seed, chain, iter, 0};
    if (!replay) for (int k = 0; k < c; ++k) z[k] = stream_normal(&s, NGO_P_Z_MU, (uint32_t)(F->col0 + k), 0, 0);
    for (int k = 0; k < c; ++k) daxpy(n, b[k], F->data + (int64_t)k * n, e);                       /* :42 / :49 */
    if (c == 1) {
        const double rhs = ddot(n, F->Xp ? F->Xp : F->data, e) * iVarE + F->rhs0;                 /* :43 Xp */
        const double lhs = F->xpx[0] * iVarE + F->lhs0;                                           /* :44 */
        b[0] = rhs / lhs + sqrt(1.0 / lhs) * z[0];                                               /* :45-46 */
    } else {
        double Yi[64], bVec[64];
        for (int k = 0; k < c; ++k) { Yi[k] = ddot(n, (F->Xp ? F->Xp : F->data) + (int64_t)k * n, e) * iVarE; bVec[k] = b[k]; }   /* :25 Xp */
        for (int i = 0; i < c; ++i) {                                                            /* :27-34 */
            bVec[i] = 0.0;
            double dt = 0.0;
            for (int k = 0; k < c; ++k) dt += F->xpx[i * c + k] * bVec[k];
            const double rhsb = Yi[i] - dt * iVarE;
            const double lhsb = F->xpx[i * c + i] * iVarE;
            const double invLhsb = 1.0 / lhsb;
            bVec[i] = invLhsb * rhsb + sqrt(invLhsb) * z[i];
        }
        for (int k = 0; k < c; ++k) b[k] = bVec[k];
    }
    for (int k = 0; k < c; ++k) daxpy(n, -b[k], F->data + (int64_t)k * n, e);                      /* :47 / :51 */
}

/* ------------------------------------------------------------------------- */
/* BayesR: functions.jl:238-289 (sampleBayesR!), :518-520 (sampleVarBetaR),     */
/* :536-538 (samplePi(::Vector) = Dirichlet(nLoci .+ 1)); wiring mme.jl:374-383 */
/* nc variance classes with scales v_class (first is usually 0).  Quirk 11     */
/* (SURVEY App. D): findfirst(x -> x >= rand(), cumProbs) draws a FRESH uniform */
/* for every comparison, so the variate log holds nc uniforms per locus.       */
/* ------------------------------------------------------------------------- */
typedef struct {
    int64_t n, p;
    const double* X;          /* n x p column-major, centred                      */
    const double* mpm;        /* p                                                */
    const double* lhs0;       /* p or NULL                                        */
    const double* rhs0;       /* p or NULL                                        */
    int32_t n_class, est_pi;
    const double* v_class;    /* n_class (M.vClass)                               */
    double df, scale;
    int32_t set_id, pad_;
    const double* Mp;         /* (x_j .* w) columns for E.str == "D" (mme.jl:303) or NULL -> X */
} ngo_r_set;

typedef struct {
    double* beta;             /* p                                                */
    int64_t* delta;           /* p: class of the locus, 1-based (functions.jl:261) */
    double* varBeta;          /* 1                                                */
    double* piHat;            /* n_class                                          */
    double* logPi;            /* n_class                                          */
} ngo_r_state;

typedef struct {
    int32_t replay, pad_;
    uint64_t seed;
    uint32_t chain, iter;
    double* u;                /* p x n_class (row j = the uniforms of locus j)    */
    double* z;                /* p                                                */
    double* chi2_b;           /* 1                                                */
    double* dir_pi;           /* n_class: the Dirichlet draw itself               */
} ngo_r_variates;

void ngo_r_fill_variates(const ngo_r_set* S, ngo_r_variates* V)
{
    ngo_stream s = {V->seed, V->chain, V->iter, (uint32_t)S->set_id};
    const int nc = S->n_class;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < S->p; ++j) {
        for (int v = 0; v < nc; ++v) V->u[j * nc + v] = stream_uniform(&s, NGO_P_U, (uint32_t)j, 0, (uint32_t)v);
        V->z[j] = stream_normal(&s, NGO_P_Z, (uint32_t)j, 0, 0);
    }
}

int ngo_r_sweep(const ngo_r_set* S, ngo_r_state* T, double* e, double varE, ngo_r_variates* V)
{
    const int64_t n = S->n, p = S->p;
    const int nc = S->n_class;
    if (nc < 1 || nc > 16) return -1;
    double varc[16], lhs[16], ExpLogL[16];
    int64_t nLoci[16];
    int64_t nNonZero = 0;
    double sumS = 0.0;
    const double iVarE = 1.0 / varE;
    for (int v = 0; v < nc; ++v) { varc[v] = T->varBeta[0] * S->v_class[v]; nLoci[v] = 0; }          /* :244 */
    for (int64_t j = 0; j < p; ++j) {
        const double* x = S->X + j * n;
        daxpy(n, T->beta[j], x, e);                                                              /* :249 */
        const double rhs = ddot(n, S->Mp ? S->Mp + j * n : x, e) * iVarE + (S->rhs0 ? S->rhs0[j] : 0.0);   /* :250 Mp */
        double tot = 0.0;
        for (int v = 0; v < nc; ++v) {                                                           /* :253-257 */
            lhs[v] = varc[v] == 0.0 ? 0.0 : S->mpm[j] * iVarE + (S->lhs0 ? S->lhs0[j] : 0.0) + 1.0 / varc[v];
            const double logLc = varc[v] == 0.0 ? T->logPi[v]
                                                : -0.5 * (log(varc[v] * lhs[v]) - ((rhs * rhs) / lhs[v])) + T->logPi[v];
            ExpLogL[v] = exp(logLc);
            tot += ExpLogL[v];
        }
        int cls = -1;
        double cum = 0.0;
        for (int v = 0; v < nc; ++v) {                                                           /* :259-261 */
            cum += ExpLogL[v] / tot;
            if (cum >= V->u[j * nc + v]) { cls = v; break; }
        }
        if (cls < 0) return -2;                                    /* findfirst returned nothing: the reference errors */
        T->delta[j] = cls + 1;
        nLoci[cls] += 1;
        if (varc[cls] != 0.0) {                                                                  /* :265-274 */
            nNonZero += 1;
            const double meanBeta = rhs / lhs[cls];
            const double b = meanBeta + sqrt(1.0 / lhs[cls]) * V->z[j];
            T->beta[j] = b;
            daxpy(n, -1.0 * b, x, e);
            sumS += (b * b) / S->v_class[cls];
        } else {
            T->beta[j] = 0.0;                                                                    /* :275 */
        }
    }
    ngo_stream s = {V->seed, V->chain, V->iter, (uint32_t)S->set_id};
    if (!V->replay) V->chi2_b[0] = stream_chisq(&s, NGO_P_CHI2_B, 0, 0, S->df + (double)nNonZero);
    T->varBeta[0] = (S->scale * S->df + sumS) / V->chi2_b[0];                                     /* :281, :518-520 */
    if (S->est_pi) {                                                                             /* :284-288 */
        if (!V->replay) {
            double g[16], tg = 0.0;
            for (int v = 0; v < nc; ++v) { g[v] = stream_gamma(&s, NGO_P_PI_A, 0, (uint32_t)v, (double)nLoci[v] + 1.0); tg += g[v]; }
            for (int v = 0; v < nc; ++v) V->dir_pi[v] = g[v] / tg;
        }
        for (int v = 0; v < nc; ++v) { T->piHat[v] = V->dir_pi[v]; T->logPi[v] = log(V->dir_pi[v]); }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* BayesRCpi (functions.jl:291-360) and BayesRCplus (functions.jl:362-419),   */
/* wiring mme.jl:385-418: annotations (p x nAnnot integer matrix), one common  */
/* variance and one vector of class proportions PER ANNOTATION.                */
/* Variates (explicit, like everywhere in this oracle):                        */
/*   RCpi  : u_annot[p] (Categorical: Distributions' sampler draws ONE uniform */
/*           and walks the cumulative probabilities while cp <= u),            */
/*           dirp[p][nA] (the RESULT of sampleProb = Dirichlet(annotInput +    */
/*           e_chosen) on the non-zero annotations of the locus, 0 elsewhere), */
/*           u[p][nc] (a fresh uniform per cumulative comparison, :327), z[p]  */
/*   RCplus: u[p][nA][nc], z[p][nA]                                            */
/*   both  : chi2_b[nA], dir_pi[nA][nc]                                        */
/* ------------------------------------------------------------------------- */
#define NGO_RC_MAXA 16
#define NGO_RC_MAXC 16
typedef struct {
    int64_t n, p;
    const double* X;          /* centred, n x p col-major */
    const double* mpm;
    const double* lhs0;       /* may be NULL */
    const double* rhs0;       /* may be NULL */
    int32_t n_class, n_annot, est_pi, plus;      /* plus: 0 = BayesRCpi, 1 = BayesRCplus */
    const double* v_class;    /* [n_class] */
    const int32_t* annot;     /* [p][n_annot] annotInput (mme.jl:394) */
    double df, scale;
    int32_t set_id, pad_;
} ngo_rc_set;

typedef struct {
    double* beta;             /* [p] */
    int64_t* delta;           /* [p] 1-based class */
    int64_t* annot_cat;       /* [p] 1-based annotation (RCpi; mme.jl:403) */
    double* varBeta;          /* [n_annot] */
    double* piHat;            /* [n_annot][n_class] */
    double* logPi;            /* [n_annot][n_class] */
    double* annot_prob;       /* [p][n_annot] (RCpi; mme.jl:395) */
} ngo_rc_state;

typedef struct {
    int32_t replay, pad_;
    uint64_t seed;
    uint32_t chain, iter;
    double* u_annot;
    double* dirp;
    double* u;
    double* z;
    double* chi2_b;
    double* dir_pi;
} ngo_rc_variates;

int ngo_rc_sweep(const ngo_rc_set* S, ngo_rc_state* T, double* e, double varE, ngo_rc_variates* V)
{
    const int64_t n = S->n, p = S->p;
    const int nc = S->n_class, nA = S->n_annot;
    if (nc < 1 || nc > NGO_RC_MAXC || nA < 1 || nA > NGO_RC_MAXA) return -1;
    double varc[NGO_RC_MAXA][NGO_RC_MAXC], lhs[NGO_RC_MAXA][NGO_RC_MAXC], ExpLogL[NGO_RC_MAXA][NGO_RC_MAXC], sumS[NGO_RC_MAXA];
    int64_t nLoci[NGO_RC_MAXA][NGO_RC_MAXC], nNonZero[NGO_RC_MAXA];
    const double iVarE = 1.0 / varE;
    ngo_stream s = {V->seed, V->chain, V->iter, (uint32_t)S->set_id};
    for (int a = 0; a < nA; ++a) {
        sumS[a] = 0.0; nNonZero[a] = 0;
        for (int v = 0; v < nc; ++v) { varc[a][v] = T->varBeta[a] * S->v_class[v]; nLoci[a][v] = 0; }      /* :298 / :369 */
    }
    for (int64_t j = 0; j < p; ++j) {
        const double* x = S->X + j * n;
        const int32_t* an = S->annot + j * nA;
        daxpy(n, T->beta[j], x, e);                                                              /* :303 / :374 */
        if (!S->plus) {
            const double rhs = ddot(n, x, e) * iVarE + (S->rhs0 ? S->rhs0[j] : 0.0);                /* :304 */
            double Sa[NGO_RC_MAXA], pa[NGO_RC_MAXA], tot2 = 0.0;
            for (int a = 0; a < nA; ++a) {
                Sa[a] = 0.0;
                for (int v = 0; v < nc; ++v) { lhs[a][v] = 0.0; ExpLogL[a][v] = 0.0; }               /* zeros(nAnnot,nVarClass) :305-306 */
                if (an[a] == 0) continue;                                                        /* annotNonZeroPos :307 */
                for (int v = 0; v < nc; ++v) {
                    lhs[a][v] = varc[a][v] == 0.0 ? 0.0 : S->mpm[j] * iVarE + (S->lhs0 ? S->lhs0[j] : 0.0) + 1.0 / varc[a][v];
                    const double logLv = varc[a][v] == 0.0 ? T->logPi[a * nc + v]
                                                           : -0.5 * (log(varc[a][v] * lhs[a][v]) - ((rhs * rhs) / lhs[a][v])) + T->logPi[a * nc + v];
                    ExpLogL[a][v] = exp(logLv);
                    Sa[a] += ExpLogL[a][v];                                                      /* sum(ExpLogL,dims=2) :315 */
                }
            }
            for (int a = 0; a < nA; ++a) { pa[a] = T->annot_prob[j * nA + a] * Sa[a]; tot2 += pa[a]; }    /* :315-316 */
            for (int a = 0; a < nA; ++a) pa[a] /= tot2;                                           /* :317 */
            if (!V->replay) V->u_annot[j] = stream_uniform(&s, NGO_P_U_ANNOT, (uint32_t)j, 0, 0);
            /* rand(Categorical(probAnnot)) :319 — Distributions.jl: draw = rand(); cp = p[1]; i = 1; while cp <= draw && i < n: i += 1; cp += p[i] */
            int A = 0;
            { double cp = pa[0]; while (cp <= V->u_annot[j] && A < nA - 1) { ++A; cp += pa[A]; } }
            if (!(tot2 == tot2) || an[A] == 0) return -3;             /* findfirst(isequal(A), annotNonZeroPos) finds nothing: the reference errors (:320) */
            /* sampleProb (:322, :541-544): Dirichlet(annotInput[locus, nonzero] with +1 at the chosen annotation) */
            if (!V->replay) {
                double g[NGO_RC_MAXA], tg = 0.0;
                for (int a = 0; a < nA; ++a) {
                    g[a] = 0.0;
                    if (an[a] == 0) continue;
                    g[a] = stream_gamma(&s, NGO_P_G_ANNOT, (uint32_t)j, (uint32_t)a, (double)an[a] + (a == A ? 1.0 : 0.0));
                    tg += g[a];
                }
                for (int a = 0; a < nA; ++a) V->dirp[j * nA + a] = an[a] == 0 ? 0.0 : g[a] / tg;
            }
            for (int a = 0; a < nA; ++a) if (an[a] != 0) T->annot_prob[j * nA + a] = V->dirp[j * nA + a];
            if (!V->replay) for (int v = 0; v < nc; ++v) V->u[j * nc + v] = stream_uniform(&s, NGO_P_U, (uint32_t)j, 0, (uint32_t)v);
            int cls = -1;
            double cum = 0.0;
            for (int v = 0; v < nc; ++v) {                                                       /* :325-327 */
                cum += ExpLogL[A][v] / Sa[A];
                if (cum >= V->u[j * nc + v]) { cls = v; break; }
            }
            if (cls < 0) return -2;
            T->delta[j] = cls + 1;                                                               /* :329 */
            T->annot_cat[j] = A + 1;                                                             /* :330 */
            nLoci[A][cls] += 1;
            if (varc[A][cls] != 0.0) {                                                           /* :333-341 */
                nNonZero[A] += 1;
                if (!V->replay) V->z[j] = stream_normal(&s, NGO_P_Z, (uint32_t)j, 0, 0);
                const double b = rhs / lhs[A][cls] + sqrt(1.0 / lhs[A][cls]) * V->z[j];
                T->beta[j] = b;
                daxpy(n, -1.0 * b, x, e);
                sumS[A] += (b * b) / S->v_class[cls];
            } else {
                T->beta[j] = 0.0;                                                                /* :342 */
            }
        } else {
            double tempBeta = 0.0;                                                               /* :377 */
            for (int a = 0; a < nA; ++a) {
                if (an[a] == 0) continue;                                                        /* :378 */
                const double rhs = ddot(n, x, e) * iVarE + (S->rhs0 ? S->rhs0[j] : 0.0);            /* :379 (ycorr already holds the earlier annotations' axpys) */
                double tot = 0.0;
                for (int v = 0; v < nc; ++v) {
                    lhs[a][v] = varc[a][v] == 0.0 ? 0.0 : S->mpm[j] * iVarE + (S->lhs0 ? S->lhs0[j] : 0.0) + 1.0 / varc[a][v];
                    const double logLv = varc[a][v] == 0.0 ? T->logPi[a * nc + v]
                                                           : -0.5 * (log(varc[a][v] * lhs[a][v]) - ((rhs * rhs) / lhs[a][v])) + T->logPi[a * nc + v];
                    ExpLogL[a][v] = exp(logLv);
                    tot += ExpLogL[a][v];
                }
                if (!V->replay) for (int v = 0; v < nc; ++v) V->u[(j * nA + a) * nc + v] = stream_uniform(&s, NGO_P_U, (uint32_t)j, (uint32_t)a, (uint32_t)v);
                int cls = -1;
                double cum = 0.0;
                for (int v = 0; v < nc; ++v) {                                                   /* :385-387 */
                    cum += ExpLogL[a][v] / tot;
                    if (cum >= V->u[(j * nA + a) * nc + v]) { cls = v; break; }
                }
                if (cls < 0) return -2;
                T->delta[j] = cls + 1;                                                           /* :388 (the last annotation's class stays) */
                nLoci[a][cls] += 1;
                double b = 0.0;
                if (varc[a][cls] != 0.0) {                                                       /* :391-397 */
                    nNonZero[a] += 1;
                    if (!V->replay) V->z[j * nA + a] = stream_normal(&s, NGO_P_Z, (uint32_t)j, (uint32_t)a, 0);
                    b = rhs / lhs[a][cls] + sqrt(1.0 / lhs[a][cls]) * V->z[j * nA + a];
                    sumS[a] += (b * b) / S->v_class[cls];
                }
                tempBeta += b;                                                                   /* :400 */
                daxpy(n, -1.0 * b, x, e);                                                        /* :401 */
            }
            T->beta[j] = tempBeta;                                                               /* :403 */
        }
    }
    for (int a = 0; a < nA; ++a) {                                                               /* :347-349 / :408-410 */
        if (!V->replay) V->chi2_b[a] = stream_chisq(&s, NGO_P_CHI2_B, (uint32_t)a, 0, S->df + (double)nNonZero[a]);
        T->varBeta[a] = (S->scale * S->df + sumS[a]) / V->chi2_b[a];
    }
    if (S->est_pi) {                                                                             /* :352-359 / :413-418 */
        for (int a = 0; a < nA; ++a) {
            if (!V->replay) {
                double g[NGO_RC_MAXC], tg = 0.0;
                for (int v = 0; v < nc; ++v) { g[v] = stream_gamma(&s, NGO_P_PI_A, (uint32_t)a, (uint32_t)v, (double)nLoci[a][v] + 1.0); tg += g[v]; }
                for (int v = 0; v < nc; ++v) V->dir_pi[a * nc + v] = g[v] / tg;
            }
            for (int v = 0; v < nc; ++v) { T->piHat[a * nc + v] = V->dir_pi[a * nc + v]; T->logPi[a * nc + v] = log(V->dir_pi[a * nc + v]); }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* synthetic genotype generator (SURVEY Appendix C / DESIGN.md §synthetic):    */
/* code(i,j) from Philox word (i&3) of counter (i>>2, j, 0, 0x47454e4f) under  */
/* key = seed, compared with per-column 32-bit thresholds.                    */
/* ------------------------------------------------------------------------- */
void ngo_synth_codes(uint64_t seed, int64_t n, int64_t j0, int64_t j1,
                     const uint32_t* thr0, const uint32_t* thr1, int8_t* out /* n x (j1-j0) col-major */)
{
    const uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
#pragma omp parallel for schedule(static)
    for (int64_t j = j0; j < j1; ++j) {
        int8_t* g = out + (j - j0) * n;
        for (int64_t i4 = 0; i4 < n; i4 += 4) {
            uint32_t ctr[4] = {(uint32_t)(i4 >> 2), (uint32_t)j, 0u, 0x47454e4fu}, w[4];
            ngo_philox4x32_10(ctr, key, w);
            for (int t = 0; t < 4 && i4 + t < n; ++t)
                g[i4 + t] = (int8_t)((w[t] >= thr0[j]) + (w[t] >= thr1[j]));
        }
    }
}
