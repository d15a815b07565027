// ngp_joint.cuh — multi-breed ("Tuple") BayesPR on device: sampleBayesPR!(mSet::Tuple, ...) of
// /root/reference/src/functions.jl:140-154 with the variance draw of functions.jl:513-516 and the data layout that
// mme.getMME! builds for a tuple of marker sets (mme.jl:448-467: per locus an n x k matrix M_j, mpm_j = M_j'M_j).
//
// k marker sets ("breeds", k <= 8) share the loci; per locus the k effects are drawn JOINTLY:
//     e += M_j b_j ;  RHS = M_j'e / varE ;  C = (M_j'M_j / varE + inv(Sigma_r))^-1 ;  b_j ~ MVN(C RHS, C) ;  e -= M_j b_j
// and after region r:  Sigma_r ~ InvWishart(df + |r|, scale + B_r'B_r).
//
// Kernel design: the per-locus variant of the sweep (one grid-wide reduction per locus) widened to k
// columns.  Worker CTA t owns row panel t of e (shared memory) and of every breed's genotype tiles; per locus it forms
// the k partial dots, RED-adds them as fixed-point integers into k self-synchronising accumulators (value<<8 | arrivals),
// every CTA then evaluates the k x k solve redundantly (bit-identical everywhere: the integer sums are order-free) and
// applies  e -= sum_b dbeta_b (g_b - mean_b)  to its rows.  The add-back is fused:
//     M_j'(e + M_j b_old) = M_j'e + (M_j'M_j) b_old        (M_j'M_j: centred cross-products, precomputed once on device).
// The MVN draw is  C RHS + chol(C) z  and the inverse-Wishart draw is Bartlett's decomposition, exactly as the CPU oracle
// defines them (oracle/ngp_oracle.c: ngo_mb_sweep), so replayed variates reproduce the oracle's chain.
#pragma once
#include "ngp_sweep.cuh"

namespace ngp {

struct JointDev {
    int32_t k, stream_set;            // components; set id used to address the variate stream (first member)
    int32_t set[kMaxK];               // member marker sets (index into Params::sets): breed b = set[b]
    int64_t p, n_regions;
    double df;                        // 3 + k (mme.jl:493)
    double scale[kMaxK * kMaxK];      // k x k: v .* (df - k - 1) (mme.jl:501)
    double* varBeta;                  // [n_regions][k][k]
    const int64_t* region_off;        // [n_regions + 1]
    const double* mtm;                // [p][k][k] centred cross-products M_j'M_j (mme.jl:464)
    const double* rp_z;               // replay: [iter][p][k]
    const double* rp_iw_chi2;         //         [iter][n_regions][k]
    const double* rp_iw_z;            //         [iter][n_regions][k][k] (strict lower triangle used)
};

// centred cross-products of the k columns of every locus: mtm[j][a][b] = sum_i g_a g_b - cs_a cs_b / n   (one warp per locus)
struct JointGeno {
    const uint8_t* geno[kMaxK];
    const int32_t* colsum[kMaxK];
};
static __global__ void joint_mtm_kernel(const JointGeno G, int k, int Tw, int R, int B, int64_t nblk, int64_t n, int64_t p, double* __restrict__ mtm)
{
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p) return;
    const int64_t kb = j / B;
    const int q = (int)(j - kb * B);
    const int64_t tile_bytes = (int64_t)B * R;
    const int nwords = R >> 2;
    long long s[kMaxK * (kMaxK + 1) / 2];
    for (int i = 0; i < kMaxK * (kMaxK + 1) / 2; ++i) s[i] = 0;
    for (int t = 0; t < Tw; ++t) {
        for (int wr = lane; wr < nwords; wr += 32) {
            uint32_t w[kMaxK];
            const int off = word_off(B, q, wr);
            for (int b = 0; b < k; ++b) w[b] = __ldg(reinterpret_cast<const uint32_t*>(G.geno[b] + ((int64_t)t * nblk + kb) * tile_bytes) + off);
            int x = 0;
            for (int a = 0; a < k; ++a)
                for (int b = a; b < k; ++b, ++x) s[x] += (long long)__dp4a(w[a], w[b], 0u);
        }
    }
    int x = 0;
    for (int a = 0; a < k; ++a)
        for (int b = a; b < k; ++b, ++x) {
            long long v = s[x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                const double ca = (double)G.colsum[a][j], cb = (double)G.colsum[b][j], nn = (double)n;
                const double c = ((double)v * nn - ca * cb) / nn;          // integer-exact numerator
                mtm[(j * k + a) * k + b] = c;
                mtm[(j * k + b) * k + a] = c;
            }
        }
}

// ----------------------------------------------------------------------------- the kernel
template <int KK>
__global__ void __launch_bounds__(kThreads, 1) joint_kernel(const Params P, const JointDev J)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    const int Tw = P.Tw, R = P.R, B = P.B;
    constexpr int k = KK;                            // components (J.k == KK)
    const bool is_chain = (t == Tw);                 // owns no rows; writes the outputs
    double* misc = reinterpret_cast<double*>(smem);                          // [0..27] block_sum scratch, [32..] scalars, [40..] dbeta / K
    double* invB = misc + 96;                                                // [64] inv(Sigma_r)
    long long* lprev = reinterpret_cast<long long*>(misc + 160);             // [kSlots][kMaxK] previous accumulator values
    double* e_s = misc + 160 + kSlots * kMaxK;
    SyncArea* sy = P.sync;
    GridSync gs{&sy->err, &sy->counter, 0ull, (unsigned)(Tw + 1), 0ull, 1};
    const int64_t row0 = (int64_t)t * R;
    const int nrow = is_chain ? 0 : (int)max((int64_t)0, min((int64_t)R, P.n - row0));
    const int nwords = R >> 2;
    const int nblk = (int)(P.sets[J.set[0]].p_pad / B);
    const int64_t tile_bytes = (int64_t)B * R;

    __shared__ const uint8_t* s_geno[kMaxK];
    __shared__ const double* s_mean[kMaxK];
    __shared__ double* s_beta[kMaxK];
    if (tid < kMaxK) {
        const SetDev& S = P.sets[J.set[tid < k ? tid : 0]];
        s_geno[tid] = S.geno; s_mean[tid] = S.mean; s_beta[tid] = S.beta;
    }
    if (!is_chain) for (int r = tid; r < R; r += kThreads) e_s[r] = (r < nrow) ? P.e[row0 + r] : 0.0;
    for (int i = tid; i < kSlots * kMaxK; i += kThreads) lprev[i] = sy->acc[(size_t)(i / kMaxK) * kMaxB * kAccStride + (i % kMaxK) * kAccStride];
    __syncthreads();

    unsigned rk = 0;                 // running locus count: indexes the accumulator ring
    double mu = P.sc->mu;
    const long long iter0 = P.sc->iter;

    for (int it = 0; it < P.n_iter; ++it) {
        const uint32_t iter = (uint32_t)(iter0 + it + 1);
        const int64_t rp_row = (int64_t)iter - 1 - P.replay_base;

        // ------------------------------------------------------------------ phase 0: varE, intercept (samplers.jl:32-47)
        double ee = 0.0, se = 0.0;
        if (!is_chain) for (int r = tid; r < R; r += kThreads) { const double x = e_s[r]; ee = fma(x, x, ee); se += x; }
        block_sum2(ee, se, misc);
        if (tid == 0) {
            sy->part[2 * t] = ee; sy->part[2 * t + 1] = se;
            gs.nbar++; gs.arrive(P);
        } else gs.nbar++;
        if (warp == 0) {
            gs.wait_warp();
            double a = 0.0, b = 0.0;
            for (int c = lane; c < Tw; c += 32) { a += __ldcg(&sy->part[2 * c]); b += __ldcg(&sy->part[2 * c + 1]); }
            a = warp_sum(a); b = warp_sum(b);
            if (lane == 0) {
                double varE = P.varE_in;
                if (P.do_varE) {
                    Stream st{P.key0, P.key1, P.chain, iter, 0u};
                    const double chi2 = P.replay ? P.rp_chi2_e[rp_row] : stream_chisq(st, P_CHI2_E, 0, 0, P.df_e + (double)P.n);
                    varE = (P.df_e * P.scale_e + a) / chi2;                       // functions.jl:524
                }
                double dmu = 0.0;
                if (P.has_mu && P.do_mu) {                                        // functions.jl:39-47
                    Stream st{P.key0, P.key1, P.chain, iter, 0u};
                    const double zmu = P.replay ? P.rp_z_mu[rp_row] : stream_normal(st, P_Z_MU, 0);
                    const double iVarE = 1.0 / varE;
                    const double rhs = (b + (double)P.n * mu) * iVarE + P.mu_rhs0;
                    const double lhs = (double)P.n * iVarE + P.mu_lhs0;
                    const double mu_new = rhs / lhs + sqrt(1.0 / lhs) * zmu;
                    dmu = mu - mu_new;
                    mu = mu_new;
                }
                const double nn = (double)P.n;
                double M = 2.0 * sqrt(nn) * (sqrt(a) + sqrt(nn) * fabs(dmu));     // |sum g e| <= 2 sqrt(n) ||e||
                if (!(M > 1e-300)) M = 1e-300;
                int ex; (void)frexp(M, &ex);
                int sh = 62 - kCntBits - 4 - ex;
                sh = max(-1000, min(1000, sh));
                misc[32] = varE; misc[33] = dmu; misc[34] = b + nn * dmu; misc[35] = (double)sh; misc[36] = mu;
            }
        }
        __syncthreads();
        const double varE = misc[32];
        const double dmu = misc[33];
        const double Stot = misc[34];                 // 1'e: invariant under marker updates (centred columns)
        const int sh = (int)misc[35];
        mu = misc[36];
        const double fx_scale = ldexp(1.0, sh), fx_inv = ldexp(1.0, -sh);
        const double iVarE = 1.0 / varE;
        if (dmu != 0.0) for (int r = tid; r < nrow; r += kThreads) e_s[r] += dmu;
        __syncthreads();

        // ------------------------------------------------------------------ the joint sweep (functions.jl:140-154)
        for (int64_t rg = 0; rg < J.n_regions; ++rg) {
            if (tid == 0) {                                                       // invB = inv(varBeta[mSet][r]), functions.jl:143
                double Sg[kMaxK * kMaxK], Iv[kMaxK * kMaxK];
                for (int i = 0; i < k * k; ++i) Sg[i] = __ldcg(&J.varBeta[rg * k * k + i]);
                if (!jt_inv_spd<k>(Sg, Iv)) { atomicOr(&sy->err, 2); for (int i = 0; i < k * k; ++i) Iv[i] = 0.0; }
                for (int i = 0; i < k * k; ++i) invB[i] = Iv[i];
            }
            __syncthreads();
            const int64_t j0 = J.region_off[rg], j1 = J.region_off[rg + 1];
            for (int64_t j = j0; j < j1; ++j, ++rk) {
                const int kb = (int)(j / B), q = (int)(j % B);
                const int slot = (int)(rk & (kSlots - 1));
                long long* acc = sy->acc + (size_t)slot * kMaxB * kAccStride;
                const int64_t toff = ((int64_t)t * nblk + kb) * tile_bytes;
                // k partial dots of this panel
                if (!is_chain) {
                    double a[kMaxK];
#pragma unroll
                    for (int b = 0; b < kMaxK; ++b) a[b] = 0.0;
                    for (int wr = tid; wr < nwords; wr += kThreads) {
                        const double* ep = e_s + 4 * wr;
                        const double e0 = ep[0], e1 = ep[1], e2 = ep[2], e3 = ep[3];
                        const int off = word_off(B, q, wr);
#pragma unroll
                        for (int b = 0; b < kMaxK; ++b)
                            if (b < k) {
                                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(s_geno[b] + toff) + off);
                                a[b] = fma((double)(w & 0xff), e0, a[b]); a[b] = fma((double)((w >> 8) & 0xff), e1, a[b]);
                                a[b] = fma((double)((w >> 16) & 0xff), e2, a[b]); a[b] = fma((double)(w >> 24), e3, a[b]);
                            }
                    }
#pragma unroll
                    for (int b = 0; b < kMaxK; b += 2)
                        if (b < k) {
                            block_sum2(a[b], a[b + 1], misc);
                            if (tid == 0) {
                                const double xs = a[b] * fx_scale;
                                if (!(fabs(xs) < 9007199254740992.0)) atomicOr(&sy->err, 1);
                                red_add_u64(acc + b * kAccStride, (long long)((unsigned long long)__double2ll_rn(xs) << kCntBits) + 1);
                                if (b + 1 < k) {
                                    const double ys = a[b + 1] * fx_scale;
                                    if (!(fabs(ys) < 9007199254740992.0)) atomicOr(&sy->err, 1);
                                    red_add_u64(acc + (b + 1) * kAccStride, (long long)((unsigned long long)__double2ll_rn(ys) << kCntBits) + 1);
                                }
                            }
                        }
                }
                if (warp == 0) {
                    // everything that does not depend on the sums is fetched / drawn while they are in flight: lane b draws z_b
                    Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)J.stream_set};
                    double zb = 0.0, meanb = 0.0, boldb = 0.0;
                    if (lane < k) {
                        zb = P.replay ? J.rp_z[(rp_row * J.p + j) * k + lane] : stream_normal(st, P_Z, (uint32_t)j, 0, (uint32_t)lane);
                        meanb = s_mean[lane][j];
                        boldb = __ldcg(&s_beta[lane][j]);
                    }
                    double MtM[k * k];
#pragma unroll
                    for (int i = 0; i < k * k; ++i) MtM[i] = __ldg(&J.mtm[j * k * k + i]);
                    // lane b waits for accumulator b to hold all Tw partial sums
                    double rb = 0.0;
                    if (lane < k) {
                        long long* pv = lprev + slot * kMaxK + lane;
                        long long cur;
                        do { cur = ld_relaxed_s64(acc + lane * kAccStride); } while (((cur - *pv) & 0xFF) != (long long)Tw);
                        const double A = (double)((cur - *pv - (long long)Tw) >> kCntBits) * fx_inv;
                        *pv = cur;
                        rb = A - meanb * Stot;                                    // x_b'e of the centred column
                    }
                    double r[k], z[k], bold[k], mean[k];
#pragma unroll
                    for (int b = 0; b < k; ++b) {
                        r[b] = __shfl_sync(0xffffffffu, rb, b); z[b] = __shfl_sync(0xffffffffu, zb, b);
                        bold[b] = __shfl_sync(0xffffffffu, boldb, b); mean[b] = __shfl_sync(0xffffffffu, meanb, b);
                    }
                    if (lane == 0) {
                        double LHS[k * k], C[k * k], Lc[k * k], bn[k], rhs[k];
#pragma unroll
                        for (int a = 0; a < k; ++a)
#pragma unroll
                            for (int b = 0; b < k; ++b) LHS[a * k + b] = MtM[a * k + b] * iVarE + invB[a * k + b];
                        const bool ok = jt_inv_spd<k>(LHS, C) && jt_chol<k>(C, Lc);   // functions.jl:147
                        if (!ok) atomicOr(&sy->err, 2);
#pragma unroll
                        for (int b = 0; b < k; ++b) {
                            double rr = r[b];
#pragma unroll
                            for (int a = 0; a < k; ++a) rr = fma(MtM[b * k + a], bold[a], rr);       // add-back fused (functions.jl:145)
                            rhs[b] = rr * iVarE;                                                    // functions.jl:146
                        }
                        double Kc = 0.0;
#pragma unroll
                        for (int a = 0; a < k; ++a) {
                            double m = 0.0;
#pragma unroll
                            for (int b = 0; b < k; ++b) m += C[a * k + b] * rhs[b];                  // functions.jl:148
                            double sdraw = m;
#pragma unroll
                            for (int b = 0; b <= a; ++b) sdraw += Lc[a * k + b] * z[b];              // functions.jl:149: MvNormal(mean, C)
                            bn[a] = ok ? sdraw : bold[a];
                            const double db = bn[a] - bold[a];
                            misc[40 + a] = db;
                            Kc = fma(db, mean[a], Kc);
                            if (is_chain) s_beta[a][j] = bn[a];
                        }
                        misc[40 + kMaxK] = Kc;
                    }
                }
                __syncthreads();
                if (!is_chain) {                                                  // e -= M_j (b_new - b_old), functions.jl:150
                    double db[kMaxK];
                    bool any = false;
#pragma unroll
                    for (int b = 0; b < kMaxK; ++b) { db[b] = (b < k) ? misc[40 + b] : 0.0; any = any || (db[b] != 0.0); }
                    const double K = misc[40 + kMaxK];
                    if (any) {
                        for (int wr = tid; wr < nwords; wr += kThreads) {
                            const int off = word_off(B, q, wr);
                            double s0 = -K, s1 = -K, s2 = -K, s3 = -K;
#pragma unroll
                            for (int b = 0; b < kMaxK; ++b)
                                if (b < k) {
                                    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(s_geno[b] + toff) + off);
                                    s0 = fma(db[b], (double)(w & 0xff), s0); s1 = fma(db[b], (double)((w >> 8) & 0xff), s1);
                                    s2 = fma(db[b], (double)((w >> 16) & 0xff), s2); s3 = fma(db[b], (double)(w >> 24), s3);
                                }
                            double* ep = e_s + 4 * wr;
                            const int lim = nrow - 4 * wr;
                            if (lim > 0) ep[0] -= s0;
                            if (lim > 1) ep[1] -= s1;
                            if (lim > 2) ep[2] -= s2;
                            if (lim > 3) ep[3] -= s3;
                        }
                    }
                }
                __syncthreads();
            }
        }

        // ------------------------------------------------------------------ phase 3: Sigma_r (functions.jl:152, 513-516), posterior sums
        __syncthreads();
        gs.nbar++;
        if (tid == 0) gs.arrive(P);
        if (warp == 0) gs.wait_warp();
        __syncthreads();
        if (P.accumulate) {
            for (int b = 0; b < k; ++b) {
                const SetDev& S = P.sets[J.set[b]];
                for (int64_t j = (int64_t)t * kThreads + tid; j < J.p; j += (int64_t)(Tw + 1) * kThreads) {
                    const double bj = __ldcg(&S.beta[j]);
                    S.sum_beta[j] += bj;
                    S.sum_beta2[j] = fma(bj, bj, S.sum_beta2[j]);
                    S.sum_delta[j] += 1.0;
                }
            }
        }
        {
            Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)J.stream_set};
            for (int64_t rg = (int64_t)t * kWarps + warp; rg < J.n_regions; rg += (int64_t)(Tw + 1) * kWarps) {
                const int64_t j0 = J.region_off[rg], j1 = J.region_off[rg + 1];
                double Sb[kMaxK * kMaxK];
                for (int i = 0; i < k * k; ++i) Sb[i] = 0.0;
                for (int64_t j = j0 + lane; j < j1; j += 32) {
                    double bj[kMaxK];
                    for (int b = 0; b < k; ++b) bj[b] = __ldcg(&s_beta[b][j]);
                    for (int a = 0; a < k; ++a)
                        for (int b = a; b < k; ++b) Sb[a * k + b] = fma(bj[a], bj[b], Sb[a * k + b]);
                }
                for (int a = 0; a < k; ++a)
                    for (int b = a; b < k; ++b) { const double v = warp_sum(Sb[a * k + b]); Sb[a * k + b] = v; Sb[b * k + a] = v; }
                if (lane == 0) {
                    double Psi[kMaxK * kMaxK], chi2[kMaxK], zl[kMaxK * kMaxK], Sg[kMaxK * kMaxK];
                    const double dfr = J.df + (double)(j1 - j0);
                    for (int i = 0; i < k * k; ++i) { Psi[i] = J.scale[i] + Sb[i]; zl[i] = 0.0; }
                    for (int i = 0; i < k; ++i) {
                        chi2[i] = P.replay ? J.rp_iw_chi2[(rp_row * J.n_regions + rg) * k + i]
                                           : stream_chisq(st, P_IW, (uint32_t)rg, (uint32_t)(i * k + i), dfr - (double)i);
                        for (int jj = 0; jj < i; ++jj)
                            zl[i * k + jj] = P.replay ? J.rp_iw_z[((rp_row * J.n_regions + rg) * k + i) * k + jj]
                                                      : stream_normal(st, P_IW, (uint32_t)rg, 0, (uint32_t)(i * k + jj));
                    }
                    if (jt_inv_wishart<k>(Psi, chi2, zl, Sg)) { for (int i = 0; i < k * k; ++i) J.varBeta[rg * k * k + i] = Sg[i]; }
                    else atomicOr(&sy->err, 2);
                }
            }
        }
        if (is_chain && tid == 0) {
            P.sc->mu = mu; P.sc->varE = varE; P.sc->iter = iter0 + it + 1;
            if (P.accumulate) P.sc->n_post += 1;
        }
    }
    __syncthreads();
    for (int r = tid; r < nrow; r += kThreads) P.e[row0 + r] = e_s[r];
}

}  // namespace ngp
