// ngp_sweep.cuh — the persistent cooperative Gibbs kernel of libngp (sm_100a).
//
// One launch runs n_iter whole iterations of samplers.runSampler!'s loop body
// (/root/reference/src/samplers.jl:32-53) on device:
//   phase 0  e'e and 1'e  -> varE (functions.jl:523-525), intercept (functions.jl:39-47)
//   phase 1  per-marker constants + variates for every marker set (grid-parallel)
//   phase 2  the marker sweep (functions.jl:118-137 / 157-195 / 197-236), CTA t owning
//            row panel t of the genotypes and of the residual e (resident in smem)
//   phase 3  variance components and pi (functions.jl:509-511, 531-533), posterior sums
//
// Phase 2, blocked exact sweep with one block of look-ahead.
// Warp 0 of every CTA is the "chain warp", warps 1..8 are "workers".  In step k
//   workers : e -= X_{k-1} dbeta_{k-1}            (tile k-1 still in smem: column read from HBM once)
//             A_{k+1} = sum_i (1 + g/4) e_i        (tile k+1, prefetched by TMA; one PRMT + one DFMA per code)
//             fixed point, RED (value<<8)+1 into the global int64 accumulators of block k+1:
//             integer adds are associative (bit-reproducible) and the low byte counts arrivals,
//             so an accumulator is its own barrier - no fence, no counter, no grid-wide stall;
//   chain   : polls the accumulators of block k (their REDs were issued one step earlier),
//             r_j = 4(A_j - S) - m_j S - [cross-Gram rows of block k-1] dbeta_{k-1} - [Gram rows of block k] dbeta_k,
//             runs the B dependent scalar updates redundantly in every CTA (identical inputs and code =>
//             identical results, no broadcast); mixture priors: 32 lanes evaluate 32 markers speculatively
//             and serialise only on markers whose effect actually changes.
// The dots of block k+1 therefore never wait for the chain of block k: its effect on them is restored
// exactly by G_{k+1,k} dbeta_k (integer Gram, precomputed once).
// The "literal" variant does dot / reduce / draw / axpy per marker from registers with one grid-wide
// reduction per marker (the north-star baseline whose sync cost we report).
#pragma once
#include "ngp_device.cuh"

namespace ngp {

struct SmemLayout {
    int off_e, off_tile, off_blk, off_red, off_prev, off_nzdb, off_nzcs, off_nzidx, off_outb, off_outv, off_outi, off_misc, off_mbar, total;
    int tile_bytes, blk_bytes;
};

__host__ __device__ inline SmemLayout smem_layout(int R, int B, int stages)
{
    SmemLayout L;
    int o = 0;
    L.tile_bytes = B * R;
    L.blk_bytes = blk_bytes(B);
    L.off_e = o;      o += R * 8;
    L.off_tile = o;   o += stages * L.tile_bytes;
    L.off_blk = o;    o += stages * L.blk_bytes;
    L.off_red = o;    o += 2 * kWorkerWarps * kMaxB * 8;
    L.off_prev = o;   o += kSlots * (kMaxB + 1) * 8;
    L.off_nzdb = o;   o += 2 * kMaxB * 8;
    L.off_nzcs = o;   o += 2 * kMaxB * 8;
    L.off_nzidx = o;  o += 2 * kMaxB * 4;
    L.off_outb = o;   o += 2 * kMaxB * 8;
    L.off_outv = o;   o += 2 * kMaxB * 8;
    L.off_outi = o;   o += 2 * kMaxB * 4;
    L.off_misc = o;   o += 64 * 8;
    L.off_mbar = o;   o += 8 * 8;
    L.total = o;
    return L;
}

// ----------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;   // identical on every lane (x+y == y+x)
}

// deterministic CTA sum of two values; result valid in all threads
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch)
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { scratch[2 * w] = a; scratch[2 * w + 1] = b; }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) { sa += scratch[2 * i]; sb += scratch[2 * i + 1]; }
    __syncthreads();
    a = sa; b = sb;
}

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkerWarps * 32) : "memory"); }

struct GridSync {
    unsigned long long* counter;
    unsigned long long nbar;     // barriers this CTA has taken part in
    unsigned int T;
    __device__ __forceinline__ void arrive()        // one thread, after a __syncthreads
    {
        arrive_release(counter);
    }
    __device__ __forceinline__ void wait_warp()     // whole warp polls
    {
        const unsigned long long target = nbar * (unsigned long long)T;
        while ((unsigned long long)ld_relaxed_s64(reinterpret_cast<const long long*>(counter)) < target) { }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");      // acquire once, not one CCTL.IVALL per poll
    }
};

// ----------------------------------------------------------------------------- phase 1: marker constants
// Everything in the scalar update of marker j that does not depend on the running
// residual is hoisted here.  With rr = x_j'e + d_j*beta_old_j the update is
//   BayesPR : beta_new = (rr*iVarE + rho_j)/lhs + z/sqrt(lhs)              (functions.jl:129-132)
//   BayesB/C: include iff u < 1/(1+exp(logDelta0-logDelta1))                (functions.jl:169-174, 209-216)
//             <=> A_j + B_j*rr^2 < t_j , A_j = (log v1 - log v0)/2 + logPi0 - logPi1,
//                 B_j = (1/v1 - 1/v0)/2, t_j = log(1/u - 1)
//             beta_new = (rr*iVarE [+ rho_j, BayesB only])/lhs + z/sqrt(lhs)  (functions.jl:177-180, 219-222)
// stored as beta_new = rr*C_j + QSZ_j.
__device__ __forceinline__ void prep_marker(const Params& P, const SetDev& S, int sidx, int64_t j, double varE,
                                            uint32_t iter, int64_t rp_row)
{
    const int B = P.B;
    double* c = reinterpret_cast<double*>(S.blk + (j / B) * (int64_t)blk_bytes(B) + B * B * 4) + (j % B);
    double fA = 0.0, fB = 0.0, fT = -INFINITY, fC = 0.0, fQSZ = 0.0, fD = 0.0, fBOLD = 0.0, fMEAN = 0.0, fCS = 0.0, fCHI = 1.0;
    if (j < S.p) {
        const double d = S.d[j];
        const double bold = __ldcg(&S.beta[j]);
        const double iVarE = 1.0 / varE;
        const double l0 = S.lhs0 ? S.lhs0[j] : 0.0;
        const double r0 = S.rhs0 ? S.rhs0[j] : 0.0;
        Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)sidx};
        const double z = P.replay ? S.rp_z[rp_row * S.p + j] : stream_normal(st, P_Z, (uint32_t)j);
        double vb;
        if (S.method == 0) vb = __ldcg(&S.varBeta[S.region_of ? S.region_of[j] : 0]);
        else if (S.method == 1) vb = __ldcg(&S.varBeta[j]);
        else vb = __ldcg(&S.varBeta[0]);
        const double lhs = d * iVarE + l0 + 1.0 / vb;          // 1/0 -> Inf (BayesB quirk, functions.jl:186)
        const double ilhs = 1.0 / lhs;
        fC = iVarE * ilhs;
        fQSZ = sqrt(ilhs) * z + ((S.method == 2) ? 0.0 : r0 * ilhs);
        fD = d; fBOLD = bold; fMEAN = S.mean[j]; fCS = (double)S.colsum[j];
        if (S.method != 0) {
            const double u = P.replay ? S.rp_u[rp_row * S.p + j] : stream_uniform(st, P_U, (uint32_t)j);
            const double v0 = d * varE;
            const double v1 = (d * d) * vb + v0;
            fA = 0.5 * (log(v1) - log(v0)) + (__ldcg(&S.pi[2]) - __ldcg(&S.pi[3]));
            fB = 0.5 * (1.0 / v1 - 1.0 / v0);
            fT = log(1.0 / u - 1.0);
            if (S.method == 1)
                fCHI = P.replay ? S.rp_chi2b[rp_row * S.nvar + j] : stream_chisq(st, P_CHI2_B, (uint32_t)j, 0, S.df + 1.0);
        } else {
            fT = INFINITY;   // BayesPR: always "included"
        }
    }
    c[F_A * B] = fA; c[F_B * B] = fB; c[F_T * B] = fT; c[F_C * B] = fC; c[F_QSZ * B] = fQSZ;
    c[F_D * B] = fD; c[F_BOLD * B] = fBOLD; c[F_MEAN * B] = fMEAN; c[F_CS * B] = fCS; c[F_CHI * B] = fCHI;
}

// ----------------------------------------------------------------------------- the kernel
template <int B>
__global__ void __launch_bounds__(kThreads, 1) gibbs_kernel(const Params P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    const int R = P.R, S_ = P.stages;
    const SmemLayout L = smem_layout(R, B, S_);
    double* e_s = reinterpret_cast<double*>(smem + L.off_e);
    double* red = reinterpret_cast<double*>(smem + L.off_red);
    long long* prev = reinterpret_cast<long long*>(smem + L.off_prev);
    double* nz_db = reinterpret_cast<double*>(smem + L.off_nzdb);      // [2][kMaxB]  4*dbeta of the changed markers of a block
    double* nz_cs = reinterpret_cast<double*>(smem + L.off_nzcs);      // [2][kMaxB]  their column sums
    int* nz_idx = reinterpret_cast<int*>(smem + L.off_nzidx);          // [2][kMaxB]  their index inside the block
    double* out_b = reinterpret_cast<double*>(smem + L.off_outb);      // [2][kMaxB] staged outputs of a block (written to HBM by the producer warp of CTA 0)
    double* out_v = reinterpret_cast<double*>(smem + L.off_outv);
    int* out_i = reinterpret_cast<int*>(smem + L.off_outi);
    double* misc = reinterpret_cast<double*>(smem + L.off_misc);       // [0..17] block_sum scratch, [32..] scalars, [40+2i] nnz, [41+2i] K
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L.off_mbar);
    SyncArea* sy = P.sync;
    constexpr int NB = B >> 5;             // 32-marker groups per block (lane <-> marker b*32 + lane)

    GridSync gs{&sy->counter, 0ull, (unsigned)P.T};
    const int64_t row0 = (int64_t)t * R;
    const int nrow = (int)max((int64_t)0, min((int64_t)R, P.n - row0));   // real rows of this panel

    for (int r = tid; r < R; r += kThreads) e_s[r] = (r < nrow) ? P.e[row0 + r] : 0.0;
    for (int q = tid; q < kSlots * (kMaxB + 1); q += kThreads) prev[q] = 0;
    if (tid == 0) {
        for (int s = 0; s < S_; ++s) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // cycle counters (thread 0 = chain warp, thread 32 = first worker warp), see ngp_get_profile
    long long pf[kProf];
#pragma unroll
    for (int i = 0; i < kProf; ++i) pf[i] = 0;
    long long tc = clock64();
#define NGP_TICK(i) do { const long long now__ = clock64(); pf[i] += now__ - tc; tc = now__; } while (0)
    unsigned gk = 0;         // tile ring position in [0, 2*stages): stage = gk mod stages, mbarrier parity = (gk div stages) & 1
    unsigned rk = 0;         // running reduction count (accumulator slot = rk & (kSlots-1))
    double mu = P.sc->mu;
    const long long iter0 = P.sc->iter;

    for (int it = 0; it < P.n_iter; ++it) {
        const uint32_t iter = (uint32_t)(iter0 + it + 1);
        const int64_t rp_row = (int64_t)iter - 1 - P.replay_base;
        tc = clock64();

        // ------------------------------------------------------------------ phase 0
        double ee = 0.0, se = 0.0;
        for (int r = tid; r < R; r += kThreads) { const double x = e_s[r]; ee = fma(x, x, ee); se += x; }
        block_sum2(ee, se, misc);
        if (tid == 0) {
            sy->part[2 * t] = ee; sy->part[2 * t + 1] = se;
            gs.nbar++; gs.arrive();
        } else gs.nbar++;
        if (warp == 0) {
            gs.wait_warp();
            double a = 0.0, b = 0.0;
            for (int c = lane; c < P.T; c += 32) { a += __ldcg(&sy->part[2 * c]); b += __ldcg(&sy->part[2 * c + 1]); }
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
                // fixed-point scale of the sweep reductions: |sum g e| <= 2 sqrt(n) ||e||
                const double nn = (double)P.n;
                double M = 2.0 * sqrt(nn) * (sqrt(a) + sqrt(nn) * fabs(dmu));
                if (!(M > 1e-300)) M = 1e-300;
                int ex; (void)frexp(M, &ex);
                int sh = 62 - kCntBits - 8 - ex;          // 8 count bits + 2^8 headroom for growth of ||e|| inside the iteration
                sh = max(-1000, min(1000, sh));
                misc[32] = varE; misc[33] = dmu; misc[34] = b + nn * dmu; misc[35] = (double)sh; misc[36] = mu;
            }
        }
        __syncthreads();
        const double varE = misc[32];
        const double dmu = misc[33];
        const double Stot = misc[34];                 // 1'e after the intercept update; invariant under marker updates
        const int sh = (int)misc[35];
        mu = misc[36];
        const double fx_scale = ldexp(1.0, sh), fx_inv = ldexp(1.0, -sh);
        if (dmu != 0.0) for (int r = tid; r < nrow; r += kThreads) e_s[r] += dmu;
        NGP_TICK(8);

        // ------------------------------------------------------------------ phase 1
        for (int s = 0; s < P.n_sets; ++s) {
            if (!((P.set_mask >> s) & 1)) continue;
            const SetDev& S = P.sets[s];
            for (int64_t j = (int64_t)t * kThreads + tid; j < S.p_pad; j += (int64_t)P.T * kThreads)
                prep_marker(P, S, s, j, varE, iter, rp_row);
        }
        __syncthreads();
        gs.nbar++;
        if (tid == 0) gs.arrive();
        if (warp == 0) gs.wait_warp();
        __syncthreads();
        NGP_TICK(9);

        // ------------------------------------------------------------------ phase 2 + 3 per marker set
        for (int s = 0; s < P.n_sets; ++s) {
            if (!((P.set_mask >> s) & 1)) continue;
            const SetDev& S = P.sets[s];
            const int nblk = (int)(S.p_pad / B);
            const double inv_n = 1.0 / (double)P.n;
            // hot-loop copies of the set descriptor (it lives in global memory; the volatile polls would force reloads)
            const int method = S.method;
            const int64_t p_real = S.p;
            double* const beta_g = S.beta;
            int32_t* const delta_g = S.delta;
            double* const vb_g = S.varBeta;
            const int32_t* const gramx_g = S.gramx;
            const uint8_t* const blk_g = S.blk;
            const double sdf = S.scale * S.df;
            double acc_bb = 0.0, acc_n = 0.0;          // warp 0: per-lane partials of beta'beta and nLoci
            tc = clock64();

            if (P.kernel == 0) {
                // ============================ blocked exact sweep, one block of look-ahead ============================
                const uint8_t* gbase = S.geno + (int64_t)t * S.p_pad * R;
                // stages is 3 or 4: no runtime integer division on the per-block path
                auto stage_of = [&](int k) { const unsigned x = gk + (unsigned)k; return (int)(S_ == 4 ? (x & 3u) : (x % 3u)); };
                auto parity_of = [&](int k) { const unsigned x = gk + (unsigned)k; return (uint32_t)((S_ == 4 ? (x >> 2) : (x / 3u)) & 1u); };
                auto issue = [&](int k) {
                    const int stg = stage_of(k);
                    mbar_expect_tx(&mbar[stg], (uint32_t)(L.tile_bytes + L.blk_bytes));
                    bulk_g2s(smem + L.off_tile + stg * L.tile_bytes, gbase + (int64_t)k * B * R, (uint32_t)L.tile_bytes, &mbar[stg]);
                    bulk_g2s(smem + L.off_blk + stg * L.blk_bytes, blk_g + (int64_t)k * L.blk_bytes, (uint32_t)L.blk_bytes, &mbar[stg]);
                };
                if (tid == kProducerWarp * 32) {
                    fence_proxy_async();     // consts were written through the generic proxy by other CTAs
                    for (int k = 0; k < min(S_, nblk); ++k) issue(k);
                }
                if (tid == 0) { misc[40] = 0.0; misc[42] = 0.0; }      // nnz of the two nz lists
                const int wtid = tid - 32;                  // worker thread id (0..255), negative in the chain warp
                const int ww = warp - 1;                    // worker warp id
                const int ngrp = R >> 3;
                const int g0 = (ww * ngrp) / kWorkerWarps, g1 = ((ww + 1) * ngrp) / kWorkerWarps;

                int zl[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) zl[i] = __ldg(&sy->zero16[i]);      // runtime zeros, see dec_byte_z
                const int r_lo = 8 * g0, r_hi = min(8 * g1, nrow);               // the rows this worker warp owns (dot AND axpy)

                // partial sums of block kk over this warp's rows: lane <-> markers lane, 32+lane
                auto worker_dot = [&](int kk) {
                    const int stg = stage_of(kk);
                    mbar_wait(&mbar[stg], parity_of(kk));
                    if (tid == 32) NGP_TICK(12);
                    const uint8_t* tile = smem + L.off_tile + stg * L.tile_bytes;
                    double* redk = red + (kk & 1) * (kWorkerWarps * kMaxB);     // double-buffered: one barrier per block
                    double a[NB][4];
#pragma unroll
                    for (int b = 0; b < NB; ++b) { a[b][0] = a[b][1] = a[b][2] = a[b][3] = 0.0; }
                    for (int g = g0; g < g1; ++g) {
                        const double2* ep = reinterpret_cast<const double2*>(e_s + 8 * g);
                        const double2 e01 = ep[0], e23 = ep[1], e45 = ep[2], e67 = ep[3];
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const uint2 w = *reinterpret_cast<const uint2*>(tile + (b * 32 + lane) * R + 8 * g);
                            a[b][0] = fma(dec_byte_z(w.x, 0, zl[0]), e01.x, a[b][0]);
                            a[b][1] = fma(dec_byte_z(w.x, 1, zl[1]), e01.y, a[b][1]);
                            a[b][2] = fma(dec_byte_z(w.x, 2, zl[2]), e23.x, a[b][2]);
                            a[b][3] = fma(dec_byte_z(w.x, 3, zl[3]), e23.y, a[b][3]);
                            a[b][0] = fma(dec_byte_z(w.y, 0, zl[4]), e45.x, a[b][0]);
                            a[b][1] = fma(dec_byte_z(w.y, 1, zl[5]), e45.y, a[b][1]);
                            a[b][2] = fma(dec_byte_z(w.y, 2, zl[6]), e67.x, a[b][2]);
                            a[b][3] = fma(dec_byte_z(w.y, 3, zl[7]), e67.y, a[b][3]);
                        }
                    }
#pragma unroll
                    for (int b = 0; b < NB; ++b) redk[ww * B + b * 32 + lane] = (a[b][0] + a[b][1]) + (a[b][2] + a[b][3]);
                    if (tid == 32) NGP_TICK(13);
                    worker_bar();
                    if (tid == 32) NGP_TICK(14);
                    if (wtid < B) {
                        double A = 0.0;
#pragma unroll
                        for (int c = 0; c < kWorkerWarps; ++c) A += redk[c * B + wtid];
                        const double xs = A * fx_scale;
                        if (!(fabs(xs) < 9007199254740992.0)) atomicOr(&sy->err, 1);     // 2^53 << 8 still fits
                        long long* acc = sy->acc + (int64_t)((rk + (unsigned)kk) & (kSlots - 1)) * (kMaxB + 1) * kAccStride;
                        red_add_u64(acc + wtid * kAccStride, (__double2ll_rn(xs) << kCntBits) + 1);
                    }
                };
                // e -= sum_q dbeta_q (g_q - m_q) for the changed markers of block kk, from its smem tile;
                // every worker warp updates exactly the rows its own dots read, so no CTA-wide barrier is needed
                auto worker_axpy = [&](int kk) {
                    const int li = kk & 1;
                    const int nnz = (int)misc[40 + 2 * li];
                    if (nnz == 0) return;
                    const double K = misc[41 + 2 * li];
                    const uint8_t* tile = smem + L.off_tile + stage_of(kk) * L.tile_bytes;
                    for (int r = r_lo + lane; r < r_hi; r += 32) {
                        double sacc = 0.0;
                        for (int q = 0; q < nnz; ++q) {
                            const uint32_t byte = tile[nz_idx[li * kMaxB + q] * R + r];
                            sacc = fma(nz_db[li * kMaxB + q], dec_byte(byte, 0), sacc);
                        }
                        e_s[r] -= (sacc - K);
                    }
                    __syncwarp();
                };

                // staged outputs of block kk -> HBM (producer warp of CTA 0; plain coalesced stores)
                auto write_out = [&](int kk) {
                    const int li = kk & 1;
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const int q = b * 32 + lane;
                        const int64_t j = (int64_t)kk * B + q;
                        if (j < p_real) {
                            beta_g[j] = out_b[li * kMaxB + q];
                            if (method != 0) delta_g[j] = out_i[li * kMaxB + q];
                            if (method == 1) vb_g[j] = out_v[li * kMaxB + q];
                        }
                    }
                };

                // ---- prologue: partial sums of block 0
                if (warp >= 1 && warp <= kWorkerWarps) worker_dot(0);
                __syncthreads();
                if (tid == 32) NGP_TICK(1);

                for (int k = 0; k < nblk; ++k) {
                    if (warp == kProducerWarp) {
                        // ================= TMA producer: tile k-2 was applied in step k-1, refill its stage =================
                        if (lane == 0 && k >= 2 && k - 2 + S_ < nblk) issue(k - 2 + S_);
                        if (t == 0 && k >= 1) write_out(k - 1);
                    } else if (warp != 0) {
                        // ================= workers =================
                        if (k >= 1) worker_axpy(k - 1);
                        if (tid == 32) NGP_TICK(2);
                        if (k + 1 < nblk) worker_dot(k + 1);
                        if (tid == 32) NGP_TICK(1);
                    } else {
                        // ================= chain warp =================
                        const int stg = stage_of(k);
                        tc = clock64();
                        mbar_wait(&mbar[stg], parity_of(k));
                        NGP_TICK(0);
                        const int32_t* gram = reinterpret_cast<const int32_t*>(smem + L.off_blk + stg * L.blk_bytes);
                        const double* cst = reinterpret_cast<const double*>(smem + L.off_blk + stg * L.blk_bytes + B * B * 4);
                        const int slot = (int)((rk + (unsigned)k) & (kSlots - 1));
                        const long long* acc = sy->acc + (int64_t)slot * (kMaxB + 1) * kAccStride;
                        const int lp = (k & 1) ^ 1;                  // nz list of block k-1
                        const int npend = (k >= 1) ? (int)misc[40 + 2 * lp] : 0;
                        double r[NB], bold[NB], dd[NB], cs[NB], bnew[NB];
                        bool inc[NB];
                        long long curv[NB];
                        // cross-Gram rows of the markers changed in block k-1: issue the loads BEFORE polling so that
                        // the two L2 round trips overlap
                        constexpr int MAXP = 4;
                        int gxv[MAXP][NB];
                        const int32_t* gx = gramx_g + (int64_t)k * B * B;
#pragma unroll
                        for (int q0 = 0; q0 < MAXP; ++q0)
#pragma unroll
                            for (int b = 0; b < NB; ++b)
                                gxv[q0][b] = (q0 < npend) ? __ldg(gx + nz_idx[lp * kMaxB + q0] * B + b * 32 + lane) : 0;
                        {   // every lane polls its own accumulators until all T CTAs have added theirs
                            bool done;
                            do {
                                done = true;
#pragma unroll
                                for (int b = 0; b < NB; ++b) {
                                    const int q = b * 32 + lane;
                                    curv[b] = ld_relaxed_s64(acc + q * kAccStride);
                                    done = done && (((curv[b] - prev[slot * (kMaxB + 1) + q]) & 0xFF) == (long long)P.T);
                                }
                            } while (!__all_sync(0xffffffffu, done));
                        }
                        NGP_TICK(3);
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int q = b * 32 + lane;
                            long long* pv = prev + slot * (kMaxB + 1) + q;
                            const double A = (double)((curv[b] - *pv - (long long)P.T) >> kCntBits) * fx_inv;
                            *pv = curv[b];
                            cs[b] = cst[F_CS * B + q];
                            r[b] = 4.0 * (A - Stot) - cst[F_MEAN * B + q] * Stot;     // x_j'e for the e of one block ago
                            bold[b] = cst[F_BOLD * B + q];
                            dd[b] = cst[F_D * B + q];
                            bnew[b] = 0.0; inc[b] = false;
                        }
                        // the partial sums were formed before block k-1 was applied to e: restore exactly with the cross Gram
#pragma unroll
                        for (int q0 = 0; q0 < MAXP; ++q0) {
                            if (q0 < npend) {
                                const double dbf = 0.25 * nz_db[lp * kMaxB + q0];
                                const double csf = nz_cs[lp * kMaxB + q0];
#pragma unroll
                                for (int b = 0; b < NB; ++b) r[b] = fma(-((double)gxv[q0][b] - csf * cs[b] * inv_n), dbf, r[b]);
                            }
                        }
                        for (int q0 = MAXP; q0 < npend; ++q0) {
                            const int f = nz_idx[lp * kMaxB + q0];
                            const double dbf = 0.25 * nz_db[lp * kMaxB + q0];
                            const double csf = nz_cs[lp * kMaxB + q0];
#pragma unroll
                            for (int b = 0; b < NB; ++b) {
                                const double gc = (double)__ldg(gx + f * B + b * 32 + lane) - csf * cs[b] * inv_n;
                                r[b] = fma(-gc, dbf, r[b]);
                            }
                        }
                        int nnz = 0;
                        double K = 0.0;
                        const int li = k & 1;
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int q = b * 32 + lane;
                            const double cA = cst[F_A * B + q], cB = cst[F_B * B + q], cT = cst[F_T * B + q];
                            const double cC = cst[F_C * B + q], cQ = cst[F_QSZ * B + q];
                            const double mq = cst[F_MEAN * B + q];
                            int start = 0;
                            while (start < 32) {
                                pf[7]++;
                                const double rr = fma(dd[b], bold[b], r[b]);        // add-back fused: x'(e + x b) = x'e + d b
                                const double dl = fma(cB, rr * rr, cA);
                                const bool in = dl < cT;                            // NaN -> excluded, like rand() < NaN
                                const double bn = in ? fma(rr, cC, cQ) : 0.0;
                                const double db = bn - bold[b];
                                const bool act = lane >= start;
                                const unsigned m = __ballot_sync(0xffffffffu, act && (db != 0.0));
                                const int f = m ? (__ffs(m) - 1) : 32;
                                if (act && lane <= f) { bnew[b] = bn; inc[b] = in; }
                                if (f == 32) break;
                                const double dbf = __shfl_sync(0xffffffffu, db, f);
                                const double csf = __shfl_sync(0xffffffffu, cs[b], f);
                                const double mf = __shfl_sync(0xffffffffu, mq, f);
                                const int32_t* grow = gram + (b * 32 + f) * B;
#pragma unroll
                                for (int bb = 0; bb < NB; ++bb) {
                                    if (bb >= b) {
                                        const double gc = (double)grow[bb * 32 + lane] - csf * cs[bb] * inv_n;
                                        if (bb > b || lane > f) r[bb] = fma(-gc, dbf, r[bb]);
                                    }
                                }
                                if (lane == 0) {
                                    nz_idx[li * kMaxB + nnz] = b * 32 + f;
                                    nz_db[li * kMaxB + nnz] = 4.0 * dbf;
                                    nz_cs[li * kMaxB + nnz] = csf;
                                }
                                K = fma(4.0 * dbf, 1.0 + 0.25 * mf, K);
                                ++nnz;
                                start = f + 1;
                            }
                        }
                        if (lane == 0) { misc[40 + 2 * li] = (double)nnz; misc[41 + 2 * li] = K; }
                        // outputs of the block: staged in smem, written to HBM by the producer warp of CTA 0 in the next step
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int q = b * 32 + lane;
                            acc_bb = fma(bnew[b], bnew[b], acc_bb);
                            if (method != 0) acc_n += inc[b] ? 1.0 : 0.0;
                            if (t == 0) {
                                out_b[li * kMaxB + q] = bnew[b];
                                out_i[li * kMaxB + q] = inc[b] ? 1 : 0;
                                if (method == 1)                                    // functions.jl:182,186
                                    out_v[li * kMaxB + q] = inc[b] ? (sdf + bnew[b] * bnew[b]) / cst[F_CHI * B + q] : 0.0;
                            }
                        }
                        pf[6] += nnz;
                        NGP_TICK(4);
                    }
                    __syncthreads();
                    if (tid == 0) NGP_TICK(5);
                    if (tid == 32) NGP_TICK(10);
                    // tile k-1 is no longer needed (its axpy ran in this step): refill its stage
                }
                // ---- epilogue: apply the last block
                if (warp >= 1 && warp <= kWorkerWarps) worker_axpy(nblk - 1);
                if (warp == kProducerWarp && t == 0) write_out(nblk - 1);
                __syncthreads();
                gk = (gk + (unsigned)nblk) % (2u * (unsigned)S_);
                rk += (unsigned)nblk;
            } else {
                // ============================ literal per-marker sweep ============================
                const int ngrp = R >> 3;
                for (int64_t j = 0; j < S.p; ++j, ++rk) {
                    const uint8_t* col = S.geno + ((int64_t)t * S.p_pad + j) * R;
                    const int slot = (int)(rk & (kSlots - 1));
                    long long* acc = sy->acc + (int64_t)slot * (kMaxB + 1) * kAccStride;
                    uint2 w0 = make_uint2(0xF0F0F0F0u, 0xF0F0F0F0u);
                    double a = 0.0, dummy = 0.0;
                    for (int g = tid; g < ngrp; g += kThreads) {
                        const uint2 w = __ldg(reinterpret_cast<const uint2*>(col + 8 * g));
                        if (g == tid) w0 = w;
                        const double* ep = e_s + 8 * g;
                        a = fma(dec_byte(w.x, 0), ep[0], a); a = fma(dec_byte(w.x, 1), ep[1], a);
                        a = fma(dec_byte(w.x, 2), ep[2], a); a = fma(dec_byte(w.x, 3), ep[3], a);
                        a = fma(dec_byte(w.y, 0), ep[4], a); a = fma(dec_byte(w.y, 1), ep[5], a);
                        a = fma(dec_byte(w.y, 2), ep[6], a); a = fma(dec_byte(w.y, 3), ep[7], a);
                    }
                    block_sum2(a, dummy, misc);
                    if (tid == 0) {
                        const double xs = a * fx_scale;
                        if (!(fabs(xs) < 9007199254740992.0)) atomicOr(&sy->err, 1);
                        red_add_u64(acc, (__double2ll_rn(xs) << kCntBits) + 1);
                    }
                    if (warp == 0) {
                        const double* c = reinterpret_cast<const double*>(S.blk + (j / B) * (int64_t)blk_bytes(B) + B * B * 4) + (j % B);
                        const double cA = __ldcg(c + F_A * B), cB = __ldcg(c + F_B * B), cT = __ldcg(c + F_T * B);
                        const double cC = __ldcg(c + F_C * B), cQ = __ldcg(c + F_QSZ * B), d = __ldcg(c + F_D * B);
                        const double bold = __ldcg(c + F_BOLD * B), mean = __ldcg(c + F_MEAN * B), chi = __ldcg(c + F_CHI * B);
                        long long* pv = prev + slot * (kMaxB + 1);
                        long long cur;
                        do { cur = ld_relaxed_s64(acc); } while (((cur - *pv) & 0xFF) != (long long)P.T);
                        const double A = (double)((cur - *pv - (long long)P.T) >> kCntBits) * fx_inv;
                        __syncwarp();
                        if (lane == 0) *pv = cur;
                        const double r = 4.0 * (A - Stot) - mean * Stot;
                        const double rr = fma(d, bold, r);
                        const double dl = fma(cB, rr * rr, cA);
                        const bool in = dl < cT;
                        const double bn = in ? fma(rr, cC, cQ) : 0.0;
                        if (lane == 0) {
                            misc[40] = bn - bold; misc[41] = mean;
                            acc_bb = fma(bn, bn, acc_bb);
                            if (S.method != 0) acc_n += in ? 1.0 : 0.0;
                            if (t == 0) {
                                S.beta[j] = bn;
                                if (S.method != 0) S.delta[j] = in ? 1 : 0;
                                if (S.method == 1) S.varBeta[j] = in ? (S.scale * S.df + bn * bn) / chi : 0.0;
                            }
                        }
                    }
                    __syncthreads();
                    const double db = misc[40];
                    if (db != 0.0) {
                        const double db4 = 4.0 * db, K = db4 * (1.0 + 0.25 * misc[41]);
                        for (int g = tid; g < ngrp; g += kThreads) {
                            const uint2 w = (g == tid) ? w0 : __ldg(reinterpret_cast<const uint2*>(col + 8 * g));   // first pass from registers
                            double* ep = e_s + 8 * g;
                            const int lim = nrow - 8 * g;
                            if (lim > 0) ep[0] -= fma(db4, dec_byte(w.x, 0), -K);
                            if (lim > 1) ep[1] -= fma(db4, dec_byte(w.x, 1), -K);
                            if (lim > 2) ep[2] -= fma(db4, dec_byte(w.x, 2), -K);
                            if (lim > 3) ep[3] -= fma(db4, dec_byte(w.x, 3), -K);
                            if (lim > 4) ep[4] -= fma(db4, dec_byte(w.y, 0), -K);
                            if (lim > 5) ep[5] -= fma(db4, dec_byte(w.y, 1), -K);
                            if (lim > 6) ep[6] -= fma(db4, dec_byte(w.y, 2), -K);
                            if (lim > 7) ep[7] -= fma(db4, dec_byte(w.y, 3), -K);
                        }
                    }
                    __syncthreads();
                }
                // (only lane 0 accumulated acc_bb / acc_n in the literal path; the other lanes hold 0)
            }
            if (tid == 0) { tc = clock64(); }

            // ------------------------------------------------------------------ phase 3
            const bool regional = (S.method == 0 && S.n_regions > 1);
            if (warp == 0 && !regional) {
                const double bb = warp_sum(acc_bb);
                const double nl = warp_sum(acc_n);
                if (t == 0 && lane == 0) {
                    Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                    if (S.method == 0) {            // one region: functions.jl:135
                        const double chi2 = P.replay ? S.rp_chi2b[rp_row * S.nvar] : stream_chisq(st, P_CHI2_B, 0, 0, S.df + (double)S.p);
                        S.varBeta[0] = (S.scale * S.df + bb) / chi2;
                    } else if (S.method == 2) {     // functions.jl:230
                        const double chi2 = P.replay ? S.rp_chi2b[rp_row * S.nvar] : stream_chisq(st, P_CHI2_B, 0, 0, S.df + nl);
                        S.varBeta[0] = (S.scale * S.df + bb) / chi2;
                    }
                    if (S.method != 0 && S.est_pi) {    // functions.jl:189-193, 231-235, 531-533
                        const double piIn = P.replay ? S.rp_betapi[rp_row] : stream_beta(st, nl + 1.0, (double)S.p - nl + 1.0);
                        S.pi[0] = 1.0 - piIn; S.pi[1] = piIn;
                        S.pi[2] = log(1.0 - piIn); S.pi[3] = log(piIn);
                    }
                }
            }
            if (regional || P.accumulate) {
                // beta / delta of this sweep were written by CTA 0 (plain stores, off the critical path): one grid
                // barrier, then the posterior sums and the region variances are spread over the whole grid
                __syncthreads();
                gs.nbar++;
                if (tid == 0) gs.arrive();
                if (warp == 0) gs.wait_warp();
                __syncthreads();
            }
            if (P.accumulate) {
                for (int64_t j = (int64_t)t * kThreads + tid; j < S.p; j += (int64_t)P.T * kThreads) {
                    const double bj = __ldcg(&S.beta[j]);
                    S.sum_beta[j] += bj;
                    S.sum_beta2[j] = fma(bj, bj, S.sum_beta2[j]);
                    S.sum_delta[j] += (S.method == 0) ? 1.0 : (double)__ldcg(&S.delta[j]);
                }
            }
            if (regional) {
                Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                for (int64_t rg = (int64_t)t * kWarps + warp; rg < S.n_regions; rg += (int64_t)P.T * kWarps) {
                    const int64_t j0 = S.region_off[rg], j1 = S.region_off[rg + 1];
                    double bb = 0.0;
                    for (int64_t j = j0 + lane; j < j1; j += 32) { const double bj = __ldcg(&S.beta[j]); bb = fma(bj, bj, bb); }
                    bb = warp_sum(bb);
                    if (lane == 0) {
                        const double chi2 = P.replay ? S.rp_chi2b[rp_row * S.nvar + rg]
                                                     : stream_chisq(st, P_CHI2_B, (uint32_t)rg, 0, S.df + (double)(j1 - j0));
                        S.varBeta[rg] = (S.scale * S.df + bb) / chi2;              // functions.jl:135
                    }
                }
            }
            if (tid == 0) NGP_TICK(11);
        }   // sets

        if (t == 0 && tid == 0) {
            P.sc->mu = mu; P.sc->varE = varE; P.sc->iter = iter0 + it + 1;
            if (P.accumulate) P.sc->n_post += 1;
        }
    }   // iterations

    __syncthreads();
    for (int r = tid; r < nrow; r += kThreads) P.e[row0 + r] = e_s[r];
    if (tid == 0) {
        const int own[] = {0, 3, 4, 5, 6, 7, 8, 9, 11};
        for (int i : own) sy->prof[t * kProf + i] = pf[i];
    }
    if (tid == 32) {
        const int own[] = {1, 2, 10, 12, 13, 14};
        for (int i : own) sy->prof[t * kProf + i] = pf[i];
    }
#undef NGP_TICK
}

}  // namespace ngp
