// ngp_sweep.cuh — the persistent cooperative Gibbs kernel of libngp (sm_100a).
//
// One launch runs n_iter whole iterations of samplers.runSampler!'s loop body
// (/root/reference/src/samplers.jl:32-53) on device:
//   phase 0  e'e and 1'e  -> varE (functions.jl:523-525), intercept (functions.jl:39-47)
//   phase 1  per-marker constants + variates for every marker set (grid-parallel)
//   phase 2  the marker sweep (functions.jl:118-137 / 157-195 / 197-236)
//   phase 3  variance components and pi (functions.jl:509-511, 531-533), posterior sums
//
// Phase 2, blocked exact sweep with D blocks of look-ahead.  The grid is Tw worker CTAs + 1 chain CTA.
//   worker CTA t : owns row panel t of the genotypes (TMA tile ring in smem, each column read from HBM once per sweep)
//                  and of the residual e.  Roles inside the CTA run decoupled, handing work over through mbarriers:
//                  * updater warps hold e in REGISTERS (4 consecutive rows per thread), apply the changed effects of
//                    block ja,  e -= dbeta (g - mean),  from one 32-bit tile word per changed marker, and write a new
//                    VERSION of the fixed-point residual (8 signed byte limbs per row, B-operand order of the MMA);
//                  * dot warps each own whole blocks: for block j the warp reads version j-D-1 of the limbs and forms
//                    A_j[q] = sum_i g_iq e_i  as an exact integer dot on the INT8 tensor cores (mma.sync.m16n8k32:
//                    A = 16 markers x 32 rows of codes, B = 32 rows x 8 limbs), combines the limbs to int64 and RED-adds
//                    (value<<8)+1 into a global accumulator: integer adds are associative (bit-reproducible) and the low
//                    byte counts arrivals, so an accumulator is its own barrier.  8 blocks are in flight per CTA;
//                  * one lane drives the TMA tile ring, one warp receives the changed-effect lists.
//   chain CTA    : prep warps poll the accumulators of block m, r = A/2^s - mean*1'e, and restore exactly the effect of
//                  the blocks m-D..m-2 that the dots have not seen:  r -= Gc[a][q] dbeta_a  with the integer cross-Gram
//                  (precomputed once): distances DN+1..D from HBM/L2, distances 2..DN from the TMA'd block record.
//                  The chain warp applies distance 1 and the block's own Gram, runs the B dependent scalar updates
//                  (mixture priors: 32 lanes evaluate 32 markers speculatively and serialise only on markers whose effect
//                  actually changes) and publishes the changed effects {q, dbeta, dbeta*mean} to the worker CTAs through a
//                  global ring of 8-byte {payload, sequence} words (no fence on either side).
// The "literal" variant does dot / reduce / draw / axpy per marker with one grid-wide reduction per marker (the
// north-star baseline whose sync cost we report).
#pragma once
#include "ngp_device.cuh"
#include "ngp_small_la.cuh"

namespace ngp {

template <int B>
struct NzListT {                // changed effects of one block
    int nnz, pad;
    int idx[B];
    double db[B];               // dbeta
    double aux[B];              // worker side: dbeta * mean ; chain side: column sum of the marker
};

struct SmemLayout {
    // worker CTA
    int w_e, w_limb, w_tile, w_nz, w_vbuf, w_bar;
    // chain CTA
    int c_rec, c_rbase, c_nz, c_prev, c_bar, c_outb, c_outi;
    int misc, total;
    int tile_bytes, rec_bytes, limb_bytes, nz_bytes;
};

__host__ __device__ inline SmemLayout smem_layout(int R, int B, int NT, int DN, int NR, int NV, int store2 = 0)
{
    SmemLayout L;
    L.tile_bytes = store2 ? (B * R) >> 2 : B * R;
    L.rec_bytes = (1 + DN) * gram_bytes(B) + consts_bytes(B);
    L.limb_bytes = R * 8;
    L.nz_bytes = 8 + 20 * B;
    int o = 0;
    L.misc = o;    o += 96 * 8;
    const int base = o;
    // worker
    L.w_e = o;     o += R * 8;
    L.w_limb = o;  o += NV * L.limb_bytes;
    L.w_nz = o;    o += kNzSmem * L.nz_bytes;
    L.w_vbuf = o;  o += kNzRing * 4;
    L.w_bar = o;   o += (2 * NT + 2 * kNzSmem + 2 * kNzRing) * 8;
    o = (o + 127) & ~127;
    L.w_tile = o;  o += NT * L.tile_bytes;
    const int wtot = o;
    // chain
    o = base;
    L.c_bar = o;   o += (2 * kRecStages + 3 * kPrepWarps + kNzRing) * 8;
    L.c_outb = o;  o += kPrepWarps * 64 * 8;
    L.c_outi = o;  o += kPrepWarps * 64 * 4;
    L.c_rbase = o; o += kPrepWarps * B * 8;
    L.c_prev = o;  o += kSlots * B * 8;
    L.c_nz = o;    o += kNzRing * L.nz_bytes;
    o = (o + 127) & ~127;
    L.c_rec = o;   o += NR * L.rec_bytes;
    L.total = o > wtot ? o : wtot;
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

// Grid barrier.  Sharded chain (n_ranks > 1): an arrival is pushed to the counter of every rank (release at system scope
// covers this CTA's earlier peer stores), every rank polls its own counter; the counter is monotonic across launches.
// Row-sharded chain: every wait for another rank gives up after kShardTimeoutNs (or as soon as any wait of this rank has given up) and
// raises bit 8 of the error word; the launch then runs to its end on garbage and the host reports NGP_ECUDA-free failure NGP_ETIMEOUT.
constexpr unsigned long long kShardTimeoutNs = 4000000000ull;
__device__ __forceinline__ bool shard_wait_expired(int* err, unsigned long long& t0)
{
    if (*reinterpret_cast<volatile int*>(err) & 8) return true;
    const unsigned long long now = global_ns();
    if (t0 == 0) { t0 = now; return false; }
    if (now - t0 > kShardTimeoutNs) { atomicOr(err, 8); return true; }
    return false;
}

struct GridSync {
    int* err;
    unsigned long long* counter;
    unsigned long long nbar;     // barriers this CTA has taken part in (this launch)
    unsigned int T;              // CTAs of all ranks
    unsigned long long base;     // arrivals before this launch
    int n_ranks;
    __device__ __forceinline__ void arrive(const Params& P)        // one thread, after a __syncthreads
    {
        if (n_ranks > 1) {
            for (int r = 0; r < n_ranks; ++r) arrive_release_sys(&P.peer[r]->counter);
        } else arrive_release(counter);
    }
    __device__ __forceinline__ void wait_warp()     // whole warp polls
    {
        const unsigned long long target = base + nbar * (unsigned long long)T;
        if (n_ranks > 1) {
            // bounded: a rank that never arrives (its launch failed, its GPU is gone) must not hang the others for ever
            unsigned spins = 0;
            unsigned long long t0 = 0;
            while ((unsigned long long)ld_relaxed_s64_sys(reinterpret_cast<const long long*>(counter)) < target) {
                if ((++spins & 0x3ffu) == 0u && shard_wait_expired(err, t0)) break;
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
        } else {
            while ((unsigned long long)ld_relaxed_s64(reinterpret_cast<const long long*>(counter)) < target) { }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");      // acquire once, not one CCTL.IVALL per poll
        }
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
    double* c = S.consts + (j / B) * (int64_t)(kNF * B) + (j % B);
    double fA = 0.0, fB = 0.0, fT = -INFINITY, fC = 0.0, fQSZ = 0.0, fD = 0.0, fBOLD = 0.0, fMEAN = 0.0, fCS = 0.0, fCHI = 1.0;
    if (j < S.p) {
        const double d = S.d[j];
        const double bold = __ldcg(&S.beta[j]);
        const double iVarE = 1.0 / varE;
        const double l0 = S.lhs0 ? S.lhs0[j] : 0.0;
        const double r0 = S.rhs0 ? S.rhs0[j] : 0.0;
        Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)(S.group_k ? S.stream_set : sidx)};
        // tuple (method 4): column j = locus j / k, breed j % k; the oracle addresses its normal as (locus, component)
        const double z = P.replay ? S.rp_z[rp_row * S.p + j]
                                  : (S.group_k ? stream_normal(st, P_Z, (uint32_t)(j / S.group_k), 0, (uint32_t)(j % S.group_k)) : stream_normal(st, P_Z, (uint32_t)j));
        double vb;
        if (S.method == 4) vb = 1.0;
        else if (S.method == 0) vb = __ldcg(&S.varBeta[S.region_of ? S.region_of[j] : 0]);
        else if (S.method == 1) vb = __ldcg(&S.varBeta[j]);
        else vb = __ldcg(&S.varBeta[0]);
        const double lhs = d * iVarE + l0 + 1.0 / vb;          // 1/0 -> Inf (BayesB quirk, functions.jl:186)
        const double ilhs = 1.0 / lhs;
        fC = iVarE * ilhs;
        fQSZ = sqrt(ilhs) * z + ((S.method == 2) ? 0.0 : r0 * ilhs);
        fD = d; fBOLD = bold; fMEAN = S.mean[j]; fCS = (double)S.colsum[j];
        if (S.method == 1 || S.method == 2) {
            const double u = P.replay ? S.rp_u[rp_row * S.p + j] : stream_uniform(st, P_U, (uint32_t)j);
            const double v0 = d * varE;
            const double v1 = (d * d) * vb + v0;
            fA = 0.5 * (log(v1) - log(v0)) + (__ldcg(&S.pi[2]) - __ldcg(&S.pi[3]));
            fB = 0.5 * (1.0 / v1 - 1.0 / v0);
            fT = log(1.0 / u - 1.0);
            if (S.method == 1)
                fCHI = P.replay ? S.rp_chi2b[rp_row * S.nvar + j] : stream_chisq(st, P_CHI2_B, (uint32_t)j, 0, S.df + 1.0);
        } else if (S.method == 0) {
            fT = INFINITY;   // BayesPR: always "included"
        } else if (S.method == 4) {
            fC = 0.0; fQSZ = z; fT = (double)(S.region_of ? S.region_of[j] : 0);     // the joint solve runs in the chain warp: variate + region index
        } else {
            // BayesR: F_QSZ carries the normal variate, the class algebra runs in the sweep.  The blocked sweep (<= 4 classes) also finds the
            // uniforms of the cumulative comparisons (a fresh one per comparison, functions.jl:261) here: F_A, F_B, F_T, F_C = u_0 .. u_3
            fC = 0.0; fQSZ = z;
            if (S.n_class <= 4 && S.method == 3) {
                double uv[4] = {0.0, 0.0, 0.0, 0.0};
                for (int v = 0; v < S.n_class; ++v)
                    uv[v] = P.replay ? S.rp_u[(rp_row * S.p + j) * S.n_class + v] : stream_uniform(st, P_U, (uint32_t)j, 0, (uint32_t)v);
                fA = uv[0]; fB = uv[1]; fT = uv[2]; fC = uv[3];
            }
        }
    }
    c[F_A * B] = fA; c[F_B * B] = fB; c[F_T * B] = fT; c[F_C * B] = fC; c[F_QSZ * B] = fQSZ;
    c[F_D * B] = fD; c[F_BOLD * B] = fBOLD; c[F_MEAN * B] = fMEAN; c[F_CS * B] = fCS; c[F_CHI * B] = fCHI;
}

// ----------------------------------------------------------------------------- fixed-point residual limbs
// Rows 4*rg .. 4*rg+3 of e -> 8 signed byte limbs per row in the B-fragment order of mma.m16n8k32: inside the 256-byte
// record of a 32-row chunk, word [half][n][tt] (XOR-swizzled by half) holds limb n of rows 32c + 16*half + 4*tt + (0..3).
// e_fx = rint(e * 2^sh) = sum_n d_n 256^n (mod 2^64) with d_n in [-128,127]: the balanced digits are the bytes of
// (e_fx + 0x0080808080808080) ^ 0x0080808080808080.
__device__ __forceinline__ void quantise4(const double (&e)[4], uint32_t* limb, int rg, double fx_scale)
{
    const int c = rg >> 3, half = (rg >> 2) & 1, tt = rg & 3;
    const unsigned long long bias = 0x0080808080808080ull;
    const unsigned long long x0 = ((unsigned long long)__double2ll_rn(e[0] * fx_scale) + bias) ^ bias;
    const unsigned long long x1 = ((unsigned long long)__double2ll_rn(e[1] * fx_scale) + bias) ^ bias;
    const unsigned long long x2 = ((unsigned long long)__double2ll_rn(e[2] * fx_scale) + bias) ^ bias;
    const unsigned long long x3 = ((unsigned long long)__double2ll_rn(e[3] * fx_scale) + bias) ^ bias;
    uint32_t* dst = limb + c * 64 + half * 32;
    const int sw = half << 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {          // low / high 32 bits: limbs 4h .. 4h+3
        const uint32_t a0 = (uint32_t)(x0 >> (32 * h)), a1 = (uint32_t)(x1 >> (32 * h));
        const uint32_t a2 = (uint32_t)(x2 >> (32 * h)), a3 = (uint32_t)(x3 >> (32 * h));
        const uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a2, a3, 0x5140);
        const uint32_t t2 = __byte_perm(a0, a1, 0x7362), t3 = __byte_perm(a2, a3, 0x7362);
        dst[(((4 * h + 0) << 2) | tt) ^ sw] = __byte_perm(t0, t1, 0x5410);
        dst[(((4 * h + 1) << 2) | tt) ^ sw] = __byte_perm(t0, t1, 0x7632);
        dst[(((4 * h + 2) << 2) | tt) ^ sw] = __byte_perm(t2, t3, 0x5410);
        dst[(((4 * h + 3) << 2) | tt) ^ sw] = __byte_perm(t2, t3, 0x7632);
    }
}

// ----------------------------------------------------------------------------- tuple step of the chain warp
// One step (SM = 32 NS columns) of a tuple of K interleaved marker sets: the K lanes gl .. gl+K-1 of a slot hold the K breeds of one
// locus.  Per locus: complete the add-back with the cross terms  sum_{a != b} Gc[a][b] beta_old_a  (rr already holds x'e + d beta_old),
// gather the K values, solve  C = (M'M/varE + invB_r)^-1,  beta = C rhs + chol(C) z  in every lane (uniform inputs, no divergence),
// then restore the later columns of the step from the Gram rows of the K changed columns.  Every column becomes a list entry.
template <int B, int NS, int K, class NzList>
__device__ __forceinline__ void joint_step(double (&rr)[NS], const double (&bold)[NS], const double (&cs)[NS], const double (&cQ)[NS],
                                           const double (&cT)[NS], double (&bnew)[NS], bool (&inc)[NS], const int32_t* (&gram)[NS],
                                           const int (&kb)[NS], const int (&qb)[NS], NzList* cnz, unsigned g0, int lane, double inv_n,
                                           double iVarE, const double* jinvB, SyncArea* sy)
{
    const unsigned FULL = 0xffffffffu;
    double mydb[NS];
    int cur_rg = -1;
    double invB[K * K];
#pragma unroll
    for (int x = 0; x < K * K; ++x) invB[x] = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        mydb[i] = 0.0;
#pragma unroll 1
        for (int gl = 0; gl < 32; gl += K) {
            const int m0 = 32 * i + gl;                      // first column of the locus within the step
            const int ka = m0 / B, qa0 = m0 % B;             // its block of the step and column inside the block (K divides B: no straddling)
            // region of the locus -> inverse covariance (uniform across the warp)
            const double rgd = __shfl_sync(FULL, cT[i], gl);
            if (!(rgd >= 0.0)) continue;                     // padding columns past the last locus (F_T = -inf): nothing to draw
            const int rg = (int)rgd;
            if (rg != cur_rg) {
#pragma unroll
                for (int x = 0; x < K * K; ++x) invB[x] = __ldcg(&jinvB[(size_t)rg * K * K + x]);
                cur_rg = rg;
            }
            double bo[K], csg[K], zg[K], rhs[K];
#pragma unroll
            for (int a = 0; a < K; ++a) {
                bo[a] = __shfl_sync(FULL, bold[i], gl + a);
                csg[a] = __shfl_sync(FULL, cs[i], gl + a);
                zg[a] = __shfl_sync(FULL, cQ[i], gl + a);
            }
            // centred Gram of the locus' K columns, M_j'M_j: the distance-0 Gram of their block, read through the record pointer of the
            // lane that owns the locus' first column (the records of a step's blocks are different stages of the ring)
            const unsigned long long owner_ptr = __shfl_sync(FULL, (unsigned long long)(uintptr_t)gram[i], gl);
            const int32_t* gk = reinterpret_cast<const int32_t*>((uintptr_t)owner_ptr);
            double MtM[K * K];
#pragma unroll
            for (int a = 0; a < K; ++a)
#pragma unroll
                for (int b = 0; b < K; ++b)
                    MtM[a * K + b] = (double)gk[(qa0 + a) * B + (qa0 + b)] - csg[a] * csg[b] * inv_n;
            // x_b'(e + M beta_old) = rr_b (own add-back included) + sum_{a != b} MtM[a][b] beta_old_a
#pragma unroll
            for (int b = 0; b < K; ++b) {
                double v = __shfl_sync(FULL, rr[i], gl + b);
#pragma unroll
                for (int a = 0; a < K; ++a) if (a != b) v = fma(MtM[a * K + b], bo[a], v);
                rhs[b] = v * iVarE;                                                        // RHS (functions.jl:146)
            }
            double LHS[K * K], C[K * K], Lc[K * K], bn[K], db[K];
#pragma unroll
            for (int x = 0; x < K * K; ++x) LHS[x] = MtM[x] * iVarE + invB[x];
            bool ok;
            if constexpr (K == 2) {
                // 2 x 2 in closed form (the generic route costs eight fp64 divisions / square roots on the dependent path)
                const double det = fma(LHS[0], LHS[3], -(LHS[1] * LHS[2]));
                const double idet = 1.0 / det;
                C[0] = LHS[3] * idet; C[1] = -LHS[1] * idet; C[2] = C[1]; C[3] = LHS[0] * idet;      // functions.jl:147
                Lc[0] = sqrt(C[0]); Lc[1] = 0.0; Lc[2] = C[2] / Lc[0];
                const double s22 = fma(-Lc[2], Lc[2], C[3]);
                Lc[3] = sqrt(s22);
                ok = (LHS[0] > 0.0) && (det > 0.0) && (s22 > 0.0);
            } else {
                ok = jt_inv_spd<K>(LHS, C) && jt_chol<K>(C, Lc);                           // functions.jl:147
            }
            if (!ok && lane == 0) atomicOr(&sy->err, 2);
#pragma unroll
            for (int a = 0; a < K; ++a) {
                double mval = 0.0;
#pragma unroll
                for (int b = 0; b < K; ++b) mval += C[a * K + b] * rhs[b];                    // functions.jl:148
                double sdraw = mval;
#pragma unroll
                for (int b = 0; b <= a; ++b) sdraw += Lc[a * K + b] * zg[b];                  // functions.jl:149
                bn[a] = ok ? sdraw : bo[a];
                db[a] = bn[a] - bo[a];
            }
#pragma unroll
            for (int a = 0; a < K; ++a)
                if (lane == gl + a) { bnew[i] = bn[a]; inc[i] = true; mydb[i] = db[a]; }
            // restore the later columns of the step (everything after the locus)
            const int mlast = m0 + K - 1;
#pragma unroll
            for (int i2 = i; i2 < NS; ++i2)
                if (32 * i2 + lane > mlast) {
                    double acc = 0.0;
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        const double gc = (double)gram[i2][(kb[i2] - ka) * B * B + (qa0 + a) * B + qb[i2]] - csg[a] * cs[i2] * inv_n;
                        acc = fma(gc, db[a], acc);
                    }
                    rr[i2] -= acc;
                }
        }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        NzList& ml = cnz[(g0 + (unsigned)kb[i]) & (kNzRing - 1)];
        ml.idx[qb[i]] = qb[i]; ml.db[qb[i]] = mydb[i]; ml.aux[qb[i]] = cs[i];
    }
}

// ----------------------------------------------------------------------------- the kernel
// LIT: the per-marker ("literal") sweep instead of the blocked one — a separate instantiation, so that neither variant carries the other's code
// TUP: the instantiation that also sweeps a tuple of interleaved marker sets (method 4); kept apart so that the k x k algebra does not
//      weigh on the register allocation of the single-trait sweep
// SH: the blocked sweep of a row-sharded chain (system-scope REDs into every rank's accumulators, runtime arrival-count bits, rank-local
// pre-reduction).  Compiled out of the one-GPU instantiations: its code in the dot / prep warps' loops cost the plain sweep 5 % (int8) to
// 18 % (2-bit) through instruction-cache pressure alone (A/B of round 2).  The per-marker sweep (LIT) always carries it.
// SHT (with BIGR): refetch ring of ONE tile whose dots all 8 dot warps share (see the dot warps).
template <int B, bool PROF, bool DBG, bool LIT, bool TUP, bool BIGR = false, bool BR = false, bool SH = false, bool SHT = false>
__device__ __forceinline__ void gibbs_body(const Params& P, const int t)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NB = (B + 31) / 32;      // markers per lane of a prep warp (lane <-> marker b*32 + lane)
    constexpr int NS = (B == 16) ? 1 : 2;  // markers per lane of a chain step: a step covers SM = 32 NS markers (32 for blocks of 16, else 64)
    constexpr int SM = 32 * NS;
    constexpr int SB = SM / B;             // blocks per step of the chain warp (B = 16, 32: 2; B = 64: 1)
    constexpr int SBS = (SB == 1) ? 0 : (SB == 2) ? 1 : 2;
    constexpr int RBS = kPrepWarps / SB;   // super-block slots of the r_base ring
    constexpr int kHelperWarp = kFirstPrepWarp + kPrepWarps;     // chain CTA: publishes the lists and writes the outputs
    constexpr int MG = B / 16;             // 16-marker MMA groups per block
    constexpr int UG = (B == 16 || BIGR) ? kUpdGroups : 1;   // 4-row groups per updater thread: panels of up to 512 rows, or (blocks of 16 / the BIGR instantiation) 2048
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Tw = P.Tw;
    const bool is_chain = (t == Tw);
    const int R = P.R, D = P.D, DN = P.DN, NT = P.NT, NR = P.NR;
    const int NV = P.NV;
    const SmemLayout L = smem_layout(R, B, NT, DN, NR, NV, P.store2);
    // timing experiments only (results are garbage): 1 workers ignore the lists, 2 prep warps skip the accumulator poll,
    // 4 chain warp skips corrections + scalar updates, 8 workers skip the dots and the RED
    const int dbg = DBG ? P.debug : 0;
    double* misc = reinterpret_cast<double*>(smem + L.misc);     // [0..27] block_sum scratch, [32..] scalars, [44] chain progress, [48..] literal-kernel prev
    volatile unsigned* cprog = reinterpret_cast<volatile unsigned*>(misc + 44);     // chain CTA: global number of blocks whose lists are complete
    SyncArea* sy = P.sync;

    typedef NzListT<B> NzList;
    // worker views
    double* e_s = reinterpret_cast<double*>(smem + L.w_e);
    uint32_t* limb = reinterpret_cast<uint32_t*>(smem + L.w_limb);                          // [NV][R*2] words
    NzList* wnz = reinterpret_cast<NzList*>(smem + L.w_nz);
    int* vbuf = reinterpret_cast<int*>(smem + L.w_vbuf);                                    // [kNzRing] limb version holding the state after block g
    uint64_t* tile_full = reinterpret_cast<uint64_t*>(smem + L.w_bar);
    uint64_t* tile_free = tile_full + NT;
    uint64_t* nz_full = tile_free + NT;
    uint64_t* nz_free = nz_full + kNzSmem;
    uint64_t* ver_full = nz_free + kNzSmem;                                                 // [kNzRing] block g applied to e
    uint64_t* dot_done = ver_full + kNzRing;                                                // [kNzRing] dots of block g formed
    // chain views
    uint64_t* rec_full = reinterpret_cast<uint64_t*>(smem + L.c_bar);
    uint64_t* rec_free = rec_full + kRecStages;
    uint64_t* rb_full = rec_free + kRecStages;
    uint64_t* rb_free = rb_full + kPrepWarps;
    uint64_t* nzc_full = rb_free + kPrepWarps;
    uint64_t* step_done = nzc_full + kNzRing;                                               // [kPrepWarps] a 64-marker step of the chain warp
    double* out_b = reinterpret_cast<double*>(smem + L.c_outb);                             // [kPrepWarps][64] new effects of a step
    int* out_i = reinterpret_cast<int*>(smem + L.c_outi);                                   // [kPrepWarps][64] inclusion indicators
    double* rbase = reinterpret_cast<double*>(smem + L.c_rbase);                            // [kPrepWarps][B]
    long long* prev = reinterpret_cast<long long*>(smem + L.c_prev);                        // [kSlots][B]
    NzList* cnz = reinterpret_cast<NzList*>(smem + L.c_nz);

    constexpr bool kSh = LIT || SH;
    const bool sharded = kSh && P.n_ranks > 1;
    const int cb = kSh ? P.cnt_bits : kCntBits;              // arrival-count bits of the accumulators
    const long long cmask = (1LL << cb) - 1;
    const long long arr_all = kSh ? (long long)P.Tw_all : (long long)Tw;     // worker CTAs of all ranks add into every rank's accumulators
    GridSync gs{&sy->err, &sy->counter, 0ull, (unsigned)P.T_all, P.bar_base, P.n_ranks};
    const int64_t row0 = (int64_t)t * R;
    const int nrow = is_chain ? 0 : (int)max((int64_t)0, min((int64_t)R, P.n - row0));    // real rows of this panel
    const int nchunk = R >> 5;

    if (!is_chain) {
        for (int r = tid; r < R; r += kThreads) e_s[r] = (r < nrow) ? P.e[row0 + r] : 0.0;
    } else {
        // the accumulators are monotonic across launches: start from their current values
        for (int q = tid; q < kSlots * B; q += kThreads) prev[q] = sy->acc[((size_t)(q / B) * kMaxB + (q % B)) * kAccStride];
    }
    if (tid < kSlots) reinterpret_cast<long long*>(misc + 48)[tid] = sy->acc[(size_t)tid * kMaxB * kAccStride];      // literal kernel: previous accumulator values
    // weighted residuals (per-marker kernel): the second accumulator of a slot carries the UNWEIGHTED dot; lane l of warp 0 keeps
    // its previous value for slot l (kSlots == 32)
    const bool wt = LIT && P.w != nullptr;
    long long prev_u = 0;
    if (LIT && warp == 0) prev_u = sy->acc[((size_t)lane * kMaxB + 1) * kAccStride];
    if (tid == 0) *cprog = 0u;
    if (tid == 0) {
        if (!is_chain) {
            for (int s = 0; s < NT; ++s) { mbar_init(&tile_full[s], 1); mbar_init(&tile_free[s], P.refetch ? 1 : kUpdWarps); }
            for (int s = 0; s < kNzSmem; ++s) { mbar_init(&nz_full[s], 1); mbar_init(&nz_free[s], kUpdWarps); }
            for (int s = 0; s < kNzRing; ++s) { mbar_init(&ver_full[s], kUpdWarps); mbar_init(&dot_done[s], 1); }
        } else {
            for (int s = 0; s < kRecStages; ++s) { mbar_init(&rec_full[s], 1); mbar_init(&rec_free[s], 1); }
            for (int s = 0; s < kPrepWarps; ++s) { mbar_init(&rb_full[s], ((B == 16) ? 32 : 64) / B); mbar_init(&rb_free[s], 1); mbar_init(&step_done[s], 1); }
            for (int s = 0; s < kNzRing; ++s) mbar_init(&nzc_full[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    __shared__ double fx_s[2 * kMaxFxCols];          // fixed effects: [0..32) the effects, [32..64) Yi of the set being sampled
    __shared__ long long comb[SHT ? kDotWarps * B : 1];      // SHT: [8][B] partial sums of the dot warps sharing a tile (static: only that instantiation pays)
    __shared__ int ccnt_s[1];                                // SHT: warps that have stored their sums
    if constexpr (SHT) { if (tid == 0) ccnt_s[0] = 0; __syncthreads(); }
    __shared__ long long step_end_clk[64];
    __shared__ unsigned long long pub_ns[kNzRing];   // instrumented kernel: global time at which a list was published (worker CTA copy / chain CTA copy)        // instrumented kernel: clock at which the chain warp finished a step
    // cycle counters, see ngp_get_profile
    long long pf[kProf];
#pragma unroll
    for (int i = 0; i < kProf; ++i) pf[i] = 0;
    long long tc = 0;
#define NGP_TICK(i) do { if constexpr (PROF) { const long long now__ = clock64(); pf[i] += now__ - tc; tc = now__; } } while (0)

    // ring cursors: every role advances its cursors once per block, so they stay consistent across sweeps
    Ring r_ta{0, 0}, r_tp{0, 0};                  // tile ring: updater (release), TMA producer
    Ring r_nzw{0, 0}, r_nzp{0, 0};                // worker nz list ring: consumer (updaters), producer (poll warp)
    Ring r_rec{0, 0};                             // chain CTA record ring (TMA producer cursor)
    unsigned tiles_issued = 0, recs_issued = 0, nz_recv = 0;
    unsigned gblk = 0;       // blocks processed before the current sweep (all sets, all iterations of this launch): indexes every ring
    const uint32_t seq0 = P.gblk0;     // blocks swept by earlier launches: sequence number of a list word = seq0 + block number + 1
    unsigned rk = 0;         // literal kernel: running reduction count
    double mu = P.sc->mu;
    const long long iter0 = P.sc->iter;

    for (int it = 0; it < P.n_iter; ++it) {
        const uint32_t iter = (uint32_t)(iter0 + it + 1);
        const int64_t rp_row = (int64_t)iter - 1 - P.replay_base;
        if constexpr (PROF) tc = clock64();

        // ------------------------------------------------------------------ phase 0
        double ee = 0.0, se = 0.0;
        if (wt) {
            // E.str == "D": e'We (functions.jl:526-528) and 1'We (Xp of the intercept, mme.jl:136); the plain 1'e goes along for the
            // unweighted inclusion dots of BayesB/C
            double su = 0.0, dm = 0.0;
            if (!is_chain) for (int r = tid; r < R; r += kThreads) { const double x = e_s[r], wv = __ldg(&P.w[row0 + r]); ee = fma(wv * x, x, ee); se = fma(wv, x, se); su += x; }
            block_sum2(su, dm, misc);
            if (tid == 0) sy->part_u[t] = su;
        } else if (!is_chain) for (int r = tid; r < R; r += kThreads) { const double x = e_s[r]; ee = fma(x, x, ee); se += x; }
        block_sum2(ee, se, misc);
        if (tid == 0) {
            if (sharded) {
                for (int r = 0; r < P.n_ranks; ++r) { P.peer[r]->part[2 * (P.cta_off + t)] = ee; P.peer[r]->part[2 * (P.cta_off + t) + 1] = se; }
            } else { sy->part[2 * t] = ee; sy->part[2 * t + 1] = se; }
            gs.nbar++; gs.arrive(P);
        } else gs.nbar++;
        if (warp == 0) {
            gs.wait_warp();
            double a = 0.0, b = 0.0, bu = 0.0;
            const int nparts = sharded ? P.T_all : Tw;     // same order on every rank: identical sums, identical draws
            for (int c = lane; c < nparts; c += 32) { a += __ldcg(&sy->part[2 * c]); b += __ldcg(&sy->part[2 * c + 1]); if (wt) bu += __ldcg(&sy->part_u[c]); }
            a = warp_sum(a); b = warp_sum(b);
            if (wt) bu = warp_sum(bu);
            if (lane == 0) {
                double varE = P.varE_in;
                if (P.do_varE) {
                    Stream st{P.key0, P.key1, P.chain, iter, 0u};
                    const double chi2 = P.replay ? P.rp_chi2_e[rp_row] : stream_chisq(st, P_CHI2_E, 0, 0, P.df_e + (double)P.n_total);
                    varE = (P.df_e * P.scale_e + a) / chi2;                       // functions.jl:524
                }
                double dmu = 0.0;
                if (P.has_mu && P.do_mu) {                                        // functions.jl:39-47
                    Stream st{P.key0, P.key1, P.chain, iter, 0u};
                    const double zmu = P.replay ? P.rp_z_mu[rp_row] : stream_normal(st, P_Z_MU, 0);
                    const double iVarE = 1.0 / varE;
                    const double xpx = wt ? P.w_sum : (double)P.n_total;               // mme.jl:135 / :138
                    const double rhs = (b + xpx * mu) * iVarE + P.mu_rhs0;
                    const double lhs = xpx * iVarE + P.mu_lhs0;
                    const double mu_new = rhs / lhs + sqrt(1.0 / lhs) * zmu;
                    dmu = mu - mu_new;
                    mu = mu_new;
                }
                // fixed-point scale of the sweep reductions: |sum_i g_i e_i| <= 2 sqrt(n) ||e|| (Cauchy-Schwarz), 2^4 headroom
                // for growth of ||e|| inside the iteration, 8 count bits: the grid total stays below 2^55
                const double nn = (double)P.n_total;
                double M = 2.0 * sqrt(nn) * (sqrt(a) + sqrt(nn) * fabs(dmu));
                if (wt) M = 2.0 * sqrt(nn) * (sqrt(a / P.w_min) + sqrt(nn) * fabs(dmu)) * fmax(1.0, P.w_max);    // |sum g w e| <= 2 sqrt(n) max(w) ||e||
                if (!(M > 1e-300)) M = 1e-300;
                int ex; (void)frexp(M, &ex);
                int sh = 62 - P.cnt_bits - 4 - ex;
                sh = max(-1000, min(1000, sh));
                misc[32] = varE; misc[33] = dmu; misc[34] = b + (wt ? P.w_sum : nn) * dmu; misc[35] = (double)sh; misc[36] = mu;
                if (wt) misc[37] = bu + nn * dmu;
            }
        }
        __syncthreads();
        const double varE = misc[32];
        const double dmu = misc[33];
        double Stot = misc[34];                       // 1'e after the intercept update; invariant under marker updates
                                                      // (weighted residuals: 1'We, which every changed effect moves by wcs_j)
        double Stot_u = wt ? misc[37] : 0.0;          // weighted residuals: the plain 1'e (invariant under marker updates)
        const int sh = (int)misc[35];
        mu = misc[36];
        const double fx_scale = ldexp(1.0, sh), fx_inv = ldexp(1.0, -sh);
        if (dmu != 0.0) for (int r = tid; r < nrow; r += kThreads) e_s[r] += dmu;
        // ------------------------------------------------------------------ fixed effects besides the intercept (functions.jl:22-54)
        if (P.fx.n_cols > 0 && P.do_mu) {
            const FxDev& X = P.fx;
            const int64_t ldx = (int64_t)Tw * R;
            const double iVarE = 1.0 / varE;
            __syncthreads();
            if (tid < X.n_cols) fx_s[tid] = __ldcg(&X.b[tid]);
            __syncthreads();
            for (int xs = 0; xs < X.n_sets; ++xs) {
                const int c0 = X.first[xs], c1 = X.first[xs + 1];
                // ycorr += X b (functions.jl:42, :49)
                if (!is_chain)
                    for (int r = tid; r < nrow; r += kThreads) {
                        double a = 0.0;
                        for (int c = c0; c < c1; ++c) a = fma(X.data[(int64_t)c * ldx + row0 + r], fx_s[c], a);
                        e_s[r] += a;
                    }
                __syncthreads();
                // Xp * ycorr: per-CTA partials, one grid barrier, summed in CTA order by everybody
                for (int c = c0; c < c1; c += 2) {
                    double a0 = 0.0, a1 = 0.0;
                    if (!is_chain)
                        for (int r = tid; r < nrow; r += kThreads) {
                            const double ev = wt ? e_s[r] * __ldg(&P.w[row0 + r]) : e_s[r];       // Xp = (X .* w)' (mme.jl:136)
                            a0 = fma(X.data[(int64_t)c * ldx + row0 + r], ev, a0);
                            if (c + 1 < c1) a1 = fma(X.data[(int64_t)(c + 1) * ldx + row0 + r], ev, a1);
                        }
                    block_sum2(a0, a1, misc);
                    if (tid == 0) { sy->part_fx[t * kMaxFxCols + c] = a0; if (c + 1 < c1) sy->part_fx[t * kMaxFxCols + c + 1] = a1; }
                }
                __syncthreads();
                gs.nbar++;
                if (tid == 0) gs.arrive(P);
                if (warp == 0) {
                    gs.wait_warp();
                    double yi = 0.0;
                    if (c0 + lane < c1) { for (int cta = 0; cta < Tw; ++cta) yi += __ldcg(&sy->part_fx[cta * kMaxFxCols + c0 + lane]); }
                    fx_s[kMaxFxCols + (lane & (kMaxFxCols - 1))] = yi * iVarE;                       // Yi (functions.jl:25)
                    __syncwarp();
                    if (lane == 0) {
                        Stream st{P.key0, P.key1, P.chain, iter, 0u};
                        const int nc = c1 - c0;
                        const double* xpx = X.xpx + X.xoff[xs];
                        double dS = 0.0, dSw = 0.0;
                        for (int i = 0; i < nc; ++i) {
                            const double z = P.replay ? X.rp_z[rp_row * X.n_cols + c0 + i] : stream_normal(st, P_Z_MU, (uint32_t)(1 + c0 + i));
                            const double bold_i = fx_s[c0 + i];
                            double bn;
                            if (nc == 1) {                                                            // functions.jl:43-46
                                const double rhs = fx_s[kMaxFxCols] + X.rhs0[xs];
                                const double lhs = xpx[0] * iVarE + X.lhs0[xs];
                                bn = rhs / lhs + sqrt(1.0 / lhs) * z;
                            } else {                                                                  // Wang's trick, functions.jl:27-34
                                fx_s[c0 + i] = 0.0;
                                double dt = 0.0;
                                for (int kk = 0; kk < nc; ++kk) dt += xpx[i * nc + kk] * fx_s[c0 + kk];
                                const double rhsb = fx_s[kMaxFxCols + i] - dt * iVarE;
                                const double invLhsb = 1.0 / (xpx[i * nc + i] * iVarE);
                                bn = invLhsb * rhsb + sqrt(invLhsb) * z;
                            }
                            fx_s[c0 + i] = bn;
                            dS = fma(X.colsum[c0 + i], bn - bold_i, dS);
                            if (wt) dSw = fma(X.colsum_w[c0 + i], bn - bold_i, dSw);
                            if (is_chain) X.b[c0 + i] = bn;
                        }
                        // e was restored to e + X b_old before the dots: 1'e of the final e = 1'e_before - sum_c colsum_c (b_new - b_old)
                        misc[34] -= wt ? dSw : dS;
                        if (wt) misc[37] -= dS;
                    }
                }
                __syncthreads();
                // ycorr -= X b (functions.jl:47, :51)
                if (!is_chain)
                    for (int r = tid; r < nrow; r += kThreads) {
                        double a = 0.0;
                        for (int c = c0; c < c1; ++c) a = fma(X.data[(int64_t)c * ldx + row0 + r], fx_s[c], a);
                        e_s[r] -= a;
                    }
                __syncthreads();
            }
            Stot = misc[34];
            if (wt) Stot_u = misc[37];
        }
        if (tid == 0) NGP_TICK(8);

        // ------------------------------------------------------------------ phase 1
        for (int s = 0; s < P.n_sets; ++s) {
            if (!((P.set_mask >> s) & 1)) continue;
            const SetDev& S = P.sets[s];
            for (int64_t j = (int64_t)t * kThreads + tid; j < S.p_pad; j += (int64_t)(Tw + 1) * kThreads)
                prep_marker(P, S, s, j, varE, iter, rp_row);
            if (TUP && S.method == 4) {                            // invB = inv(varBeta[mSet][r]) (functions.jl:143), one thread per region
                const int k = S.group_k;
                for (int64_t rg = (int64_t)t * kThreads + tid; rg < S.n_regions; rg += (int64_t)(Tw + 1) * kThreads) {
                    double Sg[16], Iv[16];
                    for (int i = 0; i < k * k; ++i) Sg[i] = __ldcg(&S.jvar[rg * k * k + i]);
                    const bool ok = (k == 2) ? jt_inv_spd<2>(Sg, Iv) : jt_inv_spd<4>(Sg, Iv);
                    if (!ok) { atomicOr(&sy->err, 2); for (int i = 0; i < k * k; ++i) Iv[i] = 0.0; }
                    for (int i = 0; i < k * k; ++i) S.jinvB[rg * k * k + i] = Iv[i];
                }
            }
        }
        __syncthreads();
        gs.nbar++;
        if (tid == 0) gs.arrive(P);
        if (warp == 0) gs.wait_warp();
        __syncthreads();
        if (tid == 0) NGP_TICK(9);

        // ------------------------------------------------------------------ phase 2 + 3 per marker set
        for (int s = 0; s < P.n_sets; ++s) {
            if (!((P.set_mask >> s) & 1)) continue;
            const SetDev& S = P.sets[s];
            const int nblk = (int)(S.p_pad / B);
            const double inv_n = 1.0 / (double)P.n_total;
            const int method = S.method;
            // BayesR in the blocked sweep (the BR instantiation): class variances, log proportions (uniform)
            double br_varc[4] = {0.0, 0.0, 0.0, 0.0}, br_logpi[4] = {0.0, 0.0, 0.0, 0.0}, br_vcls[4] = {1.0, 1.0, 1.0, 1.0};
            const int br_nc = BR ? S.n_class : 0;
            const double br_iVarE = 1.0 / varE;
            bool br_bad = false;
            if constexpr (BR) {
                const double vb0 = __ldcg(&S.varBeta[0]);
                for (int v = 0; v < 4; ++v)
                    if (v < br_nc) { br_vcls[v] = S.v_class[v]; br_varc[v] = vb0 * S.v_class[v]; br_logpi[v] = __ldcg(&S.pi_class[br_nc + v]); }      // functions.jl:244
            }
            const int64_t p_real = S.p;
            double acc_bb = 0.0, acc_n = 0.0;          // chain warp: per-lane partials of beta'beta and nLoci
            double acc_cls = 0.0;                      // BayesR: loci assigned to class `lane`
            double rc_nloci_keep = 0.0, rc_sumS_keep = 0.0, rc_nnz_keep = 0.0;      // BayesRCpi / BayesRCplus counters of warp 0 (per-marker kernel)
            if constexpr (PROF) tc = clock64();

            if constexpr (!LIT) {
                // ======================================================================================================
                //                                  blocked exact sweep, look-ahead D
                // ======================================================================================================
                if (!is_chain) {
                    // ====================================================================== worker CTA
                    const uint8_t* gbase = S.geno + (int64_t)t * nblk * L.tile_bytes;
                    unsigned char* tiles = smem + L.w_tile;
                    const int lwords = R * 2;                                 // words per limb version
                    // residual rows of the updater threads: e_s -> registers, version 0 of the limbs
                    double er[UG][4];
                    if (warp < kUpdWarps) {
#pragma unroll
                        for (int k = 0; k < UG; ++k) {
                            const int rg = tid + k * kUpdThreads;
                            if (4 * rg < R) {
                                const double2 a01 = *reinterpret_cast<const double2*>(e_s + 4 * rg), a23 = *reinterpret_cast<const double2*>(e_s + 4 * rg + 2);
                                er[k][0] = a01.x; er[k][1] = a01.y; er[k][2] = a23.x; er[k][3] = a23.y;
                                quantise4(er[k], limb, rg, fx_scale);
                            }
                        }
                    }
                    __syncthreads();
                    if (warp < kUpdWarps) {
                        // ------------------------------------------------------------------ updater warps
                        int cur = 0;                       // limb version holding the current state of e
                        int sup[kLimbVers];                // block at which a version was superseded (its last reader is the dot of block sup + D)
#pragma unroll
                        for (int v = 0; v < kLimbVers; ++v) sup[v] = -1;
                        for (int ja = 0; ja < nblk; ++ja) {
                            const unsigned gidx = gblk + (unsigned)ja;
                            if (!(dbg & 1)) mbar_wait(&nz_full[r_nzw.s], r_nzw.ph);
                            else mbar_wait(&dot_done[gidx & (kNzRing - 1)], (gidx / kNzRing) & 1u);     // stay behind the dots like the real protocol
                            if (tid == 0) NGP_TICK(10);
                            const NzList& nl = wnz[r_nzw.s];
                            const int nnz = (dbg & 1) ? 0 : nl.nnz;
                            if (nnz > 0) {
                                // e -= sum_q dbeta_q (g_q - mean_q): one tile word = the 4 codes of this thread's rows
                                const int nv = (cur + 1 == NV) ? 0 : cur + 1;
                                int sv = -1;
#pragma unroll
                                for (int v = 0; v < kLimbVers; ++v) if (v == nv) sv = sup[v];
                                if (sv >= 0) {
                                    const int jl = min(sv + D, nblk - 1);          // last dot that reads version nv
                                    if (jl > ja) {                                  // dots <= ja are complete (their list exists)
                                        const unsigned gl = gblk + (unsigned)jl;
                                        mbar_wait(&dot_done[gl & (kNzRing - 1)], (gl / kNzRing) & 1u);
                                    }
                                }
                                if (tid == 0) NGP_TICK(5);
                                // one 32-bit word = the 4 codes of this thread's rows: from the tile, still resident in shared memory, or (big
                                // panels: the ring only covers the dots) re-read from L2 / HBM
                                const uint32_t* tw = P.refetch ? reinterpret_cast<const uint32_t*>(gbase + (int64_t)ja * L.tile_bytes)
                                                               : reinterpret_cast<const uint32_t*>(tiles + r_ta.s * L.tile_bytes);
#pragma unroll
                                for (int k = 0; k < UG; ++k) {
                                    const int rg = tid + k * kUpdThreads;
                                    if (4 * rg < R) {
                                        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                                        for (int i0 = 0; i0 < nnz; i0 += 4) {
                                            uint32_t w4[4];
#pragma unroll
                                            for (int u = 0; u < 4; ++u)
                                                w4[u] = (i0 + u >= nnz) ? 0u
                                                      : P.store2 ? expand2((uint32_t)reinterpret_cast<const uint8_t*>(tw)[word_off(B, nl.idx[i0 + u], rg)])
                                                                 : tw[word_off(B, nl.idx[i0 + u], rg)];
#pragma unroll
                                            for (int u = 0; u < 4; ++u)
                                                if (i0 + u < nnz) {
                                                    const uint32_t w = w4[u];
                                                    const double db = nl.db[i0 + u], kk = nl.aux[i0 + u];
                                                    s0 += fma(db, (double)(w & 0xff), -kk);
                                                    s1 += fma(db, (double)((w >> 8) & 0xff), -kk);
                                                    s2 += fma(db, (double)((w >> 16) & 0xff), -kk);
                                                    s3 += fma(db, (double)(w >> 24), -kk);
                                                }
                                        }
                                        const int r0 = 4 * rg;
                                        if (r0 < nrow) er[k][0] -= s0;
                                        if (r0 + 1 < nrow) er[k][1] -= s1;
                                        if (r0 + 2 < nrow) er[k][2] -= s2;
                                        if (r0 + 3 < nrow) er[k][3] -= s3;
                                        quantise4(er[k], limb + nv * lwords, rg, fx_scale);
                                    }
                                }
#pragma unroll
                                for (int v = 0; v < kLimbVers; ++v) if (v == cur) sup[v] = ja;
                                cur = nv;
                            }
                            if (tid == 0) vbuf[gidx & (kNzRing - 1)] = cur;
                            if constexpr (PROF) { if (tid == 0) { pf[26] += (long long)(global_ns() - pub_ns[gidx & (kNzRing - 1)]); pf[27] += 1; } }
                            __syncwarp();
                            if (lane == 0) {
                                mbar_arrive(&ver_full[gidx & (kNzRing - 1)]);
                                if (!P.refetch) mbar_arrive(&tile_free[r_ta.s]);
                                if (!(dbg & 1)) mbar_arrive(&nz_free[r_nzw.s]);
                            }
                            r_ta.adv(NT); r_nzw.adv(kNzSmem);
                            if (tid == 0) NGP_TICK(2);
                        }
                        // registers -> e_s (phases 0 and 3 and the next sweep read the master copy from smem)
#pragma unroll
                        for (int k = 0; k < UG; ++k) {
                            const int rg = tid + k * kUpdThreads;
                            if (4 * rg < R) {
                                *reinterpret_cast<double2*>(e_s + 4 * rg) = make_double2(er[k][0], er[k][1]);
                                *reinterpret_cast<double2*>(e_s + 4 * rg + 2) = make_double2(er[k][2], er[k][3]);
                            }
                        }
                    } else if (warp < kFirstDotWarp + kDotWarps) {
                        // ------------------------------------------------------------------ dot warps: whole blocks, 8 in flight
                        const int dw = warp - kFirstDotWarp;
                        const int g = lane >> 2, tt = lane & 3;
                        const long long plim = (1LL << (63 - cb)) / arr_all;
                        // refetch mode: a stage is always consumed by the same warp (the dot warps taking part divide NT), so an mbarrier
                        // waiter is never more than one fill behind
                        const int ND = !P.refetch ? kDotWarps : (NT >= 8) ? 8 : (NT >= 4) ? 4 : (NT >= 2) ? 2 : 1;
                        // BIGR on a refetch ring of ONE tile (a tile of 64 markers x > 1024 rows: two do not fit): all 8 dot warps share every tile — warp wsub
                        // takes the 32-row chunks wsub, wsub + 8, ... — and the last of them to finish adds up the 8 partial sums, pushes them and
                        // frees the tile (one warp needs longer for the dots of 88 KB than their load takes: C5 int8 3.24 -> 2.53 ms/sweep).  With two
                        // or four tiles in the ring sharing did not pay (C3 int8 14.1 -> 14.5, C5 2-bit 1.87 -> 2.05 ms: profiles/r2/tune_*_split_v1.jsonl).
                        constexpr int W = SHT ? kDotWarps : 1;                // (the launch picks the SHT instantiation for BIGR on a refetch ring with NT == 1)
                        const int grp = (W > 1) ? 0 : dw, wsub = (W > 1) ? dw : 0;
                        int* ccnt = ccnt_s;
                        const int j0 = (grp < ND) ? (int)((unsigned)(grp - (int)(gblk & (unsigned)(ND - 1))) & (unsigned)(ND - 1)) : nblk;    // first j with (gblk+j) % ND == grp
                        int tslot = (int)((gblk + (unsigned)j0) % (unsigned)NT);
                        uint32_t tph = ((gblk + (unsigned)j0) / (unsigned)NT) & 1u;
                        constexpr int CH = (MG >= 4) ? 1 : 4 / MG;          // independent accumulator chains per marker group
                        for (int j = j0; j < nblk; j += ND) {
                            const unsigned gidx = gblk + (unsigned)j;
                            const int ja = j - D - 1;
                            int ver = 0;
                            if (ja >= 0) {
                                const unsigned ga = gblk + (unsigned)ja;
                                mbar_wait(&ver_full[ga & (kNzRing - 1)], (ga / kNzRing) & 1u);
                                ver = vbuf[ga & (kNzRing - 1)];
                            }
                            if (tid == kFirstDotWarp * 32) NGP_TICK(14);
                            bool pusher = true;              // (BIGR with a shared tile: only the last warp to finish pushes the sums)
                            mbar_wait(&tile_full[tslot], tph);
                            if (tid == kFirstDotWarp * 32) NGP_TICK(20);
                            if (!(dbg & 8)) {
                                const unsigned char* tile = tiles + tslot * L.tile_bytes;
                                const uint32_t* lv = limb + ver * lwords;
                                int acc[MG][CH][4];
#pragma unroll
                                for (int mg = 0; mg < MG; ++mg)
#pragma unroll
                                    for (int ch = 0; ch < CH; ++ch) { acc[mg][ch][0] = acc[mg][ch][1] = acc[mg][ch][2] = acc[mg][ch][3] = 0; }
                                if (P.store2) {
                                    // 2-bit tiles: one 32-bit word per lane and MMA atom, expanded to the four A registers on chip
                                    const uint32_t* tile2 = reinterpret_cast<const uint32_t*>(tile);
                                    for (int c0 = wsub; c0 < nchunk; c0 += W * CH) {
#pragma unroll
                                        for (int ch = 0; ch < CH; ++ch) {
                                            const int c = c0 + ch * W;
                                            if (c < nchunk) {
                                                const uint32_t b0 = lv[c * 64 + lane], b1 = lv[c * 64 + 32 + (lane ^ 4)];
#pragma unroll
                                                for (int mg = 0; mg < MG; ++mg) {
                                                    const uint4 a = expand2_word(tile2[(c * MG + mg) * 32 + lane]);
                                                    imma16832(acc[mg][ch], a, b0, b1);
                                                }
                                            }
                                        }
                                    }
                                } else
                                for (int c0 = wsub; c0 < nchunk; c0 += W * CH) {
#pragma unroll
                                    for (int ch = 0; ch < CH; ++ch) {
                                        const int c = c0 + ch * W;
                                        if (c < nchunk) {
                                            const uint32_t b0 = lv[c * 64 + lane], b1 = lv[c * 64 + 32 + (lane ^ 4)];
#pragma unroll
                                            for (int mg = 0; mg < MG; ++mg) {
                                                const uint4 a = *reinterpret_cast<const uint4*>(tile + ((size_t)(c * MG + mg) * 32 + lane) * 16);
                                                imma16832(acc[mg][ch], a, b0, b1);
                                            }
                                        }
                                    }
                                }
                                if (tid == kFirstDotWarp * 32) NGP_TICK(1);
                                long long* accg = sy->acc + (size_t)(gidx & (kSlots - 1)) * kMaxB * kAccStride;
                                if constexpr (SHT) {
                                    long long part[MG];              // lanes tt = 0 / 1: this warp's sums of markers 16 mg + g / + 8
#pragma unroll
                                    for (int mg = 0; mg < MG; ++mg) {
                                        int c4[4];
#pragma unroll
                                        for (int x = 0; x < 4; ++x) {
                                            c4[x] = acc[mg][0][x];
#pragma unroll
                                            for (int ch = 1; ch < CH; ++ch) c4[x] += acc[mg][ch][x];
                                        }
                                        unsigned long long v0 = ((unsigned long long)(long long)c4[0] << (16 * tt)) + ((unsigned long long)(long long)c4[1] << (16 * tt + 8));
                                        unsigned long long v1 = ((unsigned long long)(long long)c4[2] << (16 * tt)) + ((unsigned long long)(long long)c4[3] << (16 * tt + 8));
                                        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                                        v0 += __shfl_xor_sync(0xffffffffu, v0, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
                                        part[mg] = (long long)(tt ? v1 : v0);
                                    }
                                    if (W > 1) {
                                        // (the tile is freed by the last warp, after it has read every row of sums: no warp is a tile ahead of the others)
                                        if (tt < 2) {
#pragma unroll
                                            for (int mg = 0; mg < MG; ++mg) comb[wsub * B + mg * 16 + g + 8 * tt] = part[mg];
                                        }
                                        __syncwarp();
                                        int old = 0;
                                        if (lane == 0) { __threadfence_block(); old = atomicAdd(ccnt, 1); }
                                        old = __shfl_sync(0xffffffffu, old, 0);
                                        pusher = (old == W - 1);
                                        if (pusher) {
                                            __threadfence_block();
                                            if (tt < 2) {
#pragma unroll
                                                for (int mg = 0; mg < MG; ++mg) {
                                                    long long tsum = 0;
                                                    for (int w = 0; w < W; ++w) tsum += comb[w * B + mg * 16 + g + 8 * tt];
                                                    part[mg] = tsum;
                                                }
                                            }
                                            __syncwarp();
                                            if (lane == 0) *ccnt = 0;
                                        }
                                    }
                                    if (pusher && tt < 2) {
#pragma unroll
                                        for (int mg = 0; mg < MG; ++mg) {
                                            const long long sa = part[mg];
                                            if (sa >= plim || sa <= -plim) atomicOr(&sy->err, 1);
                                            const long long rv = (long long)((unsigned long long)sa << cb) + 1;
                                            if (sharded && P.hier) {
                                                red_add_u64(sy->acc2 + (size_t)(gidx & (kSlots - 1)) * kMaxB + (size_t)(mg * 16 + g + 8 * tt), rv);
                                            } else if (sharded) {
                                                const size_t aoff = (size_t)(gidx & (kSlots - 1)) * kMaxB * kAccStride + (size_t)(mg * 16 + g + 8 * tt) * kAccStride;
                                                for (int r = 0; r < P.n_ranks; ++r) red_add_u64_sys(P.peer[r]->acc + aoff, rv);
                                            } else red_add_u64(accg + (mg * 16 + g + 8 * tt) * kAccStride, rv);
                                        }
                                    }
                                } else
#pragma unroll
                                for (int mg = 0; mg < MG; ++mg) {
                                    int c4[4];
#pragma unroll
                                    for (int x = 0; x < 4; ++x) {
                                        c4[x] = acc[mg][0][x];
#pragma unroll
                                        for (int ch = 1; ch < CH; ++ch) c4[x] += acc[mg][ch][x];
                                    }
                                    // lane holds limbs 2tt, 2tt+1 of markers g and g+8: sum_n c_n 256^n (mod 2^64)
                                    unsigned long long v0 = ((unsigned long long)(long long)c4[0] << (16 * tt)) + ((unsigned long long)(long long)c4[1] << (16 * tt + 8));
                                    unsigned long long v1 = ((unsigned long long)(long long)c4[2] << (16 * tt)) + ((unsigned long long)(long long)c4[3] << (16 * tt + 8));
                                    v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                                    v0 += __shfl_xor_sync(0xffffffffu, v0, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
                                    // lanes tt = 0 / tt = 1 add the sums of markers g / g+8 of this CTA's panel
                                    if (tt < 2) {
                                        const long long sa = (long long)(tt ? v1 : v0);
                                        if (sa >= plim || sa <= -plim) atomicOr(&sy->err, 1);
                                        const long long rv = (long long)((unsigned long long)sa << cb) + 1;
                                        if (sharded && P.hier) {
                                            // row-sharded chain, many ranks: reduce inside the rank first (its prep warps forward one total per marker)
                                            red_add_u64(sy->acc2 + (size_t)(gidx & (kSlots - 1)) * kMaxB + (size_t)(mg * 16 + g + 8 * tt), rv);
                                        } else if (sharded) {
                                            // row-sharded chain: the partial sums of the block go into the accumulator ring of EVERY rank (NVLink peer memory)
                                            const size_t aoff = (size_t)(gidx & (kSlots - 1)) * kMaxB * kAccStride + (size_t)(mg * 16 + g + 8 * tt) * kAccStride;
                                            for (int r = 0; r < P.n_ranks; ++r) red_add_u64_sys(P.peer[r]->acc + aoff, rv);
                                        } else red_add_u64(accg + (mg * 16 + g + 8 * tt) * kAccStride, rv);
                                    }
                                }
                            }
                            if constexpr (PROF) { if (tid == kFirstDotWarp * 32 && ja >= 0) { pf[28] += (long long)(global_ns() - pub_ns[(gblk + (unsigned)ja) & (kNzRing - 1)]); pf[29] += 1; } }
                            __syncwarp();
                            if (lane == 0 && pusher) { mbar_arrive(&dot_done[gidx & (kNzRing - 1)]); if (P.refetch) mbar_arrive(&tile_free[tslot]); }
                            tslot += ND;
                            while (tslot >= NT) { tslot -= NT; tph ^= 1u; }
                            if (tid == kFirstDotWarp * 32) NGP_TICK(15);
                        }
                    } else if (warp == kTmaWarp) {
                        // ------------------------------------------------------------------ TMA producer of the tile ring
                        if (lane == 0) {
                            const bool ef = (P.opt & 1) != 0;
                            const uint64_t pol = policy_evict_first();
                            for (int i = 0; i < nblk; ++i) {
                                if (tiles_issued >= (unsigned)NT) mbar_wait(&tile_free[r_tp.s], r_tp.ph ^ 1u);
                                mbar_expect_tx(&tile_full[r_tp.s], (uint32_t)L.tile_bytes);
                                if (ef) bulk_g2s_hint(tiles + r_tp.s * L.tile_bytes, gbase + (int64_t)i * L.tile_bytes, (uint32_t)L.tile_bytes, &tile_full[r_tp.s], pol);
                                else bulk_g2s(tiles + r_tp.s * L.tile_bytes, gbase + (int64_t)i * L.tile_bytes, (uint32_t)L.tile_bytes, &tile_full[r_tp.s]);
                                r_tp.adv(NT); ++tiles_issued;
                            }
                        }
                    } else {
                        // ------------------------------------------------------------------ poll warp: changed-effect lists
                        // one L2 round trip probes the first 8 words (header + one entry) of 4 consecutive slots
                        int ja_cur = 0;
                        const unsigned long long* llc = sy->ll[t % kLLCopies];      // this CTA's replica of the list ring
                        auto deliver_begin = [&]() -> NzList& {
                            if (nz_recv >= (unsigned)kNzSmem) mbar_wait(&nz_free[r_nzp.s], r_nzp.ph ^ 1u);
                            return wnz[r_nzp.s];
                        };
                        auto deliver_end = [&](NzList& nl, int nnz) {
                            if constexpr (PROF) {
                                if (lane == 0) {
                                    const unsigned gq = gblk + (unsigned)ja_cur;
                                    unsigned long long w;
                                    do { w = ld_relaxed_u64(llc + (size_t)(gq & (kNzRing - 1)) * kLLSlotWords + kLLSlotWords - 1); } while ((uint32_t)(w >> 32) != seq0 + gq + 1u);
                                    const unsigned long long tn = global_ns();
                                    const unsigned long long tp = (tn & ~0xffffffffull) | (w & 0xffffffffull);
                                    pub_ns[gq & (kNzRing - 1)] = tp;
                                    pf[24] += (long long)(tn - tp); pf[25] += 1;
                                }
                            }
                            if (lane == 0) nl.nnz = nnz;
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&nz_full[r_nzp.s]);
                            r_nzp.adv(kNzSmem); ++nz_recv;
                        };
                        const int g4 = lane >> 3, wi8 = lane & 7;
                        int ja = (dbg & 1) ? nblk : 0;
                        // kPollPipe probes are kept in flight, one issued every kPollGap cycles: a list is seen about half a gap
                        // (not half an L2 round trip, which is > 1 us under the streaming load) after it lands in L2
                        auto probe = [&](int tag) -> unsigned long long {
                            const unsigned gme = gblk + (unsigned)(tag + g4);
                            return (tag + g4 < nblk) ? ld_relaxed_u64(llc + (size_t)(gme & (kNzRing - 1)) * kLLSlotWords + wi8) : 0ull;
                        };
                        unsigned long long pwv[kPollPipe];
                        int ptag[kPollPipe];
#pragma unroll
                        for (int i = 0; i < kPollPipe; ++i) { pwv[i] = 0ull; ptag[i] = -1000; }     // tag -1000: an empty pipeline slot
                        long long t_issue = clock64() - kPollGap;
                        while (ja < nblk) {
                            ja_cur = ja;
                            // oldest probe out, new probe in (issued before the old one is looked at)
                            const unsigned long long w = pwv[0];
                            const int tag = ptag[0];
#pragma unroll
                            for (int i = 0; i + 1 < kPollPipe; ++i) { pwv[i] = pwv[i + 1]; ptag[i] = ptag[i + 1]; }
                            while (clock64() - t_issue < kPollGap) { }
                            t_issue = clock64();
                            pwv[kPollPipe - 1] = probe(ja); ptag[kPollPipe - 1] = ja;
                            if (tag + 4 <= ja) continue;                                   // empty slot, or everything it probed was delivered meanwhile
                            const bool ok = (tag + g4 < nblk) && ((uint32_t)(w >> 32) == seq0 + gblk + (unsigned)(tag + g4) + 1u);
                            const unsigned okm_all = __ballot_sync(0xffffffffu, ok);
                            const int skip = ja - tag;                                     // lane groups whose list was delivered by an earlier probe
                            const unsigned okm = okm_all >> (8 * skip);
                            for (int k = 0; k + skip < 4 && ja < nblk; ++k) {
                                if (!((okm >> (8 * k)) & 1u)) break;                       // header of the next list not there yet: next probe
                                const int nnz = (int)(uint32_t)__shfl_sync(0xffffffffu, w, 8 * (k + skip));
                                if (nnz == 0) {
                                    NzList& nl = deliver_begin();
                                    deliver_end(nl, 0);
                                } else if (nnz == 1) {
                                    if (((okm >> (8 * k)) & 0x3Eu) != 0x3Eu) break;         // its entry is still in flight
                                    NzList& nl = deliver_begin();
                                    if (g4 == k + skip && wi8 >= 1 && wi8 <= kLLEntryWords) {
                                        const uint32_t pay = (uint32_t)w;
                                        if (wi8 == 1) nl.idx[0] = (int)pay;
                                        else if (wi8 <= 3) reinterpret_cast<uint32_t*>(&nl.db[0])[wi8 - 2] = pay;
                                        else reinterpret_cast<uint32_t*>(&nl.aux[0])[wi8 - 4] = pay;
                                    }
                                    deliver_end(nl, 1);
                                } else {
                                    // long list: read the whole slot, 32 words per round trip
                                    const unsigned gidx = gblk + (unsigned)ja;
                                    const uint32_t seq = seq0 + gidx + 1u;
                                    const unsigned long long* sw = llc + (size_t)(gidx & (kNzRing - 1)) * kLLSlotWords;
                                    NzList& nl = deliver_begin();
                                    const int need = 1 + kLLEntryWords * nnz;
                                    for (int base = 0; base < need; base += 32) {
                                        unsigned long long w2;
                                        for (;;) {
                                            w2 = ld_relaxed_u64(sw + base + lane);
                                            if (__all_sync(0xffffffffu, ((uint32_t)(w2 >> 32) == seq) || (base + lane >= need))) break;
                                        }
                                        const int wi = base + lane;
                                        if (wi >= 1 && wi < need) {
                                            const int en = (wi - 1) / kLLEntryWords, f = (wi - 1) - en * kLLEntryWords;
                                            const uint32_t pay = (uint32_t)w2;
                                            if (f == 0) nl.idx[en] = (int)pay;
                                            else if (f <= 2) reinterpret_cast<uint32_t*>(&nl.db[en])[f - 1] = pay;
                                            else reinterpret_cast<uint32_t*>(&nl.aux[en])[f - 3] = pay;
                                        }
                                    }
                                    deliver_end(nl, nnz);
                                }
                                ++ja; ja_cur = ja;
                            }
                        }
                    }
                } else {
                    // ====================================================================== chain CTA
                    unsigned char* recs = smem + L.c_rec;
                    const int32_t* const gx_g = S.gx;
                    const size_t gx_blk = (size_t)(D + 1) * B * B;
                    const int gofs = (1 + DN) * gram_bytes(B);          // constants follow the Gram matrices in a record
                    if (warp == 0) {
                        // ------------------------------------------------------------------ chain warp: SM = 32 NS markers per step
                        // slot i of a lane = marker mi = 32 i + lane of the step: block kb = mi / B of the step, column qb = mi % B
                        int kb[NS], qb[NS];
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            if constexpr (B >= 32) { kb[i] = (32 * i) / B; qb[i] = (32 * i) % B + lane; }      // the block of a slot is a compile-time constant
                            else { kb[i] = (32 * i + lane) / B; qb[i] = (32 * i + lane) % B; }
                        }
                        int nn[SB];                                         // changed effects per block of the step
#pragma unroll
                        for (int k = 0; k < SB; ++k) nn[k] = 0;
                        for (int s0 = 0; s0 < nblk; s0 += SB) {
                            const unsigned g0 = gblk + (unsigned)s0;            // global number of the first block of the step
                            const unsigned sg = g0 >> SBS;                      // global step number
                            const int ss = (int)(sg & (RBS - 1));
                            if constexpr (PROF) tc = clock64();
                            // r_base of the step's blocks is published after their records have landed (the prep warps waited for them)
                            mbar_wait(&rb_full[ss], (sg / RBS) & 1u);
                            if constexpr (PROF) {
                                const long long now_ = clock64();
                                if (lane == 0 && (s0 >> SBS) < 2048) { sy->trace[2 * (s0 >> SBS)] = tc; sy->trace[2 * (s0 >> SBS) + 1] = now_ - tc; }
                            }
                            NGP_TICK(3);
                            const int32_t* gram[NS];
                            const double* cst[NS];
                            // rr = x'e + d*beta_old (add-back fused: x'(e + x b) = x'e + d b) is the running quantity; bnz = (beta_old != 0)
                            double rr[NS], bold[NS], cs[NS], bnew[NS], cA[NS], cB[NS], cT[NS], cC[NS], cQ[NS];
                            bool inc[NS], bnz[NS];
#pragma unroll
                            for (int i = 0; i < NS; ++i) {
                                const unsigned gk = g0 + (unsigned)kb[i];
                                const unsigned char* rec = recs + (gk & (unsigned)(NR - 1)) * L.rec_bytes;
                                gram[i] = reinterpret_cast<const int32_t*>(rec);
                                cst[i] = reinterpret_cast<const double*>(rec + gofs) + qb[i];
                                const double r0 = rbase[(gk & (kPrepWarps - 1)) * B + qb[i]];
                                cs[i] = cst[i][F_CS * B]; bold[i] = cst[i][F_BOLD * B];
                                rr[i] = fma(cst[i][F_D * B], bold[i], r0);
                                cA[i] = cst[i][F_A * B]; cB[i] = cst[i][F_B * B]; cT[i] = cst[i][F_T * B];
                                cC[i] = cst[i][F_C * B]; cQ[i] = cst[i][F_QSZ * B];
                                bnew[i] = 0.0; inc[i] = false; bnz[i] = bold[i] != 0.0;
                            }
                            // BayesR (functions.jl:238-289), classes v < nc <= 4: everything of the class likelihoods that does not depend on the running
                            // dot, per marker:  1/lhs_v  and  A_v = -0.5 log(varc_v lhs_v) + logPi_v  (a class of variance 0: 1/lhs = 0, A = logPi)
                            double r_il[BR ? NS : 1][4], r_A[BR ? NS : 1][4];
                            int ocls[NS];
                            if constexpr (BR) {
#pragma unroll
                                for (int i = 0; i < NS; ++i) {
                                    const double dj = cst[i][F_D * B];
#pragma unroll
                                    for (int v = 0; v < 4; ++v) {
                                        const double vc = br_varc[v];
                                        const bool zero = !(v < br_nc) || vc == 0.0;
                                        const double lhs_v = fma(dj, br_iVarE, 1.0 / vc);
                                        r_il[i][v] = zero ? 0.0 : 1.0 / lhs_v;
                                        r_A[i][v] = (v < br_nc) ? (zero ? br_logpi[v] : fma(-0.5, log(vc * lhs_v), br_logpi[v])) : -INFINITY;
                                    }
                                    ocls[i] = 0;
                                }
                            }
                            NGP_TICK(4);
                            // cross-Gram corrections from the blocks of the previous step (distances 1 .. 2 SB - 1 <= DN), oldest first
                            if (s0 > 0 && !(dbg & 4)) {
#pragma unroll
                                for (int kk = SB; kk >= 1; --kk) {
                                    const int np = nn[SB - kk];                 // list sizes of the previous step are still in registers
                                    if (np == 0) continue;
                                    const NzList& pl = cnz[(g0 - (unsigned)kk) & (kNzRing - 1)];
                                    // four entries at a time: the list reads and the Gram loads of a group are issued together (a dense
                                    // list — BayesPR, tuple — has B entries; one entry at a time costs a full LDS -> LDS -> I2F -> DFMA chain each)
                                    // (the tuple instantiation keeps one entry at a time: its chain warp is register-bound by the joint draw
                                    //  and the grouped form cost C4 45.6 -> 57.4 ms/sweep)
                                    if constexpr (TUP) {
                                        for (int e = 0; e < np; ++e) {
                                            const int a = pl.idx[e];
                                            const double dbf = pl.db[e], csf = pl.aux[e];
#pragma unroll
                                            for (int i = 0; i < NS; ++i)
                                                rr[i] = fma(-((double)gram[i][(kb[i] + kk) * B * B + a * B + qb[i]] - csf * cs[i] * inv_n), dbf, rr[i]);
                                        }
                                    } else
                                    for (int e0 = 0; e0 < np; e0 += 4) {
                                        int a4[4]; double db4[4], cs4[4]; int g4_[4][NS];
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            const bool on = e0 + u < np;
                                            a4[u] = on ? pl.idx[e0 + u] : 0; db4[u] = on ? pl.db[e0 + u] : 0.0; cs4[u] = on ? pl.aux[e0 + u] : 0.0;
                                        }
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
#pragma unroll
                                            for (int i = 0; i < NS; ++i) g4_[u][i] = gram[i][(kb[i] + kk) * B * B + a4[u] * B + qb[i]];
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
#pragma unroll
                                            for (int i = 0; i < NS; ++i)
                                                rr[i] = fma(-((double)g4_[u][i] - cs4[u] * cs[i] * inv_n), db4[u], rr[i]);
                                    }
                                }
                            }
                            NGP_TICK(16);
#pragma unroll
                            for (int k = 0; k < SB; ++k) nn[k] = 0;
                            int pos = (dbg & 4) ? SM : 0;                       // markers [0, pos) of the step are committed
                            if (TUP && method == 4 && pos < SM) {
                                // Tuple of k interleaved marker sets (functions.jl:140-154): lanes g0 .. g0+k-1 of a slot hold the k breeds of one locus.
                                // Strictly sequential over the loci; per locus the k x k solve is evaluated by every lane (uniform inputs).
                                if (S.group_k == 2) joint_step<B, NS, 2>(rr, bold, cs, cQ, cT, bnew, inc, gram, kb, qb, cnz, g0, lane, inv_n, 1.0 / varE, S.jinvB, sy);
                                else joint_step<B, NS, 4>(rr, bold, cs, cQ, cT, bnew, inc, gram, kb, qb, cnz, g0, lane, inv_n, 1.0 / varE, S.jinvB, sy);
#pragma unroll
                                for (int k = 0; k < SB; ++k) nn[k] = B;
                                if constexpr (PROF) pf[7] += SM;
                                pos = SM;
                            }
                            if (method == 0 && pos < SM) {
                                // BayesPR: every effect is redrawn, so there is nothing to speculate on and the positions of the changes are
                                // known: the dependent chain per marker is  beta_new = rr*C + QSZ -> dbeta -> shuffle -> rr of the later markers;
                                // the Gram values of marker m against the later ones do not depend on dbeta and are fetched ahead of it.
                                double mydb[NS];
#pragma unroll
                                for (int i = 0; i < NS; ++i) {
#pragma unroll 4
                                    for (int f = 0; f < 32; ++f) {
                                        const int m = 32 * i + f;
                                        const int ka = m / B, qa = m % B;
                                        const double bn = fma(rr[i], cC[i], cQ[i]);
                                        const double dbl = bn - bold[i];
                                        const double dbf = __shfl_sync(0xffffffffu, dbl, f);
                                        const double csf = __shfl_sync(0xffffffffu, cs[i], f);
                                        if (lane == f) { bnew[i] = bn; inc[i] = true; mydb[i] = dbl; }
#pragma unroll
                                        for (int i2 = i; i2 < NS; ++i2)
                                            if (32 * i2 + lane > m) {
                                                const double gc = (double)gram[i2][(kb[i2] - ka) * B * B + qa * B + qb[i2]] - csf * cs[i2] * inv_n;
                                                rr[i2] = fma(-gc, dbf, rr[i2]);
                                            }
                                    }
                                }
                                // every marker of the step is a list entry, in marker order: each lane writes its own
#pragma unroll
                                for (int i = 0; i < NS; ++i) {
                                    NzList& ml = cnz[(g0 + (unsigned)kb[i]) & (kNzRing - 1)];
                                    ml.idx[qb[i]] = qb[i]; ml.db[qb[i]] = mydb[i]; ml.aux[qb[i]] = cs[i];
                                }
#pragma unroll
                                for (int k = 0; k < SB; ++k) nn[k] = B;
                                if constexpr (PROF) pf[7] += SM;
                                pos = SM;
                            }
                            while (pos < SM) {
                                if constexpr (PROF) pf[7]++;
                                double bnv[NS];
                                bool inv[NS];
                                unsigned mk[NS];
#pragma unroll
                                int clsv[NS];
#pragma unroll
                                for (int i = 0; i < NS; ++i) {
                                    clsv[i] = 0;
                                    if constexpr (BR) {
                                        // class likelihoods exp(A_v + rhs^2 / (2 lhs_v)), proportions in class order, first class whose cumulative
                                        // probability reaches ITS uniform (functions.jl:250-262); beta from the class's lhs (functions.jl:266-275)
                                        const double rhs = rr[i] * br_iVarE, hq = 0.5 * rhs * rhs;
                                        const double uu[4] = {cA[i], cB[i], cT[i], cC[i]};
                                        double ex[4], tot = 0.0;
#pragma unroll
                                        for (int v = 0; v < 4; ++v) { ex[v] = (v < br_nc) ? exp(fma(hq, r_il[i][v], r_A[i][v])) : 0.0; tot += ex[v]; }
                                        double cum = 0.0, ilc = 0.0;
                                        int cls = -1;
#pragma unroll
                                        for (int v = 0; v < 4; ++v) {
                                            cum += ex[v] / tot;
                                            if (cls < 0 && v < br_nc && cum >= uu[v]) { cls = v; ilc = r_il[i][v]; }
                                        }
                                        if (cls < 0) { if (32 * i + lane >= pos && bold[i] == bold[i]) br_bad = true; cls = 0; ilc = r_il[i][0]; }
                                        clsv[i] = cls;
                                        bnv[i] = (ilc != 0.0) ? fma(rhs, ilc, sqrt(ilc) * cQ[i]) : 0.0;
                                        inv[i] = true;
                                    } else {
                                    const double dl = fma(cB[i], rr[i] * rr[i], cA[i]);
                                    bnv[i] = fma(rr[i], cC[i], cQ[i]);                  // evaluated alongside the inclusion test
                                    inv[i] = dl < cT[i];                                // NaN -> excluded, like rand() < NaN
                                    }
                                    const bool ch = inv[i] ? (bnv[i] != bold[i]) : bnz[i];      // effect changes <=> beta_new - beta_old != 0
                                    mk[i] = __ballot_sync(0xffffffffu, (32 * i + lane >= pos) && ch);
                                }
                                const int ci = (NS == 1 || mk[0]) ? 0 : 1;              // slot of the first marker whose effect changes
                                const unsigned mm = (NS == 1 || mk[0]) ? mk[0] : mk[NS - 1];
                                const int f = mm ? (__ffs(mm) - 1) : 32;
                                const int last = mm ? 32 * ci + f : SM - 1;                 // commit markers [pos, last]
#pragma unroll
                                for (int i = 0; i < NS; ++i) {
                                    const int mi = 32 * i + lane;
                                    if (mi >= pos && mi <= last) { bnew[i] = inv[i] ? bnv[i] : 0.0; inc[i] = inv[i]; if constexpr (BR) ocls[i] = clsv[i]; }
                                }
                                if (!mm) break;
                                const double mydb = ci ? ((inv[NS - 1] ? bnv[NS - 1] : 0.0) - bold[NS - 1]) : ((inv[0] ? bnv[0] : 0.0) - bold[0]);
                                const double dbf = __shfl_sync(0xffffffffu, mydb, f);
                                const double csf = __shfl_sync(0xffffffffu, ci ? cs[NS - 1] : cs[0], f);
                                const int ka = (32 * ci + f) / B, qa = (32 * ci + f) % B;
                                // restore the later markers of the step: Gram row of (block ka, column qa) against their block
#pragma unroll
                                for (int i = 0; i < NS; ++i) {
                                    if (32 * i + lane > last) {
                                        const double gc = (double)gram[i][(kb[i] - ka) * B * B + qa * B + qb[i]] - csf * cs[i] * inv_n;
                                        rr[i] = fma(-gc, dbf, rr[i]);
                                    }
                                }
                                if (lane == 0) {
                                    NzList& ml = cnz[(g0 + (unsigned)ka) & (kNzRing - 1)];
                                    int at = 0;
#pragma unroll
                                    for (int k = 0; k < SB; ++k) if (k == ka) at = nn[k];
                                    ml.idx[at] = qa; ml.db[at] = dbf; ml.aux[at] = csf;
                                }
#pragma unroll
                                for (int k = 0; k < SB; ++k) if (k == ka) nn[k]++;
                                pos = last + 1;
                            }
                            NGP_TICK(17);
                            if constexpr (BR) { if (br_bad) { atomicOr(&sy->err, 4); br_bad = false; } }       // findfirst found nothing: the reference throws here
                            // hand the step over: lists (prep warps, helper warp), new effects and indicators (helper warp)
                            const int os = (int)(sg & (kPrepWarps - 1));
#pragma unroll
                            for (int i = 0; i < NS; ++i) { out_b[os * 64 + 32 * i + lane] = bnew[i]; out_i[os * 64 + 32 * i + lane] = BR ? ocls[i] : (inc[i] ? 1 : 0); }
                            if (lane == 0) {
#pragma unroll
                                for (int k = 0; k < SB; ++k) cnz[(g0 + (unsigned)k) & (kNzRing - 1)].nnz = nn[k];
                            }
                            __syncwarp();
                            if (lane == 0) {
                                *cprog = g0 + (unsigned)SB;                     // prep warps skip the barriers of lists known to be complete
                                mbar_arrive(&step_done[os]);
#pragma unroll
                                for (int k = 0; k < SB; ++k) mbar_arrive(&nzc_full[(g0 + (unsigned)k) & (kNzRing - 1)]);
                            }
                            if constexpr (PROF) {
#pragma unroll
                                for (int k = 0; k < SB; ++k) pf[6] += nn[k];
                                if (lane == 0) step_end_clk[sg & 63u] = clock64();
                            }
                            NGP_TICK(18);
                        }
                    } else if (warp == 1) {
                        // ------------------------------------------------------------------ TMA producer of the block records
                        if (lane == 0) {
                            fence_proxy_async();     // consts were written through the generic proxy by other CTAs
                            const uint32_t gbytes = (uint32_t)gofs, cbytes = (uint32_t)consts_bytes(B);
                            for (int m = 0; m < nblk; ++m) {
                                if (recs_issued >= (unsigned)NR) mbar_wait(&rec_free[r_rec.s], r_rec.ph ^ 1u);
                                unsigned char* dst = recs + r_rec.s * L.rec_bytes;
                                mbar_expect_tx(&rec_full[r_rec.s], gbytes + cbytes);
                                bulk_g2s(dst, gx_g + (size_t)m * gx_blk, gbytes, &rec_full[r_rec.s]);
                                bulk_g2s(dst + gbytes, S.consts + (size_t)m * kNF * B, cbytes, &rec_full[r_rec.s]);
                                r_rec.adv(NR); ++recs_issued;
                            }
                        }
                    } else if (warp < kFirstPrepWarp + kPrepWarps) {
                        // ------------------------------------------------------------------ prep warps
                        const int pw = warp - kFirstPrepWarp;
                        int m0 = (int)((unsigned)(pw - (int)(gblk & (kPrepWarps - 1))) & (kPrepWarps - 1));   // first m with (gblk+m) % kPrepWarps == pw
                        const bool hier = sharded && P.hier;
                        const long long arr_exp = hier ? (long long)P.n_ranks : arr_all;      // arrivals per accumulator of this rank
                        for (int m = m0; m < nblk; m += kPrepWarps) {
                            const unsigned gidx = gblk + (unsigned)m;
                            const unsigned sg = gidx >> SBS;
                            const int ss = (int)(sg & (RBS - 1));
                            if (warp == kFirstPrepWarp) if constexpr (PROF) tc = clock64();
                            double far[NB], cs[NB], mean[NB];
                            bool live[NB];
#pragma unroll
                            for (int b = 0; b < NB; ++b) {
                                const int q = b * 32 + lane;
                                live[b] = q < B;
                                const int64_t j = (int64_t)m * B + (live[b] ? q : 0);
                                cs[b] = (double)__ldg(&S.colsum[j]);
                                mean[b] = __ldg(&S.mean[j]);
                                far[b] = 0.0;
                            }
                            // cross-Gram corrections from the blocks before the previous step, oldest block first:
                            //   distances > DN: rows fetched on demand from HBM/L2;  distances <= DN: rows from the block record
                            bool have_rec = false;
                            const unsigned rcs = gidx & (unsigned)(NR - 1);
                            const uint32_t rcp = (gidx / (unsigned)NR) & 1u;
                            const int32_t* gram = reinterpret_cast<const int32_t*>(recs + rcs * L.rec_bytes);
                            // An mbarrier wait only tells the parity of the last completed phase: before waiting for the record of block m
                            // make sure the previous occupant of its stage (block m - NR) is gone, i.e. that block has been through the chain warp
                            auto wait_record = [&]() {
                                if (m >= NR) {
                                    const unsigned gp = gidx - (unsigned)NR;
                                    mbar_wait(&nzc_full[gp & (kNzRing - 1)], (gp / kNzRing) & 1u);
                                }
                                mbar_wait(&rec_full[rcs], rcp);
                            };
                            auto apply_list = [&](int sblk) {
                                const NzList& pl = cnz[(gblk + (unsigned)sblk) & (kNzRing - 1)];
                                const int np = pl.nnz;
                                const int d = m - sblk;
                                if (d > DN) {
                                    const int32_t* gd = gx_g + (size_t)m * gx_blk + (size_t)d * B * B;
                                    for (int i0 = 0; i0 < np; i0 += 4) {
                                        int gv[4][NB];
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
#pragma unroll
                                            for (int b = 0; b < NB; ++b)
                                                gv[u][b] = (i0 + u < np && live[b]) ? __ldg(gd + pl.idx[i0 + u] * B + b * 32 + lane) : 0;
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
                                            if (i0 + u < np) {
                                                const double dbf = pl.db[i0 + u], csf = pl.aux[i0 + u];
#pragma unroll
                                                for (int b = 0; b < NB; ++b) far[b] = fma(-((double)gv[u][b] - csf * cs[b] * inv_n), dbf, far[b]);
                                            }
                                    }
                                } else {
                                    if (!have_rec) { wait_record(); have_rec = true; }
                                    const int32_t* gd = gram + d * B * B;
                                    for (int i = 0; i < np; ++i) {
                                        const int a = pl.idx[i];
                                        const double dbf = pl.db[i], csf = pl.aux[i];
#pragma unroll
                                        for (int b = 0; b < NB; ++b)
                                            if (live[b]) far[b] = fma(-((double)gd[a * B + b * 32 + lane] - csf * cs[b] * inv_n), dbf, far[b]);
                                    }
                                }
                            };
                            const int s_first = max(0, m - D);
                            const int s_last = (m & ~(SB - 1)) - SB - 1;          // last block whose list is not the chain warp's business
                            // pass 1, without blocking: the lists the chain warp has already completed (one ballot finds the non-empty ones)
                            int s_next = s_first;
                            {
                                const int done = (int)(*cprog - gblk);              // blocks of this sweep with complete lists
                                const int avail = min(s_last, done - 1);
                                const int sb_ = s_first + lane;
                                const bool has = (sb_ <= avail) && (cnz[(gblk + (unsigned)sb_) & (kNzRing - 1)].nnz > 0);
                                unsigned msk = __ballot_sync(0xffffffffu, has);
                                // the far rows of all these lists are prefetched together (one 128-byte line per changed marker and
                                // 32 columns), so that the in-order accumulation below does not pay one L2 round trip per list
                                for (unsigned mp = msk; mp; mp &= mp - 1) {
                                    const int sb2 = s_first + __ffs(mp) - 1, d2 = m - sb2;
                                    if (d2 > DN) {
                                        const NzList& pl = cnz[(gblk + (unsigned)sb2) & (kNzRing - 1)];
                                        const int32_t* gd = gx_g + (size_t)m * gx_blk + (size_t)d2 * B * B;
                                        const int np = pl.nnz;
                                        for (int e = 0; e < np; ++e)
#pragma unroll
                                            for (int b = 0; b < NB; ++b)
                                                if (live[b]) asm volatile("prefetch.global.L1 [%0];" ::"l"(gd + pl.idx[e] * B + b * 32 + lane));
                                    }
                                }
                                while (msk) { const int l = __ffs(msk) - 1; msk &= msk - 1; apply_list(s_first + l); }
                                if (avail >= s_first) s_next = avail + 1;
                            }
                            // poll the accumulators of block m until all Tw worker CTAs have added their partial sums
                            const int slot = (int)(gidx & (kSlots - 1));
                            const long long* acc = sy->acc + (size_t)slot * kMaxB * kAccStride;
                            long long cur[NB];
#pragma unroll
                            for (int b = 0; b < NB; ++b) cur[b] = 0;
                            if (hier) {
                                // stage 1: wait for the Tw worker CTAs of THIS rank, forward the rank's total of every marker to every rank.  The staging
                                // accumulators start from zero and are reset here: nobody adds to a slot again before its block has been consumed
                                // (a worker is never kSlots blocks ahead of the chain)
                                long long* a2 = sy->acc2 + (size_t)slot * kMaxB;
                                long long c2[NB];
                                for (;;) {
                                    bool done = true;
#pragma unroll
                                    for (int b = 0; b < NB; ++b) {
                                        c2[b] = live[b] ? ld_relaxed_s64(a2 + b * 32 + lane) : 0;
                                        if (live[b]) done = done && ((c2[b] & cmask) == (long long)Tw);
                                    }
                                    if (__all_sync(0xffffffffu, done)) break;
                                }
#pragma unroll
                                for (int b = 0; b < NB; ++b)
                                    if (live[b]) {
                                        st_relaxed_u64(reinterpret_cast<unsigned long long*>(a2 + b * 32 + lane), 0ull);
                                        const long long v = (c2[b] - (long long)Tw) >> cb;                  // exact integer sum of this rank's rows
                                        const long long rv = (long long)((unsigned long long)v << cb) + 1;
                                        const size_t aoff = (size_t)slot * kMaxB * kAccStride + (size_t)(b * 32 + lane) * kAccStride;
                                        for (int r = 0; r < P.n_ranks; ++r) red_add_u64_sys(P.peer[r]->acc + aoff, rv);
                                    }
                            }
                            if (!(dbg & 2)) {
                                // pipelined like the list polling of the worker CTAs: kPollPipe probes in flight, kPollGap cycles apart
                                long long pv_[kPollPipe][NB];
                                bool pvalid[kPollPipe];
#pragma unroll
                                for (int i = 0; i < kPollPipe; ++i) pvalid[i] = false;
                                long long t_issue = clock64() - kPollGap;
                                unsigned sh_spins = 0;
                                unsigned long long sh_t0 = 0;
                                for (;;) {
                                    bool done = pvalid[0];
#pragma unroll
                                    for (int b = 0; b < NB; ++b) {
                                        cur[b] = pv_[0][b];
                                        if (live[b] && pvalid[0]) done = done && (((cur[b] - prev[slot * B + b * 32 + lane]) & cmask) == arr_exp);
                                    }
                                    const bool valid0 = pvalid[0];
#pragma unroll
                                    for (int i = 0; i + 1 < kPollPipe; ++i) {
                                        pvalid[i] = pvalid[i + 1];
#pragma unroll
                                        for (int b = 0; b < NB; ++b) pv_[i][b] = pv_[i + 1][b];
                                    }
                                    if (valid0 && __all_sync(0xffffffffu, done)) break;
                                    while (clock64() - t_issue < kPollGap) { }
                                    t_issue = clock64();
#pragma unroll
                                    for (int b = 0; b < NB; ++b)
                                        pv_[kPollPipe - 1][b] = !live[b] ? 0 : sharded ? ld_relaxed_s64_sys(acc + (b * 32 + lane) * kAccStride) : ld_relaxed_s64(acc + (b * 32 + lane) * kAccStride);
                                    pvalid[kPollPipe - 1] = true;
                                    if (sharded && (++sh_spins & 0x3ffu) == 0u && shard_wait_expired(&sy->err, sh_t0)) break;      // a rank that never arrives
                                }
                            }
                            if (warp == kFirstPrepWarp) NGP_TICK(13);
                            if constexpr (PROF) {
                                // loop latency: end of the chain step that released the dots of block m -> their sums are complete
                                if (warp == kFirstPrepWarp && m - D - 1 >= 0) {
                                    const long long lat = clock64() - step_end_clk[((gblk + (unsigned)(m - D - 1)) >> SBS) & 63u];
                                    pf[21] += lat; pf[22] += 1; if (lat > pf[23]) pf[23] = lat;
                                    pf[30] += (long long)(global_ns() - pub_ns[(gblk + (unsigned)(m - D - 1)) & (kNzRing - 1)]); pf[31] += 1;
                                }
                            }
                            // pass 2: the remaining lists, as the chain warp completes them
                            for (int sblk = s_next; sblk <= s_last; ++sblk) {
                                const unsigned gs_ = gblk + (unsigned)sblk;
                                mbar_wait(&nzc_full[gs_ & (kNzRing - 1)], (gs_ / kNzRing) & 1u);
                                if (cnz[gs_ & (kNzRing - 1)].nnz > 0) apply_list(sblk);
                            }
                            if (!have_rec) wait_record();     // the chain warp relies on it
                            if (warp == kFirstPrepWarp) NGP_TICK(12);
                            if (sg >= (unsigned)RBS) mbar_wait(&rb_free[ss], ((sg / RBS) - 1u) & 1u);      // the helper warp is done with the slot
#pragma unroll
                            for (int b = 0; b < NB; ++b) {
                                const int q = b * 32 + lane;
                                if (live[b]) {
                                    long long* pv = prev + slot * B + q;
                                    const double A = (double)((cur[b] - *pv - arr_exp) >> cb) * fx_inv;
                                    if (!(dbg & 2)) *pv = cur[b];
                                    rbase[pw * B + q] = (A - mean[b] * Stot) + far[b];       // x_q'e as the dots saw it + the corrections above
                                }
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&rb_full[ss]);
                        }
                    } else if (warp == kHelperWarp) {
                        // ------------------------------------------------------------------ helper warp: publish the lists, write the outputs
                        double* const beta_g = S.beta;
                        int32_t* const delta_g = S.delta;
                        double* const vb_g = S.varBeta;
                        const double sdf = S.scale * S.df;
                        int kb[NS], qb[NS];
#pragma unroll
                        for (int i = 0; i < NS; ++i) { kb[i] = (32 * i + lane) / B; qb[i] = (32 * i + lane) % B; }
                        for (int s0 = 0; s0 < nblk; s0 += SB) {
                            const unsigned g0 = gblk + (unsigned)s0;
                            const unsigned sg = g0 >> SBS;
                            const int os = (int)(sg & (kPrepWarps - 1));
                            mbar_wait(&step_done[os], (sg / kPrepWarps) & 1u);
                            // changed effects -> worker CTAs: {payload32, seq32} words, any order, no fence
#pragma unroll
                            for (int k = 0; k < SB; ++k) {
                                const unsigned gk = g0 + (unsigned)k;
                                const NzList& ml = cnz[gk & (kNzRing - 1)];
                                const int nnz = ml.nnz;
                                const size_t lofs = (size_t)(gk & (kNzRing - 1)) * kLLSlotWords;
                                const unsigned long long seqhi = (unsigned long long)(seq0 + gk + 1u) << 32;
                                const double* cm = reinterpret_cast<const double*>(recs + (gk & (unsigned)(NR - 1)) * L.rec_bytes + gofs) + F_MEAN * B;
                                // (entry, replica) pairs are spread over the lanes; the header of a replica goes last (any order is fine)
                                for (int x = lane; x < nnz * kLLCopies; x += 32) {
                                    const int e = x / kLLCopies, cpy_ = x % kLLCopies;
                                    const int q = ml.idx[e];
                                    const double dbf = ml.db[e];
                                    const unsigned long long dbb = (unsigned long long)__double_as_longlong(dbf);
                                    const unsigned long long kkb = (unsigned long long)__double_as_longlong(dbf * cm[q]);
                                    unsigned long long* ew = sy->ll[cpy_] + lofs + 1 + kLLEntryWords * e;
                                    st_relaxed_u64(ew + 0, seqhi | (unsigned long long)(uint32_t)q);
                                    st_relaxed_u64(ew + 1, seqhi | (dbb & 0xffffffffull));
                                    st_relaxed_u64(ew + 2, seqhi | (dbb >> 32));
                                    st_relaxed_u64(ew + 3, seqhi | (kkb & 0xffffffffull));
                                    st_relaxed_u64(ew + 4, seqhi | (kkb >> 32));
                                }
                                if (lane < kLLCopies) st_relaxed_u64(sy->ll[lane] + lofs, seqhi | (unsigned long long)(uint32_t)nnz);
                                if (P.opt & 2) {
                                    // the far cross-Gram rows of a changed effect (distances DN+1 .. D) will be wanted by the prep warps a few
                                    // blocks from now: start them on their way from HBM into L2
                                    const int nd = D - DN;
                                    const int mblk = s0 + k;
                                    for (int x = lane; x < nnz * nd; x += 32) {
                                        const int e = x / nd, d = DN + 1 + x % nd;
                                        const int m = mblk + d;
                                        if (m < nblk) {
                                            const int32_t* row = gx_g + (size_t)m * gx_blk + (size_t)d * B * B + ml.idx[e] * B;
                                            prefetch_l2(row);
                                            if (B == 64) prefetch_l2(row + 32);
                                        }
                                    }
                                }
                                if constexpr (PROF) {
                                    const unsigned long long tn = global_ns();
                                    if (lane == 0) pub_ns[gk & (kNzRing - 1)] = tn;
                                    if (lane < kLLCopies) st_relaxed_u64(sy->ll[lane] + lofs + kLLSlotWords - 1, seqhi | (tn & 0xffffffffull));
                                }
                            }
                            // outputs of the step (plain coalesced stores)
#pragma unroll
                            for (int i = 0; i < NS; ++i) {
                                const unsigned gk = g0 + (unsigned)kb[i];
                                const int64_t j = (int64_t)(s0 + kb[i]) * B + qb[i];
                                const double bn = out_b[os * 64 + 32 * i + lane];
                                if constexpr (BR) {
                                    // BayesR: delta = class (1-based, functions.jl:263), nLoci per class, sumS = sum beta^2 / vClass and the
                                    // number of loci in classes of non-zero variance (functions.jl:264-273)
                                    const int cls = out_i[os * 64 + 32 * i + lane];
                                    const bool real = j < p_real;
#pragma unroll
                                    for (int v = 0; v < 4; ++v) {
                                        const unsigned mv = __ballot_sync(0xffffffffu, real && cls == v);
                                        if (lane == v) acc_cls += (double)__popc(mv);
                                    }
                                    if (real) {
                                        double vcl = br_vcls[0], vcc = br_varc[0];
#pragma unroll
                                        for (int v = 1; v < 4; ++v) if (cls == v) { vcl = br_vcls[v]; vcc = br_varc[v]; }
                                        if (vcc != 0.0) { acc_bb += (bn * bn) / vcl; acc_n += 1.0; }
                                        beta_g[j] = bn; delta_g[j] = cls + 1;
                                    }
                                    continue;
                                }
                                const bool in = out_i[os * 64 + 32 * i + lane] != 0;
                                acc_bb = fma(bn, bn, acc_bb);
                                if (method != 0) acc_n += in ? 1.0 : 0.0;
                                if (j < p_real) {
                                    beta_g[j] = bn;
                                    if (method != 0) delta_g[j] = in ? 1 : 0;
                                    if (method == 1) {                                      // functions.jl:182,186
                                        const double chi = reinterpret_cast<const double*>(recs + (gk & (unsigned)(NR - 1)) * L.rec_bytes + gofs)[F_CHI * B + qb[i]];
                                        vb_g[j] = in ? (sdf + bn * bn) / chi : 0.0;
                                    }
                                }
                            }
                            __syncwarp();
                            if (lane == 0) {
#pragma unroll
                                for (int k = 0; k < SB; ++k) mbar_arrive(&rec_free[(g0 + (unsigned)k) & (unsigned)(NR - 1)]);
                                mbar_arrive(&rb_free[(int)(sg & (RBS - 1))]);
                            }
                        }
                    }
                }
                __syncthreads();
                gblk += (unsigned)nblk;
            } else {
                // ============================ literal per-marker sweep ============================
                // every CTA (the chain CTA owns no rows) evaluates the scalar update redundantly; CTA Tw writes the outputs
                long long* lprev = reinterpret_cast<long long*>(misc + 48);
                const int nwords = R >> 2;
                const long long arrivals = arr_all;
                // BayesR (functions.jl:238-289): lane v of warp 0 owns variance class v
                const int nc = (S.method == 3) ? S.n_class : 0;
                double varc_v = 0.0, logpi_v = 0.0, vcls_v = 0.0;
                if (nc && lane < nc) {
                    vcls_v = S.v_class[lane];
                    varc_v = __ldcg(&S.varBeta[0]) * vcls_v;                            // functions.jl:244
                    logpi_v = __ldcg(&S.pi_class[nc + lane]);
                }
                // BayesRCpi / BayesRCplus (functions.jl:291-419): lane a * ncR + v of warp 0 owns (annotation a, class v)
                const bool rc_pi = S.method == 5, rc_plus = S.method == 6;
                const int nA = (rc_pi || rc_plus) ? S.n_annot : 0, ncR = (rc_pi || rc_plus) ? S.n_class : 1;
                const int a_l = lane / ncR, v_l = lane % ncR;
                const bool lane_on = nA && lane < nA * ncR;
                double rc_varc = 0.0, rc_logpi = 0.0, rc_vcls = 1.0;
                double rc_nloci = 0.0, rc_sumS = 0.0, rc_nnz = 0.0;          // nLoci[a][v] in lane (a,v); sumS[a], nNonZero[a] in lane (a,0)
                if (lane_on) {
                    rc_vcls = S.v_class[v_l];
                    rc_varc = __ldcg(&S.varBeta[a_l]) * rc_vcls;                         // functions.jl:298 / :369
                    rc_logpi = __ldcg(&S.pi_class[nA * ncR + lane]);
                }
                const double* ap_rd = nA ? S.annot_prob + (size_t)(iter & 1u) * S.p * nA : nullptr;         // annotProb as the previous iteration left it
                double* ap_wr = nA ? S.annot_prob + (size_t)((iter + 1u) & 1u) * S.p * nA : nullptr;
                for (int64_t j = 0; j < S.p; ++j, ++rk) {
                    const int k = (int)(j / B), q = (int)(j % B);
                    const uint8_t* tile = S.geno + ((int64_t)t * nblk + k) * L.tile_bytes;
                    const int slot = (int)(rk & (kSlots - 1));
                    long long* acc = sy->acc + (size_t)slot * kMaxB * kAccStride;
                    uint32_t w0 = 0;
                    double a = 0.0, dummy = 0.0;
                    const bool two = wt && (S.method == 1 || S.method == 2);      // BayesB/C: inclusion from the unweighted dot, mean from Mp_j'e
                    if (!is_chain) {
                        if (wt) {
                            // a = (x_j .* w)'e (Mp, mme.jl:303), dummy = x_j'e
                            for (int wr = tid; wr < nwords; wr += kThreads) {
                                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(tile) + word_off(B, q, wr));
                                if (wr == tid) w0 = w;
                                const double* ep = e_s + 4 * wr;
                                const double* wp = P.w + row0 + 4 * wr;
                                const double c0 = (double)(w & 0xff), c1 = (double)((w >> 8) & 0xff), c2 = (double)((w >> 16) & 0xff), c3 = (double)(w >> 24);
                                a = fma(c0 * __ldg(wp), ep[0], a); a = fma(c1 * __ldg(wp + 1), ep[1], a);
                                a = fma(c2 * __ldg(wp + 2), ep[2], a); a = fma(c3 * __ldg(wp + 3), ep[3], a);
                                dummy = fma(c0, ep[0], dummy); dummy = fma(c1, ep[1], dummy); dummy = fma(c2, ep[2], dummy); dummy = fma(c3, ep[3], dummy);
                            }
                        } else
                        for (int wr = tid; wr < nwords; wr += kThreads) {
                            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(tile) + word_off(B, q, wr));
                            if (wr == tid) w0 = w;
                            const double* ep = e_s + 4 * wr;
                            a = fma((double)(w & 0xff), ep[0], a); a = fma((double)((w >> 8) & 0xff), ep[1], a);
                            a = fma((double)((w >> 16) & 0xff), ep[2], a); a = fma((double)(w >> 24), ep[3], a);
                        }
                    }
                    block_sum2(a, dummy, misc);
                    if (tid == 0 && !is_chain) {
                        if (two) {
                            const double xu = dummy * fx_scale;
                            if (!(fabs(xu) < 9007199254740992.0)) atomicOr(&sy->err, 1);
                            red_add_u64(acc + kAccStride, (long long)((unsigned long long)__double2ll_rn(xu) << cb) + 1);
                        }
                        const double xs = a * fx_scale;
                        if (!(fabs(xs) < 9007199254740992.0)) atomicOr(&sy->err, 1);
                        const long long v = (long long)((unsigned long long)__double2ll_rn(xs) << cb) + 1;
                        if (sharded) {
                            // the per-marker scalar reduction over NVLink peer memory: one RED into every rank's accumulator
                            for (int r = 0; r < P.n_ranks; ++r) red_add_u64_sys(P.peer[r]->acc + (size_t)slot * kMaxB * kAccStride, v);
                        } else red_add_u64(acc, v);
                    }
                    if (warp == 0) {
                        const double* c = S.consts + (int64_t)k * (kNF * B) + q;
                        const double cA = __ldcg(c + F_A * B), cB = __ldcg(c + F_B * B), cT = __ldcg(c + F_T * B);
                        const double cC = __ldcg(c + F_C * B), cQ = __ldcg(c + F_QSZ * B), d = __ldcg(c + F_D * B);
                        const double bold = __ldcg(c + F_BOLD * B), mean = __ldcg(c + F_MEAN * B), chi = __ldcg(c + F_CHI * B);
                        double u_v = 0.0;                                               // BayesR: a fresh uniform per comparison (functions.jl:261)
                        if (nc && lane < nc) {
                            Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                            u_v = P.replay ? S.rp_u[(rp_row * S.p + j) * nc + lane] : stream_uniform(st, P_U, (uint32_t)j, 0, (uint32_t)lane);
                        }
                        long long* pv = lprev + slot;
                        long long cur;
                        if (sharded) {
                            unsigned spins = 0;
                            unsigned long long t0 = 0;
                            do {
                                cur = ld_relaxed_s64_sys(acc);
                                if ((++spins & 0x3ffu) == 0u && shard_wait_expired(&sy->err, t0)) break;
                            } while (((cur - *pv) & cmask) != arrivals);
                        }
                        else { do { cur = ld_relaxed_s64(acc); } while (((cur - *pv) & cmask) != arrivals); }
                        const double A = (double)((cur - *pv - arrivals) >> cb) * fx_inv;
                        __syncwarp();
                        if (lane == 0) *pv = cur;
                        const double r = A - mean * Stot;
                        const double rr = fma(d, bold, r);
                        double rq = rr;                                                 // the dot of the inclusion test
                        if (two) {
                            const long long pu = __shfl_sync(0xffffffffu, prev_u, slot);
                            long long cu;
                            do { cu = ld_relaxed_s64(acc + kAccStride); } while (((cu - pu) & cmask) != arrivals);
                            if (lane == slot) prev_u = cu;
                            const double Au = (double)((cu - pu - arrivals) >> cb) * fx_inv;
                            rq = fma(__ldg(&S.d_unw[j]), bold, Au - mean * Stot_u);       // functions.jl:168, :208: view(data,:,locus)'ycorr
                        }
                        if (nA) {
                            const unsigned FULL = 0xffffffffu;
                            Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                            const double iVarE = 1.0 / varE;
                            const double l0 = S.lhs0 ? S.lhs0[j] : 0.0, r0 = S.rhs0 ? S.rhs0[j] : 0.0;
                            const int an_l = lane_on ? __ldg(&S.annot[j * nA + a_l]) : 0;                    // annotInput of this lane's annotation
                            const int an_a = (lane < nA) ? __ldg(&S.annot[j * nA + lane]) : 0;               // ... and of annotation `lane`
                            const double lhs_l = (rc_varc == 0.0) ? 0.0 : d * iVarE + l0 + 1.0 / rc_varc;
                            double bn = 0.0;
                            int cls_out = 0;
                            if (rc_pi) {
                                const double rhs = fma(rr, iVarE, r0);                                                 // :304
                                double ex = 0.0;                                                                       // ExpLogL[a][v], zero for annotations the locus does not have
                                if (lane_on && an_l != 0) ex = exp((rc_varc == 0.0) ? rc_logpi : -0.5 * (log(rc_varc * lhs_l) - (rhs * rhs) / lhs_l) + rc_logpi);
                                double Sa = 0.0;                                                                       // sum(ExpLogL, dims = 2)[a] in class order
                                for (int v = 0; v < ncR; ++v) Sa += __shfl_sync(FULL, ex, (a_l * ncR + v) & 31);
                                const double pa1 = lane_on ? __ldcg(&ap_rd[j * nA + a_l]) * Sa : 0.0;                  // :315
                                double tot2 = 0.0;
                                for (int a = 0; a < nA; ++a) tot2 += __shfl_sync(FULL, pa1, a * ncR);
                                const double u_an = P.replay ? S.rp_u_annot[rp_row * S.p + j] : stream_uniform(st, P_U_ANNOT, (uint32_t)j);
                                // rand(Categorical(probAnnot)) :319 — one uniform; cp = p[1]; while cp <= draw && i < n: i += 1; cp += p[i]
                                int A = 0;
                                double cp = __shfl_sync(FULL, pa1, 0) / tot2;
                                for (int a = 1; a < nA; ++a) {
                                    const double pa = __shfl_sync(FULL, pa1, a * ncR) / tot2;
                                    if (A == a - 1 && cp <= u_an) { A = a; cp += pa; }
                                }
                                const int an_A = __shfl_sync(FULL, an_a, A);
                                if (!(tot2 == tot2) || an_A == 0) atomicOr(&sy->err, 4);                              // findfirst(isequal(A), nonzero) finds nothing: the reference errors
                                // sampleProb (:322, :541-544): Dirichlet(annotInput[nonzero] + e_A); lane a draws the gamma of annotation a
                                double g = 0.0;
                                if (!P.replay && lane < nA && an_a != 0) g = stream_gamma(st, P_G_ANNOT, (uint32_t)j, (uint32_t)lane, (double)an_a + (lane == A ? 1.0 : 0.0));
                                double tg = 0.0;
                                for (int a = 0; a < nA; ++a) tg += __shfl_sync(FULL, g, a);
                                if (is_chain && lane < nA && an_a != 0) ap_wr[j * nA + lane] = P.replay ? S.rp_dirp[(rp_row * S.p + j) * nA + lane] : g / tg;
                                // class inside the chosen annotation (:324-327): a fresh uniform per cumulative comparison
                                const double u_v = (lane < ncR) ? (P.replay ? S.rp_u[(rp_row * S.p + j) * ncR + lane] : stream_uniform(st, P_U, (uint32_t)j, 0, (uint32_t)lane)) : 0.0;
                                const double S_A = __shfl_sync(FULL, Sa, A * ncR);
                                double cum = 0.0;
                                int cls = -1;
                                for (int v = 0; v < ncR; ++v) {
                                    cum += __shfl_sync(FULL, ex, A * ncR + v) / S_A;
                                    const double uv = __shfl_sync(FULL, u_v, v);
                                    if (cls < 0 && cum >= uv) cls = v;
                                }
                                if (cls < 0) { atomicOr(&sy->err, 4); cls = 0; }
                                const int lc = A * ncR + cls;
                                const double varc_c = __shfl_sync(FULL, rc_varc, lc), lhs_c = __shfl_sync(FULL, lhs_l, lc), vcls_c = __shfl_sync(FULL, rc_vcls, lc);
                                if (varc_c != 0.0) {                                                                   // :333-341
                                    const double z = P.replay ? S.rp_z[rp_row * S.p + j] : stream_normal(st, P_Z, (uint32_t)j);
                                    bn = rhs / lhs_c + sqrt(1.0 / lhs_c) * z;
                                    if (lane == A * ncR) { rc_sumS += (bn * bn) / vcls_c; rc_nnz += 1.0; }
                                }
                                if (lane == lc) rc_nloci += 1.0;                                                       // nLoci[A, class] += 1
                                cls_out = cls;
                                if (is_chain && lane == 0) S.annot_cat[j] = A + 1;                                     // :330
                            } else {
                                // BayesRCplus: one effect per annotation of the locus, in annotation order; ycorr loses every one of them before the
                                // next is drawn (:379, :401):  x'(e - x b) = x'e - (x'x) b
                                double rr_run = rr, temp = 0.0;
                                for (int a = 0; a < nA; ++a) {
                                    if (__shfl_sync(FULL, an_a, a) == 0) continue;                                     // :378
                                    const double rhs = fma(rr_run, iVarE, r0);
                                    double ex = 0.0;
                                    if (lane_on && a_l == a) ex = exp((rc_varc == 0.0) ? rc_logpi : -0.5 * (log(rc_varc * lhs_l) - (rhs * rhs) / lhs_l) + rc_logpi);
                                    double tot = 0.0;
                                    for (int v = 0; v < ncR; ++v) tot += __shfl_sync(FULL, ex, a * ncR + v);
                                    const double u_v = (lane < ncR) ? (P.replay ? S.rp_u[((rp_row * S.p + j) * nA + a) * ncR + lane]
                                                                                : stream_uniform(st, P_U, (uint32_t)j, (uint32_t)a, (uint32_t)lane)) : 0.0;
                                    double cum = 0.0;
                                    int cls = -1;
                                    for (int v = 0; v < ncR; ++v) {
                                        cum += __shfl_sync(FULL, ex, a * ncR + v) / tot;
                                        const double uv = __shfl_sync(FULL, u_v, v);
                                        if (cls < 0 && cum >= uv) cls = v;
                                    }
                                    if (cls < 0) { atomicOr(&sy->err, 4); cls = 0; }
                                    const int lc = a * ncR + cls;
                                    const double varc_c = __shfl_sync(FULL, rc_varc, lc), lhs_c = __shfl_sync(FULL, lhs_l, lc), vcls_c = __shfl_sync(FULL, rc_vcls, lc);
                                    double b = 0.0;
                                    if (varc_c != 0.0) {                                                               // :391-397
                                        const double z = P.replay ? S.rp_z[(rp_row * S.p + j) * nA + a] : stream_normal(st, P_Z, (uint32_t)j, (uint32_t)a);
                                        b = rhs / lhs_c + sqrt(1.0 / lhs_c) * z;
                                        if (lane == a * ncR) { rc_sumS += (b * b) / vcls_c; rc_nnz += 1.0; }
                                    }
                                    if (lane == lc) rc_nloci += 1.0;
                                    temp += b;                                                                         // :400
                                    rr_run = fma(-d, b, rr_run);
                                    cls_out = cls;                                                                     // :388: the last annotation's class stays in delta
                                }
                                bn = temp;                                                                             // :403
                            }
                            if (lane == 0) {
                                misc[40] = bn - bold; misc[41] = mean;
                                if (is_chain) { S.beta[j] = bn; S.delta[j] = cls_out + 1; }
                            }
                        } else if (nc) {
                            // class likelihoods (functions.jl:250-258), one class per lane
                            const double iVarE = 1.0 / varE;
                            const double l0 = S.lhs0 ? S.lhs0[j] : 0.0, r0 = S.rhs0 ? S.rhs0[j] : 0.0;
                            const double rhs = fma(rr, iVarE, r0);
                            const double lhs_v = (varc_v == 0.0) ? 0.0 : d * iVarE + l0 + 1.0 / varc_v;
                            double ex = 0.0;
                            if (lane < nc) ex = exp((varc_v == 0.0) ? logpi_v : -0.5 * (log(varc_v * lhs_v) - (rhs * rhs) / lhs_v) + logpi_v);
                            double tot = 0.0;                                           // sum in class order, like sum(ExpLogL)
                            for (int v = 0; v < nc; ++v) tot += __shfl_sync(0xffffffffu, ex, v);
                            const double pr = ex / tot;
                            double cum = 0.0;                                           // cumsum(probs)
                            for (int v = 0; v < nc; ++v) { const double pv_ = __shfl_sync(0xffffffffu, pr, v); if (v <= lane) cum += pv_; }
                            const unsigned hit = __ballot_sync(0xffffffffu, lane < nc && cum >= u_v);
                            if (!hit) atomicOr(&sy->err, 4);                            // findfirst found nothing: the reference throws here
                            const int cls = hit ? (__ffs(hit) - 1) : 0;
                            const double varc_c = __shfl_sync(0xffffffffu, varc_v, cls);
                            const double lhs_c = __shfl_sync(0xffffffffu, lhs_v, cls);
                            const double vcls_c = __shfl_sync(0xffffffffu, vcls_v, cls);
                            const double bn = (varc_c != 0.0) ? rhs / lhs_c + sqrt(1.0 / lhs_c) * cQ : 0.0;      // functions.jl:266-268, :275
                            if (wt) Stot -= (bn - bold) * __ldg(&S.wcs[j]);               // 1'We after e -= x_j (bn - bold)
                            if (lane == cls) acc_cls += 1.0;                            // nLoci[classSNP] += 1
                            if (lane == 0) {
                                misc[40] = bn - bold; misc[41] = mean;
                                if (varc_c != 0.0) { acc_bb += (bn * bn) / vcls_c; acc_n += 1.0; }               // sumS, nNonZero
                                if (is_chain) { S.beta[j] = bn; S.delta[j] = cls + 1; }
                            }
                        } else {
                        const double dl = fma(cB, rq * rq, cA);
                        const bool in = dl < cT;
                        const double bn = in ? fma(rr, cC, cQ) : 0.0;
                        if (wt) Stot -= (bn - bold) * __ldg(&S.wcs[j]);                   // 1'We after e -= x_j (bn - bold)
                        if (lane == 0) {
                            misc[40] = bn - bold; misc[41] = mean;
                            acc_bb = fma(bn, bn, acc_bb);
                            if (S.method != 0) acc_n += in ? 1.0 : 0.0;
                            if (is_chain) {
                                S.beta[j] = bn;
                                if (S.method != 0) S.delta[j] = in ? 1 : 0;
                                if (S.method == 1) S.varBeta[j] = in ? (S.scale * S.df + bn * bn) / chi : 0.0;
                            }
                        }
                        }
                    }
                    __syncthreads();
                    const double db = misc[40];
                    if (db != 0.0 && !is_chain) {
                        const double K = db * misc[41];
                        for (int wr = tid; wr < nwords; wr += kThreads) {
                            const uint32_t w = (wr == tid) ? w0 : __ldg(reinterpret_cast<const uint32_t*>(tile) + word_off(B, q, wr));   // first pass from registers
                            double* ep = e_s + 4 * wr;
                            const int lim = nrow - 4 * wr;
                            if (lim > 0) ep[0] -= fma(db, (double)(w & 0xff), -K);
                            if (lim > 1) ep[1] -= fma(db, (double)((w >> 8) & 0xff), -K);
                            if (lim > 2) ep[2] -= fma(db, (double)((w >> 16) & 0xff), -K);
                            if (lim > 3) ep[3] -= fma(db, (double)(w >> 24), -K);
                        }
                    }
                    __syncthreads();
                }
                // (only lane 0 accumulated acc_bb / acc_n in the literal path; the other lanes hold 0)
                rc_nloci_keep = rc_nloci; rc_sumS_keep = rc_sumS; rc_nnz_keep = rc_nnz;
            }
            if constexpr (PROF) tc = clock64();

            // ------------------------------------------------------------------ phase 3 (chain CTA)
            const bool regional = (S.method == 0 && S.n_regions > 1) || S.method == 4;
            const int p3warp = LIT ? 0 : kHelperWarp;     // the warp that accumulated beta'beta and nLoci
            if (LIT && is_chain && warp == 0 && (S.method == 5 || S.method == 6)) {
                // per annotation: variance (functions.jl:347-349 / :408-410) and, with estimatePi, Dirichlet(nLoci[a, :] + 1) (:352-359 / :413-418)
                const int nA3 = S.n_annot, nc3 = S.n_class;
                const int a3 = lane / nc3, v3 = lane % nc3;
                const bool on3 = lane < nA3 * nc3;
                Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                if (on3 && v3 == 0) {
                    const double chi2 = P.replay ? S.rp_chi2b[rp_row * S.nvar + a3] : stream_chisq(st, P_CHI2_B, (uint32_t)a3, 0, S.df + rc_nnz_keep);
                    S.varBeta[a3] = (S.scale * S.df + rc_sumS_keep) / chi2;
                }
                if (S.est_pi) {
                    double gv = 0.0;
                    if (on3) gv = P.replay ? S.rp_betapi[(rp_row * nA3 + a3) * nc3 + v3] : stream_gamma(st, P_PI_A, (uint32_t)a3, (uint32_t)v3, rc_nloci_keep + 1.0);
                    double tg = 0.0;
                    for (int v = 0; v < nc3; ++v) tg += __shfl_sync(0xffffffffu, gv, (a3 * nc3 + v) & 31);
                    if (on3) {
                        const double ph = P.replay ? gv : gv / tg;
                        S.pi_class[lane] = ph; S.pi_class[nA3 * nc3 + lane] = log(ph);
                    }
                }
            } else if (is_chain && warp == p3warp && S.method == 3) {
                const double sumS = warp_sum(acc_bb);
                const double nnz = warp_sum(acc_n);
                const int nc = S.n_class;
                Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                if (lane == 0) {                    // functions.jl:281, 518-520
                    const double chi2 = P.replay ? S.rp_chi2b[rp_row * S.nvar] : stream_chisq(st, P_CHI2_B, 0, 0, S.df + nnz);
                    S.varBeta[0] = (S.scale * S.df + sumS) / chi2;
                }
                if (S.est_pi) {                     // functions.jl:284-288, 536-538: Dirichlet(nLoci .+ 1) = normalised gammas
                    double gv = 0.0;
                    if (lane < nc) gv = P.replay ? S.rp_betapi[rp_row * nc + lane] : stream_gamma(st, P_PI_A, 0, (uint32_t)lane, acc_cls + 1.0);
                    double tg = 0.0;
                    for (int v = 0; v < nc; ++v) tg += __shfl_sync(0xffffffffu, gv, v);
                    if (lane < nc) {
                        const double ph = P.replay ? gv : gv / tg;
                        S.pi_class[lane] = ph; S.pi_class[nc + lane] = log(ph);
                    }
                }
            } else if (is_chain && warp == p3warp && !regional) {
                const double bb = warp_sum(acc_bb);
                const double nl = warp_sum(acc_n);
                if (lane == 0) {
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
                // beta / delta of this sweep were written by the chain CTA (plain stores): one grid barrier, then the
                // posterior sums and the region variances are spread over the whole grid
                __syncthreads();
                gs.nbar++;
                if (tid == 0) gs.arrive(P);
                if (warp == 0) gs.wait_warp();
                __syncthreads();
            }
            if (P.accumulate) {
                for (int64_t j = (int64_t)t * kThreads + tid; j < S.p; j += (int64_t)(Tw + 1) * kThreads) {
                    const double bj = __ldcg(&S.beta[j]);
                    S.sum_beta[j] += bj;
                    S.sum_beta2[j] = fma(bj, bj, S.sum_beta2[j]);
                    S.sum_delta[j] += (S.method == 0) ? 1.0 : (double)__ldcg(&S.delta[j]);
                }
            }
            if (TUP && S.method == 4) {
                // Sigma_r ~ InvWishart(df + |r|, scale + B_r'B_r) (functions.jl:152, 513-516), one warp per region; effects are interleaved
                const int k = S.group_k;
                Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)S.stream_set};
                for (int64_t rg = (int64_t)t * kWarps + warp; rg < S.n_regions; rg += (int64_t)(Tw + 1) * kWarps) {
                    const int64_t c0 = S.region_off[rg], c1 = S.region_off[rg + 1];          // column offsets = locus offsets * k
                    double Sb[16];
                    for (int x = 0; x < k * k; ++x) Sb[x] = 0.0;
                    for (int64_t c = c0 + (int64_t)lane * k; c < c1; c += 32 * k) {
                        double bj[4];
                        for (int b = 0; b < k; ++b) bj[b] = __ldcg(&S.beta[c + b]);
                        for (int a = 0; a < k; ++a)
                            for (int b = a; b < k; ++b) Sb[a * k + b] = fma(bj[a], bj[b], Sb[a * k + b]);
                    }
                    for (int a = 0; a < k; ++a)
                        for (int b = a; b < k; ++b) { const double v = warp_sum(Sb[a * k + b]); Sb[a * k + b] = v; Sb[b * k + a] = v; }
                    if (lane == 0) {
                        double Psi[16], chi2[4], zl[16], Sg[16];
                        const double dfr = S.df + (double)((c1 - c0) / k);
                        for (int x = 0; x < k * k; ++x) { Psi[x] = S.jscale[x] + Sb[x]; zl[x] = 0.0; }
                        for (int i = 0; i < k; ++i) {
                            chi2[i] = P.replay ? S.rp_iw_chi2[(rp_row * S.n_regions + rg) * k + i]
                                               : stream_chisq(st, P_IW, (uint32_t)rg, (uint32_t)(i * k + i), dfr - (double)i);
                            for (int jj = 0; jj < i; ++jj)
                                zl[i * k + jj] = P.replay ? S.rp_iw_z[((rp_row * S.n_regions + rg) * k + i) * k + jj]
                                                          : stream_normal(st, P_IW, (uint32_t)rg, 0, (uint32_t)(i * k + jj));
                        }
                        const bool ok = (k == 2) ? jt_inv_wishart<2>(Psi, chi2, zl, Sg) : jt_inv_wishart<4>(Psi, chi2, zl, Sg);
                        if (ok) { for (int x = 0; x < k * k; ++x) S.jvar[rg * k * k + x] = Sg[x]; }
                        else atomicOr(&sy->err, 2);
                    }
                }
            } else if (regional) {
                Stream st{P.key0, P.key1, P.chain, iter, (uint32_t)s};
                for (int64_t rg = (int64_t)t * kWarps + warp; rg < S.n_regions; rg += (int64_t)(Tw + 1) * kWarps) {
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

        if (is_chain && tid == 0) {
            P.sc->mu = mu; P.sc->varE = varE; P.sc->iter = iter0 + it + 1;
            if (P.accumulate) P.sc->n_post += 1;
        }
    }   // iterations

    __syncthreads();
    for (int r = tid; r < nrow; r += kThreads) P.e[row0 + r] = e_s[r];
    // profile: thread 0 of every CTA (worker warp 0 / chain warp), plus the first prep warp of the chain CTA
    if constexpr (PROF) {
        // thread 0 = updater (worker CTA) / chain warp (chain CTA); first dot warp; first prep warp
        const bool dotlane = !is_chain && tid == kFirstDotWarp * 32, preplane = is_chain && tid == kFirstPrepWarp * 32;
        for (int i = 0; i < kProf; ++i) {
            const bool dot_i = (i == 1 || i == 14 || i == 15 || i == 20 || i == 28 || i == 29), poll_i = (i == 24 || i == 25), prep_i = (i == 12 || i == 13 || (i >= 21 && i <= 23) || i >= 30);
            if (tid == 0 && !((dot_i || poll_i) && !is_chain) && !(prep_i && is_chain)) sy->prof[t * kProf + i] = pf[i];
            if (!is_chain && tid == kPollWarp * 32 && poll_i) sy->prof[t * kProf + i] = pf[i];
            if (dotlane && dot_i) sy->prof[t * kProf + i] = pf[i];
            if (preplane && prep_i) sy->prof[t * kProf + i] = pf[i];
        }
    }
#undef NGP_TICK
}

template <int B, bool PROF, bool DBG, bool LIT, bool TUP, bool BIGR = false, bool BR = false, bool SH = false, bool SHT = false>
__global__ void __launch_bounds__(kThreads, 1) gibbs_kernel(const Params P)
{
    gibbs_body<B, PROF, DBG, LIT, TUP, BIGR, BR, SH, SHT>(P, (int)blockIdx.x);
}

// All ranks of a row-sharded chain whose shards live on ONE device, as ONE cooperative grid (the only legal way to run kernels that wait
// for one another on one GPU: separate launches are not guaranteed to be co-resident).  CTA b belongs to the rank r with
// cta_off(r) <= b < cta_off(r) + Tw(r) + 1 and runs that rank's per-marker sweep on that rank's Params — the same code, the same
// system-scope REDs and polls as between GPUs.
struct GroupParams { const Params* ranks; int n_ranks; };
template <int B, bool LIT>
__global__ void __launch_bounds__(kThreads, 1) gibbs_group_kernel(const GroupParams G)
{
    __shared__ Params Ps;
    __shared__ int rsel;
    if (threadIdx.x == 0) {
        int r = 0;
        while (r + 1 < G.n_ranks && (int)blockIdx.x >= G.ranks[r + 1].cta_off) ++r;
        rsel = r;
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(&G.ranks[rsel]);
        int* dst = reinterpret_cast<int*>(&Ps);
        for (int i = threadIdx.x; i < (int)(sizeof(Params) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    gibbs_body<B, false, false, LIT, false, false, false, !LIT>(Ps, (int)blockIdx.x - Ps.cta_off);
}

}  // namespace ngp
