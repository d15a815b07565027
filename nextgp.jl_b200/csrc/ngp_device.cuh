// ngp_device.cuh — device-side data structures and small PTX wrappers of libngp.
//
// HBM layout of one marker set ("row-panelled, block-tiled, fragment-ordered"):
//   the n individuals are cut into Tw row panels of R rows (R = multiple of 32), panel t is owned by worker CTA t
//   of the persistent sweep kernel; the p markers are cut into blocks of B.  The B x R codes of (panel t, block k)
//   are ONE contiguous B*R-byte tile
//       geno[(t * nblk + k) * B * R + byte_off(B, q, r)]      marker k*B + q, row t*R + r
//   that a single cp.async.bulk (TMA) brings into shared memory, so every column is read from HBM exactly
//   once per sweep.  Inside a tile the bytes are stored in the operand order of the INT8 tensor-core
//   instruction mma.sync.m16n8k32 (A = 16 markers x 32 rows, row-major): the four 32-bit A registers of a lane
//   are 16 contiguous bytes, so the dot phase loads its operands with one conflict-free LDS.128 per MMA.
//   A stored byte is the plain code 0/1/2.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "ngp_rng.cuh"

namespace ngp {

// ---- CTA geometry of the sweep kernel: worker CTAs 0..Tw-1 own row panels, CTA Tw runs the scalar chain
constexpr int kUpdWarps = 4;           // worker CTA: warps 0..3 "updaters": residual rows in registers, axpy + re-quantise
constexpr int kUpdThreads = kUpdWarps * 32;
constexpr int kUpdGroups = 4;          //             4-row groups per updater thread: R <= 4 * kUpdThreads * kUpdGroups = 2048
constexpr int kDotWarps = 8;           //             warps 4..11 "dot warps": each owns whole blocks (IMMA dots + RED)
constexpr int kFirstDotWarp = kUpdWarps;
constexpr int kTmaWarp = 12;           //             warp 12 lane 0: TMA producer of the tile ring
constexpr int kPollWarp = 13;          //             warp 13: receives the changed-effect lists of the chain CTA
constexpr int kWarps = 14;
constexpr int kThreads = kWarps * 32;  // chain CTA: warp 0 chain, warp 1 TMA producer of the block records, warps 2..9 "prep" warps
constexpr int kPrepWarps = 8;          //            (poll accumulators, cross-Gram corrections of distance >= 2 -> r_base)
constexpr int kFirstPrepWarp = 2;
constexpr int kLimbVers = 16;          // maximum number of versions of the fixed-point residual kept per worker CTA

constexpr int kMaxClass = 8;           // BayesR variance classes
constexpr int kMaxFxCols = 32;         // columns of all fixed-effect sets besides the intercept
constexpr int kMaxFxSets = 4;
constexpr int kMaxRanks = 8;           // row shards of one chain (GPUs of one NVSwitch box)
constexpr int kProf = 32;              // cycle counters per CTA (ngp_get_profile)
constexpr int kMaxB = 64;              // markers per block: 16, 32 or 64
constexpr int kNF = 10;                // per-marker constant fields
constexpr int kMaxD = 24;              // maximum look-ahead depth (blocks)
constexpr int kSlots = 32;             // accumulator ring (blocks in flight <= D+1), multiple of kPrepWarps
constexpr int kMaxCtas = 1280;         // CTAs of all ranks of a row-sharded chain (8 ranks x 160); one rank: <= 160
constexpr int kCntBits = 8;            // low bits of an accumulator that count arrivals (Params::cnt_bits: 8, or 11 when more than 255 worker CTAs of all ranks add into it)
constexpr int kAccStride = 1;          // int64 units between two accumulators (contiguous measured no slower than 256 B apart)
constexpr int kLLCopies = 8;           // replicas of the changed-effect list ring (worker CTA t polls replica t % 8): no L2 hot spot
constexpr int kNzRing = 32;            // rings indexed by the global block number (lists, versions, dot-done flags): >= D+2
constexpr int kNzSmem = 8;             // worker-CTA smem ring of received lists
constexpr int kRecStages = 8;          // chain CTA: maximum stages of the block-record ring (TMA)
constexpr int kPollPipe = 1;           // polling loops (lists, accumulators): probes kept in flight
constexpr int kPollGap = 0;            //                                          cycles between two probes
constexpr int kLLEntryWords = 5;       // idx, dbeta lo/hi, K lo/hi
constexpr int kLLSlotWords = 1 + kLLEntryWords * kMaxB + 3;   // header + entries, padded to a multiple of 4 words

// fields of the per-iteration marker constants, stored [block][field][B]
enum Field { F_A = 0, F_B = 1, F_T = 2, F_C = 3, F_QSZ = 4, F_D = 5, F_BOLD = 6, F_MEAN = 7, F_CS = 8, F_CHI = 9 };

struct SetDev {
    int64_t p, p_pad;
    int32_t method, est_pi;
    int64_t n_regions, nvar;
    double df, scale;
    const uint8_t* geno;       // [Tw][p_pad/B] tiles of B*R bytes
    const int32_t* gx;         // [p_pad/B][D+1][B][B]: gx[k][d][a][b] = raw sum_i g_a g_b, a in block k-d, b in block k (zeros for k < d)
    double* consts;            // [p_pad/B][kNF][B] per-iteration marker constants (phase 1)
    const int32_t* colsum;     // [p_pad]
    const double* d;           // [p_pad] mpm
    const double* mean;        // [p_pad]
    double* beta;              // [p_pad]
    int32_t* delta;            // [p_pad]
    double* varBeta;           // [nvar]
    double* pi;                // [4] piHat0 piHat1 logPi0 logPi1
    const int32_t* region_of;  // [p_pad] (BayesPR with >1 region) or null
    const int64_t* region_off; // [n_regions+1] or null
    const double* lhs0;        // [p] or null
    const double* rhs0;        // [p] or null
    const double* rp_u;        // replay arrays (device) or null
    const double* rp_z;
    const double* rp_chi2b;
    const double* rp_betapi;
    int32_t n_class, pad_cls;  // BayesR (method 3): variance classes (functions.jl:241)
    double v_class[kMaxClass]; //   M.vClass
    double* pi_class;          //   [2 * n_class] piHat, logPi (mme.jl:375,383)
    // BayesRCpi / BayesRCplus (methods 5, 6; per-marker kernel): annotations.  pi_class = [n_annot][n_class] piHat then [n_annot][n_class] logPi
    int32_t n_annot, pad_an;
    const int32_t* annot;      // [p][n_annot] annotInput (mme.jl:394)
    double* annot_prob;        // [2][p][n_annot] annotProb (mme.jl:395), double-buffered by iteration parity: every CTA reads the values of the
                               // previous iteration while the chain CTA writes this iteration's Dirichlet draws
    int32_t* annot_cat;        // [p] annotCat (RCpi)
    const double* rp_u_annot;  // replay [iter][p]
    const double* rp_dirp;     // replay [iter][p][n_annot]
    // tuple of k marker sets swept by the blocked kernel (method 4): the breeds' columns are INTERLEAVED, column j*k + b = breed b
    // of locus j, and the k effects of a locus are drawn jointly (functions.jl:140-154)
    int32_t group_k, stream_set;   // k (2 or 4; 0 = ordinary set); set id that addresses the variate stream (first member)
    double jscale[16];             // k x k prior scale (mme.jl:501)
    double* jvar;                  // [n_regions][k][k] covariances (varBeta of the tuple)
    double* jinvB;                 // [n_regions][k][k] their inverses, refreshed in phase 1
    const double* rp_iw_chi2;      // replay of the inverse-Wishart draw: [iter][n_regions][k]
    const double* rp_iw_z;         //                                     [iter][n_regions][k][k]
    // weighted residuals (E.str == "D", mme.jl:299-303; per-marker kernel only): d then points at the WEIGHTED mpm
    const double* d_unw;       // [p_pad] unweighted mpm x_j'x_j (the BayesB/C inclusion dot is unweighted: functions.jl:168, :208) or null
    const double* wcs;         // [p_pad] sum_i w_i (g_ij - mean_j): change of sum_i w_i e_i per unit change of the effect
    double* sum_beta;          // posterior sums [p_pad]
    double* sum_beta2;
    double* sum_delta;
};

// fixed effects with one or several columns (covariates, factor levels): X[xSet] of getMME!, sampled after the intercept
// (samplers.jl:37-39, functions.jl:22-54)
struct FxDev {
    int32_t n_sets, n_cols;              // sets, columns of all sets
    int32_t first[kMaxFxSets + 1];       // columns [first[s], first[s+1]) belong to set s
    int32_t xoff[kMaxFxSets];            // offset of the set's c x c block in xpx
    const double* data;                  // [n_cols][Tw*R] column-major, pad rows zero
    const double* xpx;                   // X'X per set (X[xSet].xpx)
    const double* colsum;                // [n_cols] 1'x_c: keeps 1'e current
    const double* colsum_w;              // [n_cols] w'x_c (weighted residuals: keeps 1'We current; xpx is then X'WX, mme.jl:135) or null
    double* b;                           // [n_cols] the effects
    double lhs0[kMaxFxSets], rhs0[kMaxFxSets];   // single-column sets: X[xSet].lhs / .rhs
    const double* rp_z;                  // replay [iter][n_cols]
};

struct SyncArea {
    // ---- head: zeroed before every launch (counter, error word, then the phase-0 partials of the CTAs in use)
    unsigned long long counter;                 // grid-barrier arrivals, monotonic within a launch
    unsigned long long pad0[15];
    int err;
    int pad1[31];
    double part[kMaxCtas * 2];                  // phase-0 partials (e'e, sum e) per CTA
    // ---- body: persists across launches (monotonic accumulators, sequence-numbered list words)
    long long acc[kSlots * kMaxB * kAccStride]; // fixed-point reduction accumulators (monotonic; low bits count arrivals)
    long long acc2[kSlots * kMaxB];             // row-sharded chain, rank-local pre-reduction: this rank's worker CTAs add here, its prep warps forward the total
    unsigned long long ll[kLLCopies][kNzRing * kLLSlotWords];   // changed-effect lists chain CTA -> worker CTAs (copy = CTA % kLLCopies);
                                                // every 8-byte word = {payload32, seq32}, seq = global block number + 1
    long long prof[160 * kProf];                // per-CTA cycle counters of the last launch (this rank's CTAs): see ngp_get_profile
    long long trace[2 * 2048];                  // instrumented kernel: (start clock, cycles waited for r_base) of the chain warp's first 2048 steps
    double part_fx[160 * kMaxFxCols];           // per-CTA partial dots x_c'e of the fixed-effect columns (not with row sharding)
    double part_u[160];                         // weighted residuals: per-CTA partial of the plain 1'e (phase 0; not with row sharding)
};
constexpr size_t kSyncHeadFixed = 16 * 8 + 32 * 4;                       // counter + error word; the launch also clears part[0 .. 2 T_all)
constexpr size_t kSyncHeadBytes = kSyncHeadFixed + kMaxCtas * 2 * 8;

struct Scalars {          // device-resident chain scalars
    double mu, varE;
    long long iter;       // iterations completed
    long long n_post;     // posterior samples accumulated
};

struct Params {
    int64_t n;
    int32_t Tw;                // worker CTAs (row panels); the grid has Tw + 1 CTAs
    int32_t R, B, n_sets, kernel;
    int32_t NV;                // versions of the fixed-point residual per worker CTA (<= kLimbVers)
    int32_t D, DN, NT, NR;     // look-ahead depth (blocks), near depth (cross-Grams kept in the block record), tile ring stages,
                               // block-record ring stages of the chain CTA (power of two <= kRecStages)
    double* e;                 // [Tw*R]
    const SetDev* sets;        // device array
    Scalars* sc;
    SyncArea* sync;
    double df_e, scale_e;
    int32_t has_mu, do_varE, do_mu, set_mask;
    double mu_lhs0, mu_rhs0;
    double varE_in;            // used when do_varE == 0
    int32_t n_iter, replay;
    int64_t replay_base;       // iteration number of replay row 0 minus 1
    const double* rp_chi2_e;
    const double* rp_z_mu;
    uint32_t key0, key1, chain;
    int32_t accumulate;        // add to posterior sums
    uint32_t gblk0;            // global number of the first block of this launch (lists and accumulator slots are numbered globally)
    int32_t debug;             // timing experiments (NGP_CFG_DEBUG), see ngp_sweep.cuh
    int32_t opt;               // schedule options (NGP_CFG_OPT): 1 tiles streamed evict-first, 2 far Gram rows prefetched at change time
    // ---- row-sharded chain: rank r holds rows [row_0(r), row_0(r) + n) of X and e; every grid-wide quantity
    //      (barrier arrivals, phase-0 partials, per-marker fixed-point sums) is PUSHED into the SyncArea of every rank over
    //      NVLink peer memory, and every poll is local.  n_ranks == 1: not sharded (peer[0] == sync).
    int32_t n_ranks, rank;
    int32_t cta_off, T_all;    // index of this rank's first CTA in the all-rank CTA numbering; CTAs of all ranks
    int32_t Tw_all;            // worker CTAs of all ranks (arrivals per accumulator)
    int32_t cnt_bits;          // arrival-count bits of an accumulator: 8, or 11 when Tw_all > 255
    int32_t hier;              // row-sharded blocked sweep: 1 = the worker CTAs of a rank reduce locally (acc2) and the rank's prep warps push ONE total per
                               // marker to every rank (n_ranks arrivals per accumulator instead of Tw_all: many ranks are atomic-throughput-bound otherwise)
    int32_t store2;            // 1: genotypes stored as 2-bit codes (NGP_STORE_2BIT), expanded to the INT8 operands on chip
    int32_t refetch;           // 1: tiles leave shared memory once their dots are formed; the columns of changed effects are re-read from L2/HBM
    int64_t n_total;           // individuals over all ranks (n is the local row count)
    unsigned long long bar_base;   // barrier arrivals counted before this launch (sharded: the counter is never reset)
    SyncArea* peer[kMaxRanks];
    FxDev fx;
    // weighted residuals: w = E.iVarStr = inv.(D) (mme.jl:73), zero beyond row n; null = "I"
    const double* w;           // [Tw*R]
    double w_sum, w_min, w_max;
};

// ----------------------------------------------------------------------------- tile layout
// 32-bit word (rows 4*wr .. 4*wr+3 of the panel) of marker q of a block, in units of words.
// mma.m16n8k32 A fragment: lane = 4*g + t holds a0 = A[g][4t..4t+3], a1 = A[g+8][4t..], a2 = A[g][16+4t..], a3 = A[g+8][16+4t..]
// The same word is a B-fragment register of the transposed product (Gram set-up kernel): b0/b1 = rows 4t.. / 16+4t.. of column g.
__host__ __device__ __forceinline__ int word_off(int B, int q, int wr)
{
    const int c = wr >> 3, wq = wr & 7;                 // chunk of 32 rows, word inside the chunk
    const int mg = q >> 4, qq = q & 15, g = qq & 7, hi8 = qq >> 3;
    const int a_idx = ((wq >> 2) << 1) | hi8, t = wq & 3;
    return (((c * (B >> 4) + mg) * 32 + g * 4 + t) << 2) | a_idx;
}
__host__ __device__ __forceinline__ int byte_off(int B, int q, int r) { return (word_off(B, q, r >> 2) << 2) | (r & 3); }

// 2-bit device storage (NGP_STORE_2BIT): the same tile with every 32-bit word (the codes of 4 consecutive rows of one marker) squeezed
// into ONE byte, code of row 4w + i in bits 2i .. 2i+1.  The byte index of (marker q, rows 4 wr ..) is word_off(B, q, wr): a lane's four
// A registers of an MMA atom become the four bytes of ONE 32-bit word, expanded on chip.
//   expand2(x): byte of 4 codes -> 4 bytes (c0 | c1 << 8 | c2 << 16 | c3 << 24).  The low and the high nibble are moved 16 bits apart, so
//   that the shifted copies of the multiplication by (1 + 2^6) neither overlap nor carry.
__host__ __device__ __forceinline__ uint32_t expand2(uint32_t x)
{
    const uint32_t v = (x & 0x0Fu) | ((x & 0xF0u) << 12);
    return (v * 0x41u) & 0x03030303u;
}
__host__ __device__ __forceinline__ uint4 expand2_word(uint32_t w)
{
    uint4 a;
    a.x = expand2(w & 0xffu); a.y = expand2((w >> 8) & 0xffu); a.z = expand2((w >> 16) & 0xffu); a.w = expand2(w >> 24);
    return a;
}
__host__ __device__ __forceinline__ int64_t tile_bytes_of(int B, int R, int store2) { return store2 ? ((int64_t)B * R) >> 2 : (int64_t)B * R; }

__host__ __device__ __forceinline__ int consts_bytes(int B) { return kNF * B * 8; }
__host__ __device__ __forceinline__ int gram_bytes(int B) { return B * B * 4; }

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void red_add_u64(long long* addr, long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void arrive_release(unsigned long long* c)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(c) : "memory");
}
// system-scope variants for peer memory over NVLink (row-sharded chain)
__device__ __forceinline__ void red_add_u64_sys(long long* addr, long long v)
{
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void arrive_release_sys(unsigned long long* c)
{
    asm volatile("red.release.sys.global.add.u64 [%0], 1;" ::"l"(c) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_s64_sys(const long long* p)
{
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p)     // LDG.STRONG.GPU, no L1, no fence
{
    long long v;
    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "NGP_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NGP_DONE_%=;\n\t"
        "bra NGP_WAIT_%=;\n\t"
        "NGP_DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 cache policy (the genotype stream is read once per sweep: evict-first keeps it from displacing the Gram band)
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// INT8 tensor-core MMA, D(16x8,s32) += A(16x32,s8,row) * B(32x8,s8,col)   (SASS: IMMA.16832.S8.S8)
__device__ __forceinline__ void imma16832(int (&c)[4], const uint4& a, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// A ring cursor: stage index + mbarrier phase parity, advanced once per use
struct Ring {
    int s;
    uint32_t ph;
    __device__ __forceinline__ void adv(int n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

}  // namespace ngp
