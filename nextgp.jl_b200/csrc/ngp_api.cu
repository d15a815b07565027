// ngp_api.cu — C ABI of libngp.so (include/ngp.h) and the set-up kernels:
// genotype packing into the row-panelled layout, column statistics (mean, mpm =
// X_j'X_j, replacing /root/reference/src/mme.jl:305-307), block Gram matrices,
// synthetic-genotype generator, unpacking for round-trip tests.
//
// There is no CPU code path for any sampler arithmetic in this library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <unistd.h>

#include "../../include/ngp.h"
#include "ngp_sweep.cuh"
#include "ngp_joint.cuh"
#include "ngp_kernels.h"

using namespace ngp;

static_assert(kSyncHeadBytes == offsetof(SyncArea, acc) && kSyncHeadFixed == offsetof(SyncArea, part), "the per-launch memset covers the head of SyncArea: counter, error word, the partials of the CTAs in use");

// ============================================================================= set-up kernels
namespace {

__device__ __forceinline__ int block_sum_int(int v, int* scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += scratch[i];
    __syncthreads();
    return s;
}

// one CTA per column of the chunk: validate, scatter into the tiled fragment-ordered layout, column sums
template <int FMT>
__global__ void pack_kernel(const void* __restrict__ src, int64_t ld, int64_t n, int64_t j0, int R, int B, int64_t nblk,
                            uint8_t* __restrict__ geno, int32_t* __restrict__ colsum, int32_t* __restrict__ colsumsq,
                            int* __restrict__ err, int64_t jmul = 1, int64_t jadd = 0, int store2 = 0)
{
    __shared__ int scratch[32];
    const int64_t jc = blockIdx.x, j = (j0 + jc) * jmul + jadd;      // destination column (jmul, jadd: interleaving of a tuple's sets)
    const int64_t k = j / B;
    const int q = (int)(j - k * B);
    const int64_t tile_bytes = tile_bytes_of(B, R, store2);
    int s = 0, ss = 0, bad = 0;
    auto code_at = [&](int64_t i) -> int {
        int g;
        if (FMT == NGP_GENO_I8) {
            g = reinterpret_cast<const int8_t*>(src)[jc * ld + i];
        } else if (FMT == NGP_GENO_F64) {
            const double x = reinterpret_cast<const double*>(src)[jc * ld + i];
            g = (x == 0.0) ? 0 : (x == 1.0) ? 1 : (x == 2.0) ? 2 : -1;
        } else {
            g = (reinterpret_cast<const uint8_t*>(src)[jc * ld + (i >> 2)] >> (2 * (int)(i & 3))) & 3;
        }
        if (g < 0 || g > 2) { bad = 1; g = 0; }
        s += g; ss += g * g;
        return g;
    };
    if (store2) {
        // a thread owns the 4 consecutive rows of one stored byte (R is a multiple of 32: they lie in one panel)
        for (int64_t i4 = (int64_t)threadIdx.x * 4; i4 < n; i4 += (int64_t)blockDim.x * 4) {
            uint32_t byte = 0;
            for (int u = 0; u < 4 && i4 + u < n; ++u) byte |= (uint32_t)code_at(i4 + u) << (2 * u);
            const int64_t t = i4 / R;
            const int r = (int)(i4 - t * R);
            geno[(t * nblk + k) * tile_bytes + word_off(B, q, r >> 2)] = (uint8_t)byte;
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const int g = code_at(i);
            const int64_t t = i / R;
            const int r = (int)(i - t * R);
            geno[(t * nblk + k) * tile_bytes + byte_off(B, q, r)] = (uint8_t)g;
        }
    }
    s = block_sum_int(s, scratch);
    ss = block_sum_int(ss, scratch);
    if (bad) atomicOr(err, 1);
    if (threadIdx.x == 0) { colsum[j] = s; colsumsq[j] = ss; }
}

__global__ void colstats_kernel(int64_t n, int64_t p, int64_t p_pad, const int32_t* colsum, const int32_t* colsumsq,
                                double* mean, double* d)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p_pad) return;
    if (j < p) {
        const double s = (double)colsum[j], ss = (double)colsumsq[j], nn = (double)n;
        mean[j] = s / nn;                         // mean(thisM,dims=1), prepMatVec.jl:129
        d[j] = (nn * ss - s * s) / nn;            // sum (g - mean)^2 = dot(c,c), mme.jl:306 (integer-exact numerator)
    } else { mean[j] = 0.0; d[j] = 0.0; }
}

// raw banded Gram over all panels, one CTA per block k of B markers, on the INT8 tensor cores (integer-exact):
//   gx[k][d][a][b] = sum_i g_{(k-d)B+a,i} g_{kB+b,i},   d = 0..D   (zeros for k < d)
// A work item is one 16x8 output tile (d, 16 markers a, 8 markers b); both operands come straight from the
// fragment-ordered tiles (the A atom is 512 contiguous bytes; a B register is one tile word of the other block).
template <int B>
__global__ void __launch_bounds__(256) gram_kernel(const uint8_t* __restrict__ geno, int Tw, int R, int64_t nblk, int D,
                                                   int32_t* __restrict__ gx, int store2)
{
    constexpr int MG = B / 16, NBG = B / 8;
    const int64_t k = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tt = lane & 3;
    const int nchunk = R >> 5;
    const int64_t tile_bytes = tile_bytes_of(B, R, store2);
    const int items = (D + 1) * MG * NBG;
    for (int item = warp; item < items; item += 8) {
        const int d = item / (MG * NBG), rem = item - d * (MG * NBG), mg = rem / NBG, nb = rem - mg * NBG;
        int acc[4] = {0, 0, 0, 0};
        if (k - d >= 0) {
            const int qb = nb * 8 + g;
            for (int t = 0; t < Tw; ++t) {
                const uint8_t* ta = geno + ((int64_t)t * nblk + (k - d)) * tile_bytes;
                const uint32_t* tb = reinterpret_cast<const uint32_t*>(geno + ((int64_t)t * nblk + k) * tile_bytes);
                if (store2) {
                    const uint8_t* tb8 = reinterpret_cast<const uint8_t*>(tb);
#pragma unroll 4
                    for (int c = 0; c < nchunk; ++c) {
                        const uint4 a = expand2_word(__ldg(reinterpret_cast<const uint32_t*>(ta) + (size_t)(c * MG + mg) * 32 + lane));
                        const uint32_t b0 = expand2((uint32_t)__ldg(tb8 + word_off(B, qb, 8 * c + tt)));
                        const uint32_t b1 = expand2((uint32_t)__ldg(tb8 + word_off(B, qb, 8 * c + 4 + tt)));
                        imma16832(acc, a, b0, b1);
                    }
                } else {
#pragma unroll 4
                for (int c = 0; c < nchunk; ++c) {
                    const uint4 a = __ldg(reinterpret_cast<const uint4*>(ta + ((size_t)(c * MG + mg) * 32 + lane) * 16));
                    const uint32_t b0 = __ldg(tb + word_off(B, qb, 8 * c + tt));
                    const uint32_t b1 = __ldg(tb + word_off(B, qb, 8 * c + 4 + tt));
                    imma16832(acc, a, b0, b1);
                }
                }
            }
        }
        int32_t* out = gx + ((k * (D + 1) + d) * B) * (int64_t)B;
        const int a0 = mg * 16 + g, b0i = nb * 8 + 2 * tt;
        out[a0 * B + b0i] = acc[0]; out[a0 * B + b0i + 1] = acc[1];
        out[(a0 + 8) * B + b0i] = acc[2]; out[(a0 + 8) * B + b0i + 1] = acc[3];
    }
}

__global__ void synth_kernel(uint32_t key0, uint32_t key1, int64_t n, int64_t row0, int64_t j0, int64_t ncols,
                             const uint32_t* __restrict__ thr0, const uint32_t* __restrict__ thr1, int8_t* __restrict__ out)
{
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t jc = blockIdx.y;
    if (i4 >= n || jc >= ncols) return;
    const int64_t j = j0 + jc;
    uint32_t w[4];
    philox4x32_10((uint32_t)((row0 + i4) >> 2), (uint32_t)j, 0u, 0x47454e4fu, key0, key1, w);     // row0 is a multiple of 4
    const uint32_t a = thr0[j], b = thr1[j];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (i4 + q < n) out[jc * n + i4 + q] = (int8_t)((w[q] >= a) + (w[q] >= b));
}

__global__ void unpack_kernel(const uint8_t* __restrict__ geno, int64_t n, int R, int B, int64_t nblk, int64_t j0, int64_t ncols,
                              int8_t* __restrict__ out, int store2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t jc = blockIdx.y;
    if (i >= n || jc >= ncols) return;
    const int64_t j = j0 + jc, k = j / B, t = i / R;
    const int r = (int)(i - t * R);
    const uint8_t* tile = geno + (t * nblk + k) * tile_bytes_of(B, R, store2);
    out[jc * n + i] = store2 ? (int8_t)((tile[word_off(B, (int)(j - k * B), r >> 2)] >> (2 * (r & 3))) & 3)
                             : (int8_t)tile[byte_off(B, (int)(j - k * B), r)];
}

// weighted residuals (E.str == "D"): mpm_j = sum_i w_i x_ij^2 with x = g - mean (mme.jl:299-301) and sum_i w_i x_ij; one CTA per marker
__global__ void __launch_bounds__(256) weighted_moments_kernel(const uint8_t* __restrict__ geno, int64_t n, int R, int B, int64_t nblk, int64_t p,
                                                                const double* __restrict__ mean, const double* __restrict__ w,
                                                                double* __restrict__ dw, double* __restrict__ wcs)
{
    const int64_t j = blockIdx.x;
    if (j >= p) return;
    const int64_t k = j / B;
    const int q = (int)(j - k * B);
    const double m = mean[j];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t t = i / R;
        const double x = (double)geno[(t * nblk + k) * ((int64_t)B * R) + byte_off(B, q, (int)(i - t * R))] - m;
        const double wx = w[i] * x;
        s1 += wx; s2 = fma(wx, x, s2);
    }
    __shared__ double sh[2 * 8];
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { sh[2 * (threadIdx.x >> 5)] = s1; sh[2 * (threadIdx.x >> 5) + 1] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int v = 0; v < 8; ++v) { a += sh[2 * v]; b += sh[2 * v + 1]; }
        wcs[j] = a; dw[j] = b;
    }
}

__global__ void region_of_kernel(const int64_t* region_off, int64_t n_regions, int32_t* region_of)
{
    for (int64_t r = blockIdx.x; r < n_regions; r += gridDim.x)
        for (int64_t j = region_off[r] + threadIdx.x; j < region_off[r + 1]; j += blockDim.x) region_of[j] = (int32_t)r;
}

// effects of a tuple: member arrays beta_b[j]  <->  interleaved copy beta[j*k + b]
struct TupleBeta { double* member[8]; };
__global__ void tuple_beta_kernel(TupleBeta M, double* inter, int k, int64_t p, int to_inter)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    for (int b = 0; b < k; ++b) {
        if (to_inter) inter[j * k + b] = M.member[b][j];
        else M.member[b][j] = inter[j * k + b];
    }
}

__global__ void fill_kernel(double* x, int64_t n, double v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void fill_i32_kernel(int32_t* x, int64_t n, int32_t v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

__global__ void debug_variates_kernel(uint32_t key0, uint32_t key1, uint32_t chain, uint32_t iter, uint32_t set_id,
                                      int purpose, double df, int64_t n, double* out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Stream st{key0, key1, chain, iter, set_id};
    double v;
    if (purpose == P_U) v = stream_uniform(st, P_U, (uint32_t)i);
    else if (purpose == P_Z) v = stream_normal(st, P_Z, (uint32_t)i);
    else if (purpose == P_PI_A) v = stream_beta(st, df, (double)(i + 1));
    else v = stream_chisq(st, P_CHI2_B, (uint32_t)i, 0, df);
    out[i] = v;
}

}  // namespace

// ============================================================================= handle
struct SetHost {
    bool have_geno = false, have_prior = false, joint_member = false;
    bool gram_global = false;          // row-sharded chain: the banded Gram has been summed over the ranks (ngp_set_gram)
    int64_t p = 0, p_pad = 0, nvar = 0, n_regions = 0;
    int method = 0, est_pi = 0, storage = 0;
    double df = 4.0, scale = 0.0, var_init = 0.0, pi_in = 0.0;
    uint8_t* geno = nullptr;
    double* consts = nullptr;
    int32_t *gx = nullptr, *colsum = nullptr, *colsumsq = nullptr, *delta = nullptr, *region_of = nullptr;
    double *d = nullptr, *mean = nullptr, *beta = nullptr, *varBeta = nullptr, *pi = nullptr, *pi_class = nullptr, *jinvB = nullptr;
    int group_k = 0, stream_set = 0;
    bool joint_copy = false;           // the interleaved copy of a tuple's member sets
    int n_class = 0, n_annot = 0;
    int32_t *annot = nullptr, *annot_cat = nullptr;       // BayesRCpi / BayesRCplus
    double *annot_prob = nullptr, *rp_u_annot = nullptr, *rp_dirp = nullptr;
    double v_class[kMaxClass] = {0.0};
    double *lhs0 = nullptr, *rhs0 = nullptr, *sum_beta = nullptr, *sum_beta2 = nullptr, *sum_delta = nullptr;
    int64_t* region_off = nullptr;
    double *rp_u = nullptr, *rp_z = nullptr, *rp_chi2b = nullptr, *rp_betapi = nullptr;
    double *dw = nullptr, *wcs = nullptr;      // weighted residuals: weighted mpm, weighted centred column sums
    bool w_ready = false;
};

struct JointHost {      // tuple of marker sets with jointly drawn effects (mme.jl:448-489)
    bool active = false;
    int k = 0, set[kMaxK] = {0};
    int64_t p = 0, n_regions = 0;
    double df = 0.0, scale[kMaxK * kMaxK] = {0.0};
    double *varBeta = nullptr, *mtm = nullptr;
    int64_t* region_off = nullptr;
    double *rp_z = nullptr, *rp_iw_chi2 = nullptr, *rp_iw_z = nullptr;
    int replay_iters = 0;
    int blocked_set = -1;     // slot of the interleaved copy swept by the blocked kernel (k = 2, 4, 8), or -1
    bool state_in_copy = false;   // the current effects live in the interleaved copy (else in the member sets)
};

struct ngp_handle {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    // geometry
    int64_t n = 0;
    int refetch = 0, store2 = -1;      // store2: device storage of ALL marker sets of the handle (-1 until the first upload)
    int Tw = 0, R = 0, B = 0, D = 0, DN = 0, NT = 0, NR = 0, NV = 0;      // worker CTAs (grid = Tw + 1), rows per panel, block, look-ahead, near depth, tile stages
    int cfg_kernel = NGP_KERNEL_BLOCKED, cfg_block = 0, cfg_min_rows = 128, cfg_max_ctas = 0, cfg_lookahead = 0, cfg_tile_stages = 0, cfg_near = 0, cfg_profile = 0, cfg_debug = 0, cfg_versions = 0, cfg_refetch = -1, cfg_opt = -1;     // cfg_opt -1: by tile-ring mode (resident: 3, refetch: 2 — refetched columns should still be in L2)
    SmemLayout L{};
    // model
    SetHost sets[NGP_MAX_SETS];
    JointHost joint;
    int n_sets = 0;
    double* e = nullptr;
    double* w = nullptr;           // residual weights E.iVarStr (mme.jl:73), [Tw*R], or null
    double w_sum = 0.0, w_min = 0.0, w_max = 0.0;
    bool have_y = false;
    double df_e = 4.0, scale_e = 0.0;
    int has_mu = 0;
    double mu_lhs0 = 0.0, mu_rhs0 = 0.0;
    SetDev* sets_dev = nullptr;
    bool sets_dirty = true;
    Scalars* sc = nullptr;
    SyncArea* sync = nullptr;
    // variates
    uint64_t seed = 0;
    uint32_t chain = 0;
    int replay = 0, replay_iters = 0;
    int64_t replay_base = 0;
    double *rp_chi2_e = nullptr, *rp_z_mu = nullptr;
    // fixed effects besides the intercept
    FxDev fx{};
    double *fx_data = nullptr, *fx_xpx = nullptr, *fx_colsum = nullptr, *fx_b = nullptr, *fx_rp_z = nullptr;
    double *fx_xpx_w = nullptr, *fx_colsum_w = nullptr;      // weighted residuals: X'WX and w'x_c (mme.jl:135), built at the first launch
    std::vector<double> fx_host, w_host;                      // host copies ([n_cols][n] columns, weights) to build them from
    bool fx_w_ready = false;
    int fx_replay_iters = 0;
    // row-sharded chain
    int shard_rank = 0, shard_world = 1;
    bool shard_attached = false, shard_same_device = false;   // same_device: another rank of the chain lives on this device (one grid: ngp_run_group)
    SyncArea* peer[kMaxRanks] = {nullptr};
    bool peer_ipc[kMaxRanks] = {false};
    int shard_Tw[kMaxRanks] = {0};
    int64_t n_total = 0;
    uint64_t bar_count = 0;     // grid-barrier rounds so far (the sharded counter is monotonic across launches)
    // pinned staging of the sweep-level call (delta as int32, pi, error flag): one stream synchronisation per ngp_sweep
    unsigned char* stage = nullptr;
    size_t stage_bytes = 0;
    const void* ready_kfn = nullptr;     // sweep-kernel variant whose launch attributes are set
    // stats
    int64_t launches = 0;
    int* err_pinned = nullptr;  // pinned landing place of the kernel's error word: copied on the stream right behind the launch, ONE synchronisation per run
    double sum_run_ms = 0.0;    // device time of all sweep launches so far
    int last_variant = -1;      // kernel variant of the last launch (ngp_timing.kernel_variant)
    uint64_t gblk = 0;          // blocks swept so far by the blocked kernel (numbers the list words and accumulator slots)
    bool timed = false;
    bool dense_rings = false;   // the rings are sized for sets whose every effect changes in every sweep (apply_ring_geometry)
};

static std::string g_create_err;

static int fail(ngp_handle* h, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(h, NGP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// every copy / memset is ordered on the handle's stream (own stream is non-blocking, so the
// legacy default stream would NOT be ordered against the sweep kernel)
static cudaError_t cpy(ngp_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, h->stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(h->stream);
}
static cudaError_t zero(ngp_handle* h, void* dst, int v, size_t bytes) { return cudaMemsetAsync(dst, v, bytes, h->stream); }

template <class Tp>
static cudaError_t dalloc(Tp** p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(Tp));
}

static void free_joint(JointHost& j)
{
    cudaFree(j.varBeta); cudaFree(j.mtm); cudaFree(j.region_off); cudaFree(j.rp_z); cudaFree(j.rp_iw_chi2); cudaFree(j.rp_iw_z);
    j = JointHost();
}

static void free_set(SetHost& s)
{
    cudaFree(s.geno); cudaFree(s.consts); cudaFree(s.gx); cudaFree(s.colsum); cudaFree(s.colsumsq); cudaFree(s.delta); cudaFree(s.region_of);
    cudaFree(s.d); cudaFree(s.mean); cudaFree(s.beta); cudaFree(s.varBeta); cudaFree(s.pi); cudaFree(s.pi_class); cudaFree(s.jinvB);
    cudaFree(s.lhs0); cudaFree(s.rhs0); cudaFree(s.sum_beta); cudaFree(s.sum_beta2); cudaFree(s.sum_delta);
    cudaFree(s.region_off); cudaFree(s.rp_u); cudaFree(s.rp_z); cudaFree(s.rp_chi2b); cudaFree(s.rp_betapi);
    cudaFree(s.dw); cudaFree(s.wcs);
    cudaFree(s.annot); cudaFree(s.annot_cat); cudaFree(s.annot_prob); cudaFree(s.rp_u_annot); cudaFree(s.rp_dirp);
    s = SetHost();
}

extern "C" {

int ngp_abi_version(void) { return NGP_ABI_VERSION; }

int ngp_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

const char* ngp_last_error(const ngp_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int ngp_create(int device, ngp_handle** out)
{
    ngp_handle* h = nullptr;
    if (!out) return fail(h, NGP_EINVAL, "ngp_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(h, NGP_ECUDA, "ngp_create: no CUDA device (%s); libngp has no CPU fallback",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    }
    if (device < 0 || device >= count) return fail(h, NGP_EINVAL, "ngp_create: device %d out of range [0,%d)", device, count);
    ngp_handle* nh = new ngp_handle();
    nh->device = device;
    h = nullptr;
    CU(cudaSetDevice(device));
    CU(cudaGetDeviceProperties(&nh->prop, device));
    if (nh->prop.major < 10) {
        int maj = nh->prop.major, mnr = nh->prop.minor;
        delete nh;
        return fail(h, NGP_EUNSUPPORTED, "ngp_create: device is sm_%d%d; libngp is built for sm_100a only", maj, mnr);
    }
    h = nh;
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaMalloc((void**)&h->sc, sizeof(Scalars)));
    CU(zero(h, h->sc, 0, sizeof(Scalars)));
    CU(cudaMalloc((void**)&h->sync, sizeof(SyncArea)));
    CU(zero(h, h->sync, 0, sizeof(SyncArea)));
    CU(cudaMalloc((void**)&h->sets_dev, sizeof(SetDev) * NGP_MAX_SETS));
    CU(cudaMallocHost((void**)&h->err_pinned, sizeof(int)));
    *out = h;
    return NGP_OK;
}

int ngp_destroy(ngp_handle* h)
{
    if (!h) return NGP_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& s : h->sets) free_set(s);
    free_joint(h->joint);
    for (int r = 0; r < h->shard_world; ++r) if (h->peer_ipc[r] && h->peer[r]) cudaIpcCloseMemHandle(h->peer[r]);
    cudaFree(h->e); cudaFree(h->w); cudaFree(h->sc); cudaFree(h->sync); cudaFree(h->sets_dev);
    cudaFree(h->rp_chi2_e); cudaFree(h->rp_z_mu);
    if (h->stage) cudaFreeHost(h->stage);
    if (h->err_pinned) cudaFreeHost(h->err_pinned);
    cudaFree(h->fx_data); cudaFree(h->fx_xpx); cudaFree(h->fx_colsum); cudaFree(h->fx_b); cudaFree(h->fx_rp_z);
    cudaFree(h->fx_xpx_w); cudaFree(h->fx_colsum_w);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return NGP_OK;
}

int ngp_configure(ngp_handle* h, int key, int64_t value)
{
    if (!h) return NGP_EINVAL;
    switch (key) {
    case NGP_CFG_KERNEL:
        if (value != NGP_KERNEL_BLOCKED && value != NGP_KERNEL_LITERAL) return fail(h, NGP_EINVAL, "ngp_configure: unknown kernel %lld", (long long)value);
        h->cfg_kernel = (int)value; return NGP_OK;
    case NGP_CFG_BLOCK:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: block size must be set before the first upload");
        if (value != 0 && value != 16 && value != 32 && value != 64) return fail(h, NGP_EINVAL, "ngp_configure: block must be 0, 16, 32 or 64");
        h->cfg_block = (int)value; return NGP_OK;
    case NGP_CFG_MIN_ROWS:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: min rows must be set before the first upload");
        if (value < 8) return fail(h, NGP_EINVAL, "ngp_configure: min rows must be >= 8");
        h->cfg_min_rows = (int)value; return NGP_OK;
    case NGP_CFG_MAX_CTAS:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: max CTAs must be set before the first upload");
        if (value < 0 || value == 1) return fail(h, NGP_EINVAL, "ngp_configure: max CTAs must be 0 (one per SM) or >= 2 (workers + the chain CTA)");
        h->cfg_max_ctas = (int)value; return NGP_OK;
    case NGP_CFG_VERSIONS:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: residual versions must be set before the first upload");
        if (value < 0 || value > kLimbVers || value == 1) return fail(h, NGP_EINVAL, "ngp_configure: residual versions must be 0 (auto) or in [2,%d]", kLimbVers);
        h->cfg_versions = (int)value; return NGP_OK;
    case NGP_CFG_DEBUG:
        // only the decoupling experiments that keep every ring protocol alive are accepted (2: prep warps skip the accumulator poll,
        // 4: the chain warp skips corrections + scalar updates); the others stall the pipeline they were written for
        if (value != 0 && value != 2 && value != 4 && value != 6) return fail(h, NGP_EINVAL, "ngp_configure: debug must be 0, 2, 4 or 6");
        h->cfg_debug = (int)value; return NGP_OK;
    case NGP_CFG_REFETCH:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: the tile-ring mode must be set before the first upload");
        if (value < -1 || value > 1) return fail(h, NGP_EINVAL, "ngp_configure: refetch must be -1 (auto), 0 or 1");
        h->cfg_refetch = (int)value; return NGP_OK;
    case NGP_CFG_PROFILE:
        h->cfg_profile = value ? 1 : 0; return NGP_OK;
    case NGP_CFG_OPT:
        if (value < -1 || value > 0xffff) return fail(h, NGP_EINVAL, "ngp_configure: option mask out of range (-1 = auto)");
        h->cfg_opt = (int)value; return NGP_OK;
    case NGP_CFG_LOOKAHEAD:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: look-ahead must be set before the first upload");
        if (value < 0 || value > kMaxD) return fail(h, NGP_EINVAL, "ngp_configure: look-ahead must be in [0,%d] (0 = auto)", kMaxD);
        h->cfg_lookahead = (int)value; return NGP_OK;
    case NGP_CFG_TILE_STAGES:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: tile stages must be set before the first upload");
        if (value < 0 || value > kMaxD + 8) return fail(h, NGP_EINVAL, "ngp_configure: tile stages must be in [0,%d] (0 = auto)", kMaxD + 8);
        h->cfg_tile_stages = (int)value; return NGP_OK;
    case NGP_CFG_NEAR:
        if (h->Tw) return fail(h, NGP_EINVAL, "ngp_configure: near depth must be set before the first upload");
        if (value < 0 || value > kMaxD) return fail(h, NGP_EINVAL, "ngp_configure: near depth must be in [0,%d] (0 = auto)", kMaxD);
        h->cfg_near = (int)value; return NGP_OK;
    default: return fail(h, NGP_EINVAL, "ngp_configure: unknown key %d", key);
    }
}

int ngp_set_stream(ngp_handle* h, void* cuda_stream)
{
    if (!h) return NGP_EINVAL;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return NGP_OK;
}

// ----------------------------------------------------------------------------- geometry
static int choose_geometry(ngp_handle* h, int64_t n, int store2)
{
    const int sms = h->prop.multiProcessorCount;
    int maxc = h->cfg_max_ctas ? std::min(h->cfg_max_ctas, sms) : sms;
    maxc = std::min(maxc, std::min(160, kMaxCtas / h->shard_world));      // phase-0 partials of all ranks share one array
    if (maxc < 2) return fail(h, NGP_EUNSUPPORTED, "the sweep kernel needs at least 2 co-resident CTAs (device has %d SMs)", sms);
    const int maxw = maxc - 1;                     // one CTA runs the scalar chain
    const int64_t want = (n + h->cfg_min_rows - 1) / h->cfg_min_rows;
    int Tw = (int)std::max<int64_t>(1, std::min<int64_t>(maxw, want));
    const int64_t R = 32 * ((n + 32LL * Tw - 1) / (32LL * Tw));      // rows per panel: whole 32-row MMA chunks
    Tw = (int)((n + R - 1) / R);                                    // no empty panels
    const size_t cap = h->prop.sharedMemPerBlockOptin;
    int B = h->cfg_block;
    // Blocks of 64 markers halve the number of trips round the feedback loop (list -> residual version -> dots -> sums) against blocks of
    // 32, and the chain warp steps over 64 markers anyway.  Panels of up to 512 rows: 2-bit tiles of 64 markers (16 R bytes) stay
    // resident with 12 blocks of look-ahead; int8 tiles (64 R bytes) use the refetch ring.  Larger panels (the BIGR instantiation): blocks of 64 for 2-bit tiles and for int8 panels of up to
    // 1024 rows (refetch ring of 2 stages: C3 21.3 -> 15.3 ms/sweep); beyond, int8, a ring of ONE tile whose dots all 8 dot warps share (C5 3.24 with
    // blocks of 16 -> 2.53 ms/sweep; sweeps in profiles/r2/tune_*.jsonl).
    const int64_t maxR = 4LL * kUpdThreads * kUpdGroups;     // residual rows the updater warps hold in registers
    if (R > maxR)
        return fail(h, NGP_EUNSUPPORTED, "n = %lld needs %lld rows per CTA; at most %lld are supported", (long long)n, (long long)R, (long long)maxR);
    const bool auto_B = (B == 0);
    if (auto_B) B = 64;
  for (;;) {                                                 // (blocks of 64, else — automatic choice only — blocks of 16 where a tile of 64 markers does not fit)
    const int dn_min = 2 * ((B == 16 ? 32 : 64) / B) - 1;    // the chain warp steps over 64 markers (32 for blocks of 16): distances inside two steps come from the records
    // Two tile-ring modes.  resident: a tile stays in shared memory until its block has been applied to e (NT >= D + 2), so the rare
    // residual update reads it there.  refetch: a tile stays only until its dots are formed (NT = 8 / 4 / 2 / 1 stages, consumed by as many
    // dot warps) and the columns of changed effects are re-read from L2 / HBM: the look-ahead is no longer bounded by shared memory.
    // Small panels use resident; refetch is chosen when resident would leave fewer than kMinResidentD blocks of look-ahead.
    // defaults from the sweeps in profiles/r1/tune_*_r1j.jsonl (C2: D 14 / near 4 = 0.971 ms against 0.994 at D 13 / near 3; near 5 falls off a
    // cliff at 1.13 ms; C1 with blocks of 64: near 3 = 0.582 ms against 0.593 at near 1 in the sweep but no gain in bench.py, left at the minimum)
    const int want_D = h->cfg_lookahead ? h->cfg_lookahead : (B == 64 ? (n < 20000 ? 6 : 12) : B == 32 ? 14 : 20);     // (few rows: the loop is short, look-ahead only costs corrections)
    constexpr int kMinResidentD = 10;
    auto search = [&](int refetch) -> bool {
        int DN = std::max(dn_min, h->cfg_near ? h->cfg_near : (B == 32 ? 4 : 0));
        int D = std::min(refetch ? std::min(want_D, 13) : want_D, kNzRing - 2);
        if (refetch && h->cfg_lookahead) D = std::min(h->cfg_lookahead, kNzRing - 2);
        auto legal_nt = [](int v) { return v >= 8 ? 8 : v >= 4 ? 4 : v >= 2 ? 2 : 1; };
        for (;; ) {
            if (D < dn_min) break;
            DN = std::max(dn_min, std::min(DN, D));
            const int nt_min = refetch ? 1 : D + 2;
            int NT = refetch ? legal_nt(h->cfg_tile_stages ? h->cfg_tile_stages : (B == 64 ? ((R <= 512 || store2) ? 4 : 2) : 8)) : (h->cfg_tile_stages ? std::max(h->cfg_tile_stages, nt_min) : D + 4);
            // shrink the record ring of the chain CTA, the tile ring, then the look-ahead, until both CTA roles fit
            for (;;) {
                for (int NR = kRecStages; NR >= 2; NR >>= 1) {
                    const int NV = h->cfg_versions ? h->cfg_versions : std::max(4, std::min(kLimbVers, D / 2 + 3));
                    SmemLayout L = smem_layout((int)R, B, NT, DN, NR, NV, store2);
                    // static shared memory of the kernels: 1.5 KB, + 4.1 KB in the instantiation whose dot warps share the one tile of the ring
                    const size_t stat = 2048 + ((refetch && NT == 1 && R > 4 * kUpdThreads && B != 16) ? (size_t)kDotWarps * B * 8 + 64 : 0);
                    if ((size_t)L.total + stat <= cap) {
                        h->n = n; h->Tw = Tw; h->R = (int)R; h->B = B; h->D = D; h->DN = DN; h->NT = NT; h->NR = NR; h->NV = NV; h->L = L;
                        h->refetch = refetch;
                        return true;
                    }
                }
                if (NT > nt_min) NT = refetch ? legal_nt(NT - 1) : NT - 1; else break;
            }
            if (DN > dn_min) { --DN; continue; }
            if (D > dn_min) { --D; continue; }
            break;
        }
        return false;
    };
    if (h->cfg_refetch == 0) { if (search(0)) return NGP_OK; }
    else if (h->cfg_refetch == 1) { if (search(1)) return NGP_OK; }
    else {
        if (search(0) && h->D >= std::min(want_D, kMinResidentD)) return NGP_OK;
        if (search(1)) return NGP_OK;
        if (search(0)) return NGP_OK;
    }
    if (auto_B && B == 64) { B = 16; continue; }
    break;
  }
    return fail(h, NGP_EUNSUPPORTED, "n = %lld needs %lld rows per CTA: the panel tiles (block %d) do not fit in %zu bytes of shared memory",
                (long long)n, (long long)R, B, cap);
}

static int build_gram(ngp_handle* h, SetHost& S)
{
    const int64_t nblk = S.p_pad / h->B;
    cudaFree(S.gx); S.gx = nullptr;
    CU(dalloc(&S.gx, (size_t)nblk * (h->D + 1) * h->B * h->B));
    if (h->B == 64) gram_kernel<64><<<(unsigned)nblk, 256, 0, h->stream>>>(S.geno, h->Tw, h->R, nblk, h->D, S.gx, h->store2);
    else if (h->B == 32) gram_kernel<32><<<(unsigned)nblk, 256, 0, h->stream>>>(S.geno, h->Tw, h->R, nblk, h->D, S.gx, h->store2);
    else gram_kernel<16><<<(unsigned)nblk, 256, 0, h->stream>>>(S.geno, h->Tw, h->R, nblk, h->D, S.gx, h->store2);
    CU(cudaGetLastError());
    return NGP_OK;
}

static int choose_geometry(ngp_handle* h, int64_t n, int store2);

// Dense-update sets.  Every effect of a BayesPR set (BayesRR, region-wise or per-locus variances) changes in every sweep, so the look-ahead
// buys nothing but cross-Gram corrections (B changed columns per block and distance): when all marker sets of the handle are BayesPR the
// rings are re-sized to a short look-ahead served from the block records (D = DN = 3, the tuple sweep's geometry: c4rr 15.1 -> 10.8 ms,
// C1 0.60 -> 0.55 ms per sweep, profiles/r2/tune_c*_dense.jsonl), and back when a spike-and-slab set joins.  Tw, R, B and the tile-ring
// mode stay; the banded Gram is laid out by look-ahead and is rebuilt.  Automatic geometry on one GPU only.  (Panels of more than 512 rows:
// dense sets are better served by blocks of 16 with resident tiles — 100k x 600k BayesRR 135 ms per sweep against 206 with blocks of 64 on
// the refetch ring, where every column is read twice — but the block size is fixed at the first upload: the host mirrors that know the priors
// before they upload, api.getMME and bench.py, pass NGP_CFG_BLOCK = 16 for such models; see INTEGRATION.md.)
static int apply_ring_geometry(ngp_handle* h)
{
    if (h->Tw == 0 || h->cfg_lookahead || h->cfg_near || h->shard_world > 1 || h->joint.active) return NGP_OK;
    bool any = false, all_pr = true;
    for (int s = 0; s < h->n_sets; ++s) {
        const SetHost& S = h->sets[s];
        if (!S.have_geno || !S.have_prior) continue;
        any = true;
        if (S.method != NGP_BAYESPR) all_pr = false;
    }
    const bool want = any && all_pr && h->B == 64;      // (blocks of 16 / 32 want a longer look-ahead even for dense sets: c4rr 16:4 18.4 ms, 16:8 14.0; 32:3 15.1, 32:6 11.4)
    if (want == h->dense_rings) return NGP_OK;
    const int sv_l = h->cfg_lookahead, sv_n = h->cfg_near, sv_r = h->cfg_refetch, sv_b = h->cfg_block;
    const int Tw0 = h->Tw, R0 = h->R, D0 = h->D;
    h->cfg_block = h->B; h->cfg_refetch = h->refetch;
    if (want) { h->cfg_lookahead = 3; h->cfg_near = 3; }
    const int rc = choose_geometry(h, h->n, h->store2 > 0);
    h->cfg_lookahead = sv_l; h->cfg_near = sv_n; h->cfg_refetch = sv_r; h->cfg_block = sv_b;
    if (rc) return rc;
    if (h->Tw != Tw0 || h->R != R0) return fail(h, NGP_EINVAL, "internal: the tile layout changed while re-sizing the rings");
    h->dense_rings = want;
    h->ready_kfn = nullptr;
    h->sets_dirty = true;
    CU(cudaMemsetAsync(h->sync, 0, sizeof(SyncArea), h->stream));      // clean rings for the new geometry
    h->gblk = 0;
    if (h->D != D0)
        for (int s = 0; s < h->n_sets; ++s)
            if (h->sets[s].have_geno) { int rg = build_gram(h, h->sets[s]); if (rg) return rg; }
    CU(cudaStreamSynchronize(h->stream));
    return NGP_OK;
}

static int finish_upload(ngp_handle* h, SetHost& S)
{
    const int64_t p_pad = S.p_pad;
    CU(dalloc(&S.mean, p_pad));
    CU(dalloc(&S.d, p_pad));
    colstats_kernel<<<(unsigned)((p_pad + 255) / 256), 256, 0, h->stream>>>(h->n, S.p, p_pad, S.colsum, S.colsumsq, S.mean, S.d);
    CU(cudaGetLastError());
    const int64_t nblk = p_pad / h->B;
    CU(dalloc(&S.consts, (size_t)nblk * kNF * h->B));
    CU(zero(h, S.consts, 0, sizeof(double) * (size_t)nblk * kNF * h->B));
    { int rg = build_gram(h, S); if (rg) return rg; }
    CU(dalloc(&S.beta, p_pad));
    CU(dalloc(&S.delta, p_pad));
    CU(cudaMemsetAsync(S.beta, 0, sizeof(double) * p_pad, h->stream));
    fill_i32_kernel<<<(unsigned)((p_pad + 255) / 256), 256, 0, h->stream>>>(S.delta, p_pad, 1);   // delta = ones (mme.jl:444)
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    S.have_geno = true;
    S.have_prior = false;
    h->sets_dirty = true;
    return NGP_OK;
}

static int begin_upload(ngp_handle* h, int set_id, int64_t n, int64_t p, int storage)
{
    if (!h) return NGP_EINVAL;
    if (set_id < 0 || set_id >= NGP_MAX_SETS) return fail(h, NGP_EINVAL, "set_id %d out of range", set_id);
    if (n <= 0 || p <= 0) return fail(h, NGP_EINVAL, "n and p must be positive");
    if (p > 0x7fffffffLL || n > 0x7fffffffLL) return fail(h, NGP_EINVAL, "n and p must fit in 31 bits");
    if (storage != NGP_STORE_I8 && storage != NGP_STORE_2BIT) return fail(h, NGP_EINVAL, "unknown device storage %d", storage);
    if (h->store2 >= 0 && h->store2 != (storage == NGP_STORE_2BIT)) return fail(h, NGP_EINVAL, "all marker sets of a handle share one device storage format");
    if (storage == NGP_STORE_2BIT && h->shard_world > 1) return fail(h, NGP_EUNSUPPORTED, "2-bit device storage: blocked sweep on one GPU only (the row-sharded chain runs the per-marker kernel)");
    CU(cudaSetDevice(h->device));
    if (h->Tw == 0) {
        int rc = choose_geometry(h, n, storage == NGP_STORE_2BIT);
        if (rc) return rc;
        h->store2 = (storage == NGP_STORE_2BIT);
        CU(dalloc(&h->e, (size_t)h->Tw * h->R));
        CU(zero(h, h->e, 0, sizeof(double) * (size_t)h->Tw * h->R));
    } else if (n != h->n) {
        return fail(h, NGP_EINVAL, "all marker sets of a handle must have n = %lld individuals (got %lld)", (long long)h->n, (long long)n);
    }
    SetHost& S = h->sets[set_id];
    if (S.joint_member) {                       // re-uploading a member dissolves the tuple
        for (int b = 0; b < h->joint.k; ++b) { h->sets[h->joint.set[b]].joint_member = false; h->sets[h->joint.set[b]].have_prior = false; }
        free_joint(h->joint);
    }
    free_set(S);
    S.p = p;
    S.p_pad = ((p + kMaxB - 1) / kMaxB) * kMaxB;
    S.storage = storage;
    const size_t gbytes = ((size_t)h->Tw * S.p_pad * h->R) >> (storage == NGP_STORE_2BIT ? 2 : 0);
    CU(dalloc(&S.geno, gbytes));
    CU(cudaMemsetAsync(S.geno, 0, gbytes, h->stream));             // code 0 everywhere (pad rows / pad markers)
    CU(dalloc(&S.colsum, S.p_pad));
    CU(dalloc(&S.colsumsq, S.p_pad));
    CU(cudaMemsetAsync(S.colsum, 0, sizeof(int32_t) * S.p_pad, h->stream));
    CU(cudaMemsetAsync(S.colsumsq, 0, sizeof(int32_t) * S.p_pad, h->stream));
    h->n_sets = std::max(h->n_sets, set_id + 1);
    return NGP_OK;
}

int ngp_upload_genotypes(ngp_handle* h, int set_id, int64_t n, int64_t p, const void* data, int fmt, int64_t ld, int storage)
{
    if (!h) return NGP_EINVAL;
    if (!data) return fail(h, NGP_EINVAL, "ngp_upload_genotypes: data is NULL");
    if (fmt != NGP_GENO_I8 && fmt != NGP_GENO_F64 && fmt != NGP_GENO_PACKED2) return fail(h, NGP_EINVAL, "unknown genotype format %d", fmt);
    const int64_t min_ld = (fmt == NGP_GENO_PACKED2) ? (n + 3) / 4 : n;
    if (ld < min_ld) return fail(h, NGP_EINVAL, "ld = %lld is smaller than a column (%lld)", (long long)ld, (long long)min_ld);
    int rc = begin_upload(h, set_id, n, p, storage);
    if (rc) return rc;
    SetHost& S = h->sets[set_id];
    const size_t elem = (fmt == NGP_GENO_F64) ? 8 : 1;
    const size_t colbytes = (size_t)ld * elem;
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(p, (int64_t)((256u << 20) / colbytes)));
    chunk = std::min<int64_t>(chunk, 65535);
    void* stage = nullptr;
    int* derr = nullptr;
    CU(cudaMalloc(&stage, colbytes * chunk));
    CU(cudaMalloc((void**)&derr, sizeof(int)));
    CU(cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
    for (int64_t j0 = 0; j0 < p; j0 += chunk) {
        const int64_t nc = std::min(chunk, p - j0);
        CU(cudaMemcpyAsync(stage, (const char*)data + (size_t)j0 * colbytes, colbytes * nc, cudaMemcpyHostToDevice, h->stream));
        if (fmt == NGP_GENO_I8) pack_kernel<NGP_GENO_I8><<<(unsigned)nc, 256, 0, h->stream>>>(stage, ld, n, j0, h->R, h->B, S.p_pad / h->B, S.geno, S.colsum, S.colsumsq, derr, 1, 0, h->store2);
        else if (fmt == NGP_GENO_F64) pack_kernel<NGP_GENO_F64><<<(unsigned)nc, 256, 0, h->stream>>>(stage, ld, n, j0, h->R, h->B, S.p_pad / h->B, S.geno, S.colsum, S.colsumsq, derr, 1, 0, h->store2);
        else pack_kernel<NGP_GENO_PACKED2><<<(unsigned)nc, 256, 0, h->stream>>>(stage, ld, n, j0, h->R, h->B, S.p_pad / h->B, S.geno, S.colsum, S.colsumsq, derr, 1, 0, h->store2);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));     // the pageable source buffer is borrowed per chunk
    }
    int herr = 0;
    CU(cpy(h, &herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(stage); cudaFree(derr);
    if (herr) { free_set(S); return fail(h, NGP_EDATA, "genotype value outside {0,1,2} (missing data must be removed before upload)"); }
    return finish_upload(h, S);
}

int ngp_synth_genotypes(ngp_handle* h, int set_id, int64_t n, int64_t p, uint64_t seed, const uint32_t* thr0, const uint32_t* thr1, int storage)
{
    return ngp_synth_genotypes_rows(h, set_id, 0, n, p, seed, thr0, thr1, storage);
}

int ngp_synth_genotypes_rows(ngp_handle* h, int set_id, int64_t row0, int64_t n, int64_t p, uint64_t seed, const uint32_t* thr0, const uint32_t* thr1, int storage)
{
    if (!h) return NGP_EINVAL;
    if (!thr0 || !thr1) return fail(h, NGP_EINVAL, "ngp_synth_genotypes: thresholds are NULL");
    if (row0 < 0 || (row0 & 3)) return fail(h, NGP_EINVAL, "ngp_synth_genotypes_rows: row0 must be a non-negative multiple of 4");
    int rc = begin_upload(h, set_id, n, p, storage);
    if (rc) return rc;
    SetHost& S = h->sets[set_id];
    uint32_t *d0 = nullptr, *d1 = nullptr;
    int8_t* stage = nullptr;
    int* derr = nullptr;
    CU(cudaMalloc((void**)&d0, sizeof(uint32_t) * p));
    CU(cudaMalloc((void**)&d1, sizeof(uint32_t) * p));
    CU(cpy(h, d0, thr0, sizeof(uint32_t) * p, cudaMemcpyHostToDevice));
    CU(cpy(h, d1, thr1, sizeof(uint32_t) * p, cudaMemcpyHostToDevice));
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(p, (int64_t)((512u << 20) / (size_t)n)));
    chunk = std::min<int64_t>(chunk, 65535);
    CU(cudaMalloc((void**)&stage, (size_t)n * chunk));
    CU(cudaMalloc((void**)&derr, sizeof(int)));
    CU(cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
    for (int64_t j0 = 0; j0 < p; j0 += chunk) {
        const int64_t nc = std::min(chunk, p - j0);
        dim3 grid((unsigned)(((n + 3) / 4 + 255) / 256), (unsigned)nc);
        synth_kernel<<<grid, 256, 0, h->stream>>>((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), n, row0, j0, nc, d0, d1, stage);
        CU(cudaGetLastError());
        pack_kernel<NGP_GENO_I8><<<(unsigned)nc, 256, 0, h->stream>>>(stage, n, n, j0, h->R, h->B, S.p_pad / h->B, S.geno, S.colsum, S.colsumsq, derr, 1, 0, h->store2);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(stage); cudaFree(derr); cudaFree(d0); cudaFree(d1);
    return finish_upload(h, S);
}

int ngp_download_genotypes(ngp_handle* h, int set_id, int64_t j0, int64_t j1, int8_t* out)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_download_genotypes: no such set");
    SetHost& S = h->sets[set_id];
    if (j0 < 0 || j1 > S.p || j0 >= j1 || !out) return fail(h, NGP_EINVAL, "ngp_download_genotypes: bad column range");
    CU(cudaSetDevice(h->device));
    const int64_t n = h->n;
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(j1 - j0, (int64_t)((256u << 20) / (size_t)n)));
    chunk = std::min<int64_t>(chunk, 65535);
    int8_t* stage = nullptr;
    CU(cudaMalloc((void**)&stage, (size_t)n * chunk));
    for (int64_t c0 = j0; c0 < j1; c0 += chunk) {
        const int64_t nc = std::min(chunk, j1 - c0);
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)nc);
        unpack_kernel<<<grid, 256, 0, h->stream>>>(S.geno, n, h->R, h->B, S.p_pad / h->B, c0, nc, stage, h->store2);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + (size_t)(c0 - j0) * n, stage, (size_t)n * nc, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    cudaFree(stage);
    return NGP_OK;
}

static int ensure_weighted_moments(ngp_handle* h, int s);
int ngp_get_column_stats(ngp_handle* h, int set_id, double* mean, double* mpm)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_get_column_stats: no such set");
    SetHost& S = h->sets[set_id];
    CU(cudaSetDevice(h->device));
    if (mean) CU(cpy(h, mean, S.mean, sizeof(double) * S.p, cudaMemcpyDeviceToHost));
    if (h->w) { int rcw = ensure_weighted_moments(h, set_id); if (rcw) return rcw; }      // weighted mpm as soon as weights are set, whatever the call order
    if (mpm) CU(cpy(h, mpm, (h->w && S.w_ready) ? S.dw : S.d, sizeof(double) * S.p, cudaMemcpyDeviceToHost));
    return NGP_OK;
}

// ----------------------------------------------------------------------------- host 2-bit codec
int ngp_pack2(const int8_t* codes, int64_t n, int64_t p, int64_t ld_in, uint8_t* out, int64_t ld_out)
{
    if (!codes || !out || n < 0 || p < 0 || ld_in < n || ld_out < (n + 3) / 4) return NGP_EINVAL;
    for (int64_t j = 0; j < p; ++j) {
        const int8_t* c = codes + j * ld_in;
        uint8_t* o = out + j * ld_out;
        memset(o, 0, (size_t)ld_out);
        for (int64_t i = 0; i < n; ++i) {
            const int g = c[i];
            if (g < 0 || g > 2) return NGP_EDATA;
            o[i >> 2] |= (uint8_t)(g << (2 * (i & 3)));
        }
    }
    return NGP_OK;
}

int ngp_unpack2(const uint8_t* packed, int64_t n, int64_t p, int64_t ld_in, int8_t* out, int64_t ld_out)
{
    if (!packed || !out || n < 0 || p < 0 || ld_in < (n + 3) / 4 || ld_out < n) return NGP_EINVAL;
    for (int64_t j = 0; j < p; ++j) {
        const uint8_t* c = packed + j * ld_in;
        int8_t* o = out + j * ld_out;
        for (int64_t i = 0; i < n; ++i) {
            const int g = (c[i >> 2] >> (2 * (i & 3))) & 3;
            if (g > 2) return NGP_EDATA;
            o[i] = (int8_t)g;
        }
    }
    return NGP_OK;
}

// ----------------------------------------------------------------------------- model
int ngp_set_phenotype(ngp_handle* h, const double* y, int64_t n)
{
    if (!h || !y) return fail(h, NGP_EINVAL, "ngp_set_phenotype: NULL argument");
    if (h->Tw == 0) return fail(h, NGP_EINVAL, "ngp_set_phenotype: upload a marker set first (it fixes n)");
    if (n != h->n) return fail(h, NGP_EINVAL, "ngp_set_phenotype: n = %lld but the genotypes have %lld rows", (long long)n, (long long)h->n);
    CU(cudaSetDevice(h->device));
    CU(zero(h, h->e, 0, sizeof(double) * (size_t)h->Tw * h->R));
    CU(cpy(h, h->e, y, sizeof(double) * n, cudaMemcpyHostToDevice));     // ycorr = deepcopy(Y), mme.jl:57
    Scalars z{};
    CU(cpy(h, h->sc, &z, sizeof z, cudaMemcpyHostToDevice));
    if (h->fx_b) CU(zero(h, h->fx_b, 0, sizeof(double) * kMaxFxCols));       // b = zeros, like every effect getMME! allocates
    h->have_y = true;
    return NGP_OK;
}

int ngp_set_residual_prior(ngp_handle* h, double df_e, double scale_e)
{
    if (!h) return NGP_EINVAL;
    if (!(df_e > 0.0) || !(scale_e >= 0.0)) return fail(h, NGP_EINVAL, "ngp_set_residual_prior: df must be > 0 and scale >= 0");
    h->df_e = df_e; h->scale_e = scale_e;
    return NGP_OK;
}

// E.str == "D" (mme.jl:70-73): w = E.iVarStr = inv.(priorVCV[:e].str), one positive weight per individual; NULL returns to "I".
// Replaces sum(e.^2) by sum(w.*e.^2) in sampleVarE (functions.jl:526-528), xpx / Xp of the intercept (mme.jl:135-136) and
// mpm / Mp of every marker set (mme.jl:299-303).  Sampled by the per-marker kernel (the weighted dots are not integer sums of codes).
int ngp_set_residual_weights(ngp_handle* h, const double* w, int64_t n)
{
    if (!h) return NGP_EINVAL;
    CU(cudaSetDevice(h->device));
    for (auto& S : h->sets) S.w_ready = false;
    h->sets_dirty = true;
    h->fx_w_ready = false;
    if (!w) { cudaFree(h->w); h->w = nullptr; h->w_host.clear(); return NGP_OK; }
    if (h->Tw == 0) return fail(h, NGP_EINVAL, "ngp_set_residual_weights: upload a marker set first (it fixes n)");
    if (n != h->n) return fail(h, NGP_EINVAL, "ngp_set_residual_weights: n = %lld but the genotypes have %lld rows", (long long)n, (long long)h->n);
    double sum = 0.0, lo = INFINITY, hi = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (!(w[i] > 0.0) || !std::isfinite(w[i])) return fail(h, NGP_EINVAL, "ngp_set_residual_weights: weight %lld is not a positive finite number", (long long)i);
        sum += w[i]; lo = std::min(lo, w[i]); hi = std::max(hi, w[i]);
    }
    if (!h->w) CU(dalloc(&h->w, (size_t)h->Tw * h->R));
    CU(zero(h, h->w, 0, sizeof(double) * (size_t)h->Tw * h->R));
    CU(cpy(h, h->w, w, sizeof(double) * n, cudaMemcpyHostToDevice));
    h->w_sum = sum; h->w_min = lo; h->w_max = hi;
    h->w_host.assign(w, w + n);
    return NGP_OK;
}

int ngp_set_intercept(ngp_handle* h, int enabled, double lhs0, double rhs0)
{
    if (!h) return NGP_EINVAL;
    h->has_mu = enabled ? 1 : 0; h->mu_lhs0 = lhs0; h->mu_rhs0 = rhs0;
    return NGP_OK;
}

int ngp_set_fixed_effects(ngp_handle* h, int n_sets, const ngp_fixed_set* sets)
{
    if (!h) return NGP_EINVAL;
    if (h->Tw == 0) return fail(h, NGP_EINVAL, "ngp_set_fixed_effects: upload a marker set first (it fixes n and the row panels)");
    if (h->joint.active) return fail(h, NGP_EUNSUPPORTED, "ngp_set_fixed_effects: not available together with a tuple of marker sets");
    if (n_sets < 0 || n_sets > kMaxFxSets || (n_sets > 0 && !sets)) return fail(h, NGP_EINVAL, "ngp_set_fixed_effects: 0..%d sets", kMaxFxSets);
    CU(cudaSetDevice(h->device));
    cudaFree(h->fx_data); cudaFree(h->fx_xpx); cudaFree(h->fx_colsum); cudaFree(h->fx_b); cudaFree(h->fx_rp_z);
    cudaFree(h->fx_xpx_w); cudaFree(h->fx_colsum_w);
    h->fx_data = h->fx_xpx = h->fx_colsum = h->fx_b = h->fx_rp_z = h->fx_xpx_w = h->fx_colsum_w = nullptr;
    h->fx_host.clear(); h->fx_w_ready = false;
    h->fx = FxDev{};
    h->fx_replay_iters = 0;
    if (n_sets == 0) return NGP_OK;
    FxDev F{};
    int cols = 0, xsz = 0;
    for (int s = 0; s < n_sets; ++s) {
        if (sets[s].n_cols < 1 || !sets[s].data) return fail(h, NGP_EINVAL, "ngp_set_fixed_effects: set %d has no columns", s);
        F.first[s] = cols; F.xoff[s] = xsz; F.lhs0[s] = sets[s].n_cols == 1 ? sets[s].lhs0 : 0.0; F.rhs0[s] = sets[s].n_cols == 1 ? sets[s].rhs0 : 0.0;
        cols += sets[s].n_cols; xsz += sets[s].n_cols * sets[s].n_cols;
    }
    if (cols > kMaxFxCols) return fail(h, NGP_EUNSUPPORTED, "ngp_set_fixed_effects: %d columns exceed %d", cols, kMaxFxCols);
    for (int s = n_sets; s <= kMaxFxSets; ++s) F.first[s] = cols;
    F.n_sets = n_sets; F.n_cols = cols;
    const int64_t n = h->n, ldx = (int64_t)h->Tw * h->R;
    std::vector<double> data((size_t)cols * ldx, 0.0), xpx((size_t)xsz, 0.0), cs((size_t)cols, 0.0);
    for (int s = 0; s < n_sets; ++s) {
        const int nc = sets[s].n_cols;
        for (int c = 0; c < nc; ++c) {
            const double* src = sets[s].data + (int64_t)c * n;
            double* dst = data.data() + (size_t)(F.first[s] + c) * ldx;
            double sum = 0.0;
            for (int64_t i = 0; i < n; ++i) { dst[i] = src[i]; sum += src[i]; }
            cs[(size_t)(F.first[s] + c)] = sum;
            h->fx_host.insert(h->fx_host.end(), src, src + n);
        }
        for (int a = 0; a < nc; ++a)                                                  // X[xSet].xpx = X'X
            for (int b = a; b < nc; ++b) {
                const double *xa = sets[s].data + (int64_t)a * n, *xb = sets[s].data + (int64_t)b * n;
                double d = 0.0;
                for (int64_t i = 0; i < n; ++i) d += xa[i] * xb[i];
                xpx[(size_t)F.xoff[s] + a * nc + b] = d; xpx[(size_t)F.xoff[s] + b * nc + a] = d;
            }
    }
    CU(dalloc(&h->fx_data, data.size())); CU(cpy(h, h->fx_data, data.data(), sizeof(double) * data.size(), cudaMemcpyHostToDevice));
    CU(dalloc(&h->fx_xpx, xpx.size())); CU(cpy(h, h->fx_xpx, xpx.data(), sizeof(double) * xpx.size(), cudaMemcpyHostToDevice));
    CU(dalloc(&h->fx_colsum, cs.size())); CU(cpy(h, h->fx_colsum, cs.data(), sizeof(double) * cs.size(), cudaMemcpyHostToDevice));
    CU(dalloc(&h->fx_b, (size_t)kMaxFxCols)); CU(zero(h, h->fx_b, 0, sizeof(double) * kMaxFxCols));
    CU(cudaStreamSynchronize(h->stream));
    F.data = h->fx_data; F.xpx = h->fx_xpx; F.colsum = h->fx_colsum; F.b = h->fx_b;
    h->fx = F;
    return NGP_OK;
}

int ngp_get_fixed_effects(ngp_handle* h, double* b)
{
    if (!h || !b) return fail(h, NGP_EINVAL, "ngp_get_fixed_effects: NULL argument");
    if (!h->fx.n_cols) return fail(h, NGP_EINVAL, "ngp_get_fixed_effects: the handle has no fixed effects besides the intercept");
    CU(cudaSetDevice(h->device));
    CU(cpy(h, b, h->fx_b, sizeof(double) * h->fx.n_cols, cudaMemcpyDeviceToHost));
    return NGP_OK;
}

int ngp_set_fixed_replay(ngp_handle* h, int32_t n_iter, const double* z)
{
    if (!h) return NGP_EINVAL;
    if (!h->fx.n_cols || n_iter <= 0 || !z) return fail(h, NGP_EINVAL, "ngp_set_fixed_replay: no fixed effects, or empty log");
    CU(cudaSetDevice(h->device));
    CU(dalloc(&h->fx_rp_z, (size_t)n_iter * h->fx.n_cols));
    CU(cpy(h, h->fx_rp_z, z, sizeof(double) * (size_t)n_iter * h->fx.n_cols, cudaMemcpyHostToDevice));
    h->fx_replay_iters = n_iter;
    return NGP_OK;
}

int ngp_set_prior(ngp_handle* h, int set_id, const ngp_prior* pr)
{
    if (!h || !pr) return fail(h, NGP_EINVAL, "ngp_set_prior: NULL argument");
    if (set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_set_prior: upload genotypes of set %d first", set_id);
    SetHost& S = h->sets[set_id];
    if (S.joint_member) return fail(h, NGP_EINVAL, "ngp_set_prior: set %d is a member of a tuple (ngp_set_joint_prior)", set_id);
    if (h->joint.active) return fail(h, NGP_EUNSUPPORTED, "ngp_set_prior: a handle with a tuple of marker sets samples only the tuple");
    if (pr->method != NGP_BAYESPR && pr->method != NGP_BAYESB && pr->method != NGP_BAYESC && pr->method != NGP_BAYESR) return fail(h, NGP_EINVAL, "ngp_set_prior: unknown method %d", pr->method);
    if (!(pr->df > 0.0) || !(pr->var_init >= 0.0)) return fail(h, NGP_EINVAL, "ngp_set_prior: df must be > 0 and var_init >= 0");
    if ((pr->method == NGP_BAYESB || pr->method == NGP_BAYESC) && !(pr->pi_in > 0.0 && pr->pi_in < 1.0)) return fail(h, NGP_EINVAL, "ngp_set_prior: pi_in must be inside (0,1)");
    if (pr->method == NGP_BAYESR) {
        if (pr->n_class < 1 || pr->n_class > kMaxClass || !pr->v_class || !pr->pi_class) return fail(h, NGP_EINVAL, "ngp_set_prior: BayesR needs 1..%d classes with v_class and pi_class", kMaxClass);
        for (int v = 0; v < pr->n_class; ++v)
            if (!(pr->v_class[v] >= 0.0) || !(pr->pi_class[v] > 0.0)) return fail(h, NGP_EINVAL, "ngp_set_prior: BayesR class %d: scale must be >= 0 and proportion > 0", v);
    }
    CU(cudaSetDevice(h->device));
    S.method = pr->method; S.est_pi = pr->est_pi ? 1 : 0; S.df = pr->df; S.scale = pr->scale; S.var_init = pr->var_init; S.pi_in = pr->pi_in;
    cudaFree(S.region_off); S.region_off = nullptr;
    cudaFree(S.region_of); S.region_of = nullptr;
    S.n_regions = 0;
    if (pr->method == NGP_BAYESPR) {
        if (pr->region_off && pr->n_regions >= 1) {
            if (pr->region_off[0] != 0 || pr->region_off[pr->n_regions] != S.p) return fail(h, NGP_EINVAL, "ngp_set_prior: region offsets must start at 0 and end at p");
            for (int64_t r = 0; r < pr->n_regions; ++r)
                if (pr->region_off[r + 1] <= pr->region_off[r]) return fail(h, NGP_EINVAL, "ngp_set_prior: empty or unordered region %lld", (long long)r);
            S.n_regions = pr->n_regions;
        } else S.n_regions = 1;
        S.nvar = S.n_regions;
        if (S.n_regions > 1) {
            CU(dalloc(&S.region_off, S.n_regions + 1));
            CU(cpy(h, S.region_off, pr->region_off, sizeof(int64_t) * (S.n_regions + 1), cudaMemcpyHostToDevice));
            CU(dalloc(&S.region_of, S.p_pad));
            CU(zero(h, S.region_of, 0, sizeof(int32_t) * S.p_pad));
            region_of_kernel<<<(unsigned)std::min<int64_t>(S.n_regions, 4096), 128, 0, h->stream>>>(S.region_off, S.n_regions, S.region_of);
            CU(cudaGetLastError());
        }
    } else if (pr->method == NGP_BAYESB) S.nvar = S.p;
    else S.nvar = 1;
    S.n_class = 0;
    if (pr->method == NGP_BAYESR) {
        S.n_class = pr->n_class;
        double pc[2 * kMaxClass];
        for (int v = 0; v < S.n_class; ++v) { S.v_class[v] = pr->v_class[v]; pc[v] = pr->pi_class[v]; pc[S.n_class + v] = log(pr->pi_class[v]); }   // mme.jl:375,383
        CU(dalloc(&S.pi_class, 2 * kMaxClass));
        CU(cudaMemcpyAsync(S.pi_class, pc, sizeof(double) * 2 * S.n_class, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(dalloc(&S.varBeta, S.nvar));
    fill_kernel<<<(unsigned)((S.nvar + 255) / 256), 256, 0, h->stream>>>(S.varBeta, S.nvar, pr->var_init);     // mme.jl:516
    CU(cudaGetLastError());
    CU(dalloc(&S.pi, 4));
    double pi4[4] = {0.0, 1.0, -INFINITY, 0.0};
    if (pr->method == NGP_BAYESB || pr->method == NGP_BAYESC) { pi4[0] = 1.0 - pr->pi_in; pi4[1] = pr->pi_in; pi4[2] = log(1.0 - pr->pi_in); pi4[3] = log(pr->pi_in); }   // mme.jl:351,359
    CU(cudaMemcpyAsync(S.pi, pi4, sizeof pi4, cudaMemcpyHostToDevice, h->stream));
    cudaFree(S.lhs0); S.lhs0 = nullptr; cudaFree(S.rhs0); S.rhs0 = nullptr;
    if (pr->lhs0) { CU(dalloc(&S.lhs0, S.p)); CU(cudaMemcpyAsync(S.lhs0, pr->lhs0, sizeof(double) * S.p, cudaMemcpyHostToDevice, h->stream)); }
    if (pr->rhs0) { CU(dalloc(&S.rhs0, S.p)); CU(cudaMemcpyAsync(S.rhs0, pr->rhs0, sizeof(double) * S.p, cudaMemcpyHostToDevice, h->stream)); }
    CU(dalloc(&S.sum_beta, S.p_pad)); CU(dalloc(&S.sum_beta2, S.p_pad)); CU(dalloc(&S.sum_delta, S.p_pad));
    CU(cudaMemsetAsync(S.sum_beta, 0, sizeof(double) * S.p_pad, h->stream));
    CU(cudaMemsetAsync(S.sum_beta2, 0, sizeof(double) * S.p_pad, h->stream));
    CU(cudaMemsetAsync(S.sum_delta, 0, sizeof(double) * S.p_pad, h->stream));
    CU(cudaMemsetAsync(S.beta, 0, sizeof(double) * S.p_pad, h->stream));                                          // mme.jl:443
    fill_i32_kernel<<<(unsigned)((S.p_pad + 255) / 256), 256, 0, h->stream>>>(S.delta, S.p_pad, 1);               // mme.jl:444
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    S.have_prior = true;
    h->sets_dirty = true;
    return apply_ring_geometry(h);
}

// Replace only the per-marker prior information (lhs0 / rhs0) of a set; the chain state is kept.  GRN.jl:150-164 (sampleΛ2!) changes
// the right-hand-side offset alpha*pMeans[g] for every gene g while the markers stay the same.
int ngp_set_marker_summary(ngp_handle* h, int set_id, const double* lhs0, const double* rhs0)
{
    if (!h) return NGP_EINVAL;
    if (set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_prior)
        return fail(h, NGP_EINVAL, "ngp_set_marker_summary: marker set %d has no prior yet", set_id);
    CU(cudaSetDevice(h->device));
    SetHost& S = h->sets[set_id];
    if (S.method == NGP_BAYESR && (lhs0 || rhs0)) return fail(h, NGP_EUNSUPPORTED, "ngp_set_marker_summary: BayesR takes no summary statistics (functions.jl:232-281)");
    const double* src[2] = {lhs0, rhs0};
    double** dst[2] = {&S.lhs0, &S.rhs0};
    for (int i = 0; i < 2; ++i) {
        if (!src[i]) { if (*dst[i]) { CU(cudaStreamSynchronize(h->stream)); cudaFree(*dst[i]); *dst[i] = nullptr; h->sets_dirty = true; } continue; }
        if (!*dst[i]) { CU(dalloc(dst[i], S.p)); h->sets_dirty = true; }
        CU(cpy(h, *dst[i], src[i], sizeof(double) * S.p, cudaMemcpyHostToDevice));
    }
    return NGP_OK;
}

// Overwrite the effect variances of one set (nvar doubles) and nothing else.  BayesLV (functions.jl:421-486) keeps the variance model of
// the log-variances on the host: after every BayesPR sweep with one region per locus the caller replaces the device's draw by its own.
int ngp_set_var_beta(ngp_handle* h, int set_id, const double* varBeta)
{
    if (!h) return NGP_EINVAL;
    if (set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_prior || !h->sets[set_id].varBeta || !varBeta)
        return fail(h, NGP_EINVAL, "ngp_set_var_beta: marker set %d has no prior yet, or NULL argument", set_id);
    CU(cudaSetDevice(h->device));
    SetHost& S = h->sets[set_id];
    CU(cpy(h, S.varBeta, varBeta, sizeof(double) * S.nvar, cudaMemcpyHostToDevice));
    return NGP_OK;
}

int ngp_set_rng(ngp_handle* h, uint64_t seed, uint32_t chain_id)
{
    if (!h) return NGP_EINVAL;
    h->seed = seed; h->chain = chain_id;
    return NGP_OK;
}

int ngp_set_replay(ngp_handle* h, const ngp_replay* log)
{
    if (!h) return NGP_EINVAL;
    CU(cudaSetDevice(h->device));
    h->replay = 0; h->replay_iters = 0;
    if (!log) { h->sets_dirty = true; return NGP_OK; }
    if (log->n_iter <= 0 || !log->chi2_e) return fail(h, NGP_EINVAL, "ngp_set_replay: n_iter must be > 0 and chi2_e non-NULL");
    const int ni = log->n_iter;
    CU(dalloc(&h->rp_chi2_e, ni));
    CU(dalloc(&h->rp_z_mu, ni));
    CU(cpy(h, h->rp_chi2_e, log->chi2_e, sizeof(double) * ni, cudaMemcpyHostToDevice));
    if (log->z_mu) CU(cpy(h, h->rp_z_mu, log->z_mu, sizeof(double) * ni, cudaMemcpyHostToDevice));
    else if (h->has_mu) return fail(h, NGP_EINVAL, "ngp_set_replay: z_mu is required when the intercept is enabled");
    for (int s = 0; s < h->n_sets; ++s) {
        SetHost& S = h->sets[s];
        if (!S.have_geno || S.joint_member || S.joint_copy) continue;
        if (S.method == NGP_BAYESRCPI || S.method == NGP_BAYESRCPLUS) continue;        // their log comes through ngp_set_rc_replay
        if (!S.have_prior) return fail(h, NGP_EINVAL, "ngp_set_replay: set the prior of set %d first", s);
        if (s >= log->n_sets || !log->z[s] || !log->chi2_b[s]) return fail(h, NGP_EINVAL, "ngp_set_replay: z / chi2_b missing for set %d", s);
        CU(dalloc(&S.rp_z, (size_t)ni * S.p));
        CU(cpy(h, S.rp_z, log->z[s], sizeof(double) * (size_t)ni * S.p, cudaMemcpyHostToDevice));
        CU(dalloc(&S.rp_chi2b, (size_t)ni * S.nvar));
        CU(cpy(h, S.rp_chi2b, log->chi2_b[s], sizeof(double) * (size_t)ni * S.nvar, cudaMemcpyHostToDevice));
        if (S.method != NGP_BAYESPR) {
            if (!log->u[s]) return fail(h, NGP_EINVAL, "ngp_set_replay: u missing for set %d", s);
            const size_t per_u = (S.method == NGP_BAYESR) ? (size_t)S.n_class : 1, per_pi = per_u;
            CU(dalloc(&S.rp_u, (size_t)ni * S.p * per_u));
            CU(cpy(h, S.rp_u, log->u[s], sizeof(double) * (size_t)ni * S.p * per_u, cudaMemcpyHostToDevice));
            if (S.est_pi) {
                if (!log->beta_pi[s]) return fail(h, NGP_EINVAL, "ngp_set_replay: beta_pi missing for set %d", s);
                CU(dalloc(&S.rp_betapi, ni * per_pi));
                CU(cpy(h, S.rp_betapi, log->beta_pi[s], sizeof(double) * ni * per_pi, cudaMemcpyHostToDevice));
            }
        }
    }
    Scalars sc;
    CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
    h->replay = 1; h->replay_iters = ni; h->replay_base = sc.iter;
    h->sets_dirty = true;
    return NGP_OK;
}

// ----------------------------------------------------------------------------- launch
static int sync_sets(ngp_handle* h)
{
    if (!h->sets_dirty) return NGP_OK;
    SetDev sd[NGP_MAX_SETS];
    memset(sd, 0, sizeof sd);
    for (int s = 0; s < h->n_sets; ++s) {
        const SetHost& S = h->sets[s];
        if (!S.have_geno || !S.have_prior) continue;
        SetDev& D = sd[s];
        D.p = S.p; D.p_pad = S.p_pad; D.method = S.method; D.est_pi = S.est_pi; D.n_regions = S.n_regions; D.nvar = S.nvar;
        D.df = S.df; D.scale = S.scale; D.group_k = S.group_k; D.stream_set = S.stream_set; D.jinvB = S.jinvB;
        if (S.group_k) { D.jvar = h->joint.varBeta; for (int a = 0; a < S.group_k * S.group_k; ++a) D.jscale[a] = h->joint.scale[a]; D.rp_iw_chi2 = h->joint.rp_iw_chi2; D.rp_iw_z = h->joint.rp_iw_z; D.rp_z = h->joint.rp_z; }
        D.n_class = S.n_class; D.pi_class = S.pi_class; memcpy(D.v_class, S.v_class, sizeof D.v_class);
        D.n_annot = S.n_annot; D.annot = S.annot; D.annot_prob = S.annot_prob; D.annot_cat = S.annot_cat; D.rp_u_annot = S.rp_u_annot; D.rp_dirp = S.rp_dirp;
        D.geno = S.geno; D.gx = S.gx; D.consts = S.consts; D.colsum = S.colsum; D.d = S.d; D.mean = S.mean;
        if (h->w && S.w_ready) { D.d = S.dw; D.d_unw = S.d; D.wcs = S.wcs; }
        D.beta = S.beta; D.delta = S.delta; D.varBeta = S.varBeta; D.pi = S.pi; D.region_of = S.region_of; D.region_off = S.region_off;
        D.lhs0 = S.lhs0; D.rhs0 = S.rhs0;
        D.rp_u = S.rp_u; D.rp_z = S.group_k ? h->joint.rp_z : S.rp_z; D.rp_chi2b = S.rp_chi2b; D.rp_betapi = S.rp_betapi;
        D.sum_beta = S.sum_beta; D.sum_beta2 = S.sum_beta2; D.sum_delta = S.sum_delta;
    }
    CU(cudaMemcpyAsync(h->sets_dev, sd, sizeof sd, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->sets_dirty = false;
    return NGP_OK;
}

static void fill_params(ngp_handle* h, Params& P, int n_iter, int set_mask, int do_varE, int do_mu, double varE_in, int accumulate)
{
    P.n = h->n; P.Tw = h->Tw; P.R = h->R; P.B = h->B; P.n_sets = h->n_sets; P.kernel = h->cfg_kernel;
    P.D = h->D; P.DN = h->DN; P.NT = h->NT; P.NR = h->NR; P.NV = h->NV; P.refetch = h->refetch; P.store2 = h->store2 > 0;
    P.e = h->e; P.sets = h->sets_dev; P.sc = h->sc; P.sync = h->sync;
    P.df_e = h->df_e; P.scale_e = h->scale_e; P.has_mu = h->has_mu; P.do_varE = do_varE; P.do_mu = do_mu; P.set_mask = set_mask;
    P.mu_lhs0 = h->mu_lhs0; P.mu_rhs0 = h->mu_rhs0; P.varE_in = varE_in; P.n_iter = n_iter; P.replay = h->replay;
    P.replay_base = h->replay_base; P.rp_chi2_e = h->rp_chi2_e; P.rp_z_mu = h->rp_z_mu;
    P.key0 = (uint32_t)(h->seed & 0xffffffffu); P.key1 = (uint32_t)(h->seed >> 32); P.chain = h->chain; P.accumulate = accumulate;
    P.debug = h->cfg_debug; P.opt = h->cfg_opt >= 0 ? h->cfg_opt : (h->refetch ? 2 : 3);
    P.fx = h->fx; P.fx.rp_z = h->fx_rp_z;
    P.w = h->w; P.w_sum = h->w_sum; P.w_min = h->w_min; P.w_max = h->w_max;
    if (h->w && h->fx.n_cols && h->fx_w_ready) { P.fx.xpx = h->fx_xpx_w; P.fx.colsum_w = h->fx_colsum_w; }
    P.n_ranks = h->shard_world; P.rank = h->shard_rank; P.n_total = h->shard_world > 1 ? h->n_total : h->n;
    P.cta_off = 0; P.T_all = h->Tw + 1; P.Tw_all = h->Tw; P.bar_base = 0; P.cnt_bits = kCntBits;
    P.peer[0] = h->sync;
    if (h->shard_world > 1) {
        P.T_all = 0; P.Tw_all = 0;
        for (int r = 0; r < h->shard_world; ++r) {
            if (r == h->shard_rank) P.cta_off = P.T_all;
            P.T_all += h->shard_Tw[r] + 1; P.Tw_all += h->shard_Tw[r];
            P.peer[r] = h->peer[r];
        }
        P.bar_base = h->bar_count * (unsigned long long)P.T_all;
        if (P.Tw_all > 255) P.cnt_bits = 11;
        P.hier = !(h->cfg_opt >= 0 && (h->cfg_opt & 512));   // rank-local pre-reduction: on (C5 over 2 GPUs: 1.85 -> 1.60 ms/sweep; over 8: 2.70 -> 0.89)              // up to 2047 arrivals per accumulator (three bits less of fixed-point resolution)
    }
}

static int kernel_error_code(ngp_handle* h, int kerr);
static int check_kernel_error(ngp_handle* h)
{
    int kerr = 0;
    CU(cpy(h, &kerr, &h->sync->err, sizeof(int), cudaMemcpyDeviceToHost));
    return kernel_error_code(h, kerr);
}
static int kernel_error_code(ngp_handle* h, int kerr)
{
    if (kerr & 8) return fail(h, NGP_ETIMEOUT, "row-sharded chain: a rank did not arrive within 4 s (results invalid)");
    if (kerr & 2) return fail(h, NGP_ENUMERIC, "a covariance matrix of the tuple sampler is not positive definite");
    if (kerr & 4) return fail(h, NGP_ENUMERIC, "BayesR: no class reached its uniform (the reference's findfirst returns nothing here)");
    if (kerr) return fail(h, NGP_ERANGE, "fixed-point reduction range exceeded (residual grew by more than 2^4 within an iteration)");
    return NGP_OK;
}

// whole iterations (or one sweep) of the tuple sampler: ngp_joint.cuh
static int launch_joint(ngp_handle* h, int n_iter, int do_varE, int do_mu, double varE_in, int accumulate)
{
    CU(cudaSetDevice(h->device));
    JointHost& Jh = h->joint;
    if (!Jh.active) return fail(h, NGP_EINVAL, "no tuple of marker sets (ngp_set_joint_prior)");
    if (h->store2 > 0) return fail(h, NGP_EUNSUPPORTED, "the tuple sampler reads int8 device storage (upload with NGP_STORE_I8)");
    if (!h->have_y) return fail(h, NGP_EINVAL, "no phenotype / residual on the device (ngp_set_phenotype or ngp_joint_sweep)");
    if (h->replay) {
        Scalars sc;
        CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
        if (sc.iter - h->replay_base + n_iter > h->replay_iters || !Jh.rp_z || Jh.replay_iters != h->replay_iters)
            return fail(h, NGP_EINVAL, "replay log of the tuple missing or exhausted (ngp_set_joint_replay after ngp_set_replay)");
    }
    int rc = sync_sets(h);
    if (rc) return rc;
    Params P{};
    fill_params(h, P, n_iter, 0, do_varE, do_mu, varE_in, accumulate);
    JointDev J{};
    J.k = Jh.k; J.stream_set = Jh.set[0]; J.p = Jh.p; J.n_regions = Jh.n_regions; J.df = Jh.df;
    for (int b = 0; b < Jh.k; ++b) J.set[b] = Jh.set[b];
    memcpy(J.scale, Jh.scale, sizeof J.scale);
    J.varBeta = Jh.varBeta; J.region_off = Jh.region_off; J.mtm = Jh.mtm;
    J.rp_z = Jh.rp_z; J.rp_iw_chi2 = Jh.rp_iw_chi2; J.rp_iw_z = Jh.rp_iw_z;
    const size_t smem = sizeof(double) * (size_t)(160 + kSlots * kMaxK + h->R);
    const void* kfn = ngp_joint_kernel(Jh.k);
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kThreads, smem));
    if (per_sm * h->prop.multiProcessorCount < h->Tw + 1)
        return fail(h, NGP_EUNSUPPORTED, "cooperative grid of %d CTAs does not fit (%d per SM x %d SMs)", h->Tw + 1, per_sm, h->prop.multiProcessorCount);
    CU(cudaMemsetAsync(h->sync, 0, kSyncHeadFixed + sizeof(double) * 2 * (size_t)(h->Tw + 1), h->stream));
    void* args[] = {&P, &J};
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(cudaLaunchCooperativeKernel(kfn, dim3(h->Tw + 1), dim3(kThreads), args, smem, h->stream));
    CU(cudaEventRecord(h->ev1, h->stream));
    h->launches += 1;
    h->timed = true;
    CU(cudaStreamSynchronize(h->stream));
    return check_kernel_error(h);
}

// weighted residuals: mpm_j = sum_i w_i x_ij^2 and sum_i w_i x_ij of a marker set (mme.jl:299-303), once per weight vector
static int ensure_weighted_moments(ngp_handle* h, int s)
{
    SetHost& S = h->sets[s];
    if (!h->w || !S.have_geno || S.w_ready) return NGP_OK;
    if (h->store2 > 0) return NGP_OK;                  // (2-bit tiles: the weighted path is refused at launch)
    if (!S.dw) { CU(dalloc(&S.dw, S.p_pad)); CU(dalloc(&S.wcs, S.p_pad)); }
    CU(zero(h, S.dw, 0, sizeof(double) * S.p_pad));
    CU(zero(h, S.wcs, 0, sizeof(double) * S.p_pad));
    weighted_moments_kernel<<<(unsigned)S.p, 256, 0, h->stream>>>(S.geno, h->n, h->R, h->B, S.p_pad / h->B, S.p, S.mean, h->w, S.dw, S.wcs);
    CU(cudaGetLastError());
    S.w_ready = true;
    h->sets_dirty = true;
    return NGP_OK;
}

// everything up to the launch itself: checks, weighted set-up, Params, kernel variant, bookkeeping of the global block / barrier counters
static int launch_prepare(ngp_handle* h, int n_iter, int set_mask, int do_varE, int do_mu, double varE_in, int accumulate, Params& P, const void*& kfn, bool group)
{
    CU(cudaSetDevice(h->device));
    if (h->joint.active && !(h->joint.blocked_set >= 0 && set_mask == (1 << h->joint.blocked_set)))
        return fail(h, NGP_EINVAL, "this handle samples a tuple of marker sets: use ngp_run / ngp_joint_sweep");
    const bool sharded = h->shard_world > 1;
    { int rg = apply_ring_geometry(h); if (rg) return rg; }
    if (sharded && !h->shard_attached) return fail(h, NGP_EINVAL, "row-sharded handle: call ngp_shard_attach before sampling");
    if (sharded && h->fx.n_cols) return fail(h, NGP_EUNSUPPORTED, "fixed effects besides the intercept are not available on a row-sharded handle");
    if (h->replay && h->fx.n_cols && do_mu && (!h->fx_rp_z || h->fx_replay_iters != h->replay_iters))
        return fail(h, NGP_EINVAL, "replay log of the fixed effects missing (ngp_set_fixed_replay after ngp_set_replay)");
    if (sharded && h->cfg_kernel != NGP_KERNEL_LITERAL)
        for (int s = 0; s < h->n_sets; ++s)
            if (((set_mask >> s) & 1) && !h->sets[s].gram_global)
                return fail(h, NGP_EINVAL, "row-sharded chain on the blocked kernel: the banded Gram of set %d holds this rank's rows only — sum it over "
                                           "the ranks (ngp_get_gram -> all-reduce -> ngp_set_gram), or use NGP_KERNEL_LITERAL", s);
    if (sharded && h->shard_same_device && !group)
        return fail(h, NGP_EUNSUPPORTED, "shards of one chain on the same device wait for one another: separate launches are not guaranteed to be "
                                         "co-resident, run them as one grid with ngp_run_group");
    if (!h->have_y) return fail(h, NGP_EINVAL, "no phenotype / residual on the device (ngp_set_phenotype or ngp_sweep)");
    int active = 0;
    for (int s = 0; s < h->n_sets; ++s)
        if ((set_mask >> s) & 1) {
            if (!h->sets[s].have_geno || !h->sets[s].have_prior) return fail(h, NGP_EINVAL, "marker set %d has no genotypes or no prior", s);
            ++active;
        }
    if (!active) return fail(h, NGP_EINVAL, "no marker set to sample");
    if (h->replay) {
        Scalars sc;
        CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
        if (sc.iter - h->replay_base + n_iter > h->replay_iters) return fail(h, NGP_EINVAL, "replay log exhausted (%d iterations)", h->replay_iters);
    }
    if (h->w) {
        if (sharded || h->joint.active)
            return fail(h, NGP_EUNSUPPORTED, "weighted residuals: available for fixed effects + marker-set models on one GPU (no tuple, no row sharding)");
        if (h->fx.n_cols && !h->fx_w_ready) {              // X[xSet].xpx = X'(w .* X), and w'x_c for the running 1'We (mme.jl:135)
            const FxDev& F = h->fx;
            const int64_t n = h->n;
            std::vector<double> xw((size_t)kMaxFxCols * kMaxFxCols, 0.0), cw((size_t)F.n_cols, 0.0);
            for (int s = 0; s < F.n_sets; ++s) {
                const int c0 = F.first[s], nc = F.first[s + 1] - c0;
                for (int a = 0; a < nc; ++a) {
                    const double* xa = h->fx_host.data() + (size_t)(c0 + a) * n;
                    double sw = 0.0;
                    for (int64_t i = 0; i < n; ++i) sw += h->w_host[i] * xa[i];
                    cw[(size_t)(c0 + a)] = sw;
                    for (int b = a; b < nc; ++b) {
                        const double* xb = h->fx_host.data() + (size_t)(c0 + b) * n;
                        double d = 0.0;
                        for (int64_t i = 0; i < n; ++i) d += xa[i] * (h->w_host[i] * xb[i]);
                        xw[(size_t)F.xoff[s] + a * nc + b] = d; xw[(size_t)F.xoff[s] + b * nc + a] = d;
                    }
                }
            }
            if (!h->fx_xpx_w) { CU(dalloc(&h->fx_xpx_w, xw.size())); CU(dalloc(&h->fx_colsum_w, (size_t)kMaxFxCols)); }
            CU(cpy(h, h->fx_xpx_w, xw.data(), sizeof(double) * xw.size(), cudaMemcpyHostToDevice));
            CU(cpy(h, h->fx_colsum_w, cw.data(), sizeof(double) * cw.size(), cudaMemcpyHostToDevice));
            CU(cudaStreamSynchronize(h->stream));
            h->fx_w_ready = true;
        }
        for (int s = 0; s < h->n_sets; ++s)
            if ((set_mask >> s) & 1) { int rcw = ensure_weighted_moments(h, s); if (rcw) return rcw; }
    }
    int rc = sync_sets(h);
    if (rc) return rc;
    P = Params{};
    fill_params(h, P, n_iter, set_mask, do_varE, do_mu, varE_in, accumulate);
    if (h->w) P.kernel = NGP_KERNEL_LITERAL;            // weighted dots are not integer sums of codes: per-marker sweep
    // BayesR: the blocked sweep has its own instantiation (class algebra per lane in the chain warp) for <= 4 classes without summary-statistic
    // priors, when every set of the launch is BayesR; anything else goes to the per-marker sweep
    bool any_r = false, all_r = true, r_ok = true;
    for (int s = 0; s < h->n_sets; ++s)
        if ((set_mask >> s) & 1) {
            const SetHost& S = h->sets[s];
            if (S.method == NGP_BAYESR) { any_r = true; if (S.n_class > 4 || S.lhs0 || S.rhs0) r_ok = false; } else all_r = false;
        }
    const bool r_blocked = any_r && all_r && r_ok && P.kernel == NGP_KERNEL_BLOCKED && !h->w && !sharded && !h->cfg_debug && !h->cfg_profile &&
                           !(h->R > 4 * kUpdThreads && h->B != 16);
    if (any_r && !r_blocked) P.kernel = NGP_KERNEL_LITERAL;
    for (int s = 0; s < h->n_sets; ++s)
        if (((set_mask >> s) & 1) && (h->sets[s].method == NGP_BAYESRCPI || h->sets[s].method == NGP_BAYESRCPLUS)) {
            if (h->w || sharded) return fail(h, NGP_EUNSUPPORTED, "BayesRCpi / BayesRCplus: no weighted residuals, no row sharding");
            if (h->replay && (!h->sets[s].rp_u || !h->sets[s].rp_z)) return fail(h, NGP_EINVAL, "replay log of the RC set %d missing (ngp_set_rc_replay after ngp_set_replay)", s);
            P.kernel = NGP_KERNEL_LITERAL;          // annotation algebra lives in the per-marker sweep
        }
    bool tuple_mask = false;
    for (int s = 0; s < h->n_sets; ++s) if (((set_mask >> s) & 1) && h->sets[s].group_k) tuple_mask = true;
    int variant = tuple_mask ? NGP_KV_TUP : (P.kernel == NGP_KERNEL_LITERAL) ? NGP_KV_LIT : h->cfg_debug ? NGP_KV_DBG : h->cfg_profile ? NGP_KV_PROF : NGP_KV_PLAIN;
    if (r_blocked) variant = NGP_KV_R;
    if (h->R > 4 * kUpdThreads && h->B != 16 && variant != NGP_KV_LIT) {   // more than 512 rows per CTA with blocks of 32 / 64: the instantiation with 4 row groups per updater thread (the per-marker sweep keeps e in shared memory: any panel)
        if (variant != NGP_KV_PLAIN) return fail(h, NGP_EUNSUPPORTED, "%d rows per CTA with blocks of %d: only the plain blocked sweep is built for this geometry (use blocks of 16)", h->R, h->B);
        variant = (h->refetch && h->NT == 1 && !sharded) ? NGP_KV_BIGR1 : NGP_KV_BIGR;      // one tile in the ring: its dots are shared by the 8 dot warps
    }
    if (sharded && variant != NGP_KV_LIT) {               // the row-sharded blocked sweep is its own instantiation (SH): the one-GPU kernels do not carry its code
        if (variant != NGP_KV_PLAIN && variant != NGP_KV_BIGR) return fail(h, NGP_EUNSUPPORTED, "row-sharded chain: the plain blocked sweep or the per-marker sweep (no profile / debug / tuple / BayesR-blocked instantiation)");
        if (!group) variant = (variant == NGP_KV_BIGR) ? NGP_KV_SHARD_BIGR : NGP_KV_SHARD;
    }
    if (h->store2 > 0 && (variant == NGP_KV_LIT || variant == NGP_KV_TUP))   // (the blocked BayesR sweep reads 2-bit tiles like the plain one)
        return fail(h, NGP_EUNSUPPORTED, "2-bit device storage serves the blocked sweep of BayesPR / BayesB / BayesC sets (not the per-marker kernel: "
                                         "BayesR, weighted residuals, row sharding; not the tuple sampler): upload with NGP_STORE_I8");
    if (group && variant != NGP_KV_LIT && variant != NGP_KV_PLAIN) return fail(h, NGP_EUNSUPPORTED, "ngp_run_group runs the plain blocked or the per-marker sweep");
    const int gvariant = group ? (variant == NGP_KV_LIT ? NGP_KV_GROUP : NGP_KV_GROUPB) : variant;
    kfn = ngp_gibbs_kernel(h->B, gvariant);
    h->last_variant = gvariant;
    if (!group && h->ready_kfn != kfn) {                      // once per kernel variant: attribute + co-residency check of the cooperative grid
        // the attribute belongs to the function, not to the handle: always the device maximum, so that handles with different
        // geometries never lower it under each other
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, kfn));
        CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kThreads, h->L.total));
        if (per_sm * h->prop.multiProcessorCount < h->Tw + 1)
            return fail(h, NGP_EUNSUPPORTED, "cooperative grid of %d CTAs does not fit (%d per SM x %d SMs)", h->Tw + 1, per_sm, h->prop.multiProcessorCount);
        h->ready_kfn = kfn;
    }
    uint64_t blocks = 0;
    for (int s = 0; s < h->n_sets; ++s) if ((set_mask >> s) & 1) blocks += (uint64_t)(h->sets[s].p_pad / h->B);
    blocks *= (uint64_t)n_iter;
    if (h->gblk + blocks >= 0xffff0000ull) {          // 32-bit sequence numbers: start over with clean rings
        CU(cudaMemsetAsync(h->sync, 0, sizeof(SyncArea), h->stream));
        h->gblk = 0;
    }
    P.gblk0 = (uint32_t)h->gblk;
    h->gblk += blocks;
    if (sharded) {
        // peers may already be arriving on this rank's counter: never reset it; account for the barrier rounds of this launch
        CU(cudaMemsetAsync(&h->sync->err, 0, sizeof(int), h->stream));
        uint64_t rounds = 2;
        for (int s = 0; s < h->n_sets; ++s)
            if ((set_mask >> s) & 1) rounds += ((h->sets[s].method == NGP_BAYESPR && h->sets[s].n_regions > 1) || accumulate) ? 1 : 0;
        h->bar_count += rounds * (uint64_t)n_iter;
    } else CU(cudaMemsetAsync(h->sync, 0, kSyncHeadFixed + sizeof(double) * 2 * (size_t)(h->Tw + 1), h->stream));
    return NGP_OK;
}

static int launch(ngp_handle* h, int n_iter, int set_mask, int do_varE, int do_mu, double varE_in, int accumulate, bool defer_sync = false)
{
    Params P{};
    const void* kfn = nullptr;
    int rc = launch_prepare(h, n_iter, set_mask, do_varE, do_mu, varE_in, accumulate, P, kfn, false);
    if (rc) return rc;
    void* args[] = {&P};
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(cudaLaunchCooperativeKernel(kfn, dim3(h->Tw + 1), dim3(kThreads), args, (size_t)h->L.total, h->stream));
    CU(cudaEventRecord(h->ev1, h->stream));
    h->launches += 1;
    h->timed = true;
    if (defer_sync) return NGP_OK;                 // the caller queues its device-to-host copies first and synchronises once
    CU(cudaMemcpyAsync(h->err_pinned, &h->sync->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    { float ms = 0.f; if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->sum_run_ms += ms; }
    return kernel_error_code(h, *h->err_pinned);
}

// the tuple's effects live either in the member sets (per-locus kernel) or in the interleaved copy (blocked kernel)
static int tuple_sync_state(ngp_handle* h, bool want_copy)
{
    JointHost& J = h->joint;
    if (J.blocked_set < 0 || J.state_in_copy == want_copy) return NGP_OK;
    TupleBeta M{};
    for (int b = 0; b < J.k; ++b) M.member[b] = h->sets[J.set[b]].beta;
    tuple_beta_kernel<<<(unsigned)((J.p + 255) / 256), 256, 0, h->stream>>>(M, h->sets[J.blocked_set].beta, J.k, J.p, want_copy ? 1 : 0);
    CU(cudaGetLastError());
    J.state_in_copy = want_copy;
    return NGP_OK;
}

static bool tuple_use_blocked(const ngp_handle* h) { return h->joint.blocked_set >= 0 && h->cfg_kernel == NGP_KERNEL_BLOCKED; }

static int tuple_launch(ngp_handle* h, int n_iter, int do_varE, int do_mu, double varE_in, int accumulate)
{
    if (h->w) return fail(h, NGP_EUNSUPPORTED, "weighted residuals are not available for a tuple of marker sets");
    JointHost& J = h->joint;
    if (tuple_use_blocked(h)) {
        if (h->replay && (!J.rp_z || J.replay_iters != h->replay_iters))
            return fail(h, NGP_EINVAL, "replay log of the tuple missing (ngp_set_joint_replay after ngp_set_replay)");
        int rc = tuple_sync_state(h, true);
        if (rc) return rc;
        h->sets_dirty = true;                       // replay / covariance pointers of the copy follow the tuple's
        return launch(h, n_iter, 1 << J.blocked_set, do_varE, do_mu, varE_in, accumulate);
    }
    int rc = tuple_sync_state(h, false);
    if (rc) return rc;
    return launch_joint(h, n_iter, do_varE, do_mu, varE_in, accumulate);
}

int ngp_run(ngp_handle* h, int32_t n_iter)
{
    if (!h) return NGP_EINVAL;
    if (n_iter <= 0) return fail(h, NGP_EINVAL, "ngp_run: n_iter must be positive");
    if (h->joint.active) return tuple_launch(h, n_iter, 1, 1, 0.0, 1);
    int mask = 0;
    for (int s = 0; s < h->n_sets; ++s) if (h->sets[s].have_geno) mask |= 1 << s;
    return launch(h, n_iter, mask, 1, 1, 0.0, 1);
}

static int push_set_state(ngp_handle* h, int s, const double* beta, const int64_t* delta, const double* varBeta, const double* piHat)
{
    SetHost& S = h->sets[s];
    if (beta) CU(cpy(h, S.beta, beta, sizeof(double) * S.p, cudaMemcpyHostToDevice));
    if (delta) {
        std::vector<int32_t> d32((size_t)S.p);
        for (int64_t j = 0; j < S.p; ++j) d32[(size_t)j] = (int32_t)delta[j];
        CU(cpy(h, S.delta, d32.data(), sizeof(int32_t) * S.p, cudaMemcpyHostToDevice));
    }
    if (varBeta && S.varBeta) CU(cpy(h, S.varBeta, varBeta, sizeof(double) * S.nvar, cudaMemcpyHostToDevice));
    if (piHat && S.method == NGP_BAYESR) {
        double pc[2 * kMaxClass];
        for (int v = 0; v < S.n_class; ++v) { pc[v] = piHat[v]; pc[S.n_class + v] = log(piHat[v]); }
        CU(cpy(h, S.pi_class, pc, sizeof(double) * 2 * S.n_class, cudaMemcpyHostToDevice));
    } else if (piHat && S.method != NGP_BAYESPR) {
        double pi4[4] = {piHat[0], piHat[1], log(piHat[0]), log(piHat[1])};
        CU(cpy(h, S.pi, pi4, sizeof pi4, cudaMemcpyHostToDevice));
    }
    return NGP_OK;
}

static int pull_set_state(ngp_handle* h, int s, double* beta, int64_t* delta, double* varBeta, double* piHat)
{
    SetHost& S = h->sets[s];
    if (beta) CU(cpy(h, beta, S.beta, sizeof(double) * S.p, cudaMemcpyDeviceToHost));
    if (delta) {
        std::vector<int32_t> d32((size_t)S.p);
        CU(cpy(h, d32.data(), S.delta, sizeof(int32_t) * S.p, cudaMemcpyDeviceToHost));
        for (int64_t j = 0; j < S.p; ++j) delta[j] = d32[(size_t)j];
    }
    if (varBeta && S.varBeta) CU(cpy(h, varBeta, S.varBeta, sizeof(double) * S.nvar, cudaMemcpyDeviceToHost));
    if (piHat && S.method == NGP_BAYESR && S.pi_class) {
        CU(cpy(h, piHat, S.pi_class, sizeof(double) * S.n_class, cudaMemcpyDeviceToHost));
    } else if (piHat && S.pi) {
        double pi4[4];
        CU(cpy(h, pi4, S.pi, sizeof pi4, cudaMemcpyDeviceToHost));
        piHat[0] = pi4[0]; piHat[1] = pi4[1];
    }
    return NGP_OK;
}

int ngp_sweep(ngp_handle* h, int set_id, double* ycorr, double varE, double* beta, int64_t* delta, double* varBeta, double* piHat)
{
    if (!h) return NGP_EINVAL;
    if (set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno || !h->sets[set_id].have_prior)
        return fail(h, NGP_EINVAL, "ngp_sweep: marker set %d is not ready", set_id);
    if (!ycorr || !(varE > 0.0)) return fail(h, NGP_EINVAL, "ngp_sweep: ycorr must be non-NULL and varE > 0");
    if (h->sets[set_id].method == NGP_BAYESRCPI || h->sets[set_id].method == NGP_BAYESRCPLUS)
        return fail(h, NGP_EUNSUPPORTED, "ngp_sweep: BayesRCpi / BayesRCplus sets are sampled at run level (ngp_run; state through ngp_get_state / ngp_get_rc_state)");
    CU(cudaSetDevice(h->device));
    SetHost& S = h->sets[set_id];
    // Everything is queued on the handle's stream — host-to-device copies, the sweep, device-to-host copies — and the stream is
    // synchronised ONCE.  delta (Int64 on the Julia side, int32 on device), pi and the error flag go through a pinned staging buffer.
    const size_t npi = (S.method == NGP_BAYESR) ? 2 * (size_t)S.n_class : 4;
    const size_t need = sizeof(int32_t) * (size_t)S.p_pad + sizeof(double) * 2 * kMaxClass * 2 + 64;
    if (h->stage_bytes < need) {
        if (h->stage) cudaFreeHost(h->stage);
        h->stage = nullptr; h->stage_bytes = 0;
        CU(cudaMallocHost((void**)&h->stage, need));
        h->stage_bytes = need;
    }
    int32_t* st_delta = reinterpret_cast<int32_t*>(h->stage);
    double* st_pi_in = reinterpret_cast<double*>(h->stage + sizeof(int32_t) * (size_t)S.p_pad);
    double* st_pi_out = st_pi_in + 2 * kMaxClass;
    int* st_err = reinterpret_cast<int*>(st_pi_out + 2 * kMaxClass);
    double* dev_pi = (S.method == NGP_BAYESR) ? S.pi_class : S.pi;
    CU(cudaMemcpyAsync(h->e, ycorr, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
    if (beta) CU(cudaMemcpyAsync(S.beta, beta, sizeof(double) * S.p, cudaMemcpyHostToDevice, h->stream));
    if (delta) {
        for (int64_t j = 0; j < S.p; ++j) st_delta[j] = (int32_t)delta[j];
        CU(cudaMemcpyAsync(S.delta, st_delta, sizeof(int32_t) * S.p, cudaMemcpyHostToDevice, h->stream));
    }
    if (varBeta && S.varBeta) CU(cudaMemcpyAsync(S.varBeta, varBeta, sizeof(double) * S.nvar, cudaMemcpyHostToDevice, h->stream));
    if (piHat && S.method != NGP_BAYESPR && dev_pi) {
        const size_t half = npi / 2;
        for (size_t v = 0; v < half; ++v) { st_pi_in[v] = piHat[v]; st_pi_in[half + v] = log(piHat[v]); }
        CU(cudaMemcpyAsync(dev_pi, st_pi_in, sizeof(double) * npi, cudaMemcpyHostToDevice, h->stream));
    }
    h->have_y = true;
    int rc = launch(h, 1, 1 << set_id, 0, 0, varE, 0, /*defer_sync=*/true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ycorr, h->e, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
    if (beta) CU(cudaMemcpyAsync(beta, S.beta, sizeof(double) * S.p, cudaMemcpyDeviceToHost, h->stream));
    if (delta) CU(cudaMemcpyAsync(st_delta, S.delta, sizeof(int32_t) * S.p, cudaMemcpyDeviceToHost, h->stream));
    if (varBeta && S.varBeta) CU(cudaMemcpyAsync(varBeta, S.varBeta, sizeof(double) * S.nvar, cudaMemcpyDeviceToHost, h->stream));
    if (piHat && dev_pi) CU(cudaMemcpyAsync(st_pi_out, dev_pi, sizeof(double) * npi, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(st_err, &h->sync->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (*st_err & 2) return fail(h, NGP_ENUMERIC, "a covariance matrix of the tuple sampler is not positive definite");
    if (*st_err & 4) return fail(h, NGP_ENUMERIC, "BayesR: no class reached its uniform (the reference's findfirst returns nothing here)");
    if (*st_err) return fail(h, NGP_ERANGE, "fixed-point reduction range exceeded (residual grew by more than 2^4 within an iteration)");
    if (delta) for (int64_t j = 0; j < S.p; ++j) delta[j] = st_delta[j];
    if (piHat && dev_pi) { const size_t half = npi / 2; for (size_t v = 0; v < half; ++v) piHat[v] = st_pi_out[v]; }
    return NGP_OK;
}

// ----------------------------------------------------------------------------- row-sharded chain
int ngp_shard_init(ngp_handle* h, int rank, int world)
{
    if (!h) return NGP_EINVAL;
    if (h->Tw) return fail(h, NGP_EINVAL, "ngp_shard_init: must be called before the first upload");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(h, NGP_EINVAL, "ngp_shard_init: rank %d / world %d out of range (<= %d)", rank, world, kMaxRanks);
    h->shard_rank = rank; h->shard_world = world; h->shard_attached = false;
    return NGP_OK;
}

// All shards of a chain that live on ONE device: one cooperative grid over all ranks' row slices (ngp::gibbs_group_kernel).
int ngp_run_group(ngp_handle** hs, int n_handles, int32_t n_iter)
{
    if (!hs || n_handles < 1 || !hs[0]) return NGP_EINVAL;
    ngp_handle* h = hs[0];
    if (n_iter <= 0) return fail(h, NGP_EINVAL, "ngp_run_group: n_iter must be positive");
    if (n_handles != h->shard_world) return fail(h, NGP_EINVAL, "ngp_run_group: %d handles for a chain of %d shards", n_handles, h->shard_world);
    int total = 0;
    size_t smem = 0;
    for (int r = 0; r < n_handles; ++r) {
        ngp_handle* g = hs[r];
        if (!g || g->device != h->device || g->shard_world != n_handles || g->shard_rank != r || g->B != h->B)
            return fail(h, NGP_EINVAL, "ngp_run_group: handle %d is not rank %d of this chain on device %d (same block size)", r, r, h->device);
        if (g->joint.active) return fail(h, NGP_EUNSUPPORTED, "ngp_run_group: no tuple sampler on a row-sharded chain");
        total += g->Tw + 1;
        smem = std::max(smem, (size_t)g->L.total);
    }
    CU(cudaSetDevice(h->device));
    std::vector<Params> Ps((size_t)n_handles);
    const void* kfn = nullptr;
    // validate every shard before anything is launched
    for (int r = 0; r < n_handles; ++r) {
        ngp_handle* g = hs[r];
        int mask = 0;
        for (int s = 0; s < g->n_sets; ++s) if (g->sets[s].have_geno) mask |= 1 << s;
        int rc = launch_prepare(g, n_iter, mask, 1, 1, 0.0, 1, Ps[(size_t)r], kfn, true);
        if (rc) { if (g != h) h->err = g->err; return rc; }
        if (g->stream != h->stream) CU(cudaStreamSynchronize(g->stream));      // its memsets / copies are ordered before the launch on hs[0]'s stream
    }
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kfn));
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kThreads, smem));
    if (per_sm * h->prop.multiProcessorCount < total)
        return fail(h, NGP_EUNSUPPORTED, "cooperative grid of %d CTAs (all shards) does not fit (%d per SM x %d SMs)", total, per_sm, h->prop.multiProcessorCount);
    Params* dPs = nullptr;
    CU(cudaMalloc((void**)&dPs, sizeof(Params) * (size_t)n_handles));
    CU(cudaMemcpyAsync(dPs, Ps.data(), sizeof(Params) * (size_t)n_handles, cudaMemcpyHostToDevice, h->stream));
    GroupParams G{dPs, n_handles};
    void* args[] = {&G};
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(cudaLaunchCooperativeKernel(kfn, dim3((unsigned)total), dim3(kThreads), args, smem, h->stream));
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(dPs);
    for (int r = 0; r < n_handles; ++r) {
        hs[r]->launches += 1;
        hs[r]->timed = (r == 0);
        int rc = check_kernel_error(hs[r]);
        if (rc) { if (hs[r] != h) h->err = hs[r]->err; return rc; }
    }
    return NGP_OK;
}

int ngp_shard_export(ngp_handle* h, ngp_shard_info* out)
{
    if (!h || !out) return fail(h, NGP_EINVAL, "ngp_shard_export: NULL argument");
    if (!h->Tw) return fail(h, NGP_EINVAL, "ngp_shard_export: upload the rank's rows first (fixes its CTA geometry)");
    CU(cudaSetDevice(h->device));
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t ih;
    CU(cudaIpcGetMemHandle(&ih, h->sync));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ngp_shard_info.ipc holds a cudaIpcMemHandle_t");
    memcpy(out->ipc, &ih, 64);
    out->n_local = h->n; out->worker_ctas = h->Tw; out->device = h->device; out->pid = (int64_t)getpid();
    out->local_ptr = (uint64_t)(uintptr_t)h->sync;
    return NGP_OK;
}

int ngp_shard_attach(ngp_handle* h, const ngp_shard_info* all)
{
    if (!h || !all) return fail(h, NGP_EINVAL, "ngp_shard_attach: NULL argument");
    if (h->shard_world < 2) return fail(h, NGP_EINVAL, "ngp_shard_attach: the handle is not sharded (ngp_shard_init)");
    if (!h->Tw) return fail(h, NGP_EINVAL, "ngp_shard_attach: upload the rank's rows first");
    CU(cudaSetDevice(h->device));
    int64_t ntot = 0, ctas = 0;
    for (int r = 0; r < h->shard_world; ++r) {
        const ngp_shard_info& I = all[r];
        if (I.worker_ctas <= 0 || I.n_local <= 0) return fail(h, NGP_EINVAL, "ngp_shard_attach: rank %d published no geometry", r);
        ntot += I.n_local; ctas += I.worker_ctas + 1;
        h->shard_Tw[r] = I.worker_ctas;
        if (r == h->shard_rank) { h->peer[r] = h->sync; h->peer_ipc[r] = false; continue; }
        if (I.pid == (int64_t)getpid()) {                       // same process: the pointer is valid here; map the peer device if needed
            if (I.device != h->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, h->device, I.device));
                if (!can) return fail(h, NGP_EUNSUPPORTED, "device %d cannot access device %d (no NVLink / P2P)", h->device, I.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(I.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
                cudaGetLastError();
            }
            h->peer[r] = (SyncArea*)(uintptr_t)I.local_ptr; h->peer_ipc[r] = false;
            if (I.device == h->device) h->shard_same_device = true;
        } else {                                               // one process per GPU: CUDA IPC mapping over NVLink
            cudaIpcMemHandle_t ih;
            memcpy(&ih, I.ipc, 64);
            void* ptr = nullptr;
            CU(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
            h->peer[r] = (SyncArea*)ptr; h->peer_ipc[r] = true;
        }
    }
    if (ctas > kMaxCtas) return fail(h, NGP_EUNSUPPORTED, "%lld CTAs over all ranks exceed %d", (long long)ctas, kMaxCtas);
    h->n_total = ntot;
    h->shard_attached = true;
    return NGP_OK;
}

int ngp_get_column_sums(ngp_handle* h, int set_id, int64_t* colsum, int64_t* colsumsq)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_get_column_sums: no such set");
    SetHost& S = h->sets[set_id];
    CU(cudaSetDevice(h->device));
    std::vector<int32_t> a((size_t)S.p), b((size_t)S.p);
    CU(cpy(h, a.data(), S.colsum, sizeof(int32_t) * S.p, cudaMemcpyDeviceToHost));
    CU(cpy(h, b.data(), S.colsumsq, sizeof(int32_t) * S.p, cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < S.p; ++j) { if (colsum) colsum[j] = a[(size_t)j]; if (colsumsq) colsumsq[j] = b[(size_t)j]; }
    return NGP_OK;
}

int ngp_set_column_sums(ngp_handle* h, int set_id, int64_t n_total, const int64_t* colsum, const int64_t* colsumsq)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_set_column_sums: no such set");
    if (!colsum || !colsumsq || n_total < h->n) return fail(h, NGP_EINVAL, "ngp_set_column_sums: bad argument");
    SetHost& S = h->sets[set_id];
    CU(cudaSetDevice(h->device));
    std::vector<int32_t> a((size_t)S.p), b((size_t)S.p);
    for (int64_t j = 0; j < S.p; ++j) {
        if (colsum[j] < 0 || colsum[j] > 0x7fffffffLL || colsumsq[j] < 0 || colsumsq[j] > 0x7fffffffLL) return fail(h, NGP_EINVAL, "ngp_set_column_sums: sum out of range at column %lld", (long long)j);
        a[(size_t)j] = (int32_t)colsum[j]; b[(size_t)j] = (int32_t)colsumsq[j];
    }
    CU(cpy(h, S.colsum, a.data(), sizeof(int32_t) * S.p, cudaMemcpyHostToDevice));
    CU(cpy(h, S.colsumsq, b.data(), sizeof(int32_t) * S.p, cudaMemcpyHostToDevice));
    // mean and mpm of the WHOLE column (all ranks' rows): prepMatVec.jl:129, mme.jl:305-307
    colstats_kernel<<<(unsigned)((S.p_pad + 255) / 256), 256, 0, h->stream>>>(n_total, S.p, S.p_pad, S.colsum, S.colsumsq, S.mean, S.d);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    return NGP_OK;
}

// Row-sharded chain on the blocked kernel: the banded raw Gram gx[k][d][a][b] = sum_i g_a g_b is a sum over individuals, so every rank needs the
// sum of all ranks' Grams (int32, (p_pad / B) * (D + 1) * B * B values; all ranks share B and D).  Host round trip at set-up.
int ngp_gram_size(ngp_handle* h, int set_id, int64_t* count)
{
    if (!h || !count || set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_gram_size: bad argument");
    *count = (h->sets[set_id].p_pad / h->B) * (int64_t)(h->D + 1) * h->B * h->B;
    return NGP_OK;
}
int ngp_get_gram(ngp_handle* h, int set_id, int32_t* out)
{
    int64_t cnt = 0;
    int rc = ngp_gram_size(h, set_id, &cnt);
    if (rc || !out) return rc ? rc : fail(h, NGP_EINVAL, "ngp_get_gram: out is NULL");
    CU(cudaSetDevice(h->device));
    CU(cpy(h, out, h->sets[set_id].gx, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost));
    return NGP_OK;
}
int ngp_set_gram(ngp_handle* h, int set_id, const int32_t* in)
{
    int64_t cnt = 0;
    int rc = ngp_gram_size(h, set_id, &cnt);
    if (rc || !in) return rc ? rc : fail(h, NGP_EINVAL, "ngp_set_gram: in is NULL");
    CU(cudaSetDevice(h->device));
    CU(cpy(h, h->sets[set_id].gx, in, sizeof(int32_t) * (size_t)cnt, cudaMemcpyHostToDevice));
    h->sets[set_id].gram_global = true;
    return NGP_OK;
}

int ngp_set_joint_prior(ngp_handle* h, const ngp_joint_prior* pr)
{
    if (!h || !pr) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: NULL argument");
    const int k = pr->k;
    if (k < 2 || k > kMaxK) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: k must be in [2,%d]", kMaxK);
    if (h->store2 > 0) return fail(h, NGP_EUNSUPPORTED, "the tuple sampler reads int8 device storage (upload with NGP_STORE_I8)");
    if (!pr->scale || !pr->var_init || !(pr->df > 0.0)) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: scale / var_init / df missing");
    int64_t p = 0;
    unsigned member_mask = 0;
    for (int b = 0; b < k; ++b) {
        const int s = pr->set_id[b];
        if (s < 0 || s >= NGP_MAX_SETS || !h->sets[s].have_geno) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: upload genotypes of set %d first", s);
        if ((member_mask >> s) & 1u) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: set %d listed twice", s);
        member_mask |= 1u << s;
        if (b == 0) p = h->sets[s].p;
        else if (h->sets[s].p != p) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: member sets must have the same number of loci (mme.jl:453)");
    }
    for (int s = 0; s < h->n_sets; ++s)
        if (h->sets[s].have_geno && !h->sets[s].joint_copy && !((member_mask >> s) & 1u))
            return fail(h, NGP_EUNSUPPORTED, "ngp_set_joint_prior: every marker set of the handle must be a member of the tuple (set %d is not)", s);
    int64_t R = 1;
    std::vector<int64_t> ro;
    if (pr->region_off && pr->n_regions >= 1) {
        R = pr->n_regions;
        if (pr->region_off[0] != 0 || pr->region_off[R] != p) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: region offsets must start at 0 and end at p");
        for (int64_t r = 0; r < R; ++r)
            if (pr->region_off[r + 1] <= pr->region_off[r]) return fail(h, NGP_EINVAL, "ngp_set_joint_prior: empty or unordered region %lld", (long long)r);
        ro.assign(pr->region_off, pr->region_off + R + 1);
    } else ro = {0, p};
    CU(cudaSetDevice(h->device));
    for (int b = 0; b < h->joint.k; ++b) h->sets[h->joint.set[b]].joint_member = false;
    if (h->joint.blocked_set >= 0) free_set(h->sets[h->joint.blocked_set]);
    free_joint(h->joint);
    JointHost& J = h->joint;
    J.k = k; J.p = p; J.n_regions = R; J.df = pr->df;
    for (int b = 0; b < k; ++b) J.set[b] = pr->set_id[b];
    for (int i = 0; i < k * k; ++i) J.scale[i] = pr->scale[i];
    CU(dalloc(&J.region_off, (size_t)R + 1));
    CU(cpy(h, J.region_off, ro.data(), sizeof(int64_t) * (R + 1), cudaMemcpyHostToDevice));
    std::vector<double> vb((size_t)R * k * k);
    for (int64_t r = 0; r < R; ++r) for (int i = 0; i < k * k; ++i) vb[(size_t)r * k * k + i] = pr->var_init[i];      // mme.jl:516
    CU(dalloc(&J.varBeta, vb.size()));
    CU(cpy(h, J.varBeta, vb.data(), sizeof(double) * vb.size(), cudaMemcpyHostToDevice));
    CU(dalloc(&J.mtm, (size_t)p * k * k));
    JointGeno G{};
    for (int b = 0; b < k; ++b) { G.geno[b] = h->sets[J.set[b]].geno; G.colsum[b] = h->sets[J.set[b]].colsum; }
    joint_mtm_kernel<<<(unsigned)((p + 7) / 8), 256, 0, h->stream>>>(G, k, h->Tw, h->R, h->B, h->sets[J.set[0]].p_pad / h->B, h->n, p, J.mtm);
    CU(cudaGetLastError());
    for (int b = 0; b < k; ++b) {
        SetHost& S = h->sets[J.set[b]];
        S.method = NGP_BAYESPR; S.est_pi = 0; S.nvar = 0; S.n_regions = R; S.df = pr->df; S.scale = 0.0;
        CU(dalloc(&S.sum_beta, S.p_pad)); CU(dalloc(&S.sum_beta2, S.p_pad)); CU(dalloc(&S.sum_delta, S.p_pad));
        CU(cudaMemsetAsync(S.sum_beta, 0, sizeof(double) * S.p_pad, h->stream));
        CU(cudaMemsetAsync(S.sum_beta2, 0, sizeof(double) * S.p_pad, h->stream));
        CU(cudaMemsetAsync(S.sum_delta, 0, sizeof(double) * S.p_pad, h->stream));
        CU(cudaMemsetAsync(S.beta, 0, sizeof(double) * S.p_pad, h->stream));                                      // mme.jl:458
        S.have_prior = true; S.joint_member = true;
    }
    CU(cudaStreamSynchronize(h->stream));
    J.active = true;
    h->sets_dirty = true;
    // Blocked path (k = 2 or 4, dividing the block size, a free set slot, one GPU): an INTERLEAVED copy of the members' codes, column
    // j*k + b = breed b of locus j, is swept by the look-ahead kernel with the joint k x k draw in the chain warp (ngp_sweep.cuh: joint_step).
    int slot = -1;
    for (int s = 0; s < NGP_MAX_SETS; ++s) if (!h->sets[s].have_geno) { slot = s; break; }
    // (panels of more than 512 rows with blocks of 32 / 64 have no tuple instantiation of the blocked sweep: the per-locus kernel serves them)
    if (slot >= 0 && (k == 2 || k == 4) && h->shard_world == 1 && p * k <= 0x7fffffffLL && !(h->R > 4 * kUpdThreads && h->B != 16)) {
        // Every effect of a tuple changes in every sweep, so the look-ahead buys nothing but cross-Gram corrections (B changed columns per
        // block and distance): a short look-ahead served entirely from the block records is the fast geometry (C1 / C4 measurements).
        // Tw, R and B — the tile layout of the sets already uploaded — stay as they are; only the kernel's rings are re-sized.
        if (!h->cfg_lookahead && !h->cfg_near) {
            const int sv_l = h->cfg_lookahead, sv_n = h->cfg_near, sv_r = h->cfg_refetch, sv_b = h->cfg_block;
            h->cfg_lookahead = 3; h->cfg_near = 3; h->cfg_refetch = 0; h->cfg_block = h->B;
            const int Tw0 = h->Tw, R0 = h->R;
            int rcg = choose_geometry(h, h->n, 0);
            h->cfg_lookahead = sv_l; h->cfg_near = sv_n; h->cfg_refetch = sv_r; h->cfg_block = sv_b;
            if (rcg) return rcg;
            if (h->Tw != Tw0 || h->R != R0) return fail(h, NGP_EINVAL, "internal: the tile layout changed while re-sizing the rings");
            h->ready_kfn = nullptr;
        }
        int rc = begin_upload(h, slot, h->n, p * k, NGP_STORE_I8);
        if (rc) return rc;
        SetHost& T = h->sets[slot];
        int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(p, (int64_t)((256u << 20) / (size_t)h->n)));
        chunk = std::min<int64_t>(chunk, 65535);
        int8_t* stage = nullptr;
        int* derr = nullptr;
        CU(cudaMalloc((void**)&stage, (size_t)h->n * chunk));
        CU(cudaMalloc((void**)&derr, sizeof(int)));
        CU(cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
        for (int b = 0; b < k; ++b) {
            const SetHost& Mb = h->sets[J.set[b]];
            for (int64_t c0 = 0; c0 < p; c0 += chunk) {
                const int64_t nc = std::min(chunk, p - c0);
                dim3 grid((unsigned)((h->n + 255) / 256), (unsigned)nc);
                unpack_kernel<<<grid, 256, 0, h->stream>>>(Mb.geno, h->n, h->R, h->B, Mb.p_pad / h->B, c0, nc, stage, h->store2);
                CU(cudaGetLastError());
                pack_kernel<NGP_GENO_I8><<<(unsigned)nc, 256, 0, h->stream>>>(stage, h->n, h->n, c0, h->R, h->B, T.p_pad / h->B, T.geno, T.colsum, T.colsumsq, derr, (int64_t)k, (int64_t)b);
                CU(cudaGetLastError());
            }
        }
        CU(cudaStreamSynchronize(h->stream));
        cudaFree(stage); cudaFree(derr);
        rc = finish_upload(h, T);
        if (rc) return rc;
        T.method = 4; T.est_pi = 0; T.nvar = 0; T.n_regions = R; T.df = pr->df; T.scale = 0.0; T.group_k = k; T.stream_set = J.set[0]; T.joint_copy = true;
        std::vector<int64_t> rok(ro);
        for (auto& v : rok) v *= k;                                   // locus offsets -> column offsets of the interleaved copy
        CU(dalloc(&T.region_off, (size_t)R + 1));
        CU(cpy(h, T.region_off, rok.data(), sizeof(int64_t) * (R + 1), cudaMemcpyHostToDevice));
        if (R > 1) {
            CU(dalloc(&T.region_of, T.p_pad));
            CU(zero(h, T.region_of, 0, sizeof(int32_t) * T.p_pad));
            region_of_kernel<<<(unsigned)std::min<int64_t>(R, 4096), 128, 0, h->stream>>>(T.region_off, R, T.region_of);
            CU(cudaGetLastError());
        }
        CU(dalloc(&T.jinvB, (size_t)R * k * k));
        CU(dalloc(&T.sum_beta, T.p_pad)); CU(dalloc(&T.sum_beta2, T.p_pad)); CU(dalloc(&T.sum_delta, T.p_pad));
        CU(cudaMemsetAsync(T.sum_beta, 0, sizeof(double) * T.p_pad, h->stream));
        CU(cudaMemsetAsync(T.sum_beta2, 0, sizeof(double) * T.p_pad, h->stream));
        CU(cudaMemsetAsync(T.sum_delta, 0, sizeof(double) * T.p_pad, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        T.have_prior = true;
        J.blocked_set = slot; J.state_in_copy = false;
        h->sets_dirty = true;
    }
    return NGP_OK;
}

int ngp_set_joint_replay(ngp_handle* h, int32_t n_iter, const double* z, const double* iw_chi2, const double* iw_z)
{
    if (!h) return NGP_EINVAL;
    JointHost& J = h->joint;
    if (!J.active) return fail(h, NGP_EINVAL, "ngp_set_joint_replay: no tuple of marker sets");
    if (n_iter <= 0 || !z || !iw_chi2 || !iw_z) return fail(h, NGP_EINVAL, "ngp_set_joint_replay: n_iter must be > 0 and the logs non-NULL");
    CU(cudaSetDevice(h->device));
    const size_t k = (size_t)J.k, nz = (size_t)n_iter * J.p * k, nc = (size_t)n_iter * J.n_regions * k, nl = nc * k;
    CU(dalloc(&J.rp_z, nz)); CU(cpy(h, J.rp_z, z, sizeof(double) * nz, cudaMemcpyHostToDevice));
    CU(dalloc(&J.rp_iw_chi2, nc)); CU(cpy(h, J.rp_iw_chi2, iw_chi2, sizeof(double) * nc, cudaMemcpyHostToDevice));
    CU(dalloc(&J.rp_iw_z, nl)); CU(cpy(h, J.rp_iw_z, iw_z, sizeof(double) * nl, cudaMemcpyHostToDevice));
    J.replay_iters = n_iter;
    h->sets_dirty = true;
    return NGP_OK;
}

int ngp_get_joint_state(ngp_handle* h, double* beta, double* varBeta)
{
    if (!h) return NGP_EINVAL;
    JointHost& J = h->joint;
    if (!J.active) return fail(h, NGP_EINVAL, "ngp_get_joint_state: no tuple of marker sets");
    CU(cudaSetDevice(h->device));
    if (beta) {
        int rc = tuple_sync_state(h, false);          // effects back into the member sets if the blocked kernel ran last
        if (rc) return rc;
        for (int b = 0; b < J.k; ++b) CU(cpy(h, beta + (size_t)b * J.p, h->sets[J.set[b]].beta, sizeof(double) * J.p, cudaMemcpyDeviceToHost));
    }
    if (varBeta) CU(cpy(h, varBeta, J.varBeta, sizeof(double) * (size_t)J.n_regions * J.k * J.k, cudaMemcpyDeviceToHost));
    return NGP_OK;
}

int ngp_joint_sweep(ngp_handle* h, double* ycorr, double varE, double* beta, double* varBeta)
{
    if (!h) return NGP_EINVAL;
    JointHost& J = h->joint;
    if (!J.active) return fail(h, NGP_EINVAL, "ngp_joint_sweep: no tuple of marker sets");
    if (!ycorr || !(varE > 0.0)) return fail(h, NGP_EINVAL, "ngp_joint_sweep: ycorr must be non-NULL and varE > 0");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->e, ycorr, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
    if (beta) for (int b = 0; b < J.k; ++b) CU(cudaMemcpyAsync(h->sets[J.set[b]].beta, beta + (size_t)b * J.p, sizeof(double) * J.p, cudaMemcpyHostToDevice, h->stream));
    if (varBeta) CU(cudaMemcpyAsync(J.varBeta, varBeta, sizeof(double) * (size_t)J.n_regions * J.k * J.k, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_y = true;
    if (beta) J.state_in_copy = false;                 // the caller's effects are in the member sets now
    int rc = tuple_launch(h, 1, 0, 0, varE, 0);
    if (rc) return rc;
    CU(cpy(h, ycorr, h->e, sizeof(double) * h->n, cudaMemcpyDeviceToHost));
    return ngp_get_joint_state(h, beta, varBeta);
}

int ngp_get_class_pi(ngp_handle* h, int set_id, double* piHat)
{
    if (!h || !piHat || set_id < 0 || set_id >= NGP_MAX_SETS || h->sets[set_id].method != NGP_BAYESR || !h->sets[set_id].pi_class)
        return fail(h, NGP_EINVAL, "ngp_get_class_pi: set %d is not a BayesR set (RC sets: ngp_get_rc_state)", set_id);
    CU(cudaSetDevice(h->device));
    CU(cpy(h, piHat, h->sets[set_id].pi_class, sizeof(double) * h->sets[set_id].n_class, cudaMemcpyDeviceToHost));
    return NGP_OK;
}

int ngp_set_rc_prior(ngp_handle* h, int set_id, const ngp_rc_prior* pr)
{
    if (!h || !pr) return fail(h, NGP_EINVAL, "ngp_set_rc_prior: NULL argument");
    if (set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].have_geno) return fail(h, NGP_EINVAL, "ngp_set_rc_prior: upload genotypes of set %d first", set_id);
    if (pr->n_class < 1 || pr->n_class > kMaxClass || pr->n_annot < 1 || pr->n_annot * pr->n_class > 32 || !pr->v_class || !pr->pi_class || !pr->annot)
        return fail(h, NGP_EINVAL, "ngp_set_rc_prior: 1..%d classes, n_annot * n_class <= 32, v_class / pi_class / annot non-NULL", kMaxClass);
    for (int v = 0; v < pr->n_class; ++v)
        if (!(pr->v_class[v] >= 0.0) || !(pr->pi_class[v] > 0.0)) return fail(h, NGP_EINVAL, "ngp_set_rc_prior: class %d: scale must be >= 0 and proportion > 0", v);
    SetHost& S0 = h->sets[set_id];
    const int64_t p = S0.p;
    const int nA = pr->n_annot, nc = pr->n_class;
    for (int64_t j = 0; j < p; ++j) {
        int64_t rs = 0;
        for (int a = 0; a < nA; ++a) { if (pr->annot[j * nA + a] < 0) return fail(h, NGP_EINVAL, "ngp_set_rc_prior: negative annotation count"); rs += pr->annot[j * nA + a]; }
        if (rs == 0) return fail(h, NGP_EDATA, "ngp_set_rc_prior: locus %lld has no annotation (annotProb would be NaN, mme.jl:395-399)", (long long)j);
    }
    // the BayesR plumbing (classes, variances, state arrays), then the annotation arrays
    ngp_prior q{};
    q.method = NGP_BAYESR; q.est_pi = pr->est_pi; q.df = pr->df; q.scale = pr->scale; q.var_init = pr->var_init;
    q.n_class = pr->n_class; q.v_class = pr->v_class; q.pi_class = pr->pi_class;
    int rc = ngp_set_prior(h, set_id, &q);
    if (rc) return rc;
    SetHost& S = h->sets[set_id];
    S.method = pr->plus ? NGP_BAYESRCPLUS : NGP_BAYESRCPI;
    S.n_annot = nA;
    S.nvar = nA;
    CU(dalloc(&S.varBeta, S.nvar));
    fill_kernel<<<1, 256, 0, h->stream>>>(S.varBeta, S.nvar, pr->var_init);                                  // mme.jl:516
    CU(cudaGetLastError());
    std::vector<double> pc((size_t)2 * nA * nc), ap((size_t)2 * p * nA);
    for (int a = 0; a < nA; ++a)
        for (int v = 0; v < nc; ++v) { pc[(size_t)a * nc + v] = pr->pi_class[v]; pc[(size_t)nA * nc + a * nc + v] = log(pr->pi_class[v]); }      // mme.jl:390-392
    for (int64_t j = 0; j < p; ++j) {
        double rs = 0.0;
        for (int a = 0; a < nA; ++a) rs += (double)pr->annot[j * nA + a];
        for (int a = 0; a < nA; ++a) ap[(size_t)j * nA + a] = ap[(size_t)(p + j) * nA + a] = (double)pr->annot[j * nA + a] / rs;                 // mme.jl:395
    }
    CU(dalloc(&S.pi_class, pc.size()));
    CU(cpy(h, S.pi_class, pc.data(), sizeof(double) * pc.size(), cudaMemcpyHostToDevice));
    CU(dalloc(&S.annot, (size_t)p * nA));
    CU(cpy(h, S.annot, pr->annot, sizeof(int32_t) * (size_t)p * nA, cudaMemcpyHostToDevice));
    CU(dalloc(&S.annot_prob, ap.size()));
    CU(cpy(h, S.annot_prob, ap.data(), sizeof(double) * ap.size(), cudaMemcpyHostToDevice));
    CU(dalloc(&S.annot_cat, (size_t)S.p_pad));
    CU(zero(h, S.annot_cat, 0, sizeof(int32_t) * S.p_pad));                                                  // mme.jl:403
    CU(cudaStreamSynchronize(h->stream));
    h->sets_dirty = true;
    return NGP_OK;
}

int ngp_get_rc_state(ngp_handle* h, int set_id, int64_t* annot_cat, double* annot_prob, double* pi_hat)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || (h->sets[set_id].method != NGP_BAYESRCPI && h->sets[set_id].method != NGP_BAYESRCPLUS))
        return fail(h, NGP_EINVAL, "ngp_get_rc_state: set %d is not a BayesRCpi / BayesRCplus set", set_id);
    SetHost& S = h->sets[set_id];
    CU(cudaSetDevice(h->device));
    if (annot_cat) {
        std::vector<int32_t> t((size_t)S.p);
        CU(cpy(h, t.data(), S.annot_cat, sizeof(int32_t) * S.p, cudaMemcpyDeviceToHost));
        for (int64_t j = 0; j < S.p; ++j) annot_cat[j] = t[(size_t)j];
    }
    if (annot_prob) {
        Scalars sc;
        CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
        // iteration it reads buffer (it & 1) and writes buffer ((it + 1) & 1): after sc.iter iterations the current values are in buffer ((sc.iter + 1) & 1)
        const size_t off = (size_t)((sc.iter + 1) & 1) * S.p * S.n_annot;
        CU(cpy(h, annot_prob, S.annot_prob + off, sizeof(double) * (size_t)S.p * S.n_annot, cudaMemcpyDeviceToHost));
    }
    if (pi_hat) CU(cpy(h, pi_hat, S.pi_class, sizeof(double) * (size_t)S.n_annot * S.n_class, cudaMemcpyDeviceToHost));
    return NGP_OK;
}

int ngp_set_rc_replay(ngp_handle* h, int set_id, int32_t n_iter, const double* u_annot, const double* dirp, const double* u,
                      const double* z, const double* chi2_b, const double* dir_pi)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || (h->sets[set_id].method != NGP_BAYESRCPI && h->sets[set_id].method != NGP_BAYESRCPLUS))
        return fail(h, NGP_EINVAL, "ngp_set_rc_replay: set %d is not a BayesRCpi / BayesRCplus set", set_id);
    SetHost& S = h->sets[set_id];
    if (!h->replay || n_iter != h->replay_iters) return fail(h, NGP_EINVAL, "ngp_set_rc_replay: call ngp_set_replay first, with the same number of iterations");
    const bool plus = S.method == NGP_BAYESRCPLUS;
    if (!u || !z || !chi2_b || (!plus && (!u_annot || !dirp)) || (S.est_pi && !dir_pi)) return fail(h, NGP_EINVAL, "ngp_set_rc_replay: missing arrays");
    CU(cudaSetDevice(h->device));
    const size_t ni = (size_t)n_iter, p = (size_t)S.p, nA = (size_t)S.n_annot, nc = (size_t)S.n_class;
    const size_t nu = plus ? ni * p * nA * nc : ni * p * nc, nz = plus ? ni * p * nA : ni * p;
    CU(dalloc(&S.rp_u, nu)); CU(cpy(h, S.rp_u, u, sizeof(double) * nu, cudaMemcpyHostToDevice));
    CU(dalloc(&S.rp_z, nz)); CU(cpy(h, S.rp_z, z, sizeof(double) * nz, cudaMemcpyHostToDevice));
    CU(dalloc(&S.rp_chi2b, ni * nA)); CU(cpy(h, S.rp_chi2b, chi2_b, sizeof(double) * ni * nA, cudaMemcpyHostToDevice));
    if (!plus) {
        CU(dalloc(&S.rp_u_annot, ni * p)); CU(cpy(h, S.rp_u_annot, u_annot, sizeof(double) * ni * p, cudaMemcpyHostToDevice));
        CU(dalloc(&S.rp_dirp, ni * p * nA)); CU(cpy(h, S.rp_dirp, dirp, sizeof(double) * ni * p * nA, cudaMemcpyHostToDevice));
    }
    if (S.est_pi) { CU(dalloc(&S.rp_betapi, ni * nA * nc)); CU(cpy(h, S.rp_betapi, dir_pi, sizeof(double) * ni * nA * nc, cudaMemcpyHostToDevice)); }
    h->sets_dirty = true;
    return NGP_OK;
}

int ngp_get_state(ngp_handle* h, ngp_state* out)
{
    if (!h || !out) return fail(h, NGP_EINVAL, "ngp_get_state: NULL argument");
    CU(cudaSetDevice(h->device));
    Scalars sc;
    CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
    out->n = h->n; out->n_sets = h->n_sets; out->mu = sc.mu; out->varE = sc.varE; out->iter = sc.iter;
    if (out->e && h->e) CU(cpy(h, out->e, h->e, sizeof(double) * h->n, cudaMemcpyDeviceToHost));
    for (int s = 0; s < h->n_sets; ++s) {
        if (!h->sets[s].have_geno || !h->sets[s].have_prior) continue;
        int rc = pull_set_state(h, s, out->beta[s], out->delta[s], out->varBeta[s], h->sets[s].method == NGP_BAYESR ? nullptr : out->pi[s]);   // class proportions: ngp_get_class_pi
        if (rc) return rc;
    }
    return NGP_OK;
}

int ngp_set_state(ngp_handle* h, const ngp_state* in)
{
    if (!h || !in) return fail(h, NGP_EINVAL, "ngp_set_state: NULL argument");
    if (h->Tw == 0) return fail(h, NGP_EINVAL, "ngp_set_state: upload a marker set first");
    CU(cudaSetDevice(h->device));
    Scalars sc;
    CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
    sc.mu = in->mu; sc.varE = in->varE; sc.iter = in->iter;
    CU(cpy(h, h->sc, &sc, sizeof sc, cudaMemcpyHostToDevice));
    if (in->e) { CU(cpy(h, h->e, in->e, sizeof(double) * h->n, cudaMemcpyHostToDevice)); h->have_y = true; }
    for (int s = 0; s < h->n_sets; ++s) {
        if (!h->sets[s].have_geno || !h->sets[s].have_prior) continue;
        int rc = push_set_state(h, s, in->beta[s], in->delta[s], in->varBeta[s], h->sets[s].method == NGP_BAYESR ? nullptr : in->pi[s]);
        if (rc) return rc;
    }
    return NGP_OK;
}

int ngp_reset_posterior(ngp_handle* h)
{
    if (!h) return NGP_EINVAL;
    CU(cudaSetDevice(h->device));
    for (int s = 0; s < h->n_sets; ++s) {
        SetHost& S = h->sets[s];
        if (!S.sum_beta) continue;
        CU(zero(h, S.sum_beta, 0, sizeof(double) * S.p_pad));
        CU(zero(h, S.sum_beta2, 0, sizeof(double) * S.p_pad));
        CU(zero(h, S.sum_delta, 0, sizeof(double) * S.p_pad));
    }
    Scalars sc;
    CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
    sc.n_post = 0;
    CU(cpy(h, h->sc, &sc, sizeof sc, cudaMemcpyHostToDevice));
    return NGP_OK;
}

int ngp_get_posterior(ngp_handle* h, int set_id, int64_t* n_samples, double* sum_beta, double* sum_beta2, double* sum_delta)
{
    if (!h || set_id < 0 || set_id >= NGP_MAX_SETS || !h->sets[set_id].sum_beta) return fail(h, NGP_EINVAL, "ngp_get_posterior: no such set");
    CU(cudaSetDevice(h->device));
    SetHost& S = h->sets[set_id];
    Scalars sc;
    CU(cpy(h, &sc, h->sc, sizeof sc, cudaMemcpyDeviceToHost));
    if (n_samples) *n_samples = sc.n_post;
    if (S.joint_member && h->joint.blocked_set >= 0 && h->cfg_kernel == NGP_KERNEL_BLOCKED) {
        // the blocked tuple sweep accumulates in the interleaved copy: column j*k + b
        const JointHost& J = h->joint;
        const SetHost& T = h->sets[J.blocked_set];
        int b = 0;
        for (int x = 0; x < J.k; ++x) if (J.set[x] == set_id) b = x;
        std::vector<double> tmp((size_t)T.p);
        double* outs[3] = {sum_beta, sum_beta2, sum_delta};
        const double* srcs[3] = {T.sum_beta, T.sum_beta2, T.sum_delta};
        for (int q = 0; q < 3; ++q)
            if (outs[q]) {
                CU(cpy(h, tmp.data(), srcs[q], sizeof(double) * T.p, cudaMemcpyDeviceToHost));
                for (int64_t j = 0; j < J.p; ++j) outs[q][j] = tmp[(size_t)(j * J.k + b)];
            }
        return NGP_OK;
    }
    if (sum_beta) CU(cpy(h, sum_beta, S.sum_beta, sizeof(double) * S.p, cudaMemcpyDeviceToHost));
    if (sum_beta2) CU(cpy(h, sum_beta2, S.sum_beta2, sizeof(double) * S.p, cudaMemcpyDeviceToHost));
    if (sum_delta) CU(cpy(h, sum_delta, S.sum_delta, sizeof(double) * S.p, cudaMemcpyDeviceToHost));
    return NGP_OK;
}

int ngp_get_timing(ngp_handle* h, ngp_timing* out)
{
    if (!h || !out) return fail(h, NGP_EINVAL, "ngp_get_timing: NULL argument");
    memset(out, 0, sizeof *out);
    out->launches = h->launches; out->ctas = h->Tw ? h->Tw + 1 : 0; out->threads = kThreads; out->block = h->B; out->rows_per_cta = h->R;
    out->smem_bytes = h->L.total; out->lookahead = h->D; out->near_depth = h->DN; out->tile_stages = h->NT; out->record_stages = h->NR;
    out->kernel_variant = h->last_variant; out->refetch = h->refetch; out->storage_2bit = h->store2 > 0; out->sum_run_ms = h->sum_run_ms;
    if (h->timed) {
        float ms = 0.f;
        CU(cudaEventSynchronize(h->ev1));
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        out->last_run_ms = ms;
    }
    return NGP_OK;
}

int ngp_get_profile(ngp_handle* h, int64_t* out, int32_t max_ctas)
{
    if (!h || !out || max_ctas <= 0) return fail(h, NGP_EINVAL, "ngp_get_profile: bad argument");
    CU(cudaSetDevice(h->device));
    const int nc = std::min<int>(max_ctas, h->Tw + 1);
    CU(cpy(h, out, h->sync->prof, sizeof(long long) * kProf * nc, cudaMemcpyDeviceToHost));
    return nc;
}

int ngp_get_trace(ngp_handle* h, int64_t* out, int32_t n)
{
    if (!h || !out || n <= 0) return fail(h, NGP_EINVAL, "ngp_get_trace: bad argument");
    CU(cudaSetDevice(h->device));
    const int m = std::min<int>(n, 2 * 2048);
    CU(cpy(h, out, h->sync->trace, sizeof(long long) * m, cudaMemcpyDeviceToHost));
    return m;
}

int ngp_debug_variates(ngp_handle* h, int set_id, uint32_t iter, int purpose, double df, int64_t n, double* out)
{
    if (!h || !out || n <= 0) return fail(h, NGP_EINVAL, "ngp_debug_variates: bad argument");
    CU(cudaSetDevice(h->device));
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, sizeof(double) * n));
    debug_variates_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>((uint32_t)(h->seed & 0xffffffffu), (uint32_t)(h->seed >> 32),
                                                                              h->chain, iter, (uint32_t)set_id, purpose, df, n, d);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    return NGP_OK;
}

}  // extern "C"
