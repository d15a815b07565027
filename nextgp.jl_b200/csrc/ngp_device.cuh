// ngp_device.cuh — device-side data structures and small PTX wrappers of libngp.
//
// HBM layout of one marker set ("row-panelled column-major"):
//   the n individuals are cut into T row panels of R rows (R = 8*odd), panel t is
//   owned by CTA t of the persistent sweep kernel; inside a panel the markers are
//   consecutive and each marker's R codes are contiguous:
//       geno[(t * p_pad + j) * R + r]      row i = t*R + r, marker j
//   so a block of B consecutive markers of one panel is ONE contiguous B*R-byte
//   chunk that a single cp.async.bulk (TMA) brings into shared memory, and every
//   column is read from HBM exactly once per sweep.
//   A stored byte is 0xF0 | (code << 2): placed in bits 16..23 of the high word
//   of an fp64 (under 0x3F in bits 24..31) it IS the double 1 + code/4, so one
//   PRMT turns a code into an FMA operand (no I2F on the hot path).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "ngp_rng.cuh"

namespace ngp {

constexpr int kThreads = 320;          // sweep kernel: warp 0 = chain warp, warps 1..8 = workers, warp 9 = TMA producer
constexpr int kWarps = kThreads / 32;
constexpr int kWorkerWarps = 8;
constexpr int kProducerWarp = kWorkerWarps + 1;
constexpr int kProf = 16;              // cycle counters per CTA (ngp_get_profile)
constexpr int kMaxB = 64;              // markers per block (32 or 64)
constexpr int kNF = 10;                // per-marker constant fields
constexpr int kSlots = 4;              // reduction accumulator ring (look-ahead 1 needs >= 4)
constexpr int kAccStride = 32;         // int64 units between accumulators (256 B: distinct L2 slices)
constexpr int kMaxCtas = 160;          // < 256: the low byte of an accumulator counts arrivals
constexpr int kCntBits = 8;

// fields of the per-iteration marker constants, stored [block][field][B]
enum Field { F_A = 0, F_B = 1, F_T = 2, F_C = 3, F_QSZ = 4, F_D = 5, F_BOLD = 6, F_MEAN = 7, F_CS = 8, F_CHI = 9 };

struct SetDev {
    int64_t p, p_pad;
    int32_t method, est_pi;
    int64_t n_regions, nvar;
    double df, scale;
    const uint8_t* geno;       // [T][p_pad][R]
    uint8_t* blk;              // [p_pad/B] records {int32 gram[B][B]; double consts[kNF][B]}: raw Gram inside block k
                               //   (sum_i g_a g_b) + the per-iteration marker constants, one TMA bulk copy per block
    const int32_t* gramx;      // [p_pad/B][B][B] raw sum g_a g_b, a in block k-1, b in block k (block 0: zeros)
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
    double* sum_beta;          // posterior sums [p_pad]
    double* sum_beta2;
    double* sum_delta;
};

struct SyncArea {
    unsigned long long counter;                 // grid-barrier arrivals, monotonic within a launch
    unsigned long long pad0[15];
    long long acc[kSlots * (kMaxB + 1) * kAccStride];   // fixed-point reduction accumulators (monotonic)
    double part[kMaxCtas * 2];                  // phase-0 partials (e'e, sum e) per CTA
    int zero16[16];                             // runtime zeros (low words of the fp64 code operands, see dec_byte_z)
    long long prof[kMaxCtas * kProf];           // per-CTA cycle counters of the last launch: see ngp_get_profile
    int err;
};

struct Scalars {          // device-resident chain scalars
    double mu, varE;
    long long iter;       // iterations completed
    long long n_post;     // posterior samples accumulated
};

struct Params {
    int64_t n;
    int32_t T, R, B, n_sets, stages, kernel;
    double* e;                 // [T*R]
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
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void red_add_u64(long long* addr, long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void arrive_release(unsigned long long* c)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(c) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* c)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p)     // LDG.STRONG.GPU, no L1, no fence
{
    long long v;
    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// stored byte k of word w -> the double 1 + code/4
__device__ __forceinline__ double dec_byte(uint32_t w, int k)
{
    const uint32_t hi = __byte_perm(w, 0x3F000000u, 0x7044u | ((uint32_t)k << 8));
    return __hiloint2double((int)hi, 0);
}
// same with the low word taken from a register that holds a RUNTIME zero: the compiler cannot fold it, so it keeps
// the zero in the even register of the operand pair and the PRMT writes the odd one in place (no MOV per code)
__device__ __forceinline__ double dec_byte_z(uint32_t w, int k, int z)
{
    const uint32_t hi = __byte_perm(w, 0x3F000000u, 0x7044u | ((uint32_t)k << 8));
    return __hiloint2double((int)hi, z);
}
__host__ __device__ __forceinline__ int blk_bytes(int B) { return B * B * 4 + kNF * B * 8; }
__host__ __device__ __forceinline__ uint8_t enc_code(int g) { return (uint8_t)(0xF0 | (g << 2)); }
__host__ __device__ __forceinline__ int dec_code(uint8_t b) { return (b >> 2) & 3; }

}  // namespace ngp
