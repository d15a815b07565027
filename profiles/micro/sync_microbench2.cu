// Microbenchmark 2 of the grid-wide "self-synchronising accumulator" reduction of ngp::gibbs_kernel.
// Explores what bounds the RED -> visible latency L and the sustainable round rate:
//   shards G : CTA c adds into shard (c % G) of every accumulator; the poller sums G shards (expects T/G arrivals each)
//   lag      : the round polled is r - lag (look-ahead depth of the sweep)
//   pollmode : 0 = every lane polls its own accumulators until complete
//              1 = poll ONE sentinel accumulator (staggered per CTA) until complete, then read the others (re-poll if late)
// Build: nvcc -arch=sm_100a -O3 -o sync_microbench2 sync_microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_add_u64(long long* addr, long long v)
{ asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p)
{ long long v; asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

constexpr int SLOTS = 8, MAXNV = 128, MAXG = 4, STRIDE = 32;   // 256 B between accumulators

__global__ void bench(long long* acc, int nv, int rounds, int lag, int G, int pollmode, long long* out, int work)
{
    const int T = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    __shared__ long long prev[SLOTS][MAXG][MAXNV];
    __shared__ double sink;
    for (int i = tid; i < SLOTS * MAXG * MAXNV; i += blockDim.x) (&prev[0][0][0])[i] = 0;
    __syncthreads();
    const int shard = c % G;
    // arrivals expected in shard g
    long long t0 = clock64(), tpoll = 0;
    double x = 1.0;
    for (int r = 0; r < rounds + lag; ++r) {
        for (int w = 0; w < work; ++w) x = fma(x, 1.0000001, 0.5);
        const int slot = r % SLOTS;
        if (r < rounds && tid >= 32 && tid < 32 + nv)
            red_add_u64(acc + (((long long)slot * MAXG + shard) * MAXNV + (tid - 32)) * STRIDE, (1LL << 8) + 1);
        const int rp = r - lag;
        if (warp == 0 && rp >= 0) {
            const int sp = rp % SLOTS;
            long long tt = clock64();
            if (pollmode == 1) {
                // sentinel: accumulator (c % nv) of shard (c % G)
                const int q = c % nv, g = c % G;
                const int exp_g = (T - g + G - 1) / G;
                const long long* a = acc + (((long long)sp * MAXG + g) * MAXNV + q) * STRIDE;
                long long cur;
                do { cur = ld_relaxed_s64(a); } while (((cur - prev[sp][g][q]) & 0xFF) != exp_g);
            }
            bool done;
            long long cur[MAXNV / 32 * MAXG];
            do {
                done = true;
                int k = 0;
                for (int g = 0; g < G; ++g) {
                    const int exp_g = (T - g + G - 1) / G;
                    for (int b = 0; b < nv / 32; ++b, ++k) {
                        const int q = b * 32 + lane;
                        cur[k] = ld_relaxed_s64(acc + (((long long)sp * MAXG + g) * MAXNV + q) * STRIDE);
                        done = done && (((cur[k] - prev[sp][g][q]) & 0xFF) == exp_g);
                    }
                }
            } while (!__all_sync(0xffffffffu, done));
            int k = 0;
            for (int g = 0; g < G; ++g)
                for (int b = 0; b < nv / 32; ++b, ++k) prev[sp][g][b * 32 + lane] = cur[k];
            tpoll += clock64() - tt;
        }
        __syncthreads();
    }
    if (tid == 0) { out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = tpoll; sink = x; }
}

int main()
{
    cudaSetDevice(0);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int T = pr.multiProcessorCount;
    long long *acc, *out;
    const size_t bytes = (size_t)SLOTS * MAXG * MAXNV * STRIDE * 8;
    cudaMalloc(&acc, bytes); cudaMalloc(&out, T * 16);
    const int rounds = 2000;
    printf("T=%d rounds=%d\n", T, rounds);
    for (int pollmode = 0; pollmode < 2; ++pollmode)
        for (int nv : {32, 64, 128})
            for (int G : {1, 2, 4})
                for (int lag : {0, 1, 2, 3}) {
                    if (nv / 32 * G > 16) continue;
                    cudaMemset(acc, 0, bytes);
                    int rr = rounds, work = 0, nvv = nv, gg = G, ll = lag, pm = pollmode;
                    void* args[] = {&acc, &nvv, &rr, &ll, &gg, &pm, &out, &work};
                    cudaLaunchCooperativeKernel((void*)bench, dim3(T), dim3(32 + MAXNV), args, 0, 0);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[2 * 256]; cudaMemcpy(h, out, T * 16, cudaMemcpyDeviceToHost);
                    double tot = 0, poll = 0; for (int i = 0; i < T; ++i) { tot += h[2 * i]; poll += h[2 * i + 1]; }
                    printf("pollmode=%d nv=%3d G=%d lag=%d : %.0f cycles/round, poll %.0f  (%s)\n", pollmode, nv, G, lag,
                           tot / T / rounds, poll / T / rounds, cudaGetErrorString(e)); fflush(stdout);
                }
    return 0;
}
