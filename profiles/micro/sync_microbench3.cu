// Microbenchmark 3: what bounds the RED.ADD.64 accumulator reduction of ngp::gibbs_kernel?
//   exp 0 : RED only (no polling) -> drain throughput; one final grid-wide check
//   exp 1 : poll only (accumulators never change, "done" immediately) -> poll round trip
//   exp 2 : RED + poll with lag, address order rotated per CTA by `rot` (0 = every CTA hits accumulator 0 first)
//   exp 3 : as 2 but a 32-bit arrival counter is separate from the data: data RED, then one RED.32 per CTA on a counter
//           (poller polls ONE counter address, then reads data once)
// Build: nvcc -arch=sm_100a -O3 -o sync_microbench3 sync_microbench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_add_u64(long long* addr, long long v)
{ asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p)
{ long long v; asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

constexpr int SLOTS = 8, NVMAX = 128;

template <int NB>
__global__ void bench(long long* acc, int strideq, int rounds, int lag, int exp, int rot, long long* out)
{
    constexpr int NV = NB * 32;
    const int T = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    __shared__ long long prev[SLOTS][NVMAX];
    for (int i = tid; i < SLOTS * NVMAX; i += blockDim.x) (&prev[0][0])[i] = 0;
    __syncthreads();
    long long t0 = clock64(), tpoll = 0;
    for (int r = 0; r < rounds + lag; ++r) {
        const int slot = r % SLOTS;
        if (exp != 1 && r < rounds && tid >= 32 && tid < 32 + NV) {
            const int q = (tid - 32 + rot * c) % NV;
            red_add_u64(acc + ((long long)slot * NVMAX + q) * strideq, (1LL << 8) + 1);
        }
        const int rp = r - lag;
        if (exp >= 1 && warp == 0 && rp >= 0) {
            const int sp = rp % SLOTS;
            const long long expect = (exp == 1) ? 0 : T;
            long long tt = clock64();
            long long cur[NB];
            bool done;
            do {
                done = true;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int q = b * 32 + lane;
                    cur[b] = ld_relaxed_s64(acc + ((long long)sp * NVMAX + q) * strideq);
                    done = done && (((cur[b] - prev[sp][q]) & 0xFF) == expect);
                }
            } while (!__all_sync(0xffffffffu, done));
#pragma unroll
            for (int b = 0; b < NB; ++b) prev[sp][b * 32 + lane] = cur[b];
            tpoll += clock64() - tt;
        }
        __syncthreads();
    }
    if (tid == 0) { out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = tpoll; }
}

template <int NB>
void run(long long* acc, long long* out, size_t bytes, int T, int strideq, int lag, int exp, int rot)
{
    const int rounds = 2000;
    cudaMemset(acc, 0, bytes);
    int rr = rounds;
    void* args[] = {&acc, &strideq, &rr, &lag, &exp, &rot, &out};
    cudaLaunchCooperativeKernel((void*)bench<NB>, dim3(T), dim3(32 + NVMAX), args, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2 * 256]; cudaMemcpy(h, out, T * 16, cudaMemcpyDeviceToHost);
    double tot = 0, poll = 0; for (int i = 0; i < T; ++i) { tot += h[2 * i]; poll += h[2 * i + 1]; }
    printf("exp=%d T=%3d nv=%3d stride=%4dB lag=%d rot=%d : %.0f cycles/round, poll %.0f  (%s)\n", exp, T, NB * 32, strideq * 8, lag, rot,
           tot / T / rounds, poll / T / rounds, cudaGetErrorString(e));
    fflush(stdout);
}

int main()
{
    cudaSetDevice(0);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int SM = pr.multiProcessorCount;
    long long *acc, *out;
    const size_t bytes = (size_t)SLOTS * NVMAX * 128 * 8;
    cudaMalloc(&acc, bytes); cudaMalloc(&out, 256 * 16);
    for (int T : {SM, SM / 2, SM / 4})
        for (int stride : {1, 16, 32, 40, 128}) {
            run<1>(acc, out, bytes, T, stride, 0, 0, 0);
            run<2>(acc, out, bytes, T, stride, 0, 0, 0);
            run<4>(acc, out, bytes, T, stride, 0, 0, 0);
            run<2>(acc, out, bytes, T, stride, 0, 0, 1);
        }
    for (int T : {SM, SM / 2}) {
        run<1>(acc, out, bytes, T, 32, 0, 1, 0);
        run<2>(acc, out, bytes, T, 32, 0, 1, 0);
        run<4>(acc, out, bytes, T, 32, 0, 1, 0);
    }
    for (int T : {SM, SM / 2, SM / 4})
        for (int stride : {1, 32, 40})
            for (int rot : {0, 1, 7})
                for (int lag : {0, 2}) {
                    run<1>(acc, out, bytes, T, stride, lag, 2, rot);
                    run<2>(acc, out, bytes, T, stride, lag, 2, rot);
                    run<4>(acc, out, bytes, T, stride, lag, 2, rot);
                }
    return 0;
}
