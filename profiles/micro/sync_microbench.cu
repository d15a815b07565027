// Microbenchmark of the grid-wide "self-synchronising accumulator" reduction used by ngp::gibbs_kernel:
// every CTA adds NV values with RED.ADD.64 (value<<8)+1, then one warp polls until all T arrivals are visible.
// Reports cycles per round for several layouts.  Build: nvcc -arch=sm_100a -O3 -o sync_microbench sync_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

__device__ __forceinline__ void red_add_u64(long long* addr, long long v)
{ asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p)
{ long long v; asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// mode 0: all CTAs RED then poll the same round (no look-ahead);  mode 1: poll round r-1 while RED round r (lag 1)
// mode 2: like 0 but only lane 0..NV/32-1 groups: single leader CTA polls and publishes a flag that others poll
__global__ void bench(long long* acc, int stride, int nv, int rounds, int mode, int slots, long long* out, int work)
{
    const int T = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ long long prev[4][64];
    __shared__ double sink;
    for (int i = tid; i < 4 * 64; i += blockDim.x) (&prev[0][0])[i] = 0;
    __syncthreads();
    long long t0 = clock64(), tpoll = 0;
    double x = 1.0;
    for (int r = 0; r < rounds; ++r) {
        // some independent work to emulate the dot phase
        for (int w = 0; w < work; ++w) x = fma(x, 1.0000001, 0.5);
        const int slot = r % slots;
        if (tid >= 32 && tid < 32 + nv) red_add_u64(acc + ((long long)slot * 65 + (tid - 32)) * stride, (1LL << 8) + 1);
        const int rp = (mode == 1) ? r - 1 : r;
        if (warp == 0 && rp >= 0) {
            const int sp = rp % slots;
            long long tt = clock64();
            for (int b = 0; b < nv / 32; ++b) {
                const int q = b * 32 + lane;
                long long cur;
                do { cur = ld_relaxed_s64(acc + ((long long)sp * 65 + q) * stride); } while (((cur - prev[sp][q]) & 0xFF) != T);
                prev[sp][q] = cur;
            }
            tpoll += clock64() - tt;
        }
        __syncthreads();
    }
    if (tid == 0) { out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = tpoll; sink = x; }
}

int main()
{
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int T = pr.multiProcessorCount;
    long long *acc, *out;
    cudaMalloc(&acc, 4 * 65 * 64 * 8 * 8); cudaMalloc(&out, T * 16);
    const int rounds = 2000;
    printf("T=%d rounds=%d\n", T, rounds);
    for (int mode = 0; mode < 2; ++mode)
        for (int nv : {32, 64})
            for (int stride : {1, 8, 32})
                for (int work : {0, 400}) {
                    cudaMemset(acc, 0, 4 * 65 * 64 * 8 * 8);
                    int slots = 4, rr = rounds;
                    void* args[] = {&acc, &stride, &nv, &rr, &mode, &slots, &out, &work};
                    cudaLaunchCooperativeKernel((void*)bench, dim3(T), dim3(288), args, 0, 0);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[2 * 256]; cudaMemcpy(h, out, T * 16, cudaMemcpyDeviceToHost);
                    double tot = 0, poll = 0; for (int i = 0; i < T; ++i) { tot += h[2 * i]; poll += h[2 * i + 1]; }
                    printf("mode=%d nv=%d stride=%3dB work=%3d : %.0f cycles/round, poll %.0f  (%s)\n", mode, nv, stride * 8, work,
                           tot / T / rounds, poll / T / rounds, cudaGetErrorString(e));
                }
    return 0;
}
