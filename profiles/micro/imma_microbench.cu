// Does legacy INT8 mma.sync (IMMA.16832.S8.S8) run at a useful rate on B200?  One warp / 4 warps per SM, dependent and
// independent accumulator chains.  Build: nvcc -arch=sm_100a -O3 -o imma_microbench imma_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma(int (&c)[4], const uint4& a, const uint2& b)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}
template <int CH>
__global__ void bench(int iters, int* out, long long* cyc)
{
    int c[CH][4];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
    uint4 a = make_uint4(0x01020001u + threadIdx.x, 0x02010002u, 0x00010201u, 0x01000102u);
    uint2 b = make_uint2(0x7f80ff01u ^ threadIdx.x, 0x01fe8003u);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) imma(c[i], a, b);
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    int* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int warps : {1, 4, 8}) {
        long long h[148];
        bench<1><<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps/SM=%d chains=1 : %.1f cycles per IMMA per warp (dependent)\n", warps, (double)h[0] / iters);
        bench<4><<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps/SM=%d chains=4 : %.1f cycles per IMMA per warp (4 independent chains)  %s\n", warps, (double)h[0] / iters / 4, cudaGetErrorString(cudaGetLastError()));
    }
    // correctness: one IMMA with known operands
    return 0;
}
