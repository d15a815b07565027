// Dependent-issue latency of the instructions on the chain warp's critical path (one warp, one SM).
// Build: nvcc -arch=sm_100a -O3 -o lat_microbench lat_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double a, double b, int* ibuf)
{
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 7 + 1) & 1023;
    __syncwarp();
    double x = a; long long t0, t1; int q = threadIdx.x; unsigned m = 0; float f = (float)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    t1 = clock64(); cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * b;
    t1 = clock64(); cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + b;
    t1 = clock64(); cyc[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { q = __double2int_rn(x) & 1023; x = (double)q + 0.25; }      // F2I + I2F + DADD
    t1 = clock64(); cyc[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) q = sm[q];                                                  // LDS dependent
    t1 = clock64(); cyc[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) q = __shfl_sync(0xffffffffu, q, (q + 1) & 31);              // SHFL dependent
    t1 = clock64(); cyc[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { m = __ballot_sync(0xffffffffu, (q + (int)m) & 1); q += __ffs(m); }   // VOTE + FLO + add
    t1 = clock64(); cyc[6] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { x = (x > a) ? x * b : x + b; }                            // DSETP + select + op
    t1 = clock64(); cyc[7] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
    t1 = clock64(); cyc[8] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { f = (float)x; x = (double)f * b; }                        // F2F down + F2F up + DMUL
    t1 = clock64(); cyc[9] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { x = (double)sm[q] + x; q = (q + 1) & 1023; }              // LDS + I2F + DADD (LDS not dependent on x)
    t1 = clock64(); cyc[10] = t1 - t0;
    out[threadIdx.x] = x + q + m + f; ibuf[threadIdx.x] = q;
}
int main()
{
    double* out; long long* cyc; int* ib;
    cudaMalloc(&out, 256); cudaMalloc(&cyc, 128); cudaMalloc(&ib, 128);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, 1.000001, 0.999999, ib);
    long long h[16]; cudaMemcpy(h, cyc, 88, cudaMemcpyDeviceToHost);
    const char* nm[] = {"DFMA", "DMUL", "DADD", "F2I+I2F+DADD", "LDS(dep)", "SHFL(dep)", "VOTE+FFS+IADD", "DSETP+sel+D-op", "FFMA", "F2F.dn+F2F.up+DMUL", "I2F+DADD (LDS indep)"};
    for (int i = 0; i < 11; ++i) printf("%-24s %.1f cycles per iteration\n", nm[i], (double)h[i] / N);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
