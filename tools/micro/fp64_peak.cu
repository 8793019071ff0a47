// fp64 peak of this GPU: vector DFMA and tensor DMMA (mma.sync.m8n8k4.f64).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu  (the binary is git-ignored; results: profiles/r02u_fp64_peak.txt)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters, double x)
{
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, 1e-9);
    double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters, double x)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double* d; cudaMalloc(&d, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int tpb : {128, 256, 512, 1024}) {
        const int grid = 148 * (2048 / tpb), iters = 4096;
        for (int which = 0; which < 2; ++which) {
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_dfma<<<grid, tpb>>>(d, iters, 0.999); else k_dmma<<<grid, tpb>>>(d, iters, 0.999);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double flop = which == 0 ? 2.0 * 16 * iters * (double)grid * tpb : 2.0 * 256 * 8 * iters * (double)grid * (tpb / 32);
            printf("%s tpb %4d: %.3f ms  %.2f TFLOP/s\n", which == 0 ? "DFMA" : "DMMA", tpb, best, flop / best * 1e-9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
