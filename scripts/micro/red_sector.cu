// micro-benchmark: throughput of 64-bit global REDs when the 32 lanes of a warp instruction fall into 32 / 16 / 8 distinct
// 32-byte sectors (4 accumulators per sector).  Question behind it: is K3 (column -> link volume scatter) bound per lane
// operation or per sector request?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_sector red_sector.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void red_kernel(unsigned long long* acc, const int* idx, int n_per_thread, int total)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < n_per_thread; ++k) {
        const int i = (int)(((long long)k * gridDim.x * blockDim.x + t) % total);
        atomicAdd(&acc[idx[i]], 12345ULL + k);
    }
}
int main()
{
    const int L = 300000, total = 1 << 24;
    unsigned long long* acc; int* idx;
    cudaMalloc(&acc, sizeof(unsigned long long) * L); cudaMemset(acc, 0, sizeof(unsigned long long) * L);
    cudaMalloc(&idx, sizeof(int) * total);
    int* h = new int[total];
    for (int per_sector = 1; per_sector <= 4; per_sector *= 2) {
        // every aligned group of `per_sector` consecutive lanes shares one sector; groups are scattered over the array
        unsigned s = 12345u;
        for (int i = 0; i < total; i += per_sector) {
            s = s * 1664525u + 1013904223u;
            const int sector = (int)((s >> 8) % (L / 4));
            for (int j = 0; j < per_sector; ++j) h[i + j] = sector * 4 + j;
        }
        cudaMemcpy(idx, h, sizeof(int) * total, cudaMemcpyHostToDevice);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int blocks = 148 * 8, threads = 256, npt = 64;
        red_kernel<<<blocks, threads>>>(acc, idx, 4, total);
        cudaEventRecord(a);
        red_kernel<<<blocks, threads>>>(acc, idx, npt, total);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double ops = (double)blocks * threads * npt;
        printf("lanes per sector %d: %.3f ms, %.1f G lane-REDs/s, %.1f G sector requests/s\n", per_sector, ms, ops / ms * 1e-6, ops / per_sector / ms * 1e-6);
    }
    return 0;
}
