// Micro-benchmarks that drive the K1 design: dependent-access latency of the memory operations the SSSP kernel uses.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <numeric>
#include <algorithm>
#include <random>

__global__ void chase_ldcg(const unsigned long long* a, int steps, long long* out, unsigned long long* sink)
{
    unsigned long long i = threadIdx.x * 977 + blockIdx.x * 131071;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) i = __ldcg((const long long*)&a[i]);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / steps;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = i;
}
__global__ void chase_ldg128(const int4* a, int steps, long long* out, unsigned long long* sink)
{
    unsigned i = threadIdx.x * 977 + blockIdx.x * 131071;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) i = __ldg(&a[i]).x;
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / steps;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = i;
}
__global__ void chase_atom(unsigned long long* a, int steps, long long* out, unsigned long long* sink)
{
    unsigned long long i = threadIdx.x * 977 + blockIdx.x * 131071;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) i = atomicMin(&a[i], ~0ULL); // returns the stored index, changes nothing
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / steps;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = i;
}
// throughput: every thread issues `per` independent random accesses, then waits
template <int MODE>
__global__ void tput(unsigned long long* a, size_t n, int iters, unsigned long long* sink)
{
    unsigned long long x = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 12345;
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
        size_t i = (x >> 20) % n;
        if (MODE == 0) acc += __ldcg((const long long*)&a[i]);
        if (MODE == 1) acc += atomicMin(&a[i], ~0ULL);
        if (MODE == 2) atomicMin(&a[i], ~0ULL);
        if (MODE == 3) a[i] = x | (1ULL << 62);
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv)
{
    const size_t mb = argc > 1 ? atoi(argv[1]) : 32;
    const size_t n = mb * 1024 * 1024 / 8;
    std::vector<unsigned long long> h(n);
    std::iota(h.begin(), h.end(), 0ULL);
    std::mt19937_64 rng(1);
    std::shuffle(h.begin(), h.end(), rng); // random permutation: a[i] = next index (may have short cycles; fine)
    unsigned long long *d, *sink;
    long long* out;
    cudaMalloc(&d, n * 8); cudaMalloc(&sink, 148 * 64 * 1024 * 8); cudaMalloc(&out, 4096 * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    std::vector<int4> h4(n / 2);
    for (size_t i = 0; i < n / 2; ++i) h4[i] = make_int4((int)(h[i] % (n / 2)), 0, 0, 0);
    int4* d4; cudaMalloc(&d4, n / 2 * 16); cudaMemcpy(d4, h4.data(), n / 2 * 16, cudaMemcpyHostToDevice);
    long long ho[4096];
    auto report = [&](const char* name, int blocks, int threads) {
        cudaDeviceSynchronize();
        cudaMemcpy(ho, out, blocks * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int b = 0; b < blocks; ++b) s += ho[b];
        printf("%-28s %3zu MB  blocks %4d x %3d threads: %7.0f cycles per dependent access\n", name, mb, blocks, threads, s / blocks);
    };
    for (int threads : {1, 32}) {
        for (int blocks : {1, 148, 1184}) {
            // warm L2
            chase_ldcg<<<blocks, threads>>>(d, 2000, out, sink);
            chase_ldcg<<<blocks, threads>>>(d, 2000, out, sink); report("ld.cg 8B chase", blocks, threads);
            chase_ldg128<<<blocks, threads>>>(d4, 2000, out, sink); report("ld.nc 16B chase", blocks, threads);
            chase_atom<<<blocks, threads>>>(d, 2000, out, sink); report("atom.min.u64 chase", blocks, threads);
        }
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"ld.cg random 8B", "atom.min.u64 (return)", "red.min.u64", "st random 8B"};
    for (int mode = 0; mode < 4; ++mode) for (int bpsm : {2, 8}) {
        int blocks = 148 * bpsm, threads = 256, iters = 2000;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) tput<0><<<blocks, threads>>>(d, n, iters, sink);
            if (mode == 1) tput<1><<<blocks, threads>>>(d, n, iters, sink);
            if (mode == 2) tput<2><<<blocks, threads>>>(d, n, iters, sink);
            if (mode == 3) tput<3><<<blocks, threads>>>(d, n, iters, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("%-28s %3zu MB  %d CTAs/SM x 256: %8.2f G accesses/s\n", names[mode], mb, bpsm, (double)blocks * threads * iters / ms / 1e6);
        if (mode == 3) cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
