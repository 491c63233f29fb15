// Micro-benchmark (not product code): what does a random gather from a text-sized array cost on B200,
// as a function of the bytes taken per access and their alignment?  Decides the refinement kernel's
// fetch shape.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// MODE 0: one u64 at a random word
// MODE 1: two consecutive u64 at a random word (unaligned 16 B)
// MODE 2: aligned 32 B (2 x 128-bit)
// MODE 3: four consecutive u64 at a random word (unaligned 32 B)
// MODE 4: aligned 64 B (4 x 128-bit)
// MODE 5: aligned 16 B (1 x 128-bit)
// MODE 6: aligned 128 B (8 x 128-bit)
// FLAV 0: ld.global.nc  1: ld.global  2: ld.global.cg  3: ld.global.nc.L2::64B  4: ld.global.cv  5: ld.global.nc.L1::no_allocate
template <int FLAV> __device__ __forceinline__ uint64_t ld64(const uint64_t *p)
{
    uint64_t v;
    if (FLAV == 0) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (FLAV == 1) asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (FLAV == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (FLAV == 3) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (FLAV == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
template <int FLAV, int N>
__global__ void __launch_bounds__(256) gflav(const uint64_t *__restrict__ a, uint64_t words, uint64_t *__restrict__ out,
                                             uint64_t total)
{
    const uint64_t t0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x);
    if (t0 >= total) return;
    uint64_t acc = 0;
    uint64_t w = mix(t0) % (words - 16);
#pragma unroll
    for (int k = 0; k < N; ++k) acc ^= ld64<FLAV>(a + w + k);
    if (acc == 0x1234567) out[0] = acc;
}
template <int FLAV, int N> static void runf(const char *name, const uint64_t *a, uint64_t words, uint64_t *out, uint64_t total)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((total + 255) / 256);
    gflav<FLAV, N><<<grid, 256>>>(a, words, out, total);
    cudaEventRecord(e0);
    gflav<FLAV, N><<<grid, 256>>>(a, words, out, total);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("flavour %-22s words %d array %6.0f MB  %8.2f ms  %7.2f G gathers/s\n", name, N, words * 8 / 1e6, ms,
           total / ms / 1e6);
}

template <int MODE, int PER>
__global__ void __launch_bounds__(256) gather(const uint64_t *__restrict__ a, uint64_t words, uint64_t *__restrict__ out,
                                              uint64_t total)
{
    const uint64_t t0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * PER;
    if (t0 >= total) return;
    uint64_t acc = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        uint64_t w = mix(t0 + u) % (words - 16);
        if (MODE == 0) {
            acc ^= __ldg(a + w);
        } else if (MODE == 1) {
            acc ^= __ldg(a + w) + __ldg(a + w + 1);
        } else if (MODE == 2) {
            const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(a + (w & ~3ull));
            ulonglong2 x = __ldg(p), y = __ldg(p + 1);
            acc ^= x.x + x.y + y.x + y.y;
        } else if (MODE == 3) {
            acc ^= __ldg(a + w) + __ldg(a + w + 1) + __ldg(a + w + 2) + __ldg(a + w + 3);
        } else if (MODE == 4) {
            const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(a + (w & ~7ull));
            ulonglong2 x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2), v = __ldg(p + 3);
            acc ^= x.x + x.y + y.x + y.y + z.x + z.y + v.x + v.y;
        } else if (MODE == 5) {
            const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(a + (w & ~1ull));
            ulonglong2 x = __ldg(p);
            acc ^= x.x + x.y;
        } else {
            const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(a + (w & ~15ull));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ulonglong2 x = __ldg(p + k);
                acc ^= x.x + x.y;
            }
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

template <int MODE, int PER> static void run(const char *name, const uint64_t *a, uint64_t words, uint64_t *out, uint64_t total)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((total / PER + 255) / 256);
    gather<MODE, PER><<<grid, 256>>>(a, words, out, total);
    cudaEventRecord(e0);
    gather<MODE, PER><<<grid, 256>>>(a, words, out, total);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s per-thread %d array %6.0f MB  %8.2f ms  %7.2f G gathers/s\n", name, PER, words * 8 / 1e6, ms,
           total / ms / 1e6);
}

int main(int argc, char **argv)
{
    const uint64_t total = 1ull << 30;
    if (argc > 1) {
        size_t g = 0;
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("L2 fetch granularity limit: %zu\n", g);
    }
    {
        const uint64_t words = 757500000ull / 8;
        uint64_t *a, *out;
        cudaMalloc(&a, words * 8);
        cudaMalloc(&out, 8);
        cudaMemset(a, 1, words * 8);
        runf<0, 1>("ld.global.nc", a, words, out, total);
        runf<1, 1>("ld.global", a, words, out, total);
        runf<2, 1>("ld.global.cg", a, words, out, total);
        runf<3, 1>("ld.global.nc.L2::64B", a, words, out, total);
        runf<4, 1>("ld.global.cv", a, words, out, total);
        runf<5, 1>("ld.global.nc.L1::no_alloc", a, words, out, total);
        runf<0, 2>("ld.global.nc", a, words, out, total);
        runf<2, 2>("ld.global.cg", a, words, out, total);
        runf<0, 4>("ld.global.nc", a, words, out, total);
        runf<2, 4>("ld.global.cg", a, words, out, total);
        runf<0, 8>("ld.global.nc", a, words, out, total);
        runf<0, 16>("ld.global.nc", a, words, out, total);
        cudaFree(a);
        cudaFree(out);
    }
    for (int pass = 0; pass < (argc > 2 ? 2 : 0); ++pass) {
        const uint64_t words = pass == 0 ? 757500000ull / 8 : 64000000ull / 8;
        uint64_t *a, *out;
        cudaMalloc(&a, words * 8);
        cudaMalloc(&out, 8);
        cudaMemset(a, 1, words * 8);
        run<0, 1>("u64", a, words, out, total);
        run<0, 4>("u64", a, words, out, total);
        run<1, 1>("2 x u64 unaligned", a, words, out, total);
        run<1, 4>("2 x u64 unaligned", a, words, out, total);
        run<5, 4>("16 B aligned", a, words, out, total);
        run<2, 1>("32 B aligned", a, words, out, total);
        run<2, 4>("32 B aligned", a, words, out, total);
        run<3, 1>("4 x u64 unaligned", a, words, out, total);
        run<3, 4>("4 x u64 unaligned", a, words, out, total);
        run<4, 1>("64 B aligned", a, words, out, total);
        run<4, 2>("64 B aligned", a, words, out, total);
        run<6, 1>("128 B aligned", a, words, out, total);
        cudaFree(a);
        cudaFree(out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
