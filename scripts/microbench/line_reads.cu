// Micro-benchmark for the access pattern that bounds the tile solve: every warp issues 32-byte loads, four lanes per
// 128-byte line, eight different lines per warp instruction, lines picked pseudo-randomly from a buffer of a given
// size (L2-resident or HBM-resident), B loads in flight per lane before the first use.  Prints lines/s and GB/s for a
// sweep of buffer sizes, batch sizes and resident warps, i.e. the ceiling for "read K/2 finished neighbours per state".
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o line_reads line_reads.cu && ./line_reads
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld4(const double* p, double (&f)[4])
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(f[0]), "=d"(f[1]), "=d"(f[2]), "=d"(f[3]) : "l"(p));
}

template <int B>
__global__ void k_lines(const double* __restrict__ buf, uint64_t n_lines, int iters, double* __restrict__ sink)
{
    const int lane = threadIdx.x & 31, lc = lane & 3, lg = lane >> 2;
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t state = warp * 0x9E3779B97F4A7C15ull + (uint64_t)lg * 0xBF58476D1CE4E5B9ull + 12345;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        double v[B][4];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t line = (state >> 20) % n_lines;
            ld4(buf + line * 16 + lc * 4, v[b]);
        }
#pragma unroll
        for (int b = 0; b < B; ++b) acc += v[b][0] + v[b][1] + v[b][2] + v[b][3];
    }
    if (acc == 1.2345e300) sink[0] = acc;
}

template <int B>
static void run(const double* buf, uint64_t n_lines, int ctas_per_sm, int nsm, double* sink)
{
    const int iters = 2048 / B;
    const int grid = nsm * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_lines<B><<<grid, 256>>>(buf, n_lines, iters, sink);            // warm-up (and L2 fill for small buffers)
    cudaEventRecord(a);
    k_lines<B><<<grid, 256>>>(buf, n_lines, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double lines = (double)grid * 8 /*warps*/ * 8 /*lines per warp load*/ * B * iters;
    printf("buffer %7.0f MB  batch %d  warps/SM %2d : %7.2f G lines/s  %7.2f TB/s\n", n_lines * 128.0 / 1e6, B, ctas_per_sm * 8,
           lines / ms / 1e6, lines * 128.0 / ms / 1e9);
    cudaEventDestroy(a); cudaEventDestroy(b);
}

int main()
{
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    const size_t max_bytes = (size_t)4 << 30;
    double *buf = nullptr, *sink = nullptr;
    cudaMalloc(&buf, max_bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, max_bytes);
    printf("%s, %d SMs\n", prop.name, nsm);
    for (size_t mb : {16, 64, 256, 4096}) {
        const uint64_t n_lines = (uint64_t)mb * 1024 * 1024 / 128;
        for (int ctas : {2, 3, 4, 8}) {
            run<1>(buf, n_lines, ctas, nsm, sink);
            run<3>(buf, n_lines, ctas, nsm, sink);
            run<6>(buf, n_lines, ctas, nsm, sink);
        }
    }
    cudaFree(buf); cudaFree(sink);
    return 0;
}
