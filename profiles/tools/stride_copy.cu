// Microbenchmark: is HBM slower for the PL kernel's access pattern (512-byte row segments at a 2 KB pitch, read + write)
// than for a contiguous copy?  Same bytes, same persistent 148-CTA shape.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(512) copy_contig(const double2* __restrict__ X, double2* __restrict__ G, size_t n2) {
    const size_t per = (n2 + gridDim.x - 1) / gridDim.x;
    const size_t lo = blockIdx.x * per, hi = min(n2, lo + per);
    for (size_t i = lo + threadIdx.x; i < hi; i += 512 * 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) if (i + u * 512 < hi) v[u] = __ldcs(X + i + u * 512);
#pragma unroll
        for (int u = 0; u < 8; u++) if (i + u * 512 < hi) __stcs(G + i + u * 512, v[u]);
    }
}

// item k = chunk * ntiles + tile (chunk-major, contiguous ranges per CTA as in pl_desc_kernel); a tile = TR consecutive rows;
// a warp moves one 512-byte row segment per instruction
template <int TR>
__global__ void __launch_bounds__(512) copy_tiles(const double* __restrict__ X, double* __restrict__ G, int P, int B, int ntiles, int nchunks) {
    const long long W = (long long)ntiles * nchunks, Gd = gridDim.x;
    const long long k_lo = blockIdx.x * W / Gd, k_hi = (blockIdx.x + 1) * W / Gd;
    const int w = threadIdx.x >> 5, tx = threadIdx.x & 31;
    for (long long k = k_lo; k < k_hi; k++) {
        const int chunk = (int)(k / ntiles), tile = (int)(k % ntiles);
        const int r0 = tile * TR;
        const size_t col = (size_t)chunk * 64 + 2 * tx;
        double2 v[TR / 16];
#pragma unroll
        for (int u = 0; u < TR / 16; u++) {
            const int r = r0 + w + 16 * u;
            if (r < P) v[u] = __ldcs(reinterpret_cast<const double2*>(X + (size_t)r * B + col));
        }
#pragma unroll
        for (int u = 0; u < TR / 16; u++) {
            const int r = r0 + w + 16 * u;
            if (r < P) __stcs(reinterpret_cast<double2*>(G + (size_t)r * B + col), v[u]);
        }
    }
}

int main() {
    const int P = 299996;
    for (int B : {64, 256, 1024}) {
        const size_t n = (size_t)P * B;
        double *X, *G;
        CK(cudaMalloc(&X, n * 8)); CK(cudaMalloc(&G, n * 8));
        CK(cudaMemset(X, 1, n * 8)); CK(cudaMemset(G, 0, n * 8));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int TR = 128, ntiles = (P + TR - 1) / TR, nchunks = B / 64;
        for (int variant = 0; variant < 3; variant++) {
            const int grid = variant == 2 ? 296 : 148;
            auto launch = [&]() {
                if (variant == 0) copy_contig<<<148, 512>>>((const double2*)X, (double2*)G, n / 2);
                else copy_tiles<TR><<<grid, 512>>>(X, G, P, B, ntiles, nchunks);
            };
            for (int i = 0; i < 5; i++) launch();
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            for (int i = 0; i < 20; i++) launch();
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
            printf("{\"B\": %d, \"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", B, variant == 0 ? "contiguous" : (variant == 1 ? "row-segments 148 CTAs" : "row-segments 296 CTAs"), ms, 2.0 * n * 8 / ms * 1e-6);
        }
        cudaFree(X); cudaFree(G);
    }
    return 0;
}
