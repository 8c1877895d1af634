// Micro-benchmark: FP64 FMA throughput vs resident warps per SM and independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int ILP, int MIX>
__global__ void k(double* out, int iters, double seed) {
    double a[ILP];
    int z[4] = {threadIdx.x, 1, 2, 3};
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i + threadIdx.x * 1e-9;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
            if (MIX) {
#pragma unroll
                for (int q = 0; q < MIX; ++q) z[q & 3] = (z[q & 3] ^ (z[(q + 1) & 3] + u)) * 3 + it;  // integer filler
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456 || z[0] + z[1] + z[2] + z[3] == 123456789) out[0] = s;
}

template <int ILP, int MIX>
int run(double* d, int warps_per_sm) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<ILP, MIX><<<148, warps_per_sm * 32>>>(d, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fmas = (double)iters * 8 * ILP * 148.0 * warps_per_sm * 32;
    printf("ILP %d mix %d warps/SM %2d : %.2f TFLOP/s (%.1f%% of 36.7)\n", ILP, MIX, warps_per_sm, 2 * fmas / best / 1e9, 2 * fmas / best / 1e9 / 36.7 * 100);
    return 0;
}

int main() {
    double* d; CHECK(cudaMalloc(&d, 64));
    for (int w : {4, 8, 12, 16, 32}) {
        run<1, 0>(d, w); run<2, 0>(d, w); run<4, 0>(d, w); run<8, 0>(d, w);
        run<4, 2>(d, w); run<8, 4>(d, w); run<8, 8>(d, w);
    }
    return 0;
}
