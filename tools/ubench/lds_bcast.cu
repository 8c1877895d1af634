// Micro-benchmark: cost of broadcast-style shared-memory loads for different lane->address patterns.
#include <cstdio>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int WIDTH /*8 or 16 bytes*/, int ITER>
__global__ void k(const int* offs /*32 per-lane byte offsets*/, double* out, long long* cyc, int nwarps_active) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const char* base = reinterpret_cast<const char*>(sm) + offs[lane];
    double acc0 = 0, acc1 = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (WIDTH == 16) {
                double2 v = *reinterpret_cast<const double2*>(base + u * 256 + (it & 1) * 16);
                acc0 += v.x; acc1 += v.y;
            } else {
                double v = *reinterpret_cast<const double*>(base + u * 256 + (it & 1) * 16);
                acc0 += v;
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
}

int main() {
    int *d_offs; double* d_out; long long* d_cyc;
    CHECK(cudaMalloc(&d_offs, 128)); CHECK(cudaMalloc(&d_out, 8 * 1024 * 148)); CHECK(cudaMalloc(&d_cyc, 8 * 32 * 148));
    struct Pat { const char* name; int offs[32]; } pats[8];
    int np = 0;
    auto add = [&](const char* name, auto f) { pats[np].name = name; for (int l = 0; l < 32; ++l) pats[np].offs[l] = f(l); ++np; };
    add("all lanes same address", [](int l) { return 0; });
    add("2 groups of 16 (halves), banks 0/4", [](int l) { return (l / 16) * 16 + (l / 16) * 8192; });
    add("4 groups = quarters, stagger 16B", [](int l) { return (l / 8) * 16 + (l / 8) * 4096; });
    add("4 groups interleaved (l%4), stagger 16B", [](int l) { return (l % 4) * 16 + (l % 4) * 4096; });
    add("8 groups interleaved (l%8), stagger 16B", [](int l) { return (l % 8) * 16 + (l % 8) * 2048; });
    add("distinct per lane, 16B stride (no bcast)", [](int l) { return l * 16; });
    add("4 groups = quarters, same bank (conflict)", [](int l) { return (l / 8) * 4096; });
    const int ITER = 256;
    for (int width : {16, 8}) {
        for (int warps : {1, 4, 8}) {
            for (int p = 0; p < np; ++p) {
                CHECK(cudaMemcpy(d_offs, pats[p].offs, 128, cudaMemcpyHostToDevice));
                if (width == 16) k<16, ITER><<<148, warps * 32, 40960>>>(d_offs, d_out, d_cyc, warps);
                else k<8, ITER><<<148, warps * 32, 40960>>>(d_offs, d_out, d_cyc, warps);
                CHECK(cudaDeviceSynchronize());
                long long h[8]; CHECK(cudaMemcpy(h, d_cyc, 8 * warps, cudaMemcpyDeviceToHost));
                long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
                printf("LDS.%d warps/SM=%d  %-45s  %.2f cyc per LDS per warp, %.2f cyc/LDS SM-wide\n", width * 8, warps, pats[p].name,
                       (double)mx / (ITER * 16), (double)mx / (ITER * 16) / warps);
            }
        }
    }
    return 0;
}
