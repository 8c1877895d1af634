// Micro-benchmark ("decide cfg2 with a measurement"): does a DMMA stream keep its rate with
// integer / select work interleaved, unlike DFMA (which drops to 65 % / 42 %, dfma_occ.cu)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/dmma_issue tools/ubench/dmma_issue.cu
//
// All f64 mma.sync shapes lower to DMMA.8x8x4 on sm_100a (cuobjdump), 256 FMA per instruction:
// 16 FP64-pipe cycles per SM sub-partition at the measured 37 TFLOP/s, against 2 cycles / 32 FMA for a DFMA.
#include <cstdio>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ILP independent accumulator tiles; per DMMA: MIX integer ops and NF dependent-free DFMAs
template <int ILP, int MIX, int NF>
__global__ void k_dmma(double* out, int iters, double seed) {
    double c[ILP][2], f[4];
    int z[4] = {(int)threadIdx.x, 1, 2, 3};
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = seed + i;
    const double a = seed * 1e-3 + threadIdx.x * 1e-9, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                dmma(c[i], a, b);
#pragma unroll
                for (int q = 0; q < MIX; ++q) z[q & 3] = (z[q & 3] ^ (z[(q + 1) & 3] + u)) * 3 + it;
#pragma unroll
                for (int q = 0; q < NF; ++q) f[q & 3] = fma(f[q & 3], 1.0000001, 1e-9);
            }
        }
    }
    double s = f[0] + f[1] + f[2] + f[3];
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456 || z[0] + z[1] + z[2] + z[3] == 123456789) out[0] = s;
}

// warp-specialised mix: even warps run a DFMA + integer stream (the R phase's profile), odd warps a DMMA stream
template <int MIXR>
__global__ void k_split(double* out, int iters, double seed) {
    const int w = threadIdx.x >> 5;
    double s = 0;
    if (w & 1) {
        double c[4][2] = {};
        const double a = seed * 1e-3 + threadIdx.x * 1e-9, b = 1e-3;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(c[i], a, b);
        for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    } else {
        double f[4] = {seed, seed + 1, seed + 2, seed + 3};
        int z[4] = {(int)threadIdx.x, 1, 2, 3};
        for (int it = 0; it < iters * 4; ++it)   // x4: the same FP64-pipe time as the DMMA warps (64 vs 256 cycles per iteration)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < 4; ++i) f[i] = fma(f[i], 1.0000001, 1e-9);
#pragma unroll
                for (int q = 0; q < MIXR; ++q) z[q & 3] = (z[q & 3] ^ (z[(q + 1) & 3] + u)) * 3 + it;
            }
        s = f[0] + f[1] + f[2] + f[3] + (z[0] + z[1] + z[2] + z[3] == 123456789);
    }
    if (s == 123.456) out[0] = s;
}

// dependent DMMA chain on one warp: cycles per instruction
__global__ void k_lat(double* out, long long* cyc, double seed) {
    double c[2] = {0, 0};
    const double a = seed * 1e-3, b = 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) dmma(c, a, b);
    }
    long long t1 = clock64();
    // chain through the A operand (accumulator of one feeds the A operand of the next: the W -> T W -> update pattern)
    double d[2] = {0, 0}, aa = a;
    long long t2 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { d[0] = 0; d[1] = 0; dmma(d, aa, b); aa = d[0]; }
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
    out[threadIdx.x] = c[0] + c[1] + aa;
}

static float time_it(void (*launch)(int), int arg) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); launch(arg); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

static double* g_d;
static const int ITERS = 1024;
template <int ILP, int MIX, int NF> void launch_dmma(int w) { k_dmma<ILP, MIX, NF><<<148, w * 32>>>(g_d, ITERS, 1.0); }
template <int MIXR> void launch_split(int w) { k_split<MIXR><<<148, w * 32>>>(g_d, ITERS, 1.0); }

template <int ILP, int MIX, int NF> void run(int w) {
    float ms = time_it(launch_dmma<ILP, MIX, NF>, w);
    double n = (double)ITERS * 4 * ILP * 148.0 * w;           // DMMA (and groups) executed
    double tf = 2.0 * n * (256.0 + 32.0 * NF) / ms / 1e9;
    printf("DMMA ILP %d  +%2d int +%d DFMA per DMMA, warps/SM %2d : %6.2f TFLOP/s total  (DMMA part %5.1f%% of 37.0)\n", ILP, MIX, NF, w,
           tf, 2.0 * n * 256.0 / ms / 1e9 / 37.0 * 100);
}
template <int MIXR> void run_split(int w) {
    float ms = time_it(launch_split<MIXR>, w);
    double nd = (double)ITERS * 16 * 148.0 * (w / 2), nf = (double)ITERS * 4 * 32 * 148.0 * (w / 2);
    printf("split: %2d warps/SM (half DFMA+%d int per 4 DFMA, half DMMA): DMMA %6.2f + DFMA %6.2f TFLOP/s\n", w, MIXR,
           2.0 * nd * 256 / ms / 1e9, 2.0 * nf * 32 / ms / 1e9);
}

int main() {
    CHECK(cudaMalloc(&g_d, 4096));
    for (int w : {4, 8, 16}) {
        run<1, 0, 0>(w); run<2, 0, 0>(w); run<4, 0, 0>(w); run<8, 0, 0>(w);
        run<4, 4, 0>(w); run<4, 8, 0>(w); run<4, 12, 0>(w); run<4, 16, 0>(w);
        run<4, 0, 2>(w); run<4, 0, 4>(w); run<4, 4, 4>(w);
    }
    for (int w : {8, 16}) { run_split<0>(w); run_split<6>(w); run_split<12>(w); }
    long long* cyc; CHECK(cudaMalloc(&cyc, 64));
    k_lat<<<1, 32>>>(g_d, cyc, 1.0);
    long long h[2]; CHECK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("dependent DMMA.8x8x4: %.1f cycles via accumulator, %.1f cycles via A operand (incl. 2 moves)\n", h[0] / 1024.0, h[1] / 1024.0);
    return 0;
}
