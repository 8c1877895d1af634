// Latency micro-benchmarks behind the panel kernel's column step (one CTA of 256 threads, like the panel kernel):
// cycles per dependent round trip of the primitives the step is made of.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double rsqrt_seed(double x) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rcp_seed(double x) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
constexpr int REPS = 64;
__global__ void __cluster_dims__(1, 1, 1) __launch_bounds__(256, 1) lat_kernel(double* out, long long* cyc, int active_warps, double one) {
    __shared__ __align__(16) double buf[8][64];
    __shared__ double part[8][32];
    __shared__ double recv[32];
    __shared__ uint64_t bar;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool rec = threadIdx.x == 0;
    double x = one + lane * 1e-9, acc = 0.0;
    long long t0, t1;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool act = w < active_warps;
    // 0: dependent DFMA
    __syncthreads(); t0 = clock64();
    if (act) { for (int i = 0; i < REPS; ++i) x = fma(x, one, 1e-30); }
    t1 = clock64(); if (rec) cyc[0] = t1 - t0;
    // 1: dependent DADD
    __syncthreads(); t0 = clock64();
    if (act) { for (int i = 0; i < REPS; ++i) x = x + 1e-30; }
    t1 = clock64(); if (rec) cyc[1] = t1 - t0;
    // 2: 64-bit SHFL.BFLY + DADD
    __syncthreads(); t0 = clock64();
    if (act) { for (int i = 0; i < REPS; ++i) x += __shfl_xor_sync(0xffffffffu, x, 8) * 1e-30; }
    t1 = clock64(); if (rec) cyc[2] = t1 - t0;
    // 3: STS.128 -> syncwarp -> LDS.128
    __syncthreads(); t0 = clock64();
    if (act) {
        for (int i = 0; i < REPS; ++i) {
            if ((lane & 7) == (i & 7)) *reinterpret_cast<double2*>(&buf[w][(lane >> 3) * 16]) = make_double2(x, x);
            __syncwarp();
            const double2 v = *reinterpret_cast<const double2*>(&buf[w][(lane >> 3) * 16]);
            x = v.x * one;
            __syncwarp();
        }
    }
    t1 = clock64(); if (rec) cyc[3] = t1 - t0;
    // 4: STS -> __syncthreads -> 8 LDS + add tree  (all 8 warps take part)
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < REPS; ++i) {
        part[w][lane] = x;
        __syncthreads();
        double s0 = part[0][lane] + part[1][lane], s1 = part[2][lane] + part[3][lane];
        double s2 = part[4][lane] + part[5][lane], s3 = part[6][lane] + part[7][lane];
        x = ((s0 + s1) + (s2 + s3)) * 0.125;
        __syncthreads();
    }
    t1 = clock64(); if (rec) cyc[4] = t1 - t0;
    // 5: rsqrt seed chain
    __syncthreads(); t0 = clock64();
    if (act) { for (int i = 0; i < REPS; ++i) x = rsqrt_seed(x); }
    t1 = clock64(); if (rec) cyc[5] = t1 - t0;
    // 6: rcp seed chain
    __syncthreads(); t0 = clock64();
    if (act) { for (int i = 0; i < REPS; ++i) x = rcp_seed(x); }
    t1 = clock64(); if (rec) cyc[6] = t1 - t0;
    // 7: st.async to the own CTA + mbarrier wait (warp 0 only pushes 32 values, everyone waits)
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < REPS; ++i) {
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(256) : "memory");
        if (w == 0) {
            uint32_t ra, rb;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&recv[lane])), "r"(0));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(&bar)), "r"(0));
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(ra), "d"(x), "r"(rb) : "memory");
        }
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)(i & 1)) : "memory");
        }
        x = recv[lane] * one;
    }
    t1 = clock64(); if (rec) cyc[7] = t1 - t0;
    // 8: full Newton rsqrt (seed + 2 steps) + rcp (seed + 2 steps) dependent pair, as in the panel kernel
    __syncthreads(); t0 = clock64();
    if (act) {
        for (int i = 0; i < REPS; ++i) {
            double y = rsqrt_seed(x);
            const double hx = 0.5 * x;
            for (int it = 0; it < 2; ++it) { const double e = fma(-hx * y, y, 0.5); y = fma(y, e, y); }
            const double D = fma(x, y, 1.0);
            double u = rcp_seed(D);
            for (int it = 0; it < 2; ++it) { const double e = fma(-D, u, 1.0); u = fma(u, e, u); }
            x = u * 2.0;
        }
    }
    t1 = clock64(); if (rec) cyc[8] = t1 - t0;
    // 9: empty __syncthreads pair
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < REPS; ++i) { __syncthreads(); }
    t1 = clock64(); if (rec) cyc[9] = t1 - t0;
    // 10: 32 independent DFMA (one panel update's worth) per rep, dependent across reps
    double r8[32];
    for (int k = 0; k < 32; ++k) r8[k] = x + k;
    __syncthreads(); t0 = clock64();
    if (act) { for (int i = 0; i < REPS; ++i) { for (int k = 0; k < 32; ++k) r8[k] = fma(r8[k], one, x); } }
    t1 = clock64(); if (rec) cyc[10] = t1 - t0;
    for (int k = 0; k < 32; ++k) acc += r8[k];
    out[threadIdx.x] = x + acc;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 16 * 8);
    const char* names[] = {"DFMA dependent", "DADD dependent", "SHFL64+DMUL+DADD", "STS.128->syncwarp->LDS.128 (+DMUL, +syncwarp)", "STS->bar->8 LDS+3 DADD+DMUL->bar", "MUFU.RSQ64H chain", "MUFU.RCP64H chain", "arm + STAS(local) + try_wait + LDS", "rsqrt NR2 + rcp NR2 chain", "__syncthreads", "32 indep DFMA per rep"};
    for (int aw : {1, 8}) {
        for (int rep = 0; rep < 2; ++rep) lat_kernel<<<1, 256>>>(out, cyc, aw, 1.0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("active warps %d:\n", aw);
        for (int k = 0; k < 11; ++k) printf("  %-50s %8.1f cycles/rep\n", names[k], (double)h[k] / REPS);
    }
    return 0;
}
