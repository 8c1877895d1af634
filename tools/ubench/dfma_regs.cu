// Micro-benchmark: does the DFMA rate depend on how many DISTINCT register operands an instruction reads?
// (operand-collector / register-bank pressure: the kernels' FMAs read three different 64-bit registers, dfma_occ.cu's
// only one.)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/dfma_regs tools/ubench/dfma_regs.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, const double* in, int iters) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[i + threadIdx.x]; b[i] = in[8 + i + threadIdx.x]; c[i] = in[16 + i + threadIdx.x]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) c[i] = fma(a[0], b[0], c[i]);            // one varying operand (accumulator)
                if (MODE == 1) c[i] = fma(a[i], b[0], c[i]);            // two varying
                if (MODE == 2) c[i] = fma(a[i], b[i], c[i]);            // three varying
                if (MODE == 3) c[i] = fma(a[i], b[(i + u) & 7], c[i]);  // three varying, no fixed pairing
                if (MODE == 4) c[i] = fma(a[(i + u) & 7], b[i >> 1], c[i]);  // update-loop pattern: b shared by 2 neighbours
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(double* out, double* in, int w) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k<MODE><<<148, w * 32>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fmas = (double)iters * 32 * 148.0 * w * 32;
    printf("mode %d warps/SM %2d : %6.2f TFLOP/s\n", MODE, w, 2 * fmas / best / 1e9);
}
int main() {
    double *out, *in; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    for (int w : {4, 8, 16}) { run<0>(out, in, w); run<1>(out, in, w); run<2>(out, in, w); run<3>(out, in, w); run<4>(out, in, w); }
    return 0;
}
