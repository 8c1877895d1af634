// Batched QR / least-squares entry points (a1-a4 of SURVEY.md section 8).
#include <algorithm>

#include "../../include/linalg_b200.h"
#include "batched_qr32.cuh"
#include "batched_qr32_pipe.cuh"
#include "batched_qr32_dmma.cuh"
#include "batched_qr32_ll.cuh"
#include "batched_qr32_c8.cuh"
#include "batched_mgs32_c8.cuh"
#include "batched_small.cuh"
#include "ctx.cuh"
#include "ops.cuh"

namespace lq {

// ------------------------------------------------------------------ 32x32 fast kernels
template <int P, int C, int WARPS, bool KEEPV, int MINB, int NR = 3>
static int launch_hh32(Ctx* c, cudaStream_t st, const double* A, long long batch, double* Q, double* R) {
    using D = Dist32<P, C>;
    auto kern = hh_qr32_kernel<P, C, WARPS, KEEPV, MINB, NR>;
    const size_t smem = (size_t)WARPS * D::MPW * D::SMEM_DOUBLES * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(c->device);
    }
    const long long per_block = (long long)WARPS * D::MPW;
    const long long blocks = (batch + per_block - 1) / per_block;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, batch);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

// software-pipelined kernel (R phase of pair n+1 interleaved with the Q phase of pair n), persistent grid
template <int WARPS, int MINB, int NR>
static int launch_hh32_pipe(Ctx* c, cudaStream_t st, const double* A, long long batch, double* Q, double* R) {
    auto kern = hh_qr32_pipe_kernel<WARPS, MINB, NR>;
    const size_t smem = (size_t)WARPS * Pipe32::WARP_DOUBLES * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(c->device);
    }
    const long long pairs = (batch + 1) / 2;
    const long long blocks = std::min<long long>((long long)c->sm_count * MINB, (pairs + WARPS - 1) / WARPS);
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, batch);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

// R phase on DFMA, Q formation as compact-WY block reflectors on the FP64 tensor pipe (DMMA.8x8x4)
template <int WARPS, int MINB, int PHASES = 3>
static int launch_hh32_dmma(Ctx* c, cudaStream_t st, const double* A, long long batch, double* Q, double* R) {
    auto kern = hh_qr32_dmma_kernel<WARPS, MINB, PHASES>;
    const size_t smem = (size_t)WARPS * Dmma32::warp_doubles() * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(c->device);
    }
    const long long per_block = (long long)WARPS * 2;
    const long long blocks = (batch + per_block - 1) / per_block;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, batch);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

// register-light left-looking R phase (NS column stages, packed reflectors) + DMMA Q phase
template <int NS, int WARPS, int MINB, int PHASES = 3, bool KEEPV = false>
static int launch_hh32_ll(Ctx* c, cudaStream_t st, const double* A, long long batch, double* Q, double* R) {
    auto kern = hh_qr32_ll_kernel<NS, WARPS, MINB, PHASES, KEEPV>;
    const size_t smem = (size_t)WARPS * Pack32::WARP_DOUBLES * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(c->device);
    }
    const long long per_block = (long long)WARPS * 2;
    const long long blocks = (batch + per_block - 1) / per_block;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, batch);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

// lane = column, four matrices per warp, left-looking panels + DMMA Q phase
template <int WARPS, int MINB, int PHASES = 3, bool KEEPV = false, bool PF = false, bool PF2 = false, typename LAY = Col8>
static int launch_hh32_c8(Ctx* c, cudaStream_t st, const double* A, long long batch, double* Q, double* R) {
    auto kern = hh_qr32_c8_kernel<WARPS, MINB, PHASES, KEEPV, PF, PF2, LAY>;
    const size_t smem = (size_t)WARPS * LAY::WARP_DOUBLES * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(c->device);
    }
    const long long per_block = (long long)WARPS * 4;
    const long long blocks = (batch + per_block - 1) / per_block;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, batch);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

template <int P, int C, int WARPS, int MINB>
static int launch_mgs32(Ctx* c, cudaStream_t st, const double* A, long long batch, int reorth, double* Q, double* R,
                        int* info) {
    using D = Dist32<P, C>;
    auto kern = mgs_qr32_kernel<P, C, WARPS, MINB>;
    const size_t smem = (size_t)WARPS * D::MPW * D::MGS_DOUBLES * sizeof(double);
    const long long per_block = (long long)WARPS * D::MPW;
    const long long blocks = (batch + per_block - 1) / per_block;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, info, batch, reorth);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

// lane = column MGS (round 2): four matrices per warp, seven warps per SM
template <int WARPS, int MINB, bool KEEPV>
static int launch_mgs32_c8(Ctx* c, cudaStream_t st, const double* A, long long batch, int reorth, double* Q, double* R, int* info) {
    auto kern = mgs_qr32_c8_kernel<WARPS, MINB, KEEPV>;
    const size_t smem = (size_t)WARPS * MgsCol8::WARP_DOUBLES * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set(c->device);
    }
    const long long per_block = (long long)WARPS * 4;
    const long long blocks = (batch + per_block - 1) / per_block;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, Q, R, info, batch, reorth);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

int hh_qr_batched_stream(Ctx* c, cudaStream_t st, const double* A, long long batch, int m, int n, double* Q,
                         double* R, int variant) {
    if (batch == 0) return LQ_OK;
    if (m == 32 && n == 32 && variant >= 0) {
        switch (variant) {
            case 0:   // default (round 2): lane = column, left-looking panels, Q formation on DMMA, one 8-warp CTA per SM
            case 59: return launch_hh32_c8<8, 1, 3, true>(c, st, A, batch, Q, R);
            case 14: return launch_hh32<2, 4, 2, true, 4, -1>(c, st, A, batch, Q, R);  // round-1 default: reciprocal seeded from the raw rsqrt
#ifdef LQ_ALL_VARIANTS  // design-space variants measured in profiles/ (build with LINALG_B200_ALL_VARIANTS=1; tools/sweep_hh32.py)
            case 82: return launch_hh32_c8<8, 1, 3, true, false, false, Col8P>(c, st, A, batch, Q, R);  // padded reflector storage
            case 80: return launch_hh32_c8<8, 1, 3, true, false, true>(c, st, A, batch, Q, R);   // next panel prefetched into registers
            case 81: return launch_hh32_c8<8, 1, 3, false, false, true>(c, st, A, batch, Q, R);
            case 6: return launch_hh32<2, 4, 2, true, 4, 2>(c, st, A, batch, Q, R);   // 2 Newton steps, reciprocal behind the norm
            case 13: return launch_hh32_pipe<2, 4, 2>(c, st, A, batch, Q, R);
            case 21: return launch_hh32_dmma<4, 2>(c, st, A, batch, Q, R);
            case 52: return launch_hh32_c8<4, 2>(c, st, A, batch, Q, R);
            case 60: return launch_hh32_c8<8, 1, 3, false>(c, st, A, batch, Q, R);
            case 36: return launch_hh32_ll<2, 4, 3, 3, true>(c, st, A, batch, Q, R);
            case 5: return launch_hh32<2, 4, 2, true, 4>(c, st, A, batch, Q, R);
            case 1: return launch_hh32<1, 1, 4, false, 3>(c, st, A, batch, Q, R);
            case 2: return launch_hh32<1, 2, 4, false, 2>(c, st, A, batch, Q, R);
            case 3: return launch_hh32<2, 2, 4, false, 4>(c, st, A, batch, Q, R);
            case 4: return launch_hh32<2, 4, 2, false, 5>(c, st, A, batch, Q, R);
            case 7: return launch_hh32<2, 4, 1, true, 8>(c, st, A, batch, Q, R);
            case 8: return launch_hh32<4, 4, 4, true, 4>(c, st, A, batch, Q, R);
            case 9: return launch_hh32<4, 4, 4, true, 3>(c, st, A, batch, Q, R);
            case 10: return launch_hh32<4, 4, 4, false, 5>(c, st, A, batch, Q, R);
            case 11: return launch_hh32<2, 2, 4, true, 3>(c, st, A, batch, Q, R);
            case 12: return launch_hh32<4, 4, 2, true, 8, 2>(c, st, A, batch, Q, R);
            case 20: return launch_hh32_dmma<2, 4>(c, st, A, batch, Q, R);
            case 22: return launch_hh32_dmma<1, 8>(c, st, A, batch, Q, R);
            case 23: return launch_hh32_dmma<4, 2, 1>(c, st, A, batch, Q, R);
            case 50: return launch_hh32_c8<2, 5>(c, st, A, batch, Q, R);
            case 51: return launch_hh32_c8<2, 4>(c, st, A, batch, Q, R);
            case 53: return launch_hh32_c8<2, 5, 1>(c, st, A, batch, Q, R);
            case 54: return launch_hh32_c8<2, 5, 2>(c, st, A, batch, Q, R);
            case 55: return launch_hh32_c8<2, 5, 3, true>(c, st, A, batch, Q, R);
            case 56: return launch_hh32_c8<2, 4, 3, true>(c, st, A, batch, Q, R);
            case 57: return launch_hh32_c8<2, 5, 1, true>(c, st, A, batch, Q, R);
            case 58: return launch_hh32_c8<4, 2, 3, true>(c, st, A, batch, Q, R);
            case 64: return launch_hh32_c8<8, 1, 3, true, true>(c, st, A, batch, Q, R);   // + L2 prefetch of the later panels
            case 65: return launch_hh32_c8<8, 1, 3, false, true>(c, st, A, batch, Q, R);
            case 66: return launch_hh32_c8<4, 2, 3, true, true>(c, st, A, batch, Q, R);
            case 61: return launch_hh32_c8<4, 2, 1, true>(c, st, A, batch, Q, R);
            case 62: return launch_hh32_c8<4, 2, 1, false>(c, st, A, batch, Q, R);
            case 63: return launch_hh32_c8<4, 2, 2, false>(c, st, A, batch, Q, R);
            case 30: return launch_hh32_ll<2, 4, 4>(c, st, A, batch, Q, R);
            case 32: return launch_hh32_ll<2, 4, 4, 1>(c, st, A, batch, Q, R);
            case 34: return launch_hh32_ll<2, 4, 3>(c, st, A, batch, Q, R);
            case 37: return launch_hh32_ll<2, 4, 3, 1, true>(c, st, A, batch, Q, R);
            case 38: return launch_hh32_ll<1, 4, 2, 3, true>(c, st, A, batch, Q, R);
            case 39: return launch_hh32_ll<1, 4, 2, 1, true>(c, st, A, batch, Q, R);
            case 40: return launch_hh32_ll<2, 4, 4, 2>(c, st, A, batch, Q, R);
            case 41: return launch_hh32_ll<2, 2, 6, 3, true>(c, st, A, batch, Q, R);
            case 24: return launch_hh32_dmma<4, 2, 2>(c, st, A, batch, Q, R);
#endif
            default: break;
        }
        set_error(c, "householder_qr_batched: unknown kernel variant %d (the design-space variants need a library built with "
                     "LINALG_B200_ALL_VARIANTS=1)", variant);
        return LQ_ERR_ARG;
    }
    const size_t smem = small_hh_smem_doubles(m, n, 0) * sizeof(double);
    if (smem <= (size_t)c->max_smem) {
        static DeviceLatch configured;
        if (!configured.test(c->device)) {
            LQ_CUDA(c, cudaFuncSetAttribute(small_hh_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
            configured.set(c->device);
        }
        small_hh_kernel<0><<<(unsigned)batch, 256, smem, st>>>(A, nullptr, Q, R, nullptr, m, n, 0);
        LQ_CHECK_LAUNCH(c);
        LQ_COUNT_LAUNCH(c);
        return LQ_OK;
    }
    // too large for one CTA: blocked path, matrix by matrix (stream `st` must be the context stream)
    for (long long b = 0; b < batch; ++b)
        LQ_TRY(blocked_householder_qr(c, A + b * (long long)m * n, m, n, Q + b * (long long)m * n,
                                      R + b * (long long)n * n));
    return LQ_OK;
}

int mgs_qr_batched_stream(Ctx* c, cudaStream_t st, const double* A, long long batch, int m, int n, int reorth,
                          double* Q, double* R, int* info) {
    if (batch == 0) return LQ_OK;
    if (m == 32 && n == 32) {
        static const int mgs_variant = getenv("LINALG_B200_MGS_VARIANT") ? atoi(getenv("LINALG_B200_MGS_VARIANT")) : 0;  // read once
        switch (mgs_variant) {
#ifdef LQ_ALL_VARIANTS
            // lane = column form (round-2 experiment, profiles/ubench_r2.md): 109 M matrices/s against 135 for the default
            case 1: return launch_mgs32_c8<7, 1, true>(c, st, A, batch, reorth, Q, R, info);
            case 2: return launch_mgs32_c8<7, 1, false>(c, st, A, batch, reorth, Q, R, info);
            case 3: return launch_mgs32_c8<3, 2, true>(c, st, A, batch, reorth, Q, R, info);
#endif
            default: return launch_mgs32<2, 4, 2, 4>(c, st, A, batch, reorth, Q, R, info);
        }
    }
    const size_t smem = small_mgs_smem_doubles(m, n, 0) * sizeof(double);
    if (smem <= (size_t)c->max_smem) {
        static DeviceLatch configured;
        if (!configured.test(c->device)) {
            LQ_CUDA(c, cudaFuncSetAttribute(small_mgs_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
            configured.set(c->device);
        }
        small_mgs_kernel<0><<<(unsigned)batch, 256, smem, st>>>(A, nullptr, Q, R, nullptr, info, m, n, 0, reorth);
        LQ_CHECK_LAUNCH(c);
        LQ_COUNT_LAUNCH(c);
        return LQ_OK;
    }
    for (long long b = 0; b < batch; ++b)
        LQ_TRY(large_mgs_qr(c, A + b * (long long)m * n, m, n, reorth, Q + b * (long long)m * n,
                            R + b * (long long)n * n, info ? info + b : nullptr));
    return LQ_OK;
}

// info (device, may be null): per system 0, or 1 + the first column whose R[j][j] is exactly 0 (np.linalg.solve raises
// LinAlgError("Singular matrix") there, linalg/qr.py:134); reported by the warp-per-system kernel (n <= 64, nrhs <= 16),
// 0 for the other shapes
int lstsq_hh_batched_stream(Ctx* c, cudaStream_t st, const double* A, const double* B, long long batch, int m, int n,
                            int nrhs, double* X, int* info = nullptr) {
    if (batch == 0) return LQ_OK;
    int rc = lstsq_tile_kernel_launch(c, st, A, B, batch, m, n, nrhs, X, info, 2);
    if (rc != LQ_ERR_UNSUPPORTED) return rc;
    if (info) LQ_CUDA(c, cudaMemsetAsync(info, 0, sizeof(int) * (size_t)batch, st));
    rc = lstsq_stream_kernel_launch(c, st, A, B, batch, m, n, nrhs, X);
    if (rc != LQ_ERR_UNSUPPORTED) return rc;
    const size_t smem = small_hh_smem_doubles(m, n, nrhs) * sizeof(double);
    if (smem <= (size_t)c->max_smem) {
        static DeviceLatch configured;
        if (!configured.test(c->device)) {
            LQ_CUDA(c, cudaFuncSetAttribute(small_hh_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
            configured.set(c->device);
        }
        small_hh_kernel<1><<<(unsigned)batch, 256, smem, st>>>(A, B, nullptr, nullptr, X, m, n, nrhs);
        LQ_CHECK_LAUNCH(c);
        LQ_COUNT_LAUNCH(c);
        return LQ_OK;
    }
    for (long long b = 0; b < batch; ++b)
        LQ_TRY(large_lstsq_householder(c, A + b * (long long)m * n, B + b * (long long)m * nrhs, m, n, nrhs,
                                       X + b * (long long)n * nrhs));
    return LQ_OK;
}

int lstsq_mgs_batched_stream(Ctx* c, cudaStream_t st, const double* A, const double* B, long long batch, int m, int n,
                             int nrhs, double* X, int* info) {
    if (batch == 0) return LQ_OK;
    // Only X and the dependence report are observable (linalg/qr.py:103-119): the solution is the one of the
    // Householder entry point and |R[j][j]| does not depend on how the QR factorisation was computed.
    {
        const int rc = lstsq_tile_kernel_launch(c, st, A, B, batch, m, n, nrhs, X, info, 1);
        if (rc != LQ_ERR_UNSUPPORTED) return rc;
    }
    const size_t smem = small_mgs_smem_doubles(m, n, nrhs) * sizeof(double);
    if (smem <= (size_t)c->max_smem) {
        static DeviceLatch configured;
        if (!configured.test(c->device)) {
            LQ_CUDA(c, cudaFuncSetAttribute(small_mgs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
            configured.set(c->device);
        }
        small_mgs_kernel<1><<<(unsigned)batch, 256, smem, st>>>(A, B, nullptr, nullptr, X, info, m, n, nrhs, 0);
        LQ_CHECK_LAUNCH(c);
        LQ_COUNT_LAUNCH(c);
        return LQ_OK;
    }
    for (long long b = 0; b < batch; ++b)
        LQ_TRY(large_lstsq_mgs(c, A + b * (long long)m * n, B + b * (long long)m * nrhs, m, n, nrhs,
                               X + b * (long long)n * nrhs, info ? info + b : nullptr));
    return LQ_OK;
}

// ------------------------------------------------------------------ host-pointer pipelines
// Chunks of the batch flow through two lanes (stream + device buffers each): H2D, kernel and D2H
// of neighbouring chunks overlap on the copy engines.  With pinned host buffers the copies are
// truly asynchronous; pageable buffers still work (the runtime stages them).
struct ChunkPlan {
    long long chunk = 0;
};
static long long pick_chunk(long long batch, size_t bytes_per_item) {
    const size_t target = (size_t)192 << 20;  // ~192 MiB of device traffic per chunk
    long long ch = (long long)std::max<size_t>(1, target / std::max<size_t>(1, bytes_per_item));
    return std::min(batch, ch);
}

template <typename Fn>
static int run_chunked(Ctx* c, long long batch, size_t in_a, size_t in_b, size_t out_a, size_t out_b, size_t out_i,
                       const void* hA, const void* hB, void* hOa, void* hOb, void* hInfo, Fn&& fn) {
    if (batch == 0) return LQ_OK;
    LQ_CUDA(c, cudaSetDevice(c->device));
    const long long chunk = pick_chunk(batch, in_a + in_b + out_a + out_b + out_i);
    DevBuf dA[2], dB[2], dOa[2], dOb[2], dI[2];
    for (int l = 0; l < 2; ++l) {
        LQ_TRY(dA[l].alloc(c, in_a * chunk, c->lane[l]));
        LQ_TRY(dB[l].alloc(c, in_b * chunk, c->lane[l]));
        LQ_TRY(dOa[l].alloc(c, out_a * chunk, c->lane[l]));
        LQ_TRY(dOb[l].alloc(c, out_b * chunk, c->lane[l]));
        LQ_TRY(dI[l].alloc(c, out_i * chunk, c->lane[l]));
    }
    int k = 0;
    for (long long off = 0; off < batch; off += chunk, ++k) {
        const int l = k & 1;
        cudaStream_t st = c->lane[l];
        const long long cnt = std::min(chunk, batch - off);
        LQ_CUDA(c, cudaMemcpyAsync(dA[l].p, (const char*)hA + off * in_a, in_a * cnt, cudaMemcpyHostToDevice, st));
        if (in_b)
            LQ_CUDA(c, cudaMemcpyAsync(dB[l].p, (const char*)hB + off * in_b, in_b * cnt, cudaMemcpyHostToDevice, st));
        LQ_TRY(fn(st, cnt, dA[l].p, dB[l].p, dOa[l].p, dOb[l].p, dI[l].p));
        if (out_a)
            LQ_CUDA(c, cudaMemcpyAsync((char*)hOa + off * out_a, dOa[l].p, out_a * cnt, cudaMemcpyDeviceToHost, st));
        if (out_b)
            LQ_CUDA(c, cudaMemcpyAsync((char*)hOb + off * out_b, dOb[l].p, out_b * cnt, cudaMemcpyDeviceToHost, st));
        if (out_i && hInfo)
            LQ_CUDA(c, cudaMemcpyAsync((char*)hInfo + off * out_i, dI[l].p, out_i * cnt, cudaMemcpyDeviceToHost, st));
    }
    for (int l = 0; l < 2; ++l) LQ_CUDA(c, cudaStreamSynchronize(c->lane[l]));
    return LQ_OK;
}

static bool fits_small(Ctx* c, size_t doubles) { return doubles * sizeof(double) <= (size_t)c->max_smem; }

}  // namespace lq

using namespace lq;

#define LQ_ARGS_QR(c, A, batch, m, n)                                                                       \
    LQ_REQUIRE(c, c != nullptr, LQ_ERR_ARG, "null context");                                               \
    LQ_REQUIRE(c, batch >= 0 && m >= 1 && n >= 1, LQ_ERR_SHAPE, "bad shape batch=%lld m=%d n=%d",        \
               (long long)batch, m, n);                                                                    \
    LQ_REQUIRE(c, batch == 0 || A != nullptr, LQ_ERR_ARG, "null input pointer")

extern "C" {

int lq_householder_qr_batched_dev(lq_ctx* h, const double* A, int64_t batch, int m, int n, double* Q, double* R,
                                  int variant) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, m >= n, LQ_ERR_SHAPE, "householder_qr needs m >= n (got %d x %d), linalg/qr.py:52", m, n);
    LQ_CUDA(c, cudaSetDevice(c->device));
    return hh_qr_batched_stream(c, c->stream, A, batch, m, n, Q, R, variant);
}

int lq_householder_qr_batched(lq_ctx* h, const double* A, int64_t batch, int m, int n, double* Q, double* R) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, m >= n, LQ_ERR_SHAPE, "householder_qr needs m >= n (got %d x %d), linalg/qr.py:52", m, n);
    const size_t mn = (size_t)m * n * 8, nn = (size_t)n * n * 8;
    if (!(m == 32 && n == 32) && !fits_small(c, small_hh_smem_doubles(m, n, 0))) {
        // large matrices: one at a time through the blocked path on the context stream
        for (int64_t b = 0; b < batch; ++b)
            LQ_TRY(lq_householder_qr(h, A + b * (size_t)m * n, m, n, Q + b * (size_t)m * n, R + b * (size_t)n * n));
        return LQ_OK;
    }
    return run_chunked(c, batch, mn, 0, mn, nn, 0, A, nullptr, Q, R, nullptr,
                       [&](cudaStream_t st, long long cnt, void* dA, void*, void* dQ, void* dR, void*) {
                           return hh_qr_batched_stream(c, st, (const double*)dA, cnt, m, n, (double*)dQ, (double*)dR, 0);
                       });
}

int lq_mgs_qr_batched_dev(lq_ctx* h, const double* A, int64_t batch, int m, int n, int reorth, double* Q, double* R,
                          int32_t* info) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_CUDA(c, cudaSetDevice(c->device));
    return mgs_qr_batched_stream(c, c->stream, A, batch, m, n, reorth, Q, R, info);
}

int lq_mgs_qr_batched(lq_ctx* h, const double* A, int64_t batch, int m, int n, int reorth, double* Q, double* R,
                      int32_t* info) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    const size_t mn = (size_t)m * n * 8, nn = (size_t)n * n * 8;
    if (!(m == 32 && n == 32) && !fits_small(c, small_mgs_smem_doubles(m, n, 0))) {
        for (int64_t b = 0; b < batch; ++b)
            LQ_TRY(lq_mgs_qr(h, A + b * (size_t)m * n, m, n, reorth, Q + b * (size_t)m * n, R + b * (size_t)n * n,
                             info ? info + b : nullptr));
        return LQ_OK;
    }
    return run_chunked(c, batch, mn, 0, mn, nn, sizeof(int32_t), A, nullptr, Q, R, info,
                       [&](cudaStream_t st, long long cnt, void* dA, void*, void* dQ, void* dR, void* dI) {
                           return mgs_qr_batched_stream(c, st, (const double*)dA, cnt, m, n, reorth, (double*)dQ,
                                                        (double*)dR, (int*)dI);
                       });
}

int lq_lstsq_householder_batched_dev(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n,
                                     int nrhs, double* X) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, m >= n && nrhs >= 1, LQ_ERR_SHAPE, "least squares needs m >= n and nrhs >= 1 (got %d x %d, %d rhs)", m,
               n, nrhs);
    LQ_CUDA(c, cudaSetDevice(c->device));
    return lstsq_hh_batched_stream(c, c->stream, A, B, batch, m, n, nrhs, X);
}

int lq_lstsq_householder_batched_info_dev(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n,
                                          int nrhs, double* X, int32_t* info) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, m >= n && nrhs >= 1, LQ_ERR_SHAPE, "least squares needs m >= n and nrhs >= 1 (got %d x %d, %d rhs)", m,
               n, nrhs);
    LQ_CUDA(c, cudaSetDevice(c->device));
    return lstsq_hh_batched_stream(c, c->stream, A, B, batch, m, n, nrhs, X, info);
}

static int lstsq_hh_host(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n, int nrhs, double* X,
                         int32_t* info);

int lq_lstsq_householder_batched(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n, int nrhs,
                                 double* X) {
    return lstsq_hh_host(h, A, B, batch, m, n, nrhs, X, nullptr);
}
int lq_lstsq_householder_batched_info(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n, int nrhs,
                                      double* X, int32_t* info) {
    return lstsq_hh_host(h, A, B, batch, m, n, nrhs, X, info);
}

static int lstsq_hh_host(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n, int nrhs, double* X,
                         int32_t* info) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, m >= n && nrhs >= 1, LQ_ERR_SHAPE, "least squares needs m >= n and nrhs >= 1 (got %d x %d, %d rhs)", m,
               n, nrhs);
    const size_t mn = (size_t)m * n * 8, mk = (size_t)m * nrhs * 8, nk = (size_t)n * nrhs * 8;
    if (!fits_small(c, small_hh_smem_doubles(m, n, nrhs)) && !lstsq_stream_kernel_supported(m, n, nrhs) &&
        !lstsq_tile_kernel_supported(m, n, nrhs)) {
        LQ_CUDA(c, cudaSetDevice(c->device));
        for (int64_t b = 0; b < batch; ++b) {
            DevBuf dA, dB, dX;
            LQ_TRY(dA.alloc(c, mn));
            LQ_TRY(dB.alloc(c, mk));
            LQ_TRY(dX.alloc(c, nk));
            LQ_CUDA(c, cudaMemcpyAsync(dA.p, A + b * (size_t)m * n, mn, cudaMemcpyHostToDevice, c->stream));
            LQ_CUDA(c, cudaMemcpyAsync(dB.p, B + b * (size_t)m * nrhs, mk, cudaMemcpyHostToDevice, c->stream));
            LQ_TRY(large_lstsq_householder(c, dA.as<double>(), dB.as<double>(), m, n, nrhs, dX.as<double>()));
            LQ_CUDA(c, cudaMemcpyAsync(X + b * (size_t)n * nrhs, dX.p, nk, cudaMemcpyDeviceToHost, c->stream));
            LQ_CUDA(c, cudaStreamSynchronize(c->stream));
            if (info) info[b] = 0;
        }
        return LQ_OK;
    }
    return run_chunked(c, batch, mn, mk, nk, 0, info ? sizeof(int32_t) : 0, A, B, X, nullptr, info,
                       [&](cudaStream_t st, long long cnt, void* dA, void* dB, void* dX, void*, void* dI) {
                           return lstsq_hh_batched_stream(c, st, (const double*)dA, (const double*)dB, cnt, m, n, nrhs,
                                                          (double*)dX, info ? (int*)dI : nullptr);
                       });
}

int lq_lstsq_mgs_batched_dev(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n, int nrhs,
                             double* X, int32_t* info) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, nrhs >= 1, LQ_ERR_SHAPE, "nrhs must be >= 1");
    LQ_CUDA(c, cudaSetDevice(c->device));
    return lstsq_mgs_batched_stream(c, c->stream, A, B, batch, m, n, nrhs, X, info);
}

int lq_lstsq_mgs_batched(lq_ctx* h, const double* A, const double* B, int64_t batch, int m, int n, int nrhs, double* X,
                         int32_t* info) {
    Ctx* c = as_ctx(h);
    LQ_ARGS_QR(c, A, batch, m, n);
    LQ_REQUIRE(c, nrhs >= 1, LQ_ERR_SHAPE, "nrhs must be >= 1");
    const size_t mn = (size_t)m * n * 8, mk = (size_t)m * nrhs * 8, nk = (size_t)n * nrhs * 8;
    if (!fits_small(c, small_mgs_smem_doubles(m, n, nrhs)) && !lstsq_tile_kernel_supported(m, n, nrhs)) {
        LQ_CUDA(c, cudaSetDevice(c->device));
        for (int64_t b = 0; b < batch; ++b) {
            DevBuf dA, dB, dX, dI;
            LQ_TRY(dA.alloc(c, mn));
            LQ_TRY(dB.alloc(c, mk));
            LQ_TRY(dX.alloc(c, nk));
            LQ_TRY(dI.alloc(c, 16));
            LQ_CUDA(c, cudaMemcpyAsync(dA.p, A + b * (size_t)m * n, mn, cudaMemcpyHostToDevice, c->stream));
            LQ_CUDA(c, cudaMemcpyAsync(dB.p, B + b * (size_t)m * nrhs, mk, cudaMemcpyHostToDevice, c->stream));
            LQ_TRY(large_lstsq_mgs(c, dA.as<double>(), dB.as<double>(), m, n, nrhs, dX.as<double>(), dI.as<int>()));
            LQ_CUDA(c, cudaMemcpyAsync(X + b * (size_t)n * nrhs, dX.p, nk, cudaMemcpyDeviceToHost, c->stream));
            if (info) LQ_CUDA(c, cudaMemcpyAsync(info + b, dI.p, 4, cudaMemcpyDeviceToHost, c->stream));
            LQ_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        return LQ_OK;
    }
    return run_chunked(c, batch, mn, mk, nk, 0, sizeof(int32_t), A, B, X, nullptr, info,
                       [&](cudaStream_t st, long long cnt, void* dA, void* dB, void* dX, void*, void* dI) {
                           return lstsq_mgs_batched_stream(c, st, (const double*)dA, (const double*)dB, cnt, m, n, nrhs,
                                                           (double*)dX, (int*)dI);
                       });
}

}  // extern "C"
