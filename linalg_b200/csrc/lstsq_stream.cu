// K3 launch glue: streaming Householder least squares (stream_qr.cuh) for n <= 64, nrhs <= 32.
#include "../../include/linalg_b200.h"
#include "ops.cuh"
#include "stream_qr.cuh"

namespace lq {

bool lstsq_stream_kernel_supported(int m, int n, int nrhs) {
    if (getenv("LINALG_B200_NO_STREAM_LSTSQ")) return false;
    return n >= 1 && n <= 64 && nrhs >= 1 && nrhs <= 32 && m >= n;
}

int lstsq_stream_kernel_launch(Ctx* c, cudaStream_t st, const double* A, const double* B, long long batch, int m, int n,
                               int nrhs, double* X) {
    if (!lstsq_stream_kernel_supported(m, n, nrhs)) return LQ_ERR_UNSUPPORTED;
    constexpr int RPT = 16, WARPS = 4;
    using Cfg = StreamCfg<3, RPT, WARPS>;
    auto kern = lstsq_stream_kernel<RPT, WARPS>;
    const size_t smem = Cfg::smem_doubles(n) * sizeof(double);
    static bool configured[64] = {};
    if (!configured[c->device]) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
        configured[c->device] = true;
    }
    kern<<<(unsigned)batch, WARPS * 32, smem, st>>>(A, B, X, m, n, nrhs);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

}  // namespace lq
