// K3 launch glue: streaming Householder least squares (stream_qr.cuh) for n <= 64, nrhs <= 32.
#include "../../include/linalg_b200.h"
#include "ops.cuh"
#include "stream_qr.cuh"

namespace lq {

bool lstsq_stream_kernel_supported(int m, int n, int nrhs) {
    if (getenv("LINALG_B200_NO_STREAM_LSTSQ")) return false;
    return n >= 1 && n <= 64 && nrhs >= 1 && nrhs <= 32 && m >= n;
}

template <int RPT, int WARPS>
static int launch_variant(Ctx* c, cudaStream_t st, const double* A, const double* B, long long batch, int m, int n, int nrhs,
                          double* X) {
    using Cfg = StreamCfg<3, RPT, WARPS>;
    auto kern = lstsq_stream_kernel<RPT, WARPS>;
    const size_t smem = Cfg::smem_doubles(n) * sizeof(double);
    static DeviceLatch configured;
    if (!configured.test(c->device)) {
        LQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
        configured.set(c->device);
    }
    kern<<<(unsigned)batch, WARPS * 32, smem, st>>>(A, B, X, m, n, nrhs);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

// rows per streamed block = RPT * WARPS.  Taller blocks mean fewer column steps per system (every step pays the
// publish -> dots -> barrier -> norm / reciprocal chain once, whatever the block height).
int lstsq_stream_kernel_launch(Ctx* c, cudaStream_t st, const double* A, const double* B, long long batch, int m, int n,
                               int nrhs, double* X) {
    if (!lstsq_stream_kernel_supported(m, n, nrhs)) return LQ_ERR_UNSUPPORTED;
    static const int variant = getenv("LINALG_B200_LSTSQ_VARIANT") ? atoi(getenv("LINALG_B200_LSTSQ_VARIANT")) : -1;
    int v = variant;
    if (v < 0) v = 0;
    switch (v) {
        case 1: return launch_variant<16, 8>(c, st, A, B, batch, m, n, nrhs, X);
        case 2: return launch_variant<8, 8>(c, st, A, B, batch, m, n, nrhs, X);
        case 3: return launch_variant<16, 16>(c, st, A, B, batch, m, n, nrhs, X);
        case 4: return launch_variant<8, 16>(c, st, A, B, batch, m, n, nrhs, X);
        default: return launch_variant<16, 4>(c, st, A, B, batch, m, n, nrhs, X);
    }
}

}  // namespace lq
