// K2 (round 2): batched 32x32 modified Gram-Schmidt QR, "one lane = one column" form (the layout of batched_qr32_c8.cuh).
//
// Reference semantics: linalg/qr.py:14-49 (qr) for every A[b] of a (batch, 32, 32) array -- left-looking MGS,
// R[k, j] = q_k . v, v -= R[k, j] q_k for k < j in increasing k, R[j, j] = ||v|| > 0, "linearly dependent" when
// R[j, j] < 1e-12 (reported per matrix in info[], 1-based column), reorth=True = a second sweep over Q whose R is returned.
//
// Four matrices per warp, eight lanes each; lane c of a group owns column 8 p + c of the current 8-column panel with all 32
// rows in registers.  A panel is loaded, the finished vectors of the earlier panels are projected out (vector broadcast
// from shared memory; dot product and update are lane-local: no shuffles, exactly the reference's order of operations per
// column), then its eight columns are orthogonalised against each other (one 64-bit shuffle per column for ||v||^2).
// Vectors are kept UNNORMALISED in shared memory, their 1 / ||v|| in the owner lanes' registers (no second publish; the projection
// coefficient is (v_k . a) / ||v_k||^2), Q is scaled when it is stored.  There is no separate Q phase: the round-1 kernel
// (two matrices per warp, 2-D lane layout) needs 2.4 x the shuffles and selects per matrix.
#pragma once

#include "batched_qr32.cuh"

namespace lq {

struct MgsCol8 {
    static constexpr int MAT = 1026;   // 32 vectors x 32 rows + 2: 513 16-byte words = 1 (mod 8), the four matrices of a warp read
                                       // four different 16-byte bank groups with one broadcast LDS.128
    static constexpr int WARP_DOUBLES = 4 * MAT;   // 32.8 KB: seven warps per SM (the reciprocal norms live in registers)
};

template <int WARPS, int MINB, bool KEEPV>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    mgs_qr32_c8_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, int* __restrict__ info,
                       long long batch, int reorth) {
    constexpr int N = 32;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g4 = lane >> 3, c = lane & 7;
    double* wbase = smem + (size_t)warp * MgsCol8::WARP_DOUBLES;
    double* vb = wbase + g4 * MgsCol8::MAT;            // vector k at vb + 32 k
    const long long mat = ((long long)blockIdx.x * WARPS + warp) * 4 + g4;
    const bool valid = mat < batch;
    const long long matc = valid ? mat : (batch - 1);
    int bad = 0;
    double myrinv[4] = {1.0, 1.0, 1.0, 1.0};  // 1 / ||v|| of my column in each panel (lane c holds columns c, 8 + c, 16 + c, 24 + c)

    const int nsweep = reorth ? 2 : 1;
#pragma unroll 1
    for (int sweep = 0; sweep < nsweep; ++sweep) {
        const bool last = sweep == nsweep - 1;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int col = 8 * p + c;
            double a[N], rr[N];
            if (sweep == 0) {
                const double* Ag = A + matc * (N * N) + col;
#pragma unroll
                for (int i = 0; i < N; ++i) a[i] = ld_stream(Ag + i * N);
            } else {
                // second sweep (linalg/qr.py:46-47): the input is the first sweep's Q = v / ||v||, still in shared memory
                const double rs = myrinv[p];
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    const double2 vv = *reinterpret_cast<const double2*>(vb + 32 * col + i);
                    a[i] = vv.x * rs;
                    a[i + 1] = vv.y * rs;
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < N; ++k) rr[k] = 0.0;

            // ---- project out the vectors of the earlier panels: r = (v_k . a) / ||v_k||, a -= (r / ||v_k||) v_k
#pragma unroll
            for (int k = 0; k < 8 * p; ++k) {
                const double* vk_ = vb + 32 * k;
                double d[4] = {0.0, 0.0, 0.0, 0.0};
                double vk[KEEPV ? N : 2];
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    const double2 vv = *reinterpret_cast<const double2*>(vk_ + i);
                    if (KEEPV) vk[i] = vv.x, vk[i + 1] = vv.y;
                    d[(i >> 1) & 1] = fma(vv.x, a[i], d[(i >> 1) & 1]);
                    d[2 + ((i >> 1) & 1)] = fma(vv.y, a[i + 1], d[2 + ((i >> 1) & 1)]);
                }
                const double ri = __shfl_sync(0xffffffffu, myrinv[k >> 3], k & 7, 8);
                const double r = ((d[0] + d[1]) + (d[2] + d[3])) * ri;
                rr[k] = r;
                const double s = r * ri;
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    double2 vv;
                    if (KEEPV) vv = make_double2(vk[i], vk[i + 1]);
                    else vv = *reinterpret_cast<const double2*>(vk_ + i);
                    a[i] = fma(-s, vv.x, a[i]);
                    a[i + 1] = fma(-s, vv.y, a[i + 1]);
                }
            }

            // ---- the 8 columns of this panel against each other
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = 8 * p + jj;
                double* vj = vb + 32 * j;
                if (c == jj) {
#pragma unroll
                    for (int i = 0; i < N; i += 2) *reinterpret_cast<double2*>(vj + i) = make_double2(a[i], a[i + 1]);
                }
                __syncwarp();
                double d[4] = {0.0, 0.0, 0.0, 0.0};
                double vk[KEEPV ? N : 2];
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    const double2 vv = *reinterpret_cast<const double2*>(vj + i);
                    if (KEEPV) vk[i] = vv.x, vk[i + 1] = vv.y;
                    d[(i >> 1) & 1] = fma(vv.x, a[i], d[(i >> 1) & 1]);
                    d[2 + ((i >> 1) & 1)] = fma(vv.y, a[i + 1], d[2 + ((i >> 1) & 1)]);
                }
                const double dot = (d[0] + d[1]) + (d[2] + d[3]);            // v_j . a_c
                const double ss = __shfl_sync(0xffffffffu, dot, jj, 8);      // ||v_j||^2 = the owner column's own dot product
                double rinv;
                const double nrm = sqrt_nr_t<2>(fmax(ss, 1e-300), rinv);
                if (nrm < kEps && bad == 0) bad = j + 1;                     // linalg/qr.py:40-41
                if (c == jj) myrinv[p] = rinv;
                const double r = (c > jj) ? dot * rinv : ((c == jj) ? nrm : 0.0);
                rr[j] = (c >= jj) ? r : rr[j];
                const double s = (c > jj) ? r * rinv : 0.0;
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    double2 vv;
                    if (KEEPV) vv = make_double2(vk[i], vk[i + 1]);
                    else vv = *reinterpret_cast<const double2*>(vj + i);
                    a[i] = fma(-s, vv.x, a[i]);
                    a[i + 1] = fma(-s, vv.y, a[i + 1]);
                }
            }
            __syncwarp();

            // ---- store Q (scaled) and the R column of this panel (rows below the diagonal are exact zeros)
            if (valid && last) {
                double* Qg = Q + mat * (N * N) + col;
                double* Rg = R + mat * (N * N) + col;
#pragma unroll
                for (int i = 0; i < N; ++i) st_stream(Qg + i * N, a[i] * myrinv[p]);
#pragma unroll
                for (int i = 0; i < N; ++i) st_stream(Rg + i * N, (i <= col) ? rr[i] : 0.0);
            }
        }
    }
    if (valid && info != nullptr && c == 0) info[mat] = bad;
}

}  // namespace lq
