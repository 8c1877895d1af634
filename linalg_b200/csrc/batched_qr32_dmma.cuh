// K1 (round 2): batched 32x32 Householder QR with the Q formation on the FP64 tensor pipe.
//
// Reference semantics: linalg/qr.py:52-100 (householder_qr) for every A[b] of a (batch, 32, 32) array.
//
// Why: the round-1 kernel (batched_qr32.cuh) is bound by instruction issue, not by HBM or the FP64 datapath: a
// DFMA holds a scheduler's dispatch port for 2 cycles, every select / shuffle / LDS for 1, and the Q phase alone
// issued ~1 930 FP64 + ~450 other instructions per matrix.  A DMMA.8x8x4 (the only f64 MMA shape sm_100a has; every
// mma.sync f64 shape lowers to it) does 256 FMAs for ONE issue slot, so the Q phase is re-formulated as compact-WY
// block reflectors on 8x8 tiles:
//
//   R phase  unchanged in structure (two matrices per warp, 16 lanes each, rank-1 DFMA updates, scalar chain per
//            column); the reflectors v_j (pivot element patched to v0), and beta_j stay in shared memory.
//   Q phase  per matrix, all 32 lanes: Q^T lives in 16 accumulator tiles (32 doubles / lane).
//            G_p = V_p^T V_p (4 panels of 8 reflectors)                         20 DMMA
//            T_p by the dlarft row recurrence, lane (g, t) = row g of panel t    ~36 DFMA for all four panels at once
//            for p = 3..0:  W^T = Q^T V_p ; W2^T = W^T (-T_p)^T ; Q^T += W2^T V_p^T   28 + 20 + 60 DMMA
//            An accumulator fragment IS a valid A operand when the contraction index is permuted (k = t <-> column
//            2t+i), so W^T and W2^T never leave registers; V is read from shared memory directly in B-fragment order.
//
// Tile / fragment conventions (lane l, g = l >> 2, t = l & 3), mma.m8n8k4.f64:
//   A[g][t], B[t][g], C[g][2t], C[g][2t+1].
//   qt[cb][rb][i] = Q[8 rb + 2t + i][8 cb + g]          (tile (cb, rb) of Q^T)
//   F1(p, rb, i)  = V[8 rb + 2t + i][8 p + g]           (B operand of W^T = Q^T V, A and B operand of the Gram)
//   F3(p, rb, i)  = V[8 rb + g][8 p + 2t + i]           (B operand of the update; the panel's own W^T tile)
#pragma once

#include "batched_qr32.cuh"

namespace lq {

struct Dmma32 {
    using D = Dist32<2, 4>;
    static constexpr int SCRATCH = 512;  // per warp: G (4 x 64) and -T (4 x 64)
    static constexpr int warp_doubles() { return D::MPW * D::SMEM_DOUBLES + SCRATCH; }
};

// PHASES: 3 = product; 1 = R phase only, 2 = Q phase only (timing diagnostics: the other output is garbage)
template <int WARPS, int MINB, int PHASES = 3>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    hh_qr32_dmma_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, long long batch) {
    using D = Dist32<2, 4>;
    constexpr int P = 2, C = 4;
    constexpr int N = 32, RPL = D::RPL, LC = D::LC, L = D::L;
    constexpr int ROWP = D::ROWP, PSTRIDE = D::PSTRIDE;
    extern __shared__ __align__(16) double smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wbase = smem + (size_t)warp * Dmma32::warp_doubles();
    const long long mat0 = ((long long)blockIdx.x * WARPS + warp) * D::MPW;

    if (PHASES & 1) {
        // ================= R phase (two matrices per warp, 16 lanes each) =================
        const int g = lane / L, lm = lane % L, p = lm / LC, lc = lm % LC;
        const long long mat = mat0 + g;
        const bool valid = mat < batch;
        const long long matc = valid ? mat : (batch - 1);
        double* vb = wbase + (size_t)g * D::SMEM_DOUBLES;
        double* betas = vb + N * ROWP;

        int colv[C];
#pragma unroll
        for (int s = 0; s < C; ++s) colv[s] = D::col(s, lc);

        double r[C][RPL];
        {
            const double* Ag = A + matc * (N * N) + p * N;
#pragma unroll
            for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
                for (int s = 0; s < C; ++s) r[s][ii] = ld_stream(Ag + ii * (P * N) + colv[s]);
        }

#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int so = D::owner_slot(j), lo = D::owner_lc(j);
            const int iib = j / P, jp = j % P;
            const int ii0 = iib & ~1;
            double* vj = vb + j * ROWP + p * PSTRIDE;

            if (lc == lo) {
#pragma unroll
                for (int ii = ii0; ii < RPL; ii += 2)
                    *reinterpret_cast<double2*>(vj + ii) = make_double2(r[so][ii], r[so][ii + 1]);
            }
            __syncwarp();

            double d[C], d2[C];
            double vkeep[RPL];
#pragma unroll
            for (int s = 0; s < C; ++s) d[s] = 0.0, d2[s] = 0.0;
#pragma unroll
            for (int ii = ii0; ii < RPL; ii += 2) {
                double2 vv = *reinterpret_cast<const double2*>(vj + ii);
                if (ii < iib) vv.x = 0.0;
                if (ii == iib) vv.x = (p >= jp) ? vv.x : 0.0;
                if (ii + 1 == iib) vv.y = (p >= jp) ? vv.y : 0.0;
                vkeep[ii] = vv.x;
                vkeep[ii + 1] = vv.y;
#pragma unroll
                for (int s = so; s < C; ++s) {
                    d[s] = fma(vv.x, r[s][ii], d[s]);
                    if (C - so >= 3) d[s] = fma(vv.y, r[s][ii + 1], d[s]);
                    else d2[s] = fma(vv.y, r[s][ii + 1], d2[s]);
                }
            }
            if (C - so < 3) {
#pragma unroll
                for (int s = so; s < C; ++s) d[s] += d2[s];
            }
            double ss = group_sum<P, C>(d[so]);
            ss = __shfl_sync(0xffffffffu, ss, lo, L);
            const double x0 = vb[j * ROWP + jp * PSTRIDE + iib];

            const double ssc = fmax(ss, 1e-300);
            // y = 1/||x||, beta = 2 / v^T v = y^2 / (1 + |x0| y); the reciprocal is seeded from the UNREFINED y
            const double ax0 = fabs(x0);
            double y = rsqrt_seed(ssc);
            double u = rcp_seed(fma(ax0, y, 1.0));
            const double hx = 0.5 * ssc;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const double e = fma(-hx * y, y, 0.5);
                y = fma(y, e, y);
            }
            const double nrm = ssc * y;
            const double Dn = fma(ax0, y, 1.0);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const double e = fma(-Dn, u, 1.0);
                u = fma(u, e, u);
            }
            double beta = (y * y) * u;
            const bool skip = nrm < kEps;  // qr.py:79-80
            const double alpha = copysign(nrm, x0);
            const double v0 = x0 + alpha;
            beta = skip ? 0.0 : beta;
            __syncwarp();  // every lane has read x0 before the pivot slot is patched for the Q phase
            if (lm == 0) {
                betas[j] = beta;
                vb[j * ROWP + jp * PSTRIDE + iib] = v0;
            }
            const bool piv = (p == jp);
            const double alpha_m = piv ? alpha : 0.0;

#pragma unroll
            for (int s = so; s < C; ++s) {
                const double part = group_sum<P, C>(fma(alpha_m, r[s][iib], d[s]));
                d[s] = beta * part;
            }
#pragma unroll
            for (int ii = ii0; ii < RPL; ii += 2) {
                double2 vv;
                vv.x = vkeep[ii];
                vv.y = vkeep[ii + 1];
                if (ii == iib) vv.x = piv ? v0 : vv.x;
                if (ii + 1 == iib) vv.y = piv ? v0 : vv.y;
#pragma unroll
                for (int s = so; s < C; ++s) {
                    r[s][ii] = fma(-d[s], vv.x, r[s][ii]);
                    r[s][ii + 1] = fma(-d[s], vv.y, r[s][ii + 1]);
                }
            }
            if (lc == lo && piv && !skip) r[so][iib] = -alpha;
        }

        if (valid) {
            double* Rg = R + mat * (N * N) + p * N;
#pragma unroll
            for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
                for (int s = 0; s < C; ++s) {
                    const int i = P * ii + p;
                    st_stream(Rg + ii * (P * N) + colv[s], (colv[s] >= i) ? r[s][ii] : 0.0);
                }
        }
    }
    __syncwarp();

    // ================= Q phase: Q = (I - V0 T0 V0^T) ... (I - V3 T3 V3^T), backward accumulation on DMMA ==========
    const int gq = lane >> 2, tq = lane & 3;
    double* Gs = wbase + D::MPW * D::SMEM_DOUBLES;
    double* Ts = Gs + 256;
    const bool m1 = gq >= 2 * tq, m1b = gq >= 2 * tq + 1;   // F3 diagonal-tile masks (row g >= column 2t+i)
    const bool m0 = 2 * tq >= gq, m0b = 2 * tq + 1 >= gq;   // F1 diagonal-tile masks (row 2t+i >= column g)

    if (PHASES & 2)
#pragma unroll 1
    for (int mi = 0; mi < D::MPW; ++mi) {
        const double* vb = wbase + (size_t)mi * D::SMEM_DOUBLES;
        const double* betas = vb + N * ROWP;
        const double* f1base = vb + gq * ROWP + tq;                          // + 8p*ROWP + i*PSTRIDE + 4rb
        const double* f3base = vb + (2 * tq) * ROWP + (gq & 1) * PSTRIDE + (gq >> 1);  // + (8p+i)*ROWP + 4rb

        auto F1 = [&](int p, int rb, int i) -> double {
            double v = f1base[8 * p * ROWP + i * PSTRIDE + 4 * rb];
            if (rb == p) v = (i ? m0b : m0) ? v : 0.0;
            return v;
        };
        auto F3 = [&](int p, int rb, int i) -> double {
            double v = f3base[(8 * p + i) * ROWP + 4 * rb];
            if (rb == p) v = (i ? m1b : m1) ? v : 0.0;
            return v;
        };

        // ---- Gram matrices of the four panels
        double G[4][2];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            G[p][0] = G[p][1] = 0.0;
#pragma unroll
            for (int rb = p; rb < 4; ++rb)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double f = F1(p, rb, i);
                    dmma_8x8x4(G[p], f, f);
                }
        }
        __syncwarp();  // previous matrix's readers of Gs / Ts are done
#pragma unroll
        for (int p = 0; p < 4; ++p) *reinterpret_cast<double2*>(Gs + p * 64 + gq * 8 + 2 * tq) = make_double2(G[p][0], G[p][1]);
        __syncwarp();

        // ---- T of panel tq, row gq (dlarft, forward / columnwise): T[g][k] = -beta_k sum_{m=g}^{k-1} T[g][m] G[m][k]
        {
            double Trow[8], bk[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const double2 b2 = *reinterpret_cast<const double2*>(betas + 8 * tq + k);
                bk[k] = b2.x;
                bk[k + 1] = b2.y;
            }
            double acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                Trow[k] = (k == gq) ? bk[k] : ((k > gq) ? -bk[k] * acc[k] : 0.0);
                // push row k of G into the pending sums of the later columns
                if (k < 7) {
#pragma unroll
                    for (int kk = (k + 1) & ~1; kk < 8; kk += 2) {
                        const double2 g2 = *reinterpret_cast<const double2*>(Gs + tq * 64 + k * 8 + kk);
                        if (kk > k) acc[kk] = fma(Trow[k], g2.x, acc[kk]);
                        acc[kk + 1] = fma(Trow[k], g2.y, acc[kk + 1]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k += 2)
                *reinterpret_cast<double2*>(Ts + tq * 64 + gq * 8 + k) = make_double2(-Trow[k], -Trow[k + 1]);
        }
        __syncwarp();

        // ---- backward accumulation
        double qt[4][4][2];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                qt[cb][rb][0] = (cb == rb && 2 * tq == gq) ? 1.0 : 0.0;
                qt[cb][rb][1] = (cb == rb && 2 * tq + 1 == gq) ? 1.0 : 0.0;
            }

#pragma unroll
        for (int p = 3; p >= 0; --p) {
            double f3[4][2], f1[4][2], w[4][2], w2[4][2];
#pragma unroll
            for (int rb = p; rb < 4; ++rb) {
                f3[rb][0] = F3(p, rb, 0);
                f3[rb][1] = F3(p, rb, 1);
                if (rb > p) {
                    f1[rb][0] = F1(p, rb, 0);
                    f1[rb][1] = F1(p, rb, 1);
                }
            }
            const double2 tt = *reinterpret_cast<const double2*>(Ts + p * 64 + gq * 8 + 2 * tq);
            w[p][0] = f3[p][0];
            w[p][1] = f3[p][1];
#pragma unroll
            for (int cb = p + 1; cb < 4; ++cb) w[cb][0] = w[cb][1] = 0.0;
#pragma unroll
            for (int rb = p + 1; rb < 4; ++rb)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int cb = p + 1; cb < 4; ++cb) dmma_8x8x4(w[cb], qt[cb][rb][i], f1[rb][i]);
#pragma unroll
            for (int cb = p; cb < 4; ++cb) w2[cb][0] = w2[cb][1] = 0.0;
#pragma unroll
            for (int cb = p; cb < 4; ++cb) dmma_8x8x4(w2[cb], w[cb][0], tt.x);
#pragma unroll
            for (int cb = p; cb < 4; ++cb) dmma_8x8x4(w2[cb], w[cb][1], tt.y);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int cb = p; cb < 4; ++cb)
#pragma unroll
                    for (int rb = p; rb < 4; ++rb) dmma_8x8x4(qt[cb][rb], w2[cb][i], f3[rb][i]);
        }

        const long long mat = mat0 + mi;
        if (mat < batch) {
            double* Qg = Q + mat * (N * N) + (2 * tq) * N + gq;
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) st_stream(Qg + (8 * rb + i) * N + 8 * cb, qt[cb][rb][i]);
        }
    }
}

}  // namespace lq
