// K1 (round 2, default): batched 32x32 Householder QR, register-light left-looking form.
//
// Reference semantics: linalg/qr.py:52-100 (householder_qr) for every A[b] of a (batch, 32, 32) array.
//
// What round 1 measured (profiles/r2_hh32_notes.md): the one-shot kernel keeps the whole matrix pair in registers
// (255 registers, 8 warps / SM) and every column step is a ~400-cycle dependent chain (publish -> dot -> shuffle ->
// rsqrt / reciprocal Newton -> shuffle -> update), so with two warps per scheduler the R phase alone runs at 30 % issue
// utilisation and takes 75 % of the kernel.  The cure is more resident warps, i.e. fewer registers per matrix:
//
//   R phase  left-looking over NS column stages (NS = 2: 16 columns = 64 registers per matrix pair).  A stage loads
//            its columns, applies the finished reflectors of the earlier stages (vectors and betas come from shared
//            memory; no scalar chain in those steps), then factors its own columns exactly like round 1 (two
//            matrices per warp, 16 lanes each, rank-1 DFMA updates).  Same flops as right-looking, same operation
//            order per column.  Shorter critical path per step: every partial dot is reduced over the row-parity
//            pair BEFORE the scalar chain (the pivot-row element is pre-broadcast), the dot products use two
//            accumulators, the reflector is re-read from shared memory in the update instead of kept in registers.
//   storage  reflectors packed (rows >= 4*(j/4) only): 592 doubles per matrix instead of 1 152, pivot element
//            patched to v0 and the rows above the pivot zeroed after the scalar chain (off the critical path), so
//            later stages and the Q phase read them without masks.
//   Q phase  compact-WY on DMMA.8x8x4 as in batched_qr32_dmma.cuh (Q^T in 16 accumulator tiles, ~100 registers).
//
// => <= 128 registers, 16 warps / SM, 12.2 KB of shared memory per warp.
#pragma once

#include "batched_qr32.cuh"

namespace lq {

// packed reflector storage of one matrix: vector j holds rows 2*ii0(j) .. 31 as [parity][ii - ii0]
struct Pack32 {
    __host__ __device__ static constexpr int ii0(int j) { return 2 * (j >> 2); }
    __host__ __device__ static constexpr int plen(int j) { return 16 - 2 * (j >> 2); }
    // the two parity sub-rows must not start in the same or the opposite bank quad (LDS.128 broadcast of the four
    // (matrix, parity) groups of a warp): plen/2 mod 8 is 0 for j < 4 and 4 for 16 <= j < 20 -> pad those by one word
    __host__ __device__ static constexpr int pad(int j) { return ((j >> 2) == 0 || (j >> 2) == 4) ? 2 : 0; }
    __host__ __device__ static constexpr int pstr(int j) { return plen(j) + pad(j); }
    __host__ __device__ static constexpr int vsize(int j) { return 2 * plen(j) + pad(j); }
    __host__ __device__ static constexpr int voff(int j) {  // sum of vsize(t), t < j, in closed form
        return 128 * (j >> 2) - 8 * (j >> 2) * ((j >> 2) - 1) + ((j >> 2) > 0 ? 8 : 0) + ((j >> 2) > 4 ? 8 : 0) + (j & 3) * vsize(j);
    }
    __host__ __device__ static constexpr int voff_slow(int j) {
        int o = 0;
        for (int t = 0; t < j; ++t) o += vsize(t);
        return o;
    }
    static constexpr int VDOUBLES = 592;
    static constexpr int BETA = VDOUBLES;  // betas[32]
    static constexpr int MAT = 632;        // matrix stride: 316 16-byte words = 4 (mod 8): the two matrices of a warp never collide
    static constexpr int SCRATCH = 256;    // per warp: G (4 x 64), then overwritten by -T (4 x 64)
    static constexpr int WARP_DOUBLES = 2 * MAT + SCRATCH;
    // run-time offset (Q phase: the reflector index depends on the lane)
    __device__ static __forceinline__ int voff_rt(int k) { return voff(k); }
    __device__ static __forceinline__ int pstr_rt(int k) { return pstr(k); }
};
static_assert(Pack32::voff_slow(32) == Pack32::VDOUBLES, "packed size");
static_assert(Pack32::voff(5) == Pack32::voff_slow(5) && Pack32::voff(17) == Pack32::voff_slow(17) && Pack32::voff(31) == Pack32::voff_slow(31) &&
                  Pack32::voff(22) == Pack32::voff_slow(22) && Pack32::voff(3) == Pack32::voff_slow(3),
              "closed-form offsets");

__device__ __forceinline__ double2 lds128(const double* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)));
    return v;
}

// NS column stages (1, 2 or 4); PHASES as in hh_qr32_dmma_kernel (3 = product)
// KEEPV: the broadcast reflector stays in registers between the dot products and the update (half the LDS traffic,
// +32 registers) instead of being re-read from shared memory.
template <int NS, int WARPS, int MINB, int PHASES = 3, bool KEEPV = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    hh_qr32_ll_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, long long batch) {
    constexpr int N = 32, P = 2, LC = 8, L = 16, RPL = 16;
    constexpr int CS = 4 / NS, W = N / NS;  // column slots / columns per stage
    extern __shared__ __align__(16) double smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wbase = smem + (size_t)warp * Pack32::WARP_DOUBLES;
    const long long mat0 = ((long long)blockIdx.x * WARPS + warp) * 2;

    if (PHASES & 1) {
        // ================= R phase (two matrices per warp, 16 lanes each) =================
        const int g = lane >> 4, lm = lane & 15, p = lm >> 3, lc = lm & 7;
        const long long mat = mat0 + g;
        const bool valid = mat < batch;
        const long long matc = valid ? mat : (batch - 1);
        double* vb = wbase + g * Pack32::MAT;
        double* betas = vb + Pack32::BETA;
        const int lane_p0 = lane & ~8, lane_p1 = lane | 8;  // the two row-parity lanes of my column

#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int colv[CS];
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) colv[cs] = s * W + ((cs & 1) ? ((cs + 1) * LC - 1 - lc) : (cs * LC + lc));

            double r[CS][RPL];
            {
                const double* Ag = A + matc * (N * N) + p * N;
#pragma unroll
                for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) r[cs][ii] = ld_stream(Ag + ii * (P * N) + colv[cs]);
            }

            // ---- apply the reflectors of the earlier stages (masked + pivot-patched in shared memory)
#pragma unroll
            for (int j = 0; j < s * W; ++j) {
                const int ii0 = Pack32::ii0(j);
                const double* vj = vb + Pack32::voff(j) + p * Pack32::pstr(j) - ii0;
                double d[CS], d2[CS];
                double vk[KEEPV ? RPL : 2];
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) d[cs] = 0.0, d2[cs] = 0.0;
#pragma unroll
                for (int ii = ii0; ii < RPL; ii += 2) {
                    const double2 vv = *reinterpret_cast<const double2*>(vj + ii);
                    if (KEEPV) vk[ii] = vv.x, vk[ii + 1] = vv.y;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        d[cs] = fma(vv.x, r[cs][ii], d[cs]);
                        d2[cs] = fma(vv.y, r[cs][ii + 1], d2[cs]);
                    }
                }
                const double beta = betas[j];
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    double t = d[cs] + d2[cs];
                    t += __shfl_xor_sync(0xffffffffu, t, LC);
                    d[cs] = beta * t;
                }
#pragma unroll
                for (int ii = ii0; ii < RPL; ii += 2) {
                    const double2 vv = KEEPV ? make_double2(vk[ii], vk[ii + 1]) : lds128(vj + ii);
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        r[cs][ii] = fma(-d[cs], vv.x, r[cs][ii]);
                        r[cs][ii + 1] = fma(-d[cs], vv.y, r[cs][ii + 1]);
                    }
                }
            }

            // ---- factor the columns of this stage
#pragma unroll
            for (int jj = 0; jj < W; ++jj) {
                const int j = s * W + jj;
                const int so = jj / LC;
                const int lo = (so & 1) ? ((so + 1) * LC - 1 - jj) : (jj - so * LC);
                const int iib = j / P, jp = j % P;
                const int ii0 = Pack32::ii0(j);  // == iib & ~1
                double* vj = vb + Pack32::voff(j) + p * Pack32::pstr(j) - ii0;
                const bool piv = (p == jp);

                // pivot-row elements of the active columns, made known to both row-parity lanes ahead of the chain
                double ajc[CS];
#pragma unroll
                for (int cs = so; cs < CS; ++cs) ajc[cs] = __shfl_sync(0xffffffffu, r[cs][iib], jp ? lane_p1 : lane_p0);

                if (lc == lo) {
#pragma unroll
                    for (int ii = ii0; ii < RPL; ii += 2)
                        *reinterpret_cast<double2*>(vj + ii) = make_double2(r[so][ii], r[so][ii + 1]);
                }
                __syncwarp();

                // first row pair (rows 4*(j/4) .. 4*(j/4)+3): mask the rows above the pivot
                double2 vf = *reinterpret_cast<const double2*>(vj + ii0);
                if (ii0 < iib) vf.x = 0.0;
                if (ii0 == iib) vf.x = (p >= jp) ? vf.x : 0.0;
                if (ii0 + 1 == iib) vf.y = (p >= jp) ? vf.y : 0.0;

                double d[CS], d2[CS];
                double vk[KEEPV ? RPL : 2];
#pragma unroll
                for (int cs = so; cs < CS; ++cs) {
                    d[cs] = vf.x * r[cs][ii0];
                    d2[cs] = vf.y * r[cs][ii0 + 1];
                }
#pragma unroll
                for (int ii = ii0 + 2; ii < RPL; ii += 2) {
                    const double2 vv = *reinterpret_cast<const double2*>(vj + ii);
                    if (KEEPV) vk[ii] = vv.x, vk[ii + 1] = vv.y;
#pragma unroll
                    for (int cs = so; cs < CS; ++cs) {
                        d[cs] = fma(vv.x, r[cs][ii], d[cs]);
                        d2[cs] = fma(vv.y, r[cs][ii + 1], d2[cs]);
                    }
                }
                // x^T a_c, complete in both parity lanes
#pragma unroll
                for (int cs = so; cs < CS; ++cs) {
                    const double t = d[cs] + d2[cs];
                    d[cs] = t + __shfl_xor_sync(0xffffffffu, t, LC);
                }
                const double ss = __shfl_sync(0xffffffffu, d[so], lo, L);  // ||x||^2 = the owner column's own dot product
                const double x0 = vb[Pack32::voff(j) + jp * Pack32::pstr(j) + iib - ii0];

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

                // v^T a_c = x^T a_c + alpha * a_jc
#pragma unroll
                for (int cs = so; cs < CS; ++cs) d[cs] = beta * fma(alpha, ajc[cs], d[cs]);

                // first pair with the pivot element, stored back for the later stages and the Q phase
                if (ii0 == iib) vf.x = piv ? v0 : vf.x;
                if (ii0 + 1 == iib) vf.y = piv ? v0 : vf.y;
                __syncwarp();  // every lane has read the raw first pair / x0
                if (lc == lo) *reinterpret_cast<double2*>(vj + ii0) = vf;
                if (lm == 0) betas[j] = beta;
#pragma unroll
                for (int cs = so; cs < CS; ++cs) {
                    r[cs][ii0] = fma(-d[cs], vf.x, r[cs][ii0]);
                    r[cs][ii0 + 1] = fma(-d[cs], vf.y, r[cs][ii0 + 1]);
                }
#pragma unroll
                for (int ii = ii0 + 2; ii < RPL; ii += 2) {
                    const double2 vv = KEEPV ? make_double2(vk[ii], vk[ii + 1]) : lds128(vj + ii);
#pragma unroll
                    for (int cs = so; cs < CS; ++cs) {
                        r[cs][ii] = fma(-d[cs], vv.x, r[cs][ii]);
                        r[cs][ii + 1] = fma(-d[cs], vv.y, r[cs][ii + 1]);
                    }
                }
                // exact diagonal for the owner (mathematically the update already gives -alpha)
                if (lc == lo && piv && !skip) r[so][iib] = -alpha;
            }

            // ---- store the R columns of this stage (strict lower triangle forced to exact zeros, qr.py:97)
            if (valid) {
                double* Rg = R + mat * (N * N) + p * N;
#pragma unroll
                for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int i = P * ii + p;
                        st_stream(Rg + ii * (P * N) + colv[cs], (colv[cs] >= i) ? r[cs][ii] : 0.0);
                    }
            }
        }
    }
    __syncwarp();

    // ================= Q phase: Q = (I - V0 T0 V0^T) ... (I - V3 T3 V3^T), backward accumulation on DMMA ==========
    // Fragment conventions (lane l, g = l >> 2, t = l & 3) as in batched_qr32_dmma.cuh.
    const int gq = lane >> 2, tq = lane & 3;
    double* Gs = wbase + 2 * Pack32::MAT;
    const bool m1 = gq >= 2 * tq, m1b = gq >= 2 * tq + 1;   // F3 diagonal-tile masks (row g >= column 2t+i)
    const bool m0 = 2 * tq >= gq, m0b = 2 * tq + 1 >= gq;   // F1 diagonal-tile masks (row 2t+i >= column g)

    if (PHASES & 2)
#pragma unroll 1
    for (int mi = 0; mi < 2; ++mi) {
        const double* vb = wbase + mi * Pack32::MAT;
        const double* betas = vb + Pack32::BETA;
        // V[r][k] lives at voff(k) + (r & 1) * pstr(k) + (r >> 1) - 2 * (k >> 2)
        int f1o[4], f1s[4], f3o[4][2];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int k1 = 8 * p + gq;  // F1: r = 8 rb + 2 tq + i
            f1o[p] = Pack32::voff_rt(k1) + tq - 2 * (k1 >> 2);
            f1s[p] = Pack32::pstr_rt(k1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int k3 = 8 * p + 2 * tq + i;  // F3: r = 8 rb + gq
                f3o[p][i] = Pack32::voff_rt(k3) + (gq & 1) * Pack32::pstr_rt(k3) + (gq >> 1) - 2 * (k3 >> 2);
            }
        }
        auto F1 = [&](int p, int rb, int i) -> double {
            double v = vb[f1o[p] + i * f1s[p] + 4 * rb];
            if (rb == p) v = (i ? m0b : m0) ? v : 0.0;
            return v;
        };
        auto F3 = [&](int p, int rb, int i) -> double {
            double v = vb[f3o[p][i] + 4 * rb];
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
        __syncwarp();  // the previous matrix's readers of the scratch are done
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
                if (k < 7) {
#pragma unroll
                    for (int kk = (k + 1) & ~1; kk < 8; kk += 2) {
                        const double2 g2 = *reinterpret_cast<const double2*>(Gs + tq * 64 + k * 8 + kk);
                        if (kk > k) acc[kk] = fma(Trow[k], g2.x, acc[kk]);
                        acc[kk + 1] = fma(Trow[k], g2.y, acc[kk + 1]);
                    }
                }
            }
            __syncwarp();  // every lane has read G: the scratch now takes -T
#pragma unroll
            for (int k = 0; k < 8; k += 2)
                *reinterpret_cast<double2*>(Gs + tq * 64 + gq * 8 + k) = make_double2(-Trow[k], -Trow[k + 1]);
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
            const double2 tt = *reinterpret_cast<const double2*>(Gs + p * 64 + gq * 8 + 2 * tq);
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
