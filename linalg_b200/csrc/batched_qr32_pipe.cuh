// K1 (software-pipelined): batched 32x32 Householder QR, the R phase of matrix pair n+1 interleaved with the
// Q phase of matrix pair n in the SAME warp.
//
// Reference semantics: linalg/qr.py:52-100 on every A[b] of a (batch, 32, 32) row-major float64 array.
//
// Why: hh_qr32_kernel (batched_qr32.cuh) is bound by latency, not by a pipe: a warp issues ~185 cycles of
// instructions per column step but the step's dependent chain (publish -> read back -> dots -> shuffle -> norm ->
// rsqrt / rcp -> scale -> update) is ~670 cycles long, and the 128 registers of matrix data per thread allow only
// two warps per scheduler (54 % of the issue slots used).  More warps are impossible, so the second instruction
// stream has to come from inside the warp.  The R phase (j = 0 .. 31) touches the trailing block A[j:, j:] and the Q
// phase (j = 31 .. 0) touches Q[j:, j:]: while one shrinks the other grows, their live registers add up to about
// one matrix.  So a persistent warp walks over its matrix pairs and runs, per combined step t, the R step t of the
// next pair and the Q step 31 - t of the current pair: two independent dependency chains for the scheduler at the
// register cost of one.  R rows leave for global memory as soon as they are final (two rows at a time), Q elements
// are created (identity) when the block reaches them.
//
// Layout: Dist32<2, 4> of batched_qr32.cuh (two matrices per warp, lane (p, lc) holds rows 2 ii + p of four folded
// column slots).  The reflectors live in shared memory, triangular-packed (reflector j keeps the row pairs
// ii >= ii0(j) only) and double buffered (R phase of pair n+1 writes one buffer while the Q phase of pair n reads
// the other): 12.6 KB per matrix, 4 CTAs of 2 warps per SM.
#pragma once

#include "batched_qr32.cuh"

namespace lq {

struct Pipe32 {
    static constexpr int N = 32, P = 2, C = 4, LC = 8, L = 16, RPL = 16;
    __host__ __device__ static constexpr int ii0(int j) { return (j / 2) & ~1; }
    __host__ __device__ static constexpr int so(int j) { return j / LC; }
    // doubles between the two row-parity sub-rows of reflector j: length + 2, nudged so that the broadcast 16-byte loads
    // of the four (matrix, parity) groups of a warp never share banks (matrices are 64 bytes apart modulo 128)
    __host__ __device__ static constexpr int pstride(int j) {
        const int b = j / 4;  // four consecutive reflectors share a length: 16 - 2 b
        return 18 - 2 * b + ((b == 1 || b == 5) ? 2 : 0);
    }
    // offset of reflector j in the packed buffer (closed form: folds to a constant in the unrolled step loop)
    __host__ __device__ static constexpr int off(int j) {
        const int b = j / 4;
        return 8 * (18 * b - b * (b - 1)) + (b > 1 ? 16 : 0) + (b > 5 ? 16 : 0) + (j % 4) * 2 * pstride(j);
    }
    static constexpr int VDOUBLES = 736;                 // = off(32)
    static constexpr int SLOT = VDOUBLES + 2 * N + 8;    // reflectors + beta[32] + v0[32] + 64-byte stagger = 808
    static constexpr int WARP_DOUBLES = 4 * SLOT;        // 2 matrices x 2 buffers
    __host__ __device__ static constexpr bool live(int s, int ii, int j) { return s >= so(j) && ii >= ii0(j); }
};
static_assert(Pipe32::off(32) == Pipe32::VDOUBLES, "packed reflector size");
static_assert((Pipe32::SLOT * 8) % 128 == 64, "the two matrices of a warp are staggered by 64 bytes");

// Four combined steps: R steps j = 4 B .. 4 B + 3 of the matrix in r and Q steps j = 4 (7 - B) + 3 .. 4 (7 - B) of the
// matrix in q.  The four steps of a group share the live register ranges (column slots >= B / 2, row pairs >= 2 B), so
// they run as a ROLLED loop with a runtime step index (the few places that depend on it are selects): the code is a
// quarter of the fully unrolled form and every fetched instruction line is used four times in a row, which is what
// keeps eight warps per SM at different code positions inside the instruction cache.
template <bool DO_R, bool DO_Q, int NR, int B>
__device__ __forceinline__ void pipe32_group(double (&r)[4][16], double (&q)[4][16], double* __restrict__ vbR,
                                             const double* __restrict__ vbQ, double* __restrict__ Rg, bool validR,
                                             const int (&colv)[4], int p, int lc, int lm) {
    using Pp = Pipe32;
    constexpr int C = 4, RPL = 16, L = 16;
    constexpr int soR = B / 2, ii0R = 2 * B, offR = Pp::off(4 * B), pstR = Pp::pstride(4 * B);
    constexpr int BQ = 7 - B, soQ = BQ / 2, ii0Q = 2 * BQ, offQ = Pp::off(4 * BQ), pstQ = Pp::pstride(4 * BQ);

    if (DO_Q) {
        // elements of Q that the block reaches in this group start as the identity
#pragma unroll
        for (int s = 0; s < C; ++s)
#pragma unroll
            for (int ii = 0; ii < RPL; ++ii)
                if (Pp::live(s, ii, 4 * BQ) && !(BQ < 7 && Pp::live(s, ii, 4 * (BQ + 1)))) q[s][ii] = (2 * ii + p == colv[s]) ? 1.0 : 0.0;
    }

#pragma unroll
    for (int hiR = 0; hiR < 2; ++hiR)  // compile time: the pivot row pair ii0R + hiR is a fixed register
#pragma unroll 1
    for (int jpR = 0; jpR < 2; ++jpR) {  // run time: row parity of the pivot
        // ---------------- R step jr = 4 B + u: publish ----------------
        const int u = 2 * hiR + jpR;
        const int jr = 4 * B + u;
        const int loR = (soR & 1) ? (8 * (soR + 1) - 1 - jr) : (jr - 8 * soR);
        double* vjR = vbR + offR + u * (2 * pstR) + p * pstR - ii0R;
        if (DO_R) {
            if (lc == loR) {
#pragma unroll
                for (int ii = ii0R; ii < RPL; ii += 2) *reinterpret_cast<double2*>(vjR + ii) = make_double2(r[soR][ii], r[soR][ii + 1]);
            }
            __syncwarp();
        }

        // ---------------- Q step jq = 4 BQ + 3 - u: dots ----------------
        const int hiQ = 1 - hiR, jpQ = 1 - jpR;
        const int uq = 2 * hiQ + jpQ;
        const int jq = 4 * BQ + uq;
        const double* vjQ = vbQ + offQ + uq * (2 * pstQ) + p * pstQ - ii0Q;
        double dq[C], dq2[C], vq[RPL];
        double betaQ = 0.0;
        if (DO_Q) {
            betaQ = vbQ[Pp::VDOUBLES + jq];
            const double v0Q = vbQ[Pp::VDOUBLES + 32 + jq];
            const bool pivQ = (p == jpQ);
#pragma unroll
            for (int s = 0; s < C; ++s) dq[s] = 0.0, dq2[s] = 0.0;
#pragma unroll
            for (int ii = ii0Q; ii < RPL; ii += 2) {
                double2 vv = *reinterpret_cast<const double2*>(vjQ + ii);
                if (ii == ii0Q) {
                    const double lowx = pivQ ? v0Q : ((p > jpQ) ? vv.x : 0.0);
                    const double lowy = pivQ ? v0Q : ((p > jpQ) ? vv.y : 0.0);
                    vv.x = hiQ ? 0.0 : lowx;
                    vv.y = hiQ ? lowy : vv.y;
                }
                vq[ii] = vv.x;
                vq[ii + 1] = vv.y;
#pragma unroll
                for (int s = soQ; s < C; ++s) {
                    dq[s] = fma(vv.x, q[s][ii], dq[s]);
                    if (C - soQ >= 3) dq[s] = fma(vv.y, q[s][ii + 1], dq[s]);
                    else dq2[s] = fma(vv.y, q[s][ii + 1], dq2[s]);
                }
            }
            if (C - soQ < 3) {
#pragma unroll
                for (int s = soQ; s < C; ++s) dq[s] += dq2[s];
            }
        }

        // ---------------- R step: dots ----------------
        double dr[C], dr2[C], vr[RPL];
        if (DO_R) {
#pragma unroll
            for (int s = 0; s < C; ++s) dr[s] = 0.0, dr2[s] = 0.0;
#pragma unroll
            for (int ii = ii0R; ii < RPL; ii += 2) {
                double2 vv = *reinterpret_cast<const double2*>(vjR + ii);
                if (ii == ii0R) {
                    const bool below = (p >= jpR);
                    vv.x = (!hiR && below) ? vv.x : 0.0;
                    vv.y = (!hiR || below) ? vv.y : 0.0;
                }
                vr[ii] = vv.x;
                vr[ii + 1] = vv.y;
#pragma unroll
                for (int s = soR; s < C; ++s) {
                    dr[s] = fma(vv.x, r[s][ii], dr[s]);
                    if (C - soR >= 3) dr[s] = fma(vv.y, r[s][ii + 1], dr[s]);
                    else dr2[s] = fma(vv.y, r[s][ii + 1], dr2[s]);
                }
            }
            if (C - soR < 3) {
#pragma unroll
                for (int s = soR; s < C; ++s) dr[s] += dr2[s];
            }
        }

        // ---------------- Q step: combine the two row parities, scale ----------------
        if (DO_Q) {
#pragma unroll
            for (int s = soQ; s < C; ++s) dq[s] = betaQ * group_sum<2, 4>(dq[s]);
        }

        // ---------------- R step: norm, reflector scalars ----------------
        double alphaR = 0.0, v0R = 0.0;
        bool skipR = false, pivR = false;
        if (DO_R) {
            double ss = group_sum<2, 4>(dr[soR]);
            ss = __shfl_sync(0xffffffffu, ss, loR, L);
            const double x0 = vbR[offR + u * (2 * pstR) + jpR * pstR + hiR];
            const double ssc = fmax(ss, 1e-300);
            const double nrm = ssc * rsqrt_nr_t<NR>(ssc);
            skipR = nrm < kEps;  // qr.py:79-80
            alphaR = copysign(nrm, x0);
            v0R = x0 + alphaR;
            const double beta = skipR ? 0.0 : rcp_nr_t<NR>(nrm * fabs(v0R));  // 2 / v^T v
            if (lm == 0) {
                vbR[Pp::VDOUBLES + jr] = beta;
                vbR[Pp::VDOUBLES + 32 + jr] = v0R;
            }
            pivR = (p == jpR);
            const double alpha_m = pivR ? alphaR : 0.0;
#pragma unroll
            for (int s = soR; s < C; ++s) {
                const double rpiv = hiR ? r[s][ii0R + 1] : r[s][ii0R];
                const double part = group_sum<2, 4>(fma(alpha_m, rpiv, dr[s]));  // v^T R[:, c]
                dr[s] = beta * part;
            }
        }

        // ---------------- Q step: update ----------------
        if (DO_Q) {
#pragma unroll
            for (int ii = ii0Q; ii < RPL; ii += 2) {
#pragma unroll
                for (int s = soQ; s < C; ++s) {
                    q[s][ii] = fma(-dq[s], vq[ii], q[s][ii]);
                    q[s][ii + 1] = fma(-dq[s], vq[ii + 1], q[s][ii + 1]);
                }
            }
        }

        // ---------------- R step: update, finished rows leave ----------------
        if (DO_R) {
#pragma unroll
            for (int ii = ii0R; ii < RPL; ii += 2) {
                double vx = vr[ii], vy = vr[ii + 1];
                if (ii == ii0R) {  // pivot row carries v0 = x0 + alpha
                    vx = (pivR && !hiR) ? v0R : vx;
                    vy = (pivR && hiR) ? v0R : vy;
                }
#pragma unroll
                for (int s = soR; s < C; ++s) {
                    r[s][ii] = fma(-dr[s], vx, r[s][ii]);
                    r[s][ii + 1] = fma(-dr[s], vy, r[s][ii + 1]);
                }
            }
            {
                const bool diag = (lc == loR) && pivR && !skipR;  // exact diagonal
                r[soR][ii0R] = (diag && !hiR) ? -alphaR : r[soR][ii0R];
                r[soR][ii0R + 1] = (diag && hiR) ? -alphaR : r[soR][ii0R + 1];
            }
            if (jpR == 1 && validR) {
                // rows jr - 1 (p = 0) and jr (p = 1) of R are final; strict lower triangle exact zeros (qr.py:97)
                const int i = 2 * (ii0R + hiR) + p;
#pragma unroll
                for (int s = 0; s < C; ++s) {
                    double val = 0.0;
                    if (s >= soR) {
                        const double rv = hiR ? r[s][ii0R + 1] : r[s][ii0R];
                        val = (colv[s] >= i) ? rv : 0.0;
                    }
                    st_stream(Rg + (2 * (ii0R + hiR)) * 32 + colv[s], val);
                }
            }
        }
    }
}

// One pass of 32 combined steps.  DO_R: factor the matrix held in r (reflectors -> vbR, R rows -> Rg).
// DO_Q: accumulate Q of the pair whose reflectors are in vbQ into q (identity created on the fly).
template <bool DO_R, bool DO_Q, int NR>
__device__ __forceinline__ void pipe32_pass(double (&r)[4][16], double (&q)[4][16], double* __restrict__ vbR,
                                            const double* __restrict__ vbQ, double* __restrict__ Rg, bool validR,
                                            const int (&colv)[4], int p, int lc, int lm) {
    pipe32_group<DO_R, DO_Q, NR, 0>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 1>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 2>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 3>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 4>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 5>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 6>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    pipe32_group<DO_R, DO_Q, NR, 7>(r, q, vbR, vbQ, Rg, validR, colv, p, lc, lm);
    __syncwarp();
}

// grid: min(#SM * MINB, ceil(pairs / WARPS)) persistent CTAs of WARPS warps; warp w walks over the matrix pairs
// w, w + W, w + 2 W, ... (W = all warps of the grid).
template <int WARPS, int MINB, int NR>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    hh_qr32_pipe_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, long long batch) {
    using Pp = Pipe32;
    constexpr int N = 32, C = 4, RPL = 16, P = 2;
    extern __shared__ __align__(16) double smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / 16, lm = lane % 16, p = lm / 8, lc = lm % 8;
    const long long npairs = (batch + 1) / 2;
    const long long wstride = (long long)gridDim.x * WARPS;
    long long cur = (long long)blockIdx.x * WARPS + warp;
    if (cur >= npairs) return;  // (no block-wide barrier anywhere in this kernel)

    double* wb = smem + (size_t)warp * Pp::WARP_DOUBLES + g * Pp::SLOT;  // buffer b of my matrix: wb + b * 2 * SLOT
    int colv[C];
#pragma unroll
    for (int s = 0; s < C; ++s) colv[s] = Dist32<2, 4>::col(s, lc);

    double r[C][RPL], q[C][RPL];
    auto load_pair = [&](long long pair, bool& valid, long long& mat) {
        mat = 2 * pair + g;
        valid = mat < batch;
        const long long matc = valid ? mat : (batch - 1);
        const double* Ag = A + matc * (N * N) + p * N;
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
            for (int s = 0; s < C; ++s) r[s][ii] = ld_stream(Ag + ii * (P * N) + colv[s]);
    };
    auto prefetch_pair = [&](long long pair) {
        if (pair < npairs) {
            // 2 matrices = 16 KB = 128 lines of 128 bytes: 4 per lane (clamped inside the batch)
            const long long last = batch * (N * N) - 16;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                long long e = pair * (2 * N * N) + (long long)(k * 32 + lane) * 16;
                e = e < last ? e : last;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + e));
            }
        }
    };

    bool validR, validQ;
    long long matR, matQ;
    int b = 0;
    // prologue: R phase of the first pair alone
    load_pair(cur, validR, matR);
    prefetch_pair(cur + wstride);
    pipe32_pass<true, false, NR>(r, q, wb, wb, R + (validR ? matR : 0) * (N * N) + p * N, validR, colv, p, lc, lm);
    validQ = validR;
    matQ = matR;
    // steady state: R phase of the next pair interleaved with the Q phase of the current one
    for (long long nxt = cur + wstride; nxt < npairs; nxt += wstride) {
        load_pair(nxt, validR, matR);
        prefetch_pair(nxt + wstride);
        pipe32_pass<true, true, NR>(r, q, wb + (b ^ 1) * 2 * Pp::SLOT, wb + b * 2 * Pp::SLOT,
                                    R + (validR ? matR : 0) * (N * N) + p * N, validR, colv, p, lc, lm);
        if (validQ) {
            double* Qg = Q + matQ * (N * N) + p * N;
#pragma unroll
            for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
                for (int s = 0; s < C; ++s) st_stream(Qg + ii * (P * N) + colv[s], q[s][ii]);
        }
        validQ = validR;
        matQ = matR;
        b ^= 1;
    }
    // epilogue: Q phase of the last pair alone
    pipe32_pass<false, true, NR>(r, q, wb, wb + b * 2 * Pp::SLOT, R, false, colv, p, lc, lm);
    if (validQ) {
        double* Qg = Q + matQ * (N * N) + p * N;
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii)
#pragma unroll
            for (int s = 0; s < C; ++s) st_stream(Qg + ii * (P * N) + colv[s], q[s][ii]);
    }
}

}  // namespace lq
