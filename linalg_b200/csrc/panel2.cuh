// K4a (second generation): Householder panel factorisation inside a thread-block cluster, one
// asynchronous DSMEM exchange per column instead of a cluster barrier.
//
// Reference semantics: the column loop of linalg/qr.py:75-91 restricted to an (mp x 32) panel.
// Same register layout and same outputs as panel.cuh (R rows back into A, unit-norm reflectors
// into V, the panel's compact-WY factor T), different plumbing of the latency-bound column step:
//   * the column loop is fully unrolled (32 compile-time steps): the pivot row is read straight
//     out of the registers of the lanes that own it, finished column slots are skipped in the
//     rank-1 update, row masks are compile-time bounded;
//   * the per-CTA partial dot products travel to every CTA of the cluster with
//     st.async.shared::cluster (SASS STAS), each store signalling the receiver's mbarrier with its
//     byte count, so a CTA waits only for the DATA of the current column (one DSMEM latency), never
//     for a rendezvous of all CTAs (barrier.cluster costs ~380 cycles + an L1 flush per column);
//     two receive buffers alternate, the data flow itself is the flow control (a CTA can push
//     column j+2 only after it has received every peer's column j+1, i.e. after the peers have
//     consumed column j);
//   * all eight warps share the push (two target CTAs per warp);
//   * the T-factor recurrence of column j-1 runs in one warp while that warp would otherwise sit
//     waiting for the exchange of column j, so it never delays the column chain;
//   * 1/||v_c|| is formed once per column, not per stored element.
#pragma once

#include "common.cuh"

namespace lq {

constexpr int P2_WARPS = 8;
constexpr int P2_THREADS = P2_WARPS * 32;
constexpr int P2_MAXCS = 16;
constexpr int P2_NB = 32;

__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v),
                 "r"(remote_bar)
                 : "memory");
}

template <int RPT>
struct Panel2Cfg {
    static constexpr int C = 4;
    static constexpr int ROWS_PER_WARP = 4 * RPT;
    static constexpr int ROWS_PER_CTA = P2_WARPS * ROWS_PER_WARP;
    static constexpr int VSTRIDE = RPT + 2;  // row lanes staggered by 4 banks (broadcast LDS.128 of the 4 groups do not collide)
    static_assert(ROWS_PER_WARP >= P2_NB, "the top 32 x 32 block must live in warp 0 of CTA 0");
    __device__ __host__ static constexpr int col(int s, int lc) { return (s & 1) ? ((s + 1) * 8 - 1 - lc) : (s * 8 + lc); }
};

template <int RPT>
struct __align__(16) Panel2Smem {
    double vbuf[P2_WARPS][4][Panel2Cfg<RPT>::VSTRIDE];  // published column, per warp and row lane
    double part[2][P2_WARPS][32];                       // per-warp partial dot products (double buffered)
    double ppart[2][32];                                // pivot row (written by CTA 0, warp 0)
    double recv[2][P2_MAXCS][32];                       // exchange: per source CTA partials
    double precv[2][32];                                // exchange: pivot row
    double Tt[32][33];                                  // T_u transposed: Tt[k][i] = T_u[i][k]
    double gsave[32];                                   // g of the previous column (T warp only)
    double beta[32];                                    // 2 / v^T v  (0 if skipped)
    double v0[32];                                      // pivot entry of v
    double rdiag[32];                                   // R[j][j]
    double rn[32];                                      // 1 / ||v_c||  (0 if skipped)
    double nv[32];                                      // ||v_c||      (0 if skipped)
    double gall[32][33];                                // SOLO kernel: g of every column (T is built after the column loop)
    uint64_t bar[2];
};

// TRACE: warp 0 of CTA 0 and the T warp write clock64() stamps of the phases of every column step (diagnostics)
#define P2_STAMP(slot)                                                                  \
    do {                                                                                \
        if (TRACE && lane == 0 && (top_warp || t_warp))                                 \
            trace[((t_warp ? 1 : 0) * 32 + j) * 8 + (slot)] = clock64();                \
    } while (0)

// SOLO (round 2): the panel fits ONE CTA (mp <= 32 RPT rows).  The whole st.async / mbarrier exchange disappears (it cost ~1.5 us
// per column even for a 33-row panel: 48 us per panel whatever its height): after the one block barrier of the column every warp
// sums the eight per-warp partials itself.  The T recurrence, which the cluster kernel hides in the shadow of the exchange, is
// done once after the column loop from the saved g vectors.  Plain (non-cluster) launch.
template <int RPT, bool TRACE, bool SOLO = false>
__global__ void __launch_bounds__(P2_THREADS, 1)
    panel2_cluster_kernel(double* __restrict__ A, int lda, double* __restrict__ V, int ldv, double* __restrict__ T, int ldt,
                          int mp, long long* __restrict__ trace) {
    using Cfg = Panel2Cfg<RPT>;
    constexpr int C = Cfg::C, NB = P2_NB;
    __shared__ Panel2Smem<RPT> sm;

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = lane >> 3, lc = lane & 7;
    const int q = SOLO ? 0 : (int)cluster_ctarank();
    const int CS = SOLO ? 1 : (int)cluster_nctarank();
    const int rowbase = q * Cfg::ROWS_PER_CTA + w * Cfg::ROWS_PER_WARP;
    const bool top_warp = (q == 0) && (w == 0);              // holds rows 0 .. 4*RPT-1, i.e. the whole top block
    const bool t_warp = (q == CS - 1) && (w == P2_WARPS - 1);  // maintains the T factor

    int colv[C];
#pragma unroll
    for (int s = 0; s < C; ++s) colv[s] = Cfg::col(s, lc);

    // ---- load (zero fill below mp)
    double r[C][RPT];
#pragma unroll
    for (int ii = 0; ii < RPT; ++ii) {
        const int row = rowbase + 4 * ii + p;
#pragma unroll
        for (int s = 0; s < C; ++s) r[s][ii] = (row < mp) ? A[(long long)row * lda + colv[s]] : 0.0;
    }
    for (int e = threadIdx.x; e < 32 * 33; e += P2_THREADS) (&sm.Tt[0][0])[e] = 0.0;
    for (int e = threadIdx.x; e < 2 * P2_MAXCS * 32; e += P2_THREADS) (&sm.recv[0][0][0])[e] = 0.0;
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (!SOLO) cluster_sync_all();  // every CTA's barriers are initialised before the first remote store

    const uint32_t expect_bytes = (uint32_t)(CS * 32 * sizeof(double) + 32 * sizeof(double));
    // my two push targets
    const int tgt0 = w, tgt1 = w + P2_WARPS;
    double beta_prev = 0.0;  // T warp: beta of the previous column

#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int so = j >> 3;
        const int lo = (so & 1) ? ((so + 1) * 8 - 1 - j) : (j - so * 8);
        const int par = j & 1;

        P2_STAMP(0);
        if (!SOLO && threadIdx.x == 0) mbar_expect_tx(&sm.bar[par], expect_bytes);  // arm this column's phase

        // ---- owner lanes publish their part of column j
        if (lc == lo) {
            double* dst = &sm.vbuf[w][p][0];
#pragma unroll
            for (int ii = 0; ii < RPT; ii += 2) *reinterpret_cast<double2*>(dst + ii) = make_double2(r[so][ii], r[so][ii + 1]);
        }
        __syncwarp();
        double xv[RPT];
        {
            const double* src = &sm.vbuf[w][p][0];
#pragma unroll
            for (int ii = 0; ii < RPT; ii += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(src + ii);
                xv[ii] = t2.x;
                xv[ii + 1] = t2.y;
            }
        }
        if (top_warp) {
            // rows above the pivot do not take part; the pivot row goes out as it is
#pragma unroll
            for (int ii = 0; ii < RPT; ++ii)
                if (4 * ii < j) {
                    if (4 * ii + p < j) xv[ii] = 0.0;
                }
            if (p == (j & 3)) {
#pragma unroll
                for (int s = 0; s < C; ++s) sm.ppart[par][colv[s]] = r[s][j >> 2];
            }
        }

        // ---- partial dots x^T P[:, c]  (c > j: update, c < j: T factor, c = j: squared norm)
        double d[C];
        {
            double dB[C];
#pragma unroll
            for (int s = 0; s < C; ++s) d[s] = 0.0, dB[s] = 0.0;
#pragma unroll
            for (int ii = 0; ii < RPT; ii += 2)
#pragma unroll
                for (int s = 0; s < C; ++s) {
                    d[s] = fma(xv[ii], r[s][ii], d[s]);
                    dB[s] = fma(xv[ii + 1], r[s][ii + 1], dB[s]);
                }
#pragma unroll
            for (int s = 0; s < C; ++s) d[s] += dB[s];
        }
#pragma unroll
        for (int s = 0; s < C; ++s) {
            d[s] += __shfl_xor_sync(0xffffffffu, d[s], 8);
            d[s] += __shfl_xor_sync(0xffffffffu, d[s], 16);
        }
        if (p == 0) {
#pragma unroll
            for (int s = 0; s < C; ++s) sm.part[par][w][colv[s]] = d[s];
        }
        P2_STAMP(1);
        __syncthreads();
        P2_STAMP(2);

        // ---- every warp forms the CTA sum (lane = column) and pushes it to its two target CTAs
        if (!SOLO) {
            double s0 = sm.part[par][0][lane] + sm.part[par][1][lane];
            double s1 = sm.part[par][2][lane] + sm.part[par][3][lane];
            double s2 = sm.part[par][4][lane] + sm.part[par][5][lane];
            double s3 = sm.part[par][6][lane] + sm.part[par][7][lane];
            const double sum = (s0 + s1) + (s2 + s3);
            const uint32_t dst_local = smem_u32(&sm.recv[par][q][lane]);
            const uint32_t bar_local = smem_u32(&sm.bar[par]);
            if (tgt0 < CS) st_async_f64(mapa_shared(dst_local, (uint32_t)tgt0), sum, mapa_shared(bar_local, (uint32_t)tgt0));
            if (tgt1 < CS) st_async_f64(mapa_shared(dst_local, (uint32_t)tgt1), sum, mapa_shared(bar_local, (uint32_t)tgt1));
            if (q == 0) {
                const double pv = sm.ppart[par][lane];
                const uint32_t pdst = smem_u32(&sm.precv[par][lane]);
                if (tgt0 < CS) st_async_f64(mapa_shared(pdst, (uint32_t)tgt0), pv, mapa_shared(bar_local, (uint32_t)tgt0));
                if (tgt1 < CS) st_async_f64(mapa_shared(pdst, (uint32_t)tgt1), pv, mapa_shared(bar_local, (uint32_t)tgt1));
            }
        }

        P2_STAMP(3);
        // ---- T column of the PREVIOUS reflector, in the shadow of the exchange
        //      T_u[0:jj, jj] = -beta_jj * T_u[0:jj, 0:jj] * g[0:jj],  T_u[jj][jj] = beta_jj
        if (!SOLO && j > 0 && t_warp) {
            const int jj = j - 1;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int k = 0; k < jj; ++k) {
                const double tk = sm.Tt[k][lane], gk = sm.gsave[k];
                if ((k & 3) == 0) a0 = fma(tk, gk, a0);
                else if ((k & 3) == 1) a1 = fma(tk, gk, a1);
                else if ((k & 3) == 2) a2 = fma(tk, gk, a2);
                else a3 = fma(tk, gk, a3);
            }
            const double acc = (a0 + a1) + (a2 + a3);
            __syncwarp();
            if (lane < jj) sm.Tt[jj][lane] = -beta_prev * acc;
            if (lane == jj) sm.Tt[jj][lane] = beta_prev;
            __syncwarp();
        }

        // ---- wait for the column's data from every CTA
        P2_STAMP(4);
        if (!SOLO) mbar_wait(&sm.bar[par], (uint32_t)((j >> 1) & 1));
        P2_STAMP(5);

        // totals: lane c holds column c.  Always 16 slots (the unused ones stay zero): no loop, no branches
        double tot, prow, x0;
        if (SOLO) {
            const double s0 = sm.part[par][0][lane] + sm.part[par][1][lane];
            const double s1 = sm.part[par][2][lane] + sm.part[par][3][lane];
            const double s2 = sm.part[par][4][lane] + sm.part[par][5][lane];
            const double s3 = sm.part[par][6][lane] + sm.part[par][7][lane];
            tot = (s0 + s1) + (s2 + s3);
            prow = sm.ppart[par][lane];
            x0 = sm.ppart[par][j];
        } else {
            double tt[P2_MAXCS];
#pragma unroll
            for (int t = 0; t < P2_MAXCS; ++t) tt[t] = sm.recv[par][t][lane];
#pragma unroll
            for (int o = P2_MAXCS / 2; o > 0; o >>= 1)
#pragma unroll
                for (int t = 0; t < o; ++t) tt[t] += tt[t + o];
            tot = tt[0];
            prow = sm.precv[par][lane];
            x0 = sm.precv[par][j];                             // pivot (broadcast load)
        }
        const double ss = __shfl_sync(0xffffffffu, tot, j);   // sum_{r>=j} x_r^2
        // y = 1/||x||;  beta = 2 / v^T v = 1 / (||x|| (||x|| + |x0|)) = y^2 / (1 + |x0| y).  The reciprocal's seed comes
        // from the unrefined y, so its MUFU runs beside the Newton steps of y instead of behind them.
        const double ssc = fmax(ss, 1e-300);
        const double ax0 = fabs(x0);
        double y = rsqrt_seed(ssc);
        double u = rcp_seed(fma(ax0, y, 1.0));
        {
            const double hx = 0.5 * ssc;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const double e = fma(-hx * y, y, 0.5);
                y = fma(y, e, y);
            }
        }
        const double nrm = ssc * y;
        {
            const double D = fma(ax0, y, 1.0);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const double e = fma(-D, u, 1.0);
                u = fma(u, e, u);
            }
        }
        const bool skip = nrm < kEps;  // qr.py:79-80
        const double alpha = copysign(nrm, x0);
        const double v0 = x0 + alpha;
        const double beta = skip ? 0.0 : (y * y) * u;
        // g_c = v^T P[:, c] = x^T P[:, c] + alpha * P[j][c]
        const double gl = fma(alpha, prow, tot);
        double sc[C];
#pragma unroll
        for (int s = 0; s < C; ++s) {
            if (s >= so) {
                const double gs = __shfl_sync(0xffffffffu, gl, colv[s]);
                sc[s] = (colv[s] > j) ? beta * gs : 0.0;
            } else {
                sc[s] = 0.0;
            }
        }

        P2_STAMP(6);
        // ---- update P[j:, c] -= s_c v   (v = x except v0 on the pivot row, 0 above it); finished slots are skipped
        if (top_warp && p == (j & 3)) xv[j >> 2] = v0;
#pragma unroll
        for (int s = 0; s < C; ++s)
            if (s >= so) {
#pragma unroll
                for (int ii = 0; ii < RPT; ++ii) r[s][ii] = fma(-sc[s], xv[ii], r[s][ii]);
            }

        P2_STAMP(7);
        // ---- bookkeeping
        if (threadIdx.x == 0) {
            sm.beta[j] = beta;
            sm.v0[j] = v0;
            sm.rdiag[j] = skip ? x0 : -alpha;
        }
        if (SOLO) {
            if (w == 0) sm.gall[j][lane] = gl;
        } else if (t_warp) {
            __syncwarp();
            sm.gsave[lane] = gl;
            beta_prev = beta;
            __syncwarp();
        }
    }
    if (SOLO) {
        // T_u[0:jj, jj] = -beta_jj T_u[0:jj, 0:jj] g_jj[0:jj], T_u[jj][jj] = beta_jj: one warp, lane = row of T_u
        __syncthreads();
        if (w == 0) {
            for (int jj = 0; jj < NB; ++jj) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int k = 0; k + 3 < jj; k += 4) {
                    a0 = fma(sm.Tt[k][lane], sm.gall[jj][k], a0);
                    a1 = fma(sm.Tt[k + 1][lane], sm.gall[jj][k + 1], a1);
                    a2 = fma(sm.Tt[k + 2][lane], sm.gall[jj][k + 2], a2);
                    a3 = fma(sm.Tt[k + 3][lane], sm.gall[jj][k + 3], a3);
                }
                for (int k = jj & ~3; k < jj; ++k) a0 = fma(sm.Tt[k][lane], sm.gall[jj][k], a0);
                const double bj = sm.beta[jj];
                __syncwarp();
                if (lane < jj) sm.Tt[jj][lane] = -bj * ((a0 + a1) + (a2 + a3));
                if (lane == jj) sm.Tt[jj][lane] = bj;
                __syncwarp();
            }
        }
    }
    // T column of the last reflector
    if (!SOLO && t_warp) {
        const int jj = NB - 1;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < jj; k += 2) {
            a0 = fma(sm.Tt[k][lane], sm.gsave[k], a0);
            if (k + 1 < jj) a1 = fma(sm.Tt[k + 1][lane], sm.gsave[k + 1], a1);
        }
        const double acc = a0 + a1;
        __syncwarp();
        if (lane < jj) sm.Tt[jj][lane] = -beta_prev * acc;
        if (lane == jj) sm.Tt[jj][lane] = beta_prev;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const double bt = sm.beta[threadIdx.x];
        const double rnv = (bt > 0.0) ? sqrt(0.5 * bt) : 0.0;  // 1 / ||v||,  ||v||^2 = 2 / beta
        sm.rn[threadIdx.x] = rnv;
        sm.nv[threadIdx.x] = (bt > 0.0) ? 1.0 / rnv : 0.0;
    }
    __syncthreads();

    // ---- store: R rows (top block), normalised reflectors, T
    {
        double rnc[C], v0c[C], rdc[C];
#pragma unroll
        for (int s = 0; s < C; ++s) {
            rnc[s] = sm.rn[colv[s]];
            v0c[s] = sm.v0[colv[s]];
            rdc[s] = sm.rdiag[colv[s]];
        }
#pragma unroll
        for (int ii = 0; ii < RPT; ++ii) {
            const int row = rowbase + 4 * ii + p;
            if (row >= mp) continue;
#pragma unroll
            for (int s = 0; s < C; ++s) {
                const int c = colv[s];
                double vv = r[s][ii] * rnc[s];
                if (top_warp) {
                    if (row == c) vv = v0c[s] * rnc[s];
                    else if (row < c) vv = 0.0;
                    if (row < c) A[(long long)row * lda + c] = r[s][ii];
                    else if (row == c) A[(long long)row * lda + c] = rdc[s];
                }
                V[(long long)row * ldv + c] = vv;
            }
        }
    }
    if (q == CS - 1) {
        // T_n[i][k] = T_u[i][k] * ||v_i|| ||v_k||
        for (int e = threadIdx.x; e < NB * NB; e += P2_THREADS) {
            const int i = e >> 5, k = e & 31;
            const double val = (i <= k) ? sm.Tt[k][i] * sm.nv[i] * sm.nv[k] : 0.0;
            T[(long long)i * ldt + k] = val;
        }
    }
    if (!SOLO) cluster_sync_all();  // nobody exits while a peer may still target its shared memory
}

}  // namespace lq
