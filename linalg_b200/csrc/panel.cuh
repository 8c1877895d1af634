// K4a: Householder panel factorisation inside a thread-block cluster.
//
// Reference semantics: the column loop of linalg/qr.py:75-91 restricted to an (mp x nb) panel,
// nb <= 8*C.  The panel lives in REGISTERS: lane (p, lc) of warp w of cluster CTA q holds rows
//   row(ii) = q*ROWS_PER_CTA + w*4*RPT + 4*ii + p      (ii = 0..RPT-1, p = 0..3)
// of the C folded column slots col(s, lc) (lc = 0..7), i.e. C*RPT doubles per lane.  One column
// step costs: the owner lanes publish their column through a warp-private shared buffer, every
// lane accumulates x^T P[:, c] for all its columns (c > j feeds the update, c < j feeds the T
// factor, c = j is the squared norm), partial sums are combined over the 4 row lanes (shuffles),
// the 8 warps (shared memory) and the cluster CTAs (DSMEM push + one cluster barrier), after
// which every CTA redundantly forms alpha, v0, beta and updates its rows.  No global memory is
// touched between the initial load and the final store.
//
// Outputs: R rows of the panel (upper triangle of the top nb x nb block) back into A; unit-norm
// reflectors w_j = v_j/||v_j|| into V (explicit diagonal, zeros above it; skipped columns zero);
// the compact-WY factor T (nb x nb, upper) of  H_0 H_1 ... H_{nb-1} = I - V T V^T  (tau = 2).
#pragma once

#include "common.cuh"

namespace lq {

constexpr int PANEL_WARPS = 8;
constexpr int PANEL_THREADS = PANEL_WARPS * 32;
constexpr int PANEL_MAXCS = 16;

template <int C, int RPT>
struct PanelCfg {
    static constexpr int NBMAX = 8 * C;
    static constexpr int ROWS_PER_WARP = 4 * RPT;
    static constexpr int ROWS_PER_CTA = PANEL_WARPS * ROWS_PER_WARP;
    __device__ __host__ static constexpr int col(int s, int lc) { return (s & 1) ? ((s + 1) * 8 - 1 - lc) : (s * 8 + lc); }
};

struct PanelSmem {
    double vbuf[PANEL_WARPS][4][32];                 // published column, per warp, per row lane (RPT <= 32)
    double part[PANEL_WARPS][32];                    // per-warp partial dot products
    double ppart[32];                                // pivot row captured by the pivot warp
    double recv[2][PANEL_MAXCS][32];                 // cluster exchange: per source CTA partials
    double precv[2][32];                             // cluster exchange: pivot row (from the owning CTA)
    double Tt[32][33];                               // T_u transposed: Tt[k][i] = T_u[i][k]
    double gbuf[32];
    double beta[32];                                 // 2 / v^T v  (0 if skipped)
    double v0[32];                                   // pivot entry of v
    double rdiag[32];                                // R[j][j]
};

template <int C, int RPT>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
    panel_cluster_kernel(double* __restrict__ A, int lda, double* __restrict__ V, int ldv, double* __restrict__ T,
                         int ldt, int mp, int nb) {
    using Cfg = PanelCfg<C, RPT>;
    __shared__ PanelSmem sm;

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = lane >> 3, lc = lane & 7;
    const int q = (int)cluster_ctarank();
    const int CS = (int)cluster_nctarank();
    const int rowbase = q * Cfg::ROWS_PER_CTA + w * Cfg::ROWS_PER_WARP;  // panel-local row of (ii=0,p=0)
    const bool low_warp = rowbase < nb;  // this warp holds rows inside the top nb x nb block

    int colv[C];
#pragma unroll
    for (int s = 0; s < C; ++s) colv[s] = Cfg::col(s, lc);

    // ---- load (zero fill outside mp x nb)
    double r[C][RPT];
#pragma unroll
    for (int ii = 0; ii < RPT; ++ii) {
        const int row = rowbase + 4 * ii + p;
#pragma unroll
        for (int s = 0; s < C; ++s) {
            double val = 0.0;
            if (row < mp && colv[s] < nb) val = A[(long long)row * lda + colv[s]];
            r[s][ii] = val;
        }
    }
    for (int e = threadIdx.x; e < 32 * 33; e += PANEL_THREADS) (&sm.Tt[0][0])[e] = 0.0;
    if (threadIdx.x < 32) {
        sm.beta[threadIdx.x] = 0.0;
        sm.v0[threadIdx.x] = 0.0;
        sm.rdiag[threadIdx.x] = 0.0;
    }
    __syncthreads();
    cluster_sync_all();  // every CTA of the cluster is resident before the first DSMEM store

    for (int j = 0; j < nb; ++j) {
        const int so = j >> 3;
        const int lo = (so & 1) ? ((so + 1) * 8 - 1 - j) : (j - so * 8);
        const int par = j & 1;
        const bool piv_warp = (j >= rowbase) && (j < rowbase + Cfg::ROWS_PER_WARP);

        // ---- owner lanes publish their part of column j
        if (lc == lo) {
            double* dst = &sm.vbuf[w][p][0];
#pragma unroll
            for (int s = 0; s < C; ++s)
                if (s == so) {
#pragma unroll
                    for (int ii = 0; ii < RPT; ii += 2) *reinterpret_cast<double2*>(dst + ii) = make_double2(r[s][ii], r[s][ii + 1]);
                }
        }
        __syncwarp();

        // ---- partial dots x^T P[:, c] over rows >= j; the pivot warp also extracts row j
        double xv[RPT];
        double d[C], e[C];
#pragma unroll
        for (int s = 0; s < C; ++s) d[s] = 0.0, e[s] = 0.0;
        {
            const double* src = &sm.vbuf[w][p][0];
#pragma unroll
            for (int ii = 0; ii < RPT; ii += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(src + ii);
                xv[ii] = t2.x;
                xv[ii + 1] = t2.y;
            }
        }
        if (low_warp) {
#pragma unroll
            for (int ii = 0; ii < RPT; ++ii) {
                const int row = rowbase + 4 * ii + p;
                if (row < j) xv[ii] = 0.0;
                if (piv_warp) {
                    const double sel = (row == j) ? 1.0 : 0.0;
#pragma unroll
                    for (int s = 0; s < C; ++s) e[s] = fma(sel, r[s][ii], e[s]);
                }
            }
        }
        {
            double dB[C];  // second accumulator per column slot: halves the dependent DFMA chains
#pragma unroll
            for (int s = 0; s < C; ++s) dB[s] = 0.0;
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
        // combine the 4 row lanes
#pragma unroll
        for (int s = 0; s < C; ++s) {
            d[s] += __shfl_xor_sync(0xffffffffu, d[s], 8);
            d[s] += __shfl_xor_sync(0xffffffffu, d[s], 16);
        }
        if (piv_warp) {
#pragma unroll
            for (int s = 0; s < C; ++s) {
                e[s] += __shfl_xor_sync(0xffffffffu, e[s], 8);
                e[s] += __shfl_xor_sync(0xffffffffu, e[s], 16);
            }
        }
        if (p == 0) {
#pragma unroll
            for (int s = 0; s < C; ++s) {
                sm.part[w][colv[s]] = d[s];
                if (piv_warp) sm.ppart[colv[s]] = e[s];
            }
        }
        __syncthreads();

        // ---- CTA sum over warps, push to every CTA of the cluster (DSMEM)
        if (w == 0) {
            double sum = 0.0;
#pragma unroll
            for (int ww = 0; ww < PANEL_WARPS; ++ww) sum += sm.part[ww][lane];
            const uint32_t dst_local = smem_u32(&sm.recv[par][q][lane]);
            for (int t = 0; t < CS; ++t) st_cluster_f64(mapa_shared(dst_local, (uint32_t)t), sum);
            const int piv_cta = j / Cfg::ROWS_PER_CTA;
            if (q == piv_cta) {
                const double pv = sm.ppart[lane];
                const uint32_t pdst = smem_u32(&sm.precv[par][lane]);
                for (int t = 0; t < CS; ++t) st_cluster_f64(mapa_shared(pdst, (uint32_t)t), pv);
            }
        }
        cluster_sync_all();

        // ---- every warp forms the totals for its columns: lane c holds column c
        double tot;
        {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;  // CS is 1, 2, 4, 8 or 16
            if (CS >= 4) {
                for (int t = 0; t < CS; t += 4) {
                    t0 += sm.recv[par][t][lane];
                    t1 += sm.recv[par][t + 1][lane];
                    t2 += sm.recv[par][t + 2][lane];
                    t3 += sm.recv[par][t + 3][lane];
                }
            } else {
                for (int t = 0; t < CS; ++t) t0 += sm.recv[par][t][lane];
            }
            tot = (t0 + t1) + (t2 + t3);
        }
        const double prow = sm.precv[par][lane];
        const double ss = __shfl_sync(0xffffffffu, tot, j);    // sum_{r>=j} x_r^2
        const double x0 = __shfl_sync(0xffffffffu, prow, j);   // pivot
        double rinv_n;
        const double nrm = sqrt_nr(fmax(ss, 1e-300), rinv_n);  // MUFU seed + Newton steps, full double accuracy
        const bool skip = nrm < kEps;  // qr.py:79-80
        const double alpha = copysign(nrm, x0);
        const double v0 = x0 + alpha;
        const double beta = skip ? 0.0 : rcp_nr(nrm * fabs(v0));
        // g_c = v^T P[:, c] = x^T P[:, c] + alpha * P[j][c]
        const double gl = fma(alpha, prow, tot);
        double sc[C];
#pragma unroll
        for (int s = 0; s < C; ++s) {
            const double gs = __shfl_sync(0xffffffffu, gl, colv[s]);
            sc[s] = (colv[s] > j) ? beta * gs : 0.0;
        }

        // ---- update P[j:, c] -= s_c v   (v = x except v0 on the pivot row, 0 above it)
        if (piv_warp) {
#pragma unroll
            for (int ii = 0; ii < RPT; ++ii) {
                const int row = rowbase + 4 * ii + p;
                if (row == j) xv[ii] = v0;
            }
        }
#pragma unroll
        for (int ii = 0; ii < RPT; ++ii)
#pragma unroll
            for (int s = 0; s < C; ++s) r[s][ii] = fma(-sc[s], xv[ii], r[s][ii]);

        // ---- bookkeeping (every CTA keeps its own copy) and the T column (last warp of last CTA)
        if (threadIdx.x == 0) {
            sm.beta[j] = beta;
            sm.v0[j] = v0;
            sm.rdiag[j] = skip ? x0 : -alpha;
        }
        if (q == CS - 1 && w == PANEL_WARPS - 1) {
            // T_u[0:j, j] = -beta * T_u[0:j, 0:j] * g[0:j],  T_u[j][j] = beta
            sm.gbuf[lane] = gl;
            __syncwarp();
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = 0;
            for (; k + 4 <= j; k += 4) {
                a0 = fma(sm.Tt[k][lane], sm.gbuf[k], a0);
                a1 = fma(sm.Tt[k + 1][lane], sm.gbuf[k + 1], a1);
                a2 = fma(sm.Tt[k + 2][lane], sm.gbuf[k + 2], a2);
                a3 = fma(sm.Tt[k + 3][lane], sm.gbuf[k + 3], a3);
            }
            for (; k < j; ++k) a0 = fma(sm.Tt[k][lane], sm.gbuf[k], a0);
            const double acc = (a0 + a1) + (a2 + a3);
            __syncwarp();
            if (lane < j) sm.Tt[j][lane] = -beta * acc;
            if (lane == j) sm.Tt[j][lane] = beta;
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- store: R rows (top block), normalised reflectors, T
#pragma unroll
    for (int ii = 0; ii < RPT; ++ii) {
        const int row = rowbase + 4 * ii + p;
        if (row >= mp) continue;
#pragma unroll
        for (int s = 0; s < C; ++s) {
            const int c = colv[s];
            if (c >= nb) continue;
            const double bt = sm.beta[c];
            const double rn = sqrt(0.5 * bt);  // 1 / ||v_c||
            double vv;
            if (row > c) vv = r[s][ii] * rn;
            else if (row == c) vv = sm.v0[c] * rn;
            else vv = 0.0;
            V[(long long)row * ldv + c] = vv;
            if (row < c) A[(long long)row * lda + c] = r[s][ii];
            else if (row == c) A[(long long)row * lda + c] = sm.rdiag[c];
        }
    }
    if (q == CS - 1 && w == PANEL_WARPS - 1) {
        // T_n[i][k] = T_u[i][k] * ||v_i|| ||v_k||
        for (int k = 0; k < nb; ++k) {
            const int i = lane;
            if (i < nb) {
                const double bi = sm.beta[i], bk = sm.beta[k];
                double val = 0.0;
                if (i <= k && bi > 0.0 && bk > 0.0) val = sm.Tt[k][i] * 2.0 / sqrt(bi * bk);
                T[(long long)i * ldt + k] = val;
            }
        }
    }
    cluster_sync_all();  // nobody exits while a peer may still target its shared memory
}

}  // namespace lq
