// K5: tall-skinny paths -- Gram matrix, symmetric eigen-solver, economy SVD (linalg/svd.py:10-82)
// and TSQR, with the NCCL exchange steps of the row-sharded variants.
#include <algorithm>
#include <vector>

#include "../../include/linalg_b200.h"
#include "ops.cuh"
#include "stream_qr.cuh"

namespace lq {

namespace {

inline int grid_for(Ctx* c, long long total) {
    return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)c->sm_count * 8));
}

// ------------------------------------------------------------------ Jacobi eigen-solver
// Two-sided cyclic Jacobi with the round-robin ("circle") parallel ordering.  One CTA owns the
// matrix (shared memory when it fits, global memory otherwise); every round applies n/2 disjoint
// rotations, first to the columns then to the rows.  The (c, s) pairs are logged so that the
// eigenvectors can be accumulated afterwards by independent CTAs (one row of V each).
__device__ __forceinline__ void rr_pair(int n_even, int round, int k, int& p, int& q) {
    // players 0..n_even-2 on a circle, player n_even-1 fixed
    const int mod = n_even - 1;
    int a, b;
    if (k == 0) {
        a = n_even - 1;
        b = round % mod;
    } else {
        a = (round + k) % mod;
        b = (round - k + mod) % mod;
    }
    p = min(a, b);
    q = max(a, b);
}

__global__ void __launch_bounds__(1024) jacobi_kernel(const double* __restrict__ G, int n, double* __restrict__ Awork,
                                                     int use_global, double2* __restrict__ rotlog, int max_sweeps, double rel_tol,
                                                     int* __restrict__ nrounds_out, double* __restrict__ lambda) {
    extern __shared__ __align__(16) double sm[];
    const int ne = (n + 1) & ~1;          // padded to even; the pad index never rotates (zero coupling)
    const int ld = ne | 1;                // odd pitch
    double* Am = use_global ? Awork : sm;
    double2* cs = reinterpret_cast<double2*>(use_global ? sm : sm + (size_t)ne * ld + (ne & 1 ? 1 : 0) + 1);
    // align cs to 16 bytes
    cs = reinterpret_cast<double2*>((reinterpret_cast<uintptr_t>(cs) + 15) & ~(uintptr_t)15);
    __shared__ int s_rot;
    __shared__ double s_scale;
    __shared__ int2 pq[1024];  // pairs of the current round (n <= 2048)
    const int tid = threadIdx.x, nt = blockDim.x;
    const int half = ne / 2;

    for (int e = tid; e < ne * ne; e += nt) {
        const int i = e / ne, j = e - i * ne;
        Am[(size_t)i * ld + j] = (i < n && j < n) ? G[(size_t)i * n + j] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double g = 0.0;
        for (int i = 0; i < n; ++i) g = fmax(g, fabs(Am[(size_t)i * ld + i]));
        s_scale = g;
    }
    __syncthreads();
    const double floor_abs = 4.9303806576313238e-32 * s_scale;  // eps^2 * max|diag|

    int rounds = 0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int rd = 0; rd < ne - 1; ++rd, ++rounds) {
            // rotation parameters of this round
            for (int kk = tid; kk < half; kk += nt) {
                int p, q;
                rr_pair(ne, rd, kk, p, q);
                pq[kk] = make_int2(p, q);
                const double app = Am[(size_t)p * ld + p], aqq = Am[(size_t)q * ld + q], apq = Am[(size_t)p * ld + q];
                double cc = 1.0, ss = 0.0;
                if (fabs(apq) > floor_abs && fabs(apq) > rel_tol * sqrt(fabs(app * aqq))) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    cc = 1.0 / sqrt(1.0 + t * t);
                    ss = t * cc;
                    s_rot = 1;  // benign race: every writer stores the same value (a same-address atomic serialises 64 lanes)
                }
                cs[kk] = make_double2(cc, ss);
                rotlog[(size_t)rounds * half + kk] = make_double2(cc, ss);
            }
            __syncthreads();
            // columns: A[:, p], A[:, q] -- one warp per rotation pair, lanes over the rows; 32-bit indexing and a
            // 4-way unrolled body (the solver is instruction bound on its single SM)
            const int wid = tid >> 5, ln = tid & 31, nwp = nt >> 5;
            for (int k = wid; k < half; k += nwp) {
                const int p = pq[k].x, q = pq[k].y;
                const double2 r2 = cs[k];
                if (r2.y != 0.0) {
                    double* colp = Am + p;
                    double* colq = Am + q;
                    int i = ln;
                    for (; i + 96 < ne; i += 128) {
                        double ap[4], aq[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            ap[u] = colp[(i + 32 * u) * ld];
                            aq[u] = colq[(i + 32 * u) * ld];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            colp[(i + 32 * u) * ld] = r2.x * ap[u] - r2.y * aq[u];
                            colq[(i + 32 * u) * ld] = r2.y * ap[u] + r2.x * aq[u];
                        }
                    }
                    for (; i < ne; i += 32) {
                        const double ap = colp[i * ld], aq = colq[i * ld];
                        colp[i * ld] = r2.x * ap - r2.y * aq;
                        colq[i * ld] = r2.y * ap + r2.x * aq;
                    }
                }
            }
            __syncthreads();
            // rows: A[p, :], A[q, :]
            for (int k = wid; k < half; k += nwp) {
                const int p = pq[k].x, q = pq[k].y;
                const double2 r2 = cs[k];
                if (r2.y != 0.0) {
                    double* rowp = Am + p * ld;
                    double* rowq = Am + q * ld;
                    int j = ln;
                    for (; j + 96 < ne; j += 128) {
                        double ap[4], aq[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            ap[u] = rowp[j + 32 * u];
                            aq[u] = rowq[j + 32 * u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            rowp[j + 32 * u] = r2.x * ap[u] - r2.y * aq[u];
                            rowq[j + 32 * u] = r2.y * ap[u] + r2.x * aq[u];
                        }
                    }
                    for (; j < ne; j += 32) {
                        const double ap = rowp[j], aq = rowq[j];
                        rowp[j] = r2.x * ap - r2.y * aq;
                        rowq[j] = r2.y * ap + r2.x * aq;
                    }
                }
            }
            __syncthreads();
        }
        const int nrot = s_rot;
        __syncthreads();
        if (nrot == 0) break;
    }
    if (tid == 0) *nrounds_out = rounds;
    for (int i = tid; i < n; i += nt) lambda[i] = Am[(size_t)i * ld + i];
}


// One-sided (Hestenes) Jacobi for n <= 128: the columns of B = G J_1 J_2 ... are rotated until they are mutually
// orthogonal, then B = V diag(lambda).  Why it is faster than jacobi_kernel: a round of the two-sided kernel reads and
// writes the whole matrix twice (rows and columns: 98 k doubles of shared-memory traffic at n = 128, the 128 B/clk
// port is the limit) behind three block barriers; here a round touches every column once and the three inner
// products of a pair stay in registers.  Eight lanes own a column pair (16-byte loads, two rows per lane, 16-row
// stride), four pairs per warp, Newton-refined MUFU reciprocals / rsqrt instead of the IEEE subroutines.
//
// Ordering (block round-robin): the ne / 4 blocks of four columns play a round-robin tournament; in a super-round a
// warp owns one block pair (X, Y) and works through its 4 x 4 cross pairs in four steps that need only __syncwarp
// (group g keeps column x_g in registers and meets y_{(g + t) mod 4} in step t), so the block barrier comes once per
// FOUR rounds and the warps drift apart inside a super-round: loads, FP64 and the scalar chains of different warps
// overlap instead of marching in lock step.  The six pairs inside each block are done in three steps at the start of
// a sweep.  One sweep = 3 + 4 (ne / 4 - 1) steps of ne / 2 disjoint pairs = every pair once.  The (c, s) log is
// indexed [step][slot = 4 warp + group] (64 slots per step); jacobi1s_pair() maps it back to the columns for the eigenvector replay.
// ne = n padded to a multiple of 16 (zero columns never rotate).  B is left in Bout (column-major, pitch ne).
__host__ __device__ inline int jacobi1s_steps_per_sweep(int ne) { return 3 + 4 * (ne / 4 - 1); }
__device__ __forceinline__ bool jacobi1s_pair(int ne, int step, int slot, int& p, int& q) {
    const int nblk = ne >> 2, w = slot >> 2, g = slot & 3;
    if (step < 3) {
        const int b = w + 16 * (g >> 1);  // two blocks per warp
        if (b >= nblk) return false;
        const int gg = g & 1;
        // step 0: (0,1) (2,3); step 1: (0,2) (1,3); step 2: (0,3) (1,2)
        const int lo = (gg == 0) ? 0 : (step == 0 ? 2 : 1);
        const int hi = (gg == 0) ? (step + 1) : (step == 0 ? 3 : (step == 1 ? 3 : 2));
        p = 4 * b + lo;
        q = 4 * b + hi;
        return true;
    }
    if (w >= (nblk >> 1)) return false;
    const int sr = (step - 3) >> 2, t = (step - 3) & 3;
    int X, Y;
    rr_pair(nblk, sr, w, X, Y);
    p = 4 * X + g;
    q = 4 * Y + ((g + t) & 3);
    return true;
}

// one rotation of the column pair held by an 8-lane group: ap / aq are the group's rows of columns p and q
template <int NUMAX>
__device__ __forceinline__ void jacobi1s_rotate(double (&ap)[2 * NUMAX], double (&aq)[2 * NUMAX], int NU, bool active, double floor_abs,
                                                double rel_tol, double& cc, double& ss) {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, g0 = 0.0, g1 = 0.0;
#pragma unroll
    for (int u = 0; u < NUMAX; ++u) {
        if (u < NU) {
            a0 = fma(ap[2 * u], ap[2 * u], a0), a1 = fma(ap[2 * u + 1], ap[2 * u + 1], a1);
            b0 = fma(aq[2 * u], aq[2 * u], b0), b1 = fma(aq[2 * u + 1], aq[2 * u + 1], b1);
            g0 = fma(ap[2 * u], aq[2 * u], g0), g1 = fma(ap[2 * u + 1], aq[2 * u + 1], g1);
        }
    }
    double alpha = a0 + a1, beta = b0 + b1, gamma = g0 + g1;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
        beta += __shfl_xor_sync(0xffffffffu, beta, o);
        gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
    }
    cc = 1.0;
    ss = 0.0;
    // rotate while |p.q| > rel_tol ||p|| ||q||   (squared: gamma^2 > rel_tol^2 alpha beta)
    if (active && fabs(gamma) > floor_abs && gamma * gamma > rel_tol * rel_tol * (alpha * beta)) {
        // zeta = (beta - alpha) / (2 gamma), t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), c = 1 / sqrt(1 + t^2)
        double zeta = (beta - alpha) * (0.5 * rcp_nr_t<2>(gamma));
        zeta = fmin(fmax(zeta, -1e100), 1e100);
        const double w = fma(zeta, zeta, 1.0);
        const double sq = w * rsqrt_nr_t<2>(w);
        const double t = copysign(rcp_nr_t<2>(fabs(zeta) + sq), zeta);
        cc = rsqrt_nr_t<2>(fma(t, t, 1.0));
        ss = t * cc;
#pragma unroll
        for (int u = 0; u < 2 * NUMAX; ++u) {
            if (u < 2 * NU) {
                const double x = ap[u], y = aq[u];
                ap[u] = cc * x - ss * y;
                aq[u] = ss * x + cc * y;
            }
        }
    }
}

template <int NU_T>  // NU_T = 8: the n = 128 specialisation (no row guards); 0: ne / 16 at run time
__global__ void __launch_bounds__(512) jacobi1s_kernel(const double* __restrict__ G, int n, int ne, double2* __restrict__ rotlog,
                                                       int max_sweeps, double rel_tol, int* __restrict__ nsteps_out,
                                                       double* __restrict__ Bout, unsigned short* __restrict__ pairtab,
                                                       double* __restrict__ colnorm, int allow_direct) {
    constexpr int NUMAX = 8;
    extern __shared__ __align__(16) double Bm[];  // ne x ne, column-major
    __shared__ int s_rot;
    __shared__ double s_scale;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int wid = tid >> 5, ln = tid & 31, g = ln >> 3, r8 = ln & 7;
    const int NU = NU_T ? NU_T : ne / 16, nblk = ne / 4;
    const int slot = 4 * wid + g;

    for (int e = tid; e < ne * ne; e += nt) {
        const int j = e / ne, i = e - j * ne;
        Bm[e] = (i < n && j < n) ? G[(size_t)j * n + i] : 0.0;  // G is symmetric: read it row-wise
    }
    __syncthreads();
    if (tid == 0) {
        double gmax = 0.0;
        for (int i = 0; i < n; ++i) gmax = fmax(gmax, fabs(Bm[i * ne + i]));
        s_scale = gmax;
    }
    __syncthreads();
    const double floor_abs = 4.9303806576313238e-32 * s_scale * s_scale;  // (eps * max|diag|)^2
    // the pair schedule of one sweep for the eigenvector replay: p | q << 8, 0xffff = idle slot
    for (int e = tid; e < jacobi1s_steps_per_sweep(ne) * 64; e += nt) {
        int p, q;
        pairtab[e] = jacobi1s_pair(ne, e >> 6, e & 63, p, q) ? (unsigned short)(p | (q << 8)) : (unsigned short)0xffff;
    }

    auto load_col = [&](int col, double (&a)[2 * NUMAX]) {
        const double* cp = Bm + col * ne + 2 * r8;
#pragma unroll
        for (int u = 0; u < NUMAX; ++u)
            if (u < NU) {
                const double2 x = *reinterpret_cast<const double2*>(cp + 16 * u);
                a[2 * u] = x.x, a[2 * u + 1] = x.y;
            }
    };
    auto store_col = [&](int col, const double (&a)[2 * NUMAX]) {
        double* cp = Bm + col * ne + 2 * r8;
#pragma unroll
        for (int u = 0; u < NUMAX; ++u)
            if (u < NU) *reinterpret_cast<double2*>(cp + 16 * u) = make_double2(a[2 * u], a[2 * u + 1]);
    };

    int step_base = 0;
    const int sps = jacobi1s_steps_per_sweep(ne);
    int rotated = 0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep, step_base += sps) {
        if (tid == 0) s_rot = 0;
        rotated = 0;
        // ---- the pairs inside the blocks
#pragma unroll 1
        for (int st = 0; st < 3; ++st) {
            int p = 0, q = 0;
            const bool active = jacobi1s_pair(ne, st, slot, p, q);
            double ap[2 * NUMAX], aq[2 * NUMAX], cc, ss;
            load_col(p, ap);
            load_col(q, aq);
            jacobi1s_rotate<NUMAX>(ap, aq, NU, active, floor_abs, rel_tol, cc, ss);
            if (ss != 0.0) {
                store_col(p, ap);
                store_col(q, aq);
                rotated = 1;
            }
            if (r8 == 0) rotlog[(size_t)(step_base + st) * 64 + slot] = make_double2(cc, ss);
            __syncwarp();
        }
        __syncthreads();
        // ---- block round-robin
#pragma unroll 1
        for (int sr = 0; sr < nblk - 1; ++sr) {
            int p = 0, q = 0;
            const bool active = jacobi1s_pair(ne, 3 + 4 * sr, slot, p, q);
            const int qb = q - (g & 3);  // first column of block Y
            double ap[2 * NUMAX], aq[2 * NUMAX], cc, ss;
            load_col(p, ap);
            bool xdirty = false;
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                const int qq = active ? qb + ((g + t) & 3) : 0;
                load_col(qq, aq);
                jacobi1s_rotate<NUMAX>(ap, aq, NU, active, floor_abs, rel_tol, cc, ss);
                if (ss != 0.0) {
                    store_col(qq, aq);
                    xdirty = true;
                }
                if (r8 == 0) rotlog[(size_t)(step_base + 3 + 4 * sr + t) * 64 + slot] = make_double2(cc, ss);
                __syncwarp();
            }
            if (xdirty) {
                store_col(p, ap);
                rotated = 1;
            }
            __syncthreads();
        }
        if (rotated) s_rot = 1;  // benign race: every writer stores the same value
        __syncthreads();
        const int nrot = s_rot;
        __syncthreads();
        if (nrot == 0) {
            step_base += sps;
            break;
        }
    }
    if (tid == 0) *nsteps_out = step_base;
    for (int e = tid; e < ne * ne; e += nt) Bout[e] = Bm[e];
    // The converged columns are b_i = G v_i = lambda_i v_i: for a well-conditioned PSD matrix the eigenvectors are simply the
    // normalised columns (error ~ eps lambda_max / lambda_i), and the replay of the rotation log (0.22 ms at n = 128) can be
    // skipped.  nsteps_out[1] = 1 tells the two follow-up kernels to take that route (lambda_min > 1e-2 lambda_max: vectors to ~1e-14).
    __shared__ double s_cn[128];
    for (int col = tid >> 3; col < ne; col += nt >> 3) {
        double a[2 * NUMAX];
        load_col(col, a);
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int u = 0; u < NUMAX; ++u)
            if (u < NU) n0 = fma(a[2 * u], a[2 * u], n0), n1 = fma(a[2 * u + 1], a[2 * u + 1], n1);
        double nn = n0 + n1;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        if (r8 == 0) s_cn[col] = sqrt(nn);
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) colnorm[i] = s_cn[i];
    if (tid == 0) {
        double mn = 1e300, mx = 0.0;
        for (int i = 0; i < n; ++i) mn = fmin(mn, s_cn[i]), mx = fmax(mx, s_cn[i]);
        nsteps_out[1] = (allow_direct && mx > 0.0 && mn > 1e-2 * mx) ? 1 : 0;
    }
}

// V[:, i] = b_i / ||b_i|| when the Jacobi kernel found the matrix well conditioned (see there); V row-major n x n
__global__ void __launch_bounds__(256) eigvec_from_b_kernel(const double* __restrict__ B, int ne, const double* __restrict__ colnorm,
                                                            const int* __restrict__ flags, int n, double* __restrict__ V) {
    if (flags[1] == 0) return;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
        const int j = e / n, i = e - j * n;  // V[j][i] = B[row j of column i]
        V[e] = B[(size_t)i * ne + j] / colnorm[i];
    }
}

// V = J_1 J_2 ... for the one-sided kernel's log: one WARP per row of V (the row lives in warp-private shared memory,
// a step's ne / 2 rotations are disjoint: two slots per lane, one __syncwarp per step)
__global__ void __launch_bounds__(128) jacobi1s_apply_kernel(const double2* __restrict__ rotlog, const int* __restrict__ nsteps,
                                                            const unsigned short* __restrict__ pairtab, int n, int ne,
                                                            double* __restrict__ V) {
    __shared__ double rows[4][128];
    const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + wid;
    if (b >= n || nsteps[1] != 0) return;  // (well-conditioned matrix: the eigenvectors come straight from B)
    double* row = rows[wid];
    for (int j = ln; j < ne; j += 32) row[j] = (j == b) ? 1.0 : 0.0;
    __syncwarp();
    const int sps = jacobi1s_steps_per_sweep(ne);
    const int S = *nsteps;
    // the log lives in L2: fetch DEPTH steps at a time (rotations and pair schedule), then apply them, so that one L2
    // round trip is paid per DEPTH steps
    constexpr int DEPTH = 8;
    int sst = 0;
    for (int st0 = 0; st0 < S; st0 += DEPTH) {
        double2 cs0[DEPTH], cs1[DEPTH];
        unsigned pq0[DEPTH], pq1[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            cs0[d] = cs1[d] = make_double2(1.0, 0.0);
            pq0[d] = pq1[d] = 0xffff;
            if (st0 + d < S) {
                int sd = sst + d;
                sd = sd >= sps ? sd - sps : sd;
                cs0[d] = rotlog[(size_t)(st0 + d) * 64 + ln];
                cs1[d] = rotlog[(size_t)(st0 + d) * 64 + ln + 32];
                pq0[d] = pairtab[sd * 64 + ln];
                pq1[d] = pairtab[sd * 64 + ln + 32];
            }
        }
        sst += DEPTH;
        sst = sst >= sps ? sst - sps : sst;
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (cs0[d].y != 0.0 && pq0[d] != 0xffff) {
                const int p = pq0[d] & 255, q = pq0[d] >> 8;
                const double vp = row[p], vq = row[q];
                row[p] = cs0[d].x * vp - cs0[d].y * vq;
                row[q] = cs0[d].y * vp + cs0[d].x * vq;
            }
            if (cs1[d].y != 0.0 && pq1[d] != 0xffff) {
                const int p = pq1[d] & 255, q = pq1[d] >> 8;
                const double vp = row[p], vq = row[q];
                row[p] = cs1[d].x * vp - cs1[d].y * vq;
                row[q] = cs1[d].y * vp + cs1[d].x * vq;
            }
            __syncwarp();
        }
    }
    for (int j = ln; j < n; j += 32) V[(size_t)b * n + j] = row[j];
}

// lambda_i = v_i^T (G v_i) = v_i . b_i   (B = G V column-major with pitch ne; V row-major n x n, eigenvectors in columns)
__global__ void __launch_bounds__(256) eig_rayleigh_kernel(const double* __restrict__ B, int ne, const double* __restrict__ V, int n,
                                                           double* __restrict__ lambda) {
    const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
    for (int i = blockIdx.x * (blockDim.x >> 5) + wid; i < n; i += gridDim.x * (blockDim.x >> 5)) {
        double acc = 0.0;
        for (int j = ln; j < n; j += 32) acc = fma(V[(size_t)j * n + i], B[(size_t)i * ne + j], acc);
        acc = warp_sum(acc);
        if (ln == 0) lambda[i] = acc;
    }
}

// V = J_1 J_2 ... (rows of the identity transformed independently): CTA b owns row b of V.
__global__ void __launch_bounds__(256) jacobi_apply_kernel(const double2* __restrict__ rotlog, const int* __restrict__ nrounds,
                                                           int n, int ne, double* __restrict__ V) {
    extern __shared__ double row[];
    const int half = ne / 2;
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < ne; j += blockDim.x) row[j] = (j == b) ? 1.0 : 0.0;
    __syncthreads();
    const int R = *nrounds;
    // a round's ne / 2 rotations are disjoint: the threads of the CTA share them (any ne / 2, not just <= blockDim.x)
    for (int rd = 0; rd < R; ++rd) {
        for (int k = threadIdx.x; k < half; k += blockDim.x) {
            const double2 cur = rotlog[(size_t)rd * half + k];
            if (cur.y != 0.0) {
                int p, q;
                rr_pair(ne, rd % (ne - 1), k, p, q);
                const double vp = row[p], vq = row[q];
                row[p] = cur.x * vp - cur.y * vq;
                row[q] = cur.y * vp + cur.x * vq;
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) V[(size_t)b * n + j] = row[j];
}

// sort eigenvalues descending (rank by counting), permute the columns of V accordingly.  Any number of CTAs: every CTA
// ranks all eigenvalues itself (n^2 compares) and moves its slice of V (one CTA took 35 us at n = 128).
__global__ void __launch_bounds__(256) eig_sort_kernel(const double* __restrict__ lam_in, const double* __restrict__ Vin, int n,
                                                       double* __restrict__ lam_out, double* __restrict__ Vout) {
    extern __shared__ int rank_of[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double li = lam_in[i];
        int rk = 0;
        for (int j = 0; j < n; ++j) {
            const double lj = lam_in[j];
            rk += (lj > li) || (lj == li && j < i);
        }
        rank_of[i] = rk;
        if (blockIdx.x == 0) lam_out[rk] = li;
    }
    __syncthreads();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
        const int r = e / n, cidx = e - r * n;
        Vout[(size_t)r * n + rank_of[cidx]] = Vin[e];
    }
}

// s = sqrt(max(lambda, 0)); M2 = V diag(1/s for s > tol else 0); Vt = V^T; rank = #(s > tol).  Any number of CTAs (one CTA
// spent 48 us on 16 384 square roots and divisions at n = 128): the singular values are formed once per CTA in shared memory.
__global__ void __launch_bounds__(256) svd_post_kernel(const double* __restrict__ lam, const double* __restrict__ V, int n,
                                                       double tol, double* __restrict__ s, double* __restrict__ M2,
                                                       double* __restrict__ Vt, int* __restrict__ rank) {
    extern __shared__ double sv[];
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double sj = sqrt(fmax(lam[j], 0.0));
        sv[j] = sj;
        if (blockIdx.x == 0) {
            s[j] = sj;
            if (sj > tol) atomicAdd(&cnt, 1);
        }
    }
    __syncthreads();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
        const int i = e / n, j = e - i * n;
        const double sj = sv[j];
        const double v = V[e];
        M2[e] = (sj > tol) ? v / sj : 0.0;
        Vt[(size_t)j * n + i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *rank = cnt;
}

// ------------------------------------------------------------------ small triangular helpers (single CTA)
// Rinv = R^{-1} for an upper-triangular R (n x n, row-major); column-parallel back substitution.
__global__ void __launch_bounds__(256) triu_inverse_kernel(const double* __restrict__ R, int n, double* __restrict__ Rinv) {
    for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < n; col += gridDim.x * blockDim.x) {
        for (int i = n - 1; i >= 0; --i) {
            double acc = (i == col) ? 1.0 : 0.0;
            if (i > col) {
                Rinv[(size_t)i * n + col] = 0.0;
                continue;
            }
            for (int k = i + 1; k <= col; ++k) acc = fma(-R[(size_t)i * n + k], Rinv[(size_t)k * n + col], acc);
            Rinv[(size_t)i * n + col] = acc / R[(size_t)i * n + i];
        }
    }
}
// make diag(R) > 0: R <- S R, and return S (n) so that Q <- Q S
__global__ void __launch_bounds__(256) sign_fix_R_kernel(double* __restrict__ R, int n, double* __restrict__ sgn) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double sg = (R[(size_t)i * n + i] < 0.0) ? -1.0 : 1.0;
        sgn[i] = sg;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n;
        R[e] *= sgn[i];
    }
}
__global__ void __launch_bounds__(256) scale_cols_kernel(double* __restrict__ M, int n, const double* __restrict__ sgn) {
    // M <- M * diag(sgn)   (n x n)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) M[e] *= sgn[e % n];
}

// Upper Cholesky factor R (R^T R = G) of an n x n (n <= 128) SPD matrix, one CTA, right-looking in shared memory.
// stat[0] = 1 if a pivot is not safely positive (breakdown), stat[1] = max_j G_jj / min_j R_jj^2  (~ cond(A)^2).
__global__ void __launch_bounds__(1024) chol_upper_kernel(const double* __restrict__ G, int n, double* __restrict__ R,
                                                          double* __restrict__ stat) {
    extern __shared__ double chol_sm[];
    double (*S)[129] = reinterpret_cast<double (*)[129]>(chol_sm);
    __shared__ double s_dmax, s_pmin;
    __shared__ int s_bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < n * n; e += nt) S[e / n][e % n] = G[e];
    if (tid == 0) {
        s_dmax = 0.0;
        s_pmin = 1e300;
        s_bad = 0;
    }
    __syncthreads();
    if (tid == 0) {
        double dm = 0.0;
        for (int j = 0; j < n; ++j) dm = fmax(dm, S[j][j]);
        s_dmax = dm;
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        const double piv = S[j][j];
        if (!(piv > 0.0)) {  // also catches NaN
            if (tid == 0) s_bad = 1;
            break;           // uniform: every thread reads the same S[j][j]
        }
        const double d = sqrt(piv);
        __syncthreads();
        if (tid == 0) s_pmin = fmin(s_pmin, piv);
        for (int cidx = j + tid; cidx < n; cidx += nt) S[j][cidx] = (cidx == j) ? d : S[j][cidx] / d;
        __syncthreads();
        const int rem = n - j - 1;
        for (int e = tid; e < rem * rem; e += nt) {
            const int i = j + 1 + e / rem, cidx = j + 1 + e % rem;
            if (cidx >= i) S[i][cidx] = fma(-S[j][i], S[j][cidx], S[i][cidx]);
        }
        __syncthreads();
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, cidx = e % n;
        R[e] = (cidx >= i) ? S[i][cidx] : 0.0;
    }
    if (tid == 0) {
        stat[0] = s_bad ? 1.0 : 0.0;
        stat[1] = s_bad ? 1e300 : s_dmax / s_pmin;
    }
}

constexpr size_t CHOL_SMEM = 128 * 129 * sizeof(double);

// ---- register-resident versions for n <= 128 (the CholeskyQR2 path of TSQR runs two of each per call; the generic
// kernels above cost 0.20 + 0.36 ms, a fifth of the whole 2^20 x 128 TSQR)
//
// Inverse of an upper-triangular R: warp w solves the columns w, w + 32, w + 64, w + 96 of R X = I together (four
// independent recurrences interleaved); lane l holds rows l + 32 u of each right-hand side in registers, R sits in
// shared memory (odd pitch: a column walk is conflict free), 1 / R_kk is formed once.  Column-oriented back
// substitution: x_k is broadcast from its owner lane, every lane updates its rows above k.
// The 32 warps (4 columns each) are independent: they run as 8 CTAs of 4 warps on 8 SMs (one CTA of 32 warps was bound by
// the issue rate of a single SM: 45 us; every CTA stages its own copy of R from L2).
__global__ void __launch_bounds__(128) triu_inverse128_kernel(const double* __restrict__ R, int n, double* __restrict__ Rinv) {
    extern __shared__ double ti_sm[];
    double (*S)[129] = reinterpret_cast<double (*)[129]>(ti_sm);
    __shared__ double dinv[128];
    const int tid = threadIdx.x, w = blockIdx.x * (blockDim.x >> 5) + (tid >> 5), ln = tid & 31;
    for (int e = tid; e < 128 * 128; e += blockDim.x) {
        const int i = e >> 7, k = e & 127;
        S[i][k] = (i < n && k < n) ? R[(size_t)i * n + k] : ((i == k) ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid < 128) dinv[tid] = 1.0 / S[tid][tid];
    __syncthreads();
    double x[4][4];  // [column c = w + 32 cc][row block u]: row ln + 32 u
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int u = 0; u < 4; ++u) x[cc][u] = (ln + 32 * u == w + 32 * cc) ? 1.0 : 0.0;
#pragma unroll
    for (int ub = 3; ub >= 0; --ub) {
        for (int kk = 31; kk >= 0; --kk) {
            const int k = 32 * ub + kk;
            const double dk = dinv[k];
            double rk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) rk[u] = (u <= ub) ? S[ln + 32 * u][k] : 0.0;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                if (cc < ub) continue;  // column w + 32 cc < 32 (cc + 1) <= k: x_k = 0 (compile-time skip of whole blocks)
                const double xk = __shfl_sync(0xffffffffu, x[cc][ub], kk) * dk;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (u < ub) x[cc][u] = fma(-rk[u], xk, x[cc][u]);
                    else if (u == ub) x[cc][u] = (ln == kk) ? xk : ((ln < kk) ? fma(-rk[u], xk, x[cc][u]) : x[cc][u]);
                }
            }
        }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int col = w + 32 * cc;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = ln + 32 * u;
            if (i < n && col < n) Rinv[(size_t)i * n + col] = (i <= col) ? x[cc][u] : 0.0;
        }
    }
}

// Upper Cholesky factor for n <= 128 with the matrix in registers: thread (ty, tx) of a 32 x 32 block owns the
// elements (ty + 32 a, tx + 32 b).  Per column: the diagonal owner publishes the pivot, the owners of row j publish
// the scaled row, everybody updates its 16 elements (8 shared loads, 16 FMAs, no index arithmetic).  Same outputs
// as chol_upper_kernel.
__global__ void __launch_bounds__(1024) chol_upper128_kernel(const double* __restrict__ G, int n, double* __restrict__ R,
                                                             double* __restrict__ stat) {
    __shared__ double rowbuf[2][128];
    __shared__ double s_piv[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    double a[4][4];
#pragma unroll
    for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) {
            const int i = ty + 32 * ia, k = tx + 32 * ib;
            a[ia][ib] = (i < n && k < n) ? G[(size_t)i * n + k] : ((i == k) ? 1.0 : 0.0);
        }
    double dmax = 0.0, pmin = 1e300;
    if (threadIdx.x == 0) {
        for (int j = 0; j < n; ++j) dmax = fmax(dmax, G[(size_t)j * n + j]);
    }
    int bad = 0;
#pragma unroll
    for (int ja = 0; ja < 4; ++ja) {
        for (int jj = 0; jj < 32; ++jj) {
            const int j = 32 * ja + jj, par = j & 1;
            if (ty == jj && tx == jj) s_piv[par] = a[ja][ja];
            __syncthreads();
            const double piv = s_piv[par];
            if (!(piv > 0.0)) {  // also catches NaN; uniform
                bad = 1;
                break;
            }
            if (j < n) pmin = fmin(pmin, piv);  // (rows >= n are identity padding)
            const double rd = rsqrt_nr_t<2>(piv);
            if (ty == jj) {
                // my elements of row j, scaled; columns left of the diagonal are not part of R
#pragma unroll
                for (int ib = 0; ib < 4; ++ib) {
                    const int k = tx + 32 * ib;
                    double v = a[ja][ib] * rd;
                    if (k == j) v = piv * rd;
                    if (k < j) v = 0.0;
                    a[ja][ib] = v;
                    rowbuf[par][k] = v;
                }
            }
            __syncthreads();
            double ri[4], rk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ri[q] = rowbuf[par][ty + 32 * q];
                rk[q] = rowbuf[par][tx + 32 * q];
            }
#pragma unroll
            for (int ia = 0; ia < 4; ++ia) {
                if (ia < ja) continue;
                const int i = ty + 32 * ia;
#pragma unroll
                for (int ib = 0; ib < 4; ++ib) {
                    if (ib < ja) continue;
                    if (i > j) a[ia][ib] = fma(-ri[ia], rk[ib], a[ia][ib]);  // rows below j (columns left of j see rk = 0)
                }
            }
        }
        if (bad) break;
    }
#pragma unroll
    for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) {
            const int i = ty + 32 * ia, k = tx + 32 * ib;
            if (i < n && k < n) R[(size_t)i * n + k] = (k >= i) ? a[ia][ib] : 0.0;
        }
    if (threadIdx.x == 0) {
        stat[0] = bad ? 1.0 : 0.0;
        stat[1] = bad ? 1e300 : dmax / pmin;
    }
}

int launch_triu_inverse(Ctx* c, const double* R, int n, double* Rinv) {
    if (n <= 128) {
        static DeviceLatch configured;
        if (!configured.test(c->device)) {
            LQ_CUDA(c, cudaFuncSetAttribute(triu_inverse128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 129 * 8));
            configured.set(c->device);
        }
        triu_inverse128_kernel<<<8, 128, 128 * 129 * 8, c->stream>>>(R, n, Rinv);
    } else {
        triu_inverse_kernel<<<1, 256, 0, c->stream>>>(R, n, Rinv);
    }
    LQ_CHECK_LAUNCH(c);
    return LQ_OK;
}


__global__ void __launch_bounds__(256) scale_cols_tall_kernel(double* __restrict__ M, long long rows, int n,
                                                              const double* __restrict__ sgn) {
    const long long total = rows * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        M[e] *= sgn[e % n];
}

int eigh_configure(Ctx* c) {
    static DeviceLatch done;
    if (done.test(c->device)) return LQ_OK;
    LQ_CUDA(c, cudaFuncSetAttribute(chol_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_SMEM));
    cudaFuncAttributes fa{};
    LQ_CUDA(c, cudaFuncGetAttributes(&fa, jacobi_kernel));
    LQ_CUDA(c, cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    c->max_smem - (int)fa.sharedSizeBytes));
    LQ_CUDA(c, cudaFuncSetAttribute(jacobi1s_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8));
    LQ_CUDA(c, cudaFuncSetAttribute(jacobi1s_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8));
    LQ_CUDA(c, cudaFuncGetAttributes(&fa, tsqr_leaf_kernel<16, 8>));
    LQ_CUDA(c, cudaFuncSetAttribute(tsqr_leaf_kernel<16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    c->max_smem - (int)fa.sharedSizeBytes));
    done.set(c->device);
    return LQ_OK;
}

// R factor (n x n upper) of A (m x n, lda) by a tree of streaming Householder reductions
int tsqr_rfactor(Ctx* c, const double* A, int lda, long long m, int n, double* R /* n x n */) {
    using Cfg = StreamCfg<4, 16, 8>;
    const size_t smem = Cfg::smem_doubles(n) * sizeof(double);
    const double* cur = A;
    int cur_ld = lda;
    long long cur_m = m;
    DevBuf bufs[2];
    int flip = 0;
    for (int level = 0; level < 8; ++level) {
        // leaves of >= 4 blocks each, at most 2 CTAs per SM worth of leaves
        long long min_rows = (long long)Cfg::BLOCK_ROWS * 8;
        long long ctas = std::max<long long>(1, std::min<long long>((long long)c->sm_count, cur_m / min_rows));
        if (level > 0) ctas = std::max<long long>(1, std::min<long long>(ctas, cur_m / ((long long)n * 8)));
        long long rpc = (cur_m + ctas - 1) / ctas;
        rpc = (rpc + Cfg::BLOCK_ROWS - 1) / Cfg::BLOCK_ROWS * Cfg::BLOCK_ROWS;
        ctas = (cur_m + rpc - 1) / rpc;
        double* out;
        if (ctas == 1) {
            out = R;
        } else {
            LQ_TRY(bufs[flip].alloc(c, sizeof(double) * (size_t)ctas * n * n));
            out = bufs[flip].as<double>();
        }
        tsqr_leaf_kernel<16, 8><<<(unsigned)ctas, 256, smem, c->stream>>>(cur, cur_ld, cur_m, n, rpc, out);
        LQ_CHECK_LAUNCH(c);
        LQ_COUNT_LAUNCH(c);
        if (ctas == 1) return LQ_OK;
        cur = out;
        cur_ld = n;
        cur_m = ctas * (long long)n;
        flip ^= 1;
    }
    set_error(c, "tsqr_rfactor: reduction tree did not terminate");
    return LQ_ERR_ARG;
}

}  // namespace

// ------------------------------------------------------------------ public (internal) ops
// G = A^T A (A m x n with row stride lda): symmetric rank-k kernel (upper-triangle blocks only, syrk.cu) when the shape
// allows, the general TN GEMM otherwise
static int gram_ld(Ctx* c, const double* A, int lda, long long m, int n, double* G) {
    const int rc = syrk_tn(c, A, lda, m, n, G, n);
    if (rc != LQ_ERR_UNSUPPORTED) return rc;
    return gemm(c, true, false, n, n, (int)m, 1.0, A, lda, A, lda, 0.0, G, n);
}

int gram(Ctx* c, const double* A, long long m, int n, double* G) { return gram_ld(c, A, n, m, n, G); }

// CTAs for the n x n element-wise kernels behind the eigen-solver (two elements per thread at n = 128)
static unsigned small_grid(Ctx* c, int n) {
    return (unsigned)std::max<long long>(1, std::min<long long>(((long long)n * n + 511) / 512, (long long)c->sm_count * 2));
}

int eigh_jacobi(Ctx* c, const double* G, int n, double* lambda_desc, double* V, bool psd) {
    LQ_REQUIRE(c, n >= 1 && n <= 2048, LQ_ERR_UNSUPPORTED, "eigen-solver supports n <= 2048 (got %d)", n);
    LQ_TRY(eigh_configure(c));
    const int max_sweeps = 40;
    // a pair is rotated while |a_pq| > rel_tol * sqrt(a_pp a_qq); 4 eps stops the tail of sweeps that only chase
    // rounding noise (eigenvalues move by O(a_pq^2 / gap), the eigenvector basis stays a product of exact rotations)
    double rel_tol = 4.0 * 1.1102230246251565e-16;
    static const double env_tol = getenv("LINALG_B200_JACOBI_TOL") ? atof(getenv("LINALG_B200_JACOBI_TOL")) : 0.0;  // read once
    if (env_tol > 0.0) rel_tol = env_tol * 1.1102230246251565e-16;
    DevBuf Aw, rot, nr, lam, Vraw, ptab;
    LQ_TRY(nr.alloc(c, 16));
    LQ_TRY(lam.alloc(c, sizeof(double) * n));
    LQ_TRY(Vraw.alloc(c, sizeof(double) * (size_t)n * n));
    if (n <= 128 && !c->env_jacobi_two_sided) {
        // one-sided kernel: columns in shared memory, padded to a multiple of 16
        const int ne = (n + 15) / 16 * 16, half = ne / 2;
        LQ_TRY(Aw.alloc(c, sizeof(double) * (size_t)ne * ne));
        LQ_TRY(rot.alloc(c, sizeof(double2) * (size_t)(max_sweeps + 1) * jacobi1s_steps_per_sweep(ne) * 64));
        LQ_TRY(ptab.alloc(c, sizeof(unsigned short) * (size_t)jacobi1s_steps_per_sweep(ne) * 64));
        const int direct = (psd && !LQ_ENV_ONCE("LINALG_B200_JACOBI_ALWAYS_REPLAY")) ? 1 : 0;
        if (ne == 128)
            jacobi1s_kernel<8><<<1, 512, sizeof(double) * (size_t)ne * ne, c->stream>>>(G, n, ne, rot.as<double2>(), max_sweeps, rel_tol,
                                                                                       nr.as<int>(), Aw.as<double>(), ptab.as<unsigned short>(), lam.as<double>(), direct);
        else
            jacobi1s_kernel<0><<<1, 512, sizeof(double) * (size_t)ne * ne, c->stream>>>(G, n, ne, rot.as<double2>(), max_sweeps, rel_tol,
                                                                                       nr.as<int>(), Aw.as<double>(), ptab.as<unsigned short>(), lam.as<double>(), direct);
        LQ_CHECK_LAUNCH(c);
        jacobi1s_apply_kernel<<<(n + 3) / 4, 128, 0, c->stream>>>(rot.as<double2>(), nr.as<int>(), ptab.as<unsigned short>(), n, ne,
                                                                   Vraw.as<double>());
        LQ_CHECK_LAUNCH(c);
        if (direct)
            eigvec_from_b_kernel<<<small_grid(c, n), 256, 0, c->stream>>>(Aw.as<double>(), ne, lam.as<double>(), nr.as<int>(), n, Vraw.as<double>());
        LQ_CHECK_LAUNCH(c);
        eig_rayleigh_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(Aw.as<double>(), ne, Vraw.as<double>(), n, lam.as<double>());
        LQ_CHECK_LAUNCH(c);
        eig_sort_kernel<<<small_grid(c, n), 256, sizeof(int) * n, c->stream>>>(lam.as<double>(), Vraw.as<double>(), n, lambda_desc, V);
        LQ_CHECK_LAUNCH(c);
        c->launches += 5;
        return LQ_OK;
    }
    const int ne = (n + 1) & ~1, ld = ne | 1, half = ne / 2;
    const size_t need = ((size_t)ne * ld + 4) * sizeof(double) + (size_t)half * sizeof(double2) + 64;
    const int use_global = need + 9216 > (size_t)c->max_smem;  // 9 KB of static shared memory in the kernel
    LQ_TRY(Aw.alloc(c, use_global ? sizeof(double) * (size_t)ne * ld : 16));
    LQ_TRY(rot.alloc(c, sizeof(double2) * (size_t)max_sweeps * (ne - 1 > 0 ? ne - 1 : 1) * (half > 0 ? half : 1)));
    const size_t smem = use_global ? (size_t)half * sizeof(double2) + 64 : need;
    jacobi_kernel<<<1, 1024, smem, c->stream>>>(G, n, Aw.as<double>(), use_global, rot.as<double2>(), max_sweeps, rel_tol,
                                               nr.as<int>(), lam.as<double>());
    LQ_CHECK_LAUNCH(c);
    jacobi_apply_kernel<<<n, 256, sizeof(double) * (ne + 2), c->stream>>>(rot.as<double2>(), nr.as<int>(), n, ne, Vraw.as<double>());
    LQ_CHECK_LAUNCH(c);
    eig_sort_kernel<<<small_grid(c, n), 256, sizeof(int) * n, c->stream>>>(lam.as<double>(), Vraw.as<double>(), n, lambda_desc, V);
    LQ_CHECK_LAUNCH(c);
    c->launches += 3;
    return LQ_OK;
}

int svd_gram_local(Ctx* c, const double* A, long long m, int n, double tol, double* U, double* s, double* Vt,
                   int* rank_host, bool sharded) {
    LQ_REQUIRE(c, m >= 1 && n >= 1, LQ_ERR_SHAPE, "svd: bad shape %lld x %d", m, n);
    LQ_REQUIRE(c, m < (1LL << 31), LQ_ERR_SHAPE, "svd: more than 2^31 rows per device not supported");
    DevBuf G, lam, V, M2, rk;
    LQ_TRY(G.alloc(c, sizeof(double) * (size_t)n * n));
    LQ_TRY(lam.alloc(c, sizeof(double) * n));
    LQ_TRY(V.alloc(c, sizeof(double) * (size_t)n * n));
    LQ_TRY(M2.alloc(c, sizeof(double) * (size_t)n * n));
    LQ_TRY(rk.alloc(c, 16));
    LQ_TRY(gram(c, A, m, n, G.as<double>()));                                  // svd.py:42
    if (sharded) LQ_TRY(comm_allreduce_sum(c, G.as<double>(), (long long)n * n));
    LQ_TRY(eigh_jacobi(c, G.as<double>(), n, lam.as<double>(), V.as<double>(), true)); // svd.py:46-51 (G = A^T A is PSD)
    svd_post_kernel<<<small_grid(c, n), 256, sizeof(double) * n, c->stream>>>(lam.as<double>(), V.as<double>(), n, tol, s, M2.as<double>(), Vt,
                                                                             rk.as<int>());
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    LQ_TRY(gemm(c, false, false, m, n, n, 1.0, A, n, M2.as<double>(), n, 0.0, U, n));  // svd.py:61-64
    if (rank_host) {
        LQ_CUDA(c, cudaMemcpyAsync(rank_host, rk.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LQ_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return LQ_OK;
}

// TSQR: R by a Householder reduction tree (leaves stream their rows, R factors are combined),
// Q = A R^{-1} followed by one TSQR refinement pass (restores orthogonality to O(eps) whenever
// cond(A) * eps << 1), signs normalised so that diag(R) > 0 -- the MGS convention (qr.py:39-42).
static int tsqr_householder(Ctx* c, const double* A, long long m, int n, double* Q, double* R, bool sharded) {
    LQ_TRY(eigh_configure(c));
    DevBuf R1, R2, Rinv, Rall, sgn, Q1;
    const size_t nn = sizeof(double) * (size_t)n * n;
    LQ_TRY(R1.alloc(c, nn));
    LQ_TRY(R2.alloc(c, nn));
    LQ_TRY(Rinv.alloc(c, nn));
    LQ_TRY(sgn.alloc(c, sizeof(double) * n));
    LQ_TRY(Q1.alloc(c, sizeof(double) * (size_t)m * n));
    if (sharded && c->nranks > 1) LQ_TRY(Rall.alloc(c, nn * c->nranks));

    auto reduce_R = [&](const double* X, double* Rout) -> int {
        LQ_TRY(tsqr_rfactor(c, X, n, m, n, Rout));
        if (sharded && c->nranks > 1) {
            // exchange step: all-gather the local R factors, every rank reduces the same stack
            LQ_TRY(comm_allgather(c, Rout, Rall.as<double>(), (long long)n * n));
            LQ_TRY(tsqr_rfactor(c, Rall.as<double>(), n, (long long)c->nranks * n, n, Rout));
        }
        return LQ_OK;
    };
    // pass 1
    LQ_TRY(reduce_R(A, R1.as<double>()));
    LQ_TRY(launch_triu_inverse(c, R1.as<double>(), n, Rinv.as<double>()));
    LQ_CHECK_LAUNCH(c);
    LQ_TRY(gemm(c, false, false, m, n, n, 1.0, A, n, Rinv.as<double>(), n, 0.0, Q1.as<double>(), n));
    // pass 2 (refinement)
    LQ_TRY(reduce_R(Q1.as<double>(), R2.as<double>()));
    sign_fix_R_kernel<<<1, 256, 0, c->stream>>>(R2.as<double>(), n, sgn.as<double>());   // R2 <- S R2 (temporarily, S folded below)
    LQ_CHECK_LAUNCH(c);
    // we want R = S' (R2 R1) with diag > 0.  Compute R = R2 R1 first with the unsigned R2:
    //   sign_fix made diag(R2) > 0; R1's diagonal sign decides the final flip.
    LQ_TRY(gemm(c, false, false, n, n, n, 1.0, R2.as<double>(), n, R1.as<double>(), n, 0.0, R, n));
    // Q = Q1 (S R2)^{-1} ... then flip so that diag(R) > 0
    LQ_TRY(launch_triu_inverse(c, R2.as<double>(), n, Rinv.as<double>()));
    LQ_CHECK_LAUNCH(c);
    sign_fix_R_kernel<<<1, 256, 0, c->stream>>>(R, n, sgn.as<double>());                  // R <- S2 R
    scale_cols_kernel<<<grid_for(c, (long long)n * n), 256, 0, c->stream>>>(Rinv.as<double>(), n, sgn.as<double>());  // Rinv <- Rinv S2
    LQ_CHECK_LAUNCH(c);
    c->launches += 5;
    LQ_TRY(gemm(c, false, false, m, n, n, 1.0, Q1.as<double>(), n, Rinv.as<double>(), n, 0.0, Q, n));
    return LQ_OK;
}

// Fast path for well-conditioned tall-skinny matrices: CholeskyQR2 on the FP64 tensor pipe.
//   G1 = A^T A (+ all-reduce over the row shards) -> R1 = chol(G1) -> Q1 = A R1^{-1}
//   G2 = Q1^T Q1 (+ all-reduce)                   -> R2 = chol(G2) -> Q  = Q1 R2^{-1},  R = R2 R1
// Four streaming GEMMs instead of two column-by-column Householder sweeps; diag(R) > 0 by construction (the
// convention of linalg/qr.py:39-42).  The second pass restores orthogonality to O(eps) as long as
// cond(A)^2 * eps << 1; the Cholesky kernel reports a cond(A)^2 estimate and the Householder TSQR tree below
// takes over when it exceeds 1e10 (or a pivot breaks down).  *used = 0 means "fall back".
__global__ void tsqr_flags_kernel(double* f, double a, double b) {
    f[0] = a;
    f[1] = b;
}

// status words of one tsqr_cholqr2 call, identical on every rank (they derive from all-reduced data)
struct CholQrStatus {
    double chol1_bad = 0, cond2 = 0, chol2_bad = 0, cond2_pass2 = 0;
    double ranks_short = 0;  // number of ranks whose local block has fewer rows than columns
    bool ok() const { return chol1_bad == 0.0 && cond2 < 1e10 && chol2_bad == 0.0; }
};

// scratch of the CholeskyQR2 pipeline (allocated once per call, reused by every column panel of tsqr_wide)
struct CholQrScratch {
    DevBuf G, R1, R2, Rinv, Q1;
    int alloc(Ctx* c, long long m, int n) {
        const size_t nn = sizeof(double) * (size_t)n * n;
        LQ_TRY(G.alloc(c, nn + 2 * sizeof(double)));
        LQ_TRY(R1.alloc(c, nn));
        LQ_TRY(R2.alloc(c, nn));
        LQ_TRY(Rinv.alloc(c, nn));
        LQ_TRY(Q1.alloc(c, sizeof(double) * (size_t)m * n));
        return LQ_OK;
    }
};

// Enqueue one CholeskyQR2 of the m x n block A (row stride lda, n <= 128) -> Q (ldq), R (ldr); nothing is read back.
// stat[0..3] (device) receive chol1_bad, cond^2 estimate, chol2_bad, cond^2 of pass 2.  `short_flag` >= 0 puts that value
// into the two spare words behind G so that it is summed over the ranks by the first all-reduce (G[n*n]).
static int cholqr2_enqueue(Ctx* c, const double* A, int lda, long long m, int n, double* Q, int ldq, double* R, int ldr,
                           bool sharded, CholQrScratch& w, double* stat, double short_flag) {
    const bool new_chol = n <= 128 && !c->env_old_chol;
    double* G = w.G.as<double>();
    const bool carry = short_flag >= 0.0;
    if (carry) {
        tsqr_flags_kernel<<<1, 1, 0, c->stream>>>(G + (size_t)n * n, short_flag, 0.0);
        LQ_CHECK_LAUNCH(c);
    }
    // pass 1
    LQ_TRY(gram_ld(c, A, lda, m, n, G));
    if (sharded) LQ_TRY(comm_allreduce_sum(c, G, (long long)n * n + (carry ? 2 : 0)));
    if (new_chol) chol_upper128_kernel<<<1, 1024, 0, c->stream>>>(G, n, w.R1.as<double>(), stat);
    else chol_upper_kernel<<<1, 1024, CHOL_SMEM, c->stream>>>(G, n, w.R1.as<double>(), stat);
    LQ_CHECK_LAUNCH(c);
    LQ_TRY(launch_triu_inverse(c, w.R1.as<double>(), n, w.Rinv.as<double>()));
    LQ_TRY(gemm(c, false, false, m, n, n, 1.0, A, lda, w.Rinv.as<double>(), n, 0.0, w.Q1.as<double>(), n));
    // pass 2
    LQ_TRY(gram_ld(c, w.Q1.as<double>(), n, m, n, G));
    if (sharded) LQ_TRY(comm_allreduce_sum(c, G, (long long)n * n));
    if (new_chol) chol_upper128_kernel<<<1, 1024, 0, c->stream>>>(G, n, w.R2.as<double>(), stat + 2);
    else chol_upper_kernel<<<1, 1024, CHOL_SMEM, c->stream>>>(G, n, w.R2.as<double>(), stat + 2);
    LQ_CHECK_LAUNCH(c);
    LQ_TRY(launch_triu_inverse(c, w.R2.as<double>(), n, w.Rinv.as<double>()));
    c->launches += 4 + (carry ? 1 : 0);
    LQ_TRY(gemm(c, false, false, n, n, n, 1.0, w.R2.as<double>(), n, w.R1.as<double>(), n, 0.0, R, ldr));
    LQ_TRY(gemm(c, false, false, m, n, n, 1.0, w.Q1.as<double>(), n, w.Rinv.as<double>(), n, 0.0, Q, ldq));
    return LQ_OK;
}

static int tsqr_cholqr2(Ctx* c, const double* A, long long m, int n, double* Q, double* R, bool sharded, CholQrStatus* st) {
    // Optimistic pipeline: nothing is read back until the last kernel is queued (round 1 synchronised with the host after
    // the first Cholesky to look at the cond^2 estimate: a bubble in the middle of a 4 ms call).  The status words and the
    // "short rank" count ride in the first all-reduce / are computed from all-reduced data, so every rank reaches the same
    // verdict and a fallback is taken by all ranks together (the path decision must be collective).
    LQ_TRY(eigh_configure(c));
    CholQrScratch w;
    DevBuf stat;
    LQ_TRY(w.alloc(c, m, n));
    LQ_TRY(stat.alloc(c, 4 * sizeof(double)));
    LQ_TRY(cholqr2_enqueue(c, A, n, m, n, Q, n, R, n, sharded, w, stat.as<double>(), m < n ? 1.0 : 0.0));
    // the one read-back of the call, behind everything else on the stream
    double h[6] = {0, 0, 0, 0, 0, 0};
    LQ_CUDA(c, cudaMemcpyAsync(h, stat.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaMemcpyAsync(h + 4, w.G.as<double>() + (size_t)n * n, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaStreamSynchronize(c->stream));
    st->chol1_bad = h[0];
    st->cond2 = (h[1] == h[1]) ? h[1] : 1e300;  // NaN -> ill-conditioned
    st->chol2_bad = h[2];
    st->cond2_pass2 = h[3];
    st->ranks_short = h[4];
    return LQ_OK;
}

__global__ void add_inplace_kernel(double* __restrict__ dst, int ldd, const double* __restrict__ src, int lds, int rows, int cols) {
    const long long total = (long long)rows * cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / cols), k = (int)(e % cols);
        dst[(size_t)i * ldd + k] += src[(size_t)i * lds + k];
    }
}

// a7 for n > 128 (e.g. the reference's own benchmark shape 5000 x 1000, linalg/benchmark_qr.py:17): block classical
// Gram-Schmidt with re-orthogonalisation (BCGS2) over 128-column panels, each panel factored by CholeskyQR2 --
//   W  = A_k - Q_prev S1,  S1 = Q_prev^T A_k        (first projection)
//   W' = W   - Q_prev S2,  S2 = Q_prev^T W          (second projection: orthogonality to O(eps))
//   W' = Q_k R_kk (CholeskyQR2, diag > 0)           R[prev, k] = S1 + S2
// everything is GEMMs on the FP64 tensor pipe; row-sharded calls all-reduce S1, S2 and the panel Gram matrices.  The
// status words of all panels are read back once at the end; any ill-conditioned panel sends the whole call to the
// reflector path (single GPU) or fails alike on every rank (sharded).
static int tsqr_wide(Ctx* c, const double* A, long long m, int n, double* Q, double* R, bool sharded) {
    const bool sh = sharded && c->nranks > 1;
    constexpr int PB = 128;
    const int np = (n + PB - 1) / PB;
    bool ok = !c->env_tsqr_householder && (sh || m >= 2LL * n);
    if (ok) {
        LQ_TRY(eigh_configure(c));
        CholQrScratch w;
        DevBuf stat, S;
        LQ_TRY(w.alloc(c, m, PB));
        LQ_TRY(stat.alloc(c, sizeof(double) * 4 * np));
        LQ_TRY(S.alloc(c, sizeof(double) * (size_t)n * PB));
        LQ_CUDA(c, cudaMemsetAsync(R, 0, sizeof(double) * (size_t)n * n, c->stream));
        LQ_CUDA(c, cudaMemcpyAsync(Q, A, sizeof(double) * (size_t)m * n, cudaMemcpyDeviceToDevice, c->stream));
        for (int k = 0; k < np; ++k) {
            const int c0 = k * PB, bk = std::min(PB, n - c0);
            double* Qk = Q + c0;
            for (int pass = 0; pass < 2 && c0 > 0; ++pass) {
                LQ_TRY(gemm(c, true, false, c0, bk, (int)m, 1.0, Q, n, Qk, n, 0.0, S.as<double>(), bk));
                if (sh) LQ_TRY(comm_allreduce_sum(c, S.as<double>(), (long long)c0 * bk));
                LQ_TRY(gemm(c, false, false, m, bk, c0, -1.0, Q, n, S.as<double>(), bk, 1.0, Qk, n));
                add_inplace_kernel<<<grid_for(c, (long long)c0 * bk), 256, 0, c->stream>>>(R + c0, n, S.as<double>(), bk, c0, bk);
                LQ_CHECK_LAUNCH(c);
                LQ_COUNT_LAUNCH(c);
            }
            LQ_TRY(cholqr2_enqueue(c, Qk, n, m, bk, Qk, n, R + (size_t)c0 * n + c0, n, sh, w, stat.as<double>() + 4 * k, -1.0));
        }
        std::vector<double> h(4 * (size_t)np);
        LQ_CUDA(c, cudaMemcpyAsync(h.data(), stat.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
        LQ_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < np; ++k)
            if (!(h[4 * k] == 0.0 && h[4 * k + 1] < 1e10 && h[4 * k + 2] == 0.0)) ok = false;
        if (ok) return LQ_OK;
    }
    LQ_REQUIRE(c, !sh, LQ_ERR_UNSUPPORTED,
               "tsqr (sharded, n = %d > 128): a column panel is too ill-conditioned for CholeskyQR2 and the reflector tree "
               "supports n <= 128", n);
    LQ_REQUIRE(c, m >= n, LQ_ERR_SHAPE, "tsqr needs m >= n (got %lld x %d)", m, n);
    DevBuf sgn;
    LQ_TRY(sgn.alloc(c, sizeof(double) * n));
    LQ_TRY(blocked_householder_qr(c, A, (int)m, n, Q, R));
    sign_fix_R_kernel<<<1, 256, 0, c->stream>>>(R, n, sgn.as<double>());
    scale_cols_tall_kernel<<<grid_for(c, m * n), 256, 0, c->stream>>>(Q, m, n, sgn.as<double>());
    LQ_CHECK_LAUNCH(c);
    c->launches += 2;
    return LQ_OK;
}

// a7: thin QR of a tall-skinny matrix with diag(R) > 0.
int tsqr_local(Ctx* c, const double* A, long long m, int n, double* Q, double* R, bool sharded) {
    const bool sh = sharded && c->nranks > 1;
    // Row-sharded calls validate only what every rank sees alike (n); the rank-local row count is checked through the
    // all-reduced "short rank" count so that no rank leaves while its peers wait in a collective.
    LQ_REQUIRE(c, n >= 1 && m >= 1 && (sh || m >= n), LQ_ERR_SHAPE, "tsqr needs m >= n >= 1 (got %lld x %d)", m, n);
    LQ_REQUIRE(c, m < (1LL << 31), LQ_ERR_SHAPE, "tsqr: more than 2^31 rows per device not supported");
    if (n > 128) return tsqr_wide(c, A, m, n, Q, R, sharded);
    CholQrStatus st;
    bool have_status = false;
    if (!c->env_tsqr_householder && (sh || m >= 4LL * n)) {
        LQ_TRY(tsqr_cholqr2(c, A, m, n, Q, R, sh, &st));
        if (st.ok()) return LQ_OK;
        have_status = true;
    }
    if (!sh) {
        // robust path for ill-conditioned (or nearly square) input: the blocked compact-WY Householder QR, whose
        // Q is formed from the reflectors (residual and orthogonality O(eps) for any cond(A)), flipped to diag(R) > 0
        DevBuf sgn;
        LQ_TRY(sgn.alloc(c, sizeof(double) * n));
        LQ_TRY(blocked_householder_qr(c, A, (int)m, n, Q, R));
        sign_fix_R_kernel<<<1, 256, 0, c->stream>>>(R, n, sgn.as<double>());
        scale_cols_tall_kernel<<<grid_for(c, m * n), 256, 0, c->stream>>>(Q, m, n, sgn.as<double>());
        LQ_CHECK_LAUNCH(c);
        c->launches += 2;
        return LQ_OK;
    }
    // row-sharded and ill-conditioned: Householder reduction tree + Q = A R^{-1} with one refinement pass
    // (orthogonality O(eps), residual O(cond(A) eps)); its leaves need at least n rows on every rank
    if (!have_status) {
        DevBuf f;
        LQ_TRY(f.alloc(c, 2 * sizeof(double)));
        tsqr_flags_kernel<<<1, 1, 0, c->stream>>>(f.as<double>(), m < n ? 1.0 : 0.0, 0.0);
        LQ_CHECK_LAUNCH(c);
        LQ_TRY(comm_allreduce_sum(c, f.as<double>(), 2));
        double h[2] = {0, 0};
        LQ_CUDA(c, cudaMemcpyAsync(h, f.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        LQ_CUDA(c, cudaStreamSynchronize(c->stream));
        st.ranks_short = h[0];
    }
    LQ_REQUIRE(c, st.ranks_short == 0.0, LQ_ERR_SHAPE,
               "tsqr (sharded, reflector path): %d rank(s) hold fewer than n = %d rows; re-shard the rows", (int)st.ranks_short, n);
    return tsqr_householder(c, A, m, n, Q, R, sharded);
}


// ------------------------------------------------------------------ counter-based standard-normal generator
// Philox4x32-10 (Salmon et al., SC'11: counter = element group, key = seed) + Box-Muller; four doubles per counter.
// The stream depends only on (seed, element index): any grid shape and any device produce the same numbers.
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n2 = hi0 ^ ctr[3] ^ k1;
        ctr[0] = n0;
        ctr[1] = lo1;
        ctr[2] = n2;
        ctr[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__global__ void __launch_bounds__(256) philox_normal_kernel(double* __restrict__ out, long long count, unsigned long long seed) {
    const long long groups = (count + 3) / 4;
    for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (long long)gridDim.x * blockDim.x) {
        double z[4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t ctr[4] = {(uint32_t)gidx, (uint32_t)((unsigned long long)gidx >> 32), (uint32_t)half, 0x6c696e61u};
            philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
            // two uniforms in (0, 1] / [0, 1) with 53 / 32 bits, one Box-Muller pair
            const unsigned long long b = ((unsigned long long)ctr[0] << 32) | ctr[1];
            const double u1 = ((double)(b >> 11) + 1.0) * (1.0 / 9007199254740992.0);
            const double u2 = ((double)ctr[2] * 4294967296.0 + (double)ctr[3]) * (1.0 / 18446744073709551616.0);
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            z[2 * half] = rad * cs;
            z[2 * half + 1] = rad * sn;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (4 * gidx + e < count) out[4 * gidx + e] = z[e];
    }
}

}  // namespace lq

using namespace lq;

extern "C" {

int lq_gram_dev(lq_ctx* h, const double* A, int64_t m, int n, double* G) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, m >= 1 && n >= 1 && m < (1LL << 31), LQ_ERR_SHAPE, "gram: bad shape");
    LQ_CUDA(c, cudaSetDevice(c->device));
    return gram(c, A, m, n, G);
}
int lq_eigh_dev(lq_ctx* h, const double* G, int n, double* lambda_desc, double* V) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, n >= 1 && n <= 2048, LQ_ERR_SHAPE, "eigh: n must be in [1, 2048] (got %d)", n);
    LQ_CUDA(c, cudaSetDevice(c->device));
    return eigh_jacobi(c, G, n, lambda_desc, V);
}
int lq_svd_gram_dev(lq_ctx* h, const double* A, int64_t m, int n, double tol, double* U, double* s, double* Vt,
                    int* rank_host) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, m >= n, LQ_ERR_SHAPE, "svd_gram needs m >= n (transpose on the host side, svd.py:37-39)");
    LQ_CUDA(c, cudaSetDevice(c->device));
    return svd_gram_local(c, A, m, n, tol, U, s, Vt, rank_host, false);
}
int lq_svd_gram_sharded_dev(lq_ctx* h, const double* A, int64_t m, int n, double tol, double* U, double* s, double* Vt,
                            int* rank_host) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_CUDA(c, cudaSetDevice(c->device));
    return svd_gram_local(c, A, m, n, tol, U, s, Vt, rank_host, true);
}
int lq_svd_gram(lq_ctx* h, const double* A, int64_t m, int n, double tol, double* U, double* s, double* Vt,
                int* rank_host) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, m >= n && n >= 1, LQ_ERR_SHAPE, "svd_gram needs m >= n >= 1");
    LQ_CUDA(c, cudaSetDevice(c->device));
    const size_t mn = sizeof(double) * (size_t)m * n, nn = sizeof(double) * (size_t)n * n;
    DevBuf dA, dU, ds, dVt;
    LQ_TRY(dA.alloc(c, mn));
    LQ_TRY(dU.alloc(c, mn));
    LQ_TRY(ds.alloc(c, sizeof(double) * n));
    LQ_TRY(dVt.alloc(c, nn));
    LQ_CUDA(c, cudaMemcpyAsync(dA.p, A, mn, cudaMemcpyHostToDevice, c->stream));
    LQ_TRY(svd_gram_local(c, dA.as<double>(), m, n, tol, dU.as<double>(), ds.as<double>(), dVt.as<double>(), rank_host, false));
    LQ_CUDA(c, cudaMemcpyAsync(U, dU.p, mn, cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaMemcpyAsync(s, ds.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaMemcpyAsync(Vt, dVt.p, nn, cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaStreamSynchronize(c->stream));
    return LQ_OK;
}
// linalg/svd.py:67-76 on the device: orthonormal completion of U (host, m x n, first `rank` columns
// valid) from the candidate directions Z: drawn by the caller (host, m x (n - rank): upstream draws them from the global
// np.random) or, with a seed, generated on the device (no upload; deterministic -- SURVEY.md section 8f-1).
static int svd_complete_impl(lq_ctx* h, double* U, int64_t m, int n, int rank, const double* Z, bool seeded, unsigned long long seed) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    const int k = n - rank;
    LQ_REQUIRE(c, m >= n && rank >= 0 && k >= 1 && U && (Z || seeded), LQ_ERR_SHAPE, "svd_complete: bad arguments");
    LQ_REQUIRE(c, m < (1LL << 31), LQ_ERR_SHAPE, "svd_complete: too many rows");
    LQ_CUDA(c, cudaSetDevice(c->device));
    DevBuf dU, dZ, dQ, dR, dW;
    LQ_TRY(dU.alloc(c, sizeof(double) * (size_t)m * n));
    LQ_TRY(dZ.alloc(c, sizeof(double) * (size_t)m * k));
    LQ_TRY(dQ.alloc(c, sizeof(double) * (size_t)m * k));
    LQ_TRY(dR.alloc(c, sizeof(double) * (size_t)k * k));
    LQ_TRY(dW.alloc(c, sizeof(double) * (size_t)n * k));
    LQ_CUDA(c, cudaMemcpyAsync(dU.p, U, sizeof(double) * (size_t)m * n, cudaMemcpyHostToDevice, c->stream));
    if (seeded) {
        const long long count = (long long)m * k;
        philox_normal_kernel<<<grid_for(c, (count + 3) / 4), 256, 0, c->stream>>>(dZ.as<double>(), count, seed);
        LQ_CHECK_LAUNCH(c);
        LQ_COUNT_LAUNCH(c);
    } else {
        LQ_CUDA(c, cudaMemcpyAsync(dZ.p, Z, sizeof(double) * (size_t)m * k, cudaMemcpyHostToDevice, c->stream));
    }
    LQ_TRY(lq_householder_qr_dev(h, dZ.as<double>(), (int)m, k, dQ.as<double>(), dR.as<double>()));          // svd.py:69
    if (rank > 0) {
        // Q -= U_r (U_r^T Q)                                                                                // svd.py:71-72
        LQ_TRY(gemm(c, true, false, rank, k, (int)m, 1.0, dU.as<double>(), n, dQ.as<double>(), k, 0.0, dW.as<double>(), k));
        LQ_TRY(gemm(c, false, false, m, k, rank, -1.0, dU.as<double>(), n, dW.as<double>(), k, 1.0, dQ.as<double>(), k));
    }
    LQ_TRY(lq_householder_qr_dev(h, dQ.as<double>(), (int)m, k, dZ.as<double>(), dR.as<double>()));          // svd.py:74
    // U[:, rank:] = Z
    LQ_CUDA(c, cudaMemcpy2DAsync(U + rank, sizeof(double) * n, dZ.p, sizeof(double) * k, sizeof(double) * k, (size_t)m,
                                 cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaStreamSynchronize(c->stream));
    return LQ_OK;
}
int lq_svd_complete(lq_ctx* h, double* U, int64_t m, int n, int rank, const double* Z) {
    return svd_complete_impl(h, U, m, n, rank, Z, false, 0ULL);
}
int lq_svd_complete_seeded(lq_ctx* h, double* U, int64_t m, int n, int rank, uint64_t seed) {
    return svd_complete_impl(h, U, m, n, rank, nullptr, true, (unsigned long long)seed);
}
int lq_random_normal_dev(lq_ctx* h, double* out, int64_t count, uint64_t seed) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, count >= 0 && (count == 0 || out), LQ_ERR_ARG, "random_normal: bad arguments");
    if (count == 0) return LQ_OK;
    LQ_CUDA(c, cudaSetDevice(c->device));
    philox_normal_kernel<<<grid_for(c, (count + 3) / 4), 256, 0, c->stream>>>(out, count, (unsigned long long)seed);
    LQ_CHECK_LAUNCH(c);
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

int lq_tsqr_dev(lq_ctx* h, const double* A, int64_t m, int n, double* Q, double* R) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_CUDA(c, cudaSetDevice(c->device));
    return tsqr_local(c, A, m, n, Q, R, false);
}
int lq_tsqr_sharded_dev(lq_ctx* h, const double* A, int64_t m, int n, double* Q, double* R) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_CUDA(c, cudaSetDevice(c->device));
    return tsqr_local(c, A, m, n, Q, R, true);
}
int lq_tsqr(lq_ctx* h, const double* A, int64_t m, int n, double* Q, double* R) {
    Ctx* c = as_ctx(h);
    if (!c) return LQ_ERR_ARG;
    LQ_REQUIRE(c, m >= n && n >= 1, LQ_ERR_SHAPE, "tsqr needs m >= n >= 1");
    LQ_CUDA(c, cudaSetDevice(c->device));
    const size_t mn = sizeof(double) * (size_t)m * n, nn = sizeof(double) * (size_t)n * n;
    DevBuf dA, dQ, dR;
    LQ_TRY(dA.alloc(c, mn));
    LQ_TRY(dQ.alloc(c, mn));
    LQ_TRY(dR.alloc(c, nn));
    LQ_CUDA(c, cudaMemcpyAsync(dA.p, A, mn, cudaMemcpyHostToDevice, c->stream));
    LQ_TRY(tsqr_local(c, dA.as<double>(), m, n, dQ.as<double>(), dR.as<double>(), false));
    LQ_CUDA(c, cudaMemcpyAsync(Q, dQ.p, mn, cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaMemcpyAsync(R, dR.p, nn, cudaMemcpyDeviceToHost, c->stream));
    LQ_CUDA(c, cudaStreamSynchronize(c->stream));
    return LQ_OK;
}

}  // extern "C"
