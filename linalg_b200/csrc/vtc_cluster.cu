// W2 = op(T) * (V^T C)  in ONE launch: split-K over a thread-block cluster with a DSMEM reduction.
//
// This is the first half of every block-reflector application  C <- C - V op(T) V^T C  of the blocked
// Householder QR (linalg/qr.py:89,91 in compact-WY form).  V is mk x kb (kb <= 128 reflectors), C is mk x nc.
// The product V^T C is skinny (kb x nc) with a long contraction (mk rows), so one 128 x 128 output tile is
// computed by a CLUSTER of CS CTAs that split the rows; each CTA runs the FP64 tensor-core main loop of
// gemm.cu (mma.sync m16n8k8.f64 fed by bulk-async row copies through a 4-stage mbarrier ring), parks its
// partial tile in its own shared memory, and after one cluster barrier CTA r sums COLUMN slice r of all
// peers through distributed shared memory (ld.shared::cluster).  Owning whole columns, it can apply the
// triangular factor op(T) on the spot and write W2.  No partial sums ever reach global memory and the three
// launches (split-K GEMM, reduction, T multiply) of the generic path collapse into one -- these products sit
// on the latency-critical panel chain of the factorisation.
#include <algorithm>

#include "../../include/linalg_b200.h"
#include "ops.cuh"

namespace lq {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int PM = 132;                 // pitch of a [k][row] tile (16 one-row bulk copies per stage)
constexpr int TILE_DOUBLES = 16 * PM;   // 2112
constexpr int VTC_THREADS = 288;        // 8 consumer warps + 1 producer warp
constexpr int RED_DOUBLES = BM * BN;    // partial tile parked for the cluster reduction (reuses the ring)
constexpr int WCOL_PITCH = 33;
constexpr size_t RING_BYTES = (size_t)STAGES * 2 * TILE_DOUBLES * sizeof(double);               // 135168
constexpr size_t VTC_SMEM = RING_BYTES + 2 * STAGES * sizeof(uint64_t) + (size_t)BM * WCOL_PITCH * sizeof(double) + 64;
static_assert(RING_BYTES >= RED_DOUBLES * sizeof(double), "ring must hold the partial tile");

struct VtcArgs {
    const double* V;
    const double* C;
    const double* T;
    double* W2;
    int ldv, ldc, ldt;
    int kb, nc, mk;
    int mode;  // 0: W2 = V^T C;  1: W2 = T^T (V^T C);  2: W2 = T (V^T C)
};

__global__ void __maxnreg__(224) vtc_cluster_kernel(const VtcArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + RING_BYTES);
    uint64_t* empty = full + STAGES;
    double* wcol = reinterpret_cast<double*>(smem_raw + RING_BYTES + 2 * STAGES * sizeof(uint64_t));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CS = (int)cluster_nctarank();
    const int rank = (int)cluster_ctarank();
    const int n0 = blockIdx.x * BN;
    const int nvalid = min(BN, g.nc - n0);
    const int KT = g.mk / BK;
    const int per = (KT + CS - 1) / CS;
    const int kt0 = rank * per;
    const int nkt = max(0, min(KT, kt0 + per) - kt0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);
        }
        mbar_fence_init();
    }
    __syncthreads();

    double acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0;
    const int wm = warp >> 2, wn = warp & 3;
    const int gq = lane >> 2, tq = lane & 3;
    // row bands of the 128-row tile that hold reflectors at all (kb may be 32 / 64 / 96 / 128)
    const int im_hi = min(4, max(0, (g.kb - wm * 64 + 15) / 16));

    if (warp == 8) {
        const uint32_t bytes = (uint32_t)(BK * (g.kb + nvalid) * 8);
        for (int it = 0; it < nkt; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            double* sA = tiles + (size_t)s * 2 * TILE_DOUBLES;
            double* sB = sA + TILE_DOUBLES;
            const long long k0 = (long long)(kt0 + it) * BK;
            if (lane == 0) mbar_expect_tx(&full[s], bytes);
            __syncwarp();
            if (lane < BK) bulk_g2s(sA + lane * PM, g.V + (k0 + lane) * g.ldv, g.kb * 8, &full[s]);
            else bulk_g2s(sB + (lane - 16) * PM, g.C + (k0 + lane - 16) * g.ldc + n0, nvalid * 8, &full[s]);
        }
    } else {
        for (int it = 0; it < nkt; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            const double* sA = tiles + (size_t)s * 2 * TILE_DOUBLES;
            const double* sB = sA + TILE_DOUBLES;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                double af[4][4], bf[4][2];
                const int kA = ks * 8 + tq;
#pragma unroll
                for (int im = 0; im < 4; ++im) {
                    const int r = wm * 64 + im * 16 + gq;
                    af[im][0] = sA[kA * PM + r];
                    af[im][1] = sA[kA * PM + r + 8];
                    af[im][2] = sA[(kA + 4) * PM + r];
                    af[im][3] = sA[(kA + 4) * PM + r + 8];
                }
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
                    const int c = wn * 32 + jn * 8 + gq;
                    bf[jn][0] = sB[kA * PM + c];
                    bf[jn][1] = sB[(kA + 4) * PM + c];
                }
#pragma unroll
                for (int im = 0; im < 4; ++im) {
                    if (im < im_hi) {
#pragma unroll
                        for (int jn = 0; jn < 4; ++jn) dmma_16x8x8(acc[im][jn], af[im], bf[jn]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    __syncthreads();  // every bulk copy has landed and been consumed: the ring is free

    // ---- park the partial tile (rows >= kb or columns >= nvalid hold stale data and are never summed)
    double* red = tiles;  // [128][128]
    if (warp < 8) {
#pragma unroll
        for (int im = 0; im < 4; ++im)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = wm * 64 + im * 16 + gq + half * 8;
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
                    const int c = wn * 32 + jn * 8 + 2 * tq;
                    *reinterpret_cast<double2*>(red + r * BN + c) =
                        make_double2(acc[im][jn][half * 2 + 0], acc[im][jn][half * 2 + 1]);
                }
            }
    }
    cluster_sync_all();

    // ---- CTA `rank` sums column slice [rank * CSW, (rank + 1) * CSW) over all peers
    const int CSW = BN / CS;  // 8 .. 128 columns
    if (threadIdx.x < 256) {
        const int row = threadIdx.x & 127, half = threadIdx.x >> 7;  // 2 threads per row
        const int cper = CSW / 2;                                    // columns per thread (>= 4)
        const uint32_t base = smem_u32(red + row * BN + rank * CSW + half * cper);
        for (int c4 = 0; c4 < cper; c4 += 4) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (row < g.kb) {
                for (int p = 0; p < CS; ++p) {
                    const uint32_t a = mapa_shared(base + (uint32_t)c4 * 8u, (uint32_t)p);
                    s0 += ld_cluster_f64(a);
                    s1 += ld_cluster_f64(a + 8);
                    s2 += ld_cluster_f64(a + 16);
                    s3 += ld_cluster_f64(a + 24);
                }
            }
            double* dst = wcol + row * WCOL_PITCH + half * cper + c4;
            dst[0] = s0;
            dst[1] = s1;
            dst[2] = s2;
            dst[3] = s3;
        }
    }
    cluster_sync_all();  // peers are done reading my partial tile; wcol is complete CTA-wide

    // ---- apply op(T) to my columns and write W2
    if (threadIdx.x < 256) {
        const int i = threadIdx.x & 127, half = threadIdx.x >> 7;
        const int cper = CSW / 2;
        const int cbase = half * cper;
        if (i < g.kb) {
            for (int c4 = 0; c4 < cper; c4 += 4) {
                double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;
                if (g.mode == 0) {
                    const double* w = wcol + i * WCOL_PITCH + cbase + c4;
                    o0 = w[0]; o1 = w[1]; o2 = w[2]; o3 = w[3];
                } else {
                    // T is upper triangular: T^T[i][k] = T[k][i] (k <= i);  T[i][k] (k >= i)
                    const int klo = (g.mode == 1) ? 0 : i;
                    const int khi = (g.mode == 1) ? i + 1 : g.kb;
#pragma unroll 4
                    for (int k = klo; k < khi; ++k) {
                        const double tv = (g.mode == 1) ? g.T[(long long)k * g.ldt + i] : g.T[(long long)i * g.ldt + k];
                        const double* w = wcol + k * WCOL_PITCH + cbase + c4;
                        o0 = fma(tv, w[0], o0);
                        o1 = fma(tv, w[1], o1);
                        o2 = fma(tv, w[2], o2);
                        o3 = fma(tv, w[3], o3);
                    }
                }
                const int c = rank * CSW + cbase + c4;
                double* out = g.W2 + (long long)i * g.nc + n0 + c;
                if (c + 0 < nvalid) out[0] = o0;
                if (c + 1 < nvalid) out[1] = o1;
                if (c + 2 < nvalid) out[2] = o2;
                if (c + 3 < nvalid) out[3] = o3;
            }
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool vtc_cluster_supported(int kb, int nc, int mk, const double* V, int ldv, const double* Cm, int ldc) {
    if (LQ_ENV_ONCE("LINALG_B200_NO_VTC_CLUSTER")) return false;
    return kb >= 2 && kb <= 128 && (kb % 2 == 0) && nc >= 2 && (nc % 2 == 0) && mk >= BK && (mk % BK == 0) &&
           aligned16(V) && aligned16(Cm) && (ldv % 2 == 0) && (ldc % 2 == 0);
}

// mode 0: W2 = V^T C; 1: W2 = T^T (V^T C); 2: W2 = T (V^T C).   W2 is kb x nc with leading dimension nc.
int vtc_cluster(Ctx* c, int mode, int kb, int nc, int mk, const double* V, int ldv, const double* Cm, int ldc, const double* T,
                int ldt, double* W2) {
    static int max_cs[64] = {};
    if (!max_cs[c->device]) {
        LQ_CUDA(c, cudaFuncSetAttribute(vtc_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VTC_SMEM));
        LQ_CUDA(c, cudaFuncSetAttribute(vtc_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        int best = 1;
        for (int cs : {16, 8, 4, 2}) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(1, 1, cs);
            cfg.blockDim = dim3(VTC_THREADS);
            cfg.dynamicSmemBytes = VTC_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 1;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = cs;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, vtc_cluster_kernel, &cfg) == cudaSuccess && n >= 1) {
                best = cs;
                break;
            }
            cudaGetLastError();
        }
        if (const char* env = getenv("LINALG_B200_VTC_MAX_CLUSTER")) best = std::max(1, std::min(best, atoi(env)));
        max_cs[c->device] = best;
    }
    const int KT = mk / BK;
    int cs = max_cs[c->device];
    if (cs < 4) return LQ_ERR_UNSUPPORTED;    // the column-slice buffer assumes <= 32 columns per CTA
    while (cs > 4 && KT < 6 * cs) cs >>= 1;   // at least ~6 k-tiles per CTA (idle ranks just contribute zeros)
    VtcArgs g;
    g.V = V; g.C = Cm; g.T = T; g.W2 = W2; g.ldv = ldv; g.ldc = ldc; g.ldt = ldt;
    g.kb = kb; g.nc = nc; g.mk = mk; g.mode = mode;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((nc + BN - 1) / BN, 1, cs);
    cfg.blockDim = dim3(VTC_THREADS);
    cfg.dynamicSmemBytes = VTC_SMEM;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = cs;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    LQ_CUDA(c, cudaLaunchKernelEx(&cfg, vtc_cluster_kernel, g));
    LQ_COUNT_LAUNCH(c);
    return LQ_OK;
}

}  // namespace lq
