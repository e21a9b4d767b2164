// Damped normal-equation solve  (A + u I) x = g  for the LM step, replacing the reference's dense
// torch.linalg.cholesky + cholesky_solve (cuSOLVER potrf + 2 trsv)   /root/reference/super/LM.py:38-51,97-100.
//
// J^T J of an ED graph is block-banded once the nodes are ordered along the longer image axis
// (measured at config 1: half-bandwidth 42 node blocks = 300 scalars of n = 1862; profiles/r1a).
// A dense factorisation costs n^3/3 = 2.15 GF and, more importantly, ~60 dependent kernel phases in
// cuSOLVER (826 us + 313 us of triangular solves per LM iteration, 82 % of the frame).  Here:
//
//   * storage: lower band, row-major,  AB[i*ldab + (j - i + bw)]  for max(0,i-bw) <= j <= i;
//   * ONE thread-block cluster (8 or 16 CTAs) runs the whole right-looking blocked factorisation,
//     the forward substitution (the right-hand side rides along as an extra matrix row) and the
//     back substitution in a single launch; CTAs hand panels over through L2 and synchronise with the
//     hardware cluster barrier (~0.2 us) instead of kernel boundaries / grid syncs;
//   * per 32-column panel: CTA 0 factors the 32x32 diagonal block inside one warp's registers
//     (pivot chain: rsqrt -> shuffle broadcast -> own-row FMA, ~100 cycles per column), solves the
//     rows below with every thread owning a row, then all CTAs apply the rank-32 update to the
//     trailing window with 4x4 register tiles out of shared memory.
//
// FP64 throughout (parity tolerance on beta is 1e-4 with cond(A) up to ~1e6 late in the schedule).
#include <cooperative_groups.h>

#include "common.cuh"
#include "super_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int NB = 32;            // panel width
constexpr int PS = NB + 1;        // shared-memory row stride (doubles): conflict-free column access
constexpr int BC_THREADS = 256;

struct BandArgs {
    double* AB;        // (n, ldab) lower band, overwritten by L
    int ldab, n, bw;
    double* g;         // (n) rhs in, solution out
    const double* u;   // device scalar added to the diagonal (may be null)
    double* dinv;      // (n) scratch: reciprocal pivots
    int* info;         // device flag: set to 1 when a pivot is not positive (never cleared here)
    int debug;         // timing experiments only: 1 skip trailing update, 2 skip back substitution, 4 skip panel math
};

__device__ __forceinline__ double* ab_at(const BandArgs& a, int i, int j) {
    return a.AB + (size_t)i * a.ldab + (j - i + a.bw);
}

// ---- Phase A pieces (CTA 0) --------------------------------------------------------------------------
// Compile-time recursion instead of nested `#pragma unroll` loops: every index into the per-lane row is
// a constant, so the 32 doubles stay in registers (nvcc left a partially unrolled loop + a local-memory
// array on the pivot chain otherwise).
template <int C, int T>
struct PotrfInner {
    static __device__ __forceinline__ void run(double (&a)[NB], int lane) {
        const double ltc = __shfl_sync(0xffffffffu, a[C], T);          // l_tc lives in lane t
        if (lane >= T) a[T] = fma(-a[C], ltc, a[T]);
        PotrfInner<C, T + 1>::run(a, lane);
    }
};
template <int C>
struct PotrfInner<C, NB> {
    static __device__ __forceinline__ void run(double (&)[NB], int) {}
};

template <int C>
struct PotrfStep {
    static __device__ __forceinline__ void run(double (&a)[NB], int lane, int* info, double& my_inv) {
        // pivot lane: reciprocal square root of its (fully updated) diagonal entry
        double inv = 0.0;
        if (lane == C) {
            const double piv = a[C];
            if (!(piv > 0.0)) *info = 1;
            inv = rsqrt(piv);
            a[C] = piv * inv;            // l_cc
            my_inv = inv;
        }
        inv = __shfl_sync(0xffffffffu, inv, C);
        if (lane > C) a[C] *= inv;       // l_rc
        PotrfInner<C, C + 1>::run(a, lane);
        PotrfStep<C + 1>::run(a, lane, info, my_inv);
    }
};
template <>
struct PotrfStep<NB> {
    static __device__ __forceinline__ void run(double (&)[NB], int, int*, double&) {}
};

// Cholesky of the 32x32 diagonal block held as one row per lane.  Returns reciprocal pivot of this lane's row.
__device__ __forceinline__ double warp_potrf32(double (&a)[NB], int lane, int* info) {
    double my_inv = 0.0;
    PotrfStep<0>::run(a, lane, info, my_inv);
    return my_inv;
}

// x = row * L_kk^{-T} for one row held in registers (right-looking, independent FMAs per column).
template <int C, int T>
struct TrsmInner {
    static __device__ __forceinline__ void run(double (&row)[NB], double x, const double* __restrict__ Ps) {
        row[T] = fma(-x, Ps[T * PS + C], row[T]);
        TrsmInner<C, T + 1>::run(row, x, Ps);
    }
};
template <int C>
struct TrsmInner<C, NB> {
    static __device__ __forceinline__ void run(double (&)[NB], double, const double* __restrict__) {}
};
template <int C>
struct TrsmStep {
    static __device__ __forceinline__ void run(double (&row)[NB], const double* __restrict__ Ps,
                                               const double* __restrict__ sdinv) {
        const double x = row[C] * sdinv[C];
        row[C] = x;
        TrsmInner<C, C + 1>::run(row, x, Ps);
        TrsmStep<C + 1>::run(row, Ps, sdinv);
    }
};
template <>
struct TrsmStep<NB> {
    static __device__ __forceinline__ void run(double (&)[NB], const double* __restrict__, const double* __restrict__) {}
};

__global__ void __launch_bounds__(BC_THREADS, 1) band_chol_kernel(BandArgs a) {
    extern __shared__ double smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, bw = a.bw;
    double* Ps = smem;                               // (NB + bw + 1) x PS panel rows (+ rhs row)
    double* sdinv = smem + (size_t)(NB + bw + 1) * PS;   // NB reciprocal pivots
    const double u = a.u ? *a.u : 0.0;

    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);                       // real columns in this panel
        const int m = min(n - k0, NB + bw);                   // window rows (matrix rows k0 .. k0+m-1)
        const int mt = m - nb;                                // trailing rows below the panel
        const int mr = max(m, NB);                            // smem row that carries the right-hand side
        // ================= Phase A: CTA 0 factors the panel ==========================================
        if (rank == 0) {
            // load panel rows (zero outside the band / above the diagonal), rhs as row index mr.
            // Thread -> (column c = tid % 32, rows tid/32 + 16 k); 8 independent L2 loads in flight per thread.
            {
                const int c = tid & 31, r0 = tid >> 5;
                for (int rb = r0; rb <= mr; rb += 8 * (BC_THREADS / 32)) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int r = rb + q * (BC_THREADS / 32);
                        v[q] = 0.0;
                        if (r > mr) continue;
                        if (r == mr) {
                            if (c < nb) v[q] = __ldcg(a.g + k0 + c);
                        } else if (r >= m || c >= nb) {
                            v[q] = (r == c) ? 1.0 : 0.0;      // identity padding of a short last panel
                        } else {
                            const int i = k0 + r, j = k0 + c;
                            if (j <= i && i - j <= bw) v[q] = __ldcg(ab_at(a, i, j));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int r = rb + q * (BC_THREADS / 32);
                        if (r > mr) continue;
                        if (r < m && r == c && c < nb) v[q] += u;   // additive damping (LM.py:97)
                        Ps[r * PS + c] = v[q];
                    }
                }
            }
            __syncthreads();
            if (warp == 0 && !(a.debug & 4)) {
                double row[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) row[c] = Ps[lane * PS + c];
                const double inv = warp_potrf32(row, lane, a.info);
#pragma unroll
                for (int c = 0; c < NB; ++c) Ps[lane * PS + c] = (c <= lane) ? row[c] : 0.0;
                sdinv[lane] = inv;
                if (lane < nb) a.dinv[k0 + lane] = inv;
            }
            __syncthreads();
            // rows below the diagonal block (and the rhs row): x = row * L_kk^{-T}, one row per thread
            for (int r = NB + tid; r <= mr && !(a.debug & 4); r += BC_THREADS) {
                double row[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) row[c] = Ps[r * PS + c];
                TrsmStep<0>::run(row, Ps, sdinv);
#pragma unroll
                for (int c = 0; c < NB; ++c) Ps[r * PS + c] = row[c];
            }
            __syncthreads();
            // write L back: diagonal block + rows below (band positions only) + forward-substituted rhs
            for (int e = tid; e < (mr + 1) * NB; e += BC_THREADS) {
                const int r = e / NB, c = e % NB;
                if (c >= nb) continue;
                const double v = Ps[r * PS + c];
                if (r == mr) { a.g[k0 + c] = v; continue; }
                if (r >= m) continue;                          // padding rows of a short panel
                const int i = k0 + r, j = k0 + c;
                if (j <= i && i - j <= bw) __stcg(ab_at(a, i, j), v);
            }
        }
        cluster.sync();
        // ================= Phase B: all CTAs update the trailing window ===============================
        if (mt > 0) {
            // stage the panel rows below the diagonal block (mt x nb) + the rhs panel in shared memory
            {
                const int c = tid & 31, r0 = tid >> 5;
                for (int rb = r0; rb < mt; rb += 8 * (BC_THREADS / 32)) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int r = rb + q * (BC_THREADS / 32);
                        const int i = k0 + nb + r, j = k0 + c;
                        v[q] = (r < mt && c < nb && i - j <= bw) ? __ldcg(ab_at(a, i, j)) : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int r = rb + q * (BC_THREADS / 32);
                        if (r < mt) Ps[r * PS + c] = v[q];
                    }
                }
            }
            for (int c = tid; c < NB; c += BC_THREADS) Ps[mt * PS + c] = (c < nb) ? __ldcg(a.g + k0 + c) : 0.0;
            __syncthreads();
            // 4x4 register tiles over the lower triangle of the (mt x mt) window, round-robin over the cluster
            const int nt = (mt + 3) >> 2;                      // tiles per side
            const int ntiles = nt * (nt + 1) / 2;
            for (int t = rank * BC_THREADS + tid; t < ntiles && !(a.debug & 1); t += C * BC_THREADS) {
                // unrank t -> (ta >= tb)
                int ta = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while ((ta + 1) * (ta + 2) / 2 <= t) ++ta;
                while (ta * (ta + 1) / 2 > t) --ta;
                const int tb = t - ta * (ta + 1) / 2;
                const int a0 = ta * 4, b0 = tb * 4;
                double acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
#pragma unroll 8
                for (int c = 0; c < NB; ++c) {
                    double la[4], lb[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        la[x] = (a0 + x < mt) ? Ps[(a0 + x) * PS + c] : 0.0;
                        lb[x] = (b0 + x < mt) ? Ps[(b0 + x) * PS + c] : 0.0;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[x][y] = fma(la[x], lb[y], acc[x][y]);
                }
                double old[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int ra = a0 + x, rb = b0 + y;
                        old[x][y] = (ra < mt && rb <= ra) ? __ldcg(ab_at(a, k0 + nb + ra, k0 + nb + rb)) : 0.0;
                    }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int ra = a0 + x, rb = b0 + y;      // i - j = ra - rb <= mt - 1 <= bw - 1
                        if (ra < mt && rb <= ra) __stcg(ab_at(a, k0 + nb + ra, k0 + nb + rb), old[x][y] - acc[x][y]);
                    }
            }
            // rhs row: g[i] -= sum_c y_c * L[i][c]   (forward substitution riding along)
            if (rank == C - 1) {
                for (int r = tid; r < mt; r += BC_THREADS) {
                    double s = 0.0;
#pragma unroll 8
                    for (int c = 0; c < NB; ++c) s = fma(Ps[mt * PS + c], Ps[r * PS + c], s);
                    a.g[k0 + nb + r] = __ldcg(a.g + k0 + nb + r) - s;
                }
            }
        }
        cluster.sync();
    }

    // ================= back substitution  L^T x = y  (CTA 0) ============================================
    if (rank == 0 && !(a.debug & 2)) {
        double* part = Ps;                              // (warps x NB) partial sums
        double* Lkk = Ps + (BC_THREADS / 32) * PS;      // 32 x PS diagonal block
        double* xs = Lkk + NB * PS;                     // staged x of the rows below (<= bw)
        const int last = ((n - 1) / NB) * NB;
        for (int k0 = last; k0 >= 0; k0 -= NB) {
            const int nb = min(NB, n - k0);
            const int m = min(n - k0, NB + bw);
            const int mt = m - nb;
            for (int r = tid; r < mt; r += BC_THREADS) xs[r] = a.g[k0 + nb + r];      // written by this CTA
            for (int e = tid; e < NB * NB; e += BC_THREADS) {
                const int r = e / NB, c = e % NB;
                Lkk[r * PS + c] = (r < nb && c <= r && r - c <= bw) ? __ldcg(ab_at(a, k0 + r, k0 + c)) : 0.0;
            }
            __syncthreads();
            {   // s_c = sum_r L[k0+nb+r][k0+c] * x[k0+nb+r], 16 row-interleaved partials per column
                const int c = tid & 31, p = tid >> 5;
                double s = 0.0;
                if (c < nb)
                    for (int rb = p; rb < mt; rb += 8 * (BC_THREADS / 32)) {
                        double l[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int r = rb + q * (BC_THREADS / 32);
                            const int i = k0 + nb + r, j = k0 + c;
                            l[q] = (r < mt && i - j <= bw) ? __ldcg(ab_at(a, i, j)) : 0.0;
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int r = rb + q * (BC_THREADS / 32);
                            if (r < mt) s = fma(l[q], xs[r], s);
                        }
                    }
                part[p * PS + c] = s;
            }
            __syncthreads();
            if (warp == 0) {
                double s = (lane < nb) ? __ldcg(a.g + k0 + lane) : 0.0;
#pragma unroll
                for (int p = 0; p < BC_THREADS / 32; ++p) s -= part[p * PS + lane];
                const double inv = (lane < nb) ? a.dinv[k0 + lane] : 0.0;
                double x = 0.0;
#pragma unroll
                for (int c = NB - 1; c >= 0; --c) {
                    double xc = 0.0;
                    if (lane == c) { x = s * inv; xc = x; }
                    xc = __shfl_sync(0xffffffffu, xc, c);
                    if (lane < c) s = fma(-Lkk[c * PS + lane], xc, s);
                }
                if (lane < nb) a.g[k0 + lane] = x;
            }
            __syncthreads();
        }
    }
}

size_t band_smem_bytes(int bw) {
    const size_t panel = (size_t)(NB + bw + 1) * PS + NB;               // factorisation layout
    const size_t back = (size_t)(BC_THREADS / 32 + NB) * PS + bw + NB;  // back-substitution layout
    return (panel > back ? panel : back) * sizeof(double);
}

}  // namespace

static int g_band_debug = 0;

extern "C" {

int sb_band_debug(int flags) { g_band_debug = flags; return SB_OK; }

int sb_band_max_bw(void) { return (int)((227 * 1024 / sizeof(double) - NB) / PS) - NB - 1; }   // ~845

// cluster_size: 8 (portable) or 16
int sb_band_solve(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                  int cluster_size, void* stream) {
    if (!AB || !g || !dinv || !info || n <= 0 || bw < 0 || ldab < bw + 1) return SB_ERR_ARG;
    if (bw > sb_band_max_bw()) return SB_ERR_ARG;
    if (cluster_size != 1 && cluster_size != 2 && cluster_size != 4 && cluster_size != 8 && cluster_size != 16)
        return SB_ERR_ARG;
    BandArgs a;
    a.AB = AB; a.ldab = ldab; a.n = n; a.bw = bw; a.g = g; a.u = u; a.dinv = dinv; a.info = info; a.debug = g_band_debug;
    const size_t smem = band_smem_bytes(bw);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(band_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return SB_ERR_CUDA;
        cudaFuncSetAttribute(band_chol_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        configured = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(BC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, band_chol_kernel, a) != cudaSuccess) return SB_ERR_CUDA;
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
