// Pipelined banded Cholesky solve (v2) of (A + u I) x = g -- the LM step solve, replacing
// torch.linalg.cholesky + cholesky_solve   /root/reference/super/LM.py:38-51,97-100.
//
// Lesson of v1 (band_chol.cu, profiles/r1b): with one panel per cluster barrier, every panel pays
// 2 barriers + 4 L2 round trips + a 32-column pivot chain back to back (43 us/panel measured).  The
// factorisation is a dependency chain of n pivots; everything else is bulk work with slack.  v2 puts
// ONLY the chain on one CTA and lets the rest run behind it, synchronised by release/acquire flags
// in global memory (all CTAs belong to one thread-block cluster, so they are co-resident):
//
//   CTA 0  "P"  : per 32-column panel k: wait until the updates of panels <= k-2 have landed, load
//                 A(k,k-1), A(k,k); private trsm + syrk with L(k-1,k-1) kept in shared memory; Cholesky of
//                 the 32x32 diagonal block in one warp (registers, pivot chain = rsqrt -> shuffle -> own-row
//                 FMA; the other 31 columns' factors travel through a shared-memory column buffer);
//                 publish L(k,k) -> flag diag_done[k].
//   CTA 1  "R"  : forward substitution riding behind: y_k = L(k,k)^-1 g_k, g_rows -= L(rows,k) y_k.
//   CTAs 2.. "U": block rows are owned cyclically.  For panel p: after diag_done[p] solve the owned rows
//                 L(i,p) = A(i,p) L(p,p)^-T (count rows_done[p]); after all rows of the panel exist, apply
//                 the rank-32 update to the owned tiles (count upd_done[p]).  Tile (p+1,p+1) belongs to P;
//                 L(p+1,p) is stored out of place (Lsub) because P still needs A(p+1,p).
//   then       : back substitution on CTA 0.
#include <cooperative_groups.h>

#include "common.cuh"
#include "super_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int NB = 32;
constexpr int TS = NB + 1;          // shared-memory tile row stride
constexpr int TILE = NB * TS;       // doubles per shared tile
constexpr int THREADS = 256;
constexpr int MAX_WB = 28;          // block rows below a panel inside the band (bw <= 845)

struct Args {
    double* AB; int ldab, n, bw;
    double* g; const double* u; double* dinv; int* info;
    double* LB;        // (n, ldab) the factor L in band layout, OUT OF PLACE: AB keeps the updated-but-unfactored
                       // tiles so that any CTA can redo a 32x32 triangular solve instead of waiting for its owner
    int* flags;        // 3*NP ints: diag_done | rows_done | upd_done   (zeroed by the kernel)
    long long* prof;   // 16 cycle counters of the P role (debug & 4)
    int debug;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// whole CTA waits until *flag >= target
__device__ __forceinline__ void cta_wait(const int* flag, int target) {
    if (threadIdx.x == 0)
        while (ld_acquire(flag) < target) { __nanosleep(20); }
    __syncthreads();
}
// whole CTA has finished its global writes -> bump the flag
__device__ __forceinline__ void cta_signal(int* flag) {
    __syncthreads();
    // bar.sync orders the CTA's writes before thread 0; the release is cumulative (CUTLASS semaphore idiom).
    // An extra __threadfence() (fence.sc.gpu) here cost ~13k cycles per panel while other CTAs were polling.
    if (threadIdx.x == 0) red_release(flag, 1);
}

__device__ __forceinline__ bool in_band(const Args& a, int i, int j) { return i < a.n && j <= i && i - j <= a.bw; }
__device__ __forceinline__ double* ab_at(const Args& a, int i, int j) {
    return a.AB + (size_t)i * a.ldab + (j - i + a.bw);
}
__device__ __forceinline__ double* lb_at(const Args& a, int i, int j) {
    return a.LB + (size_t)i * a.ldab + (j - i + a.bw);
}

// Tile (I,J) of the band matrix <-> shared memory (row-major, stride TS).  256 threads: 4 elements each,
// loads issued back to back.  Entries outside the band / matrix read as 0.
__device__ __forceinline__ void load_tile(const Args& a, int I, int J, double* __restrict__ T) {
    const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 8 * q, i = NB * I + r, j = NB * J + c;
        v[q] = in_band(a, i, j) ? __ldcg(ab_at(a, i, j)) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) T[(r0 + 8 * q) * TS + c] = v[q];
}
__device__ __forceinline__ void store_tile(const Args& a, int I, int J, const double* __restrict__ T) {
    const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 8 * q, i = NB * I + r, j = NB * J + c;
        if (in_band(a, i, j)) __stcg(ab_at(a, i, j), T[r * TS + c]);
    }
}
// tile (I,P) of the factor L (out-of-place band LB)
__device__ __forceinline__ void load_L(const Args& a, int I, int P, double* __restrict__ T) {
    const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 8 * q, i = NB * I + r, j = NB * P + c;
        v[q] = in_band(a, i, j) ? __ldcg(lb_at(a, i, j)) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) T[(r0 + 8 * q) * TS + c] = v[q];
}
__device__ __forceinline__ void store_L(const Args& a, int I, int P, const double* __restrict__ T) {
    const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 8 * q, i = NB * I + r, j = NB * P + c;
        if (in_band(a, i, j)) __stcg(lb_at(a, i, j), T[r * TS + c]);
    }
}
// same, executed by a subset of the CTA: thread t of nt
__device__ __forceinline__ void load_tile_part(const Args& a, int I, int J, double* __restrict__ T, int t, int nt) {
    for (int base = 0; base < NB * NB; base += 5 * nt) {
        double v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = base + q * nt + t, i = NB * I + (e >> 5), j = NB * J + (e & 31);
            v[q] = (e < NB * NB && in_band(a, i, j)) ? __ldcg(ab_at(a, i, j)) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = base + q * nt + t;
            if (e < NB * NB) T[(e >> 5) * TS + (e & 31)] = v[q];
        }
    }
}
__device__ __forceinline__ void store_L_part(const Args& a, int I, int J, const double* __restrict__ T, int t, int nt) {
    for (int e = t; e < NB * NB; e += nt) {
        const int i = NB * I + (e >> 5), j = NB * J + (e & 31);
        if (in_band(a, i, j)) __stcg(lb_at(a, i, j), T[(e >> 5) * TS + (e & 31)]);
    }
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a_, double b_) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a_), "d"(b_));
}
__device__ __forceinline__ void bar_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// reciprocal square root for the pivot chain: float seed + one Newton step in double (rel. error ~1e-14,
// i.e. the factor of a matrix perturbed by 1e-14) -- about half the dependent latency of rsqrt(double)
__device__ __forceinline__ double rsqrt_chain(double x) {
    const double r = (double)rsqrtf((float)x);
    const double e = fma(-x * r, r, 1.0);
    return fma(0.5 * r, e, r);
}

// ---- compile-time unrolled row solve  x = row * L^{-T}  (row in registers, L and 1/diag in shared) ------
template <int C, int T>
struct TrsmInner {
    static __device__ __forceinline__ void run(double (&row)[NB], double x, const double* __restrict__ L) {
        row[T] = fma(-x, L[T * TS + C], row[T]);
        TrsmInner<C, T + 1>::run(row, x, L);
    }
};
template <int C>
struct TrsmInner<C, NB> {
    static __device__ __forceinline__ void run(double (&)[NB], double, const double* __restrict__) {}
};
template <int C>
struct TrsmStep {
    static __device__ __forceinline__ void run(double (&row)[NB], const double* __restrict__ L,
                                               const double* __restrict__ dinv) {
        const double x = row[C] * dinv[C];
        row[C] = x;
        TrsmInner<C, C + 1>::run(row, x, L);
        TrsmStep<C + 1>::run(row, L, dinv);
    }
};
template <>
struct TrsmStep<NB> {
    static __device__ __forceinline__ void run(double (&)[NB], const double* __restrict__, const double* __restrict__) {}
};
// rows of tile X (shared) solved in place by the calling warp, one row per lane.  One shared copy of the
// 1.1k-instruction unrolled body for all roles (noinline keeps the kernel's code footprint down).
__device__ __noinline__ void trsm_tile_warp(double* __restrict__ X, const double* __restrict__ L,
                                            const double* __restrict__ dinv) {
    const int lane = threadIdx.x & 31;
    double row[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) row[c] = X[lane * TS + c];
    TrsmStep<0>::run(row, L, dinv);
#pragma unroll
    for (int c = 0; c < NB; ++c) X[lane * TS + c] = row[c];
}
__device__ __forceinline__ void trsm_tile_warp0(double* __restrict__ X, const double* __restrict__ L,
                                                const double* __restrict__ dinv) {
    if (threadIdx.x < NB) trsm_tile_warp(X, L, dinv);
}

// ---- 32x32 Cholesky in one warp: row per lane in registers, column factors broadcast through shared ------
// Software-pipelined: the reciprocal square root of pivot C+1 is started right after the own-row FMA that
// completes it and overlaps the 31-C bulk updates of column C; all updates are branch-free (entries above
// the diagonal are garbage that is never read).
template <int C, int T>
struct PotrfInner {
    static __device__ __forceinline__ void run(double (&a)[NB], int lane, const double* __restrict__ col) {
        double f = col[T];
        if (T == C + 1) f = (lane == C + 1) ? 0.0 : f;       // that entry was updated on the critical chain
        a[T] = fma(-a[C], f, a[T]);
        PotrfInner<C, T + 1>::run(a, lane, col);
    }
};
template <int C>
struct PotrfInner<C, NB> {
    static __device__ __forceinline__ void run(double (&)[NB], int, const double* __restrict__) {}
};
template <int C>
struct PotrfStep {
    static __device__ __forceinline__ void run(double (&a)[NB], int lane, double inv, double& my_inv, bool& bad,
                                               double* __restrict__ col /* 2 x NB ping-pong */) {
        if (lane == C) {
            bad = bad || !(a[C] > 0.0);
            my_inv = inv;
        }
        a[C] *= inv;                                                  // l_rC (lane C: piv * rsqrt(piv) = l_CC)
        double inv_next = 0.0;
        if constexpr (C + 1 < NB) {
            const double d1 = fma(-a[C], a[C], a[C + 1]);             // completes pivot C+1 in lane C+1
            a[C + 1] = (lane == C + 1) ? d1 : a[C + 1];
        }
        double* cb = col + (C & 1) * NB;
        cb[lane] = a[C];
        __syncwarp();
        if constexpr (C + 1 < NB) inv_next = rsqrt_chain(a[C + 1]);   // overlaps the bulk below
        PotrfInner<C, C + 1>::run(a, lane, cb);
        if constexpr (C + 1 < NB) inv_next = __shfl_sync(0xffffffffu, inv_next, C + 1);
        PotrfStep<C + 1>::run(a, lane, inv_next, my_inv, bad, col);
    }
};
template <>
struct PotrfStep<NB> {
    static __device__ __forceinline__ void run(double (&)[NB], int, double, double&, bool&, double* __restrict__) {}
};
__device__ __forceinline__ double warp_potrf(double (&row)[NB], int lane, int* info, double* __restrict__ colbuf) {
    double my_inv = 0.0;
    bool bad = false;
    const double inv0 = __shfl_sync(0xffffffffu, rsqrt_chain(row[0]), 0);
    PotrfStep<0>::run(row, lane, inv0, my_inv, bad, colbuf);
    if (bad) *info = 1;
    return my_inv;
}

__global__ void __launch_bounds__(THREADS, 1) band_chol2_kernel(Args a) {
    extern __shared__ double smem[];
    // One cooperative grid (co-residency guaranteed by the cooperative launch): the roles only meet at two
    // grid-wide barriers, everything else is flag traffic, so the grid is not limited to a 16-CTA cluster.
    cg::grid_group cluster = cg::this_grid();
    const int rank = (int)blockIdx.x, C = (int)gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, bw = a.bw;
    const int NP = (n + NB - 1) / NB;
    const int WB = min((bw + NB - 1) / NB + 0, NP);          // block rows below a panel that the band reaches
    int* diag_done = a.flags;
    int* rows_done = a.flags + NP;
    int* upd_done = a.flags + 2 * NP;
    const int NU = C - 2;                                     // update CTAs (ranks 2..C-1)
    const double u = a.u ? *a.u : 0.0;

    for (int i = rank * THREADS + tid; i < 3 * NP; i += C * THREADS) a.flags[i] = 0;
    cluster.sync();

    if (rank == 0) {
        // ======================= P: the pivot chain ======================================================
        // Per panel only  trsm -> syrk -> potrf  is serial.  While warp 0 factors block k, warps 1..7 wait for
        // the updates of panels <= k-1 and stage A(k+1,k+1), A(k+1,k); while warp 0 solves the next
        // sub-diagonal tile they write L(k,k) out and raise diag_done[k].
        // buffers rotate by index arithmetic (no pointer arrays: those went to local memory and their reloads
        // cost 13k cycles per panel once the polling CTAs were invalidating L1):
        //   tiles 0..2: L(k-1,k-1) | A(k,k)->L(k,k) | next diagonal;  tiles 3..4: A(k,k-1) copy | next
        double* dvbase = smem + 5 * TILE;                           // 2 x NB reciprocal pivots
        double* colbuf = smem + 5 * TILE + 2 * NB;                  // 2 x NB
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long t0 = clock64(), t1;
#define PROF(slot) do { if (a.debug & 4) { t1 = clock64(); tacc[slot] += t1 - t0; t0 = t1; } } while (0)
        load_tile(a, 0, 0, smem + TILE);
        if (tid == 32) { a.prof[8] = 0; a.prof[9] = 0; a.prof[10] = 0; a.prof[11] = 0; a.prof[12] = 0; }
        __syncthreads();
        for (int k = 0; k < NP; ++k) {
            double* Lprev = smem + (k % 3) * TILE;
            double* D = smem + ((k + 1) % 3) * TILE;
            double* Dnext = smem + ((k + 2) % 3) * TILE;
            double* X = smem + (3 + (k & 1)) * TILE;
            double* Xnext = smem + (3 + ((k + 1) & 1)) * TILE;
            double* dprev = dvbase + ((k + 1) & 1) * NB;
            double* dcurp = dvbase + (k & 1) * NB;
            if (tid < NB) {                                   // damping + identity padding past n
                const int i = NB * k + tid;
                D[tid * TS + tid] = (i < n) ? D[tid * TS + tid] + u : 1.0;
            }
            PROF(0);
            if (k >= 1) {
                if (warp == 0) {
                    long long q0 = clock64();
                    trsm_tile_warp0(X, Lprev, dprev);
                    if ((a.debug & 4) && tid == 0) a.prof[11] += clock64() - q0;
                    PROF(1);
                } else {                                      // publish L(k-1,k-1)
                    long long q0 = clock64();
                    store_L_part(a, k - 1, k - 1, Lprev, tid - 32, THREADS - 32);
                    long long q1 = clock64();
                    bar_named(1, THREADS - 32);
                    long long q2 = clock64();
                    if (tid == 32) red_release(diag_done + (k - 1), 1);
                    long long q3 = clock64();
                    if ((a.debug & 4) && tid == 32) { a.prof[8] += q1 - q0; a.prof[9] += q2 - q1; a.prof[10] += q3 - q2; }
                }
                __syncthreads();
                PROF(2);
                {   // D -= X X^T: warp w -> rows 4w..4w+3, lane -> column (conflict-free, X row broadcast)
                    const int r0 = warp * 4, c = lane;
                    double s4[4] = {0, 0, 0, 0};
#pragma unroll 8
                    for (int t = 0; t < NB; ++t) {
                        const double xc = X[c * TS + t];
#pragma unroll
                        for (int q = 0; q < 4; ++q) s4[q] = fma(X[(r0 + q) * TS + t], xc, s4[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c <= r0 + q) D[(r0 + q) * TS + c] -= s4[q];
                }
            }
            __syncthreads();
            PROF(3);
            if (warp == 0) {
                double row[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) row[c] = D[lane * TS + c];
                const double inv = warp_potrf(row, lane, a.info, colbuf);
#pragma unroll
                for (int c = 0; c < NB; ++c) D[lane * TS + c] = (c <= lane) ? row[c] : 0.0;
                dcurp[lane] = inv;
                if (NB * k + lane < n) a.dinv[NB * k + lane] = inv;
            } else if (k + 1 < NP) {                          // stage the next panel's tiles behind the potrf
                if (tid == 32 && k >= 1)
                    while (ld_acquire(upd_done + (k - 1)) < NU) { __nanosleep(20); }
                bar_named(1, THREADS - 32);
                load_tile_part(a, k + 1, k + 1, Dnext, tid - 32, THREADS - 32);
                load_tile_part(a, k + 1, k, Xnext, tid - 32, THREADS - 32);
            }
            __syncthreads();
            PROF(4);
        }
        store_L(a, NP - 1, NP - 1, smem + (NP % 3) * TILE);
        cta_signal(diag_done + (NP - 1));
        if ((a.debug & 4) && tid == 0)
            for (int q = 0; q < 8; ++q) a.prof[q] = tacc[q];
#undef PROF
    } else if (rank == 1) {
        // ======================= R: forward substitution  L y = g ========================================
        double* Lpp = smem;
        double* ys = smem + TILE;              // NB
        double* Lrows = smem + TILE + NB;      // up to MAX_WB tiles L(I,p), staged with ONE L2 round trip
        for (int p = 0; p < NP; ++p) {
            cta_wait(diag_done + p, 1);
            load_L(a, p, p, Lpp);
            __syncthreads();
            if (warp == 0) {                    // y_p = L_pp^{-1} g_p : column-oriented forward solve
                const int i = NB * p + lane;
                double s = (i < n) ? __ldcg(a.g + i) : 0.0;
                const double inv = (i < n) ? __ldcg(a.dinv + i) : 0.0;
                double y = 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    double yc = 0.0;
                    if (lane == c) { y = s * inv; yc = y; }
                    yc = __shfl_sync(0xffffffffu, yc, c);
                    if (lane > c) s = fma(-Lpp[lane * TS + c], yc, s);
                }
                ys[lane] = y;
                if (i < n) a.g[i] = y;
            }
            const int last = min(NP - 1, p + WB);
            const int nrows = last - p;
            if (nrows > 0) {
                cta_wait(rows_done + p, NU);
                // all tiles below the panel at once: thread -> (tile q, 4 elements), loads back to back
                const int c = tid & 31, r0 = tid >> 5;
                for (int qb = 0; qb < nrows; qb += 4) {
                    double v[4][4];
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int q = qb + qq, I = p + 1 + q, r = r0 + 8 * e;
                            const int i = NB * I + r, j = NB * p + c;
                            double x = 0.0;
                            if (q < nrows && i < n && i - j <= bw) x = __ldcg(lb_at(a, i, j));
                            v[qq][e] = x;
                        }
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (qb + qq < nrows) Lrows[(size_t)(qb + qq) * TILE + (r0 + 8 * e) * TS + c] = v[qq][e];
                }
                __syncthreads();
                for (int rr = tid; rr < nrows * NB; rr += THREADS) {       // g_i -= L(i, panel p) . y_p
                    const int i = NB * (p + 1) + rr;
                    if (i < n) {
                        const double* Lr = Lrows + (size_t)(rr >> 5) * TILE + (rr & 31) * TS;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int cc = 0; cc < NB; cc += 2) {
                            s0 = fma(Lr[cc], ys[cc], s0);
                            s1 = fma(Lr[cc + 1], ys[cc + 1], s1);
                        }
                        a.g[i] = __ldcg(a.g + i) - (s0 + s1);
                    }
                }
            }
            __syncthreads();
        }
    } else {
        // ======================= U: trailing update + panel rows, work dealt round-robin ==================
        // Critical cycle of the whole solve: P publishes L(p,p) -> U applies panel p to the trailing tiles ->
        // P can stage the tiles of panel p+2.  Every dependent L2 round trip in between delays the pivot chain,
        // so a tile's owner does NOT wait for the owners of block rows I and J: it loads the unfactored
        // A(I,p), A(J,p) (AB is never overwritten by L) and redoes the two 32x32 triangular solves itself, one
        // warp per solve.  The panel rows that R and the back substitution need are solved and written to LB
        // afterwards, off the cycle.
        const int ui = rank - 2;
        constexpr int TB = 4;                  // trailing tiles per batch = 8 warps / 2
        double* Lpp = smem;
        double* dinv = smem + TILE;            // NB
        double* Ls = smem + TILE + NB;         // 2*TB tiles: slot 2b = L(I_b,p), slot 2b+1 = L(J_b,p)
        const int c4 = tid & 31, r4 = tid >> 5;
        long long uacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long ut0 = clock64(), ut1;
#define UPROF(slot) do { if (a.debug & 4) { __syncthreads(); ut1 = clock64(); uacc[slot] += ut1 - ut0; ut0 = ut1; } } while (0)
        for (int p = 0; p < NP; ++p) {
            const int last = min(NP - 1, p + WB);
            const int nrows = last - p;                                  // block rows below the panel
            const bool skip = (a.debug & 1) != 0;
            const int ntiles = nrows * (nrows + 1) / 2;
            const int q0 = ((ui - p) % NU + NU) % NU;                     // my first row / tile index
            const bool have_work = !skip && (q0 < nrows || q0 < ntiles);
            UPROF(7);
            if (have_work) {
                if (p >= 1) cta_wait(upd_done + (p - 1), NU);           // all tiles carry panels <= p-1
                UPROF(0);
                cta_wait(diag_done + p, 1);
                UPROF(1);
            }
            // ---- trailing tiles (I,J), p+1 <= J <= I <= last, enumerated row by row; tile t -> CTA (t+p) % NU
            int t = q0;
            bool first = true;
            while (t < ntiles && !skip) {
                int TI[TB], TJ[TB], nb_ = 0;      // indexed only by unrolled constants: stay in registers
#pragma unroll
                for (int b = 0; b < TB; ++b) {
                    TI[b] = -1; TJ[b] = -1;
                    while (t < ntiles) {
                        int ri = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
                        while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
                        while (ri * (ri + 1) / 2 > t) --ri;
                        const int I = p + 1 + ri, J = p + 1 + (t - ri * (ri + 1) / 2);
                        t += NU;
                        if (I == p + 1) continue;                         // tile (p+1,p+1) belongs to P
                        if (NB * (I - J) - (NB - 1) > bw) continue;       // entirely outside the band
                        TI[b] = I; TJ[b] = J; nb_ = b + 1;
                        break;
                    }
                }
                if (nb_ == 0) break;
                if (!first) __syncthreads();
                const int wb = warp >> 1, mh = warp & 1;              // my tile in the batch, my row half
                int myI = -1, myJ = -1;
#pragma unroll
                for (int b = 0; b < TB; ++b)
                    if (b == wb) { myI = TI[b]; myJ = TJ[b]; }
                // ALL global loads of the batch first (one L2 round trip): L(p,p), the unfactored A(I,p), A(J,p)
                // of every tile and the old values of my 8x8 output blocks
                double vp[4], vi[TB][4], vj[TB][4], oldv[2][4][2];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int rr = r4 + 8 * e, j = NB * p + c4;
                    vp[e] = (first && in_band(a, NB * p + rr, j)) ? __ldcg(lb_at(a, NB * p + rr, j)) : 0.0;
                }
#pragma unroll
                for (int b = 0; b < TB; ++b)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int I = TI[b], J = TJ[b], rr = r4 + 8 * e, j = NB * p + c4;
                        vi[b][e] = (I >= 0 && in_band(a, NB * I + rr, j)) ? __ldcg(ab_at(a, NB * I + rr, j)) : 0.0;
                        vj[b][e] = (I >= 0 && J != I && in_band(a, NB * J + rr, j)) ? __ldcg(ab_at(a, NB * J + rr, j)) : 0.0;
                    }
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int nb2 = 0; nb2 < 4; ++nb2)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int i = NB * myI + 8 * (2 * mh + mb) + (lane >> 2);
                            const int j = NB * myJ + 8 * nb2 + 2 * (lane & 3) + e;
                            oldv[mb][nb2][e] = (myI >= 0 && in_band(a, i, j)) ? __ldcg(ab_at(a, i, j)) : 0.0;
                        }
                if (first && tid < NB) dinv[tid] = (NB * p + tid < n) ? __ldcg(a.dinv + NB * p + tid) : 1.0;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (first) Lpp[(r4 + 8 * e) * TS + c4] = vp[e];
#pragma unroll
                for (int b = 0; b < TB; ++b)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        Ls[(size_t)(2 * b) * TILE + (r4 + 8 * e) * TS + c4] = vi[b][e];
                        Ls[(size_t)(2 * b + 1) * TILE + (r4 + 8 * e) * TS + c4] = vj[b][e];
                    }
                first = false;
                __syncthreads();
                {   // warp w solves slot w (skipped when the slot is empty or duplicates its partner)
                    const int b = warp >> 1;
                    int I_ = -1, J_ = -1;
#pragma unroll
                    for (int bb = 0; bb < TB; ++bb)
                        if (bb == b) { I_ = TI[bb]; J_ = TJ[bb]; }
                    if (I_ >= 0 && ((warp & 1) == 0 || J_ != I_)) trsm_tile_warp(Ls + (size_t)warp * TILE, Lpp, dinv);
                }
                __syncthreads();
                if (myI >= 0) {
                    const double* La = Ls + (size_t)(2 * wb) * TILE;
                    const double* Lb = (myJ == myI) ? La : Ls + (size_t)(2 * wb + 1) * TILE;
                    double acc[2][4][2];
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                        for (int nb2 = 0; nb2 < 4; ++nb2) acc[mb][nb2][0] = acc[mb][nb2][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < NB / 4; ++ks) {
                        double fa[2], fb[4];
#pragma unroll
                        for (int mb = 0; mb < 2; ++mb)
                            fa[mb] = La[(8 * (2 * mh + mb) + (lane >> 2)) * TS + 4 * ks + (lane & 3)];
#pragma unroll
                        for (int nb2 = 0; nb2 < 4; ++nb2) fb[nb2] = Lb[(8 * nb2 + (lane >> 2)) * TS + 4 * ks + (lane & 3)];
#pragma unroll
                        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                            for (int nb2 = 0; nb2 < 4; ++nb2) dmma884(acc[mb][nb2][0], acc[mb][nb2][1], fa[mb], fb[nb2]);
                    }
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                        for (int nb2 = 0; nb2 < 4; ++nb2)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int i = NB * myI + 8 * (2 * mh + mb) + (lane >> 2);
                                const int j = NB * myJ + 8 * nb2 + 2 * (lane & 3) + e;
                                if (in_band(a, i, j)) __stcg(ab_at(a, i, j), oldv[mb][nb2][e] - acc[mb][nb2][e]);
                            }
                }
            }
            UPROF(5);
            cta_signal(upd_done + p);
            UPROF(6);
            // ---- panel rows for R / back substitution: block row p+1+q -> CTA (q + p) % NU  (off the cycle)
            for (int q = q0; q < nrows && !skip; q += NU) {
                const int I = p + 1 + q;
                __syncthreads();
                if (first) {                                              // no tile batch ran: L(p,p) not staged yet
                    load_L(a, p, p, Lpp);
                    if (tid < NB) dinv[tid] = (NB * p + tid < n) ? __ldcg(a.dinv + NB * p + tid) : 1.0;
                    first = false;
                }
                load_tile(a, I, p, Ls);
                __syncthreads();
                trsm_tile_warp0(Ls, Lpp, dinv);
                __syncthreads();
                store_L(a, I, p, Ls);
            }
            UPROF(2);
            cta_signal(rows_done + p);
            UPROF(3);
        }
        if ((a.debug & 4) && tid == 0 && ui == 0)
            for (int q = 0; q < 8; ++q) a.prof[16 + q] = uacc[q];
#undef UPROF
    }
    cluster.sync();

    // ======================= back substitution  L^T x = y  (CTA 0) =========================================
    // Warp 0 does the arithmetic of panel k (column-block dot products out of shared memory + the 32-step
    // triangular solve); warps 1..7 meanwhile stream panel k-1's column block of L from L2 into the other
    // shared buffer.  x stays in shared memory.
    if (rank == 0 && !(a.debug & 2)) {
        const int MR = (WB + 1) * NB;                    // rows per column block (diagonal tile + WB tiles below)
        double* xs = smem;                               // (NP*NB) solution, indexed by matrix row
        double* buf0 = xs + (size_t)NP * NB;             // 2 x (MR x TS) column blocks
        double* buf1 = buf0 + (size_t)MR * TS;
        auto prefetch = [&](int k, double* __restrict__ dst, int t0, int nthreads) {
            // rows k0 .. k0+MR-1 of columns k0..k0+31: tile (k,k), Lsub[k] for the first tile below, band after
            const int k0 = NB * k;
            for (int e0 = t0; e0 < MR * NB; e0 += 16 * nthreads) {
                double v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int e = e0 + q * nthreads, r = e >> 5, c = e & 31;
                    const int i = k0 + r, j = k0 + c;
                    double x = 0.0;
                    if (e < MR * NB && i < n && j < n && j <= i && i - j <= bw) x = __ldcg(lb_at(a, i, j));
                    v[q] = x;
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int e = e0 + q * nthreads;
                    if (e < MR * NB) dst[(e >> 5) * TS + (e & 31)] = v[q];
                }
            }
        };
        for (int i = tid; i < NP * NB; i += THREADS) xs[i] = (i < n) ? __ldcg(a.g + i) : 0.0;   // y
        prefetch(NP - 1, buf0, tid, THREADS);
        __syncthreads();
        for (int k = NP - 1; k >= 0; --k) {
            double* cur = ((NP - 1 - k) & 1) ? buf1 : buf0;
            double* nxt = ((NP - 1 - k) & 1) ? buf0 : buf1;
            const int k0 = NB * k;
            if (warp == 0) {
                const int rows = min(MR, NB * NP - k0);                // rows of the block that exist (padded to tiles)
                double s = xs[k0 + lane], s1 = 0.0, s2 = 0.0, s3 = 0.0;      // 4 chains: FP64 latency ~20 cycles
                for (int r = NB; r + 3 < rows; r += 4) {                     // rows is a multiple of 32
                    s = fma(-cur[r * TS + lane], xs[k0 + r], s);
                    s1 = fma(-cur[(r + 1) * TS + lane], xs[k0 + r + 1], s1);
                    s2 = fma(-cur[(r + 2) * TS + lane], xs[k0 + r + 2], s2);
                    s3 = fma(-cur[(r + 3) * TS + lane], xs[k0 + r + 3], s3);
                }
                s = (s + s1) + (s2 + s3);
                const double inv = (k0 + lane < n) ? a.dinv[k0 + lane] : 0.0;
                double x = 0.0;
#pragma unroll
                for (int c = NB - 1; c >= 0; --c) {
                    double xc = 0.0;
                    if (lane == c) { x = s * inv; xc = x; }
                    xc = __shfl_sync(0xffffffffu, xc, c);
                    if (lane < c) s = fma(-cur[c * TS + lane], xc, s);
                }
                xs[k0 + lane] = x;
                if (k0 + lane < n) a.g[k0 + lane] = x;
            } else if (k >= 1) {
                prefetch(k - 1, nxt, tid - 32, THREADS - 32);
            }
            __syncthreads();
        }
    }
}

size_t smem_bytes(int bw, int n) {
    const int WB = (bw + NB - 1) / NB;
    size_t u_role = (size_t)TILE + NB + 8 * (size_t)TILE;   // L(p,p) + 1/diag + 2*TB tile slots
    size_t r_role = (size_t)TILE + NB + (size_t)(WB > 0 ? WB : 1) * TILE;
    if (r_role > u_role) u_role = r_role;
    size_t p_role = 5 * (size_t)TILE + 4 * NB;
    size_t back = (size_t)0;   // sized by the caller-visible formula below (needs n)
    const size_t NP = (size_t)(n + NB - 1) / NB;
    back = NP * NB + 2 * (size_t)(WB + 1) * NB * TS;
    size_t m = u_role > p_role ? u_role : p_role;
    if (back > m) m = back;
    return m * sizeof(double);
}

int g_debug2 = 0;

}  // namespace

extern "C" {

int sb_band2_debug(int flags) { g_debug2 = flags; return SB_OK; }

/* 1 when (n, bw) fits the pipelined kernel's shared-memory layout, else 0 (use sb_band_solve) */
int sb_band2_fits(int n, int bw) {
    return ((bw + NB - 1) / NB <= MAX_WB && smem_bytes(bw, n) <= 227 * 1024) ? 1 : 0;
}

static long long ws_bytes2(int n, int ldab) {
    const long long NP = (n + NB - 1) / NB;
    return (long long)n * ldab * (long long)sizeof(double) + 3 * NP * (long long)sizeof(int) + 1024;
}

long long sb_band2_workspace_bytes2(int n, int ldab) { return ws_bytes2(n, ldab); }


int sb_band_solve2(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                   void* workspace, long long ws_bytes, int cluster_size, void* stream) {
    if (!AB || !g || !dinv || !info || !workspace || n <= 0 || bw < 0 || ldab < bw + 1) return SB_ERR_ARG;
    if (cluster_size < 3 || cluster_size > 128) return SB_ERR_ARG;    // number of CTAs of the cooperative grid
    if ((bw + NB - 1) / NB > MAX_WB) return SB_ERR_ARG;
    if (ws_bytes < ws_bytes2(n, ldab)) return SB_ERR_WORKSPACE;
    const long long NP = (n + NB - 1) / NB;
    Args a;
    a.AB = AB; a.ldab = ldab; a.n = n; a.bw = bw; a.g = g; a.u = u; a.dinv = dinv; a.info = info;
    a.LB = (double*)workspace;
    a.flags = (int*)((char*)workspace + (size_t)n * ldab * sizeof(double));
    a.prof = (long long*)((char*)workspace + (size_t)n * ldab * sizeof(double) + 3 * NP * sizeof(int) + 64);
    a.prof = (long long*)(((size_t)a.prof + 7) & ~(size_t)7);
    a.debug = g_debug2;
    const size_t smem = smem_bytes(bw, n);
    if (smem > 227 * 1024) return SB_ERR_ARG;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(band_chol2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = smem;
    }
    void* kargs[] = {(void*)&a};
    if (cudaLaunchCooperativeKernel((const void*)band_chol2_kernel, dim3(cluster_size), dim3(THREADS), kargs, smem,
                                    (cudaStream_t)stream) != cudaSuccess)
        return SB_ERR_CUDA;
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
