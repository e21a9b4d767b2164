// Shared device helpers for the SuPer ED-tracking kernels (sm_100a).
//
// Numerical contract (see DESIGN.md "bit-exact correspondences"): the chain
//   warp T(p) -> projection (u,v) -> round / floor / ceil
// decides INTEGER outputs (pixel ids, bilinear corner ids) that must equal the reference's, so it
// is written with explicit IEEE intrinsics in the reference's operation order
// (/root/reference/super/utils.py:17-57, /root/reference/utils/utils.py:161-184); nvcc never
// contracts __dmul_rn/__dadd_rn into FMAs.  torch's CPU cross product evaluates
// a_p*b_q - a_r*b_s as fma(a_p, b_q, -(a_r*b_s)) (measured, oracle/super_oracle.py uses torch.cross),
// which cross_ref() reproduces.  Everything downstream (residuals, Jacobians) is tolerance-checked
// and free to use FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SB_OK 0
#define SB_ERR_ARG 1
#define SB_ERR_CUDA 2
#define SB_ERR_WORKSPACE 3

#define SB_KNN 4            // neighbours per surfel / per node (opt.num_neighbors, num_ED_neighbors)

#define SB_CHECK_LAUNCH()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return SB_ERR_CUDA;         \
    } while (0)

struct Cam {
    double fx, fy, cx, cy;
    int H, W;
};

struct V3 {
    double x, y, z;
};

__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ double dot3(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(const V3& a, const V3& b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// ---- reference-order arithmetic (no contraction) ---------------------------------------------
__device__ __forceinline__ double mulr(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double addr(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double subr(double a, double b) { return __dadd_rn(a, -b); }

// torch.cross on CPU: c_x = fma(a_y, b_z, -(a_z*b_y)) etc.
__device__ __forceinline__ V3 cross_ref(const V3& a, const V3& b) {
    return v3(__fma_rn(a.y, b.z, -mulr(a.z, b.y)),
              __fma_rn(a.z, b.x, -mulr(a.x, b.z)),
              __fma_rn(a.x, b.y, -mulr(a.y, b.x)));
}

// R(q) v + b in the reference's order: tv = (v + (2*qw)*cp) + 2*cross(qv,cp); tv += b
// (/root/reference/super/utils.py:53-57).  q is NOT normalised.  cp_out = qv x v.
__device__ __forceinline__ V3 quat_rot_ref(const V3& v, double qw, const V3& qv, V3& cp_out) {
    V3 cp = cross_ref(qv, v);
    V3 c2 = cross_ref(qv, cp);
    double tq = mulr(2.0, qw);
    V3 r;
    r.x = addr(addr(v.x, mulr(tq, cp.x)), mulr(2.0, c2.x));
    r.y = addr(addr(v.y, mulr(tq, cp.y)), mulr(2.0, c2.y));
    r.z = addr(addr(v.z, mulr(tq, cp.z)), mulr(2.0, c2.z));
    cp_out = cp;
    return r;
}

// Projection u = X*fx/(Z+1e-8) + cx, v = Y*fy/(Z+1e-8) + cy   (/root/reference/utils/utils.py:172-175)
__device__ __forceinline__ void project_ref(const V3& T, const Cam& cam, double& u, double& v) {
    double Zp = addr(T.z, 1e-8);
    u = addr(__ddiv_rn(mulr(T.x, cam.fx), Zp), cam.cx);
    v = addr(__ddiv_rn(mulr(T.y, cam.fy), Zp), cam.cy);
}

// torch.round(x).long(): round half to even.  Out-of-range / NaN handled by the caller's range test.
__device__ __forceinline__ long long round_ll(double x) { return __double2ll_rn(x); }

__device__ __forceinline__ int n_active(int n_cap, const int* n_dev) {
    return n_dev ? min(n_cap, *n_dev) : n_cap;
}

// Lower-triangular normal-equation matrix, dense (bw < 0: A[row*lda + col]) or band
// (A[row*lda + col - row + bw], entries with row - col > bw raise *overflow), with an optional node
// permutation (node id -> position in the solver's ordering), and the right-hand side g beside it.
//
// Accumulation mode.  shift < 0: f64 atomics (order of arrival decides the last bits: run-to-run differences ~1e-9 on
// J^T J -- kept for the dense cross-check path).  shift >= 0: FIXED POINT -- A and g are int64 arrays, every addend is
// rounded once to a multiple of 2^-shift (2^-gshift for g) and added with an integer atomic.  Integer addition is
// associative, so the sums do not depend on the order in which the warps arrive: the normal equations, hence the whole
// tracker, are bitwise reproducible.  Resolution 2^-40 = 9.1e-13 absolute on entries of magnitude 1e2..1e4 (below the
// f64 rounding of such an entry above 8192), range +-2^(62-shift); an addend beyond the range sets bit 1 of *overflow.
struct MatView {
    double* A;
    int lda, bw;
    const int* node_pos;
    int* overflow;
    double* g;
    double scale, gscale;    // fixed point: 2^shift, 2^gshift (multiplying by a power of two is exact); 0 = f64 atomics
    __host__ __device__ __forceinline__ void set_shift(int shift, int gshift) {
        scale = shift >= 0 ? (double)(1ull << shift) : 0.0;
        gscale = shift >= 0 ? (double)(1ull << gshift) : 0.0;
    }
    __device__ __forceinline__ int pos(int node) const { return node_pos ? node_pos[node] : node; }
    // MODE 0: decided at run time by the scale; 1: fixed point known at compile time (the frame loop's kernel: one atomic
    // per call site keeps its flush loop small enough for the instruction cache)
    template <int MODE = 0>
    __device__ __forceinline__ void put(double* p, double v, double sc) const {
        if (MODE == 0 && sc == 0.0) {
            asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
        } else {
            const double sv = v * sc;
            // red.global: fire-and-forget.  atomicAdd on the (generic) pointer compiles to a synchronous ATOM + a
            // shared-window test whose round trip the warp waits for -- 14 of them in a row made a flush cost ~10 k
            // cycles (ncu r2c-r2f: the Gram pass at 65-90 us for ~20 us of loads and products).
            if (fabs(sv) < 4.6e18)
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(__double2ll_rn(sv)) : "memory");
            else atomicOr(overflow, 2);
        }
    }
    template <int MODE = 0>
    __device__ __forceinline__ void add(int row, int col, double v) const {   // requires row >= col
        if (bw < 0 || row - col <= bw) {
            put<MODE>(A + (bw < 0 ? (size_t)row * lda + col : (size_t)row * lda + (col - row + bw)), v, scale);
        } else {
            atomicOr(overflow, 1);
        }
    }
    template <int MODE = 0>
    __device__ __forceinline__ void add_g(int i, double v) const { put<MODE>(g + i, v, gscale); }
};

// ---- block-wide deterministic sum (fixed tree), result valid in thread 0 -----------------------
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double* smem /* BLOCK/32 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < BLOCK / 32) ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}
