// k-nearest ED-node assignment, kNN blend weights and the quaternion+translation warp of surfels
// and nodes: find_knn (/root/reference/utils/utils.py:212-220 -> pytorch3d knn_points),
// Surfels.update_ed / update_sfed_knn (/root/reference/super/nodes.py:154-191) and Surfels.update
// (/root/reference/super/nodes.py:193-223).
//
// kNN definition (pytorch3d is a third-party dependency absent from the reference tree; the oracle
// DEFINES it): d2 = ((dx*dx + dy*dy) + dz*dz) in f64 with individually rounded products,
// ascending, ties -> lower index.  Brute force: the reference set (J <= a few thousand nodes) is
// staged through shared memory in tiles, one query per thread keeps its K best in registers.
#include "common.cuh"
#include "super_b200.h"

namespace {

constexpr int KNN_BLOCK = 128;
constexpr int KNN_TILE = 512;   // reference points per shared-memory tile (12 KB)
constexpr int KNN_MAXK = 8;

template <int K>
__global__ void __launch_bounds__(KNN_BLOCK)
knn_kernel(const double* __restrict__ q, int nq_cap, const int* nq_dev, const double* __restrict__ ref, int nref,
           int dim, double* __restrict__ out_d, int* __restrict__ out_i, const int* __restrict__ qseg,
           const int* __restrict__ rseg) {
    __shared__ double tile[KNN_TILE * 3];
    const int nq = n_active(nq_cap, nq_dev);
    const int i = blockIdx.x * KNN_BLOCK + threadIdx.x;
    double bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { bd[k] = INFINITY; bi[k] = -1; }
    double qx = 0, qy = 0, qz = 0;
    const int qc = (qseg && i < nq) ? qseg[i] : 0;      // --hard_seg: neighbours of the query's own class only
    if (i < nq) {
        qx = q[dim * (size_t)i];
        qy = q[dim * (size_t)i + 1];
        qz = dim > 2 ? q[dim * (size_t)i + 2] : 0.0;
    }
    for (int t0 = 0; t0 < nref; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, nref - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * dim; e += KNN_BLOCK) tile[e] = ref[(size_t)t0 * dim + e];
        __syncthreads();
        if (i < nq) {
            for (int j = 0; j < cnt; ++j) {
                if (rseg && rseg[t0 + j] != qc) continue;
                const double dx = subr(qx, tile[dim * j]), dy = subr(qy, tile[dim * j + 1]);
                double d2 = addr(mulr(dx, dx), mulr(dy, dy));
                if (dim > 2) {
                    const double dz = subr(qz, tile[dim * j + 2]);
                    d2 = addr(d2, mulr(dz, dz));
                }
                if (d2 < bd[K - 1]) {   // strict: an equal distance never displaces a lower index
                    bd[K - 1] = d2;
                    bi[K - 1] = t0 + j;
#pragma unroll
                    for (int k = K - 1; k > 0; --k) {
                        if (bd[k] < bd[k - 1]) {
                            const double td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                            const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (i < nq) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            out_d[(size_t)i * K + k] = (bi[k] >= 0) ? __dsqrt_rn(bd[k]) : 1e8;   // find_knn returns sqrt(d2); 1e8 / -1 when
            out_i[(size_t)i * K + k] = bi[k];                                     // the class has fewer than K points
        }
    }
}

// softmax_k(exp(-d_k / r_k)); stable = any_k(d_k <= r_k)     (nodes.py:179-191)
// radius_mode 0: r_k = radii[idx_k] (surfel -> node);  1: r_k = radii[i] (node -> node, nodes.py:164)
__global__ void knn_weights_kernel(const double* __restrict__ d, const int* __restrict__ idx, int n_cap,
                                   const int* n_dev, const double* __restrict__ radii, int radius_mode,
                                   double* __restrict__ w, unsigned char* __restrict__ stable) {
    const int n = n_active(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e[SB_KNN], mx = -INFINITY;
    bool any = false;
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) {
        const double r = radius_mode ? radii[i] : radii[idx[4 * (size_t)i + k]];
        const double dk = d[4 * (size_t)i + k];
        any |= dk <= r;
        e[k] = exp(-dk / r);
        mx = fmax(mx, e[k]);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) { e[k] = exp(e[k] - mx); s += e[k]; }
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) w[4 * (size_t)i + k] = e[k] / s;
    if (stable && !any) stable[i] = 0;
}

// Recompute weights of existing surfels from their CURRENT positions and OLD indices
// (nodes.py:480-484): d_k = || p - g_k ||  (torch.linalg.norm).
__global__ void reweight_kernel(const double* __restrict__ points, const int* __restrict__ idx, int n_cap,
                                const int* n_dev, const double* __restrict__ ed_points,
                                const double* __restrict__ radii, double* __restrict__ w) {
    const int n = n_active(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = points[3 * (size_t)i], py = points[3 * (size_t)i + 1], pz = points[3 * (size_t)i + 2];
    double e[SB_KNN], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) {
        const int n_ = idx[4 * (size_t)i + k];
        const double dx = px - ed_points[3 * n_], dy = py - ed_points[3 * n_ + 1], dz = pz - ed_points[3 * n_ + 2];
        const double dk = sqrt(dx * dx + dy * dy + dz * dz);
        e[k] = exp(-dk / radii[n_]);
        mx = fmax(mx, e[k]);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) { e[k] = exp(e[k] - mx); s += e[k]; }
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) w[4 * (size_t)i + k] = e[k] / s;
}

// Surfels.update, LM form (use_derived_gradient): points <- T(p); norms <- normalize(sum_k w_k (R(q_k) n + b_k))
// -- the translation IS added to the normal, reference quirk (nodes.py:207-213, Appendix B #4).
__global__ void warp_surfels_kernel(double* __restrict__ points, double* __restrict__ norms,
                                    const int* __restrict__ idx, const double* __restrict__ w, int n_cap,
                                    const int* n_dev, const double* __restrict__ ed_points,
                                    const double* __restrict__ beta) {
    const int n = n_active(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = v3(points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]);
    const V3 nr = v3(norms[3 * (size_t)i], norms[3 * (size_t)i + 1], norms[3 * (size_t)i + 2]);
    V3 T = v3(0, 0, 0), Nn = v3(0, 0, 0);
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) {
        const int n_ = idx[4 * (size_t)i + k];
        const double wk = w[4 * (size_t)i + k];
        const double* g = ed_points + 3 * n_;
        const double* b = beta + 7 * n_;
        const V3 gk = v3(g[0], g[1], g[2]);
        const V3 qv = v3(b[1], b[2], b[3]);
        V3 cp;
        V3 tv = quat_rot_ref(v3(subr(p.x, gk.x), subr(p.y, gk.y), subr(p.z, gk.z)), b[0], qv, cp);
        tv.x = addr(addr(tv.x, b[4]), gk.x); tv.y = addr(addr(tv.y, b[5]), gk.y); tv.z = addr(addr(tv.z, b[6]), gk.z);
        V3 tn = quat_rot_ref(nr, b[0], qv, cp);
        tn.x = addr(tn.x, b[4]); tn.y = addr(tn.y, b[5]); tn.z = addr(tn.z, b[6]);
        if (k == 0) {
            T = v3(mulr(wk, tv.x), mulr(wk, tv.y), mulr(wk, tv.z));
            Nn = v3(mulr(wk, tn.x), mulr(wk, tn.y), mulr(wk, tn.z));
        } else {
            T.x = addr(T.x, mulr(wk, tv.x)); T.y = addr(T.y, mulr(wk, tv.y)); T.z = addr(T.z, mulr(wk, tv.z));
            Nn.x = addr(Nn.x, mulr(wk, tn.x)); Nn.y = addr(Nn.y, mulr(wk, tn.y)); Nn.z = addr(Nn.z, mulr(wk, tn.z));
        }
    }
    points[3 * (size_t)i] = T.x; points[3 * (size_t)i + 1] = T.y; points[3 * (size_t)i + 2] = T.z;
    const double nn = fmax(sqrt(Nn.x * Nn.x + Nn.y * Nn.y + Nn.z * Nn.z), 1e-12);   // F.normalize eps
    norms[3 * (size_t)i] = Nn.x / nn; norms[3 * (size_t)i + 1] = Nn.y / nn; norms[3 * (size_t)i + 2] = Nn.z / nn;
}

// JSD of two class distributions (utils/utils.py:244-254): KL(P|Q) = sum P log(P/(Q+eps)+eps), eps = 1e-13
__device__ __forceinline__ double jsd_dev(const double* __restrict__ P, const double* __restrict__ Q, int C) {
    double k1 = 0.0, k2 = 0.0;
    for (int q = 0; q < C; ++q) {
        const double m = 0.5 * (P[q] + Q[q]);
        k1 += P[q] * log(P[q] / (m + 1e-13) + 1e-13);
        k2 += Q[q] * log(Q[q] / (m + 1e-13) + 1e-13);
    }
    return 0.5 * (k1 + k2);
}

// semantic kNN weights (nodes.py:183-189): softmax_k( exp(-JSD)^(1/2) * exp(-d/r)^(1/2) )
__global__ void reweight_semantic_kernel(const double* __restrict__ points, const int* __restrict__ idx, int n_cap,
                                         const int* n_dev, const double* __restrict__ ed_points,
                                         const double* __restrict__ radii, const double* __restrict__ ed_seg_conf,
                                         const double* __restrict__ seg_conf, int C, double* __restrict__ w) {
    const int n = n_active(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e[SB_KNN], mx = -INFINITY, sum = 0.0;
    for (int k = 0; k < SB_KNN; ++k) {
        const int n_ = idx[4 * (size_t)i + k];
        const double dx = points[3 * (size_t)i] - ed_points[3 * n_], dy = points[3 * (size_t)i + 1] - ed_points[3 * n_ + 1],
                     dz = points[3 * (size_t)i + 2] - ed_points[3 * n_ + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        const double js = jsd_dev(ed_seg_conf + (size_t)n_ * C, seg_conf + (size_t)i * C, C);
        e[k] = sqrt(exp(-js)) * sqrt(exp(-d / radii[n_]));
        mx = fmax(mx, e[k]);
    }
    for (int k = 0; k < SB_KNN; ++k) { e[k] = exp(e[k] - mx); sum += e[k]; }
    for (int k = 0; k < SB_KNN; ++k) w[4 * (size_t)i + k] = e[k] / sum;
}

// ED nodes: points += b; norms <- normalize(R(q) n)          (nodes.py:215-223)
__global__ void warp_nodes_kernel(double* __restrict__ ed_points, double* __restrict__ ed_norms,
                                  const double* __restrict__ beta, int J) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    const double* b = beta + 7 * j;
    V3 cp;
    const V3 tn = quat_rot_ref(v3(ed_norms[3 * j], ed_norms[3 * j + 1], ed_norms[3 * j + 2]), b[0], v3(b[1], b[2], b[3]), cp);
    const double nn = fmax(sqrt(tn.x * tn.x + tn.y * tn.y + tn.z * tn.z), 1e-12);
    ed_norms[3 * j] = tn.x / nn; ed_norms[3 * j + 1] = tn.y / nn; ed_norms[3 * j + 2] = tn.z / nn;
    ed_points[3 * j] += b[4]; ed_points[3 * j + 1] += b[5]; ed_points[3 * j + 2] += b[6];
}

}  // namespace

extern "C" {

int sb_knn(const double* query, int nq_cap, const int* nq_dev, const double* ref, int nref, int dim, int K,
           double* out_dist, int* out_idx, void* stream) {
    return sb_knn_class(query, nq_cap, nq_dev, nullptr, ref, nref, nullptr, dim, K, out_dist, out_idx, stream);
}

int sb_knn_class(const double* query, int nq_cap, const int* nq_dev, const int* qseg, const double* ref, int nref,
                 const int* rseg, int dim, int K, double* out_dist, int* out_idx, void* stream) {
    if (!query || !ref || !out_dist || !out_idx || ((qseg == nullptr) != (rseg == nullptr))) return SB_ERR_ARG;
    if (K < 1 || K > KNN_MAXK || nref < K || (dim != 2 && dim != 3)) return SB_ERR_ARG;
    if (nq_cap <= 0) return SB_OK;
    const int blocks = (nq_cap + KNN_BLOCK - 1) / KNN_BLOCK;
    cudaStream_t s = (cudaStream_t)stream;
#define SB_KNN_CASE(KK) \
    case KK: knn_kernel<KK><<<blocks, KNN_BLOCK, 0, s>>>(query, nq_cap, nq_dev, ref, nref, dim, out_dist, out_idx, qseg, rseg); break;
    switch (K) {
        SB_KNN_CASE(1) SB_KNN_CASE(2) SB_KNN_CASE(3) SB_KNN_CASE(4) SB_KNN_CASE(5) SB_KNN_CASE(6) SB_KNN_CASE(7)
        SB_KNN_CASE(8)
    }
#undef SB_KNN_CASE
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_knn_weights(const double* dist, const int* idx, int n_cap, const int* n_dev, const double* radii,
                   int radius_mode, double* w, unsigned char* stable, void* stream) {
    if (!dist || !idx || !radii || !w) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    knn_weights_kernel<<<(n_cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dist, idx, n_cap, n_dev, radii,
                                                                             radius_mode, w, stable);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_reweight(const double* points, const int* idx, int n_cap, const int* n_dev, const double* ed_points,
                const double* radii, double* w, void* stream) {
    if (!points || !idx || !ed_points || !radii || !w) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    reweight_kernel<<<(n_cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, idx, n_cap, n_dev, ed_points,
                                                                          radii, w);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_reweight_semantic(const double* points, const int* idx, int n_cap, const int* n_dev, const double* ed_points,
                         const double* radii, const double* ed_seg_conf, const double* seg_conf, int C, double* w,
                         void* stream) {
    if (!points || !idx || !ed_points || !radii || !ed_seg_conf || !seg_conf || !w || C < 1 || C > 8) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    reweight_semantic_kernel<<<(n_cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, idx, n_cap, n_dev, ed_points,
                                                                                   radii, ed_seg_conf, seg_conf, C, w);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_warp_update(double* points, double* norms, const int* idx, const double* w, int n_cap, const int* n_dev,
                   double* ed_points, double* ed_norms, const double* beta, int J, void* stream) {
    if (!points || !norms || !idx || !w || !ed_points || !ed_norms || !beta || J <= 0) return SB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_cap > 0) {
        warp_surfels_kernel<<<(n_cap + 255) / 256, 256, 0, s>>>(points, norms, idx, w, n_cap, n_dev, ed_points, beta);
        SB_CHECK_LAUNCH();
    }
    warp_nodes_kernel<<<(J + 127) / 128, 128, 0, s>>>(ed_points, ed_norms, beta, J);   // after the surfels
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"

extern "C" int sb_version(void) { return 100; }
