// Kernels behind the helper functions the reference exports by name ("helpers imported by name
// elsewhere"), and the producer-side pieces next to the path (8(f) rows 1, 3):
//   get_skew / transformQuatT / Trans_points      /root/reference/super/utils.py:4-71
//   pcd2depth, KLD / JSD                           /root/reference/utils/utils.py:161-184,244-254
//   per-residual ARAP / Rot losses                 /root/reference/super/loss.py:428-437,487-490 (forward(grad=False))
//   torch_dilate and the valid-mask morphology     /root/reference/utils/utils.py:152-157, utils/data_loader.py:394-397
//   SSIM depth confidence                          /root/reference/utils/data_loader.py:360-372,477-479 (Project3D layers.py:173-193)
//   surfel splat renderer (pulsar's role)          /root/reference/renderer/renderer.py:12-78, super/nodes.py:630-650
// The device functions are the ones the fused kernels use (common.cuh): the face and the hot path cannot drift apart.
#include "common.cuh"
#include "super_b200.h"

namespace {

// d[R(q)v]/dq as 3x4 (col 0 = d/dqw, cols 1..3 = d/dqv), reference operation order for the skew term
__device__ __forceinline__ void quat_jac34(const V3& v, double qw, const V3& qv, const V3& cp, double* J /* 12, row-major */) {
    const double qd = (qv.x * v.x + qv.y * v.y) + qv.z * v.z;
    const double q[3] = {qv.x, qv.y, qv.z}, vv[3] = {v.x, v.y, v.z};
    const double sk[3][3] = {{0, -v.z, v.y}, {v.z, 0, -v.x}, {-v.y, v.x, 0}};
    const double c[3] = {cp.x, cp.y, cp.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        J[4 * i] = 2.0 * c[i];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            J[4 * i + 1 + j] = 2.0 * ((((i == j ? qd : 0.0) + q[i] * vv[j]) - 2.0 * (vv[i] * q[j])) - qw * sk[i][j]);
    }
}

__global__ void get_skew_kernel(const double* __restrict__ a, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a1 = a[3 * i], a2 = a[3 * i + 1], a3 = a[3 * i + 2];
    // reference: stack([ [0,a3,-a2], [-a3,0,a1], [a2,-a1,0] ], dim=3) on (...,3) inputs -> out[..., r, c] = row_c[r]
    double* o = out + 9 * i;
    o[0] = 0.0; o[1] = -a3; o[2] = a2;
    o[3] = a3; o[4] = 0.0; o[5] = -a1;
    o[6] = -a2; o[7] = a1; o[8] = 0.0;
}

__global__ void transform_quat_kernel(const double* __restrict__ v, const double* __restrict__ beta, long long n, int bdim,
                                      double* __restrict__ tv, double* __restrict__ jac) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* b = beta + (size_t)bdim * i;
    const V3 x = v3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    const V3 qv = v3(b[1], b[2], b[3]);
    V3 cp;
    V3 r = quat_rot_ref(x, b[0], qv, cp);
    if (bdim == 7) { r.x = addr(r.x, b[4]); r.y = addr(r.y, b[5]); r.z = addr(r.z, b[6]); }
    tv[3 * i] = r.x; tv[3 * i + 1] = r.y; tv[3 * i + 2] = r.z;
    if (jac) quat_jac34(x, b[0], qv, cp, jac + 12 * i);
}

// Trans_points: T(p) = sum_k w_k [ R(q_k) d_k + b_k + g_k ], Jacobian blocks scaled by w_k
__global__ void trans_points_kernel(const double* __restrict__ d, const double* __restrict__ g, const double* __restrict__ beta,
                                    const double* __restrict__ w, long long n, int K, double* __restrict__ out,
                                    double* __restrict__ jac) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 T = v3(0, 0, 0);
    for (int k = 0; k < K; ++k) {
        const size_t ik = (size_t)i * K + k;
        const double* b = beta + 7 * ik;
        const V3 x = v3(d[3 * ik], d[3 * ik + 1], d[3 * ik + 2]);
        const V3 qv = v3(b[1], b[2], b[3]);
        V3 cp;
        V3 r = quat_rot_ref(x, b[0], qv, cp);
        r.x = addr(addr(r.x, b[4]), g[3 * ik]); r.y = addr(addr(r.y, b[5]), g[3 * ik + 1]); r.z = addr(addr(r.z, b[6]), g[3 * ik + 2]);
        const double wk = w ? w[ik] : 1.0;
        if (k == 0) T = v3(mulr(wk, r.x), mulr(wk, r.y), mulr(wk, r.z));
        else { T.x = addr(T.x, mulr(wk, r.x)); T.y = addr(T.y, mulr(wk, r.y)); T.z = addr(T.z, mulr(wk, r.z)); }
        if (jac) {
            double J[12];
            quat_jac34(x, b[0], qv, cp, J);
            for (int e = 0; e < 12; ++e) jac[12 * ik + e] = J[e] * wk;
        }
    }
    out[3 * i] = T.x; out[3 * i + 1] = T.y; out[3 * i + 2] = T.z;
}

__global__ void pcd2depth_kernel(const double* __restrict__ pcd, long long n, Cam cam, int margin, double* __restrict__ vf,
                                 double* __restrict__ uf, long long* __restrict__ vi, long long* __restrict__ ui,
                                 long long* __restrict__ coords, unsigned char* __restrict__ valid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u, v;
    project_ref(v3(pcd[3 * i], pcd[3 * i + 1], pcd[3 * i + 2]), cam, u, v);
    const long long ur = round_ll(u), vr = round_ll(v);
    if (vf) { vf[i] = v; uf[i] = u; }
    if (vi) { vi[i] = vr; ui[i] = ur; }
    coords[i] = vr * cam.W + ur;
    valid[i] = (vr >= margin && vr < cam.H - 1 - margin && ur >= margin && ur < cam.W - 1 - margin) ? 1 : 0;
}

__device__ __forceinline__ double kld_dev(const double* P, const double* Q, int C, double eps, int sp, int sq) {
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += P[c * sp] * log(P[c * sp] / (Q[c * sq] + eps) + eps);
    return s;
}
__global__ void kld_jsd_kernel(const double* __restrict__ P, const double* __restrict__ Q, long long n, int C, double eps,
                               int jsd, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = P + (size_t)C * i;
    const double* q = Q + (size_t)C * i;
    if (!jsd) { out[i] = kld_dev(p, q, C, eps, 1, 1); return; }
    double M[8];
    for (int c = 0; c < C; ++c) M[c] = 0.5 * (p[c] + q[c]);
    out[i] = 0.5 * (kld_dev(p, M, C, eps, 1, 1) + kld_dev(q, M, C, eps, 1, 1));
}

// per-residual losses of the regularisers: arap (J*K*3) and rot (J) squared residuals (LM_Solver sums them)
__global__ void reg_residuals_kernel(const double* __restrict__ ed_points, const int* __restrict__ ed_knn,
                                     const double* __restrict__ beta, int J, double lam_arap, double lam_rot,
                                     double* __restrict__ arap_r2, float* __restrict__ rot_r2) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (arap_r2 && tid < J * SB_KNN) {
        const int j = tid / SB_KNN, n = ed_knn[tid];
        const V3 d = v3(ed_points[3 * j] - ed_points[3 * n], ed_points[3 * j + 1] - ed_points[3 * n + 1],
                        ed_points[3 * j + 2] - ed_points[3 * n + 2]);
        const double* bn = beta + 7 * n;
        const double* bj = beta + 7 * j;
        V3 cp;
        const V3 tv = quat_rot_ref(d, bn[0], v3(bn[1], bn[2], bn[3]), cp);
        const double r[3] = {lam_arap * ((tv.x + bn[4]) - (d.x + bj[4])), lam_arap * ((tv.y + bn[5]) - (d.y + bj[5])),
                             lam_arap * ((tv.z + bn[6]) - (d.z + bj[6]))};
        for (int c = 0; c < 3; ++c) arap_r2[3 * tid + c] = r[c] * r[c];
    }
    if (rot_r2 && tid < J) {
        const float lam = (float)lam_rot;
        float q[4];
        for (int a = 0; a < 4; ++a) q[a] = (float)beta[7 * tid + a];
        const float s = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
        const float r = lam * (1.f - s);
        rot_r2[tid] = r * r;
    }
}

// torch_dilate for one channel: box filter > 0 with conv2d(padding='same') geometry: window rows [y - pl, y + k - 1 - pl],
// pl = (k - 1) / 2 (the extra element of an even kernel is on the bottom / right), zero padding
__global__ void dilate_box_kernel(const unsigned char* __restrict__ in, int H, int W, int k, int invert_in, int invert_out,
                                  unsigned char* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int pl = (k - 1) / 2;
    bool any = false;
    for (int dy = -pl; dy < k - pl && !any; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -pl; dx < k - pl; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            const bool v = (in[yy * W + xx] != 0) != (invert_in != 0);
            if (v) { any = true; break; }
        }
    }
    out[y * W + x] = (any != (invert_out != 0)) ? 1 : 0;
}

// ---- SSIM depth confidence --------------------------------------------------------------------------------------
// warp: every pixel's back-projected point (depth * inv_K [x,y,1], float32 like BackprojectDepth) is projected with
// P = (K T)[:3] (Project3D, layers.py:181-193), normalised by (W-1, H-1), and the frame's own colour image is sampled there
// (F.grid_sample defaults: bilinear, zeros padding, align_corners=False)
struct SsimArgs { const float* depth; const float* color; float ik[9]; float P[12]; int H, W; };

__global__ void ssim_warp_kernel(SsimArgs a, float* __restrict__ warp /* (3,H,W) */) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.W) return;
    const int W = a.W, H = a.H, P = W * H, p = y * W + x;
    const float d = a.depth[p];
    const float fx = (float)x, fy = (float)y;
    const float cx = d * ((a.ik[0] * fx + a.ik[1] * fy) + a.ik[2]);
    const float cy = d * ((a.ik[3] * fx + a.ik[4] * fy) + a.ik[5]);
    const float cz = d * ((a.ik[6] * fx + a.ik[7] * fy) + a.ik[8]);
    const float px = ((a.P[0] * cx + a.P[1] * cy) + a.P[2] * cz) + a.P[3];
    const float py = ((a.P[4] * cx + a.P[5] * cy) + a.P[6] * cz) + a.P[7];
    const float pz = ((a.P[8] * cx + a.P[9] * cy) + a.P[10] * cz) + a.P[11];
    float gx = (px / (pz + 1e-7f)) / (float)(W - 1), gy = (py / (pz + 1e-7f)) / (float)(H - 1);
    gx = (gx - 0.5f) * 2.f; gy = (gy - 0.5f) * 2.f;
    // grid_sample, align_corners=False: pixel = ((g + 1) * size - 1) / 2
    const float sx = ((gx + 1.f) * (float)W - 1.f) * 0.5f, sy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
    const float x0f = floorf(sx), y0f = floorf(sy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float wx1 = sx - x0f, wy1 = sy - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    for (int c = 0; c < 3; ++c) {
        float v = 0.f;
        if (isfinite(sx) && isfinite(sy)) {
            const float* img = a.color + (size_t)c * P;
            auto at = [&](int yy, int xx) { return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? img[yy * W + xx] : 0.f; };
            v = (at(y0, x0) * wx0 * wy0 + at(y0, x0 + 1) * wx1 * wy0) + (at(y0 + 1, x0) * wx0 * wy1 + at(y0 + 1, x0 + 1) * wx1 * wy1);
        } else {
            v = __int_as_float(0x7fc00000);
        }
        warp[(size_t)c * P + p] = v;
    }
}

// structural_similarity(full=True) per channel: 7x7 uniform window, reflect borders, sample covariance, K1 = 0.01,
// K2 = 0.03, data_range R; mean over the 3 channels; then confs = 0.5 confs + 0.5 sigmoid(mean)   (float64 inside like
// skimage, which promotes the images)
__global__ void ssim_conf_kernel(const float* __restrict__ xw, const float* __restrict__ yt, int H, int W, double R,
                                 float* __restrict__ confs, float* __restrict__ ssim_out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int P = H * W;
    const double C1 = (0.01 * R) * (0.01 * R), C2 = (0.03 * R) * (0.03 * R), NP = 49.0, cov_norm = NP / (NP - 1.0);
    double mean_s = 0.0;
    for (int c = 0; c < 3; ++c) {
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int dy = -3; dy <= 3; ++dy) {
            int yy = y + dy;
            yy = yy < 0 ? -yy - 1 : (yy >= H ? 2 * H - 1 - yy : yy);          // scipy 'reflect': d c b a | a b c d | d c b a
            for (int dx = -3; dx <= 3; ++dx) {
                int xx = x + dx;
                xx = xx < 0 ? -xx - 1 : (xx >= W ? 2 * W - 1 - xx : xx);
                const double a = xw[(size_t)c * P + yy * W + xx], b = yt[(size_t)c * P + yy * W + xx];
                sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
            }
        }
        const double ux = sx / NP, uy = sy / NP, uxx = sxx / NP, uyy = syy / NP, uxy = sxy / NP;
        const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
        const double A1 = 2 * ux * uy + C1, A2 = 2 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
        mean_s += (A1 * A2) / (B1 * B2);
    }
    const float s = (float)(mean_s / 3.0);
    if (ssim_out) ssim_out[y * W + x] = s;
    const float sig = 1.f / (1.f + expf(-s));
    confs[y * W + x] = 0.5f * confs[y * W + x] + 0.5f * sig;
}

// ---- surfel splat renderer ---------------------------------------------------------------------------------------
// Every surfel is a sphere of radius `rad` at its position (pulsar's model with gamma -> 0: the nearest sphere on a ray
// wins).  Pass 1: per surfel, the pixels of its projected disc compete with atomicMin on (depth bits << 32 | surfel id);
// pass 2: per pixel, the winner's colour or the background.  Depth along the ray uses the sphere's front surface.
__global__ void splat_zbuf_kernel(const double* __restrict__ points, const unsigned char* __restrict__ mask, int n_cap,
                                  const int* __restrict__ n_dev, Cam cam, double rad, unsigned long long* __restrict__ zbuf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active(n_cap, n_dev) || (mask && !mask[i])) return;
    const double X = points[3 * (size_t)i], Y = points[3 * (size_t)i + 1], Z = points[3 * (size_t)i + 2];
    if (!(Z > rad) || !(Z < 1e9)) return;
    const double u = X * cam.fx / Z + cam.cx, v = Y * cam.fy / Z + cam.cy;
    const double ru = rad * cam.fx / Z, rv = rad * cam.fy / Z;
    const int x0 = max(0, (int)ceil(u - ru)), x1 = min(cam.W - 1, (int)floor(u + ru));
    const int y0 = max(0, (int)ceil(v - rv)), y1 = min(cam.H - 1, (int)floor(v + rv));
    if (x1 - x0 > 64 || y1 - y0 > 64) return;                       // a sphere in the camera's face: not a surfel
    bool any = false;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) {
            // ray through the pixel centre: direction (dx, dy, 1); nearest intersection with the sphere
            const double dx = (xx - cam.cx) / cam.fx, dy = (yy - cam.cy) / cam.fy;
            const double dd = dx * dx + dy * dy + 1.0, dp = dx * X + dy * Y + Z, pp = X * X + Y * Y + Z * Z;
            const double disc = dp * dp - dd * (pp - rad * rad);
            if (disc < 0.0) continue;
            const float t = (float)((dp - sqrt(disc)) / dd);        // depth (z) of the hit
            if (!(t > 0.f)) continue;
            atomicMin(zbuf + (size_t)yy * cam.W + xx, ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)i);
            any = true;
        }
    if (!any) {   // sub-pixel sphere: it still owns the pixel its centre falls into (pulsar's minimum footprint)
        const long long xr = round_ll(u), yr = round_ll(v);
        if (xr >= 0 && xr < cam.W && yr >= 0 && yr < cam.H)
            atomicMin(zbuf + (size_t)yr * cam.W + xr, ((unsigned long long)__float_as_uint((float)(Z - rad)) << 32) | (unsigned)i);
    }
}

__global__ void splat_resolve_kernel(const unsigned long long* __restrict__ zbuf, const float* __restrict__ colors, int P,
                                     float bg0, float bg1, float bg2, float* __restrict__ img, float* __restrict__ depth,
                                     int* __restrict__ index) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const unsigned long long z = zbuf[p];
    if (z == ~0ull) {
        img[3 * p] = bg0; img[3 * p + 1] = bg1; img[3 * p + 2] = bg2;
        if (depth) depth[p] = 0.f;
        if (index) index[p] = -1;
    } else {
        const unsigned i = (unsigned)(z & 0xffffffffu);
        img[3 * p] = colors[3 * (size_t)i]; img[3 * p + 1] = colors[3 * (size_t)i + 1]; img[3 * p + 2] = colors[3 * (size_t)i + 2];
        if (depth) depth[p] = __uint_as_float((unsigned)(z >> 32));
        if (index) index[p] = (int)i;
    }
}

inline int blocks_for(long long n, int t) { return (int)((n + t - 1) / t); }

}  // namespace

extern "C" {

int sb_get_skew(const double* a, long long n, double* out, void* stream) {
    if (!a || !out || n < 0) return SB_ERR_ARG;
    if (n == 0) return SB_OK;
    get_skew_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, n, out);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_transform_quat(const double* v, const double* beta, long long n, int beta_dim, double* tv, double* jac, void* stream) {
    if (!v || !beta || !tv || n < 0 || (beta_dim != 4 && beta_dim != 7)) return SB_ERR_ARG;
    if (n == 0) return SB_OK;
    transform_quat_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(v, beta, n, beta_dim, tv, jac);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_trans_points(const double* d, const double* g, const double* beta, const double* w, long long n, int K, double* out,
                    double* jac, void* stream) {
    if (!d || !g || !beta || !out || n < 0 || K < 1) return SB_ERR_ARG;
    if (n == 0) return SB_OK;
    trans_points_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(d, g, beta, w, n, K, out, jac);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_pcd2depth(const double* pcd, long long n, const double* intr, int H, int W, int valid_margin, double* v_float,
                 double* u_float, long long* v_round, long long* u_round, long long* coords, unsigned char* valid,
                 void* stream) {
    if (!pcd || !intr || !coords || !valid || n < 0 || (!v_float != !u_float) || (!v_round != !u_round)) return SB_ERR_ARG;
    if (n == 0) return SB_OK;
    Cam cam;
    cam.fx = intr[0]; cam.fy = intr[1]; cam.cx = intr[2]; cam.cy = intr[3]; cam.H = H; cam.W = W;
    pcd2depth_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(pcd, n, cam, valid_margin, v_float, u_float,
                                                                          v_round, u_round, coords, valid);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_kld_jsd(const double* P, const double* Q, long long n, int C, double eps, int jsd, double* out, void* stream) {
    if (!P || !Q || !out || n < 0 || C < 1 || C > 8) return SB_ERR_ARG;
    if (n == 0) return SB_OK;
    kld_jsd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(P, Q, n, C, eps, jsd, out);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_reg_residuals(const double* ed_points, const int* ed_knn, const double* beta, int J, double lam_arap, double lam_rot,
                     double* arap_r2, float* rot_r2, void* stream) {
    if (!ed_points || !ed_knn || !beta || J <= 0 || (!arap_r2 && !rot_r2)) return SB_ERR_ARG;
    reg_residuals_kernel<<<blocks_for((long long)J * SB_KNN, 256), 256, 0, (cudaStream_t)stream>>>(ed_points, ed_knn, beta, J,
                                                                                                  lam_arap, lam_rot, arap_r2,
                                                                                                  rot_r2);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_dilate_box(const unsigned char* in, int H, int W, int kernel, int invert_in, int invert_out, unsigned char* out,
                  void* stream) {
    if (!in || !out || in == out || H <= 0 || W <= 0 || kernel < 1) return SB_ERR_ARG;
    dilate_box_kernel<<<dim3((W + 127) / 128, H), 128, 0, (cudaStream_t)stream>>>(in, H, W, kernel, invert_in, invert_out, out);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_ssim_conf(const float* depth, const float* color, const float* inv_K3x3, const float* KT3x4, int H, int W,
                 double data_range, float* warp_scratch, float* confs, float* ssim_out, void* stream) {
    if (!depth || !color || !inv_K3x3 || !KT3x4 || !warp_scratch || !confs || H < 7 || W < 7) return SB_ERR_ARG;
    SsimArgs a;
    a.depth = depth; a.color = color; a.H = H; a.W = W;
    for (int i = 0; i < 9; ++i) a.ik[i] = inv_K3x3[i];
    for (int i = 0; i < 12; ++i) a.P[i] = KT3x4[i];
    dim3 grid((W + 127) / 128, H);
    ssim_warp_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a, warp_scratch);
    SB_CHECK_LAUNCH();
    ssim_conf_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(warp_scratch, color, H, W, data_range, confs, ssim_out);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_render_splats(const double* points, const float* colors, const unsigned char* mask, int n_cap, const int* n_dev,
                     const double* intr, int H, int W, double rad, const float* bg3, unsigned long long* zbuf, float* img,
                     float* depth, int* index, void* stream) {
    if (!points || !colors || !intr || !bg3 || !zbuf || !img || H <= 0 || W <= 0 || !(rad > 0.0)) return SB_ERR_ARG;
    Cam cam;
    cam.fx = intr[0]; cam.fy = intr[1]; cam.cx = intr[2]; cam.cy = intr[3]; cam.H = H; cam.W = W;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(zbuf, 0xff, (size_t)H * W * sizeof(unsigned long long), st) != cudaSuccess) return SB_ERR_CUDA;
    if (n_cap > 0) {
        splat_zbuf_kernel<<<blocks_for(n_cap, 256), 256, 0, st>>>(points, mask, n_cap, n_dev, cam, rad, zbuf);
        SB_CHECK_LAUNCH();
    }
    splat_resolve_kernel<<<blocks_for((long long)H * W, 256), 256, 0, st>>>(zbuf, colors, H * W, bg3[0], bg3[1], bg3[2], img,
                                                                          depth, index);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
