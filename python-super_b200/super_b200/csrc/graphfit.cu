// The autograd optimiser of the reference (GraphFit, /root/reference/super/deform_mesh.py:25-379) as fused
// loss + ANALYTIC gradient kernels: no autograd tape, no (N,4,3,7) temporaries, no per-iteration renderer call.
//
//   deform_verts dv (J+1,7) f64, row J = global transform [q_g | t_g]
//   p'_i = R(q_g) T_i + t_g,   T_i = sum_k w_ik [ R(q_k)(p_i - g_k) + g_k + b_k ]          deform_mesh.py:198-230
//   point-plane   w_pp  sum_i omega_i ( n~_i . (p'_i - o~_i) )^2                            loss.py:293-401
//                 (o~, n~ bilinear samples of the new frame at the projection of p'_i, zero-filled corners, all four
//                 corners must be valid; projection validity tests the ROUNDED pixel with margin 1; omega = 1, or
//                 exp(-0.1 JSD(seg_conf_i, softmax(s~_i))) (soft), or [seg_i == argmax s~_i] (hard), detached)
//   ARAP          w_a   sum_j sum_k w^ED_jk | R(q_n)(g_j-g_n) + b_n - fl32(g_j-g_n) - b_j |^2  loss.py:458-473
//   Rot           w_r   sum_{j<=J} (1 - |q_j|^2)^2                                          loss.py:502-505
//   Face          w_f   sum_f (A_f(g') - A0_f)^2,  g'_j = R(q_g)(g_j + b_j) + t_g           deform_mesh.py:51-60
//   bn_morph      w_m   mean_{i in S} m_i,  m_i = mean_{k<2} |e_k - (x_i,y_i)|^2 > 15        deform_mesh.py:126-194
//
// Gradients (what loss.backward() produces in the reference): with a = dl/dp',
//   dl/dt_g += a,  dl/dq_g += Jq(q_g, T)^T a,  a_T = M(q_g)^T a,  dl/db_k += w_k a_T,  dl/dq_k += w_k Jq(q_k, d_k)^T a_T
// where R(q)v = v + 2 q_w (q_v x v) + 2 q_v x (q_v x v) = M(q) v (q NOT normalised) and
//   Jq(q,v)^T a = [ 2 a.(q_v x v) ;  2( (q_v.v) a + (q_v.a) v - 2 (v.a) q_v + q_w (v x a) ) ].
// Node gradients are accumulated in a per-CTA shared table (native f64 shared atomics) and flushed once.
#include "common.cuh"
#include "super_b200.h"

namespace {

constexpr int GF_BLOCK = 256;
constexpr int GF_MAXC = 8;          // segmentation classes

struct GfArgs {
    const double* points; const int* knn_idx; const double* knn_w; const unsigned char* stable;
    int n_cap; const int* n_dev;
    const double* ed_points; int J;
    const double* dv;                  // (J+1,7)
    Cam cam;
};

__device__ __forceinline__ V3 sub3(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 add3(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 scl3(double s, const V3& a) { return v3(s * a.x, s * a.y, s * a.z); }

// R(q) v (tolerance-checked path: FMAs allowed)
__device__ __forceinline__ V3 qrot(const V3& v, double qw, const V3& qv) {
    const V3 cp = cross3(qv, v), c2 = cross3(qv, cp);
    return v3(v.x + 2.0 * qw * cp.x + 2.0 * c2.x, v.y + 2.0 * qw * cp.y + 2.0 * c2.y, v.z + 2.0 * qw * cp.z + 2.0 * c2.z);
}
// M(q)^T a = a - 2 q_w (q_v x a) + 2 q_v x (q_v x a)
__device__ __forceinline__ V3 qrot_t(const V3& a, double qw, const V3& qv) {
    const V3 cp = cross3(qv, a), c2 = cross3(qv, cp);
    return v3(a.x - 2.0 * qw * cp.x + 2.0 * c2.x, a.y - 2.0 * qw * cp.y + 2.0 * c2.y, a.z - 2.0 * qw * cp.z + 2.0 * c2.z);
}
// Jq(q,v)^T a -> (gw, gv)
__device__ __forceinline__ void qgrad(const V3& v, double qw, const V3& qv, const V3& a, double& gw, V3& gv) {
    const V3 cp = cross3(qv, v);
    gw = 2.0 * dot3(a, cp);
    const double qd = dot3(qv, v), qa = dot3(qv, a), va = dot3(v, a);
    const V3 vxa = cross3(v, a);
    gv = v3(2.0 * (qd * a.x + qa * v.x - 2.0 * va * qv.x + qw * vxa.x),
            2.0 * (qd * a.y + qa * v.y - 2.0 * va * qv.y + qw * vxa.y),
            2.0 * (qd * a.z + qa * v.z - 2.0 * va * qv.z + qw * vxa.z));
}

// per-surfel forward warp; returns false for rows that do not take part (unstable)
struct Warped {
    int idx[SB_KNN];
    double w[SB_KNN];
    V3 p, T, pp;           // surfel, ED-warped, globally transformed
};
__device__ __forceinline__ bool gf_warp(const GfArgs& a, int i, Warped& s) {
    if (a.stable && !a.stable[i]) return false;
    const double* pp = a.points + 3 * (size_t)i;
    s.p = v3(pp[0], pp[1], pp[2]);
    const int4 id4 = *reinterpret_cast<const int4*>(a.knn_idx + 4 * (size_t)i);
    s.idx[0] = id4.x; s.idx[1] = id4.y; s.idx[2] = id4.z; s.idx[3] = id4.w;
    const double2 w01 = *reinterpret_cast<const double2*>(a.knn_w + 4 * (size_t)i);
    const double2 w23 = *reinterpret_cast<const double2*>(a.knn_w + 4 * (size_t)i + 2);
    s.w[0] = w01.x; s.w[1] = w01.y; s.w[2] = w23.x; s.w[3] = w23.y;
    V3 T = v3(0, 0, 0);
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) {
        const double* g = a.ed_points + 3 * s.idx[k];
        const double* b = a.dv + 7 * s.idx[k];
        const V3 gk = v3(__ldg(g), __ldg(g + 1), __ldg(g + 2));
        V3 tv = qrot(sub3(s.p, gk), __ldg(b), v3(__ldg(b + 1), __ldg(b + 2), __ldg(b + 3)));
        tv = v3(tv.x + __ldg(b + 4) + gk.x, tv.y + __ldg(b + 5) + gk.y, tv.z + __ldg(b + 6) + gk.z);
        T = add3(T, scl3(s.w[k], tv));
    }
    s.T = T;
    const double* gl = a.dv + 7 * a.J;
    const V3 r = qrot(T, gl[0], v3(gl[1], gl[2], gl[3]));
    s.pp = v3(r.x + gl[4], r.y + gl[5], r.z + gl[6]);
    return true;
}

constexpr int GF_STAGE = 28 * 33;      // per-warp staging: 28 node-gradient entries x 32 lanes (+1 pad)

// Back-propagate a = dl/dp' of the warp's 32 surfels (ok = lane takes part) into the shared node table and the
// thread's global-row partial.  A lane's 28 node-gradient entries (4 nodes x [q(4) b(3)], node slots in ASCENDING
// node-id order) are staged in shared memory; lanes that share the node SET are summed by 28 lanes (one entry each,
// 32 conflict-free loads) and leave as 28 shared atomics on distinct addresses.  (One atomic per entry and surfel
// was 28 same-address atomics per warp instruction: 340 us per pass, ncu launch list r1d.)
__device__ __forceinline__ void gf_backprop_warp(const GfArgs& a, const Warped& s, const V3& g, bool ok,
                                                 double* __restrict__ tab, double* __restrict__ stage,
                                                 double (&gl_acc)[7]) {
    const int lane = threadIdx.x & 31;
    unsigned long long key = ~0ull;
    if (ok) {
        const double* gl = a.dv + 7 * a.J;
        const double gw0 = gl[0];
        const V3 gqv = v3(gl[1], gl[2], gl[3]);
        double qgw; V3 qgv;
        qgrad(s.T, gw0, gqv, g, qgw, qgv);
        gl_acc[0] += qgw; gl_acc[1] += qgv.x; gl_acc[2] += qgv.y; gl_acc[3] += qgv.z;
        gl_acc[4] += g.x; gl_acc[5] += g.y; gl_acc[6] += g.z;
        const V3 aT = qrot_t(g, gw0, gqv);
        key = 0;
#pragma unroll
        for (int k = 0; k < SB_KNN; ++k) {
            const double* gp = a.ed_points + 3 * s.idx[k];
            const double* b = a.dv + 7 * s.idx[k];
            const V3 dk = v3(s.p.x - __ldg(gp), s.p.y - __ldg(gp + 1), s.p.z - __ldg(gp + 2));
            double kw; V3 kv;
            qgrad(dk, __ldg(b), v3(__ldg(b + 1), __ldg(b + 2), __ldg(b + 3)), aT, kw, kv);
            int slot = 0;
#pragma unroll
            for (int j = 0; j < SB_KNN; ++j) slot += (s.idx[j] < s.idx[k]) ? 1 : 0;
            key |= (unsigned long long)(unsigned)s.idx[k] << (48 - 16 * slot);
            const double wk = s.w[k];
            double* st = stage + (7 * slot) * 33 + lane;
            st[0 * 33] = wk * kw; st[1 * 33] = wk * kv.x; st[2 * 33] = wk * kv.y; st[3 * 33] = wk * kv.z;
            st[4 * 33] = wk * aT.x; st[5 * 33] = wk * aT.y; st[6 * 33] = wk * aT.z;
        }
    }
    __syncwarp();
    unsigned remaining = __ballot_sync(0xffffffffu, ok);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const unsigned long long kk = __shfl_sync(0xffffffffu, key, leader);
        const unsigned m = __ballot_sync(0xffffffffu, key == kk) & remaining;
        if (lane < 28) {
            const double* row = stage + lane * 33;
            double acc = 0.0;
#pragma unroll 8
            for (int l = 0; l < 32; ++l)
                if ((m >> l) & 1u) acc += row[l];
            const int slot = lane / 7, comp = lane - 7 * slot;
            const int node = (int)(kk >> (48 - 16 * slot)) & 0xffff;
            atomicAdd(tab + 7 * node + comp, acc);
        }
        remaining &= ~m;
    }
    __syncwarp();
}

// projection (Z + 1e-8 everywhere: this is what autograd differentiates) and its gradient rows
__device__ __forceinline__ void gf_project(const V3& pp, const Cam& cam, double& u, double& v, V3& du, V3& dvv) {
    const double Zp = pp.z + 1e-8, iz = 1.0 / Zp;
    u = pp.x * cam.fx * iz + cam.cx;
    v = pp.y * cam.fy * iz + cam.cy;
    du = v3(cam.fx * iz, 0.0, -cam.fx * pp.x * iz * iz);
    dvv = v3(0.0, cam.fy * iz, -cam.fy * pp.y * iz * iz);
}

// table flush + global-row / loss reduction shared by the surfel kernels
__device__ __forceinline__ void gf_finish(double* __restrict__ tab, int J, double (&gl_acc)[7], double loss, double cnt,
                                          double* __restrict__ grad, double* __restrict__ loss_out,
                                          double* __restrict__ cnt_out, double* red) {
    __syncthreads();
    for (int e = threadIdx.x; e < 7 * J; e += GF_BLOCK) {
        const double v = tab[e];
        if (v != 0.0) atomicAdd(grad + e, v);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const double s = block_sum<GF_BLOCK>(gl_acc[c], red);
        if (threadIdx.x == 0 && s != 0.0) atomicAdd(grad + 7 * J + c, s);
    }
    const double ls = block_sum<GF_BLOCK>(loss, red);
    if (threadIdx.x == 0 && ls != 0.0) atomicAdd(loss_out, ls);
    if (cnt_out) {
        const double cs = block_sum<GF_BLOCK>(cnt, red);
        if (threadIdx.x == 0 && cs != 0.0) atomicAdd(cnt_out, cs);
    }
}

// ---- point-to-plane (plain / soft-seg / hard-seg) -------------------------------------------------------
struct SegArgs {
    int mode;                         // 0 none, 1 soft, 2 hard
    int C;
    const int* sf_seg;                // (N,)
    const double* sf_seg_conf;        // (N,C)
    const double* trg_seg_conf;       // (P,C) dense map of the new frame (softmax of the scores)
};

__global__ void __launch_bounds__(GF_BLOCK)
gf_data_kernel(GfArgs a, const float4* __restrict__ vmap, const float4* __restrict__ nmap, SegArgs sg, double weight,
               double* __restrict__ grad, double* __restrict__ loss_out) {
    extern __shared__ double tab[];
    __shared__ double red[GF_BLOCK / 32];
    for (int e = threadIdx.x; e < 7 * a.J; e += GF_BLOCK) tab[e] = 0.0;
    __syncthreads();
    const int n = n_active(a.n_cap, a.n_dev);
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    const int H = a.cam.H, W = a.cam.W;
    double gl_acc[7] = {0, 0, 0, 0, 0, 0, 0}, loss = 0.0;
    double* stage = tab + 7 * a.J + (threadIdx.x >> 5) * GF_STAGE;
    for (int ib = i0; ib < i1; ib += GF_BLOCK) {          // warp-uniform trip count: the back-propagation is cooperative
        const int i = ib + threadIdx.x;
        Warped s;
        V3 g = v3(0, 0, 0);
        bool live = false;
        do {
        if (i >= i1 || !gf_warp(a, i, s)) break;
        double u, v; V3 du, dv_;
        gf_project(s.pp, a.cam, u, v, du, dv_);
        if (!(fabs(u) < 1e9 && fabs(v) < 1e9)) break;
        const long long ur = round_ll(u), vr = round_ll(v);
        if (vr < 1 || vr >= H - 2 || ur < 1 || ur >= W - 2) break;             // pcd2depth valid_margin = 1
        const double fv = floor(v), cv = ceil(v), fu = floor(u), cu = ceil(u);
        const int iy[2] = {(int)fv, (int)cv}, ix[2] = {(int)fu, (int)cu};
        const double dy[2] = {fv - v, cv - v}, dx[2] = {fu - u, cu - u};
        const double wy[2] = {fmax(1.0 - fabs(dy[0]), 0.0), fmax(1.0 - fabs(dy[1]), 0.0)};
        const double wx[2] = {fmax(1.0 - fabs(dx[0]), 0.0), fmax(1.0 - fabs(dx[1]), 0.0)};
        V3 o = v3(0, 0, 0), nn = v3(0, 0, 0), o_u = v3(0, 0, 0), o_v = v3(0, 0, 0), n_u = v3(0, 0, 0), n_v = v3(0, 0, 0);
        double sc[GF_MAXC];
#pragma unroll
        for (int c = 0; c < GF_MAXC; ++c) sc[c] = 0.0;
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int yi = c >> 1, xi = c & 1;
            const int pix = iy[yi] * W + ix[xi];
            const float4 pv = __ldg(vmap + pix);
            if (pv.w == 0.f) { ok = false; break; }
            const float4 nv = __ldg(nmap + pix);
            const double wgt = wy[yi] * wx[xi];
            // d|t|/dt = sign(t) (0 at 0), and the clamp passes the gradient only where 1-|t| >= 0
            const double sx = (dx[xi] > 0.0) ? 1.0 : (dx[xi] < 0.0 ? -1.0 : 0.0), sy = (dy[yi] > 0.0) ? 1.0 : (dy[yi] < 0.0 ? -1.0 : 0.0);
            const double gu = wy[yi] * sx, gv = wx[xi] * sy;
            o.x += pv.x * wgt; o.y += pv.y * wgt; o.z += pv.z * wgt;
            nn.x += nv.x * wgt; nn.y += nv.y * wgt; nn.z += nv.z * wgt;
            o_u.x += pv.x * gu; o_u.y += pv.y * gu; o_u.z += pv.z * gu;
            o_v.x += pv.x * gv; o_v.y += pv.y * gv; o_v.z += pv.z * gv;
            n_u.x += nv.x * gu; n_u.y += nv.y * gu; n_u.z += nv.z * gu;
            n_v.x += nv.x * gv; n_v.y += nv.y * gv; n_v.z += nv.z * gv;
            if (sg.mode) {
                const double* tc = sg.trg_seg_conf + (size_t)pix * sg.C;
                for (int q = 0; q < sg.C; ++q) sc[q] += tc[q] * wgt;
            }
        }
        if (!ok) break;
        double omega = 1.0;
        if (sg.mode) {
            // softmax of the sampled (already soft-maxed) confidences: loss.py:357
            double mx = -INFINITY, den = 0.0;
            for (int q = 0; q < sg.C; ++q) mx = fmax(mx, sc[q]);
            for (int q = 0; q < sg.C; ++q) { sc[q] = exp(sc[q] - mx); den += sc[q]; }
            int am = 0;
            for (int q = 0; q < sg.C; ++q) { sc[q] /= den; if (sc[q] > sc[am]) am = q; }
            if (sg.mode == 1) {
                // JSD(P,Q) = (KL(P|M) + KL(Q|M))/2, KL(P|Q) = sum P log(P/(Q+eps)+eps)    utils/utils.py:244-254
                const double* P = sg.sf_seg_conf + (size_t)i * sg.C;
                double k1 = 0.0, k2 = 0.0;
                for (int q = 0; q < sg.C; ++q) {
                    const double m = 0.5 * (P[q] + sc[q]);
                    k1 += P[q] * log(P[q] / (m + 1e-13) + 1e-13);
                    k2 += sc[q] * log(sc[q] / (m + 1e-13) + 1e-13);
                }
                omega = exp(-0.1 * 0.5 * (k1 + k2));
            } else {
                omega = (sg.sf_seg[i] == am) ? 1.0 : 0.0;
            }
        }
        const V3 diff = sub3(s.pp, o);
        const double r = dot3(nn, diff);
        loss += weight * omega * r * r;
        const double cu_ = dot3(diff, n_u) - dot3(nn, o_u), cv_ = dot3(diff, n_v) - dot3(nn, o_v);
        const double f = 2.0 * weight * omega * r;
        g = v3(f * (nn.x + cu_ * du.x + cv_ * dv_.x), f * (nn.y + cu_ * du.y + cv_ * dv_.y),
               f * (nn.z + cu_ * du.z + cv_ * dv_.z));
        live = true;
        } while (false);
        gf_backprop_warp(a, s, g, live, tab, stage, gl_acc);
    }
    gf_finish(tab, a.J, gl_acc, loss, 0.0, grad, loss_out, nullptr, red);
}

// ---- boundary morph --------------------------------------------------------------------------------------
// grid_sample(scores, (2x/W-1, 2y/H-1)) with align_corners=False, zero padding: pixel-centre coordinates x-0.5
__device__ __forceinline__ int morph_class(const double* __restrict__ scores, int C, int H, int W, double x, double y) {
    const double gx = ((x / W * 2.0 - 1.0) + 1.0) * W * 0.5 - 0.5, gy = ((y / H * 2.0 - 1.0) + 1.0) * H * 0.5 - 0.5;
    const double x0 = floor(gx), y0 = floor(gy);
    const int ix = (int)x0, iy = (int)y0;
    const double tx = gx - x0, ty = gy - y0;
    int best = 0;
    double bv = -INFINITY;
    for (int c = 0; c < C; ++c) {
        const double* sc = scores + (size_t)c * H * W;
        auto at = [&](int yy, int xx) { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? sc[yy * W + xx] : 0.0; };
        const double val = at(iy, ix) * (1.0 - tx) * (1.0 - ty) + at(iy, ix + 1) * tx * (1.0 - ty) +
                           at(iy + 1, ix) * (1.0 - tx) * ty + at(iy + 1, ix + 1) * tx * ty;
        if (val > bv) { bv = val; best = c; }
    }
    return best;
}

__global__ void __launch_bounds__(GF_BLOCK)
gf_morph_kernel(GfArgs a, const double* __restrict__ scores, int C, const int* __restrict__ sf_seg,
                const double* __restrict__ edge_pts, const int* __restrict__ edge_off, double* __restrict__ grad_m,
                double* __restrict__ acc /* [sum m, count] */) {
    extern __shared__ double tab[];
    __shared__ double red[GF_BLOCK / 32];
    for (int e = threadIdx.x; e < 7 * a.J; e += GF_BLOCK) tab[e] = 0.0;
    __syncthreads();
    const int n = n_active(a.n_cap, a.n_dev);
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    const int H = a.cam.H, W = a.cam.W;
    double gl_acc[7] = {0, 0, 0, 0, 0, 0, 0}, loss = 0.0, cnt = 0.0;
    double* stage = tab + 7 * a.J + (threadIdx.x >> 5) * GF_STAGE;
    for (int ib = i0; ib < i1; ib += GF_BLOCK) {
        const int i = ib + threadIdx.x;
        Warped s;
        V3 g = v3(0, 0, 0);
        bool live = false;
        do {
        if (i >= i1 || !gf_warp(a, i, s)) break;
        double x, y; V3 dx_, dy_;
        gf_project(s.pp, a.cam, x, y, dx_, dy_);
        const double sx = x / W * 2.0 - 1.0, sy = y / H * 2.0 - 1.0;
        if (!(sx > -1.0 && sx < 1.0 && sy > -1.0 && sy < 1.0)) break;
        const int cls = sf_seg[i];
        if (cls < 0 || cls >= C) break;
        if (morph_class(scores, C, H, W, x, y) == cls) break;
        const int e0 = edge_off[cls], e1 = edge_off[cls + 1];
        if (e1 - e0 < 2) break;
        // two nearest edge pixels of the surfel's own class (ties -> lower index)
        double b0 = INFINITY, b1 = INFINITY;
        int j0 = -1, j1 = -1;
        for (int e = e0; e < e1; ++e) {
            const double ex = edge_pts[2 * e] - x, ey = edge_pts[2 * e + 1] - y;
            const double d2 = ex * ex + ey * ey;
            if (d2 < b1) {
                if (d2 < b0) { b1 = b0; j1 = j0; b0 = d2; j0 = e; }
                else { b1 = d2; j1 = e; }
            }
        }
        // dropped when an edge pixel is farther than the image border
        const double d2e = fmin(fmin(fmin(x, y), (double)W - x), (double)H - y);
        if (sqrt(b1) > d2e || sqrt(b0) > d2e) break;
        const double m = 0.5 * (b0 + b1);
        if (!(m > 15.0)) break;
        loss += m;
        cnt += 1.0;
        const double gx = 2.0 * x - (edge_pts[2 * j0] + edge_pts[2 * j1]), gy = 2.0 * y - (edge_pts[2 * j0 + 1] + edge_pts[2 * j1 + 1]);
        g = v3(gx * dx_.x + gy * dy_.x, gx * dx_.y + gy * dy_.y, gx * dx_.z + gy * dy_.z);
        live = true;
        } while (false);
        gf_backprop_warp(a, s, g, live, tab, stage, gl_acc);
    }
    gf_finish(tab, a.J, gl_acc, loss, cnt, grad_m, acc, acc + 1, red);
}

// ---- regularisers on the graph: ARAP, Rot, Face (J, F small: one CTA) ----------------------------------------
__global__ void __launch_bounds__(GF_BLOCK)
gf_reg_kernel(const double* __restrict__ ed_points, const int* __restrict__ ed_knn, const double* __restrict__ ed_knn_w,
              int J, const int* __restrict__ tri, const double* __restrict__ areas, int F, const double* __restrict__ dv,
              double w_arap, double w_rot, double w_face, int use_arap, int use_rot, int use_face,
              double* __restrict__ grad, double* __restrict__ losses /* arap, rot, face */) {
    __shared__ double red[GF_BLOCK / 32];
    double la = 0.0, lr = 0.0, lf = 0.0;
    const double* gl = dv + 7 * J;
    const double gqw = gl[0];
    const V3 gqv = v3(gl[1], gl[2], gl[3]);
    if (use_arap) {
        for (int e = threadIdx.x; e < J * SB_KNN; e += GF_BLOCK) {
            const int j = e / SB_KNN, n_ = ed_knn[e];
            const double wjk = ed_knn_w[e];
            const V3 d = v3(ed_points[3 * j] - ed_points[3 * n_], ed_points[3 * j + 1] - ed_points[3 * n_ + 1],
                            ed_points[3 * j + 2] - ed_points[3 * n_ + 2]);
            const double* bn = dv + 7 * n_;
            const double* bj = dv + 7 * j;
            const V3 qv = v3(bn[1], bn[2], bn[3]);
            const V3 tv = qrot(d, bn[0], qv);
            // the rest vector goes through float32 in the reference (loss.py:468)
            const V3 r = v3(tv.x + bn[4] - (double)(float)d.x - bj[4], tv.y + bn[5] - (double)(float)d.y - bj[5],
                            tv.z + bn[6] - (double)(float)d.z - bj[6]);
            la += w_arap * wjk * dot3(r, r);
            const V3 g = scl3(2.0 * w_arap * wjk, r);
            double kw; V3 kv;
            qgrad(d, bn[0], qv, g, kw, kv);
            double* tn = grad + 7 * n_;
            double* tj = grad + 7 * j;
            atomicAdd(tn + 0, kw); atomicAdd(tn + 1, kv.x); atomicAdd(tn + 2, kv.y); atomicAdd(tn + 3, kv.z);
            atomicAdd(tn + 4, g.x); atomicAdd(tn + 5, g.y); atomicAdd(tn + 6, g.z);
            atomicAdd(tj + 4, -g.x); atomicAdd(tj + 5, -g.y); atomicAdd(tj + 6, -g.z);
        }
    }
    if (use_rot) {
        for (int j = threadIdx.x; j <= J; j += GF_BLOCK) {
            const double* q = dv + 7 * j;
            const double c = 1.0 - (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            lr += w_rot * c * c;
            for (int k = 0; k < 4; ++k) atomicAdd(grad + 7 * j + k, -4.0 * w_rot * c * q[k]);
        }
    }
    if (use_face) {
        double ga[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int f = threadIdx.x; f < F; f += GF_BLOCK) {
            int id[3];
            V3 base[3], vert[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                id[c] = tri[c * F + f];
                const double* b = dv + 7 * id[c];
                base[c] = v3(ed_points[3 * id[c]] + b[4], ed_points[3 * id[c] + 1] + b[5], ed_points[3 * id[c] + 2] + b[6]);
                const V3 r = qrot(base[c], gqw, gqv);
                vert[c] = v3(r.x + gl[4], r.y + gl[5], r.z + gl[6]);
            }
            const V3 e1 = sub3(vert[1], vert[0]), e2 = sub3(vert[2], vert[0]);
            const V3 cr = cross3(e1, e2);
            const double sn = sqrt(dot3(cr, cr) + 1e-13), A = 0.5 * sn;
            const double dA = A - areas[f];
            lf += w_face * dA * dA;
            const V3 gc = scl3(w_face * dA / sn, cr);                       // dL/dc = 2 w (A-A0) * 0.5 c / s
            const V3 g1 = cross3(e2, gc), g2 = cross3(gc, e1);             // dL/de1, dL/de2
            const V3 gv[3] = {v3(-g1.x - g2.x, -g1.y - g2.y, -g1.z - g2.z), g1, g2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double kw; V3 kv;
                qgrad(base[c], gqw, gqv, gv[c], kw, kv);
                ga[0] += kw; ga[1] += kv.x; ga[2] += kv.y; ga[3] += kv.z;
                ga[4] += gv[c].x; ga[5] += gv[c].y; ga[6] += gv[c].z;
                const V3 gb = qrot_t(gv[c], gqw, gqv);
                atomicAdd(grad + 7 * id[c] + 4, gb.x); atomicAdd(grad + 7 * id[c] + 5, gb.y); atomicAdd(grad + 7 * id[c] + 6, gb.z);
            }
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const double s = block_sum<GF_BLOCK>(ga[c], red);
            if (threadIdx.x == 0 && s != 0.0) atomicAdd(grad + 7 * J + c, s);
        }
    }
    const double sa = block_sum<GF_BLOCK>(la, red), sr = block_sum<GF_BLOCK>(lr, red), sf = block_sum<GF_BLOCK>(lf, red);
    if (threadIdx.x == 0) { losses[0] += sa; losses[1] += sr; losses[2] += sf; }
}

// ---- optimiser step (SGD momentum 0.9 | Adam), deform_mesh.py:272-275,325-327 ----------------------------------
// grad_total = grad + (w_morph / count) grad_morph;  grad_total[J] /= J;  step;  loss trace;  buffers re-zeroed.
// trace row (8 doubles): total, face, arap, rot, point_plane, bn_morph (NaN when no surfel qualifies), count, 0
__global__ void gf_step_kernel(double* __restrict__ dv, double* __restrict__ grad, double* __restrict__ grad_m,
                               double* __restrict__ acc /* [pp, arap, rot, face, morph_sum, morph_cnt] */, double w_morph,
                               int use_morph, int J, int optimizer, double lr, int iter, double* __restrict__ state,
                               double* __restrict__ trace, double* __restrict__ grad_out) {
    const int n = 7 * (J + 1);
    const double cnt = acc[5];
    const double ms = (use_morph && cnt > 0.0) ? w_morph / cnt : 0.0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        double g = grad[e] + (use_morph ? ms * grad_m[e] : 0.0);
        if (e >= 7 * J) g /= (double)J;
        if (grad_out) grad_out[e] = g;
        if (optimizer == 0) {                     // torch.optim.SGD(momentum=0.9): buf = g at the first step
            const double buf = (iter == 0) ? g : 0.9 * state[e] + g;
            state[e] = buf;
            dv[e] -= lr * buf;
        } else {                                  // torch.optim.Adam: betas (0.9, 0.999), eps 1e-8
            const double m = (iter == 0 ? 0.0 : 0.9 * state[e]) + 0.1 * g;
            const double v = (iter == 0 ? 0.0 : 0.999 * state[n + e]) + 0.001 * g * g;
            state[e] = m;
            state[n + e] = v;
            const double bc1 = 1.0 - pow(0.9, iter + 1), bc2 = 1.0 - pow(0.999, iter + 1);
            dv[e] -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + 1e-8);
        }
        grad[e] = 0.0;
        if (use_morph) grad_m[e] = 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double morph = use_morph ? (cnt > 0.0 ? w_morph * acc[4] / cnt : nan("")) : 0.0;
        double* t = trace + 8 * iter;
        t[1] = acc[3]; t[2] = acc[1]; t[3] = acc[2]; t[4] = acc[0]; t[5] = morph; t[6] = cnt; t[7] = 0.0;
        t[0] = acc[0] + acc[1] + acc[2] + acc[3] + (use_morph ? morph : 0.0);
    }
}
__global__ void gf_clear_acc_kernel(double* acc) { if (threadIdx.x < 6) acc[threadIdx.x] = 0.0; }

// Surfels.update with the global row (nodes.py:204-205,211-212,219-222): points += t_g (NOT rotated), normals rotated
// by q_g, nodes += t_g, node normals rotated by q_g -- applied AFTER the per-node warp (sb_warp_update).
__global__ void gf_global_update_kernel(double* __restrict__ points, double* __restrict__ norms, int n_cap, const int* n_dev,
                                        double* __restrict__ ed_points, double* __restrict__ ed_norms, int J,
                                        const double* __restrict__ gl) {
    const int n = n_active(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double qw = gl[0];
    const V3 qv = v3(gl[1], gl[2], gl[3]);
    if (i < n) {
        points[3 * (size_t)i] += gl[4]; points[3 * (size_t)i + 1] += gl[5]; points[3 * (size_t)i + 2] += gl[6];
        const V3 r = qrot(v3(norms[3 * (size_t)i], norms[3 * (size_t)i + 1], norms[3 * (size_t)i + 2]), qw, qv);
        const double nn = fmax(sqrt(dot3(r, r)), 1e-12);
        norms[3 * (size_t)i] = r.x / nn; norms[3 * (size_t)i + 1] = r.y / nn; norms[3 * (size_t)i + 2] = r.z / nn;
    }
    if (i < J) {
        ed_points[3 * i] += gl[4]; ed_points[3 * i + 1] += gl[5]; ed_points[3 * i + 2] += gl[6];
        const V3 r = qrot(v3(ed_norms[3 * i], ed_norms[3 * i + 1], ed_norms[3 * i + 2]), qw, qv);
        const double nn = fmax(sqrt(dot3(r, r)), 1e-12);
        ed_norms[3 * i] = r.x / nn; ed_norms[3 * i + 1] = r.y / nn; ed_norms[3 * i + 2] = r.z / nn;
    }
}

GfArgs make_gf(const double* points, const int* knn_idx, const double* knn_w, const unsigned char* stable, int n_cap,
               const int* n_dev, const double* ed_points, int J, const double* dv, int H, int W, const double* intr) {
    GfArgs a;
    a.points = points; a.knn_idx = knn_idx; a.knn_w = knn_w; a.stable = stable; a.n_cap = n_cap; a.n_dev = n_dev;
    a.ed_points = ed_points; a.J = J; a.dv = dv;
    a.cam.fx = intr[0]; a.cam.fy = intr[1]; a.cam.cx = intr[2]; a.cam.cy = intr[3]; a.cam.H = H; a.cam.W = W;
    return a;
}

int gf_grid(int n_cap) {
    int b = (n_cap + GF_BLOCK - 1) / GF_BLOCK;
    return b < 1 ? 1 : (b > 592 ? 592 : b);
}

template <typename Kern>
int gf_set_smem(Kern k, size_t smem) {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return SB_ERR_CUDA;
    return SB_OK;
}

}  // namespace

extern "C" {

int sb_gf_data(const double* points, const int* knn_idx, const double* knn_w, const unsigned char* stable, int n_cap,
               const int* n_dev, const double* ed_points, int J, const double* dv, const float* vmap, const float* nmap,
               int H, int W, const double* intr, double weight, int seg_mode, int C, const int* sf_seg,
               const double* sf_seg_conf, const double* trg_seg_conf, double* grad, double* acc, void* stream) {
    if (!points || !knn_idx || !knn_w || !ed_points || !dv || !vmap || !nmap || !grad || !acc) return SB_ERR_ARG;
    if (J <= 0 || ((size_t)7 * J + (GF_BLOCK / 32) * GF_STAGE) * sizeof(double) > 220 * 1024 || J > 65534) return SB_ERR_ARG;
    if (seg_mode && (C < 1 || C > GF_MAXC || !trg_seg_conf || (seg_mode == 1 && !sf_seg_conf) || (seg_mode == 2 && !sf_seg)))
        return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    GfArgs a = make_gf(points, knn_idx, knn_w, stable, n_cap, n_dev, ed_points, J, dv, H, W, intr);
    SegArgs sg;
    sg.mode = seg_mode; sg.C = C; sg.sf_seg = sf_seg; sg.sf_seg_conf = sf_seg_conf; sg.trg_seg_conf = trg_seg_conf;
    const size_t smem = ((size_t)7 * J + (GF_BLOCK / 32) * GF_STAGE) * sizeof(double);
    if (gf_set_smem(gf_data_kernel, smem) != SB_OK) return SB_ERR_CUDA;
    gf_data_kernel<<<gf_grid(n_cap), GF_BLOCK, smem, (cudaStream_t)stream>>>(
        a, reinterpret_cast<const float4*>(vmap), reinterpret_cast<const float4*>(nmap), sg, weight, grad, acc);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_gf_morph(const double* points, const int* knn_idx, const double* knn_w, const unsigned char* stable, int n_cap,
                const int* n_dev, const double* ed_points, int J, const double* dv, int H, int W, const double* intr,
                const double* scores, int C, const int* sf_seg, const double* edge_pts, const int* edge_off,
                double* grad_morph, double* acc, void* stream) {
    if (!points || !knn_idx || !knn_w || !ed_points || !dv || !scores || !sf_seg || !edge_pts || !edge_off ||
        !grad_morph || !acc)
        return SB_ERR_ARG;
    if (J <= 0 || ((size_t)7 * J + (GF_BLOCK / 32) * GF_STAGE) * sizeof(double) > 220 * 1024 || J > 65534 || C < 1 || C > GF_MAXC) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    GfArgs a = make_gf(points, knn_idx, knn_w, stable, n_cap, n_dev, ed_points, J, dv, H, W, intr);
    const size_t smem = ((size_t)7 * J + (GF_BLOCK / 32) * GF_STAGE) * sizeof(double);
    if (gf_set_smem(gf_morph_kernel, smem) != SB_OK) return SB_ERR_CUDA;
    gf_morph_kernel<<<gf_grid(n_cap), GF_BLOCK, smem, (cudaStream_t)stream>>>(a, scores, C, sf_seg, edge_pts, edge_off,
                                                                             grad_morph, acc + 4);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_gf_reg(const double* ed_points, const int* ed_knn, const double* ed_knn_w, int J, const int* triangles,
              const double* areas, int F, const double* dv, double w_arap, double w_rot, double w_face, int use_arap,
              int use_rot, int use_face, double* grad, double* acc, void* stream) {
    if (!ed_points || !dv || !grad || !acc || J <= 0) return SB_ERR_ARG;
    if (use_arap && (!ed_knn || !ed_knn_w)) return SB_ERR_ARG;
    if (use_face && F > 0 && (!triangles || !areas)) return SB_ERR_ARG;
    gf_reg_kernel<<<1, GF_BLOCK, 0, (cudaStream_t)stream>>>(ed_points, ed_knn, ed_knn_w, J, triangles, areas, F, dv,
                                                            w_arap, w_rot, w_face, use_arap, use_rot,
                                                            use_face && F > 0, grad, acc + 1);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_gf_step(double* dv, double* grad, double* grad_morph, double* acc, double w_morph, int use_morph, int J,
               int optimizer, double lr, int iter, double* state, double* trace, double* grad_out, void* stream) {
    if (!dv || !grad || !acc || !state || !trace || J <= 0 || iter < 0 || iter >= 64) return SB_ERR_ARG;
    if (optimizer != 0 && optimizer != 1) return SB_ERR_ARG;
    if (use_morph && !grad_morph) return SB_ERR_ARG;
    const int n = 7 * (J + 1);
    gf_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dv, grad, grad_morph, acc, w_morph, use_morph, J,
                                                                      optimizer, lr, iter, state, trace, grad_out);
    SB_CHECK_LAUNCH();
    gf_clear_acc_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_gf_global_update(double* points, double* norms, int n_cap, const int* n_dev, double* ed_points, double* ed_norms,
                        int J, const double* global_row, void* stream) {
    if (!points || !norms || !ed_points || !ed_norms || !global_row) return SB_ERR_ARG;
    const int m = n_cap > J ? n_cap : J;
    if (m <= 0) return SB_OK;
    gf_global_update_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, norms, n_cap, n_dev, ed_points,
                                                                             ed_norms, J, global_row);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
