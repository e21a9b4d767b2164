// Projective surfel fusion and stable-surfel compaction:
//   Surfels.fuseInputData                    /root/reference/super/nodes.py:270-541  (merge_data :301-355)
//   Surfels.prepareStableIndexNSwapAllModel  /root/reference/super/nodes.py:543-589  (state part)
//
// The reference builds up to 16 image-sized "layers" (the i-th most confident surfel of every pixel)
// with two global sorts and ~3 600 ATen launches, then merges layer by layer.  Every decision it
// takes is local to ONE pixel: a surfel projects to exactly one pixel, new data lives at one pixel.
// B200 design: bucket surfels by pixel (count -> exclusive scan -> fill: a CSR over pixels), then
// one thread per pixel orders its few surfels by (confidence desc, index asc) and replays the
// reference's layered merge logic sequentially for that pixel.  New surfels are appended in pixel
// order (scan), exactly the reference's order.  ~10 launches per frame, no host sync.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "super_b200.h"

namespace {

constexpr int MAX_LAYERS = 16;   // nodes.py:379

struct FuseWs {
    int* cnt;        // (P)   surfels per pixel
    int* offsets;    // (P+1) exclusive scan of cnt
    int* seg;        // (cap) surfel ids grouped by pixel
    int* pix;        // (cap) pixel of each live surfel or -1
    int* add_flag;   // (P)   1: new pixel passes the node-radius test and is appended
    int* add_pos;    // (P+1) exclusive scan of add_flag
    int* add_idx;    // (P,4)
    double* add_dist;  // (P,4)
    void* cub_tmp;
    size_t cub_bytes;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t cub_scan_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int*)nullptr, (int*)nullptr, n);
    return bytes;
}

size_t carve(FuseWs* w, char* base, int P, int cap) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
    int* cnt = (int*)take(sizeof(int) * (size_t)(P + 1));
    int* offsets = (int*)take(sizeof(int) * (size_t)(P + 1));
    int* seg = (int*)take(sizeof(int) * (size_t)cap);
    int* pix = (int*)take(sizeof(int) * (size_t)cap);
    int* add_flag = (int*)take(sizeof(int) * (size_t)(P + 1));
    int* add_pos = (int*)take(sizeof(int) * (size_t)(P + 1));
    int* add_idx = (int*)take(sizeof(int) * 4 * (size_t)P);
    double* add_dist = (double*)take(sizeof(double) * 4 * (size_t)P);
    size_t cb = cub_scan_bytes((P > cap ? P : cap) + 1);
    void* tmp = take(cb);
    if (w) {
        w->cnt = cnt; w->offsets = offsets; w->seg = seg; w->pix = pix; w->add_flag = add_flag;
        w->add_pos = add_pos; w->add_idx = add_idx; w->add_dist = add_dist; w->cub_tmp = tmp; w->cub_bytes = cb;
    }
    return off;
}

__device__ __forceinline__ Cam cam_of(const SbFrame& f) {
    Cam c; c.fx = f.fx; c.fy = f.fy; c.cx = f.cx; c.cy = f.cy; c.H = f.H; c.W = f.W;
    return c;
}

// ---- 1. project + count ------------------------------------------------------------------------
__global__ void project_count_kernel(SbSurfels sf, SbFrame fr, int* __restrict__ pix, int* __restrict__ cnt) {
    const int n = n_active(sf.cap, sf.n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cam cam = cam_of(fr);
    double u, v;
    project_ref(v3(sf.points[3 * (size_t)i], sf.points[3 * (size_t)i + 1], sf.points[3 * (size_t)i + 2]), cam, u, v);
    int p = -1;
    if (sf.stable[i] && fabs(u) < 1e9 && fabs(v) < 1e9) {
        const long long ur = round_ll(u), vr = round_ll(v);
        if (vr >= 0 && vr < cam.H - 1 && ur >= 0 && ur < cam.W - 1) {   // pcd2depth valid_proj, margin 0
            p = (int)(vr * cam.W + ur);
            atomicAdd(cnt + p, 1);
        }
    }
    pix[i] = p;
}

// ---- 2. fill the per-pixel segments --------------------------------------------------------------
__global__ void fill_kernel(int n_cap, const int* n_dev, const int* __restrict__ pix, int* __restrict__ cnt,
                            const int* __restrict__ offsets, int* __restrict__ seg) {
    const int n = n_active(n_cap, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pix[i];
    if (p < 0) return;
    const int slot = atomicSub(cnt + p, 1) - 1;   // any order: the pixel thread sorts its segment
    seg[offsets[p] + slot] = i;
}

// ---- 3. per-pixel merge logic ----------------------------------------------------------------------
struct Sample {
    double p[3], n[3], r;
    float c[3], w;
    int cls;                 // class (semantic state) or -1
    const double* conf;      // row of class probabilities or nullptr
};

__device__ __forceinline__ Sample load_surfel(const SbSurfels& sf, int i) {
    Sample s;
    for (int k = 0; k < 3; ++k) {
        s.p[k] = sf.points[3 * (size_t)i + k];
        s.n[k] = sf.norms[3 * (size_t)i + k];
        s.c[k] = sf.colors[3 * (size_t)i + k];
    }
    s.r = sf.radii[i];
    s.w = sf.confs[i];
    s.cls = sf.seg ? sf.seg[i] : -1;
    s.conf = sf.seg_conf ? sf.seg_conf + (size_t)i * sf.n_classes : nullptr;
    return s;
}

__device__ __forceinline__ Sample load_new(const SbFrame& fr, int p) {
    Sample s;
    const float4 pv = reinterpret_cast<const float4*>(fr.vmap)[p];
    const float4 nv = reinterpret_cast<const float4*>(fr.nmap)[p];
    s.p[0] = pv.x; s.p[1] = pv.y; s.p[2] = pv.z;
    s.n[0] = nv.x; s.n[1] = nv.y; s.n[2] = nv.z;
    const size_t P = (size_t)fr.H * fr.W;
    for (int k = 0; k < 3; ++k) s.c[k] = fr.color[k * P + p];
    s.r = fr.radii[p];
    s.w = fr.confs[p];
    s.cls = fr.seg ? fr.seg[p] : -1;
    s.conf = nullptr;        // set by the caller (needs the class count)
    return s;
}

// merge_data (nodes.py:301-355): test, then confidence-weighted blend written into surfel `dst`.
// float32 quantities (confidence, colour, blend weights) are float here too.
__device__ __forceinline__ bool try_merge(const SbSurfels& sf, int dst, const Sample& b, double th_dist,
                                          double th_cos, float time_now, bool add_new, bool class_gate) {
    const Sample a = load_surfel(sf, dst);
    if (class_gate && sf.seg && b.cls >= 0 && a.cls != b.cls) return false;            // nodes.py:314-316
    const double dx = a.p[0] - b.p[0], dy = a.p[1] - b.p[1], dz = a.p[2] - b.p[2];
    const double dist = sqrt(dx * dx + dy * dy + dz * dz);
    const double cosang = (a.n[0] * b.n[0] + a.n[1] * b.n[1]) + a.n[2] * b.n[2];
    if (!(dist < th_dist && cosang > th_cos)) return false;
    const float ws = __fadd_rn(a.w, b.w);
    const float wa = __fdiv_rn(a.w, ws), wb = __fdiv_rn(b.w, ws);
    double nn[3], nrm = 0.0;
    for (int k = 0; k < 3; ++k) {
        sf.points[3 * (size_t)dst + k] = (double)wa * a.p[k] + (double)wb * b.p[k];
        nn[k] = (double)wa * a.n[k] + (double)wb * b.n[k];
        nrm += nn[k] * nn[k];
    }
    nrm = fmax(sqrt(nrm), 1e-12);
    for (int k = 0; k < 3; ++k) sf.norms[3 * (size_t)dst + k] = nn[k] / nrm;
    sf.radii[dst] = (double)wa * a.r + (double)wb * b.r;
    sf.confs[dst] = ws;
    if (add_new) {   // the new sample's colour weight is tripled (nodes.py:337-341)
        const float wn = __fmul_rn(wb, 3.f);
        const float wsum = __fadd_rn(wa, wn);
        const float fa = __fdiv_rn(wa, wsum), fb = __fdiv_rn(wn, wsum);
        for (int k = 0; k < 3; ++k)
            sf.colors[3 * (size_t)dst + k] = __fadd_rn(__fmul_rn(fa, a.c[k]), __fmul_rn(fb, b.c[k]));
    } else {
        for (int k = 0; k < 3; ++k)
            sf.colors[3 * (size_t)dst + k] = __fadd_rn(__fmul_rn(wa, a.c[k]), __fmul_rn(wb, b.c[k]));
    }
    sf.time_stamp[dst] = time_now;
    if (sf.seg_conf && b.conf) {   // class probabilities: confidence-weighted, renormalised; class = argmax (nodes.py:345-353)
        double sc[8], sum = 0.0;
        const int C = sf.n_classes;
        for (int q = 0; q < C; ++q) { sc[q] = (double)wa * a.conf[q] + (double)wb * b.conf[q]; sum += sc[q]; }
        int am = 0;
        for (int q = 0; q < C; ++q) {
            sc[q] /= sum;
            sf.seg_conf[(size_t)dst * C + q] = sc[q];
            if (sc[q] > sc[am]) am = q;
        }
        sf.seg[dst] = am;
    }
    return true;
}

__global__ void fuse_pixels_kernel(SbSurfels sf, SbFrame fr, SbFuseParams pr, const int* __restrict__ offsets,
                                   int* __restrict__ seg, int* __restrict__ add_flag, long long* track_id,
                                   int n_track) {
    const int P = fr.H * fr.W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int off = offsets[p];
    const int m_all = offsets[p + 1] - off;
    const bool new_valid = reinterpret_cast<const float4*>(fr.vmap)[p].w != 0.f;
    int want_add = 0;
    if (m_all == 0) {
        // no layer covers this pixel: add_valid = valid & ~val_maps[0]  (nodes.py:412); the reference
        // only reaches that line when at least one layer exists anywhere (checked by the caller: n > 0)
        if (new_valid && !pr.disable_merging_new) want_add = 1;
        add_flag[p] = want_add;
        return;
    }
    // order the segment by (confidence desc, surfel index asc): the reference's two sorts with the
    // tie rule "lower index first" (nodes.py:367-371, DESIGN.md ties)
    for (int a = 1; a < m_all; ++a) {
        const int id = seg[off + a];
        const float c = sf.confs[id];
        int b = a - 1;
        while (b >= 0) {
            const int id2 = seg[off + b];
            const float c2 = sf.confs[id2];
            if (c2 > c || (c2 == c && id2 < id)) break;
            seg[off + b + 1] = id2;
            --b;
        }
        seg[off + b + 1] = id;
    }
    const int m = min(m_all, MAX_LAYERS);
    int ids[MAX_LAYERS];
    for (int l = 0; l < m; ++l) ids[l] = seg[off + l];

    // ---- merge the new sample into the first layer that accepts it (nodes.py:409-422)
    if (new_valid && !pr.disable_merging_new) {
        Sample nw = load_new(fr, p);
        if (fr.seg_conf) nw.conf = fr.seg_conf + (size_t)p * sf.n_classes;
        bool merged = false;
        for (int l = 0; l < m && !merged; ++l)
            merged = try_merge(sf, ids[l], nw, pr.th_dist, pr.th_cos, pr.time_now, true, pr.class_gate != 0);
        want_add = merged ? 0 : 1;
    }
    add_flag[p] = want_add;

    // ---- merge existing surfels that share the pixel, layer i <- layer j (nodes.py:424-460)
    if (!pr.disable_merging_exist) {
        bool present[MAX_LAYERS];
        for (int l = 0; l < m; ++l) present[l] = true;
        for (int i = 0; i < m; ++i) {
            bool cur = present[i];
            for (int j = i + 1; j < m && cur; ++j) {
                cur = cur && present[j];       // cumulative mask: val_map &= val_maps[j]
                if (!cur) break;
                const Sample sj = load_surfel(sf, ids[j]);
                if (try_merge(sf, ids[i], sj, pr.th_dist, pr.th_cos, pr.time_now, false, pr.class_gate != 0)) {
                    present[j] = false;
                    sf.stable[ids[j]] = 0;
                    for (int k = 0; k < n_track; ++k)
                        if (track_id[k] == ids[j]) track_id[k] = ids[i];
                }
            }
        }
        for (int l = MAX_LAYERS; l < m_all; ++l) {   // beyond 16 per pixel: deleted (nodes.py:402-403,460)
            const int id = seg[off + l];
            sf.stable[id] = 0;
            for (int k = 0; k < n_track; ++k)
                if (track_id[k] == id) track_id[k] = -2;
        }
    }
}

// ---- 4. new surfels: kNN against the nodes + radius test -------------------------------------------
constexpr int ADD_BLOCK = 128;
constexpr int ADD_TILE = 512;

__global__ void __launch_bounds__(ADD_BLOCK)
add_knn_kernel(SbFrame fr, const double* __restrict__ ed_points, const double* __restrict__ ed_radii, int J,
               int* __restrict__ add_flag, int* __restrict__ add_idx, double* __restrict__ add_dist,
               const int* __restrict__ ed_seg) {
    __shared__ double tile[ADD_TILE * 3];
    __shared__ float box[(ADD_TILE / 8) * 6];          // per group of 8 consecutive nodes: lo xyz, hi xyz (outward-rounded floats)
    const int P = fr.H * fr.W;
    const int p = blockIdx.x * ADD_BLOCK + threadIdx.x;
    const bool active = p < P && add_flag[p] != 0;
    if (!__syncthreads_or(active)) return;
    double bd[SB_KNN];
    int bi[SB_KNN];
    for (int k = 0; k < SB_KNN; ++k) { bd[k] = INFINITY; bi[k] = -1; }
    double qx = 0, qy = 0, qz = 0;
    const bool by_class = ed_seg != nullptr && fr.seg != nullptr;
    int qc = 0;
    if (active) {
        const float4 pv = reinterpret_cast<const float4*>(fr.vmap)[p];
        qx = pv.x; qy = pv.y; qz = pv.z;
        if (by_class) qc = fr.seg[p];
    }
    // (d2, index) lexicographic insertion: the result is the K smallest by (distance, index) whatever the visiting order,
    // i.e. ties -> lower index, exactly what the index-order scan with a strict comparison gave
    auto visit = [&](int j, int jg) {
        const double dx = subr(qx, tile[3 * j]), dy = subr(qy, tile[3 * j + 1]), dz = subr(qz, tile[3 * j + 2]);
        const double d2 = addr(addr(mulr(dx, dx), mulr(dy, dy)), mulr(dz, dz));
        if (d2 < bd[SB_KNN - 1] || (d2 == bd[SB_KNN - 1] && jg < bi[SB_KNN - 1])) {
            bd[SB_KNN - 1] = d2; bi[SB_KNN - 1] = jg;
#pragma unroll
            for (int k = SB_KNN - 1; k > 0; --k)
                if (bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bi[k] < bi[k - 1])) {
                    const double td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                    const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                }
        }
    };
    for (int t0 = 0; t0 < J; t0 += ADD_TILE) {
        const int cnt = min(ADD_TILE, J - t0);
        const int ng = (cnt + 7) >> 3;
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += ADD_BLOCK) tile[e] = ed_points[(size_t)t0 * 3 + e];
        __syncthreads();
        // Bounding boxes of groups of 8 consecutive nodes (neighbours on the node grid).  A group whose box is farther from
        // the pixel than its current K-th neighbour cannot contribute: the brute-force scan over all J nodes (the kernel was
        // FP64-issue-bound on it) becomes ~1/4 of the distance evaluations, with the same result.
        for (int g = threadIdx.x; g < ng; g += ADD_BLOCK) {
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int j = 8 * g; j < min(cnt, 8 * g + 8); ++j)
                for (int c = 0; c < 3; ++c) {
                    lo[c] = fminf(lo[c], __double2float_rd(tile[3 * j + c]));
                    hi[c] = fmaxf(hi[c], __double2float_ru(tile[3 * j + c]));
                }
            for (int c = 0; c < 3; ++c) { box[6 * g + c] = lo[c]; box[6 * g + 3 + c] = hi[c]; }
        }
        __syncthreads();
        if (active) {
            // lower bound of the squared distance to a group's box, rounded DOWN (float, then a relative margin)
            auto lower = [&](int g) {
                const float fx = (float)qx, fy = (float)qy, fz = (float)qz;
                const float ex = fmaxf(0.f, fmaxf(box[6 * g] - fx, fx - box[6 * g + 3]));
                const float ey = fmaxf(0.f, fmaxf(box[6 * g + 1] - fy, fy - box[6 * g + 4]));
                const float ez = fmaxf(0.f, fmaxf(box[6 * g + 2] - fz, fz - box[6 * g + 5]));
                return (double)((ex * ex + ey * ey + ez * ez) * 0.9999f);
            };
            int g0 = 0;                            // the group that most likely holds the nearest node goes first
            float best = INFINITY;
            for (int g = 0; g < ng; ++g) {
                const float lb = (float)lower(g);
                if (lb < best) { best = lb; g0 = g; }
            }
            for (int j = 8 * g0; j < min(cnt, 8 * g0 + 8); ++j)
                if (!(by_class && ed_seg[t0 + j] != qc)) visit(j, t0 + j);
            for (int g = 0; g < ng; ++g) {
                if (g == g0 || lower(g) > bd[SB_KNN - 1]) continue;
                for (int j = 8 * g; j < min(cnt, 8 * g + 8); ++j)
                    if (!(by_class && ed_seg[t0 + j] != qc)) visit(j, t0 + j);
            }
        }
    }
    if (!active) return;
    // --hard_seg: a class with fewer than K nodes.  The reference asserts (utils/utils.py:236); padding the tuple with
    // an unrelated node would bind the surfel to it with a weight of 0.15-0.25 and put duplicate ids into the node-set
    // key, so such a pixel is simply not added.
    if (bi[SB_KNN - 1] < 0) { add_flag[p] = 0; return; }
    bool any = false;
    for (int k = 0; k < SB_KNN; ++k) {
        bd[k] = __dsqrt_rn(bd[k]);
        any |= bd[k] <= ed_radii[bi[k]];       // nodes.py:501-502
        add_idx[4 * (size_t)p + k] = bi[k];
        add_dist[4 * (size_t)p + k] = bd[k];
    }
    if (!any) add_flag[p] = 0;
}

__global__ void add_write_kernel(SbSurfels sf, SbFrame fr, SbFuseParams pr, const double* __restrict__ ed_radii,
                                 const int* __restrict__ add_flag, const int* __restrict__ add_pos,
                                 const int* __restrict__ add_idx, const double* __restrict__ add_dist,
                                 int* __restrict__ n_out, int* __restrict__ overflow) {
    const int P = fr.H * fr.W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_old = n_active(sf.cap, sf.n_dev);
    if (p == 0) {
        const int total = n_old + add_pos[P];
        *n_out = min(total, sf.cap);
        if (total > sf.cap) *overflow = 1;
    }
    if (p >= P || !add_flag[p]) return;
    const int dst = n_old + add_pos[p];
    if (dst >= sf.cap) return;
    const Sample s = load_new(fr, p);
    double e[SB_KNN], mx = -INFINITY, sum = 0.0;
    const int C = sf.n_classes;
    const double* qc = fr.seg_conf ? fr.seg_conf + (size_t)p * C : nullptr;
    for (int k = 0; k < SB_KNN; ++k) {
        const int n_ = add_idx[4 * (size_t)p + k];
        sf.knn_idx[4 * (size_t)dst + k] = n_;
        e[k] = exp(-add_dist[4 * (size_t)p + k] / ed_radii[n_]);
        if ((pr.semantic_weights & 2) && qc) {                           // nodes.py:503-509 (not under --hard_seg)
            const double* P = pr.ed_seg_conf + (size_t)n_ * C;
            double k1 = 0.0, k2 = 0.0;
            for (int q = 0; q < C; ++q) {
                const double m = 0.5 * (P[q] + qc[q]);
                k1 += P[q] * log(P[q] / (m + 1e-13) + 1e-13);
                k2 += qc[q] * log(qc[q] / (m + 1e-13) + 1e-13);
            }
            e[k] = sqrt(exp(-0.5 * (k1 + k2))) * sqrt(e[k]);
        }
        mx = fmax(mx, e[k]);
    }
    for (int k = 0; k < SB_KNN; ++k) { e[k] = exp(e[k] - mx); sum += e[k]; }
    for (int k = 0; k < SB_KNN; ++k) sf.knn_w[4 * (size_t)dst + k] = e[k] / sum;
    for (int k = 0; k < 3; ++k) {
        sf.points[3 * (size_t)dst + k] = s.p[k];
        sf.norms[3 * (size_t)dst + k] = s.n[k];
        sf.colors[3 * (size_t)dst + k] = s.c[k];
    }
    sf.radii[dst] = s.r;
    sf.confs[dst] = s.w;
    sf.time_stamp[dst] = pr.time_now;
    sf.stable[dst] = 1;
    if (sf.seg && fr.seg) sf.seg[dst] = fr.seg[p];
    if (sf.seg_conf && qc)
        for (int q = 0; q < C; ++q) sf.seg_conf[(size_t)dst * C + q] = qc[q];
}

// n_out = n when nothing is added
__global__ void copy_count_kernel(const int* n_dev, int cap, int* n_out) { *n_out = n_active(cap, n_dev); }

// ---- compaction --------------------------------------------------------------------------------------
__global__ void keep_flag_kernel(SbSurfels sf, float time_now, float th_steps, int disable_removing,
                                 const long long* track_id, int n_track, int* __restrict__ keep) {
    const int n = n_active(sf.cap, sf.n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { keep[i] = 0; return; }
    bool k = true;
    if (!disable_removing) {
        k = sf.stable[i] && (time_now - sf.time_stamp[i] < th_steps);     // nodes.py:557
        for (int t = 0; t < n_track; ++t) k = k || (track_id[t] == i);      // nodes.py:559
    }
    keep[i] = k ? 1 : 0;
}

__global__ void compact_scatter_kernel(SbSurfels src, SbSurfels dst, SbFrame fr, const int* __restrict__ keep,
                                       const int* __restrict__ pos, int disable_removing) {
    const int n = n_active(src.cap, src.n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *dst.n_dev = pos[n];
    if (i >= n || !keep[i]) return;
    const int d = pos[i];
    double pt[3];
    for (int k = 0; k < 3; ++k) {
        pt[k] = src.points[3 * (size_t)i + k];
        dst.points[3 * (size_t)d + k] = pt[k];
        dst.norms[3 * (size_t)d + k] = src.norms[3 * (size_t)i + k];
        dst.colors[3 * (size_t)d + k] = src.colors[3 * (size_t)i + k];
    }
    dst.confs[d] = src.confs[i];
    dst.radii[d] = src.radii[i];
    dst.time_stamp[d] = src.time_stamp[i];
    reinterpret_cast<int4*>(dst.knn_idx)[d] = reinterpret_cast<const int4*>(src.knn_idx)[i];
    for (int k = 0; k < 4; ++k) dst.knn_w[4 * (size_t)d + k] = src.knn_w[4 * (size_t)i + k];
    // projdata = unrounded (u,v) as float32 (nodes.py:540-541), computed from the fused position
    double u, v;
    project_ref(v3(pt[0], pt[1], pt[2]), cam_of(fr), u, v);
    dst.projdata[2 * (size_t)d] = (float)u;
    dst.projdata[2 * (size_t)d + 1] = (float)v;
    dst.stable[d] = disable_removing ? src.stable[i] : 1;
    if (src.seg && dst.seg) dst.seg[d] = src.seg[i];
    if (src.seg_conf && dst.seg_conf)
        for (int q = 0; q < src.n_classes; ++q) dst.seg_conf[(size_t)d * src.n_classes + q] = src.seg_conf[(size_t)i * src.n_classes + q];
}

__global__ void track_remap_kernel(long long* track_id, int n_track, const int* keep, const int* pos, int disable) {
    const int t = threadIdx.x;
    if (t >= n_track || disable) return;
    const long long id = track_id[t];
    if (id >= 0) track_id[t] = keep[id] ? pos[id] : -1;     // nodes.py:576-580 (id_map is -1 for dropped rows)
}

// ---- tracked points -------------------------------------------------------------------------------------
// init_track_pts / update_track_pts as called from prepareStableIndexNSwapAllModel (nodes.py:225-265,594-599), one
// CTA: the labelled points are few (20) and the search is one pass over the surfels per unassigned point.
constexpr int TRK_BLOCK = 1024;
constexpr int TRK_MAX = 64;

struct TrackArgs {
    const double* points; const unsigned char* stable; const float* projdata; int n_cap; const int* n_dev;
    const float4* vmap; int H, W;
    const int* gt;            // (T,3) i32 [x, y, valid]
    int T;
    long long* track_id;      // (T,) i64: >= 0 surfel row, -1 not started, -2 lost
    float* out;               // (T,3) f32 track_rsts[filename]
};

__device__ void track_init(const TrackArgs& a, int n, int first_valid, double th, double* sd, int* si, long long* ids) {
    const int tid_ = threadIdx.x;
    for (int k = 0; k < a.T; ++k) {
        __syncthreads();
        if (tid_ < a.T) ids[tid_] = a.track_id[tid_];
        __syncthreads();
        const long long cur = ids[k];
        const int x = a.gt[3 * k], y = a.gt[3 * k + 1], v = a.gt[3 * k + 2];
        const int pix = y * a.W + x;
        const bool in_img = x >= 0 && x < a.W && y >= 0 && y < a.H;
        const float4 q = in_img ? a.vmap[pix] : make_float4(0, 0, 0, 0);
        // gt_id > 0: a valid pixel that is not the first valid pixel of the frame (index_map value 0), nodes.py:240
        if (cur < 0 && in_img && q.w != 0.f && pix != first_valid && v == 1) {
            bool any_inval = false, lost = false;
            for (int j = 0; j < a.T; ++j) { any_inval |= ids[j] >= 0 || ids[j] == -2; lost |= ids[j] == -2; }
            double best = INFINITY;
            int besti = 0x7fffffff;
            for (int i = tid_; i < n; i += TRK_BLOCK) {
                const double dx = a.points[3 * (size_t)i] - (double)q.x, dy = a.points[3 * (size_t)i + 1] - (double)q.y,
                             dz = a.points[3 * (size_t)i + 2] - (double)q.z;
                double d = sqrt(dx * dx + dy * dy + dz * dz);
                if (any_inval) {                                       // nodes.py:242-245
                    bool ex = !a.stable[i] || (lost && i == n - 2);    // dists[-2] = 1e13: the reference indexes with the id -2
                    for (int j = 0; j < a.T && !ex; ++j) ex = ids[j] == i;
                    if (ex) d = 1e13;
                }
                if (d < best) { best = d; besti = i; }                 // ascending i per thread: first minimum
            }
            sd[tid_] = best; si[tid_] = besti;
            __syncthreads();
            for (int o = TRK_BLOCK / 2; o > 0; o >>= 1) {
                if (tid_ < o) {
                    const double d2 = sd[tid_ + o];
                    const int i2 = si[tid_ + o];
                    if (d2 < sd[tid_] || (d2 == sd[tid_] && i2 < si[tid_])) { sd[tid_] = d2; si[tid_] = i2; }
                }
                __syncthreads();
            }
            if (tid_ == 0 && sd[0] < th) a.track_id[k] = si[0];
        }
    }
    __syncthreads();
    if (tid_ < a.T) {   // rows from projdata[track_id] (negative ids index from the end, as in the reference)
        long long id = a.track_id[tid_];
        if (id < 0) id += n;
        const bool okid = id >= 0 && id < n;
        a.out[3 * tid_] = okid ? a.projdata[2 * id] : 0.f;
        a.out[3 * tid_ + 1] = okid ? a.projdata[2 * id + 1] : 0.f;
        a.out[3 * tid_ + 2] = 1.f;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TRK_BLOCK) track_points_kernel(TrackArgs a) {
    __shared__ double sd[TRK_BLOCK];
    __shared__ int si[TRK_BLOCK];
    __shared__ long long ids[TRK_MAX];
    __shared__ int flags[2];
    const int n = n_active(a.n_cap, a.n_dev);
    // first valid pixel of the frame (index_map == 0)
    int fv = 0x7fffffff;
    for (int p = threadIdx.x; p < a.H * a.W; p += TRK_BLOCK)
        if (a.vmap[p].w != 0.f) { fv = p; break; }
    si[threadIdx.x] = fv;
    __syncthreads();
    for (int o = TRK_BLOCK / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) si[threadIdx.x] = min(si[threadIdx.x], si[threadIdx.x + o]);
        __syncthreads();
    }
    const int first_valid = si[0];
    __syncthreads();
    if (threadIdx.x < 3 * a.T) a.out[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) {
        bool any_tracked = false, any_new = false;
        for (int j = 0; j < a.T; ++j) { any_tracked |= a.track_id[j] >= 0; any_new |= a.track_id[j] == -1; }
        flags[0] = any_tracked; flags[1] = any_new;
    }
    __syncthreads();
    // update_track_pts: first record for this filename -> init_track_pts with th = 1e-2 (nodes.py:256-258)
    if (flags[0]) track_init(a, n, first_valid, 1e-2, sd, si, ids);
    __syncthreads();
    if (threadIdx.x == 0) {
        bool any_new = false;
        for (int j = 0; j < a.T; ++j) any_new |= a.track_id[j] == -1;
        flags[1] = any_new;
    }
    __syncthreads();
    if (flags[1]) {                                  // init_track_pts(th = 0.2) re-creates the record (nodes.py:233)
        if (threadIdx.x < 3 * a.T) a.out[threadIdx.x] = 0.f;
        track_init(a, n, first_valid, 0.2, sd, si, ids);
    }
}

}  // namespace

extern "C" {

int sb_track_points(const double* points, const unsigned char* stable, const float* projdata, int n_cap, const int* n_dev,
                    const float* vmap, int H, int W, const int* gt, int T, long long* track_id, float* out, void* stream) {
    if (!points || !stable || !projdata || !vmap || !gt || !track_id || !out || T < 1 || T > TRK_MAX) return SB_ERR_ARG;
    TrackArgs a;
    a.points = points; a.stable = stable; a.projdata = projdata; a.n_cap = n_cap; a.n_dev = n_dev;
    a.vmap = reinterpret_cast<const float4*>(vmap); a.H = H; a.W = W; a.gt = gt; a.T = T; a.track_id = track_id; a.out = out;
    track_points_kernel<<<1, TRK_BLOCK, 0, (cudaStream_t)stream>>>(a);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

long long sb_fuse_workspace_bytes(int H, int W, int cap) { return (long long)carve(nullptr, nullptr, H * W, cap); }

int sb_fuse(const SbSurfels* sfp, const SbFrame* frp, const double* ed_points, const double* ed_radii, int J,
            const SbFuseParams* prp, long long* track_id, int n_track, int* n_out, int* overflow, void* workspace,
            long long ws_bytes, void* stream) {
    if (!sfp || !frp || !ed_points || !ed_radii || !prp || !n_out || !overflow || !workspace) return SB_ERR_ARG;
    const SbSurfels sf = *sfp;
    const SbFrame fr = *frp;
    const SbFuseParams pr = *prp;
    const int P = fr.H * fr.W;
    FuseWs w;
    if ((long long)carve(&w, (char*)workspace, P, sf.cap) > ws_bytes) return SB_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int gs = (sf.cap + 255) / 256, gp = (P + 255) / 256;

    cudaMemsetAsync(w.cnt, 0, sizeof(int) * (size_t)(P + 1), s);
    project_count_kernel<<<gs, 256, 0, s>>>(sf, fr, w.pix, w.cnt);
    SB_CHECK_LAUNCH();
    size_t cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.cnt, w.offsets, P + 1, s);
    fill_kernel<<<gs, 256, 0, s>>>(sf.cap, sf.n_dev, w.pix, w.cnt, w.offsets, w.seg);
    SB_CHECK_LAUNCH();
    fuse_pixels_kernel<<<gp, 256, 0, s>>>(sf, fr, pr, w.offsets, w.seg, w.add_flag, track_id, track_id ? n_track : 0);
    SB_CHECK_LAUNCH();
    // weights of ALL existing surfels from their fused positions and old indices (nodes.py:480-484)
    int rc;
    if (pr.semantic_weights & 1) {                // all existing surfels: semantic-super, with or without --hard_seg
        if (!pr.ed_seg_conf || !sf.seg_conf) return SB_ERR_ARG;
        rc = sb_reweight_semantic(sf.points, sf.knn_idx, sf.cap, sf.n_dev, ed_points, ed_radii, pr.ed_seg_conf, sf.seg_conf,
                                  sf.n_classes, sf.knn_w, stream);
    } else {
        rc = sb_reweight(sf.points, sf.knn_idx, sf.cap, sf.n_dev, ed_points, ed_radii, sf.knn_w, stream);
    }
    if (rc) return rc;
    if (!pr.disable_adding_new && !pr.disable_merging_new) {
        add_knn_kernel<<<(P + ADD_BLOCK - 1) / ADD_BLOCK, ADD_BLOCK, 0, s>>>(fr, ed_points, ed_radii, J, w.add_flag,
                                                                             w.add_idx, w.add_dist, pr.ed_seg);
        SB_CHECK_LAUNCH();
        cudaMemsetAsync(w.add_flag + P, 0, sizeof(int), s);
        cb = w.cub_bytes;
        cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.add_flag, w.add_pos, P + 1, s);
        add_write_kernel<<<gp, 256, 0, s>>>(sf, fr, pr, ed_radii, w.add_flag, w.add_pos, w.add_idx, w.add_dist, n_out,
                                            overflow);
        SB_CHECK_LAUNCH();
    } else {
        copy_count_kernel<<<1, 1, 0, s>>>(sf.n_dev, sf.cap, n_out);
        SB_CHECK_LAUNCH();
    }
    return SB_OK;
}

int sb_compact(const SbSurfels* srcp, const SbSurfels* dstp, const SbFrame* frp, double time_now, int th_time_steps,
               int disable_removing, long long* track_id, int n_track, void* workspace, long long ws_bytes,
               void* stream) {
    if (!srcp || !dstp || !frp || !workspace) return SB_ERR_ARG;
    const SbSurfels src = *srcp, dst = *dstp;
    if (dst.cap < src.cap) return SB_ERR_ARG;
    FuseWs w;
    const int P = frp->H * frp->W;
    if ((long long)carve(&w, (char*)workspace, P, src.cap) > ws_bytes) return SB_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if ((size_t)src.cap + 1 > (size_t)4 * P) return SB_ERR_WORKSPACE;
    int* keep = w.add_idx;                  // (4P ints) >= cap+1, free after sb_fuse
    int* pos = (int*)w.add_dist;            // (8P ints) >= cap+1
    const int g = (src.cap + 1 + 255) / 256;
    keep_flag_kernel<<<g, 256, 0, s>>>(src, (float)time_now, (float)th_time_steps, disable_removing, track_id,
                                       track_id ? n_track : 0, keep);
    SB_CHECK_LAUNCH();
    size_t cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, keep, pos, src.cap + 1, s);
    compact_scatter_kernel<<<g, 256, 0, s>>>(src, dst, *frp, keep, pos, disable_removing);
    SB_CHECK_LAUNCH();
    if (track_id && n_track > 0) {
        track_remap_kernel<<<1, 64, 0, s>>>(track_id, n_track, keep, pos, disable_removing);
        SB_CHECK_LAUNCH();
    }
    return SB_OK;
}

}  // extern "C"
