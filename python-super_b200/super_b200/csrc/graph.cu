// ED-graph construction, once per sequence (and at every re-initialisation): init_graph + the grid_mesh branch of
// DirectDeformGraph.init_ED_nodes (/root/reference/super/graph_encoder.py:11-67,128-167) from the dense maps of the first
// frame, plus the solver's node order (no reference counterpart).  The reference does this with ~60 ATen launches, an
// O(J E) Python loop with one host sync per node for the radii (graph_encoder.py:151-154) and leaves index_map on the CPU.
//
// One CTA (the graph has a few hundred to a few thousand nodes): every step is a loop over the anchor grid with a block-wide
// scan where an order is needed, so node ids, edge order, face order and every sum are FIXED -- the graph, and with it the
// whole tracker, is bitwise reproducible (the round-1 host version summed the incident edge lengths with index_add_, i.e.
// float atomics).
//   nodes     : anchors (v, u) = (gy*step, gx*step) on valid pixels, ids in row-major order           (:16-24)
//   edges     : per anchor a with right / down / diagonal neighbours r, d, rd: (a,r) (a,rd) (a,d) (r,d) (:36-44,58-60)
//   triangles : (a,r,rd) (a,rd,d)                                                                      (:47-51,62-64)
//   --hard_seg + --mesh_face: edges / triangles across classes are dropped                             (:141-149)
//   radii     : mean length of the incident edges, in edge order; nodes without edges get the mean of the others (:151-154,164-166)
//   areas     : 0.5 sqrt(|(p1-p0) x (p2-p0)|^2 + 1e-13)                                                (:156-159)
#include "common.cuh"
#include "super_b200.h"

namespace {

constexpr int GT = 1024;

// exclusive scan of `cnt[0..n)` in place (values replaced by their prefix), returns the total; block-wide, any n
__device__ int block_exclusive_scan(int* cnt, int n, int* s_warp /* 33 ints */) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += GT) {
        const int i = base + tid;
        const int v = i < n ? cnt[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;                       // inclusive over warps
        }
        __syncthreads();
        const int before = carry + (wid > 0 ? s_warp[wid - 1] : 0) + x - v;
        if (i < n) cnt[i] = before;
        const int total = s_warp[31];
        __syncthreads();
        carry += total;
    }
    return carry;
}

struct GraphArgs {
    const float4* vmap; const float4* nmap; const double* seg_conf; int C;
    int H, W, step, gh, gw, prune;
    int* ws;                 // 3 G ints: node id per anchor | edge offsets | face offsets
    double* points; double* norms; int* anchor_uv; int* seg; double* ed_seg_conf;
    int* edges; int* faces; double* edge_lens; double* radii; double* areas; int* node_pos; int* counts;
};

__device__ __forceinline__ double len3(const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(fma(z, z, fma(y, y, x * x)));
}

__global__ void __launch_bounds__(GT) graph_build_kernel(GraphArgs g) {
    __shared__ int s_warp[33];
    __shared__ double s_red[GT / 32];
    __shared__ double s_sum;
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    const int G = g.gh * g.gw, W = g.W, s = g.step;
    int* nid = g.ws;
    int* eoff = g.ws + G;
    int* foff = g.ws + 2 * G;

    // 1. nodes: anchors on valid pixels, ids in row-major order
    for (int a = tid; a < G; a += GT) {
        const int gy = a / g.gw, gx = a - gy * g.gw;
        nid[a] = g.vmap[(size_t)(gy * s) * W + gx * s].w != 0.f ? 1 : 0;
    }
    __syncthreads();
    // the solver's node order first (it needs the raw flags): nodes sorted along the LONGER image axis, so that the block
    // half-bandwidth of J^T J is ~3 grid lines of the shorter one
    if (g.W >= g.H) {
        for (int a = tid; a < G; a += GT) {           // column-major copy of the flags
            const int gx = a / g.gh, gy = a - gx * g.gh;
            eoff[a] = nid[gy * g.gw + gx];
        }
        __syncthreads();
        block_exclusive_scan(eoff, G, s_warp);
    }
    __syncthreads();
    for (int a = tid; a < G; a += GT) foff[a] = nid[a];
    __syncthreads();
    const int J = block_exclusive_scan(nid, G, s_warp);
    __syncthreads();
    for (int a = tid; a < G; a += GT) {
        const bool is_node = foff[a] != 0;
        const int id = nid[a];
        if (is_node) {
            const int gy = a / g.gw, gx = a - gy * g.gw;
            const size_t pix = (size_t)(gy * s) * W + gx * s;
            const float4 v = g.vmap[pix], n = g.nmap[pix];
            g.points[3 * id] = v.x; g.points[3 * id + 1] = v.y; g.points[3 * id + 2] = v.z;
            g.norms[3 * id] = n.x; g.norms[3 * id + 1] = n.y; g.norms[3 * id + 2] = n.z;
            g.anchor_uv[2 * id] = gx * s; g.anchor_uv[2 * id + 1] = gy * s;
            g.node_pos[id] = g.W >= g.H ? eoff[gx * g.gh + gy] : id;
            if (g.seg_conf) {                           // graph_encoder.py:134-139: class = first maximum
                int am = 0;
                double mx = -INFINITY;
                for (int c = 0; c < g.C; ++c) {
                    const double p = g.seg_conf[pix * g.C + c];
                    g.ed_seg_conf[(size_t)id * g.C + c] = p;
                    if (p > mx) { mx = p; am = c; }
                }
                g.seg[id] = am;
            }
        }
    }
    __syncthreads();
    for (int a = tid; a < G; a += GT) nid[a] = foff[a] != 0 ? nid[a] : -1;
    __syncthreads();

    // 2. edges and triangles of every cell whose top-left anchor is a node
    auto node_at = [&](int gx, int gy) { return (gx < g.gw && gy < g.gh) ? nid[gy * g.gw + gx] : -1; };
    auto same = [&](int a, int b) { return !g.prune || g.seg[a] == g.seg[b]; };
    for (int a = tid; a < G; a += GT) {
        const int gy = a / g.gw, gx = a - gy * g.gw;
        const int na = nid[a];
        int ne = 0, nf = 0;
        if (na >= 0) {
            const int r = node_at(gx + 1, gy), d = node_at(gx, gy + 1), rd = node_at(gx + 1, gy + 1);
            ne = (r >= 0 && same(na, r)) + (rd >= 0 && same(na, rd)) + (d >= 0 && same(na, d)) +
                 (r >= 0 && d >= 0 && same(r, d));
            nf = (r >= 0 && rd >= 0 && same(na, r) && same(na, rd)) + (rd >= 0 && d >= 0 && same(na, rd) && same(na, d));
        }
        eoff[a] = ne;
        foff[a] = nf;
    }
    __syncthreads();
    const int E = block_exclusive_scan(eoff, G, s_warp);
    __syncthreads();
    const int F = block_exclusive_scan(foff, G, s_warp);
    __syncthreads();
    for (int a = tid; a < G; a += GT) {
        const int gy = a / g.gw, gx = a - gy * g.gw;
        const int na = nid[a];
        if (na < 0) continue;
        const int r = node_at(gx + 1, gy), d = node_at(gx, gy + 1), rd = node_at(gx + 1, gy + 1);
        int e = eoff[a], f = foff[a];
        auto put_e = [&](int p, int q) {
            g.edges[2 * e] = p; g.edges[2 * e + 1] = q;
            g.edge_lens[e] = len3(g.points + 3 * p, g.points + 3 * q);
            ++e;
        };
        auto put_f = [&](int p, int q, int t) {
            g.faces[3 * f] = p; g.faces[3 * f + 1] = q; g.faces[3 * f + 2] = t;
            const double* p0 = g.points + 3 * p; const double* p1 = g.points + 3 * q; const double* p2 = g.points + 3 * t;
            const V3 a1 = v3(p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]), a2 = v3(p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]);
            const V3 c = cross_ref(a1, a2);
            g.areas[f] = 0.5 * sqrt(((c.x * c.x + c.y * c.y) + c.z * c.z) + 1e-13);
            ++f;
        };
        if (r >= 0 && same(na, r)) put_e(na, r);
        if (rd >= 0 && same(na, rd)) put_e(na, rd);
        if (d >= 0 && same(na, d)) put_e(na, d);
        if (r >= 0 && d >= 0 && same(r, d)) put_e(r, d);
        if (r >= 0 && rd >= 0 && same(na, r) && same(na, rd)) put_f(na, r, rd);
        if (rd >= 0 && d >= 0 && same(na, rd) && same(na, d)) put_f(na, rd, d);
    }
    __syncthreads();

    // 3. radii: mean incident edge length, summed in edge order (one thread per node scans the edge list: O(J E / 1024))
    double my_sum = 0.0;
    int my_cnt = 0;
    for (int k = tid; k < J; k += GT) {
        double sum = 0.0;
        int cnt = 0;
        for (int e = 0; e < E; ++e)
            if (g.edges[2 * e] == k || g.edges[2 * e + 1] == k) { sum += g.edge_lens[e]; ++cnt; }
        const double r = cnt ? sum / (double)cnt : nan("");
        g.radii[k] = r;
        if (cnt) { my_sum += r; ++my_cnt; }
    }
    const double tot = block_sum<GT>(my_sum, s_red);
    const double totc = block_sum<GT>((double)my_cnt, s_red);
    if (tid == 0) { s_sum = tot; s_cnt = (int)totc; }
    __syncthreads();
    if (s_cnt < J) {
        const double mean = s_sum / (double)s_cnt;
        for (int k = tid; k < J; k += GT)
            if (isnan(g.radii[k])) g.radii[k] = mean;
    }
    if (tid == 0) { g.counts[0] = J; g.counts[1] = E; g.counts[2] = F; }
}

// block half-bandwidth the ED graph's own pairs (ARAP) need in the solver's order: max |pos[j] - pos[knn[j,k]]|
__global__ void graph_pair_span_kernel(const int* __restrict__ knn, const int* __restrict__ pos, int J, int K, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int span = 0;
    if (i < J * K) {
        const int j = i / K, n = knn[i];
        if (n >= 0) span = abs(pos[j] - pos[n]);
    }
    span = __reduce_max_sync(0xffffffffu, span);
    if ((threadIdx.x & 31) == 0 && span > 0) atomicMax(out, span);
}

}  // namespace

extern "C" {

int sb_graph_anchors(int H, int W, int step) {
    if (H < 2 || W < 2 || step < 1) return 0;
    return ((H - 1 + step - 1) / step) * ((W - 1 + step - 1) / step);
}

int sb_graph_build(const float* vmap, const float* nmap, const double* seg_conf, int C, int H, int W, int step,
                   int prune_classes, int* workspace, double* points, double* norms, int* anchor_uv, int* seg,
                   double* ed_seg_conf, int* edges, int* faces, double* edge_lens, double* radii, double* areas,
                   int* node_pos, int* counts, void* stream) {
    if (!vmap || !nmap || !workspace || !points || !norms || !anchor_uv || !edges || !faces || !edge_lens || !radii ||
        !areas || !node_pos || !counts)
        return SB_ERR_ARG;
    if (sb_graph_anchors(H, W, step) <= 0) return SB_ERR_ARG;
    if (seg_conf && (!seg || !ed_seg_conf || C < 1 || C > 8)) return SB_ERR_ARG;
    if (prune_classes && !seg_conf) return SB_ERR_ARG;
    GraphArgs g;
    g.vmap = reinterpret_cast<const float4*>(vmap); g.nmap = reinterpret_cast<const float4*>(nmap);
    g.seg_conf = seg_conf; g.C = C; g.H = H; g.W = W; g.step = step;
    g.gh = (H - 1 + step - 1) / step; g.gw = (W - 1 + step - 1) / step; g.prune = prune_classes;
    g.ws = workspace; g.points = points; g.norms = norms; g.anchor_uv = anchor_uv; g.seg = seg; g.ed_seg_conf = ed_seg_conf;
    g.edges = edges; g.faces = faces; g.edge_lens = edge_lens; g.radii = radii; g.areas = areas; g.node_pos = node_pos;
    g.counts = counts;
    graph_build_kernel<<<1, GT, 0, (cudaStream_t)stream>>>(g);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_graph_pair_span(const int* knn, const int* node_pos, int J, int K, int* out_max, void* stream) {
    if (!knn || !node_pos || !out_max || J <= 0 || K <= 0) return SB_ERR_ARG;
    graph_pair_span_kernel<<<(J * K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(knn, node_pos, J, K, out_max);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
