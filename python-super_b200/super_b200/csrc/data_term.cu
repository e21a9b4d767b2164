// Projective point-to-plane data term of the LM solver: fused
//   warp -> project -> validity -> bilinear(+gradient) -> residual -> 28-entry Jacobian row
//   -> J^T J / J^T r accumulation
// replacing DataLoss.prepare/forward + LossTool.* + torch.sparse.mm
// (/root/reference/super/loss.py:106-290, /root/reference/super/utils.py:17-71,
//  /root/reference/utils/utils.py:161-184).
//
// Reduction design (DESIGN.md "J^T J assembly"): surfels are visited in kNN-tuple-sorted order, so a
// warp's 32 surfels almost always share the same 4 ED nodes.  Each lane computes its own row
// [j_0..j_27 | r | 0 0 0]; the warp stages the 32x32 panel in shared memory (column-major, stride 36
// doubles: conflict-free) and accumulates the 32x32 Gram matrix  S += panel^T panel  with FP64 tensor
// core MMAs (mma.sync m8n8k4 f64, 10 upper-triangular 8x8 tiles).  S[0:28,0:28] is the 4x4 grid of
// 7x7 node-pair blocks, S[0:28,28] is J^T r and S[28,28] is sum r^2.  The accumulator is flushed with
// one f64 atomic per entry only when the node tuple changes.  No 3 400-op ATen chain, no COO
// Jacobian, no SpGEMM.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "lm_state.cuh"
#include "reg_terms.cuh"
#include "internal.h"
#include "super_b200.h"

namespace {

struct DataArgs {
    const double* points;   // (N,3) f64
    const int* knn_idx;     // (N,4) i32
    const double* knn_w;    // (N,4) f64
    const int* order;       // (N,) tuple-sorted surfel ids or nullptr (natural order)
    int n_cap;
    const int* n_dev;
    const double* ed_points;  // (J,3)
    const double* beta;       // (J,7)
    int J;
    const float4* vmap;  // (P,) new-frame vertex map, w = 1 valid / 0 invalid
    const float4* nmap;  // (P,) new-frame normal map
    Cam cam;
    double lambda;
};

struct Eval {
    int idx[SB_KNN];
    int flv, cev, flu, ceu;
    double r;
};

// Per-surfel evaluation.  Returns true when the surfel has a valid projective correspondence
// (the reference's valid_pair & intrpl_valid, loss.py:229-246).  jrow: 28 doubles when GRAD.
// NS: the node data (ed_points | beta, 10 doubles per node: g[3], q[4], t[3]) is read from the shared-memory copy
// `nodes_s` instead of global memory (the frame loop's evaluation pass: 80 node reads per surfel).
//
// Memory rounds per surfel (each a dependent L2/DRAM round trip at 20 warps per SM, which is what the pass' time is made
// of -- ncu r2g/r2n: long-scoreboard stalls 6.6 per issue): (1) the surfel's own row, (2) the eight corner samples, all
// issued together.  The reference's test of the ROUNDED pixel (valid_pair) needs no load of its own: round(x) is floor(x)
// or ceil(x), so that pixel is one of the four corners -- if the corners are inside the image and valid so is the
// rounded pixel, and if it is invalid or outside so is a corner; matched = corners inside & all four valid.
__device__ __forceinline__ float4 ldg_f4_volatile(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <bool GRAD, bool CANON = false, bool NS = false>
__device__ __forceinline__ bool eval_surfel(const DataArgs& a, int i, Eval& ev, double* jrow, int jstride,
                                            const double* nodes_s = nullptr) {
    const double* pp = a.points + 3 * (size_t)i;
    V3 p = v3(pp[0], pp[1], pp[2]);
    const int4 id4 = *reinterpret_cast<const int4*>(a.knn_idx + 4 * (size_t)i);
    ev.idx[0] = id4.x; ev.idx[1] = id4.y; ev.idx[2] = id4.z; ev.idx[3] = id4.w;
    const double2 w01 = *reinterpret_cast<const double2*>(a.knn_w + 4 * (size_t)i);
    const double2 w23 = *reinterpret_cast<const double2*>(a.knn_w + 4 * (size_t)i + 2);
    double w[SB_KNN] = {w01.x, w01.y, w23.x, w23.y};

    V3 T = v3(0, 0, 0);
#pragma unroll
    for (int k = 0; k < SB_KNN; ++k) {
        V3 gk;
        double qw;
        V3 qv, tk;
        if (NS) {
            const double* nd = nodes_s + 10 * ev.idx[k];
            gk = v3(nd[0], nd[1], nd[2]);
            qw = nd[3]; qv = v3(nd[4], nd[5], nd[6]); tk = v3(nd[7], nd[8], nd[9]);
        } else {
            const double* g = a.ed_points + 3 * ev.idx[k];
            const double* b = a.beta + 7 * ev.idx[k];
            gk = v3(__ldg(g), __ldg(g + 1), __ldg(g + 2));
            qw = __ldg(b); qv = v3(__ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
            tk = v3(__ldg(b + 4), __ldg(b + 5), __ldg(b + 6));
        }
        V3 cpk;
        V3 tv = quat_rot_ref(v3(subr(p.x, gk.x), subr(p.y, gk.y), subr(p.z, gk.z)), qw, qv, cpk);
        tv.x = addr(addr(tv.x, tk.x), gk.x);
        tv.y = addr(addr(tv.y, tk.y), gk.y);
        tv.z = addr(addr(tv.z, tk.z), gk.z);
        if (k == 0) {
            T = v3(mulr(w[k], tv.x), mulr(w[k], tv.y), mulr(w[k], tv.z));
        } else {
            T.x = addr(T.x, mulr(w[k], tv.x));
            T.y = addr(T.y, mulr(w[k], tv.y));
            T.z = addr(T.z, mulr(w[k], tv.z));
        }
    }
    double u, v;
    project_ref(T, a.cam, u, v);
    if (!(fabs(u) < 1e9 && fabs(v) < 1e9)) return false;
    const int H = a.cam.H, W = a.cam.W;

    double fv = floor(v), cv = ceil(v), fu = floor(u), cu = ceil(u);
    int iy[2] = {(int)fv, (int)cv}, ix[2] = {(int)fu, (int)cu};
    ev.flv = iy[0]; ev.cev = iy[1]; ev.flu = ix[0]; ev.ceu = ix[1];
    if (iy[0] < 0 || iy[1] >= H || ix[0] < 0 || ix[1] >= W) return false;
    // all eight samples in flight at once; corner order (fl_v,fl_u),(fl_v,ce_u),(ce_v,fl_u),(ce_v,ce_u)
    // (volatile: the compiler otherwise sinks the normal-map loads below the validity branch -- two rounds, not one)
    float4 pv[4], nv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int pix = iy[c >> 1] * W + ix[c & 1];
        pv[c] = ldg_f4_volatile(a.vmap + pix);
        nv[c] = ldg_f4_volatile(a.nmap + pix);
    }
    if (pv[0].w == 0.f || pv[1].w == 0.f || pv[2].w == 0.f || pv[3].w == 0.f) return false;
    double dy[2] = {fv - v, cv - v}, dx[2] = {fu - u, cu - u};
    double wy[2] = {fmax(1.0 - fabs(dy[0]), 0.0), fmax(1.0 - fabs(dy[1]), 0.0)};
    double wx[2] = {fmax(1.0 - fabs(dx[0]), 0.0), fmax(1.0 - fabs(dx[1]), 0.0)};
    V3 o = v3(0, 0, 0), n = v3(0, 0, 0);
    V3 o_u = v3(0, 0, 0), o_v = v3(0, 0, 0), n_u = v3(0, 0, 0), n_v = v3(0, 0, 0);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int yi = c >> 1, xi = c & 1;
        const double wgt_y = wy[yi], wgt_x = wx[xi];
        // reference: (U * w_y) * w_x, summed over corners in this order
        o.x += ((double)pv[c].x * wgt_y) * wgt_x; o.y += ((double)pv[c].y * wgt_y) * wgt_x; o.z += ((double)pv[c].z * wgt_y) * wgt_x;
        n.x += ((double)nv[c].x * wgt_y) * wgt_x; n.y += ((double)nv[c].y * wgt_y) * wgt_x; n.z += ((double)nv[c].z * wgt_y) * wgt_x;
        if (GRAD) {
            const double sx = dx[xi] >= 0.0 ? 1.0 : -1.0, sy = dy[yi] >= 0.0 ? 1.0 : -1.0;
            const double gu = wgt_y * sx, gv = wgt_x * sy;   // d/du, d/dv  (loss.py:147-150)
            o_u.x += pv[c].x * gu; o_u.y += pv[c].y * gu; o_u.z += pv[c].z * gu;
            o_v.x += pv[c].x * gv; o_v.y += pv[c].y * gv; o_v.z += pv[c].z * gv;
            n_u.x += nv[c].x * gu; n_u.y += nv[c].y * gu; n_u.z += nv[c].z * gu;
            n_v.x += nv[c].x * gv; n_v.y += nv[c].y * gv; n_v.z += nv[c].z * gv;
        }
    }
    V3 diff = v3(T.x - o.x, T.y - o.y, T.z - o.z);
    ev.r = a.lambda * ((n.x * diff.x + n.y * diff.y) + n.z * diff.z);
    if (GRAD) {
        // a^T = n^T (I - dO/dT) + diff^T dN/dT,  d(u,v)/dT uses Z without the 1e-8 (loss.py:161-173)
        const double cu_ = dot3(diff, n_u) - dot3(n, o_u);
        const double cv_ = dot3(diff, n_v) - dot3(n, o_v);
        const double iz = 1.0 / T.z;
        const double fxz = a.cam.fx * iz, fyz = a.cam.fy * iz;
        V3 av;
        av.x = n.x + cu_ * fxz;
        av.y = n.y + cv_ * fyz;
        av.z = n.z - (cu_ * fxz * T.x + cv_ * fyz * T.y) * iz;
        if (NS) asm volatile("" ::: "memory");      // keep the node re-reads below the corner samples (registers)
#pragma unroll
        for (int k = 0; k < SB_KNN; ++k) {
            // node data re-read (shared memory / L1 hits) instead of being kept live across the bilinear section
            V3 gk, qvk;
            double qwk;
            if (NS) {
                const double* nd = nodes_s + 10 * ev.idx[k];
                gk = v3(nd[0], nd[1], nd[2]);
                qwk = nd[3]; qvk = v3(nd[4], nd[5], nd[6]);
            } else {
                const double* g = a.ed_points + 3 * ev.idx[k];
                const double* b = a.beta + 7 * ev.idx[k];
                gk = v3(__ldg(g), __ldg(g + 1), __ldg(g + 2));
                qwk = __ldg(b); qvk = v3(__ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
            }
            const V3 dk = v3(p.x - gk.x, p.y - gk.y, p.z - gk.z);
            const V3 cpk = cross3(qvk, dk);
            const double s = a.lambda * w[k];
            const double qd = dot3(qvk, dk), aq = dot3(av, qvk), ad = dot3(av, dk);
            const V3 axd = cross3(av, dk);
            // CANON: the node blocks of the row are laid out in ascending node-id order, so that all surfels with the
            // same node SET (any kNN order) share one Gram accumulator
            int slot_k = k;
            if (CANON) {
                slot_k = 0;
#pragma unroll
                for (int j = 0; j < SB_KNN; ++j) slot_k += (ev.idx[j] < ev.idx[k]) ? 1 : 0;
            }
            double* jr = jrow + 7 * slot_k * jstride;
            jr[0] = s * 2.0 * dot3(av, cpk);
            jr[1 * jstride] = s * 2.0 * (qd * av.x + aq * dk.x - 2.0 * ad * qvk.x - qwk * axd.x);
            jr[2 * jstride] = s * 2.0 * (qd * av.y + aq * dk.y - 2.0 * ad * qvk.y - qwk * axd.y);
            jr[3 * jstride] = s * 2.0 * (qd * av.z + aq * dk.z - 2.0 * ad * qvk.z - qwk * axd.z);
            jr[4 * jstride] = s * av.x;
            jr[5 * jstride] = s * av.y;
            jr[6 * jstride] = s * av.z;
        }
    }
    return true;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int JT_STRIDE = 36;           // doubles per panel column (32 surfels + 4 pad)
constexpr int JT_DOUBLES = 32 * JT_STRIDE;
constexpr int JTJ_WARPS = 4;

// 4 node ids (distinct, < 65535) packed in ASCENDING order: the key of the node set
__device__ __forceinline__ unsigned long long pack_key(const int* idx) {
    unsigned long long key = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) rank += (idx[j] < idx[k]) ? 1 : 0;
        key |= (unsigned long long)(unsigned)idx[k] << (48 - 16 * rank);
    }
    return key;
}

constexpr int FL_DOUBLES = 436 + 8;      // flush scratch per warp: packed upper triangle (m <= n <= 28) of the Gram matrix + 16 ints of bases

// Flush one warp accumulator (10 tiles x 2 doubles per lane) into the lower-triangular A (dense or band, common.cuh
// MatView) and g (= -J^T r).  A flush costs the same whether the node set had 1 or 1 000 surfels and sits on the critical
// path of the warp that meets many small node sets, so everything that can be is moved out of its per-entry loop:
//   * per CTA, once: tables over the 435 packed upper-triangle entries e = (m, n), m = 7 km + cm, n = 7 kn + cn --
//     which of the 10 node pairs (or 4 right-hand-side segments) the entry belongs to, and its offset inside that pair's
//     7x7 block of the matrix in both orientations (cm L + cn and cn L + cm, L = the matrix' row pitch); and the packed
//     position of every accumulator fragment (st_idx).
//   * per flush, once: the 14 block origins -- lane p < 10 turns the solver positions of its node pair into the origin of
//     that block in the lower triangle and the orientation (which node is the row), lanes 10..13 the origins in g.
//   * per entry: two table reads, one select, one add, the atomic.
// (ncu r2c: 135 instructions per entry and 1 900 per flush before; the warps with ~10 node-set changes in their 128
// surfels set the kernel's duration.)
struct FlushTables {
    unsigned int off_a[435];     // cm * L + cn  (block stored with node km as the row)   | for g entries: cm
    unsigned int off_b[435];     // cn * L + cm  (block stored with node kn as the row)
    unsigned char pair[435];     // 0..9 node pair (km <= kn), 10..13 g segment of node km, 15 = the r^2 corner
    unsigned short st_idx[20][32];   // accumulator fragment (tile t, element e) of lane l -> packed entry, 0xffff = outside
};

__device__ __forceinline__ void build_flush_tables(FlushTables& T, int tid, int nthreads, int L) {
    for (int e = tid; e < 435; e += nthreads) {
        int m = (int)((59.f - sqrtf(3481.f - 8.f * (float)e)) * 0.5f);
        while ((m + 1) * 29 - (m + 1) * m / 2 <= e) ++m;
        while (m * 29 - m * (m - 1) / 2 > e) --m;
        const int n = m + (e - (m * 29 - m * (m - 1) / 2));
        const int km = m / 7, cm = m - 7 * km, kn = n / 7, cn = n - 7 * kn;
        if (kn == 4) {
            T.pair[e] = (unsigned char)(km == 4 ? 15 : 10 + km);
            T.off_a[e] = T.off_b[e] = (unsigned)cm;
        } else {
            T.pair[e] = (unsigned char)(km * 4 - km * (km - 1) / 2 + (kn - km));
            T.off_a[e] = (unsigned)(cm * L + cn);
            T.off_b[e] = (unsigned)(cn * L + cm);
        }
    }
    for (int i = tid; i < 20 * 32; i += nthreads) {
        const int t = i >> 6, el = (i >> 5) & 1, lane = i & 31;
        int ti = 0, rem = t;                                    // tile t -> (ti, tj), ti <= tj, row-major over the 10 tiles
        while (rem >= 4 - ti) { rem -= 4 - ti; ++ti; }
        const int tj = ti + rem;
        const int m = 8 * ti + (lane >> 2), n = 8 * tj + 2 * (lane & 3) + el;
        T.st_idx[2 * t + el][lane] = (unsigned short)((m <= n && n <= 28) ? m * 29 - m * (m - 1) / 2 + (n - m) : 0xffff);
    }
}

template <int MODE>
__device__ __forceinline__ void flush_acc(double (&acc)[10][2], unsigned long long key, int lane, const MatView& M,
                                          double* loss_cur, double* __restrict__ St, const FlushTables& T) {
#pragma unroll
    for (int t = 0; t < 10; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const unsigned idx = T.st_idx[2 * t + e][lane];
            if (idx != 0xffffu) St[idx] = acc[t][e];
            acc[t][e] = 0.0;
        }
    // block origins of this node set: lane k < 4 looks up the solver position of node k (ascending node id = the order of
    // the row's blocks), lanes 0..9 turn their pair into (origin << 1 | orientation), lanes 10..13 the origins in g
    int* bases = reinterpret_cast<int*>(St + 436);
    const int my_pos = (lane < 4) ? M.pos((int)(key >> (48 - 16 * lane)) & 0xffff) : 0;
    {
        int a = 0, rem = lane;
        while (a < 3 && rem >= 4 - a) { rem -= 4 - a; ++a; }
        const int b2 = lane < 10 ? a + rem : 0;
        const int pa = __shfl_sync(0xffffffffu, my_pos, lane < 10 ? a : (lane - 10) & 3);
        const int pb = __shfl_sync(0xffffffffu, my_pos, b2);
        int v;
        if (lane < 10) {
            const int pr = max(pa, pb), pc = min(pa, pb);        // row block = the later position: lower triangle
            const bool as_a = pa > pb;                           // node a is the row: offsets cm L + cn; else cn L + cm
            if (M.bw >= 0 && 7 * (pr - pc) + 6 > M.bw) {
                v = -1;
                atomicOr(M.overflow, 1);
            } else {
                const int L = M.bw < 0 ? M.lda : M.lda - 1;
                v = ((7 * pr * L + 7 * pc + (M.bw < 0 ? 0 : M.bw)) << 1) | (as_a ? 0 : 1);
            }
        } else {
            v = 7 * pa;
        }
        if (lane < 14) bases[lane] = v;
    }
    __syncwarp();
#pragma unroll 2
    for (int e = lane; e < 434; e += 32) {                       // entry 434 = sum r^2: dealt with below
        const double val = St[e];
        if (val == 0.0) continue;
        const int p = T.pair[e];
        const int base = bases[p];
        if (p >= 10) {
            M.put<MODE>(M.g + base + T.off_a[e], -val, M.gscale);
        } else if (base >= 0) {
            M.put<MODE>(M.A + (size_t)(base >> 1) + ((base & 1) ? T.off_b[e] : T.off_a[e]), val, M.scale);
        }
    }
    if (loss_cur && lane == 18) atomicAdd(loss_cur, St[434]);
    __syncwarp();
}

// Frame loop (sb_lm_frame, lm_frame.cu): the evaluation and the Gram accumulation are TWO launches.
//   data_eval_decide_kernel<true>  one thread per slot of the visiting order, 64-80 registers, 24+ warps per SM: warp ->
//       project -> bilinear -> residual -> Jacobian row, written as 29 coalesced column stores (rows, SoA) + the node-set
//       key; sum r^2 per block; the last block takes the LM decision for the beta the pass was evaluated at.
//   data_jtj_kernel<true>          (below) reads the rows back (L2), stages them in the same shared panel and runs the
//       DMMA Gram products + flush; assembles into the store the decision made current; skipped after a reject.
// The fused single-launch form needs 128 registers (16 warps per SM) and walks every warp through 4 chunks of dependent
// gathers + FP64 chains one after the other: 134 us per pass at 3.0e5 surfels (ncu r2c) against the sum of the two here.
struct GramArgs {
    double* store[2];          // (AB | g) as int64 fixed point; the pass assembles into store[st->sel]
    long long g_off;           // elements from the start of a store to its g
    const LMState* st;
    const double* rows;        // (29, row_stride) f64: column c of the Jacobian row of slot s at rows[c*row_stride + s]
    const unsigned long long* keys;   // (n,) node-set key per slot, ~0 = no correspondence
    int row_stride;
    // Gram records: instead of walking the 434 matrix entries of a finished accumulator itself (~900 instructions on the
    // critical path of the warp; warps that meet 5-7 node sets in their 128 surfels set the pass' duration), a warp
    // dumps the accumulator as one RECORD (fragment order -> packed upper triangle, 20 predicated stores per lane) and
    // moves on; jtj_scatter_kernel adds all records into the store with one thread per entry.  The adds are integer, so
    // neither the order in which records are claimed nor the order in which they are scattered changes the result.
    double* rec_vals;          // (rec_cap, REC_STRIDE) f64
    unsigned long long* rec_keys;   // (rec_cap,)
    int* rec_count;            // records claimed in this pass (zeroed by the evaluation pass' deciding block)
    int rec_cap;               // beyond it a warp falls back to adding its accumulator directly
};
constexpr int REC_STRIDE = 436;

// Stand-alone J^T J pass (sb_data_term_jtj): evaluation and Gram products in one launch, rows staged in a shared panel.
__global__ void __launch_bounds__(JTJ_WARPS * 32, 4)
data_jtj_kernel(DataArgs a, MatView M, double* loss_cur) {
    extern __shared__ double smem[];
    __shared__ FlushTables fl_tab;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Jt = smem + warp * (JT_DOUBLES + FL_DOUBLES);
    double* St = Jt + JT_DOUBLES;
    for (int c = 29; c < 32; ++c) Jt[c * JT_STRIDE + lane] = 0.0;   // padding columns stay zero
    build_flush_tables(fl_tab, threadIdx.x, JTJ_WARPS * 32, M.bw < 0 ? M.lda : M.lda - 1);
    __syncthreads();

    const int n = n_active(a.n_cap, a.n_dev);
    const int n_chunks = (n + 31) >> 5;
    const int total_warps = gridDim.x * JTJ_WARPS;
    const int gw = blockIdx.x * JTJ_WARPS + warp;
    const int per = (n_chunks + total_warps - 1) / total_warps;
    const int c0 = gw * per, c1 = min(n_chunks, c0 + per);

    double acc[10][2];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t][0] = acc[t][1] = 0.0;
    unsigned long long acc_key = 0;
    bool have = false;

    // one extra "tail" pass with a sentinel key flushes the last accumulator through the same code as a tuple
    // change inside the loop (a single inlined copy of the flush)
    for (int c = c0; c <= c1; ++c) {
        const bool tail = (c == c1);
        const int slot = c * 32 + lane;
        bool matched = false;
        unsigned long long key = ~0ull;
        Eval ev;
        if (!tail && slot < n) {
            const int sid = a.order ? a.order[slot] : slot;
            matched = eval_surfel<true, true>(a, sid, ev, Jt + lane, JT_STRIDE);   // row -> panel column `lane`
        }
        if (matched) {
            Jt[28 * JT_STRIDE + lane] = ev.r;
            key = pack_key(ev.idx);
        } else if (!tail) {
#pragma unroll
            for (int col = 0; col < 29; ++col) Jt[col * JT_STRIDE + lane] = 0.0;
        }
        __syncwarp();
        unsigned remaining = tail ? (have ? 1u : 0u) : __ballot_sync(0xffffffffu, matched);
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const unsigned long long k = __shfl_sync(0xffffffffu, key, leader);      // tail: the sentinel ~0
            const unsigned m = tail ? 0u : (__ballot_sync(0xffffffffu, key == k) & remaining);
            if (!have || k != acc_key) {
                if (have) flush_acc<0>(acc, acc_key, lane, M, loss_cur, St, fl_tab);
                acc_key = k;
                have = !tail;
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const unsigned m4 = (m >> (4 * ks)) & 0xfu;
                if (m4 == 0u) continue;   // warp-uniform
                // rows of other tuples in this k-step are masked out (1.0 / 0.0 factor)
                const double keep = ((m4 >> (lane & 3)) & 1u) ? 1.0 : 0.0;
                double x[4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    x[t] = keep * Jt[(8 * t + (lane >> 2)) * JT_STRIDE + 4 * ks + (lane & 3)];
                int t = 0;
#pragma unroll
                for (int ti = 0; ti < 4; ++ti)
#pragma unroll
                    for (int tj = ti; tj < 4; ++tj, ++t) dmma884(acc[t][0], acc[t][1], x[ti], x[tj]);
            }
            remaining = tail ? 0u : (remaining & ~m);
        }
        __syncwarp();
    }
}

// Gram pass of the frame loop: the rows of the last evaluation pass (SoA, rows[c * row_stride + slot]) are the operands
// of the FP64 tensor-core products AS THEY LIE IN MEMORY -- the m8n8k4 A/B fragment of k-step ks and tile t is
// rows[(8 t + lane/4) * stride + 32 chunk + 4 ks + lane%4]: the four lanes of a row read one full 32-byte sector, so a
// fragment load uses every byte it moves and no shared-memory panel, no staging stores and no barrier is needed.  All 32
// fragment loads of a chunk are issued before its 80 products (64 registers in flight per lane).
__global__ void __launch_bounds__(JTJ_WARPS * 32, 4)
jtj_gram_kernel(int n_cap, const int* __restrict__ n_dev, MatView M, GramArgs f) {
    extern __shared__ double smem[];
    __shared__ FlushTables fl_tab;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!f.st->last_accept) return;             // rejected step: the current system is kept, nothing to assemble
    M.A = f.store[f.st->sel];
    M.g = M.A + f.g_off;
    double* St = smem + warp * FL_DOUBLES;
    build_flush_tables(fl_tab, threadIdx.x, JTJ_WARPS * 32, M.bw < 0 ? M.lda : M.lda - 1);
    __syncthreads();

    const int n = n_active(n_cap, n_dev);
    const int n_chunks = (n + 31) >> 5;
    const int total_warps = gridDim.x * JTJ_WARPS;
    const int gw = blockIdx.x * JTJ_WARPS + warp;
    const int per = (n_chunks + total_warps - 1) / total_warps;
    const int c0 = gw * per, c1 = min(n_chunks, c0 + per);

    double acc[10][2];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t][0] = acc[t][1] = 0.0;
    unsigned long long acc_key = 0;
    bool have = false;
    const int frow = lane >> 2, fcol = lane & 3;

    // the key of the NEXT chunk is fetched one iteration ahead: its latency would otherwise sit in front of the fragment
    // loads of every chunk (ncu r2g: the shuffle that first uses the key was the kernel's top stall)
    unsigned long long key_next = (c0 < c1 && c0 * 32 + lane < n) ? __ldcg(f.keys + c0 * 32 + lane) : ~0ull;
    for (int c = c0; c <= c1; ++c) {
        const bool tail = (c == c1);
        const unsigned long long key = tail ? ~0ull : key_next;
        key_next = (c + 1 < c1 && (c + 1) * 32 + lane < n) ? __ldcg(f.keys + (c + 1) * 32 + lane) : ~0ull;
        const bool matched = key != ~0ull;
        unsigned remaining = tail ? (have ? 1u : 0u) : __ballot_sync(0xffffffffu, matched);
        // fragments of the whole chunk: x[ks][t] = row 8t + frow, surfel 4ks + fcol  (rows 29..31 do not exist: zero).
        // Slots >= n or without correspondence hold stale values: they are masked by select below, never multiplied.
        double x[8][4];
        if (remaining && !tail) {
            const double* src = f.rows + (size_t)frow * f.row_stride + (size_t)c * 32 + fcol;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    x[ks][t] = (8 * t + frow < 29) ? __ldcg(src + (size_t)(8 * t) * f.row_stride + 4 * ks) : 0.0;
        }
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const unsigned long long k = __shfl_sync(0xffffffffu, key, leader);      // tail: the sentinel ~0
            const unsigned m = tail ? 0u : (__ballot_sync(0xffffffffu, key == k) & remaining);
            if (!have || k != acc_key) {
                if (have) {
                    int rec = 0;
                    if (lane == 0) rec = atomicAdd(f.rec_count, 1);
                    rec = __shfl_sync(0xffffffffu, rec, 0);
                    if (rec < f.rec_cap) {
                        if (lane == 0) f.rec_keys[rec] = acc_key;
                        double* out = f.rec_vals + (size_t)rec * REC_STRIDE;
                        int t = 0;
#pragma unroll
                        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
                            for (int tj = ti; tj < 4; ++tj, ++t) {
                                const int mm = 8 * ti + frow;
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int nn = 8 * tj + 2 * fcol + e;
                                    if (mm <= nn && nn <= 28) __stcg(out + mm * 29 - mm * (mm - 1) / 2 + (nn - mm), acc[t][e]);
                                    acc[t][e] = 0.0;
                                }
                            }
                    } else {
                        flush_acc<1>(acc, acc_key, lane, M, nullptr, St, fl_tab);
                    }
                }
                acc_key = k;
                have = !tail;
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const unsigned m4 = (m >> (4 * ks)) & 0xfu;
                if (m4 == 0u) continue;   // warp-uniform
                const bool keep = (m4 >> fcol) & 1u;          // surfels of other node sets in this k-step are masked out
                double y[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) y[t] = keep ? x[ks][t] : 0.0;
                int t = 0;
#pragma unroll
                for (int ti = 0; ti < 4; ++ti)
#pragma unroll
                    for (int tj = ti; tj < 4; ++tj, ++t) dmma884(acc[t][0], acc[t][1], y[ti], y[tj]);
            }
            remaining = tail ? 0u : (remaining & ~m);
        }
    }
}

// entry e of the packed upper triangle -> (pair 0..9 | g segment 10..13 | 15 = r^2 corner, cm, cn): built once per block
__device__ __forceinline__ void build_entry_codes(unsigned short* code, int tid, int nthreads) {
    for (int e = tid; e < 435; e += nthreads) {
        int m = (int)((59.f - sqrtf(3481.f - 8.f * (float)e)) * 0.5f);
        while ((m + 1) * 29 - (m + 1) * m / 2 <= e) ++m;
        while (m * 29 - m * (m - 1) / 2 > e) --m;
        const int n = m + (e - (m * 29 - m * (m - 1) / 2));
        const int km = m / 7, cm = m - 7 * km, kn = n / 7, cn = n - 7 * kn;
        code[e] = (unsigned short)(km | (cm << 3) | (kn << 6) | (cn << 9));
    }
}

// One thread per (record, entry): adds the Gram records of the pass into the current store.
__global__ void __launch_bounds__(256) jtj_scatter_kernel(MatView M, GramArgs f) {
    __shared__ unsigned short code[448];
    if (!f.st->last_accept) return;
    M.A = f.store[f.st->sel];
    M.g = M.A + f.g_off;
    build_entry_codes(code, threadIdx.x, 256);
    __syncthreads();
    const int n_rec = min(*f.rec_count, f.rec_cap);
    const long long total = (long long)n_rec * REC_STRIDE;
    const int L = M.bw < 0 ? M.lda : M.lda - 1, bwoff = M.bw < 0 ? 0 : M.bw;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int rec = (int)(i / REC_STRIDE), e = (int)(i - (long long)rec * REC_STRIDE);
        if (e >= 434) continue;                              // 434 = sum r^2 (the evaluation pass owns the loss), 435 = pad
        const double val = __ldcg(f.rec_vals + i);
        if (val == 0.0) continue;
        const unsigned long long key = __ldg(f.rec_keys + rec);
        const unsigned c = code[e];
        const int km = c & 7, cm = (c >> 3) & 7, kn = (c >> 6) & 7, cn = (c >> 9) & 7;
        const int pm = M.pos((int)(key >> (48 - 16 * km)) & 0xffff);
        if (kn == 4) {
            M.put<1>(M.g + 7 * pm + cm, -val, M.gscale);
            continue;
        }
        const int pn = M.pos((int)(key >> (48 - 16 * kn)) & 0xffff);
        const int gm = 7 * pm + cm, gn = 7 * pn + cn;
        const int row = max(gm, gn), col = min(gm, gn);
        if (M.bw >= 0 && row - col > M.bw) { atomicOr(M.overflow, 1); continue; }
        M.put<1>(M.A + (size_t)row * L + col + bwoff, val, M.scale);
    }
}

constexpr int LOSS_BLOCK = 256;

// Loss-only pass (DataLoss.forward(grad=False)): sum r^2, one deterministic partial per block.
__global__ void __launch_bounds__(LOSS_BLOCK) data_loss_kernel(DataArgs a, double* __restrict__ partials) {
    __shared__ double red[LOSS_BLOCK / 32];
    const int n = n_active(a.n_cap, a.n_dev);
    double s = 0.0;
    for (int i = blockIdx.x * LOSS_BLOCK + threadIdx.x; i < n; i += gridDim.x * LOSS_BLOCK) {
        Eval ev;
        if (eval_surfel<false>(a, i, ev, nullptr, 1)) s += ev.r * ev.r;
    }
    s = block_sum<LOSS_BLOCK>(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// The same pass with the LM accept/reject step behind it IN the launch: every block delivers its partial and takes a
// ticket; the block that draws the last one sums all partials in their fixed order, evaluates the regularisers' losses
// and runs lm_decide_body (lm_state.cuh) -- what sb_lm_decide_reg does as a launch of its own.
// ROWS: the evaluation pass of the frame loop -- slots of the visiting order instead of surfel ids, and the Jacobian row
// of every slot is written out for the Gram pass (data_jtj_kernel<true>).
struct DecideArgs { LMState* st; double* beta; double* best; int n; RegLossArgs rg; int flip_sel; int adopt; int* rec_count; };
// (the regulariser blocks of the ROWS launch deliver their (arap, rot) loss sums behind the per-block data partials:
//  partials[gridDim.x + 2 b], [.. + 1] for block b < reg_blocks -- the deciding block then adds reg_blocks pairs
//  instead of re-evaluating 5 J residuals behind dependent loads)
struct RowsOut { double* rows; unsigned long long* keys; int row_stride; };
// The regularisers of the frame loop ride in the evaluation launch: its FIRST reg_blocks blocks assemble the ARAP / Rot
// normal-equation terms at the pass' beta (one item per thread, ~5 k instructions of straight-line atomics: 20-30 us for
// the one warp that runs them, which is why they are given blocks of their own instead of a place behind a surfel loop)
// into the store that is NOT current -- the one the Gram pass fills if the step is accepted; after a reject that store is
// cleared again by band_from_fixed_kernel before it is next used.
struct RegAsm { RegArgs reg; MatView M; double* store[2]; long long g_off; int reg_blocks; };
__device__ __noinline__ void eval_reg_item(const RegArgs* reg, const MatView* M, int tid, double* out) {   // all in SHARED memory
    double la, lr;
    reg_terms_item(*reg, tid, *M, true, la, lr);
    out[0] = la; out[1] = lr;
}
#ifndef EVAL_MINB
#define EVAL_MINB 5
#endif
constexpr int EVAL_BLOCK = 128;          // ROWS: 128 threads x EVAL_MINB blocks per SM
// NS: every surfel block first copies the node table (ed_points | beta -> 10 doubles per node) into shared memory; the
// launcher chooses it when the table is small enough not to cost residency (C1: 266 nodes = 21 KB per block).
constexpr int EVAL_NS_MAX_J = 400;
template <bool ROWS, bool NS = false>
__global__ void __launch_bounds__(ROWS ? EVAL_BLOCK : LOSS_BLOCK, ROWS ? EVAL_MINB : 4)
data_eval_decide_kernel(DataArgs a, double* __restrict__ partials, DecideArgs d, RowsOut ro, RegAsm ra) {
    constexpr int BLOCK = ROWS ? EVAL_BLOCK : LOSS_BLOCK;
    extern __shared__ double nodes_s[];
    __shared__ double red[BLOCK / 32];
    __shared__ bool s_last;
    const int n = n_active(a.n_cap, a.n_dev);
    double s = 0.0;
    const int rb = ROWS ? ra.reg_blocks : 0;
    if (NS && (int)blockIdx.x >= rb) {
        for (int e = threadIdx.x; e < 10 * a.J; e += BLOCK) {
            const int j = e / 10, c = e - 10 * j;
            nodes_s[e] = c < 3 ? __ldg(a.ed_points + 3 * j + c) : __ldg(a.beta + 7 * j + (c - 3));
        }
        __syncthreads();
    }
    if (ROWS && (int)blockIdx.x < rb) {
        // arguments through shared memory: taking the address of a kernel parameter would move the whole parameter
        // block into local memory for every thread of the launch
        __shared__ RegArgs s_reg;
        __shared__ MatView s_M;
        if (threadIdx.x == 0) {
            s_reg = ra.reg;
            s_M = ra.M;
            s_M.A = ra.store[1 - d.st->sel];
            s_M.g = s_M.A + ra.g_off;
        }
        __shared__ double s_lalr[BLOCK][2];
        __syncthreads();
        eval_reg_item(&s_reg, &s_M, (int)(blockIdx.x * BLOCK + threadIdx.x), s_lalr[threadIdx.x]);
        const double la = block_sum<BLOCK>(s_lalr[threadIdx.x][0], red);
        const double lr = block_sum<BLOCK>(s_lalr[threadIdx.x][1], red);
        if (threadIdx.x == 0) {
            partials[gridDim.x + 2 * blockIdx.x] = la;
            partials[gridDim.x + 2 * blockIdx.x + 1] = lr;
        }
    }
    for (int i = ((int)blockIdx.x - rb) * BLOCK + threadIdx.x; i < n && (int)blockIdx.x >= rb;
         i += ((int)gridDim.x - rb) * BLOCK) {
        Eval ev;
        if (ROWS) {
            const int sid = a.order ? a.order[i] : i;
            if (!a.order) {   // this thread's NEXT surfel (gathered arrays: consecutive rows) towards L1 while this one is evaluated
                const size_t nx = (size_t)i + (size_t)((int)gridDim.x - rb) * BLOCK;
                if (nx < (size_t)n) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.points + 3 * nx));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.knn_idx + 4 * nx));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.knn_w + 4 * nx));
                }
            }
            const bool ok = eval_surfel<true, true, NS>(a, sid, ev, ro.rows + i, ro.row_stride, nodes_s);
            if (ok) {
                ro.rows[(size_t)28 * ro.row_stride + i] = ev.r;
                s += ev.r * ev.r;
            }
            ro.keys[i] = ok ? pack_key(ev.idx) : ~0ull;
        } else if (eval_surfel<false, false, NS>(a, i, ev, nullptr, 1, nodes_s)) {
            s += ev.r * ev.r;
        }
    }
    s = block_sum<BLOCK>(s, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        const unsigned int ticket = atomicAdd(&d.st->ticket, 1u);
        s_last = ticket == gridDim.x - 1;
        if (s_last) d.st->ticket = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (d.rec_count && threadIdx.x == 0) *d.rec_count = 0;       // Gram records of the pass that follows
    lm_decide_body<BLOCK>(d.st, partials, (int)gridDim.x, nullptr, d.beta, d.best, d.n, d.rg, d.flip_sel != 0,
                          d.adopt != 0, rb > 0 ? partials + gridDim.x : nullptr, rb);
}

// Per-surfel rows for parity tests and for the drop-in DataLoss.forward face.
__global__ void data_rows_kernel(DataArgs a, unsigned char* __restrict__ matched, int* __restrict__ corners,
                                 double* __restrict__ r, double* __restrict__ jrow_out) {
    const int n = n_active(a.n_cap, a.n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Eval ev;
    ev.flv = ev.cev = ev.flu = ev.ceu = 0;
    ev.r = 0.0;
    double jrow[28];
    bool ok = jrow_out ? eval_surfel<true>(a, i, ev, jrow, 1) : eval_surfel<false>(a, i, ev, nullptr, 1);
    matched[i] = ok ? 1 : 0;
    if (corners) {
        corners[4 * i + 0] = ev.flv; corners[4 * i + 1] = ev.cev;
        corners[4 * i + 2] = ev.flu; corners[4 * i + 3] = ev.ceu;
    }
    if (r) r[i] = ok ? ev.r : 0.0;
    if (jrow_out)
        for (int c = 0; c < 28; ++c) jrow_out[28 * (size_t)i + c] = ok ? jrow[c] : 0.0;
}

// 64-bit sort keys of the J^T J visiting order: packed node 4-tuple; rows >= n sort last.
__global__ void tuple_keys_kernel(const int* __restrict__ knn_idx, int n_cap, const int* n_dev,
                                  long long* __restrict__ keys, const int* __restrict__ node_pos,
                                  int* __restrict__ block_bw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int span = 0;
    if (i < n_cap) {
        const int n = n_active(n_cap, n_dev);
        if (i >= n) {
            keys[i] = 0x7fffffffffffffffLL;
        } else {
            const int4 id = *reinterpret_cast<const int4*>(knn_idx + 4 * (size_t)i);
            const int idx[4] = {id.x, id.y, id.z, id.w};
            keys[i] = (long long)(pack_key(idx) >> 1);   // keep it non-negative for a signed sort
            if (block_bw) {
                int lo = 1 << 30, hi = -1;
                for (int k = 0; k < 4; ++k) {
                    const int p = node_pos ? node_pos[idx[k]] : idx[k];
                    lo = min(lo, p); hi = max(hi, p);
                }
                span = hi - lo;
            }
        }
    }
    if (block_bw) {   // block half-bandwidth the surfel tuples need in the solver's node order
        span = __reduce_max_sync(0xffffffffu, span);
        if ((threadIdx.x & 31) == 0 && span > 0) atomicMax(block_bw, span);
    }
}

// The same key with `bits` bits per node id (J < 2^bits), 4*bits significant bits in all, and the row index next to it:
// the radix sort behind it (sb_tuple_order) then needs ceil(4*bits/8) passes instead of the 8 of a 64-bit key.
__global__ void tuple_keys_compact_kernel(const int* __restrict__ knn_idx, int n_cap, const int* n_dev, int bits,
                                          unsigned long long* __restrict__ keys, int* __restrict__ rows,
                                          const int* __restrict__ node_pos, int* __restrict__ block_bw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int span = 0;
    if (i < n_cap) {
        const int n = n_active(n_cap, n_dev);
        rows[i] = i;
        if (i >= n) {
            keys[i] = (bits >= 16) ? ~0ull : ((1ull << (4 * bits)) - 1ull);     // an all-ones field is no node id
        } else {
            const int4 id = *reinterpret_cast<const int4*>(knn_idx + 4 * (size_t)i);
            const int idx[4] = {id.x, id.y, id.z, id.w};
            unsigned long long key = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int rank = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) rank += (idx[j] < idx[k]) ? 1 : 0;
                key |= (unsigned long long)(unsigned)idx[k] << (bits * (3 - rank));
            }
            keys[i] = key;
            if (block_bw) {
                int lo = 1 << 30, hi = -1;
                for (int k = 0; k < 4; ++k) {
                    const int p = node_pos ? node_pos[idx[k]] : idx[k];
                    lo = min(lo, p); hi = max(hi, p);
                }
                span = hi - lo;
            }
        }
    }
    if (block_bw) {
        span = __reduce_max_sync(0xffffffffu, span);
        if ((threadIdx.x & 31) == 0 && span > 0) atomicMax(block_bw, span);
    }
}

// the LM inputs of the surfels in visiting order: one gather per frame makes the ten evaluation passes coalesced
__global__ void gather_sorted_kernel(const double* __restrict__ points, const int* __restrict__ knn_idx,
                                     const double* __restrict__ knn_w, const int* __restrict__ order, int n_cap,
                                     const int* n_dev, double* __restrict__ o_points, int* __restrict__ o_idx,
                                     double* __restrict__ o_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active(n_cap, n_dev)) return;
    const int s = order[i];
    o_points[3 * (size_t)i] = points[3 * (size_t)s];
    o_points[3 * (size_t)i + 1] = points[3 * (size_t)s + 1];
    o_points[3 * (size_t)i + 2] = points[3 * (size_t)s + 2];
    reinterpret_cast<int4*>(o_idx)[i] = reinterpret_cast<const int4*>(knn_idx)[s];
    reinterpret_cast<double4*>(o_w)[i] = reinterpret_cast<const double4*>(knn_w)[s];
}

DataArgs make_args(const double* points, const int* knn_idx, const double* knn_w, const int* order, int n_cap,
                   const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                   const float* nmap, int H, int W, const double* intr, double lambda) {
    DataArgs a;
    a.points = points; a.knn_idx = knn_idx; a.knn_w = knn_w; a.order = order;
    a.n_cap = n_cap; a.n_dev = n_dev; a.ed_points = ed_points; a.beta = beta; a.J = J;
    a.vmap = reinterpret_cast<const float4*>(vmap);
    a.nmap = reinterpret_cast<const float4*>(nmap);
    a.cam.fx = intr[0]; a.cam.fy = intr[1]; a.cam.cx = intr[2]; a.cam.cy = intr[3];
    a.cam.H = H; a.cam.W = W;
    a.lambda = lambda;
    return a;
}

}  // namespace

// One resident wave of the J^T J kernel on the current device (a partial second wave was a 40 % tail, ncu r1), and the
// kernel's shared-memory opt-in: both are per DEVICE, so the cache is indexed by the device ordinal.
template <typename K>
static int jtj_resident(K kernel, size_t smem, int which) {
    static int resident[2][64] = {{0}};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (resident[which][dev] == 0) {
        int per_sm = 0, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, JTJ_WARPS * 32, smem);
        resident[which][dev] = (per_sm > 0 ? per_sm : 1) * sms;
    }
    return resident[which][dev];
}

extern "C" {

int sb_data_loss_blocks(int n_cap) {
    int b = (n_cap + LOSS_BLOCK - 1) / LOSS_BLOCK;
    return b < 1 ? 1 : (b > 592 ? 592 : b);   // <= 4 CTAs per SM on 148 SMs
}

int sb_tuple_keys(const int* knn_idx, int n_cap, const int* n_dev, long long* keys, const int* node_pos,
                  int* block_bw, void* stream) {
    if (!knn_idx || !keys) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    tuple_keys_kernel<<<(n_cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(knn_idx, n_cap, n_dev, keys, node_pos,
                                                                           block_bw);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

static int order_bits(int J) {
    int b = 1;
    while (b < 16 && (1 << b) <= J) ++b;       // J <= 2^b - 1: the all-ones field stays free for the sentinel
    return b;
}

long long sb_tuple_order_temp_bytes(int n_cap) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, n_cap < 1 ? 1 : n_cap, 0, 64, (cudaStream_t)0);
    return (long long)bytes;
}

int sb_tuple_order(const int* knn_idx, int n_cap, const int* n_dev, int J, const int* node_pos, int* block_bw,
                   unsigned long long* keys, unsigned long long* keys_alt, int* rows, int* order, void* temp,
                   long long temp_bytes, void* stream) {
    if (!knn_idx || !keys || !keys_alt || !rows || !order || !temp || J <= 0 || J > 65535) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    if (temp_bytes < sb_tuple_order_temp_bytes(n_cap)) return SB_ERR_WORKSPACE;
    const int bits = order_bits(J);
    cudaStream_t st = (cudaStream_t)stream;
    tuple_keys_compact_kernel<<<(n_cap + 255) / 256, 256, 0, st>>>(knn_idx, n_cap, n_dev, bits, keys, rows, node_pos, block_bw);
    SB_CHECK_LAUNCH();
    size_t tb = (size_t)temp_bytes;
    if (cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_alt, rows, order, n_cap, 0, 4 * bits, st) != cudaSuccess)
        return SB_ERR_CUDA;
    return SB_OK;
}

int sb_data_term_jtj(const double* points, const int* knn_idx, const double* knn_w, const int* order, int n_cap,
                     const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                     const float* nmap, int H, int W, const double* intr, double lambda, double* A, int lda,
                     int bw, const int* node_pos, int* band_overflow, double* g, double* loss_cur, int fx_shift,
                     int fx_gshift, void* stream) {
    if (!points || !knn_idx || !knn_w || !ed_points || !beta || !vmap || !nmap || !A || !g) return SB_ERR_ARG;
    if (J <= 0 || J > 65535) return SB_ERR_ARG;
    if (bw < 0 ? lda < 7 * J : (lda < bw + 1 || !band_overflow)) return SB_ERR_ARG;
    if (fx_shift >= 0 && (!band_overflow || fx_gshift < 0 || fx_shift > 60 || fx_gshift > 60)) return SB_ERR_ARG;
    MatView M;
    M.A = A; M.lda = lda; M.bw = bw; M.node_pos = node_pos; M.overflow = band_overflow; M.g = g;
    M.set_shift(fx_shift, fx_gshift);
    if (n_cap <= 0) return SB_OK;
    DataArgs a = make_args(points, knn_idx, knn_w, order, n_cap, n_dev, ed_points, beta, J, vmap, nmap, H, W,
                           intr, lambda);
    const int n_chunks = (n_cap + 31) / 32;
    const size_t smem = JTJ_WARPS * (JT_DOUBLES + FL_DOUBLES) * sizeof(double);
    const int resident = jtj_resident(data_jtj_kernel, smem, 0);
    if (resident <= 0) return SB_ERR_CUDA;
    int blocks = (n_chunks + JTJ_WARPS - 1) / JTJ_WARPS;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    data_jtj_kernel<<<blocks, JTJ_WARPS * 32, smem, (cudaStream_t)stream>>>(a, M, loss_cur);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_gather_sorted(const double* points, const int* knn_idx, const double* knn_w, const int* order, int n_cap,
                     const int* n_dev, double* out_points, int* out_idx, double* out_w, void* stream) {
    if (!points || !knn_idx || !knn_w || !order || !out_points || !out_idx || !out_w) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    gather_sorted_kernel<<<(n_cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, knn_idx, knn_w, order, n_cap, n_dev,
                                                                           out_points, out_idx, out_w);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_data_term_loss(const double* points, const int* knn_idx, const double* knn_w, int n_cap,
                      const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                      const float* nmap, int H, int W, const double* intr, double lambda, double* partials,
                      int n_partials, void* stream) {
    if (!points || !knn_idx || !knn_w || !ed_points || !beta || !vmap || !nmap || !partials) return SB_ERR_ARG;
    if (n_partials != sb_data_loss_blocks(n_cap)) return SB_ERR_WORKSPACE;
    DataArgs a = make_args(points, knn_idx, knn_w, nullptr, n_cap, n_dev, ed_points, beta, J, vmap, nmap, H, W,
                           intr, lambda);
    data_loss_kernel<<<n_partials, LOSS_BLOCK, 0, (cudaStream_t)stream>>>(a, partials);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_data_term_loss_decide(const double* points, const int* knn_idx, const double* knn_w, int n_cap,
                             const int* n_dev, const double* ed_points, const double* beta_in, int J, const float* vmap,
                             const float* nmap, int H, int W, const double* intr, double lambda, double* partials,
                             int n_partials, void* state, const int* ed_knn, double lam_arap, double lam_rot,
                             int use_arap, int use_rot, double* beta, double* best, void* stream) {
    if (!points || !knn_idx || !knn_w || !ed_points || !beta_in || !vmap || !nmap || !partials) return SB_ERR_ARG;
    if (!state || !ed_knn || !beta || !best || beta_in != beta) return SB_ERR_ARG;
    if (n_partials != sb_data_loss_blocks(n_cap)) return SB_ERR_WORKSPACE;
    DataArgs a = make_args(points, knn_idx, knn_w, nullptr, n_cap, n_dev, ed_points, beta_in, J, vmap, nmap, H, W,
                           intr, lambda);
    DecideArgs d;
    d.st = (LMState*)state; d.beta = beta; d.best = best; d.n = 7 * J;
    d.rg = RegLossArgs{ed_points, ed_knn, J, lam_arap, lam_rot, use_arap, use_rot};
    d.flip_sel = 0; d.adopt = 0; d.rec_count = nullptr;
    data_eval_decide_kernel<false><<<n_partials, LOSS_BLOCK, 0, (cudaStream_t)stream>>>(a, partials, d, RowsOut{nullptr, nullptr, 0},
                                                                                  RegAsm{});
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_data_term_rows(const double* points, const int* knn_idx, const double* knn_w, int n_cap, const int* n_dev,
                      const double* ed_points, const double* beta, int J, const float* vmap, const float* nmap,
                      int H, int W, const double* intr, double lambda, unsigned char* matched, int* corners,
                      double* r, double* jrow, void* stream) {
    if (!points || !knn_idx || !knn_w || !ed_points || !beta || !vmap || !nmap || !matched) return SB_ERR_ARG;
    if (n_cap <= 0) return SB_OK;
    DataArgs a = make_args(points, knn_idx, knn_w, nullptr, n_cap, n_dev, ed_points, beta, J, vmap, nmap, H, W,
                           intr, lambda);
    data_rows_kernel<<<(n_cap + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, matched, corners, r, jrow);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"

// ---- frame loop (lm_frame.cu) ---------------------------------------------------------------------------------
namespace sbi {

// resident blocks of the evaluation pass on the current device (one wave: the grid-stride loop balances the rest)
static int eval_resident(bool ns, size_t smem) {
    static int resident[2][64] = {{0}};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (resident[ns][dev] == 0) {
        int per_sm = 0, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (ns) {
            // the largest table the NS form is chosen for, so that the cached residency holds for every J below it
            const size_t smem_max = (size_t)10 * EVAL_NS_MAX_J * sizeof(double);
            if (cudaFuncSetAttribute(data_eval_decide_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_max) != cudaSuccess)
                return -1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, data_eval_decide_kernel<true, true>, EVAL_BLOCK, smem_max);
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, data_eval_decide_kernel<true, false>, EVAL_BLOCK, 0);
        }
        resident[ns][dev] = (per_sm > 0 ? per_sm : 1) * sms;
    }
    (void)smem;
    return resident[ns][dev];
}

// evaluation pass at f->beta: rows + keys for the Gram pass, loss partials, LM decision in the last block.
// adopt != 0: prologue (the system assembled next becomes current without a decision).
int launch_eval_decide(const SbLMFrame* f, int adopt, cudaStream_t st) {
    DataArgs a = make_args(f->points, f->knn_idx, f->knn_w, f->order, f->n_cap, f->n_dev, f->ed_points, f->beta, f->J,
                           f->vmap, f->nmap, f->H, f->W, f->intr, f->lam_data);
    DecideArgs d;
    d.st = (LMState*)f->state; d.beta = f->beta; d.best = f->best; d.n = 7 * f->J;
    d.rg = RegLossArgs{f->ed_points, f->ed_knn, f->J, f->lam_arap, f->lam_rot, f->use_arap, f->use_rot};
    d.flip_sel = 1; d.adopt = adopt; d.rec_count = f->rec_count;
    RegAsm ra;
    ra.reg = RegArgs{f->ed_points, f->ed_knn, f->beta, f->J, f->lam_arap, f->lam_rot, f->use_arap, f->use_rot};
    ra.M.A = nullptr; ra.M.g = nullptr;
    ra.M.lda = f->ldab; ra.M.bw = f->bw; ra.M.node_pos = f->node_pos; ra.M.overflow = f->band_overflow;
    ra.M.set_shift(f->fx_shift, f->fx_gshift);
    ra.store[0] = reinterpret_cast<double*>(f->fx_store[0]);
    ra.store[1] = reinterpret_cast<double*>(f->fx_store[1]);
    ra.g_off = (long long)f->n * f->ldab;
    const int reg_threads = (f->use_arap ? f->J * SB_KNN : 0) + (f->use_rot ? f->J : 0);
    ra.reg_blocks = (reg_threads + EVAL_BLOCK - 1) / EVAL_BLOCK;
    int blocks = (f->n_cap + EVAL_BLOCK - 1) / EVAL_BLOCK + ra.reg_blocks;
    const bool ns = f->J <= EVAL_NS_MAX_J;
    const size_t smem = ns ? (size_t)10 * f->J * sizeof(double) : 0;
    const int resident = eval_resident(ns, smem);
    if (resident <= 0) return SB_ERR_CUDA;
    if (blocks > resident && resident > 2 * ra.reg_blocks) blocks = resident;     // one wave: regulariser + surfel blocks
    if (blocks + 2 * ra.reg_blocks > f->n_partials_loss) blocks = f->n_partials_loss - 2 * ra.reg_blocks;
    if (blocks <= ra.reg_blocks) return SB_ERR_WORKSPACE;
    const RowsOut ro{f->rows, f->keys, f->row_stride};
    if (ns) data_eval_decide_kernel<true, true><<<blocks, EVAL_BLOCK, smem, st>>>(a, f->partials_loss, d, ro, ra);
    else data_eval_decide_kernel<true, false><<<blocks, EVAL_BLOCK, 0, st>>>(a, f->partials_loss, d, ro, ra);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

// Gram pass over the rows of the last evaluation: J^T J and -J^T r into the store the decision made current (the
// regularisers' terms are already there); returns at once on the device when the step was rejected.
int launch_gram(const SbLMFrame* f, cudaStream_t st, int which) {
    MatView M;
    M.A = nullptr; M.g = nullptr;        // chosen in the kernel from the device-resident selector
    M.lda = f->ldab; M.bw = f->bw; M.node_pos = f->node_pos; M.overflow = f->band_overflow;
    M.set_shift(f->fx_shift, f->fx_gshift);
    GramArgs ga;
    ga.store[0] = reinterpret_cast<double*>(f->fx_store[0]);
    ga.store[1] = reinterpret_cast<double*>(f->fx_store[1]);
    ga.g_off = (long long)f->n * f->ldab;
    ga.st = (const LMState*)f->state;
    ga.rows = f->rows; ga.keys = f->keys; ga.row_stride = f->row_stride;
    ga.rec_vals = f->rec_vals; ga.rec_keys = f->rec_keys; ga.rec_count = f->rec_count; ga.rec_cap = f->rec_cap;
    const size_t smem = JTJ_WARPS * FL_DOUBLES * sizeof(double);
    const int resident = jtj_resident(jtj_gram_kernel, smem, 1);
    if (resident <= 0) return SB_ERR_CUDA;
    const int n_chunks = (f->n_cap + 31) / 32;
    int blocks = (n_chunks + JTJ_WARPS - 1) / JTJ_WARPS;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    if (which & 1) {
        jtj_gram_kernel<<<blocks, JTJ_WARPS * 32, smem, st>>>(f->n_cap, f->n_dev, M, ga);
        SB_CHECK_LAUNCH();
    }
    if (which & 2) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        jtj_scatter_kernel<<<8 * sms, 256, 0, st>>>(M, ga);
        SB_CHECK_LAUNCH();
    }
    return SB_OK;
}

}  // namespace sbi
