// Device-resident LM controller state and the accept/reject step, shared by lm.cu (its own launch) and data_term.cu
// (the last block of the loss-only pass runs it: one launch less per LM iteration).
#pragma once
#include "common.cuh"

namespace {

// Device-resident controller state (mirrors LM_Solver.LM's locals u, minimal_loss, best_beta).
struct LMState {
    double u;             // additive damping (reset to 10 every frame)
    double v;             // 7.5
    double minimal_loss;  // 1e10 at frame start
    int iter;             // iterations completed
    int failed;           // Cholesky failed -> the reference breaks out of the loop, beta unchanged
    double loss[64];      // trace: loss at the trial beta of iteration i
    double loss_terms[64][3];
    int accept[64];
    double u_trace[64];
    unsigned int ticket;  // blocks of the fused loss pass that have delivered their partial (self-resetting)
    int last_accept;      // frame loop: 1 when the last decision accepted the step (the Gram pass then assembles), else 0
    int sel;              // frame loop (sb_lm_frame): which of the two fixed-point stores holds the normal equations of the
                          // CURRENT beta; the J^T J pass at a trial beta assembles into the other one, an accepted step flips
};

// squared ARAP residual of (node j, neighbour slot k) = item tid, and squared Rot residual of node j: the loss-only
// halves of reg_terms_kernel (same expressions, same float32 Rot term)
__device__ __forceinline__ double arap_loss_item(const double* __restrict__ ed_points, const int* __restrict__ ed_knn,
                                                 const double* __restrict__ beta, int tid, double lam_arap) {
    const int j = tid / SB_KNN;
    const int n = ed_knn[tid];
    const V3 gj = v3(ed_points[3 * j], ed_points[3 * j + 1], ed_points[3 * j + 2]);
    const V3 gn = v3(ed_points[3 * n], ed_points[3 * n + 1], ed_points[3 * n + 2]);
    const V3 d = v3(gj.x - gn.x, gj.y - gn.y, gj.z - gn.z);
    const double* bn = beta + 7 * n;
    const double* bj = beta + 7 * j;
    const V3 qv = v3(bn[1], bn[2], bn[3]);
    V3 cp;
    V3 tv = quat_rot_ref(d, bn[0], qv, cp);
    const double r[3] = {lam_arap * ((tv.x + bn[4]) - (d.x + bj[4])), lam_arap * ((tv.y + bn[5]) - (d.y + bj[5])),
                         lam_arap * ((tv.z + bn[6]) - (d.z + bj[6]))};
    return r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
}
__device__ __forceinline__ double rot_loss_item(const double* __restrict__ beta, int j, double lam_rot) {
    const float lam = (float)lam_rot;
    float q[4];
    for (int a = 0; a < 4; ++a) q[a] = (float)beta[7 * j + a];
    const float s = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
    const float r = lam * (1.f - s);
    return (double)(r * r);
}

struct RegLossArgs {     // ed_points == nullptr: the ARAP / Rot losses come in through loss_arap_rot (separate launch)
    const double* ed_points; const int* ed_knn; int J; double lam_arap, lam_rot; int use_arap, use_rot;
};

// loss = sum(data partials) + arap + rot; accept iff loss < minimal_loss   (LM.py:107-117)
// Block-wide (BLOCK threads, every thread of the block must call it).
// flip_sel: an accepted step makes the store assembled at the trial beta the current one (frame loop); adopt: the
// prologue of the frame loop -- the system assembled at the initial beta becomes current, no LM decision is taken.
template <int BLOCK>
__device__ __forceinline__ void lm_decide_body(LMState* st, const double* partials, int n_partials, double* loss_arap_rot,
                                               double* beta, double* best, int n, RegLossArgs rg, bool flip_sel = false,
                                               bool adopt = false, const double* reg_part = nullptr, int n_reg_part = 0) {
    if (adopt) {
        if (threadIdx.x == 0) { st->sel ^= 1; st->last_accept = 1; }
        return;
    }
    __shared__ double red[BLOCK / 32];
    __shared__ int s_accept;
    __shared__ double s_reg[2];
    // This runs on ONE block behind everybody else: every dependent L2 round trip here is on the frame's critical path
    // (r2o launch list: 11 us of the evaluation launch).  Loads are issued in batches, sums keep a fixed order.
    double s = 0.0;
    {
        constexpr int PER = 8;
        for (int base = 0; base < n_partials; base += PER * BLOCK) {
            double v[PER];
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int i = base + q * BLOCK + threadIdx.x;
                v[q] = i < n_partials ? __ldcg(partials + i) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < PER; ++q) s += v[q];
        }
    }
    s = block_sum<BLOCK>(s, red);
    if (reg_part) {                          // per-block (arap, rot) sums delivered by the regulariser blocks of this launch
        if (threadIdx.x == 0) {
            double la = 0.0, lr = 0.0;
            for (int b = 0; b < n_reg_part; ++b) { la += __ldcg(reg_part + 2 * b); lr += __ldcg(reg_part + 2 * b + 1); }
            s_reg[0] = la; s_reg[1] = lr;
        }
    } else if (rg.ed_points) {               // the regularisers' losses of the stepped beta, in this launch
        double la = 0.0, lr = 0.0;
        if (rg.use_arap)
            for (int i = threadIdx.x; i < rg.J * SB_KNN; i += BLOCK) la += arap_loss_item(rg.ed_points, rg.ed_knn, beta, i, rg.lam_arap);
        if (rg.use_rot)
            for (int j = threadIdx.x; j < rg.J; j += BLOCK) lr += rot_loss_item(beta, j, rg.lam_rot);
        la = block_sum<BLOCK>(la, red);
        lr = block_sum<BLOCK>(lr, red);
        if (threadIdx.x == 0) { s_reg[0] = la; s_reg[1] = lr; }
    }
    if (threadIdx.x == 0) {
        const int it = st->iter;
        if (st->failed) {
            s_accept = -1;
            st->last_accept = 0;
        } else {
            const bool own = reg_part || rg.ed_points;
            const double la = own ? s_reg[0] : loss_arap_rot[0], lr = own ? s_reg[1] : loss_arap_rot[1];
            const double loss = s + la + lr;
            const bool acc = loss < st->minimal_loss;
            if (it < 64) {
                st->loss[it] = loss;
                st->loss_terms[it][0] = s; st->loss_terms[it][1] = la; st->loss_terms[it][2] = lr;
                st->accept[it] = acc ? 1 : 0;
                st->u_trace[it] = st->u;
            }
            if (acc) { st->minimal_loss = loss; st->u /= st->v; if (flip_sel) st->sel ^= 1; }
            else st->u *= st->v;
            st->iter = it + 1;
            st->last_accept = acc ? 1 : 0;
            s_accept = acc ? 1 : 0;
        }
        if (loss_arap_rot) {
            loss_arap_rot[0] = 0.0;
            loss_arap_rot[1] = 0.0;
        }
    }
    __syncthreads();
    const int acc = s_accept;
    if (acc < 0) return;
    {
        const double* src = acc ? beta : best;
        double* dst = acc ? best : beta;
        constexpr int PER = 8;
        for (int base = 0; base < n; base += PER * BLOCK) {
            double v[PER];
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int i = base + q * BLOCK + threadIdx.x;
                v[q] = i < n ? __ldcg(src + i) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int i = base + q * BLOCK + threadIdx.x;
                if (i < n) dst[i] = v[q];
            }
        }
    }
}


}  // namespace
