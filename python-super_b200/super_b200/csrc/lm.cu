// LM solver pieces other than the data term: ARAP + Rot regularisers (residual, analytic Jacobian,
// normal-equation assembly), additive damping, step application and the device-resident
// accept/reject controller.  Restates /root/reference/super/loss.py:403-499 and
// /root/reference/super/LM.py:81-122 without any host synchronisation: u, minimal_loss, the
// failure flag and the per-iteration trace live in a small device struct.
#include "common.cuh"
#include "lm_state.cuh"
#include "super_b200.h"

namespace {

__device__ __forceinline__ void add_lower(const MatView& M, int r, int c, double v) {
    if (r >= c) M.add(r, c, v);
    else M.add(c, r, v);
}

// d[R(q)v]/dq as 3x4 (col 0 = d/dqw, cols 1..3 = d/dqv)   (/root/reference/super/utils.py:59-69)
__device__ __forceinline__ void quat_jac(const V3& v, double qw, const V3& qv, const V3& cp, double (&Jq)[3][4]) {
    const double qd = dot3(qv, v);
    const double q[3] = {qv.x, qv.y, qv.z}, vv[3] = {v.x, v.y, v.z};
    const double sk[3][3] = {{0, -v.z, v.y}, {v.z, 0, -v.x}, {-v.y, v.x, 0}};
    Jq[0][0] = 2.0 * cp.x; Jq[1][0] = 2.0 * cp.y; Jq[2][0] = 2.0 * cp.z;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Jq[i][1 + j] = 2.0 * ((i == j ? qd : 0.0) + q[i] * vv[j] - 2.0 * vv[i] * q[j] - qw * sk[i][j]);
}

// One thread per (node j, neighbour slot k) for ARAP; one thread per node for Rot (threads >= J*K).
// With A == nullptr only the loss partials are produced.
__global__ void reg_terms_kernel(const double* __restrict__ ed_points, const int* __restrict__ ed_knn,
                                 const double* __restrict__ beta, int J, double lam_arap, double lam_rot,
                                 int use_arap, int use_rot, MatView M,
                                 double* __restrict__ g, double* __restrict__ loss_arap_rot /* [2] */) {
    double* const A = M.A;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_arap = use_arap ? J * SB_KNN : 0;
    double la = 0.0, lr = 0.0;
    if (tid < n_arap) {
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
        // r = lam [ (R(q_n) d + b_n) - (d + b_j) ]        (loss.py:433-437)
        const double r[3] = {lam_arap * ((tv.x + bn[4]) - (d.x + bj[4])), lam_arap * ((tv.y + bn[5]) - (d.y + bj[5])),
                             lam_arap * ((tv.z + bn[6]) - (d.z + bj[6]))};
        la = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        if (A) {
            double Jq[3][4];
            quat_jac(d, bn[0], qv, cp, Jq);
            // residual row c: cols 7n+{0..3} = lam*Jq[c][.], 7n+4+c = lam, 7j+4+c = -lam   (loss.py:418-451)
            const int bn0 = 7 * M.pos(n), bj0 = 7 * M.pos(j);
            const double l = lam_arap, l2 = lam_arap * lam_arap;
            for (int a = 0; a < 4; ++a) {
                for (int b = 0; b <= a; ++b) {
                    double s = 0.0;
                    for (int c = 0; c < 3; ++c) s += Jq[c][a] * Jq[c][b];
                    M.add(bn0 + a, bn0 + b, l2 * s);
                }
                double gq = 0.0;
                for (int c = 0; c < 3; ++c) {
                    add_lower(M, bn0 + 4 + c, bn0 + a, l2 * Jq[c][a]);     // q_n x b_n
                    add_lower(M, bj0 + 4 + c, bn0 + a, -l2 * Jq[c][a]);    // q_n x b_j
                    gq += l * Jq[c][a] * r[c];
                }
                atomicAdd(g + bn0 + a, -gq);
            }
            for (int c = 0; c < 3; ++c) {
                M.add(bn0 + 4 + c, bn0 + 4 + c, l2);
                M.add(bj0 + 4 + c, bj0 + 4 + c, l2);
                add_lower(M, bj0 + 4 + c, bn0 + 4 + c, -l2);
                atomicAdd(g + bn0 + 4 + c, -l * r[c]);
                atomicAdd(g + bj0 + 4 + c, l * r[c]);
            }
        }
    } else if (use_rot && tid < n_arap + J) {
        // RotLoss in float32 like the reference (loss.py:487-497)
        const int j = tid - n_arap;
        const float lam = (float)lam_rot;
        float q[4];
        for (int a = 0; a < 4; ++a) q[a] = (float)beta[7 * j + a];
        const float s = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
        const float r = lam * (1.f - s);
        lr = (double)(r * r);
        if (A) {
            float jv[4];
            for (int a = 0; a < 4; ++a) jv[a] = -lam * 2.f * q[a];
            const int pj = 7 * M.pos(j);
            for (int a = 0; a < 4; ++a) {
                for (int b = 0; b <= a; ++b) M.add(pj + a, pj + b, (double)(jv[a] * jv[b]));
                atomicAdd(g + pj + a, -(double)(jv[a] * r));
            }
        }
    }
    if (loss_arap_rot) {
        __shared__ double red[8];
        double s = block_sum<256>(la, red);
        if (threadIdx.x == 0 && s != 0.0) atomicAdd(loss_arap_rot + 0, s);
        s = block_sum<256>(lr, red);
        if (threadIdx.x == 0 && s != 0.0) atomicAdd(loss_arap_rot + 1, s);
    }
}

__global__ void lm_begin_kernel(LMState* st, double* beta, double* best, int J, double u, double v,
                                double minimal_loss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        st->u = u; st->v = v; st->minimal_loss = minimal_loss; st->iter = 0; st->failed = 0; st->ticket = 0;
    }
    if (i < 7 * J) {
        const double val = (i % 7 == 0) ? 1.0 : 0.0;
        beta[i] = val;
        best[i] = val;
    }
}

__global__ void lm_damp_kernel(const LMState* st, double* A, int lda, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(size_t)i * lda + i] += st->u;
}

// beta += delta unless the factorisation failed (info != 0  ->  the reference prints and breaks).
// delta is in the solver's node order when node_pos is given.
__global__ void lm_step_kernel(LMState* st, const int* info, double* beta, const double* delta, int n,
                               const int* node_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = st->failed || (info && *info != 0);
    if (i < n && !bad) beta[i] += delta[node_pos ? 7 * node_pos[i / 7] + i % 7 : i];
    if (i == 0 && bad) st->failed = 1;
}

// loss = sum(data partials) + arap + rot; accept iff loss < minimal_loss   (LM.py:107-117): lm_decide_body, lm_state.cuh
__global__ void lm_decide_kernel(LMState* st, const double* partials, int n_partials, double* loss_arap_rot,
                                 double* beta, double* best, int n, RegLossArgs rg) {
    lm_decide_body<256>(st, partials, n_partials, loss_arap_rot, beta, best, n, rg);
}

}  // namespace

extern "C" {

int sb_lm_state_bytes() { return (int)sizeof(LMState); }

// Host-readable layout of the trace inside LMState (offsets in bytes): used by the Python face to
// decode a D2H copy of the state after the frame.
int sb_lm_state_offsets(int* out /* [8] */) {
    LMState* p = nullptr;
    out[0] = (int)(size_t)&p->u; out[1] = (int)(size_t)&p->minimal_loss; out[2] = (int)(size_t)&p->iter;
    out[3] = (int)(size_t)&p->failed; out[4] = (int)(size_t)&p->loss[0]; out[5] = (int)(size_t)&p->loss_terms[0][0];
    out[6] = (int)(size_t)&p->accept[0]; out[7] = (int)(size_t)&p->u_trace[0];
    return SB_OK;
}

int sb_lm_begin(void* state, double* beta, double* best, int J, double u, double v, double minimal_loss,
                void* stream) {
    if (!state || !beta || !best || J <= 0) return SB_ERR_ARG;
    lm_begin_kernel<<<(7 * J + 255) / 256, 256, 0, (cudaStream_t)stream>>>((LMState*)state, beta, best, J, u, v,
                                                                         minimal_loss);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_reg_terms(const double* ed_points, const int* ed_knn, const double* beta, int J, double lam_arap,
                 double lam_rot, int use_arap, int use_rot, double* A, int lda, int bw, const int* node_pos,
                 int* band_overflow, double* g, double* loss_arap_rot, void* stream) {
    if (!ed_points || !ed_knn || !beta || J <= 0) return SB_ERR_ARG;
    if (A && (!g || (bw < 0 ? lda < 7 * J : (lda < bw + 1 || !band_overflow)))) return SB_ERR_ARG;
    MatView M;
    M.A = A; M.lda = lda; M.bw = bw; M.node_pos = node_pos; M.overflow = band_overflow;
    const int threads = (use_arap ? J * SB_KNN : 0) + (use_rot ? J : 0);
    if (threads == 0) return SB_OK;
    reg_terms_kernel<<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        ed_points, ed_knn, beta, J, lam_arap, lam_rot, use_arap, use_rot, M, g, loss_arap_rot);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_lm_damp(const void* state, double* A, int lda, int n, void* stream) {
    if (!state || !A || n <= 0) return SB_ERR_ARG;
    lm_damp_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const LMState*)state, A, lda, n);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_lm_step(void* state, const int* info, double* beta, const double* delta, int n, const int* node_pos,
               void* stream) {
    if (!state || !beta || !delta || n <= 0) return SB_ERR_ARG;
    lm_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((LMState*)state, info, beta, delta, n,
                                                                    node_pos);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_lm_decide(void* state, const double* partials, int n_partials, double* loss_arap_rot, double* beta,
                 double* best, int n, void* stream) {
    if (!state || !partials || !loss_arap_rot || !beta || !best) return SB_ERR_ARG;
    RegLossArgs rg{nullptr, nullptr, 0, 0.0, 0.0, 0, 0};
    lm_decide_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((LMState*)state, partials, n_partials, loss_arap_rot,
                                                         beta, best, n, rg);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

int sb_lm_decide_reg(void* state, const double* partials, int n_partials, const double* ed_points, const int* ed_knn,
                     int J, double lam_arap, double lam_rot, int use_arap, int use_rot, double* beta, double* best,
                     int n, void* stream) {
    if (!state || !partials || !ed_points || !ed_knn || !beta || !best || J <= 0) return SB_ERR_ARG;
    RegLossArgs rg{ed_points, ed_knn, J, lam_arap, lam_rot, use_arap, use_rot};
    lm_decide_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((LMState*)state, partials, n_partials, nullptr, beta, best,
                                                         n, rg);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
