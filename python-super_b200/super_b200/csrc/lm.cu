// LM solver pieces other than the data term: ARAP + Rot regularisers (residual, analytic Jacobian,
// normal-equation assembly), additive damping, step application and the device-resident
// accept/reject controller.  Restates /root/reference/super/loss.py:403-499 and
// /root/reference/super/LM.py:81-122 without any host synchronisation: u, minimal_loss, the
// failure flag and the per-iteration trace live in a small device struct.
#include "common.cuh"
#include "lm_state.cuh"
#include "reg_terms.cuh"
#include "super_b200.h"

namespace {

// One thread per (node j, neighbour slot k) for ARAP; one thread per node for Rot (threads >= J*K).
// With A == nullptr only the loss partials are produced.
__global__ void reg_terms_kernel(RegArgs a, MatView M, double* __restrict__ loss_arap_rot /* [2] */) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double la, lr;
    reg_terms_item(a, tid, M, M.A != nullptr, la, lr);
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
        st->u = u; st->v = v; st->minimal_loss = minimal_loss; st->iter = 0; st->failed = 0; st->ticket = 0; st->sel = 0; st->last_accept = 0;
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
                 int* band_overflow, double* g, double* loss_arap_rot, int fx_shift, int fx_gshift, void* stream) {
    if (!ed_points || !ed_knn || !beta || J <= 0) return SB_ERR_ARG;
    if (A && (!g || (bw < 0 ? lda < 7 * J : (lda < bw + 1 || !band_overflow)))) return SB_ERR_ARG;
    if (A && fx_shift >= 0 && (!band_overflow || fx_gshift < 0 || fx_shift > 60 || fx_gshift > 60)) return SB_ERR_ARG;
    MatView M;
    M.A = A; M.lda = lda; M.bw = bw; M.node_pos = node_pos; M.overflow = band_overflow; M.g = g;
    M.set_shift(fx_shift, fx_gshift);
    const int threads = (use_arap ? J * SB_KNN : 0) + (use_rot ? J : 0);
    if (threads == 0) return SB_OK;
    RegArgs ra{ed_points, ed_knn, beta, J, lam_arap, lam_rot, use_arap, use_rot};
    reg_terms_kernel<<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ra, M, loss_arap_rot);
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
