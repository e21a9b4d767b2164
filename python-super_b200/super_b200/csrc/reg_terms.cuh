// ARAP + Rot regularisers of the LM solver (/root/reference/super/loss.py:403-499): residuals, analytic Jacobian rows and
// their normal-equation contributions, as a per-item device function shared by reg_terms_kernel (lm.cu, its own launch)
// and the fused J^T J pass of the frame loop (data_term.cu: trailing blocks of that launch).
#pragma once
#include "common.cuh"

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

struct RegArgs {
    const double* ed_points; const int* ed_knn; const double* beta; int J; double lam_arap, lam_rot; int use_arap, use_rot;
};

// item tid < J*K (when use_arap): ARAP pair (node j, neighbour slot k); the next J items: Rot of node j.  With
// assemble == false only the squared residuals are returned (la, lr).
__device__ __forceinline__ void reg_terms_item(const RegArgs& ra, int tid, const MatView& M, bool assemble, double& la,
                                               double& lr) {
    const double* __restrict__ ed_points = ra.ed_points;
    const int* __restrict__ ed_knn = ra.ed_knn;
    const double* __restrict__ beta = ra.beta;
    const int J = ra.J;
    const double lam_arap = ra.lam_arap, lam_rot = ra.lam_rot;
    const int n_arap = ra.use_arap ? J * SB_KNN : 0;
    la = 0.0; lr = 0.0;
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
        if (assemble) {
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
                M.add_g(bn0 + a, -gq);
            }
            for (int c = 0; c < 3; ++c) {
                M.add(bn0 + 4 + c, bn0 + 4 + c, l2);
                M.add(bj0 + 4 + c, bj0 + 4 + c, l2);
                add_lower(M, bj0 + 4 + c, bn0 + 4 + c, -l2);
                M.add_g(bn0 + 4 + c, -l * r[c]);
                M.add_g(bj0 + 4 + c, l * r[c]);
            }
        }
    } else if (ra.use_rot && tid < n_arap + J) {
        // RotLoss in float32 like the reference (loss.py:487-497)
        const int j = tid - n_arap;
        const float lam = (float)lam_rot;
        float q[4];
        for (int a = 0; a < 4; ++a) q[a] = (float)beta[7 * j + a];
        const float s = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
        const float r = lam * (1.f - s);
        lr = (double)(r * r);
        if (assemble) {
            float jv[4];
            for (int a = 0; a < 4; ++a) jv[a] = -lam * 2.f * q[a];
            const int pj = 7 * M.pos(j);
            for (int a = 0; a < 4; ++a) {
                for (int b = 0; b <= a; ++b) M.add(pj + a, pj + b, (double)(jv[a] * jv[b]));
                M.add_g(pj + a, -(double)(jv[a] * r));
            }
        }
    }
}

}  // namespace
