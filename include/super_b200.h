/* super_b200.h -- C ABI of libsuper_b200.so: hand-written sm_100a CUDA kernels for the per-frame
 * embedded-deformation (ED) tracking loop of SuPer (reference: ucsdarclab/Python-SuPer).
 *
 * The reference has no FFI layer (it is pure Python/PyTorch); the seam is Python method
 * calls.  Each entry point below replaces the reference code cited beside it and is what a ctypes
 * binding inside the reference's own modules would call (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every function returns int: 0 ok, 1 invalid argument, 2 CUDA launch error, 3 workspace mismatch;
 *   - pointers are DEVICE pointers borrowed for the call unless marked host; nothing is allocated,
 *     retained or freed; no host synchronisation; no global mutable state;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - row counts come as a capacity `n_cap` (grid sizing, known to the host) plus an optional device
 *     counter `n_dev` (actual count = min(n_cap, *n_dev)); pass NULL to use n_cap;
 *   - layouts are the reference's except integer indices, which are int32 here
 *     (reference: int64), and the new-frame maps, which are dense float4 images (x,y,z,valid) instead
 *     of compact (Nv,3) f64 arrays + index_map (their values are float32-exact in the reference).
 */
#ifndef SUPER_B200_H
#define SUPER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

int sb_version(void);

/* ---- kNN / weights / warp -------------------------------------------------------------------- */

/* find_knn -> pytorch3d knn_points: /root/reference/utils/utils.py:212-220.  K nearest rows of `ref`
 * (nref x dim, dim 2|3) for each query; out_dist = sqrt(d2) ascending (nq,K) f64, out_idx (nq,K) i32.
 * d2 = ((dx*dx + dy*dy) + dz*dz), ties -> lower index. */
int sb_knn(const double* query, int nq_cap, const int* nq_dev, const double* ref, int nref, int dim, int K,
           double* out_dist, int* out_idx, void* stream);

/* softmax_k(exp(-d_k/r_k)) and the "no node within its radius" test:
 * /root/reference/super/nodes.py:164-167 (radius_mode 1: r = radii[i]) and :179-191 (mode 0: r = radii[idx]).
 * stable[i] is cleared (never set) when no neighbour satisfies d_k <= r_k; may be NULL. */
int sb_knn_weights(const double* dist, const int* idx, int n_cap, const int* n_dev, const double* radii,
                   int radius_mode, double* w, unsigned char* stable, void* stream);

/* Weights of existing surfels from current positions and old indices: /root/reference/super/nodes.py:480-484 */
int sb_reweight(const double* points, const int* idx, int n_cap, const int* n_dev, const double* ed_points,
                const double* radii, double* w, void* stream);

/* Surfels.update (LM form): /root/reference/super/nodes.py:193-223 with Trans_points/transformQuatT
 * (/root/reference/super/utils.py:17-57).  In place on points, norms, ed_points, ed_norms. */
int sb_warp_update(double* points, double* norms, const int* idx, const double* w, int n_cap, const int* n_dev,
                   double* ed_points, double* ed_norms, const double* beta, int J, void* stream);

/* ---- LM data term ------------------------------------------------------------------------------ */

/* Number of per-block partial sums sb_data_term_loss writes for a given capacity. */
int sb_data_loss_blocks(int n_cap);

/* DataLoss.forward(grad=True) + LossTool.prepare_jtj_jtl: /root/reference/super/loss.py:200-205,222-288.
 * Accumulates J^T J into the LOWER triangle of dense row-major A (lda >= 7J) and -J^T r into g (7J);
 * both must be zeroed (or hold the other terms) by the caller.  `order` (n,) optional kNN-tuple-sorted
 * surfel ids.  intr = host double[4] {fx,fy,cx,cy}.  loss_cur (optional) accumulates sum r^2 at beta. */
int sb_data_term_jtj(const double* points, const int* knn_idx, const double* knn_w, const int* order, int n_cap,
                     const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                     const float* nmap, int H, int W, const double* intr, double lambda, double* A, int lda,
                     double* g, double* loss_cur, void* stream);

/* DataLoss.forward(grad=False): /root/reference/super/loss.py:222-248,289-290.  partials[b] = sum of r^2
 * over the surfels of block b (n_partials == sb_data_loss_blocks(n_cap)); deterministic. */
int sb_data_term_loss(const double* points, const int* knn_idx, const double* knn_w, int n_cap,
                      const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                      const float* nmap, int H, int W, const double* intr, double lambda, double* partials,
                      int n_partials, void* stream);

/* Per-surfel rows of the same computation (parity tests, drop-in DataLoss face): matched (n,) u8,
 * corners (n,4) i32 [floor v, ceil v, floor u, ceil u], r (n,) f64, jrow (n,28) f64; any of the last
 * three may be NULL. */
int sb_data_term_rows(const double* points, const int* knn_idx, const double* knn_w, int n_cap, const int* n_dev,
                      const double* ed_points, const double* beta, int J, const float* vmap, const float* nmap,
                      int H, int W, const double* intr, double lambda, unsigned char* matched, int* corners,
                      double* r, double* jrow, void* stream);

/* ---- LM regularisers, damping, controller ------------------------------------------------------ */

/* Size of the device-resident controller state and byte offsets of
 * {u, minimal_loss, iter, failed, loss[64], loss_terms[64][3], accept[64], u_trace[64]} (host int[8]). */
int sb_lm_state_bytes(void);
int sb_lm_state_offsets(int* out);

/* LM_Solver.LM prologue: beta = best = identity, u, v, minimal_loss: /root/reference/super/LM.py:81-92 */
int sb_lm_begin(void* state, double* beta, double* best, int J, double u, double v, double minimal_loss,
                void* stream);

/* ARAPLoss + RotLoss (Rot in float32 like the reference): /root/reference/super/loss.py:403-499.
 * With A != NULL adds J^T J (lower) and -J^T r; always adds sum r^2 to loss_arap_rot[0..1] if non-NULL. */
int sb_reg_terms(const double* ed_points, const int* ed_knn, const double* beta, int J, double lam_arap,
                 double lam_rot, int use_arap, int use_rot, double* A, int lda, double* g, double* loss_arap_rot,
                 void* stream);

/* jtj[diag] += u: /root/reference/super/LM.py:97 */
int sb_lm_damp(const void* state, double* A, int lda, int n, void* stream);

/* beta += delta unless *info != 0 (failed factorisation -> loop stops): /root/reference/super/LM.py:99-105 */
int sb_lm_step(void* state, const int* info, double* beta, const double* delta, int n, void* stream);

/* loss < minimal_loss ? accept : reject with u /= v | u *= v: /root/reference/super/LM.py:107-117 */
int sb_lm_decide(void* state, const double* partials, int n_partials, double* loss_arap_rot, double* beta,
                 double* best, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUPER_B200_H */
