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


/* ---- plain-C views of device state (host structs holding device pointers) ------------------------- */

/* Surfel arrays, reference layouts; capacity-based: rows [0, min(cap,*n_dev)) are live. */
typedef struct SbSurfels {
    double* points;         /* (cap,3) f64 */
    double* norms;          /* (cap,3) f64 */
    float* colors;          /* (cap,3) f32 */
    float* confs;           /* (cap,)  f32 */
    double* radii;          /* (cap,)  f64 */
    float* time_stamp;      /* (cap,)  f32 */
    int* knn_idx;           /* (cap,4) i32  (reference: i64) */
    double* knn_w;          /* (cap,4) f64 */
    float* projdata;        /* (cap,2) f32  [u,v] */
    unsigned char* stable;  /* (cap,)  bool */
    int cap;
    int* n_dev;             /* device row counter */
    /* Semantic-SuPer state (NULL / 0 when absent): /root/reference/super/nodes.py:58-66 */
    int* seg;               /* (cap,)   i32 class (reference: i64) */
    double* seg_conf;       /* (cap,C)  f64 class probabilities */
    int n_classes;          /* C <= 8 */
} SbSurfels;

/* One preprocessed input frame as dense per-pixel images (P = H*W). */
typedef struct SbFrame {
    const float* vmap;      /* (P,4) f32 x,y,z,valid(1/0)   = new_data.points + valid/index_map */
    const float* nmap;      /* (P,4) f32 nx,ny,nz,0         = new_data.norms */
    const double* radii;    /* (P,)  f64                    = new_data.radii scattered */
    const float* confs;     /* (P,)  f32                    = new_data.confs scattered */
    const float* color;     /* (3,P) f32 planar             = inputs[("color",0)] */
    int H, W;
    double fx, fy, cx, cy;  /* K[0,0], K[1,1], K[0,2], K[1,2] (float32 values promoted) */
    const int* seg;         /* (P,)   i32 argmax class per pixel, or NULL     = new_data.seg scattered */
    const double* seg_conf; /* (P,C)  f64 softmax(scores) per pixel, or NULL  = new_data.seg_conf scattered */
} SbFrame;

typedef struct SbFuseParams {
    double th_dist;         /* opt.th_dist 0.1 */
    double th_cos;          /* opt.th_cosine_ang 0.4 */
    float time_now;         /* sfdata.time */
    int disable_merging_new, disable_merging_exist, disable_adding_new;
    int class_gate;         /* merges need equal classes: (hard_seg or data == superv1) and seg present, nodes.py:314-316 */
    int semantic_weights;   /* kNN weights softmax(sqrt(exp(-JSD)) sqrt(exp(-d/r))): bit 0 = existing surfels every frame
                             * (nodes.py:466-484), bit 1 = appended surfels (nodes.py:503-509, not under --hard_seg) */
    const double* ed_seg_conf;  /* (J,C) f64 node class probabilities (semantic_weights) */
    const int* ed_seg;          /* (J,) i32 node classes: new surfels search their nodes inside their own class (--hard_seg, nodes.py:494-497) or NULL */
} SbFuseParams;

/* Everything one tracked frame's LM solve reads and writes (sb_lm_frame).  Device pointers unless marked host. */
typedef struct SbLMFrame {
    /* surfel model: Surfels.points / knn_indices / knn_w (/root/reference/super/nodes.py:36-91) */
    const double* points;       /* (n_cap,3) f64 */
    const int* knn_idx;         /* (n_cap,4) i32 */
    const double* knn_w;        /* (n_cap,4) f64 */
    const int* order;           /* (n_cap,) i32 visiting order (sb_tuple_keys, sorted) or NULL */
    int n_cap;
    const int* n_dev;
    /* ED graph */
    const double* ed_points;    /* (J,3) */
    const int* ed_knn;          /* (J,4) i32 */
    int J;
    /* new frame */
    const float* vmap;          /* (P,4) f32 */
    const float* nmap;          /* (P,4) f32 */
    int H, W;
    double intr[4];             /* fx, fy, cx, cy */
    /* terms: sqrt-free weights as LM_Solver passes them (/root/reference/super/LM.py:18-26) */
    double lam_data, lam_arap, lam_rot;
    int use_arap, use_rot;
    /* controller: LM(u=10, v=7.5, minimal_loss=1e10), num_optimize_iterations (/root/reference/super/LM.py:81-92) */
    int iterations;
    double u, v, minimal_loss;
    void* state;                /* sb_lm_state_bytes() */
    double* beta;               /* (J,7) out: the frame's result */
    double* best;               /* (J,7) scratch */
    double* partials_loss;      /* n_partials_loss >= max(sb_data_loss_blocks(n_cap), 1024) doubles: per-block sums of r^2 */
    int n_partials_loss;
    double* rows;               /* (29, row_stride) f64 scratch: Jacobian rows of the last evaluation pass */
    unsigned long long* keys;   /* (row_stride,) u64 scratch: node-set key per slot */
    int row_stride;             /* >= n_cap, multiple of 32 */
    double* rec_vals;           /* (rec_cap, 436) f64 scratch: Gram records of the pass (one per node-set run and warp) */
    unsigned long long* rec_keys;   /* (rec_cap,) u64 */
    int* rec_count;             /* device counter */
    int rec_cap;                /* suggested: 2 * ceil(n_cap / 32) + 4096; beyond it warps add their accumulators directly */
    /* normal equations, band storage in the solver's node order */
    int n, bw, ldab;            /* n = 7J, half bandwidth, row stride (>= bw+1) */
    const int* node_pos;        /* node id -> solver position, or NULL */
    const int* pos_node;        /* solver position -> node id, or NULL */
    long long* fx_store[2];     /* two fixed-point stores of n*ldab + n int64 each (AB | g) */
    int fx_shift, fx_gshift;    /* entries are multiples of 2^-fx_shift (AB), 2^-fx_gshift (g) */
    double* AB;                 /* (n,ldab) f64 work band the solver factors */
    double* g;                  /* (n,) f64 work right-hand side / solution */
    int* band_overflow;         /* bit 0: entry outside the band, bit 1: addend outside the fixed-point range */
    double* dinv;               /* (n,) scratch */
    int* info;                  /* factorisation status */
    void* solver_ws;            /* sb_band4_workspace_bytes(n, bw, ldab) */
    long long solver_ws_bytes;
    int n_ctas;
    /* measurement (optional, HOST array of cudaEvent_t from sb_event_create): events [2k], [2k+1] are recorded on `stream`
     * right before / after the k-th J^T J pass of the frame, for as many passes as the array covers */
    void* const* jtj_events;
    int n_jtj_events;
    void* const* solve_events;  /* the same around the k-th linear solve (from_fixed + the five solver kernels) */
    int n_solve_events;
    void* const* stage_events;  /* timeline: event 0 before the first launch, then one event after every stage of the frame,
                                 * in issue order -- lm_begin, [eval, gram, scatter] and per iteration [solve, eval | loss,
                                 * gram, scatter] -- for as many as the array covers */
    int n_stage_events;
    void** graph_cache;         /* host: address of a caller-owned opaque handle (initially NULL) or NULL.  With a handle and no
                                 * event arrays, the frame's launch sequence is stream-captured (on a private stream), the
                                 * caller's instantiated graph is updated in place (cudaGraphExecUpdate: same topology every
                                 * frame, only parameters change) and launched on `stream`: the ~85 dependent launches of a
                                 * frame then follow each other ~2.5 us closer than stream launches do.  The first call with a
                                 * fresh handle launches directly.  Free the handle with sb_lm_graph_destroy. */
} SbLMFrame;

int sb_version(void);

/* Releases what sb_lm_frame keeps behind SbLMFrame.graph_cache (graph exec + capture stream); *cache becomes NULL. */
int sb_lm_graph_destroy(void** cache);

/* The same mechanism for any fixed sequence of calls of this library (the per-frame tail: warp/update, fusion, compaction,
 * visiting order): between _begin and _end pass *use_stream to the calls instead of `stream`; _end replays what was captured
 * as one graph on `stream` (updated in place from frame to frame).  The first use of a handle runs directly.  Only calls of
 * this library (and nothing that synchronises) may be issued in between; abort_scope != 0 drops the capture. */
int sb_graph_scope_begin(void** cache, void* stream, void** use_stream);
int sb_graph_scope_end(void** cache, void* stream, int abort_scope);
/* A non-blocking CUDA stream (one that does not synchronise implicitly with the legacy default stream): the side stream of the
 * input prefetch (SuPer.prefetch). */
int sb_stream_create(void** out);
int sb_stream_destroy(void* stream);
/* device-to-device copy of n ints on the stream (a memcpy node inside a scope) */
int sb_copy_i32(int* dst, const int* src, int n, void* stream);

/* ---- kNN / weights / warp -------------------------------------------------------------------- */

/* find_knn -> pytorch3d knn_points: /root/reference/utils/utils.py:212-220.  K nearest rows of `ref`
 * (nref x dim, dim 2|3) for each query; out_dist = sqrt(d2) ascending (nq,K) f64, out_idx (nq,K) i32.
 * d2 = ((dx*dx + dy*dy) + dz*dz), ties -> lower index. */
int sb_knn(const double* query, int nq_cap, const int* nq_dev, const double* ref, int nref, int dim, int K,
           double* out_dist, int* out_idx, void* stream);

/* find_knn with num_classes > 0 (--hard_seg): /root/reference/utils/utils.py:222-242.  Neighbours are searched among the
 * reference points whose class rseg[j] equals the query's class qseg[i]; indices stay global; 1e8 / -1 when the class
 * has fewer than K points.  qseg == rseg == NULL: plain sb_knn. */
int sb_knn_class(const double* query, int nq_cap, const int* nq_dev, const int* qseg, const double* ref, int nref,
                 const int* rseg, int dim, int K, double* out_dist, int* out_idx, void* stream);

/* softmax_k(exp(-d_k/r_k)) and the "no node within its radius" test:
 * /root/reference/super/nodes.py:164-167 (radius_mode 1: r = radii[i]) and :179-191 (mode 0: r = radii[idx]).
 * stable[i] is cleared (never set) when no neighbour satisfies d_k <= r_k; may be NULL. */
int sb_knn_weights(const double* dist, const int* idx, int n_cap, const int* n_dev, const double* radii,
                   int radius_mode, double* w, unsigned char* stable, void* stream);

/* Weights of existing surfels from current positions and old indices: /root/reference/super/nodes.py:480-484 */
int sb_reweight(const double* points, const int* idx, int n_cap, const int* n_dev, const double* ed_points,
                const double* radii, double* w, void* stream);

/* Semantic variant of the weights (/root/reference/super/nodes.py:183-189, power_arg (1/2,1/2)):
 * w = softmax_k( sqrt(exp(-JSD(ed_seg_conf[idx_k], seg_conf_i))) * sqrt(exp(-d_k / r_{idx_k})) ), d from the positions */
int sb_reweight_semantic(const double* points, const int* idx, int n_cap, const int* n_dev, const double* ed_points,
                         const double* radii, const double* ed_seg_conf, const double* seg_conf, int C, double* w,
                         void* stream);

/* Surfels.update (LM form): /root/reference/super/nodes.py:193-223 with Trans_points/transformQuatT
 * (/root/reference/super/utils.py:17-57).  In place on points, norms, ed_points, ed_norms. */
int sb_warp_update(double* points, double* norms, const int* idx, const double* w, int n_cap, const int* n_dev,
                   double* ed_points, double* ed_norms, const double* beta, int J, void* stream);

/* ---- LM data term ------------------------------------------------------------------------------ */

/* Sort keys (i64, non-negative) that order surfels by their 4-tuple of ED nodes; rows >= n get the maximum
 * key.  Sorting them gives the `order` argument of sb_data_term_jtj (the reference has no counterpart: it
 * builds a COO Jacobian and calls torch.sparse.mm, /root/reference/super/loss.py:285-288,200-205). */
int sb_tuple_keys(const int* knn_idx, int n_cap, const int* n_dev, long long* keys, const int* node_pos,
                  int* block_bw, void* stream);

/* The sorted order itself in one call: keys with ceil(log2(J+1)) bits per node id and a stable LSD radix sort over just
 * those 4*bits bits (5 passes at J = 266 instead of the 8 of a 64-bit key).  keys / keys_alt (n_cap u64), rows (n_cap i32)
 * and temp (sb_tuple_order_temp_bytes) are scratch; order (n_cap i32) receives the visiting order. */
long long sb_tuple_order_temp_bytes(int n_cap);
int sb_tuple_order(const int* knn_idx, int n_cap, const int* n_dev, int J, const int* node_pos, int* block_bw,
                   unsigned long long* keys, unsigned long long* keys_alt, int* rows, int* order, void* temp,
                   long long temp_bytes, void* stream);

/* The LM inputs of the surfels gathered into visiting order (row i <- row order[i] of points / knn_idx / knn_w): done once
 * per frame, it makes every data-term pass of the frame coalesced (pass order = NULL with the gathered arrays). */
int sb_gather_sorted(const double* points, const int* knn_idx, const double* knn_w, const int* order, int n_cap,
                     const int* n_dev, double* out_points, int* out_idx, double* out_w, void* stream);

/* Number of per-block partial sums sb_data_term_loss writes for a given capacity. */
int sb_data_loss_blocks(int n_cap);

/* DataLoss.forward(grad=True) + LossTool.prepare_jtj_jtl: /root/reference/super/loss.py:200-205,222-288.
 * Accumulates J^T J into the LOWER triangle of dense row-major A (lda >= 7J) and -J^T r into g (7J);
 * both must be zeroed (or hold the other terms) by the caller.  `order` (n,) optional kNN-tuple-sorted
 * surfel ids.  intr = host double[4] {fx,fy,cx,cy}.  loss_cur (optional) accumulates sum r^2 at beta.
 * Matrix target: bw < 0 -> dense A[row*lda + col]; bw >= 0 -> lower band A[row*lda + col - row + bw] in the
 * node order node_pos (node id -> position, may be NULL); entries outside the band set bit 0 of *band_overflow.
 * fx_shift >= 0: A and g are int64 FIXED-POINT arrays (multiples of 2^-fx_shift, 2^-fx_gshift) added with integer atomics:
 * the sums are independent of the order of arrival (bitwise reproducible); sb_band_from_fixed converts; an addend outside
 * +-2^(62-shift) sets bit 1 of *band_overflow.  fx_shift < 0: f64 atomics (results vary ~1e-9 from run to run). */
int sb_data_term_jtj(const double* points, const int* knn_idx, const double* knn_w, const int* order, int n_cap,
                     const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                     const float* nmap, int H, int W, const double* intr, double lambda, double* A, int lda,
                     int bw, const int* node_pos, int* band_overflow, double* g, double* loss_cur, int fx_shift,
                     int fx_gshift, void* stream);

/* DataLoss.forward(grad=False): /root/reference/super/loss.py:222-248,289-290.  partials[b] = sum of r^2
 * over the surfels of block b (n_partials == sb_data_loss_blocks(n_cap)); deterministic. */
int sb_data_term_loss(const double* points, const int* knn_idx, const double* knn_w, int n_cap,
                      const int* n_dev, const double* ed_points, const double* beta, int J, const float* vmap,
                      const float* nmap, int H, int W, const double* intr, double lambda, double* partials,
                      int n_partials, void* stream);

/* loss-only pass + the accept/reject step of the iteration in ONE launch (sb_data_term_loss followed by sb_lm_decide_reg):
 * /root/reference/super/loss.py:212-246,290 (DataLoss.forward, grad=False), loss.py:433-437,487-497 (ARAP / Rot losses) and
 * /root/reference/super/LM.py:107-117 (accept iff loss < minimal_loss; u /= v or u *= v, revert beta).  The last block to
 * deliver its partial sums them in their fixed order and decides.  beta_in == beta (the trial beta). */
int sb_data_term_loss_decide(const double* points, const int* knn_idx, const double* knn_w, int n_cap,
                             const int* n_dev, const double* ed_points, const double* beta_in, int J, const float* vmap,
                             const float* nmap, int H, int W, const double* intr, double lambda, double* partials,
                             int n_partials, void* state, const int* ed_knn, double lam_arap, double lam_rot,
                             int use_arap, int use_rot, double* beta, double* best, void* stream);
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
                 double lam_rot, int use_arap, int use_rot, double* A, int lda, int bw, const int* node_pos,
                 int* band_overflow, double* g, double* loss_arap_rot, int fx_shift, int fx_gshift, void* stream);

/* int64 fixed-point store (AB | g, as sb_data_term_jtj / sb_reg_terms fill it with fx_shift >= 0) -> the f64 band AB (n,ldab)
 * and right-hand side g (n) the solver takes. */
int sb_band_from_fixed(const long long* store, int n, int ldab, int fx_shift, int fx_gshift, double* AB, double* g,
                       void* stream);

/* LM_Solver.LM, the whole loop of one frame: /root/reference/super/LM.py:81-122 (prepareCostTerm :53-79, Solver :38-51)
 * over DataLoss / ARAPLoss / RotLoss (/root/reference/super/loss.py:207-499), band path.  Enqueues
 * 2 + 8*iterations launches on `stream`; no host synchronisation.  On return (after the stream has run) f->beta holds the
 * result and the controller state the per-iteration trace (sb_lm_state_offsets).  A failed factorisation stops the
 * updates and leaves the last accepted beta (LM.py:99-103).  Bitwise reproducible. */
/* CUDA events for timing launches inside sb_lm_frame on the stream they run on (bench.py's roofline) */
int sb_event_create(void** ev);
int sb_event_destroy(void* ev);
int sb_event_elapsed_ms(void* ev_begin, void* ev_end, float* ms);   /* synchronises on ev_end */
int sb_lm_frame(const SbLMFrame* f, void* stream);

/* jtj[diag] += u: /root/reference/super/LM.py:97 */
int sb_lm_damp(const void* state, double* A, int lda, int n, void* stream);

/* beta += delta unless *info != 0 (failed factorisation -> loop stops): /root/reference/super/LM.py:99-105 */
int sb_lm_step(void* state, const int* info, double* beta, const double* delta, int n, const int* node_pos,
               void* stream);

/* Banded Cholesky solve of (A + u I) x = g replacing torch.linalg.cholesky + cholesky_solve:
 * /root/reference/super/LM.py:38-51,97-100.  AB (n, ldab) lower band row-major (overwritten with updated, unfactored tiles),
 * g (n) rhs in / solution out, u device scalar (NULL = 0), dinv (n) scratch, *info set to 1 on a non-positive pivot.
 *
 * sb_band_solve3: one cooperative grid of n_ctas >= 3 CTAs (clamped to the SM count) -- a pivot-chain CTA, a substitution
 * CTA and update CTAs synchronised by release/acquire counters; every 32x32 triangular solve is an FP64 tensor-core
 * product with the explicit inverse of the diagonal factor; push-style back substitution.  L and the inverses live in the
 * workspace (>= sb_band3_workspace_bytes(n, bw)).  sb_band3_update_role: 1 = one trailing tile per CTA and panel,
 * 2 = fixed tile owners (bands too wide for role 1 on n_ctas CTAs). */
long long sb_band3_workspace_bytes(int n, int bw);
int sb_band3_fits(int n, int bw);
int sb_band3_update_role(int n, int bw, int n_ctas);
int sb_band_solve3(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                   void* workspace, long long ws_bytes, int n_ctas, void* stream);
/* v4 = v3's pipeline run from BOTH ends of the band at once (two partial factorisations: top-down on the matrix, bottom-up on a
 * reversed copy), the middle block (>= bw rows) solved last, two concurrent back substitutions outwards: the pivot chain --
 * the solve's critical path -- is n/2 + bw/2 long instead of n.  Falls back to v3 for systems too small to split. */
long long sb_band4_workspace_bytes(int n, int bw, int ldab);
int sb_band_solve4(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                   void* workspace, long long ws_bytes, int n_ctas, void* stream);
/* the same solve with the LM step (sb_lm_step's contract, /root/reference/super/LM.py:99-105) folded into its last kernel:
 * beta[7 pos_node[p] + c] += x[7 p + c] (pos_node: solver position -> node id, NULL = identity) unless *info != 0 or
 * *lm_failed != 0; a failed factorisation sets *lm_failed.  g still receives x. */
int sb_band_solve4_step(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                        void* workspace, long long ws_bytes, int n_ctas, int* lm_failed, double* beta,
                        const int* pos_node, void* stream);
/* sb_band_solve4_step with the system taken from one of two int64 fixed-point stores (AB | g; *sel names the current one,
 * NULL = store 0): the conversion to the f64 work band AB / g rides in the solve's first kernel, which also clears the
 * other store when zero_other != 0 (the frame loop's next assembly target). */
int sb_band_solve4_step_fx(const long long* fx_store0, const long long* fx_store1, const int* sel, int fx_shift,
                           int fx_gshift, int zero_other, double* AB, int ldab, int n, int bw, double* g, const double* u,
                           double* dinv, int* info, void* workspace, long long ws_bytes, int n_ctas, int* lm_failed,
                           double* beta, const int* pos_node, void* stream);
int sb_band_max_bw(void);     /* widest half bandwidth the band solver takes */

#ifdef SB_DEBUG_EXPORTS
/* Timing experiments only (scripts/one_band4.py, scripts/probe_band3.py): built with SB_DEBUG_EXPORTS=1 python -m
 * super_b200.build; NOT part of the product library.  Process-global flags: 1 no trailing update, 2 no back substitution,
 * 4 cycle counters, 64 force update role 2, 256 stage events. */
int sb_band3_debug(int flags);
long long sb_band3_prof_offset(int n, int bw);
int sb_band4_stage_ms(float* out5); /* ms of the 5 stages of the last sb_band_solve4 (flag 256) */
#endif

/* loss < minimal_loss ? accept : reject with u /= v | u *= v: /root/reference/super/LM.py:107-117 */
int sb_lm_decide(void* state, const double* partials, int n_partials, double* loss_arap_rot, double* beta,
                 double* best, int n, void* stream);
/* the same with the ARAP / Rot losses of the stepped beta (loss.py:433-437,487-497) evaluated inside the launch instead of
 * by a loss-only sb_reg_terms in front of it: one launch less per LM iteration */
int sb_lm_decide_reg(void* state, const double* partials, int n_partials, const double* ed_points, const int* ed_knn,
                     int J, double lam_arap, double lam_rot, int use_arap, int use_rot, double* beta, double* best,
                     int n, void* stream);


/* ---- autograd optimiser (GraphFit) as fused loss + analytic-gradient kernels ------------------------------
 * /root/reference/super/deform_mesh.py:25-379 with the autograd forms of the terms in /root/reference/super/loss.py
 * (:9-100 bilinear_sample, :293-401 DataLoss.autograd_forward, :458-473 ARAP, :502-505 Rot).
 * dv (J+1,7) f64 deform_verts, row J = global transform.  grad (J+1,7) and acc[6] = {point_plane, arap, rot, face,
 * morph_sum, morph_count} are ACCUMULATED (zero them before the first term; sb_gf_step re-zeroes them). */

/* point-to-plane term; seg_mode 0 none | 1 soft (exp(-0.1 JSD)) | 2 hard (class equality); trg_seg_conf (P,C) dense map */
int sb_gf_data(const double* points, const int* knn_idx, const double* knn_w, const unsigned char* stable, int n_cap,
               const int* n_dev, const double* ed_points, int J, const double* dv, const float* vmap, const float* nmap,
               int H, int W, const double* intr, double weight, int seg_mode, int C, const int* sf_seg,
               const double* sf_seg_conf, const double* trg_seg_conf, double* grad, double* acc, void* stream);

/* boundary-morph term: scores (C,H,W) f64 raw class scores of the new frame, edge_pts (E,2) f64 (x,y) of all classes
 * concatenated, edge_off (C+1) i32.  grad_morph (J+1,7) holds the UN-normalised gradient (sb_gf_step divides by count) */
int sb_gf_morph(const double* points, const int* knn_idx, const double* knn_w, const unsigned char* stable, int n_cap,
                const int* n_dev, const double* ed_points, int J, const double* dv, int H, int W, const double* intr,
                const double* scores, int C, const int* sf_seg, const double* edge_pts, const int* edge_off,
                double* grad_morph, double* acc, void* stream);

/* ARAP (knn_w weighted) + Rot (J+1 rows) + Face terms on the graph; triangles (3,F) i32 */
int sb_gf_reg(const double* ed_points, const int* ed_knn, const double* ed_knn_w, int J, const int* triangles,
              const double* areas, int F, const double* dv, double w_arap, double w_rot, double w_face, int use_arap,
              int use_rot, int use_face, double* grad, double* acc, void* stream);

/* grad[J] /= J; optimizer 0 = SGD(momentum 0.9) | 1 = Adam(0.9, 0.999, 1e-8); state 2*(J+1)*7 doubles; trace row
 * `iter` (8 doubles): total, face, arap, rot, point_plane, bn_morph, morph_count, 0; grad_out (optional) = consumed grad */
int sb_gf_step(double* dv, double* grad, double* grad_morph, double* acc, double w_morph, int use_morph, int J,
               int optimizer, double lr, int iter, double* state, double* trace, double* grad_out, void* stream);

/* the global-row part of Surfels.update (/root/reference/super/nodes.py:204-205,211-212,219-222), after sb_warp_update */
int sb_gf_global_update(double* points, double* norms, int n_cap, const int* n_dev, double* ed_points, double* ed_norms,
                        int J, const double* global_row, void* stream);


/* ---- helpers the reference exports by name (Python face: super/utils.py, utils/utils.py, super/loss.py) ------------- */

/* get_skew: /root/reference/super/utils.py:4-14.  a (n,3) -> out (n,3,3) = [a]x  (the layout the reference produces for
 * its 3-D inputs) */
int sb_get_skew(const double* a, long long n, double* out, void* stream);
/* transformQuatT: /root/reference/super/utils.py:41-71.  v (n,3), beta (n,beta_dim) with beta_dim 4 (rotation only) or
 * 7 (q; b), q NOT normalised -> tv (n,3); jac (n,3,4) = d tv / d q when non-NULL (grad=True) */
int sb_transform_quat(const double* v, const double* beta, long long n, int beta_dim, double* tv, double* jac, void* stream);
/* Trans_points: /root/reference/super/utils.py:17-38.  d, g (n,K,3), beta (n,K,7), w (n,K) or NULL (weights 1) ->
 * out (n,3) = sum_k w_k [T(q_k,b_k) d_k + g_k]; jac (n,K,3,4) = w_k d[R(q_k) d_k]/dq_k when non-NULL */
int sb_trans_points(const double* d, const double* g, const double* beta, const double* w, long long n, int K, double* out,
                    double* jac, void* stream);
/* pcd2depth: /root/reference/utils/utils.py:161-184.  pcd (n,3) f64, intr = host {fx,fy,cx,cy} -> unrounded (v,u) f64
 * and/or rounded (v,u) i64 (half to even), coords = round(v) W + round(u) i64, valid = margin <= v < H-1-margin and
 * margin <= u < W-1-margin on the ROUNDED values (u8).  Either pair of (v,u) outputs may be NULL. */
int sb_pcd2depth(const double* pcd, long long n, const double* intr, int H, int W, int valid_margin, double* v_float,
                 double* u_float, long long* v_round, long long* u_round, long long* coords, unsigned char* valid,
                 void* stream);
/* KLD / JSD over the last axis: /root/reference/utils/utils.py:244-254.  P, Q (n,C) f64, C <= 8 -> out (n,) */
int sb_kld_jsd(const double* P, const double* Q, long long n, int C, double eps, int jsd, double* out, void* stream);
/* ARAPLoss / RotLoss .forward(grad=False): /root/reference/super/loss.py:428-437,452-455 and :487-490,497-499.
 * arap_r2 (J*K*3,) f64 squared residuals in (node, neighbour, xyz) order, rot_r2 (J,) f32; either may be NULL */
int sb_reg_residuals(const double* ed_points, const int* ed_knn, const double* beta, int J, double lam_arap, double lam_rot,
                     double* arap_r2, float* rot_r2, void* stream);

/* ---- producer-side pieces next to the path ------------------------------------------------------------------------ */

/* torch_dilate on one channel: /root/reference/utils/utils.py:152-157 (box filter > 0, conv2d padding='same': for an even
 * kernel the extra tap is on the bottom / right).  in, out (H,W) u8, out != in; invert_in / invert_out apply ~ to the
 * operand / result, so that the reference's open-then-dilate of the invalid mask (utils/data_loader.py:394-397) is two calls */
int sb_dilate_box(const unsigned char* in, int H, int W, int kernel, int invert_in, int invert_out, unsigned char* out,
                  void* stream);
/* SSIM depth confidence (run_semantic_super.py's default): /root/reference/utils/data_loader.py:360-372,477-479 with
 * Project3D (/root/reference/depth/monodepth2/layers.py:173-193), F.grid_sample defaults and
 * skimage.metrics.structural_similarity(channel_axis=0, full=True) defaults (7x7 uniform window, sample covariance,
 * K1 0.01, K2 0.03; data_range as given -- skimage, absent here, derives 2.0 from a float image: parity UNPINNED).
 * depth (H,W) f32, color (3,H,W) f32, inv_K3x3 host float[9], KT3x4 host float[12] = (K @ stereo_T)[:3], warp_scratch
 * (3,H,W) f32.  In place: confs <- 0.5 confs + 0.5 sigmoid(mean_c SSIM).  ssim_out (H,W) optional. */
int sb_ssim_conf(const float* depth, const float* color, const float* inv_K3x3, const float* KT3x4, int H, int W,
                 double data_range, float* warp_scratch, float* confs, float* ssim_out, void* stream);
/* Surfel splat renderer in pulsar's role: /root/reference/renderer/renderer.py:50-78 as called by
 * /root/reference/super/nodes.py:630-642.  Every (masked-in) surfel is a sphere of radius rad; the nearest sphere along a
 * pixel's ray wins (pulsar with gamma -> 0); img (H,W,3) f32 = its colour, bg3 (host float[3]) elsewhere; optional depth
 * (H,W) f32 of the hit and index (H,W) i32 of the winning surfel (-1 = background).  zbuf: H*W u64 scratch. */
int sb_render_splats(const double* points, const float* colors, const unsigned char* mask, int n_cap, const int* n_dev,
                     const double* intr, int H, int W, double rad, const float* bg3, unsigned long long* zbuf, float* img,
                     float* depth, int* index, void* stream);

/* ---- ED graph (once per sequence / at every re-initialisation) -------------------------------------------- */

/* init_graph + DirectDeformGraph.init_ED_nodes, grid_mesh branch: /root/reference/super/graph_encoder.py:11-67,128-167.
 * Nodes = anchors (k*step, l*step) on valid pixels of the frame's dense maps, ids row-major; per cell edges (a,r) (a,rd)
 * (a,d) (r,d) and triangles (a,r,rd) (a,rd,d); prune_classes (--hard_seg with --mesh_face) drops those across classes;
 * radii = mean incident edge length (nodes without edges: mean of the others); areas = 0.5 sqrt(|cross|^2 + 1e-13).
 * Capacity G = sb_graph_anchors(H, W, step): workspace 3G ints; points, norms (G,3) f64; anchor_uv (G,2) i32 [u,v];
 * seg (G,) i32 + ed_seg_conf (G,C) f64 when seg_conf (P,C) is given; edges (4G,2) i32; faces (2G,3) i32; edge_lens (4G);
 * radii (G); areas (2G); node_pos (G,) i32 = position of each node in the band solver's order (sorted along the longer
 * image axis; no reference counterpart); counts[3] = {J, E, F} (device).  One launch, fixed summation orders. */
int sb_graph_anchors(int H, int W, int step);
int sb_graph_build(const float* vmap, const float* nmap, const double* seg_conf, int C, int H, int W, int step,
                   int prune_classes, int* workspace, double* points, double* norms, int* anchor_uv, int* seg,
                   double* ed_seg_conf, int* edges, int* faces, double* edge_lens, double* radii, double* areas,
                   int* node_pos, int* counts, void* stream);
/* *out_max = max(*out_max, max_jk |node_pos[j] - node_pos[knn[j,k]]|): block half-bandwidth the ARAP pairs need */
int sb_graph_pair_span(const int* knn, const int* node_pos, int J, int K, int* out_max, void* stream);

/* ---- per-frame producer ---------------------------------------------------------------------------- */

/* depth_preprocessing for --load_depth inputs: /root/reference/utils/data_loader.py:333-523 (getN :532-583,
 * BackprojectDepth /root/reference/depth/monodepth2/layers.py:139-167).  depth (H,W) f32, color (3,H,W) f32,
 * inval (H,W) u8 optional extra invalid mask, inv_K3x3 = host float[9] (inv_K[:3,:3]), fx = K[0,0].
 * superv2 != 0 selects that dataset's validity rules.  pcd_scratch (P,4) f32.  Outputs are the SbFrame
 * images; valid_i32 (P,) optional 0/1 copy of the validity for scans. */
int sb_preprocess(const float* depth, const float* color, const unsigned char* inval, const float* inv_K3x3,
                  float fx, float divterm, int superv2, int H, int W, float* pcd_scratch, float* vmap, float* nmap,
                  double* radii, float* confs, int* valid_i32, void* stream);

/* Per-pixel segmentation maps of a frame: seg = argmax_c scores, seg_conf = softmax_c scores
 * (/root/reference/utils/data_loader.py:229-231,457).  scores (C,H,W) f64 -> seg (P,) i32, seg_conf (P,C) f64 */
int sb_seg_maps(const double* scores, int C, int H, int W, int* seg, double* seg_conf, void* stream);

/* ---- fusion / compaction ------------------------------------------------------------------------- */

/* Bytes of scratch sb_fuse / sb_compact need for an H x W image and surfel capacity `cap`. */
long long sb_fuse_workspace_bytes(int H, int W, int cap);

/* Surfels.fuseInputData: /root/reference/super/nodes.py:270-541 (merge_data :301-355).  In place on *sf;
 * appended rows go to [n, n_out).  track_id (n_track,) i64 device array or NULL.  *overflow is set to 1
 * if the capacity was too small (rows dropped).  Confidence ties on one pixel -> lower surfel index. */
int sb_fuse(const SbSurfels* sf, const SbFrame* frame, const double* ed_points, const double* ed_radii, int J,
            const SbFuseParams* params, long long* track_id, int n_track, int* n_out, int* overflow,
            void* workspace, long long ws_bytes, void* stream);

/* Surfels.prepareStableIndexNSwapAllModel (state part): /root/reference/super/nodes.py:543-589 plus the
 * projdata of :540-541.  Stable, recently-updated rows of *src are copied to *dst in order; *dst->n_dev
 * receives the new count; track ids are remapped. */
int sb_compact(const SbSurfels* src, const SbSurfels* dst, const SbFrame* frame, double time_now,
               int th_time_steps, int disable_removing, long long* track_id, int n_track, void* workspace,
               long long ws_bytes, void* stream);

/* Tracked-point bookkeeping after compaction: update_track_pts + init_track_pts as prepareStableIndexNSwapAllModel
 * calls them (/root/reference/super/nodes.py:225-265,594-599) for a frame that has labels.  gt (T,3) i32 [x,y,valid],
 * track_id (T,) i64 in/out (-1 not started, -2 lost), out (T,3) f32 = track_rsts[filename] ([u, v, 1] per point). */
int sb_track_points(const double* points, const unsigned char* stable, const float* projdata, int n_cap, const int* n_dev,
                    const float* vmap, int H, int W, const int* gt, int T, long long* track_id, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUPER_B200_H */
