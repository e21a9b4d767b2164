// The whole LM loop of one tracked frame behind ONE C call: LM_Solver.LM (/root/reference/super/LM.py:81-122) with its
// prepareCostTerm / Solver steps (LM.py:38-79) and the three terms of /root/reference/super/loss.py, on the band path.
//
// What is restructured against the reference (results unchanged):
//   * the loss of a trial beta falls out of the J^T J pass AT that beta (sum r^2 is entry (28,28) of the Gram panels), and
//     if the step is accepted that pass already IS the next iteration's normal equations.  So the loop is
//         assemble(beta_0) ; repeat { solve -> beta' = beta + delta ; assemble(beta') -> loss' ; accept / reject }
//     with one J^T J pass per iteration and no separate loss-only pass, except after the last solve where only the loss
//     is needed.  The reference evaluates J^T J at the current beta and the loss at the trial beta in two passes per
//     iteration (LM.py:93-107); after a rejected step it rebuilds the same J^T J again, here the kept copy is reused.
//   * the normal equations are accumulated in 64-bit FIXED POINT (common.cuh MatView): integer atomics commute, so the
//     assembled system is bitwise independent of the order in which warps arrive.  Two stores alternate: the current
//     system and the one being assembled at the trial beta; LMState.sel says which is which, on the device.
//   * band_from_fixed_kernel turns the current store into the f64 band the solver factors (the solver overwrites its
//     input, so a copy is needed anyway) and clears the other store for the next assembly: no memset nodes, no side stream.
//   * the data term runs as TWO launches: a high-occupancy evaluation pass that writes the Jacobian rows (and takes the
//     LM decision in its last block) and a Gram pass that reads them back; after a reject the Gram pass returns at once.
// One iteration = 8 launches: band_reverse (+ store -> band), band_chol3_dual, band_combine, band_chol3, band_backsub4 (+ step),
// data_eval_decide (+ ARAP/Rot blocks + decision), jtj_gram (records), jtj_scatter.  No host synchronisation, no allocation.
#include <cstring>

#include "common.cuh"
#include "lm_state.cuh"
#include "internal.h"
#include "super_b200.h"

namespace {

__global__ void band_from_fixed_kernel(const long long* __restrict__ s0, const long long* __restrict__ s1,
                                       const LMState* __restrict__ st, long long n_ab, long long n_tot, double inv_scale,
                                       double inv_gscale, double* __restrict__ AB, double* __restrict__ g, int zero_other) {
    const int sel = st ? st->sel : 0;
    const long long* src = sel ? s1 : s0;
    long long* oth = const_cast<long long*>(sel ? s0 : s1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tot; i += stride) {
        const long long v = src[i];
        if (i < n_ab) AB[i] = (double)v * inv_scale;          // power-of-two scale: exact
        else g[i - n_ab] = (double)v * inv_gscale;
        if (zero_other) oth[i] = 0;
    }
}

int from_fixed(const long long* s0, const long long* s1, const LMState* st, int n, int ldab, int shift, int gshift,
               double* AB, double* g, int zero_other, cudaStream_t stream) {
    const long long n_ab = (long long)n * ldab, n_tot = n_ab + n;
    long long blocks = (n_tot + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    band_from_fixed_kernel<<<(int)blocks, 256, 0, stream>>>(s0, s1, st, n_ab, n_tot, ldexp(1.0, -shift), ldexp(1.0, -gshift),
                                                          AB, g, zero_other);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // namespace

extern "C" {

int sb_band_from_fixed(const long long* store, int n, int ldab, int fx_shift, int fx_gshift, double* AB, double* g,
                       void* stream) {
    if (!store || !AB || !g || n <= 0 || ldab <= 0 || fx_shift < 0 || fx_gshift < 0 || fx_shift > 60 || fx_gshift > 60)
        return SB_ERR_ARG;
    return from_fixed(store, store, nullptr, n, ldab, fx_shift, fx_gshift, AB, g, 0, (cudaStream_t)stream);
}


int sb_event_create(void** ev) {
    if (!ev) return SB_ERR_ARG;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return SB_ERR_CUDA;
    *ev = (void*)e;
    return SB_OK;
}
int sb_event_destroy(void* ev) { return (ev && cudaEventDestroy((cudaEvent_t)ev) == cudaSuccess) ? SB_OK : SB_ERR_ARG; }
int sb_event_elapsed_ms(void* ev_begin, void* ev_end, float* ms) {
    if (!ev_begin || !ev_end || !ms) return SB_ERR_ARG;
    if (cudaEventSynchronize((cudaEvent_t)ev_end) != cudaSuccess) return SB_ERR_CUDA;
    return cudaEventElapsedTime(ms, (cudaEvent_t)ev_begin, (cudaEvent_t)ev_end) == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

// timeline event after a stage (SbLMFrame.stage_events)
static int stamp(const SbLMFrame* f, int& next, cudaStream_t st) {
    if (f->stage_events && next < f->n_stage_events &&
        cudaEventRecord((cudaEvent_t)f->stage_events[next], st) != cudaSuccess)
        return SB_ERR_CUDA;
    ++next;
    return SB_OK;
}

// k-th data-term pass of the frame: evaluation (+ decision) and Gram accumulation
static int jtj_pass(const SbLMFrame* f, int adopt, int k, cudaStream_t st, int& ev) {
    const bool timed = f->jtj_events && 2 * k + 1 < f->n_jtj_events;
    if (timed && cudaEventRecord((cudaEvent_t)f->jtj_events[2 * k], st) != cudaSuccess) return SB_ERR_CUDA;
    int rc = sbi::launch_eval_decide(f, adopt, st);
    if (rc != SB_OK) return rc;
    if (stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
    rc = sbi::launch_gram(f, st, 1);
    if (rc != SB_OK) return rc;
    if (stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
    rc = sbi::launch_gram(f, st, 2);
    if (rc != SB_OK) return rc;
    if (stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
    if (timed && cudaEventRecord((cudaEvent_t)f->jtj_events[2 * k + 1], st) != cudaSuccess) return SB_ERR_CUDA;
    return SB_OK;
}

static int lm_frame_launches(const SbLMFrame* f, void* stream) {
    if (!f || !f->points || !f->knn_idx || !f->knn_w || !f->ed_points || !f->ed_knn || !f->vmap || !f->nmap) return SB_ERR_ARG;
    if (!f->state || !f->beta || !f->best || !f->partials_loss || !f->rows || !f->keys || f->row_stride < f->n_cap ||
        (f->row_stride & 31))
        return SB_ERR_ARG;
    if (!f->rec_vals || !f->rec_keys || !f->rec_count || f->rec_cap < 1) return SB_ERR_ARG;
    if (!f->fx_store[0] || !f->fx_store[1] || !f->AB || !f->g || !f->band_overflow || !f->dinv || !f->info || !f->solver_ws)
        return SB_ERR_ARG;
    if (f->J <= 0 || f->J > 65535 || f->n != 7 * f->J || f->bw < 0 || f->ldab < f->bw + 1 || f->iterations < 1 ||
        f->iterations > 64 || f->n_cap <= 0)
        return SB_ERR_ARG;
    if (f->fx_shift < 0 || f->fx_shift > 60 || f->fx_gshift < 0 || f->fx_gshift > 60) return SB_ERR_ARG;
    const int n_loss_blocks = sb_data_loss_blocks(f->n_cap);
    if (f->n_partials_loss < n_loss_blocks) return SB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    LMState* state = (LMState*)f->state;
    const long long n_tot = (long long)f->n * f->ldab + f->n;
    // prologue: controller state, the store the first assembly goes to, the normal equations at the initial beta
    if (cudaMemsetAsync(f->fx_store[1], 0, (size_t)n_tot * sizeof(long long), st) != cudaSuccess) return SB_ERR_CUDA;
    if (cudaMemsetAsync(f->info, 0, sizeof(int), st) != cudaSuccess) return SB_ERR_CUDA;
    if (cudaMemsetAsync(f->rec_count, 0, sizeof(int), st) != cudaSuccess) return SB_ERR_CUDA;
    int ev = 0;
    if (stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
    int rc = sb_lm_begin(f->state, f->beta, f->best, f->J, f->u, f->v, f->minimal_loss, stream);
    if (rc != SB_OK) return rc;
    if (stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
    rc = jtj_pass(f, 1, 0, st, ev);
    if (rc != SB_OK) return rc;
    for (int it = 0; it < f->iterations; ++it) {
        const bool timed = f->solve_events && 2 * it + 1 < f->n_solve_events;
        if (timed && cudaEventRecord((cudaEvent_t)f->solve_events[2 * it], st) != cudaSuccess) return SB_ERR_CUDA;
        // store -> f64 band rides in the solve's first kernel (band_reverse_kernel), which also clears the other store
        rc = sb_band_solve4_step_fx(f->fx_store[0], f->fx_store[1], &state->sel, f->fx_shift, f->fx_gshift, 1, f->AB, f->ldab,
                                    f->n, f->bw, f->g, &state->u, f->dinv, f->info, f->solver_ws, f->solver_ws_bytes,
                                    f->n_ctas, &state->failed, f->beta, f->pos_node, stream);
        if (rc != SB_OK) return rc;
        if (timed && cudaEventRecord((cudaEvent_t)f->solve_events[2 * it + 1], st) != cudaSuccess) return SB_ERR_CUDA;
        if (stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
        if (it + 1 < f->iterations) {
            rc = jtj_pass(f, 0, it + 1, st, ev);
        } else {      // after the last solve only the loss of the trial beta is needed
            rc = sb_data_term_loss_decide(f->points, f->knn_idx, f->knn_w, f->n_cap, f->n_dev, f->ed_points, f->beta, f->J,
                                          f->vmap, f->nmap, f->H, f->W, f->intr, f->lam_data, f->partials_loss,
                                          n_loss_blocks, f->state, f->ed_knn, f->lam_arap, f->lam_rot, f->use_arap,
                                          f->use_rot, f->beta, f->best, stream);
            if (rc == SB_OK && stamp(f, ev, st) != SB_OK) return SB_ERR_CUDA;
        }
        if (rc != SB_OK) return rc;
    }
    return SB_OK;
}

// what sb_lm_frame keeps for a caller between frames (SbLMFrame.graph_cache)
struct LMGraphCache {
    cudaStream_t capture_stream;
    cudaGraphExec_t exec;
    int warmed;
    int capturing;      // sb_graph_scope_begin .. _end
    // sb_lm_frame: instantiated graphs by the argument block they were captured for (frames alternate between two input
    // buffers and the row bound is quantised by the caller, so the block repeats: a hit is one cudaGraphLaunch)
    static constexpr int SLOTS = 4;
    SbLMFrame key[SLOTS];
    cudaGraphExec_t slot_exec[SLOTS];
    unsigned long long last_use[SLOTS], clock;
};

static LMGraphCache* new_cache() {
    LMGraphCache* gc = new LMGraphCache();
    gc->capture_stream = nullptr; gc->exec = nullptr; gc->warmed = 0; gc->capturing = 0; gc->clock = 0;
    for (int i = 0; i < LMGraphCache::SLOTS; ++i) { gc->slot_exec[i] = nullptr; gc->last_use[i] = 0; memset(&gc->key[i], 0, sizeof(SbLMFrame)); }
    return gc;
}

int sb_lm_graph_destroy(void** cache) {
    if (!cache) return SB_ERR_ARG;
    LMGraphCache* gc = (LMGraphCache*)*cache;
    if (gc) {
        if (gc->exec) cudaGraphExecDestroy(gc->exec);
        for (int i = 0; i < LMGraphCache::SLOTS; ++i)
            if (gc->slot_exec[i]) cudaGraphExecDestroy(gc->slot_exec[i]);
        if (gc->capture_stream) cudaStreamDestroy(gc->capture_stream);
        delete gc;
        *cache = nullptr;
    }
    return SB_OK;
}

// capture ended: update the caller's instantiated graph in place (or instantiate it) and launch it on `stream`
static int graph_update_and_launch(LMGraphCache* gc, cudaGraph_t graph, cudaStream_t stream) {
    if (gc->exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(gc->exec, graph, &info) != cudaSuccess) {      // topology changed
            cudaGetLastError();
            cudaGraphExecDestroy(gc->exec);
            gc->exec = nullptr;
        }
    }
    if (!gc->exec && cudaGraphInstantiate(&gc->exec, graph, 0) != cudaSuccess) {
        cudaGraphDestroy(graph);
        cudaGetLastError();
        return SB_ERR_CUDA;
    }
    cudaGraphDestroy(graph);
    return cudaGraphLaunch(gc->exec, stream) == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int sb_graph_scope_begin(void** cache, void* stream, void** use_stream) {
    if (!cache || !use_stream) return SB_ERR_ARG;
    LMGraphCache* gc = (LMGraphCache*)*cache;
    if (!gc) {
        gc = new_cache();
        if (cudaStreamCreateWithFlags(&gc->capture_stream, cudaStreamNonBlocking) != cudaSuccess) { delete gc; return SB_ERR_CUDA; }
        *cache = gc;
    }
    if (gc->capturing) return SB_ERR_ARG;
    if (!gc->warmed) {            // first use: the scope's calls go straight to the caller's stream
        *use_stream = stream;
        return SB_OK;
    }
    if (cudaStreamBeginCapture(gc->capture_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) return SB_ERR_CUDA;
    gc->capturing = 1;
    *use_stream = (void*)gc->capture_stream;
    return SB_OK;
}

int sb_graph_scope_end(void** cache, void* stream, int abort_scope) {
    if (!cache || !*cache) return SB_ERR_ARG;
    LMGraphCache* gc = (LMGraphCache*)*cache;
    if (!gc->capturing) {         // the direct first use
        gc->warmed = 1;
        return SB_OK;
    }
    gc->capturing = 0;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(gc->capture_stream, &graph);
    if (abort_scope || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return abort_scope ? SB_OK : SB_ERR_CUDA;
    }
    return graph_update_and_launch(gc, graph, (cudaStream_t)stream);
}

int sb_stream_create(void** out) {      // non-blocking: does not synchronise with the legacy default stream
    if (!out) return SB_ERR_ARG;
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return SB_ERR_CUDA;
    *out = (void*)s;
    return SB_OK;
}
int sb_stream_destroy(void* s) { return (s && cudaStreamDestroy((cudaStream_t)s) == cudaSuccess) ? SB_OK : SB_ERR_ARG; }

int sb_copy_i32(int* dst, const int* src, int n, void* stream) {
    if (!dst || !src || n <= 0) return SB_ERR_ARG;
    return cudaMemcpyAsync(dst, src, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess
               ? SB_OK : SB_ERR_CUDA;
}

int sb_lm_frame(const SbLMFrame* f, void* stream) {
    if (!f) return SB_ERR_ARG;
    const bool events = f->jtj_events || f->solve_events || f->stage_events;
    if (!f->graph_cache || events) return lm_frame_launches(f, stream);
    LMGraphCache* gc = (LMGraphCache*)*f->graph_cache;
    if (!gc) {
        gc = new_cache();
        if (cudaStreamCreateWithFlags(&gc->capture_stream, cudaStreamNonBlocking) != cudaSuccess) { delete gc; return SB_ERR_CUDA; }
        *f->graph_cache = gc;
    }
    if (!gc->warmed) {            // first frame: direct launches (per-device launch configuration gets cached outside a capture)
        gc->warmed = 1;
        return lm_frame_launches(f, stream);
    }
    // The argument block this frame's graph depends on (host-side values and device addresses; what the kernels read through
    // those addresses -- beta, the LM state, the row count -- is not part of a graph)
    SbLMFrame key;
    memset(&key, 0, sizeof(key));
    memcpy(&key, f, sizeof(SbLMFrame));
    key.graph_cache = nullptr;
    int hit = -1, lru = 0;
    for (int i = 0; i < LMGraphCache::SLOTS; ++i) {
        if (gc->slot_exec[i] && memcmp(&gc->key[i], &key, sizeof(SbLMFrame)) == 0) hit = i;
        if (gc->last_use[i] < gc->last_use[lru]) lru = i;
    }
    if (hit >= 0) {
        gc->last_use[hit] = ++gc->clock;
        return cudaGraphLaunch(gc->slot_exec[hit], (cudaStream_t)stream) == cudaSuccess ? SB_OK : SB_ERR_CUDA;
    }
    // Miss: capture this frame's sequence (argument checks included: nothing is launched), update the least recently used
    // instantiated graph in place -- the topology is the same every frame, only kernel parameters and grids change -- and
    // launch it on the caller's stream.
    if (cudaStreamBeginCapture(gc->capture_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) return SB_ERR_CUDA;
    const int rc = lm_frame_launches(f, (void*)gc->capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(gc->capture_stream, &graph);
    if (rc != SB_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc != SB_OK ? rc : SB_ERR_CUDA;
    }
    cudaGraphExec_t keep = gc->exec;
    gc->exec = gc->slot_exec[lru];
    const int lrc = graph_update_and_launch(gc, graph, (cudaStream_t)stream);
    gc->slot_exec[lru] = gc->exec;
    gc->exec = keep;
    if (lrc == SB_OK) { gc->key[lru] = key; gc->last_use[lru] = ++gc->clock; }
    else gc->last_use[lru] = 0;
    return lrc;
}

}  // extern "C"
