// Banded Cholesky solve of (A + u I) x = g -- the LM step solve, replacing
// torch.linalg.cholesky + cholesky_solve   /root/reference/super/LM.py:38-51,97-100.
//
// The factorisation is a chain of n pivots; everything else has slack.  Roles on a cooperative grid (launched cooperatively
// only for the co-residency guarantee: there is no grid-wide barrier, every spin is bounded):
//
//   * P, one CTA: the pivot chain.  Warp 0 factors the 32x32 diagonal block (replicated 8x8 blocks in registers, no
//     shuffle or shared round trip on the chain) in its own loop over the panels; warp 1 builds L(k,k)^-1 behind it (for
//     U and R) and only ARRIVES at the next panel's barrier; four compute warps form L(k+1,k) = A(k+1,k) L(k,k)^-T by
//     forward substitution in 8-column blocks BEHIND the Cholesky (pre-phase: blocks 0..2, W_3, the products over
//     columns 0..23 of D -= L L^T), so that after the last pivot two DMMAs, one exchange and two DMMAs remain (final
//     step); two I/O warps store L(k+1,k), raise rows_done and -- with warp 6 -- stage the next panel's two tiles.
//   * U: one trailing tile per CTA and panel, three DMMA products (two "triangular solves" recomputed from the unfactored
//     tiles with the explicit inverse, one update).  U2 (wide bands): fixed tile owners, L(I,p) formed once per row.
//   * R: forward substitution y_p = Linv s_p behind P, then -- on the same CTA -- a push-style back substitution.
//
// Hand-over: inside the pivot CTA by value through NaN-armed shared buffers and named barriers; P -> U (inverse tiles, LI)
// and U -> P (the two tiles whose last update was the previous panel's: HM) BY VALUE through NaN-armed global tiles -- a
// value is its own ready flag, so neither a fence nor a flag nor a second L2 round trip sits on that path; R and the
// per-panel barrier among the update CTAs use release/acquire counters (diag_done | rows_done | upd_done per panel).
// The two-sided solve (sb_band_solve4*) runs two such pipelines from both ends of the band at once, then the middle block.
// Measurements behind every choice: DESIGN.md section 5, profiles/r1_microbench.md, profiles/r2s_chol_probe.md.
#include <cooperative_groups.h>

#include "common.cuh"
#include "super_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int NB = 32;
constexpr int S36 = 36;             // stride of DMMA operand tiles (conflict-free fragment loads)
constexpr int S33 = 33;             // stride of row-per-lane tiles
constexpr int T36 = NB * S36, T33 = NB * S33, T32 = NB * NB;
constexpr int THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_WB3 = 28;

struct Args3 {
    double* AB; int ldab, n, bw;
    double* g; const double* u; double* dinv; int* info;
    double* LB;      // L tiles below the diagonal, tile (I, d = I-J in 1..WB) at ((I*WB + d-1) * 1024), row-major 32x32
    double* LI;      // L(k,k)^-1, tile k at k*1024, row-major; NaN-armed before the launch: the update CTAs take it BY VALUE
    double* HM;      // hot-tile mailbox, NaN-armed: tiles (p+2,p+1) | (p+2,p+2) after panel p's update at (2p | 2p+1)*1024, row-major
    int* flags;      // diag_done[NP] | rows_done[NP] | upd_done[NP] | hot_done[NP]
    int NP, WB;
    int ke;            // panels to eliminate: NP = full solve; < NP = partial factorisation (two-sided solve), the trailing
                       // tiles keep the Schur complement and R stops after the forward substitution
    int rank0, ncta;   // this instance's CTAs are blockIdx.x in [rank0, rank0 + ncta)
    int two_phase;     // wide bands (more trailing tiles than update CTAs): L(I,p) once per row, then one product per tile
    long long* prof;   // 32 counters (debug & 4)
    int debug;         // 1: U skips its tile work, 2: no back substitution, 4: cycle counters
};

// timing experiments / cycle counters exist only in the SB_DEBUG_EXPORTS build: in the product build every debug test is a
// compile-time 0, so the counters, their registers and their code are gone (the pivot CTA's loop is instruction-fetch sensitive)
#ifdef SB_DEBUG_EXPORTS
#define DBGF(a) ((a).debug)
#else
#define DBGF(a) 0
#endif

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// hot-path study (debug build): sums of %globaltimer at five events of the panels 1..ke-2, prof[64..68] (+ counts at [69..73])
#define GSTAMP(a, slot, p_) do { if ((DBGF(a) & 4) && (p_) >= 1 && (p_) <= (a).ke - 2) { \
    atomicAdd((unsigned long long*)(a).prof + 64 + (slot), gtime_ns()); atomicAdd((unsigned long long*)(a).prof + 69 + (slot), 1ull); } } while (0)

// quiet-NaN test on the bit pattern: the polls of the by-value hand-over must not touch the FP64 pipe (DSETP) -- warp 4 of the
// pivot CTA polls while it shares scheduler AND FP64 pipe with the pivot chain
__device__ __forceinline__ bool is_qnan(double v) { return (__double2hiint(v) & 0x7ff80000) == 0x7ff80000; }

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// one thread spins until *flag >= target; bounded (a lost signal must not hang the GPU)
__device__ __forceinline__ void spin_until(const int* flag, int target) {
    long long it = 0;
    while (ld_acquire(flag) < target) {
        __nanosleep(32);
        if (++it > (1LL << 24)) __trap();
    }
}
// whole CTA waits (thread 0 polls)
__device__ __forceinline__ void cta_wait(const int* flag, int target) {
    if (threadIdx.x == 0) spin_until(flag, target);
    __syncthreads();
}

__device__ __forceinline__ bool in_band(const Args3& a, int i, int j) { return i < a.n && j <= i && i - j <= a.bw; }
__device__ __forceinline__ double* ab_at(const Args3& a, int i, int j) {
    return a.AB + (size_t)i * a.ldab + (j - i + a.bw);
}
__device__ __forceinline__ double* lb_tile(const Args3& a, int I, int d) {
    return a.LB + ((size_t)I * a.WB + (d - 1)) * T32;
}

// tile (I,J) of the band matrix -> shared (row-major, given stride), executed by threads t of nt; outside = 0
__device__ __forceinline__ void load_ab_tile(const Args3& a, int I, int J, double* __restrict__ T, int stride, int t,
                                             int nt) {
    for (int base = 0; base < T32; base += 4 * nt) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = base + q * nt + t, i = NB * I + (e >> 5), j = NB * J + (e & 31);
            v[q] = (e < T32 && in_band(a, i, j)) ? __ldcg(ab_at(a, i, j)) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = base + q * nt + t;
            if (e < T32) T[(e >> 5) * stride + (e & 31)] = v[q];
        }
    }
}
// two band tiles at once with EVERY load in flight before the first shared store (one L2 round trip instead of one
// per batch of four: the batched loader cost 5 k cycles per panel on the 96 I/O threads of the pivot CTA)
template <int NT>
__device__ __forceinline__ void load_ab_tiles2(const Args3& a, int I0, int J0, double* __restrict__ T0, int s0,
                                               int I1, int J1, double* __restrict__ T1, int s1, int t) {
    static_assert(NT % 32 == 0, "whole warps");
    constexpr int RSTEP = NT / 32, PER = (NB + RSTEP - 1) / RSTEP;
    // thread -> fixed column c, rows r0, r0+RSTEP, ...: one pointer per tile, advanced by RSTEP*(ldab-1) per load
    const int c = t & 31, r0 = t >> 5;
    const long long step = (long long)RSTEP * (a.ldab - 1);
    const bool has1 = I1 >= 0;
    const double* p0 = a.AB + (size_t)(NB * I0 + r0) * a.ldab + (NB * J0 + c - NB * I0 - r0 + a.bw);
    const double* p1 = has1 ? a.AB + (size_t)(NB * I1 + r0) * a.ldab + (NB * J1 + c - NB * I1 - r0 + a.bw) : p0;
    const int d0 = NB * (I0 - J0) + r0 - c, d1 = NB * (I1 - J1) + r0 - c;     // i - j at q = 0, +RSTEP per q
    const int i0 = NB * I0 + r0, i1 = NB * I1 + r0;
    double v0[PER], v1[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int dr = q * RSTEP;
        const bool ok0 = (r0 + dr < NB) && (i0 + dr < a.n) && (d0 + dr >= 0) && (d0 + dr <= a.bw);
        const bool ok1 = has1 && (r0 + dr < NB) && (i1 + dr < a.n) && (d1 + dr >= 0) && (d1 + dr <= a.bw);
        v0[q] = ok0 ? __ldcg(p0 + q * step) : 0.0;
        v1[q] = ok1 ? __ldcg(p1 + q * step) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int r = r0 + q * RSTEP;
        if (r < NB) {
            T0[r * s0 + c] = v0[q];
            if (has1) T1[r * s1 + c] = v1[q];
        }
    }
}
// The pivot CTA's two tiles of the next panel from the hot-tile mailbox (NaN-armed, written by the owners of the tiles'
// last update): same thread mapping and in-band predicate as above, every load in flight at once, re-polled until no
// in-band value is NaN.
template <int NT>
__device__ __forceinline__ void load_hot_tiles2(const Args3& a, const double* __restrict__ H0, const double* __restrict__ H1,
                                                int I0, int J0, double* __restrict__ T0, int s0, int I1, int J1,
                                                double* __restrict__ T1, int s1, int t) {
    constexpr int RSTEP = NT / 32, PER = (NB + RSTEP - 1) / RSTEP;
    const int c = t & 31, r0 = t >> 5;
    const int d0 = NB * (I0 - J0) + r0 - c, d1 = NB * (I1 - J1) + r0 - c;
    const int i0 = NB * I0 + r0, i1 = NB * I1 + r0;
    double v0[PER], v1[PER];
    int polls = 0;
    bool again;
    do {
        again = false;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int dr = q * RSTEP;
            const bool ok0 = (r0 + dr < NB) && (i0 + dr < a.n) && (d0 + dr >= 0) && (d0 + dr <= a.bw);
            const bool ok1 = (r0 + dr < NB) && (i1 + dr < a.n) && (d1 + dr >= 0) && (d1 + dr <= a.bw);
            v0[q] = ok0 ? __ldcg(H0 + (r0 + dr) * NB + c) : 0.0;
            v1[q] = ok1 ? __ldcg(H1 + (r0 + dr) * NB + c) : 0.0;
            again = again || is_qnan(v0[q]) || is_qnan(v1[q]);
        }
        if (again && (++polls & 63) == 0) {
            if (*(volatile int*)a.info != 0) again = false;
            else if (polls > (1 << 16)) { *a.info = 1; again = false; }
        }
    } while (again);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int r = r0 + q * RSTEP;
        if (r < NB) {
            T0[r * s0 + c] = v0[q];
            T1[r * s1 + c] = v1[q];
        }
    }
}
// contiguous row-major 32x32 global tile -> shared with stride
__device__ __forceinline__ void load_g_tile(const double* __restrict__ G, double* __restrict__ T, int stride, int t,
                                            int nt) {
    for (int base = 0; base < T32; base += 4 * nt) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = base + q * nt + t;
            v[q] = (e < T32) ? __ldcg(G + e) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = base + q * nt + t;
            if (e < T32) T[(e >> 5) * stride + (e & 31)] = v[q];
        }
    }
}
// The same from a NaN-armed tile another CTA is writing: a value is its own ready flag (8-byte stores are single
// transactions), so no flag, no fence and no second round trip.  Gives up when a failure has been flagged (the data may then
// be NaN for good) or after ~2^16 polls, which flags the failure itself.
__device__ __forceinline__ void load_g_tile_polled(const double* __restrict__ G, double* __restrict__ T, int stride, int t,
                                                   int nt, int* info) {
    for (int base = 0; base < T32; base += 4 * nt) {
        double v[4];
        int polls = 0;
        bool again;
        do {
            again = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = base + q * nt + t;
                v[q] = (e < T32) ? __ldcg(G + e) : 0.0;
                again = again || is_qnan(v[q]);
            }
            if (again && (++polls & 63) == 0) {
                if (*(volatile int*)info != 0) again = false;
                else if (polls > (1 << 16)) { *info = 1; again = false; }
            }
        } while (again);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = base + q * nt + t;
            if (e < T32) T[(e >> 5) * stride + (e & 31)] = v[q];
        }
    }
}
__device__ __forceinline__ void store_g_tile(double* __restrict__ G, const double* __restrict__ T, int stride, int t,
                                             int nt) {
    for (int e = t; e < T32; e += nt) __stcg(G + e, T[(e >> 5) * stride + (e & 31)]);
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a_, double b_) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a_), "d"(b_));
}

// Dst[8*bi .. +8, 8*bj .. +8] = Src[8*bi.., :] * Linv[8*bj.., :]^T for the block columns bj in `mask`;
// Linv is lower triangular, so only k < 8*bj+8 contributes (2*bj+2 k-steps of 4).  All tiles stride S36.
__device__ __forceinline__ void trsm_strip(const double* __restrict__ Src, const double* __restrict__ Linv,
                                           double* __restrict__ Dst, int bi, unsigned mask, int lane) {
    const int fr = lane >> 2, fc = lane & 3;
    double af[8], c[4][2];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) af[ks] = Src[(8 * bi + fr) * S36 + 4 * ks + fc];
#pragma unroll
    for (int bj = 0; bj < 4; ++bj) c[bj][0] = c[bj][1] = 0.0;
    // k-steps outermost: the (up to) four accumulator chains are independent and overlap in the DMMA pipe
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int bj = 0; bj < 4; ++bj)
            if (((mask >> bj) & 1u) && ks < 2 * bj + 2)
                dmma884(c[bj][0], c[bj][1], af[ks], Linv[(8 * bj + fr) * S36 + 4 * ks + fc]);
#pragma unroll
    for (int bj = 0; bj < 4; ++bj)
        if ((mask >> bj) & 1u)
            *reinterpret_cast<double2*>(Dst + (8 * bi + fr) * S36 + 8 * bj + 2 * fc) = make_double2(c[bj][0], c[bj][1]);
}
// acc = sum_k Pa[8*bi.., k] * Pb[8*bj.., k]   (full 32-deep product of two stride-S36 tiles)
__device__ __forceinline__ void mma_block(const double* __restrict__ Pa, const double* __restrict__ Pb, int bi, int bj,
                                          int lane, double& c0, double& c1) {
    const int fr = lane >> 2, fc = lane & 3;
    c0 = 0.0; c1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        dmma884(c0, c1, Pa[(8 * bi + fr) * S36 + 4 * ks + fc], Pb[(8 * bj + fr) * S36 + 4 * ks + fc]);
}

// reciprocal square root for the pivot chain: float seed + one Newton step in double (rel. error ~1e-14)
__device__ __forceinline__ double rsqrt_chain(double x) {
    const double r = (double)rsqrtf((float)x);
    const double e = fma(-x * r, r, 1.0);
    return fma(0.5 * r, e, r);
}

// the six lower 8x8 blocks right of the first block column: (1,1) (2,1) (3,1) (2,2) (3,2) (3,3)
__host__ __device__ constexpr int pair_i(int q) { return q == 0 ? 1 : q == 1 ? 2 : q == 2 ? 3 : q == 3 ? 2 : 3; }
__host__ __device__ constexpr int pair_j(int q) { return q < 3 ? 1 : q < 5 ? 2 : 3; }

// reciprocal square root for the pivot chain: MUFU.RSQ64H seed (rel. error 9e-7 measured) + one Halley step
//   e = 1 - x r0^2,  r = r0 + r0 e (1/2 + 3/8 e)      (cubic: rel. error ~1e-16; 4 dependent FP64 ops + MUFU = 53
// cycles measured, against 63 for float seed + Newton at 3e-14)
__device__ __forceinline__ double rsqrt_pivot(double x) {
    double r0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x * r0, r0, 1.0);
    const double t = fma(0.375, e, 0.5);
    return fma(r0 * e, t, r0);
}

// ---- 32x32 Cholesky in one warp, blocked by 8 columns.  D (shared, stride S33) holds the block.
// The pivot chain must not contain a shuffle or a shared-memory round trip (26 and 34 cycles measured, against 8.4
// for a dependent DFMA): every lane keeps a REPLICA of the current 8x8 diagonal block (36 registers) and factors it
// redundantly, so rsqrt -> scale -> own-pivot FMA stays inside one thread (about 70 cycles per column).  Lane r also
// carries the 8 entries of its own row r and solves them against the replica as the columns appear.  Column C of L
// goes to Lcol[C*S36 + row] and 1/L(C,C) to dinvs[C] immediately (the inverse builder reads them behind us); the
// trailing columns are updated in shared memory by DMMA products (next block column first).
// Small code on purpose: fully unrolled 32-column versions of this and of the inverse are ~100 KB of SASS.
__device__ __forceinline__ bool warp_potrf_blocked(double* __restrict__ D, double* __restrict__ Lcol,
                                                   double* __restrict__ dinvs, double* __restrict__ Ls, int lane,
                                                   long long* ts = nullptr, long long tbase = 0) {
    const int fr = lane >> 2, fc = lane & 3, jc = lane & 7;
    bool bad = false;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int c0 = 8 * b;
        double d[8][8], p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) d[i][j] = D[(c0 + i) * S33 + c0 + j];      // broadcast loads
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = D[lane * S33 + c0 + j];
        double x[8];                                                                 // column jc of T_b = L_bb^-1
        double myinv = 0.0;
        long long q0 = 0, q1;
        if (ts) { if (p[7] + d[7][7] != 1.2345e300) q0 = clock64(); ts[4] += q0 - tbase; tbase = q0; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double piv = d[c][c];
            // row c of the inverse only needs row c of L (final since step c-1): everything but the last
            // multiply is off the pivot chain and fills its idle issue slots
            double s0 = (jc == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
            for (int t = 0; t < c; ++t) {
                if (t & 1) s1 = fma(-d[c][t], x[t], s1); else s0 = fma(-d[c][t], x[t], s0);
            }
            const double inv = rsqrt_pivot(piv);
            x[c] = (s0 + s1) * inv;
            bad = bad || !(piv > 0.0);
#pragma unroll
            for (int i = c + 1; i < 8; ++i) d[i][c] *= inv;                          // l_ic of the replica
#pragma unroll
            for (int j = c + 1; j < 8; ++j)
#pragma unroll
                for (int i = j; i < 8; ++i) d[i][j] = fma(-d[i][c], d[j][c], d[i][j]);
            p[c] *= inv;                                                             // l_rc of my own row
#pragma unroll
            for (int j = c + 1; j < 8; ++j) p[j] = fma(-p[c], d[j][c], p[j]);
            if (lane == c) myinv = inv;
        }
        // The block's 8 columns and pivots leave the registers only now: every reader (inverse builder, helpers, the L(k+1,k)
        // pre-phase) waits for whole blocks, and a shared-memory store inside the loop above can stall the in-order chain
        // behind the other warps' shared-memory traffic.  dinvs[c0+7] and T_b(7,7) are the ready flags: columns first.
#pragma unroll
        for (int c = 0; c < 8; ++c) Lcol[(c0 + c) * S36 + lane] = p[c];
        if (lane < 8) dinvs[c0 + lane] = myinv;
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) Ls[(c0 + i) * S36 + c0 + lane] = bad ? 0.0 : x[i];   // (7,7) last: the block's ready flag
        }
        __syncwarp();
        if (ts) { if (x[7] + p[7] != 1.2345e300) q1 = clock64(); ts[5] += q1 - tbase; tbase = q1; }
        if (b < 3) {
            // D[i][b+1] -= sum_{t in block b} L[i][t] L[b+1][t] for the NEXT block column only (rows b+1..3); the helper
            // warps update the block columns further right behind us and have finished the previous round (barrier 4+b)
            bar_sync(4 + b, 128);
            double af[3][2];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) af[r][ks] = (r + 1 > b) ? Lcol[(c0 + 4 * ks + fc) * S36 + 8 * (r + 1) + fr] : 0.0;
            double acc[3][2];
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (r + 1 > b) {
                        // rows block r+1, columns block b+1: operand B = rows of block b+1 = af[b] (b is a run-time value)
                        const double bf = (b == 0) ? af[0][ks] : (b == 1) ? af[1][ks] : af[2][ks];
                        dmma884(acc[r][0], acc[r][1], af[r][ks], bf);
                    }
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if (r + 1 > b) {
                    double* dd = D + (8 * (r + 1) + fr) * S33 + 8 * (b + 1) + 2 * fc;
                    dd[0] -= acc[r][0];
                    dd[1] -= acc[r][1];
                }
            __syncwarp();
        }
        if (ts) { q1 = clock64(); ts[6] += q1 - tbase; tbase = q1; }
    }
    return bad;
}

// ---- The same factorisation SPLIT between the chain warp and the helpers.  The chain warp above carries all 32 rows (own-row
// solve interleaved with the replica: 224 FP64 instructions per 8 columns on a warp that issues one every 3.4 cycles) and, between
// two blocks, updates the whole next block column.  Here it factors ONLY the 8x8 diagonal blocks (188 instructions) and, between
// two blocks, forms L(b+1,b) = D(b+1,b) T_b^T and the next diagonal block D(b+1,b+1) -= L(b+1,b) L(b+1,b)^T (2 + 2 DMMAs);
// the rows below, L(r,b) = D(r,b) T_b^T for r >= b+2, and every other trailing update are formed by the three helper warps one
// block behind (helper_split), which have a whole block's chain (~700 cycles) for a handful of DMMAs.
// lready[b] = tag once L(b+1,b) is in Lcol (T_b is in Ls before that): the helpers' and the other readers' flag.
__device__ __forceinline__ bool warp_potrf_split(double* __restrict__ D, double* __restrict__ Lcol, double* __restrict__ dinvs,
                                                 double* __restrict__ Ls, volatile int* lready, int tag, int lane,
                                                 long long* ts = nullptr, long long tbase = 0) {
    const int fr = lane >> 2, fc = lane & 3, jc = lane & 7;
    bool bad = false;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int c0 = 8 * b;
        double d[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) d[i][j] = D[(c0 + i) * S33 + c0 + j];      // broadcast loads
        double x[8];                                                                 // column jc of T_b = L_bb^-1
        double myinv = 0.0;
        long long q0 = 0, q1;
        if (ts) { if (d[7][7] != 1.2345e300) q0 = clock64(); ts[4] += q0 - tbase; tbase = q0; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double piv = d[c][c];
            double s0 = (jc == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
            for (int t = 0; t < c; ++t) {
                if (t & 1) s1 = fma(-d[c][t], x[t], s1); else s0 = fma(-d[c][t], x[t], s0);
            }
            const double inv = rsqrt_pivot(piv);
            x[c] = (s0 + s1) * inv;
            bad = bad || !(piv > 0.0);
#pragma unroll
            for (int i = c + 1; i < 8; ++i) d[i][c] *= inv;
#pragma unroll
            for (int j = c + 1; j < 8; ++j)
#pragma unroll
                for (int i = j; i < 8; ++i) d[i][j] = fma(-d[i][c], d[j][c], d[i][j]);
            if (lane == c) myinv = inv;
        }
        if (lane < 8) {
            dinvs[c0 + lane] = myinv;
#pragma unroll
            for (int i = 0; i < 8; ++i) Ls[(c0 + i) * S36 + c0 + lane] = bad ? 0.0 : x[i];   // (7,7) last: T_b's ready flag
        }
        __syncwarp();
        if (ts) { if (x[7] != 1.2345e300) q1 = clock64(); ts[5] += q1 - tbase; tbase = q1; }
        if (b < 3) {
            bar_sync(4 + b, 128);       // the helpers have applied every earlier block to D(b+1,b) and D(b+1,b+1)
            const double a0 = D[(c0 + 8 + fr) * S33 + c0 + fc], a1 = D[(c0 + 8 + fr) * S33 + c0 + 4 + fc];
            const double t0 = Ls[(c0 + fr) * S36 + c0 + fc], t1 = Ls[(c0 + fr) * S36 + c0 + 4 + fc];
            double* dd = D + (c0 + 8 + fr) * S33 + c0 + 8 + 2 * fc;
            const double dd0 = dd[0], dd1 = dd[1];
            double l0 = 0.0, l1 = 0.0;
            dmma884(l0, l1, a0, t0);
            dmma884(l0, l1, a1, t1);
            Lcol[(c0 + 2 * fc) * S36 + c0 + 8 + fr] = l0;
            Lcol[(c0 + 2 * fc + 1) * S36 + c0 + 8 + fr] = l1;
            __syncwarp();
            if (lane == 0) { __threadfence_block(); lready[b] = tag; }
            const double f0 = Lcol[(c0 + fc) * S36 + c0 + 8 + fr], f1 = Lcol[(c0 + 4 + fc) * S36 + c0 + 8 + fr];
            double u0 = 0.0, u1 = 0.0;
            dmma884(u0, u1, f0, f0);
            dmma884(u0, u1, f1, f1);
            dd[0] = dd0 - u0;
            dd[1] = dd1 - u1;
            __syncwarp();
        }
        if (ts) { q1 = clock64(); ts[6] += q1 - tbase; tbase = q1; }
    }
    return bad;
}

// helper h (0..2) behind block b of warp_potrf_split: rows r >= b+2 of block column b, then the updates of every block right of
// it except the next diagonal one.  Barrier 11 among the three helpers.
__device__ __forceinline__ void helper_split(double* __restrict__ D, double* __restrict__ Lcol, const double* __restrict__ Ls,
                                             volatile int* lready, int tag, int b, int h, int lane) {
    const int fr = lane >> 2, fc = lane & 3, c0 = 8 * b;
    {
        int spins = 0;
        while (lready[b] != tag && ++spins < (1 << 20)) {}
        asm volatile("" ::: "memory");
    }
    const volatile double* Lv = Lcol;
    const volatile double* Tv = Ls;
    const int r = b + 2 + h;
    if (r <= 3) {
        const double a0 = D[(8 * r + fr) * S33 + c0 + fc], a1 = D[(8 * r + fr) * S33 + c0 + 4 + fc];
        double l0 = 0.0, l1 = 0.0;
        dmma884(l0, l1, a0, Tv[(c0 + fr) * S36 + c0 + fc]);
        dmma884(l0, l1, a1, Tv[(c0 + fr) * S36 + c0 + 4 + fc]);
        Lcol[(c0 + 2 * fc) * S36 + 8 * r + fr] = l0;
        Lcol[(c0 + 2 * fc + 1) * S36 + 8 * r + fr] = l1;
    }
    if (b >= 2) return;
    bar_sync(11, 96);
    // blocks (oi,oj): b = 0: h0 (2,1) (2,2) | h1 (3,1) (3,2) | h2 (3,3);  b = 1: h0 (3,2) | h1 (3,3)
    int oi[2] = {-1, -1}, oj[2] = {-1, -1};
    if (b == 0) {
        if (h == 0) { oi[0] = 2; oj[0] = 1; oi[1] = 2; oj[1] = 2; }
        else if (h == 1) { oi[0] = 3; oj[0] = 1; oi[1] = 3; oj[1] = 2; }
        else { oi[0] = 3; oj[0] = 3; }
    } else {
        if (h == 0) { oi[0] = 3; oj[0] = 2; }
        else if (h == 1) { oi[0] = 3; oj[0] = 3; }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (oi[q] < 0) continue;
        double u0 = 0.0, u1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            dmma884(u0, u1, Lv[(c0 + 4 * ks + fc) * S36 + 8 * oi[q] + fr], Lv[(c0 + 4 * ks + fc) * S36 + 8 * oj[q] + fr]);
        double* dd = D + (8 * oi[q] + fr) * S33 + 8 * oj[q] + 2 * fc;
        dd[0] -= u0;
        dd[1] -= u1;
    }
    __syncwarp();
}

// ---- L^-1 behind the Cholesky, blocked by 8, on two warps.  With 8x8 blocks L_ik, T_b = L_bb^-1:
//        M_bb = T_b,     M_bj = -T_b * S_bj,   S_bj = sum_{k=j}^{b-1} L_bk M_kj     (j < b)
// Warp "T" (warp_tinv_trailing) inverts the diagonal blocks row by row right behind the pivot chain; warp "M"
// (warp_linv_blocked) forms S_bj while the chain is still on block b (it only needs block columns < b), then
// M_bj and publishes block row b.  After the last pivot only T_3 's last row, three 8x8 products and one store
// remain.  Hand-over is by value: L[i][t] = Lcol[t*S36+i], 1/L[i][i] = dinvs[i] and the diagonal blocks of the
// inverse tile are NaN until written (a value is its own ready flag; a NaN in any result re-reads).
__device__ __forceinline__ void warp_tinv_trailing(const volatile double* Lcol, const volatile double* dinvs,
                                                   double* Ls, int lane, bool& failed, long long* ts = nullptr,
                                                   long long tbase = 0) {
    const int jc = lane & 7;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int c0 = 8 * b;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = 0.0;
            if (!failed) {
                int tries = 0;
                while (true) {
                    double d = dinvs[c0 + i];
                    int spins = 0;
                    while (d != d && ++spins < (1 << 18)) d = dinvs[c0 + i];
                    double s0 = (jc == i) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
                    for (int t = 0; t < i; ++t) {
                        const double l = Lcol[(c0 + t) * S36 + c0 + i];
                        if (t & 1) s1 = fma(-l, x[t], s1); else s0 = fma(-l, x[t], s0);
                    }
                    v = (s0 + s1) * d;
                    if (!__any_sync(FULL, v != v)) break;
                    if (++tries > 4) { failed = true; v = 0.0; break; }   // NaN input / indefinite block: give up
                }
            }
            x[i] = v;
            if (lane < 8) Ls[(c0 + i) * S36 + c0 + lane] = v;          // row 7 last: (7,7) is the block's ready flag
        }
        if (ts) ts[b] += clock64() - tbase;
    }
}

// block row b (8 rows) of the shared inverse tile -> global tile (what U and R read)
__device__ __forceinline__ void publish_inv_rows(const double* Ls, double* __restrict__ Lg, int b, int lane) {
    const int c0 = 8 * b;
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = ((const volatile double*)Ls)[(c0 + i) * S36 + lane];
#pragma unroll
    for (int i = 0; i < 8; ++i) __stcg(Lg + (c0 + i) * NB + lane, (r[i] == r[i]) ? r[i] : 0.0);
}

// Ls: shared inverse tile (stride S36; diagonal blocks NaN-armed, rest zero), Lg: global tile, Sc: 3 x 96 scratch.
// Block rows 0..2 are published to Lg as they complete; the LAST one is left to the caller (publish_inv_rows(.., 3, ..)
// behind the next top-of-panel barrier): its 8 loads + 8 stores sat on the pivot chain's panel-to-panel path.
template <typename OnLast>
__device__ __forceinline__ void warp_linv_blocked(const double* Lcol, const volatile double* dinvs, double* Ls,
                                                  double* __restrict__ Lg, double* __restrict__ Sc, int lane,
                                                  bool& failed, volatile int* rowflag, int tag, OnLast on_last,
                                                  long long* wacc = nullptr, long long* ts = nullptr, long long tbase = 0) {
    const int fr = lane >> 2, fc = lane & 3;
    constexpr int SS = 12;
    long long w0 = wacc ? clock64() : 0, w1;
#define WPROF(slot) do { if (wacc) { w1 = clock64(); wacc[slot] += w1 - w0; w0 = w1; } } while (0)
    auto wait_value = [&](const volatile double* p) {
        if (failed) return;
        int spins = 0;
        double dq = *p;
        while (dq != dq && ++spins < (1 << 18)) dq = *p;
        if (dq != dq) failed = true;
        asm volatile("" ::: "memory");       // the plain loads below stay behind the flag
    };
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int c0 = 8 * b;
        if (b >= 1) {
            wait_value(dinvs + c0 - 1);      // block columns < b are complete
            WPROF(0);
            for (int tries = 0; tries < 4; ++tries) {
                double sacc[3][2];
#pragma unroll
                for (int j = 0; j < 3; ++j) sacc[j][0] = sacc[j][1] = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k >= b) break;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int kk = 8 * k + 4 * ks + fc;
                        const double lf = ((const volatile double*)Lcol)[kk * S36 + c0 + fr];
#pragma unroll
                        for (int j = 0; j <= k; ++j)
                            dmma884(sacc[j][0], sacc[j][1], lf, ((const volatile double*)Ls)[kk * S36 + 8 * j + fr]);
                    }
                }
                bool nan_seen = false;
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < b) {
                        Sc[j * 8 * SS + fr * SS + 2 * fc] = sacc[j][0];
                        Sc[j * 8 * SS + fr * SS + 2 * fc + 1] = sacc[j][1];
                        nan_seen = nan_seen || (sacc[j][0] != sacc[j][0]) || (sacc[j][1] != sacc[j][1]);
                    }
                if (failed || !__any_sync(FULL, nan_seen)) break;
                if (tries == 3) failed = true;
            }
            __syncwarp();
            WPROF(1);
        }
        wait_value(Ls + (c0 + 7) * S36 + c0 + 7);      // T_b is complete (warp "T")
        if (b == 3) on_last();                         // the Cholesky of this panel is complete: what the pivot CTA itself needs exists
        WPROF(2);
        if (b >= 1) {
            // M_bj = -T_b S_bj, the three products interleaved
            for (int tries = 0; tries < 4; ++tries) {
                double m[3][2];
#pragma unroll
                for (int j = 0; j < 3; ++j) m[j][0] = m[j][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const double tf = ((const volatile double*)Ls)[(c0 + fr) * S36 + c0 + 4 * ks + fc];
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (j < b) dmma884(m[j][0], m[j][1], tf, Sc[j * 8 * SS + (4 * ks + fc) * SS + fr]);
                }
                bool nan_seen = false;
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < b) {
                        Ls[(c0 + fr) * S36 + 8 * j + 2 * fc] = -m[j][0];
                        Ls[(c0 + fr) * S36 + 8 * j + 2 * fc + 1] = -m[j][1];
                        nan_seen = nan_seen || (m[j][0] != m[j][0]) || (m[j][1] != m[j][1]);
                    }
                if (failed || !__any_sync(FULL, nan_seen)) break;
                if (tries == 3) failed = true;
            }
            __syncwarp();
            WPROF(3);
        }
        // block row b of the inverse is final in shared memory (flag for the pivot CTA's own products): publish it (the last
        // one: see above)
        __threadfence_block();
        if (lane == 0) rowflag[b] = tag;
        if (b < 3) publish_inv_rows(Ls, Lg, b, lane);
        WPROF(4);
        if (ts) ts[b] += clock64() - tbase;
    }
#undef WPROF
}

// D[8bi.., 8bj..] -= Lx[8bi..] Lx[8bj..]^T for up to two lower blocks (bi0,bj0), (bi1,bj1) (bi1 < 0: one block), with the
// damping / identity padding on diagonal entries; the two accumulator chains are interleaved
__device__ __forceinline__ void syrk_blocks(const double* __restrict__ Lx, double* __restrict__ D, bool has_prev, int bi0,
                                            int bj0, int bi1, int bj1, int k, int n, double u, int lane, bool damp = true) {
    const int fr = lane >> 2, fc = lane & 3;
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (has_prev) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            dmma884(a0, a1, Lx[(8 * bi0 + fr) * S36 + 4 * ks + fc], Lx[(8 * bj0 + fr) * S36 + 4 * ks + fc]);
            if (bi1 >= 0) dmma884(b0, b1, Lx[(8 * bi1 + fr) * S36 + 4 * ks + fc], Lx[(8 * bj1 + fr) * S36 + 4 * ks + fc]);
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q == 1 && bi1 < 0) break;
        const int i = 8 * (q ? bi1 : bi0) + fr, j = 8 * (q ? bj1 : bj0) + 2 * fc, gi = NB * k + i;
        double v0 = D[i * S33 + j] - (q ? b0 : a0), v1 = D[i * S33 + j + 1] - (q ? b1 : a1);
        if (damp && i == j) v0 = (gi < n) ? v0 + u : 1.0;
        if (damp && i == j + 1) v1 = (gi < n) ? v1 + u : 1.0;
        D[i * S33 + j] = v0;
        D[i * S33 + j + 1] = v1;
    }
}

// helper warp: once block column b of L exists (its last pivot is the ready flag), D[8oi.., 8oj..] -= L[8oi.., blk b] L[8oj.., blk b]^T
__device__ __forceinline__ void helper_trailing(double* __restrict__ D, const double* Lcol, const volatile double* dinvs, int b,
                                                int oi, int oj, int lane, bool active) {
    const int fr = lane >> 2, fc = lane & 3, c0 = 8 * b;
    int spins = 0;
    double dq = dinvs[c0 + 7];
    while (dq != dq && ++spins < (1 << 20)) dq = dinvs[c0 + 7];
    asm volatile("" ::: "memory");
    if (!active) return;
    for (int tries = 0; tries < 4; ++tries) {
        double c0_ = 0.0, c1_ = 0.0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            dmma884(c0_, c1_, ((const volatile double*)Lcol)[(c0 + 4 * ks + fc) * S36 + 8 * oi + fr],
                    ((const volatile double*)Lcol)[(c0 + 4 * ks + fc) * S36 + 8 * oj + fr]);
        if (__any_sync(FULL, (c0_ != c0_) || (c1_ != c1_)) && tries < 3) continue;     // a value not yet visible: re-read
        double* dd = D + (8 * oi + fr) * S33 + 8 * oj + 2 * fc;
        dd[0] -= c0_;
        dd[1] -= c1_;
        break;
    }
    __syncwarp();
}

// =====================================================================================================
// P: the pivot chain.  Warp roles: 0 Cholesky (+DMMA), 1 inverse builder, {0,2,3,5} DMMA products,
// {4,6,7} I/O (flags, L(k,k-1) store, staging of the next panel's tiles).
// =====================================================================================================
__device__ void role_P(const Args3& a, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, NP = a.NP, WB = a.WB;
    int* diag_done = a.flags;
    int* rows_done = a.flags + NP;
    int* upd_done = a.flags + 2 * NP;
    const int NU = a.ncta - 2;
    const int KE = a.ke;
    const int hot_expected = (WB >= 2) ? 2 : 0;
    const double u = a.u ? *a.u : 0.0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    double* Dbuf = smem;                   // 2 x T33
    double* Xbuf = Dbuf + 2 * T33;         // 2 x T36
    double* Lxbuf = Xbuf + 2 * T36;        // 2 x T36
    double* Linv = Lxbuf + 2 * T36;        // 2 x T36
    double* Lcol = Linv + 2 * T36;         // 2 x T36 (column-major L: Lcol[c*S36 + r])
    double* dinvs = Lcol + 2 * T36;        // 2 x NB
    double* Sc = dinvs + 2 * NB;           // 3 x 96 scratch of the inverse builder

    const int cw = (warp == 0) ? 0 : (warp == 2) ? 1 : (warp == 3) ? 2 : (warp == 5) ? 3 : -1;   // compute rank
    const int cr = (warp == 6) ? 0 : (cw >= 1) ? cw : -1;      // row block of the L(k,k-1) products (not warp 0: it is on the chain; not warp 4: same scheduler)
    const int iw = (warp == 4) ? 0 : (warp == 7) ? 1 : -1;                                       // I/O rank
    const int it = iw * 32 + lane;                                                               // I/O thread id (0..63)
    double a3[2] = {0.0, 0.0}, pre0[2] = {0.0, 0.0}, dold[2] = {0.0, 0.0};   // carried from a panel's pre-phase to its final step

    load_ab_tile(a, 0, 0, Dbuf, S33, tid, THREADS);
    for (int e = tid; e < T36; e += THREADS) Lcol[e] = qnan;
    if (tid < NB) dinvs[tid] = qnan;
    for (int e = tid; e < T36; e += THREADS) Xbuf[e] = 0.0;
    // inverse tiles: zero once (upper blocks and pad columns never change; lower blocks are overwritten whole every panel),
    // diagonal blocks NaN-armed per panel by the inverse builder itself before it arrives at the panel's barrier
    for (int e = tid; e < 2 * T36; e += THREADS) {
        const int r = (e % T36) / S36, c = (e % T36) % S36;
        Linv[e] = (e < T36 && c < NB && (r >> 3) == (c >> 3)) ? qnan : 0.0;
    }

    bool linv_failed = false, tinv_failed = false;
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    long long w1acc = 0;
    long long wacc[5] = {0, 0, 0, 0, 0};
    long long tsacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64(), t1;
    long long arrive_acc = 0, tA_prev = 0;
    __shared__ int linv_row_ready[4];      // panel index + 1 once block row b of that panel's inverse is complete in shared memory
    if (tid < 4) linv_row_ready[tid] = 0;
    __shared__ long long pe_smem[2];       // profiling: clock at the end of the last Cholesky
    const bool prof = (DBGF(a) & 4) != 0;
    // BAR.SYNC is issued "defer blocking": a clock read right behind it executes before the barrier completes.
    // The volatile shared load below cannot, and the clock read is made control-dependent on its value.
#define PROF(slot) do { if (prof) { if (*(volatile double*)dinvs != 1.2345e300) t1 = clock64(); tacc[slot] += t1 - t0; t0 = t1; } } while (0)
    if (prof && tid == 0) { a.prof[8] = clock64(); for (int q = 16; q < 24; ++q) a.prof[q] = 0; }
    // partial mode (KE < NP): one more, reduced, step k = KE forms L(KE,KE-1) and the un-damped D(KE,KE) and writes them back.
    // The pivot chain and the inverse builder run their OWN loops over the panels (everything between the warps goes through
    // named barriers and values in shared memory): the chain's code is a few KB that stay in the instruction cache, instead of
    // one body of ~50 KB that every warp jumps through.
    if (warp == 0) {
        for (int k = 0; k < NP && k <= KE; ++k) {
            const bool tail = (k == KE);
            double* D = Dbuf + (k & 1) * T33;
            PROF(0);
            bar_sync(7, THREADS);                          // this panel's operands are in shared memory (see the barrier below)
            PROF(1);
            if (prof && lane == 0) pe_smem[1] = clock64();
            const long long tA = t0;
            bar_sync(9, 160);                              // block column 0 of D is final
            PROF(2);
            PROF(3);
            if (tail) {
                bar_sync(10, 128);
                for (int e = lane; e < T32; e += 128) {    // this warp's quarter of the Schur complement, back to the band storage
                    const int r = e >> 5, c = e & 31, i = NB * k + r, j = NB * k + c;
                    if (c <= r && in_band(a, i, j)) __stcg(ab_at(a, i, j), D[r * S33 + c]);
                }
            } else {
                if (warp_potrf_blocked(D, Lcol + (k & 1) * T36, dinvs + (k & 1) * NB, Linv + (k & 1) * T36, lane,
                                       prof ? tsacc : nullptr, tA))
                    *a.info = 1;
                if (prof && lane == 0) pe_smem[0] = clock64();
                if (lane == 0) GSTAMP(a, 0, k);
                PROF(4);
            }
        }
    } else if (warp == 1) {
        // ---- L(k,k)^-1 behind the Cholesky, into shared memory (its diagonal blocks feed the next panel's products) and global
        // memory (U, R).  The moment the Cholesky is complete it arms the next panel's tile and ARRIVES at that panel's
        // barrier; the last block row's products, its publication and the flag follow off the panel-to-panel path. ----
        bar_arrive(7, THREADS);                            // panel 0's barrier (covers the initialisation above)
        for (int k = 0; k < NP && k < KE; ++k) {
            double* Lc = Lcol + (k & 1) * T36;
            double* dv = dinvs + (k & 1) * NB;
            double* Ls = Linv + (k & 1) * T36;
            const long long w0 = prof ? clock64() : 0;
            auto on_last = [&]() {
                if (NB * k + lane < n) a.dinv[NB * k + lane] = dv[lane];
                if (k + 1 < NP && k + 1 <= KE) {
                    double* Ln = Linv + ((k + 1) & 1) * T36;            // next panel's tile: its last reader was this panel's final step
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int e = q * 32 + lane, bb = e >> 6, r = (e >> 3) & 7, c = e & 7;
                        Ln[(8 * bb + r) * S36 + 8 * bb + c] = qnan;
                    }
                    if (prof) arrive_acc += clock64() - *(volatile long long*)pe_smem;
                    // After a failure nothing paces this warp any more (every wait returns at once): it then WAITS at the barrier
                    if (linv_failed) bar_sync(7, THREADS); else bar_arrive(7, THREADS);
                }
            };
            warp_linv_blocked(Lc, dv, Ls, a.LI + (size_t)k * T32, Sc, lane, linv_failed, linv_row_ready, k + 1, on_last,
                              prof ? wacc : nullptr, prof ? tsacc : nullptr, w0);
            if (prof) w1acc += clock64() - w0;
            if (linv_failed && lane == 0) *a.info = 1;
            // last block row of the inverse -> global, then the flag R waits for (covers dinv too; U takes the tile by value)
            publish_inv_rows(Ls, a.LI + (size_t)k * T32, 3, lane);
            __syncwarp();
            if (lane == 0 && !(DBGF(a) & 8)) red_release(diag_done + k, 1);
        }
    } else {
    for (int k = 0; k < NP && k <= KE; ++k) {
        const bool tail = (k == KE);
        double* D = Dbuf + (k & 1) * T33;
        double* X = Xbuf + (k & 1) * T36;
        double* Lx = Lxbuf + (k & 1) * T36;
        const double* Lp = Linv + ((k + 1) & 1) * T36;     // L(k-1,k-1)^-1
        double* Lc = Lcol + (k & 1) * T36;
        double* dv = dinvs + (k & 1) * NB;
        // top-of-panel barrier: staged tiles + previous inverse are in shared memory.  The inverse builder only ARRIVES (at the
        // end of its previous panel, the moment the inverse tile is complete in shared memory): it publishes the tile's last
        // block row and raises diag_done while the other seven warps are already in this panel's products.
        // Compute warps (cr >= 0) have done everything of this panel's L(k,k-1) and D(k) that does not need the last 8 columns
        // of L(k-1,k-1) during the previous panel (PRE-PHASE below), and only arrive here.
        if (prof && k >= 2) {      // arrivals AFTER the last Cholesky's end: how many, how late in all
            const long long dt = clock64() - *(volatile long long*)pe_smem;
            if (dt < 4000) { arrive_acc += dt; tA_prev += 1; }
        }
        if (k == 0 || cr < 0) bar_sync(7, THREADS); else bar_arrive(7, THREADS);
        if (cr >= 0) {
            // ---- FINAL STEP of L(k,k-1) = A(k,k-1) L(k-1,k-1)^-T and of block column 0 of D = A(k,k) - L L^T (+ u) ----
            // Forward-substitution form by 8-column blocks: Y_t = (X_t - sum_{s<t} Y_s L_ts^T) T_t^T with T_t = L_tt^-1 straight
            // from the Cholesky warp.  W_3 = X_3 - sum_{s<3} Y_s L_3s^T (operand registers a3) and the products over columns
            // 0..23 (pre0) were formed while the chain was still on block 3: two DMMAs, one exchange, two DMMAs remain.
            const int fr = lane >> 2, fc = lane & 3, rb = cr;
            double v0, v1;
            if (k >= 1) {
                const volatile double* T = Lp;
                int spins = 0;
                double dq = T[31 * S36 + 31];
                while (dq != dq && ++spins < (1 << 20)) dq = T[31 * S36 + 31];
                asm volatile("" ::: "memory");
                const long long f0 = (prof && warp == 3) ? clock64() : 0;
                double y0 = 0.0, y1 = 0.0;
                dmma884(y0, y1, a3[0], T[(24 + fr) * S36 + 24 + fc]);
                dmma884(y0, y1, a3[1], T[(24 + fr) * S36 + 28 + fc]);
                *reinterpret_cast<double2*>(Lx + (8 * rb + fr) * S36 + 24 + 2 * fc) = make_double2(y0, y1);
                bar_sync(2, 128);                          // Y_3 of all four row blocks
                if (prof && warp == 3 && *(volatile double*)Lx != 1.2345e300) tsacc[4] += clock64() - f0;
                bar_arrive(1, 192);                        // I/O warps may store L(k,k-1)
                double c0 = 0.0, c1 = 0.0;
                dmma884(c0, c1, Lx[(8 * rb + fr) * S36 + 24 + fc], Lx[fr * S36 + 24 + fc]);
                dmma884(c0, c1, Lx[(8 * rb + fr) * S36 + 28 + fc], Lx[fr * S36 + 28 + fc]);
                v0 = dold[0] - (pre0[0] + c0);
                v1 = dold[1] - (pre0[1] + c1);
                if (prof && warp == 3 && v0 != 1.2345e300) tsacc[5] += clock64() - f0;
            } else {
                bar_arrive(1, 192);
                v0 = D[(8 * rb + fr) * S33 + 2 * fc];
                v1 = D[(8 * rb + fr) * S33 + 2 * fc + 1];
            }
            {
                const int i = 8 * rb + fr, j = 2 * fc, gi = NB * k + i;
                if (!tail && i == j) v0 = (gi < n) ? v0 + u : 1.0;
                if (!tail && i == j + 1) v1 = (gi < n) ? v1 + u : 1.0;
                D[i * S33 + j] = v0;
                D[i * S33 + j + 1] = v1;
            }
            bar_arrive(9, 160);                            // block column 0 of D is final: the pivot chain may start
        }
        if (cw >= 1) {
            // helper h = cw owns blocks (h,1) [syrk only] and (2,2) | (3,2) | (3,3): rest of the syrk, then the
            // non-urgent part of the Cholesky's trailing updates (block columns right of the next one).
            // One barrier id per phase: a helper may reach its next arrive before warp 0 has consumed the previous one.
            const int oi = (cw == 1) ? 2 : 3, oj = (cw == 3) ? 3 : 2;
            syrk_blocks(Lx, D, k >= 1, cw, 1, oi, oj, k, n, u, lane, !tail);
            if (tail) {
                bar_sync(10, 128);
            } else {
                bar_arrive(4, 128);                                   // phase 0: warp 0 may update block column 1
                helper_trailing(D, Lc, dv, 0, oi, oj, lane, true);
                bar_arrive(5, 128);                                   // phase 1: ... block column 2
                helper_trailing(D, Lc, dv, 1, oi, oj, lane, cw == 3);
                bar_arrive(6, 128);                                   // phase 2: ... block (3,3)
            }
        }
        if (tail && cw >= 1) {
            // Schur complement of the first non-eliminated panel goes back to the band storage
            for (int e = cw * 32 + lane; e < T32; e += 128) {
                const int r = e >> 5, c = e & 31, i = NB * k + r, j = NB * k + c;
                if (c <= r && in_band(a, i, j)) __stcg(ab_at(a, i, j), D[r * S33 + c]);
            }
        }
        const bool has_next = (k + 1 < NP && !tail);
        if (iw >= 0) {
            // ---- I/O warps (4 and 7): NaN-arm the other column buffer for panel k+1 (its last reader finished before the
            // barrier), store L(k,k-1), raise rows_done ----
            double* Ln = Lcol + ((k + 1) & 1) * T36;
            if (!(DBGF(a) & 32) && !tail) for (int e = it; e < T36; e += 64) Ln[e] = qnan;
            if (it < NB) dinvs[((k + 1) & 1) * NB + it] = qnan;
            bar_sync(1, 192);                                              // L(k,k-1) is complete
            if (k >= 1 && WB >= 1) {
                store_g_tile(lb_tile(a, k, 1), Lx, S36, it, 64);
                if (iw == 1) {
                    bar_sync(3, 64);
                    if (lane == 0 && !(DBGF(a) & 8)) red_release(rows_done + (k - 1), 1);
                } else {
                    bar_arrive(3, 64);
                }
            }
        }
        if (has_next) {
            // ---- staging of panel k+1's two tiles by three warps (22 loads per thread, all in flight at once) ----
            if (cr <= 0) {      // warps 4, 7 and 6: the helpers are still behind the Cholesky's block columns
                const int lt = (cr == 0 ? 2 : iw) * 32 + lane;
                if (prof && warp == 7) tsacc[0] += clock64() - *(volatile long long*)(pe_smem + 1);
                if (k >= 1 && NU > 0 && hot_expected == 2)      // both tiles' last update was panel k-1's: by value from their owners
                    load_hot_tiles2<96>(a, a.HM + (size_t)(2 * (k - 1) + 1) * T32, a.HM + (size_t)(2 * (k - 1)) * T32, k + 1, k + 1,
                                        Dbuf + ((k + 1) & 1) * T33, S33, k + 1, k, Xbuf + ((k + 1) & 1) * T36, S36, lt);
                else
                    load_ab_tiles2<96>(a, k + 1, k + 1, Dbuf + ((k + 1) & 1) * T33, S33, k + 1, k, Xbuf + ((k + 1) & 1) * T36, S36, lt);
            }
            if (prof && warp == 7) tsacc[1] += clock64() - *(volatile long long*)(pe_smem + 1);
            if (warp == 7 && lane == 0) GSTAMP(a, 4, k - 1);
            if (cr < 0) {
                bar_arrive(8, 192);
            } else {
                // ---- PRE-PHASE of panel k+1 (see the final step above), behind the Cholesky of panel k ----
                bar_sync(8, 192);
                if (prof && warp == 3) tsacc[1] += clock64() - *(volatile long long*)(pe_smem + 1);
                const int fr = lane >> 2, fc = lane & 3, rb = cr;
                const double* Xn = Xbuf + ((k + 1) & 1) * T36;
                const double* Dn = Dbuf + ((k + 1) & 1) * T33;
                double* Lxn = Lxbuf + ((k + 1) & 1) * T36;
                const volatile double* T = Linv + (k & 1) * T36;      // T_t in the diagonal blocks (this panel's Cholesky writes them)
                const volatile double* Lk = Lc;                        // L(k,k) by columns
                double acc[4][2];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double2 x2 = *reinterpret_cast<const double2*>(Xn + (8 * rb + fr) * S36 + 8 * t + 2 * fc);
                    acc[t][0] = x2.x; acc[t][1] = x2.y;
                }
                dold[0] = Dn[(8 * rb + fr) * S33 + 2 * fc];
                dold[1] = Dn[(8 * rb + fr) * S33 + 2 * fc + 1];
                // block 0 needs no conversion: W_0 = X_0, operands straight from the staged tile
                double w0 = Xn[(8 * rb + fr) * S36 + fc], w1 = Xn[(8 * rb + fr) * S36 + 4 + fc];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    int spins = 0;
                    double dq = T[(8 * t + 7) * S36 + 8 * t + 7];
                    while (dq != dq && ++spins < (1 << 20)) dq = T[(8 * t + 7) * S36 + 8 * t + 7];
                    asm volatile("" ::: "memory");
                    if (prof && warp == 3 && t == 1) tsacc[0] += clock64() - *(volatile long long*)(pe_smem + 1);
                    if (prof && warp == 3 && t == 2) tsacc[3] += clock64() - *(volatile long long*)(pe_smem + 1);
                    if (t > 0) {   // W_t: accumulator layout -> operand layout through this warp's own rows of the L(k+1,k) tile
                        *reinterpret_cast<double2*>(Lxn + (8 * rb + fr) * S36 + 8 * t + 2 * fc) = make_double2(acc[t][0], acc[t][1]);
                        __syncwarp();
                        w0 = Lxn[(8 * rb + fr) * S36 + 8 * t + fc]; w1 = Lxn[(8 * rb + fr) * S36 + 8 * t + 4 + fc];
                        __syncwarp();
                    }
                    double y0 = 0.0, y1 = 0.0;
                    dmma884(y0, y1, w0, T[(8 * t + fr) * S36 + 8 * t + fc]);
                    dmma884(y0, y1, w1, T[(8 * t + fr) * S36 + 8 * t + 4 + fc]);
                    *reinterpret_cast<double2*>(Lxn + (8 * rb + fr) * S36 + 8 * t + 2 * fc) = make_double2(y0, y1);
                    __syncwarp();
                    const double ya = -Lxn[(8 * rb + fr) * S36 + 8 * t + fc], yb = -Lxn[(8 * rb + fr) * S36 + 8 * t + 4 + fc];
                    // the NEXT block's accumulator first: it is the one the chain of the substitution waits for
#pragma unroll
                    for (int uu = t + 1; uu < 4; ++uu) {

                        dmma884(acc[uu][0], acc[uu][1], ya, Lk[(8 * t + fc) * S36 + 8 * uu + fr]);
                        dmma884(acc[uu][0], acc[uu][1], yb, Lk[(8 * t + 4 + fc) * S36 + 8 * uu + fr]);
                    }
                }
                if (prof && warp == 3) tsacc[2] += clock64() - *(volatile long long*)(pe_smem + 1);
                // W_3 as operand registers
                *reinterpret_cast<double2*>(Lxn + (8 * rb + fr) * S36 + 24 + 2 * fc) = make_double2(acc[3][0], acc[3][1]);
                __syncwarp();
                a3[0] = Lxn[(8 * rb + fr) * S36 + 24 + fc];
                a3[1] = Lxn[(8 * rb + fr) * S36 + 28 + fc];
                bar_sync(2, 128);                          // Y_0..2 of all four row blocks
                pre0[0] = pre0[1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < 6; ++ks)
                    dmma884(pre0[0], pre0[1], Lxn[(8 * rb + fr) * S36 + 4 * ks + fc], Lxn[fr * S36 + 4 * ks + fc]);
            }
        }
    }
    }
    __syncthreads();
    if (prof && tid == 0) {
        a.prof[9] = clock64();
        for (int q = 0; q < 5; ++q) a.prof[q] = tacc[q];
    }
    if (prof && warp == 1 && lane == 0) a.prof[30] = w1acc;
    if (prof && warp == 1 && lane == 0) for (int q = 0; q < 5; ++q) a.prof[31 + q] = wacc[q];
    if (prof && lane == 0 && warp == 1) for (int q = 0; q < 4; ++q) a.prof[44 + q] = tsacc[q];
    if (prof && lane == 0 && warp == 3) for (int q = 0; q < 4; ++q) a.prof[58 + q] = tsacc[q];
    if (prof && lane == 0 && warp == 3) { a.prof[62] = tsacc[4]; a.prof[63] = tsacc[5]; }
    if (prof && lane == 0 && warp == 7) for (int q = 0; q < 2; ++q) a.prof[6 + q] = tsacc[q];
    if (prof && lane == 0) a.prof[36 + warp] = (warp == 1) ? arrive_acc : (tA_prev << 32) | arrive_acc;
    if (prof && tid == 0) for (int q = 4; q < 7; ++q) a.prof[44 + q] = tsacc[q];
#undef PROF
#undef IOPROF
}

// Push-style back substitution L^T x = y on one CTA; s (shared, NP*32) holds y on entry and x on return.
// Panels k >= ke are NOT solved (their x is given in s: the middle block of the two-sided solve) but still pushed.
__device__ void backsub(const Args3& a, double* s, int ke) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = a.NP, WB = a.WB;
    // ---- back substitution L^T x = y, push style.  Register-prefetched operands: job A = tile d = warp+1 (all
    // warps), job B = L(k,k)^-1 on warp 0 / tile d = warp+8 on warps 1..7; wider bands go on demand. ----
    double pa[NB], pb[NB];
    const int dA = warp + 1, dB = warp + 8;
    auto prefetch = [&](int k) {
        const int nd = min(WB, k);
        if (dA <= nd) {
            const double* G = lb_tile(a, k, dA);
#pragma unroll
            for (int r = 0; r < NB; ++r) pa[r] = __ldcg(G + r * NB + lane);
        }
        if (warp == 0 || dB <= nd) {
            const double* G = (warp == 0) ? a.LI + (size_t)k * T32 : lb_tile(a, k, dB);
#pragma unroll
            for (int r = 0; r < NB; ++r) pb[r] = __ldcg(G + r * NB + lane);
        }
    };
    __syncthreads();
    prefetch(NP - 1);
    for (int k = NP - 1; k >= 0; --k) {
        const double* sk = s + NB * k;
        if (warp == 0 && k < ke) {             // x_k = L(k,k)^-T s_k
            double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
            for (int r = 0; r < NB; r += 4) {
                x0 = fma(pb[r], sk[r], x0);
                x1 = fma(pb[r + 1], sk[r + 1], x1);
                x2 = fma(pb[r + 2], sk[r + 2], x2);
                x3 = fma(pb[r + 3], sk[r + 3], x3);
            }
            __syncwarp();
            s[NB * k + lane] = (x0 + x1) + (x2 + x3);
        }
        __syncthreads();
        const int nd = min(WB, k);
        if (dA <= nd && k - dA < ke) {         // s_{k-d} -= L(k,k-d)^T x_k   (targets that are given, not solved, stay)
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int r = 0; r < NB; r += 2) {
                c0 = fma(pa[r], sk[r], c0);
                c1 = fma(pa[r + 1], sk[r + 1], c1);
            }
            s[NB * (k - dA) + lane] -= c0 + c1;
        }
        if (warp != 0 && dB <= nd && k - dB < ke) {
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int r = 0; r < NB; r += 2) {
                c0 = fma(pb[r], sk[r], c0);
                c1 = fma(pb[r + 1], sk[r + 1], c1);
            }
            s[NB * (k - dB) + lane] -= c0 + c1;
        }
        for (int d = 16 + warp; d <= nd && k - d < ke; d += 8) {     // wide bands: remaining tiles on demand
            const double* G = lb_tile(a, k, d);
            double c0 = 0.0;
            for (int r = 0; r < NB; ++r) c0 = fma(__ldcg(G + r * NB + lane), sk[r], c0);
            s[NB * (k - d) + lane] -= c0;
        }
        if (k >= 1) prefetch(k - 1);
        __syncthreads();
    }
}

// =====================================================================================================
// R: forward substitution behind P, then the back substitution (same CTA, no grid barrier in between).
// =====================================================================================================
__device__ void role_R(const Args3& a, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, NP = a.NP, WB = a.WB;
    int* diag_done = a.flags;
    int* rows_done = a.flags + NP;
    double* s = smem;                          // NP*NB running right-hand side -> y -> x
    double* LinvS = s + (size_t)NP * NB;       // T33
    double* ys = LinvS + T33;                  // NB
    double* Lrows = ys + NB;                   // 8 x T33

    if (DBGF(a) & 8) return;
    for (int i = tid; i < NP * NB; i += THREADS) s[i] = (i < n) ? __ldcg(a.g + i) : 0.0;
    for (int p = 0; p < NP && p < a.ke; ++p) {
        cta_wait(diag_done + p, 1);
        load_g_tile(a.LI + (size_t)p * T32, LinvS, S33, tid, THREADS);
        __syncthreads();
        if (warp == 0) {                       // y_p = L(p,p)^-1 s_p
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int c = 0; c < NB; c += 2) {
                y0 = fma(LinvS[lane * S33 + c], s[NB * p + c], y0);
                y1 = fma(LinvS[lane * S33 + c + 1], s[NB * p + c + 1], y1);
            }
            __syncwarp();
            ys[lane] = y0 + y1;
            s[NB * p + lane] = y0 + y1;
        }
        const int nrows = min(NP - 1, p + WB) - p;
        if (nrows > 0) {
            if (DBGF(a) & 1) __syncthreads(); else cta_wait(rows_done + p, nrows);    // also orders ys
            for (int q0 = 0; q0 < nrows; q0 += 8) {
                const int nq = min(8, nrows - q0);
                {   // the batch's tiles with EVERY load in flight before the first shared store (one L2 round trip per batch
                    // instead of one per tile: this role's lag behind the pivot chain is the kernel's tail)
                    double v[8][4];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double* G = lb_tile(a, p + 1 + q0 + min(q, nq - 1), q0 + min(q, nq - 1) + 1);
#pragma unroll
                        for (int x = 0; x < 4; ++x) v[q][x] = (q < nq) ? __ldcg(G + x * THREADS + tid) : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q)
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const int e = x * THREADS + tid;
                            if (q < nq) Lrows[(size_t)q * T33 + (e >> 5) * S33 + (e & 31)] = v[q][x];
                        }
                }
                __syncthreads();
                const int q = tid >> 5;
                if (q < nq) {                  // s_i -= L(i, panel p) . y_p
                    const double* Lr = Lrows + (size_t)q * T33 + lane * S33;
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int c = 0; c < NB; c += 2) {
                        s0 = fma(Lr[c], ys[c], s0);
                        s1 = fma(Lr[c + 1], ys[c + 1], s1);
                    }
                    s[NB * (p + 1 + q0 + q) + lane] -= s0 + s1;
                }
                __syncthreads();
            }
        } else {
            __syncthreads();
        }
    }

    if ((DBGF(a) & 4) && tid == 0) a.prof[10] = clock64();
    if (DBGF(a) & 2) {
        for (int i = tid; i < n; i += THREADS) a.g[i] = s[i];
        return;
    }
    if (a.ke < NP) {                       // partial factorisation: y (eliminated panels) and the updated right-hand side below
        __syncthreads();
        for (int i = tid; i < n; i += THREADS) a.g[i] = s[i];
        return;
    }
    backsub(a, s, NP);
    for (int i = tid; i < n; i += THREADS) a.g[i] = s[i];
    if ((DBGF(a) & 4) && tid == 0) a.prof[11] = clock64();
}

// =====================================================================================================
// U: trailing update.  Tile t of panel p's trailing triangle goes to CTA (t + p) % NU; a tile costs three
// DMMA products: L(I,p) = A(I,p) Linv^T, L(J,p) = A(J,p) Linv^T, A(I,J) -= L(I,p) L(J,p)^T.  The diagonal tile
// (I,I) also stores L(I,p) for the substitutions.
// =====================================================================================================
__device__ void role_U(const Args3& a, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = a.NP, WB = a.WB, bw = a.bw;
    int* diag_done = a.flags;
    int* rows_done = a.flags + NP;
    int* upd_done = a.flags + 2 * NP;
    const int NU = a.ncta - 2, ui = (int)blockIdx.x - a.rank0 - 2;
    double* LinvS = smem;            // T36
    double* As = LinvS + T36;
    double* Bs = As + T36;
    double* LIs = Bs + T36;
    double* LJs = LIs + T36;
    const int fr = lane >> 2, fc = lane & 3;
    const bool uprof = (DBGF(a) & 4) && ui == 1;
    long long ut[6] = {0, 0, 0, 0, 0, 0}, u0 = clock64(), u1;
#define UPROF(slot) do { if (uprof) { if (*(volatile double*)LinvS != 1.2345e300) u1 = clock64(); ut[slot] += u1 - u0; u0 = u1; } } while (0)

    for (int p = 0; p < NP && p < a.ke; ++p) {
        const int last = min(NP - 1, p + WB);
        const int nrows = last - p;
        const int ntiles = nrows * (nrows + 1) / 2;
        int t = ((ui - p) % NU + NU) % NU;
        bool first = true;
        int rows_written = 0;
        // The off-diagonal tile P stages next, (p+2,p+1), sits on the solve's critical loop (inverse -> this update -> P's
        // staging -> P's products), and a whole tile is 288 DMMAs on a role that is DMMA-issue-bound.  When update CTAs are idle
        // in this panel, four of them take one 8-row block of it each (the tile's own CTA + three idle ones): 20 + 80 + 32
        // DMMAs instead of 288.
        const bool split = (ntiles >= 3) && (ntiles + 3 <= NU) && !(DBGF(a) & 1);
        const int part = !split ? -1 : (t == 1) ? 0 : (t >= ntiles && t < ntiles + 3) ? t - ntiles + 1 : -1;
        if (part >= 0) {
            const int I = p + 2, J = p + 1, rb = part;
            if (p >= 1) cta_wait(upd_done + (p - 1), NU);
            else __syncthreads();
            load_ab_tiles2<THREADS>(a, I, p, As, S36, J, p, Bs, S36, tid);
            double oldv[2] = {0.0, 0.0};
            if (warp < 4) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = NB * I + 8 * rb + fr, j = NB * J + 8 * warp + 2 * fc + e;
                    oldv[e] = in_band(a, i, j) ? __ldcg(ab_at(a, i, j)) : 0.0;
                }
            }
            if (part == 0 && tid == 0) GSTAMP(a, 1, p);
            load_g_tile_polled(a.LI + (size_t)p * T32, LinvS, S36, tid, THREADS, a.info);
            if (part == 0 && tid == 0) GSTAMP(a, 2, p);
            first = false;
            __syncthreads();
            if (warp < 4) trsm_strip(As, LinvS, LIs, rb, 1u << warp, lane);     // L(I,p), my 8 rows: one block column per warp
            else trsm_strip(Bs, LinvS, LJs, warp - 4, 0xfu, lane);              // L(J,p), all of it
            __syncthreads();
            if (warp < 4) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    dmma884(c0, c1, LIs[(8 * rb + fr) * S36 + 4 * ks + fc], LJs[(8 * warp + fr) * S36 + 4 * ks + fc]);
                double* hm = a.HM + (size_t)(2 * p) * T32;
                const int i = NB * I + 8 * rb + fr, j = NB * J + 8 * warp + 2 * fc;
                const double w0 = oldv[0] - c0, w1 = oldv[1] - c1;
                if (in_band(a, i, j)) { __stcg(ab_at(a, i, j), w0); __stcg(hm + (8 * rb + fr) * NB + 8 * warp + 2 * fc, w0); }
                if (in_band(a, i, j + 1)) { __stcg(ab_at(a, i, j + 1), w1); __stcg(hm + (8 * rb + fr) * NB + 8 * warp + 2 * fc + 1, w1); }
            }
            if (part == 0 && tid == 0) GSTAMP(a, 3, p);
            t = ntiles;                                                   // nothing else for this CTA in this panel
        }
        for (; t < ntiles && !(DBGF(a) & 1); t += NU) {
            if (t == 0) continue;                                         // tile (p+1,p+1) belongs to P
            int ri = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
            while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
            while (ri * (ri + 1) / 2 > t) --ri;
            const int I = p + 1 + ri, J = p + 1 + (t - ri * (ri + 1) / 2);
            if (NB * (I - J) - (NB - 1) > bw) continue;                   // entirely outside the band
            const bool diag = (I == J);
            UPROF(5);
            if (first && p >= 1) cta_wait(upd_done + (p - 1), NU);       // every tile carries panels <= p-1
            else __syncthreads();                                        // shared buffers of the previous tile are free
            UPROF(0);
            load_ab_tiles2<THREADS>(a, I, p, As, S36, diag ? -1 : J, p, Bs, S36, tid);
            // old values of my two 8x8 output blocks (fragment layout), issued with the operand loads
            const int bi = warp >> 1;
            double oldv[2][2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int bj = 2 * (warp & 1) + q;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = NB * I + 8 * bi + fr, j = NB * J + 8 * bj + 2 * fc + e;
                    oldv[q][e] = in_band(a, i, j) ? __ldcg(ab_at(a, i, j)) : 0.0;
                }
            }
            if (first) {
                UPROF(1);
                if (t == 1 && tid == 0) GSTAMP(a, 1, p);
                load_g_tile_polled(a.LI + (size_t)p * T32, LinvS, S36, tid, THREADS, a.info);   // no flag: the tile is NaN until written
                UPROF(2);
                if (t == 1 && tid == 0) GSTAMP(a, 2, p);
                first = false;
            }
            __syncthreads();
            if (diag) {
                trsm_strip(As, LinvS, LIs, warp >> 1, (warp & 1) ? 0x6u : 0x9u, lane);
            } else if (warp < 4) {
                trsm_strip(As, LinvS, LIs, warp, 0xfu, lane);
            } else {
                trsm_strip(Bs, LinvS, LJs, warp - 4, 0xfu, lane);
            }
            __syncthreads();
            const double* Lb = diag ? LIs : LJs;
            {
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const int bj0 = 2 * (warp & 1);
                const bool skip1 = diag && (bj0 + 1 > bi), skip0 = diag && (bj0 > bi);   // above the diagonal: not stored
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double af = LIs[(8 * bi + fr) * S36 + 4 * ks + fc];
                    if (!skip0) dmma884(acc[0][0], acc[0][1], af, Lb[(8 * bj0 + fr) * S36 + 4 * ks + fc]);
                    if (!skip1) dmma884(acc[1][0], acc[1][1], af, Lb[(8 * (bj0 + 1) + fr) * S36 + 4 * ks + fc]);
                }
                // (p+2,p+1), (p+2,p+2) are what P stages for panel p+2: they also go, by value, into its mailbox
                double* hm = (t == 1 || t == 2) ? a.HM + (size_t)(2 * p + (t - 1)) * T32 : nullptr;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q == 0 ? skip0 : skip1) continue;
                    const int i = NB * I + 8 * bi + fr, j = NB * J + 8 * (bj0 + q) + 2 * fc;
                    const double w0 = oldv[q][0] - acc[q][0], w1 = oldv[q][1] - acc[q][1];
                    if (in_band(a, i, j)) { __stcg(ab_at(a, i, j), w0); if (hm) __stcg(hm + (8 * bi + fr) * NB + 8 * (bj0 + q) + 2 * fc, w0); }
                    if (in_band(a, i, j + 1)) { __stcg(ab_at(a, i, j + 1), w1); if (hm) __stcg(hm + (8 * bi + fr) * NB + 8 * (bj0 + q) + 2 * fc + 1, w1); }
                }
            }
            if (t == 1 && tid == 0) GSTAMP(a, 3, p);
            if (diag) {
                store_g_tile(lb_tile(a, I, I - p), LIs, S36, tid, THREADS);
                ++rows_written;
            }
        }
        __syncthreads();
        UPROF(3);
        if (tid == 0) {
            // one release fence for both counters
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(upd_done + p), "r"(1) : "memory");
            if (rows_written)
                asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(rows_done + p), "r"(rows_written) : "memory");
        }
        UPROF(4);
    }
    if (uprof && tid == 0) for (int q = 0; q < 6; ++q) a.prof[52 + q] = ut[q];
#undef UPROF
}

// U for wide bands.  With more trailing tiles per panel than CTAs, recomputing the two "triangular solves" for
// every tile triples the work (C3/C5: 300 tiles on 146 CTAs), and a barrier over all update CTAs per panel serialises
// the panels.  Here every tile (I,J) has a FIXED owner CTA for its whole life (tile-owner data flow):
//   panel p, phase 1: the owner of (I,p) -- which applied all its updates itself, in program order -- forms
//                     L(I,p) = A(I,p) L(p,p)^-T once and stores it for everybody (rows_done[p]);
//   panel p, phase 2: the owner of (I,J), J > p, applies A(I,J) -= L(I,p) L(J,p)^T from the stored tiles (one product),
//                     next block column first.
// The only cross-CTA dependencies are diag_done[p] and rows_done[p] (+ the hot flag for P): no per-panel barrier.
// I + (WB+1) J is injective over a panel's trailing window (<= WB rows and columns), so a panel's tiles spread almost evenly
// over the CTAs: worst CTA 4.5 tile-equivalents instead of 6.5 with the former 7 I + 13 J at bw 640 on 72 CTAs (mean 3.3)
__device__ __forceinline__ int tile_owner(int I, int J, int NU, int WB) {
    return (int)(((unsigned)I + (unsigned)(WB + 1) * (unsigned)J) % (unsigned)NU);
}

__device__ void role_U2(const Args3& a, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = a.NP, WB = a.WB, bw = a.bw;
    int* diag_done = a.flags;
    int* rows_done = a.flags + NP;
    int* upd_done = a.flags + 2 * NP;
    const int NU = a.ncta - 2, ui = (int)blockIdx.x - a.rank0 - 2;
    double* LinvS = smem;            // T36
    double* As = LinvS + T36;
    double* LIs = As + T36;
    double* LJs = LIs + T36;
    __shared__ int lst1[MAX_WB3 + 4], lst2[448], n1, n2;
    const int fr = lane >> 2, fc = lane & 3;

    for (int p = 0; p < NP && p < a.ke; ++p) {
        const int last = min(NP - 1, p + WB);
        const int nrows = last - p;
        // ---- my work of this panel: rows (phase 1) and tiles (phase 2, packed I<<16|J, sorted by J then I) ----
        __syncthreads();
        if (tid == 0) { n1 = 0; n2 = 0; }
        __syncthreads();
        for (int c = tid; c < nrows * (nrows + 1) / 2 + nrows; c += THREADS) {
            if (c < nrows) {                                   // candidate row I = p+1+c of column p (row p+1 is P's)
                const int I = p + 1 + c;
                if (c >= 1 && tile_owner(I, p, NU, WB) == ui) lst1[atomicAdd(&n1, 1)] = I;
            } else {
                const int t = c - nrows;
                int ri = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
                while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
                while (ri * (ri + 1) / 2 > t) --ri;
                const int I = p + 1 + ri, J = p + 1 + (t - ri * (ri + 1) / 2);
                if (t != 0 && NB * (I - J) - (NB - 1) <= bw && tile_owner(I, J, NU, WB) == ui) lst2[atomicAdd(&n2, 1)] = (J << 16) | I;
            }
        }
        __syncthreads();
        if (tid == 0) {                                        // few entries: insertion sort by (J, I)
            for (int x = 1; x < n2; ++x) {
                const int v = lst2[x];
                int y = x - 1;
                while (y >= 0 && lst2[y] > v) { lst2[y + 1] = lst2[y]; --y; }
                lst2[y + 1] = v;
            }
        }
        __syncthreads();
        const int m1 = n1, m2 = n2;
        // ---- phase 1 ----
        if (m1 > 0) {
            load_g_tile_polled(a.LI + (size_t)p * T32, LinvS, S36, tid, THREADS, a.info);
            for (int x = 0; x < m1; ++x) {
                const int I = lst1[x];
                __syncthreads();
                load_ab_tiles2<THREADS>(a, I, p, As, S36, -1, p, As, S36, tid);
                __syncthreads();
                trsm_strip(As, LinvS, LIs, warp >> 1, (warp & 1) ? 0x6u : 0x9u, lane);
                __syncthreads();
                store_g_tile(lb_tile(a, I, I - p), LIs, S36, tid, THREADS);
            }
            __syncthreads();
            if (tid == 0) red_release(rows_done + p, m1);
        }
        // ---- phase 2 ----
        if (m2 > 0) {
            cta_wait(rows_done + p, nrows);                     // every L(.,p) is stored
            // Software pipeline over this CTA's tiles: the operands of tile x+1 (two L tiles, 4 values per thread each, and
            // the old values of its output fragments) are in flight in registers while tile x is multiplied; the role was
            // bound by three dependent L2 round trips per tile (~2.5 k cycles) around ~1 k cycles of DMMA.
            const int bi = warp >> 1, bj0 = 2 * (warp & 1);
            double nI[4], nJ[4], nold[2][2];
            auto fetch = [&](int x) {
                const int I = lst2[x] & 0xffff, J = lst2[x] >> 16;
                const double* GI = lb_tile(a, I, I - p);
                const double* GJ = lb_tile(a, J, J - p);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    nI[q] = __ldcg(GI + q * THREADS + tid);
                    nJ[q] = (I != J) ? __ldcg(GJ + q * THREADS + tid) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = NB * I + 8 * bi + fr, j = NB * J + 8 * (bj0 + q) + 2 * fc + e;
                        nold[q][e] = in_band(a, i, j) ? __ldcg(ab_at(a, i, j)) : 0.0;
                    }
            };
            fetch(0);
            for (int x = 0; x < m2; ++x) {
                const int I = lst2[x] & 0xffff, J = lst2[x] >> 16;
                const bool diag = (I == J);
                __syncthreads();                                // the previous tile's products are done with LIs / LJs
                double oldv[2][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = q * THREADS + tid;
                    LIs[(e >> 5) * S36 + (e & 31)] = nI[q];
                    if (!diag) LJs[(e >> 5) * S36 + (e & 31)] = nJ[q];
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int e = 0; e < 2; ++e) oldv[q][e] = nold[q][e];
                if (x + 1 < m2) fetch(x + 1);
                __syncthreads();
                const double* Lb = diag ? LIs : LJs;
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const bool skip1 = diag && (bj0 + 1 > bi), skip0 = diag && (bj0 > bi);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double af = LIs[(8 * bi + fr) * S36 + 4 * ks + fc];
                    if (!skip0) dmma884(acc[0][0], acc[0][1], af, Lb[(8 * bj0 + fr) * S36 + 4 * ks + fc]);
                    if (!skip1) dmma884(acc[1][0], acc[1][1], af, Lb[(8 * (bj0 + 1) + fr) * S36 + 4 * ks + fc]);
                }
                // what P stages for panel p+2 also goes, by value, into its mailbox
                double* hm = (I == p + 2 && (J == p + 1 || J == p + 2)) ? a.HM + (size_t)(2 * p + (J - p - 1)) * T32 : nullptr;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q == 0 ? skip0 : skip1) continue;
                    const int i = NB * I + 8 * bi + fr, j = NB * J + 8 * (bj0 + q) + 2 * fc;
                    const double w0 = oldv[q][0] - acc[q][0], w1 = oldv[q][1] - acc[q][1];
                    if (in_band(a, i, j)) { __stcg(ab_at(a, i, j), w0); if (hm) __stcg(hm + (8 * bi + fr) * NB + 8 * (bj0 + q) + 2 * fc, w0); }
                    if (in_band(a, i, j + 1)) { __stcg(ab_at(a, i, j + 1), w1); if (hm) __stcg(hm + (8 * bi + fr) * NB + 8 * (bj0 + q) + 2 * fc + 1, w1); }
                }
            }
        }
    }
    (void)upd_done;
}

__device__ __forceinline__ void run_roles(const Args3& a, double* smem) {
    const int r = (int)blockIdx.x - a.rank0;
    if (r == 0) role_P(a, smem);
    else if (r == 1) role_R(a, smem);
    else if (a.two_phase) role_U2(a, smem);
    else role_U(a, smem);
}

// ---- two-sided solve (sb_band_solve4) ---------------------------------------------------------------------
// The pivot chain is the solve's critical path, so it is cut in two: the top m panels are eliminated top-down and --
// on a reversed copy of the matrix -- the bottom kB panels bottom-up, AT THE SAME TIME, by two instances of the
// pipeline (each a partial factorisation that leaves its Schur complement in the trailing tiles); the middle block
// (>= bw rows, so the two ends never couple directly) receives both complements and is solved last; the two back
// substitutions run outwards from it, again concurrently.
struct TwoSided {
    int n, bw, ldab, m32, Lm, nB;       // rows: total, half bandwidth, band row length, 32*m, middle, bottom instance (= n - 32 m)
    int* flags; int nflags;             // the three instances' counters (one block), zeroed by band_reverse_kernel
    double* arm[3]; int narm[3];        // the three instances' inverse tiles + hot-tile mailboxes, NaN-armed by band_reverse_kernel
};

// Source of the system when it arrives as one of two int64 fixed-point stores (AB | g, sb_lm_frame): the store the
// device-resident selector names is converted on the fly -- this kernel then also writes the f64 work band the top
// instance factors and clears the OTHER store for the next assembly (what band_from_fixed_kernel does as a launch of
// its own: one launch and one gap less per LM iteration).
struct FxSrc {
    const long long* s0; const long long* s1; const int* sel; double inv_scale, inv_gscale; int zero_other;
};

// AB2 = the rows >= 32 m of the matrix with both index directions reversed (lower band storage again), its middle block
// zeroed (the bottom instance accumulates only its Schur complement there); g2 likewise.
__global__ void band_reverse_kernel(double* __restrict__ AB, double* __restrict__ g, double* __restrict__ AB2,
                                    double* __restrict__ g2, TwoSided t, FxSrc fx) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = t.nB * t.ldab;
    const long long n_ab = (long long)t.n * t.ldab;
    const long long* src = nullptr;
    if (fx.s0) {
        const int sel = fx.sel ? *fx.sel : 0;
        src = sel ? fx.s1 : fx.s0;
        long long* oth = const_cast<long long*>(sel ? fx.s0 : fx.s1);
        if (e < n_ab + t.n) {                                         // work copy of the whole system + clear the other store
            const long long v = src[e];
            if (e < n_ab) AB[e] = (double)v * fx.inv_scale;
            else g[e - n_ab] = (double)v * fx.inv_gscale;
            if (fx.zero_other) oth[e] = 0;
        }
    }
    if (e < t.nflags) t.flags[e] = 0;
    {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        const int stride = gridDim.x * blockDim.x;
#pragma unroll
        for (int q = 0; q < 3; ++q)
            for (int x = e; x < t.narm[q]; x += stride) t.arm[q][x] = qnan;
    }
    if (e < t.nB) {
        const int gi = t.n - 1 - e;                                   // original row
        g2[e] = (gi >= t.m32 + t.Lm) ? (src ? (double)src[n_ab + gi] * fx.inv_gscale : g[gi]) : 0.0;
    }
    if (e >= total) return;
    const int ip = e / t.ldab, dc = e - ip * t.ldab;                  // reversed row i', band column (j' - i' + bw)
    const int jp = ip - t.bw + dc;
    double v = 0.0;
    if (jp >= 0 && dc <= t.bw) {
        const int gi = t.n - 1 - ip, gj = t.n - 1 - jp;               // original (row, col) with gi <= gj: stored at row gj
        const bool mid_i = gi < t.m32 + t.Lm, mid_j = gj < t.m32 + t.Lm;
        if (!(mid_i && mid_j)) {
            const size_t at = (size_t)gj * t.ldab + (gi - gj + t.bw);
            v = src ? (double)src[at] * fx.inv_scale : AB[at];
        }
    }
    AB2[e] = v;
}

// the conversion alone, for the systems that take the one-sided path
__global__ void band_from_fixed3_kernel(FxSrc fx, long long n_ab, long long n_tot, double* __restrict__ AB,
                                        double* __restrict__ g) {
    const int sel = fx.sel ? *fx.sel : 0;
    const long long* src = sel ? fx.s1 : fx.s0;
    long long* oth = const_cast<long long*>(sel ? fx.s0 : fx.s1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tot; i += (long long)gridDim.x * blockDim.x) {
        const long long v = src[i];
        if (i < n_ab) AB[i] = (double)v * fx.inv_scale;
        else g[i - n_ab] = (double)v * fx.inv_gscale;
        if (fx.zero_other) oth[i] = 0;
    }
}

// middle system = the top instance's trailing block (original values + its complement) + the bottom instance's complement
__global__ void band_combine_kernel(const double* __restrict__ AB, const double* __restrict__ g, const double* __restrict__ AB2,
                                    const double* __restrict__ g2, double* __restrict__ ABm, double* __restrict__ gm, TwoSided t) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < t.Lm) gm[e] = g[t.m32 + e] + g2[t.n - 1 - (t.m32 + e)];
    if (e >= t.Lm * t.ldab) return;
    const int i = e / t.ldab, dc = e - i * t.ldab;
    const int j = i - t.bw + dc;                                      // middle-local column
    double v = 0.0;
    if (j >= 0 && dc <= t.bw) {
        const int gi = t.m32 + i, gj = t.m32 + j;
        v = AB[(size_t)gi * t.ldab + dc] + AB2[(size_t)(t.n - 1 - gj) * t.ldab + dc];
    }
    ABm[e] = v;
}

// CTA 0: top instance, CTA 1: bottom instance (reversed).  s = y on the eliminated rows, x of the middle block on the others.
__global__ void __launch_bounds__(THREADS, 1) band_backsub2_kernel(Args3 a0, Args3 a1, const double* __restrict__ xm,
                                                                   double* __restrict__ gout, TwoSided t) {
    extern __shared__ double smem[];
    const bool top = blockIdx.x == 0;
    const Args3& a = top ? a0 : a1;
    double* s = smem;
    const int ke32 = NB * a.ke;
    for (int i = threadIdx.x; i < a.NP * NB; i += THREADS) {
        double v = 0.0;
        if (i < ke32) v = __ldcg(a.g + i);
        else if (i < a.n) v = top ? xm[i - ke32] : xm[t.Lm - 1 - (i - ke32)];
        s[i] = v;
    }
    backsub(a, s, a.ke);
    __syncthreads();
    if (top) {
        for (int i = threadIdx.x; i < a.n; i += THREADS) gout[i] = s[i];             // rows [0, 32 m + Lm)
    } else {
        for (int i = threadIdx.x; i < ke32; i += THREADS) gout[t.n - 1 - i] = s[i];   // rows [32 m + Lm, n), un-reversed
    }
}

// ---- bulk-copy (TMA) back substitution of the two-sided solve ---------------------------------------------------
// The register-prefetched push loop above is bound by how fast ONE CTA pulls L out of L2: a panel's 11 tiles are
// 88 KB and only one panel's loads are ever in flight (profiles/r1_microbench.md: 15-107 B/cycle depending only on
// the bytes in flight) -> 3.4 k cycles per panel, 61 us for the two concurrent halves at n = 1862.  Here row k's tiles
// (contiguous in the workspace) and L(k,k)^-1 arrive by cp.async.bulk into a two-stage shared-memory ring, each stage
// with its own mbarrier, issued two panels ahead by one thread: up to 176 KB in flight, nothing of the copy on an
// issue slot, and only the tiles whose targets are eliminated rows are fetched at all.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    long long it = 0;
    while (true) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) break;
        if (++it > (1LL << 26)) __trap();
    }
}

// out[lane] = sum_r T[r][lane] * v[r]   (T row-major 32x32 in shared memory, v in shared memory: broadcast reads)
__device__ __forceinline__ double tile_tmatvec(const double* __restrict__ T, const double* __restrict__ v, int lane) {
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
    for (int r = 0; r < NB; r += 4) {
        c0 = fma(T[r * NB + lane], v[r], c0);
        c1 = fma(T[(r + 1) * NB + lane], v[r + 1], c1);
        c2 = fma(T[(r + 2) * NB + lane], v[r + 2], c2);
        c3 = fma(T[(r + 3) * NB + lane], v[r + 3], c3);
    }
    return (c0 + c1) + (c2 + c3);
}

constexpr int BS3_STAGES = 2;
size_t smem_backsub3(int NP, int WB) {      // s | ring of BS3_STAGES x (WB + 1) tiles | mbarriers
    return (((size_t)NP * NB * sizeof(double) + 127) & ~(size_t)127) + (size_t)BS3_STAGES * (WB + 1) * T32 * sizeof(double) + 64;
}

// CTA 0: top instance, CTA 1: bottom instance (reversed).  Same contract as band_backsub2_kernel.
__global__ void __launch_bounds__(THREADS, 1) band_backsub3_kernel(Args3 a0, Args3 a1, const double* __restrict__ xm,
                                                                   double* __restrict__ gout, TwoSided t) {
    extern __shared__ __align__(128) double smem_bs3[];
    double* smem = smem_bs3;
    const bool top = blockIdx.x == 0;
    const Args3& a = top ? a0 : a1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = a.NP, WB = a.WB, ke = a.ke, ke32 = NB * ke;
    double* s = smem;
    double* ring = smem + ((((size_t)NP * NB * sizeof(double) + 127) & ~(size_t)127) / sizeof(double));
    const size_t stage_elems = (size_t)(WB + 1) * T32;
    unsigned long long* bars = (unsigned long long*)(ring + BS3_STAGES * stage_elems);
    if (tid == 0) {
        for (int q = 0; q < BS3_STAGES; ++q) mbar_init(smem_u32(bars + q), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // source panels, descending: k in [0, kmax]; panel k needs L(k,k)^-1 if it is solved here (k < ke) and the tiles
    // d in [dlo, dhi] whose targets k - d are eliminated rows
    const int kmax = min(NP - 1, ke - 1 + WB);
    auto issue = [&](int k) {                    // thread 0 only
        const int stage = (kmax - k) % BS3_STAGES;
        double* slot = ring + (size_t)stage * stage_elems;
        const unsigned bar = smem_u32(bars + stage);
        const int dlo = max(1, k - ke + 1), dhi = min(WB, k);
        const int nt = max(0, dhi - dlo + 1) + (k < ke ? 1 : 0);
        mbar_expect_tx(bar, (unsigned)(nt * T32 * sizeof(double)));
        if (k < ke) bulk_g2s(smem_u32(slot), a.LI + (size_t)k * T32, T32 * sizeof(double), bar);
        for (int d = dlo; d <= dhi; ++d)
            bulk_g2s(smem_u32(slot + (size_t)d * T32), lb_tile(a, k, d), T32 * sizeof(double), bar);
    };
    __syncthreads();
    if (tid == 0)
        for (int q = 0; q < BS3_STAGES && kmax - q >= 0; ++q) issue(kmax - q);
    for (int i = tid; i < NP * NB; i += THREADS) {
        double v = 0.0;
        if (i < ke32) v = __ldcg(a.g + i);
        else if (i < a.n) v = top ? xm[i - ke32] : xm[t.Lm - 1 - (i - ke32)];
        s[i] = v;
    }
    __syncthreads();
    for (int k = kmax; k >= 0; --k) {
        const int idx = kmax - k, stage = idx % BS3_STAGES;
        const double* slot = ring + (size_t)stage * stage_elems;
        mbar_wait(smem_u32(bars + stage), (unsigned)((idx / BS3_STAGES) & 1));
        double* sk = s + NB * k;
        if (k < ke) {
            if (warp == 0) {                     // x_k = L(k,k)^-T s_k
                const double x = tile_tmatvec(slot, sk, lane);
                __syncwarp();
                sk[lane] = x;
            }
            __syncthreads();
        }
        const int dlo = max(1, k - ke + 1), dhi = min(WB, k);
        // s_{k-d} -= L(k,k-d)^T x_k; the nearest target (next on the chain) alone on warp 0
        for (int i = warp == 0 ? 0 : warp; dlo + i <= dhi; i += (warp == 0 ? 1 << 20 : 7)) {
            const int d = dlo + i;
            s[NB * (k - d) + lane] -= tile_tmatvec(slot + (size_t)d * T32, sk, lane);
        }
        __syncthreads();
        if (tid == 0 && k - BS3_STAGES >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(k - BS3_STAGES);
        }
    }
    if (top) {
        for (int i = tid; i < a.n; i += THREADS) gout[i] = s[i];                 // rows [0, 32 m + Lm)
    } else {
        for (int i = tid; i < ke32; i += THREADS) gout[t.n - 1 - i] = s[i];       // rows [32 m + Lm, n), un-reversed
    }
}

// ---- cluster back substitution of the two-sided solve --------------------------------------------------------------
// One CTA pulls L out of L2 at ~45 B/cycle here whatever the mechanism (register prefetch or bulk copies: 2.1 k cycles
// per 88 KB panel, profiles/r1f_band4_stages.md), so the panel's tiles are spread over a 4-CTA cluster and only what is
// on the chain stays on the leader:
//   leader  warp 0: s_k = y_k - sum_d mail[k][d];  x_k = L(k,k)^-T s_k -> xbuf of all four CTAs (DSMEM stores);
//                   mail[k-1][1] = L(k,k-1)^T x_k  (the next panel's last input)
//           warp 1: mail[k-2][2] = L(k,k-2)^T x_k
//   helper h = 1..3, warp w = 0..2: tile d = 3 + (h-1) + 3 w:  mail[k-d][d] = L(k,k-d)^T x_k, stored into the LEADER's
//                   shared memory; a contribution with distance d has d-1 panels of slack before the leader reads it.
// Every hand-over is by value: xbuf and mail are NaN-armed and a word is its own ready flag (no barrier, no fence on
// the chain).  Each CTA's tiles arrive through its own bulk-copy ring (one producer lane, full/empty mbarriers).
constexpr int BS4_C = 4, BS4_TILES = 3, BS4_MAXWB = 2 + 3 * (BS4_C - 1);
// optional LM step folded into the solve's last kernel (lm_step_kernel's contract, lm.cu): beta[7 pos_node[p] + c] += x[7 p + c]
// unless the factorisation failed (*info != 0) or an earlier iteration did (*failed != 0); a failure sets *failed.
struct StepArgs { int* failed; double* beta; const int* pos_node; };
struct Bs4Layout { size_t xbuf, ybuf, pg, mail, xs, ring, bars, total; int stages; };
Bs4Layout bs4_layout(int NP, int WB, int ke) {
    Bs4Layout L;
    size_t o = 0;
    L.xbuf = o; o += (size_t)NP * NB * 8;
    L.ybuf = o; o += (size_t)ke * NB * 8;
    L.pg = o; o += (size_t)ke * NB * 8;
    L.mail = o; o += (size_t)ke * WB * NB * 8;
    L.xs = o; o += 8 * NB * 8;
    o = (o + 127) & ~(size_t)127;
    L.ring = o;
    const size_t stage = (size_t)BS4_TILES * T32 * 8, avail = 227 * 1024 - 256;
    int stages = o + 2 * stage <= avail ? (int)((avail - o) / stage) : 0;
    if (stages > 8) stages = 8;
    L.stages = stages;
    o += (size_t)stages * stage;
    L.bars = o; o += 2 * 8 * 8;
    L.total = o;
    return L;
}

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// spin until the word is not NaN; after one timeout the caller stops waiting (a NaN that is a VALUE -- a failed
// factorisation upstream -- must flow through, not hang or trap)
__device__ __forceinline__ double poll_value(const double* p, bool& poisoned) {
    double v = *(const volatile double*)p;
    if (v == v || poisoned) return v;
    for (long long it = 0; it < (1LL << 19); ++it) {
        const long long t0 = clock64();                  // back off between polls: a tight LDS loop on one warp slows the
        while (clock64() - t0 < 48) {}                   // shared-memory products of the others (measured)
        v = *(const volatile double*)p;
        if (v == v) return v;
    }
    poisoned = true;
    return v;
}

__global__ void __launch_bounds__(THREADS, 1) band_backsub4_kernel(Args3 a0, Args3 a1, const double* __restrict__ xm,
                                                                   double* __restrict__ gout, TwoSided t, Bs4Layout L,
                                                                   StepArgs step) {
    extern __shared__ __align__(128) unsigned char smem_bs4[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const bool top = (int)blockIdx.x < BS4_C;
    const Args3& a = top ? a0 : a1;
    if (*(const volatile int*)a.info != 0) {                // failed factorisation: the caller ignores the step (LM.py:99-103)
        if (step.failed && blockIdx.x == 0 && threadIdx.x == 0) *step.failed = 1;
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_poison;                                 // a poll of the chain timed out (time slicing, a debugger): report it
    if (tid == 0) s_poison = 0;
    const bool prof = (DBGF(a) & 1024) && top && rank == 0 && tid == 0;
    if (prof) { a.prof[56] = clock64(); for (int d = 0; d < BS4_MAXWB; ++d) a.prof[70 + d] = 0; }
    const int NP = a.NP, WB = a.WB, ke = a.ke, ke32 = NB * ke;
    double* xbuf = (double*)(smem_bs4 + L.xbuf);
    double* ybuf = (double*)(smem_bs4 + L.ybuf);
    double* pg = (double*)(smem_bs4 + L.pg);                 // y_j - sum_{d>=2} mail[j][d], NaN-armed
    double* mail = (double*)(smem_bs4 + L.mail);             // [target j < ke][d - 1][lane]
    double* xs = (double*)(smem_bs4 + L.xs) + warp * NB;
    double* ring = (double*)(smem_bs4 + L.ring);
    unsigned long long* full = (unsigned long long*)(smem_bs4 + L.bars);
    unsigned long long* empty = full + 8;
    const int S = L.stages;
    const int n_cons = 3;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int kmax = min(NP - 1, ke - 1 + WB);
    // this CTA's tile in ring slot q of source panel k: leader 0 = L(k,k)^-1, 1 = d 1, 2 = d 2; helper q -> d = 2 + rank + 3 q
    auto slot_d = [&](int q) { return rank == 0 ? q : 2 + rank + 3 * q; };
    auto wanted = [&](int k, int q) {
        const int d = slot_d(q);
        if (d == 0) return k < ke;
        return d >= max(1, k - ke + 1) && d <= min(WB, k);
    };
    // the producer lane owns the barriers (starting its copies under the buffers' initialisation was tried: the lane
    // then reaches the cluster barrier 4 k cycles late, which costs more than the first tiles' latency it hides)
    auto produce = [&](int k) {
        const int idx = kmax - k, stage = idx % S;
        mbar_wait(smem_u32(empty + stage), (unsigned)(((idx / S) & 1) ^ 1));
        double* slot = ring + (size_t)stage * BS4_TILES * T32;
        const unsigned bar = smem_u32(full + stage);
        int nt = 0;
        for (int q = 0; q < BS4_TILES; ++q) nt += wanted(k, q) ? 1 : 0;
        if (nt == 0) { mbar_arrive(bar); return; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (unsigned)(nt * T32 * sizeof(double)));
        for (int q = 0; q < BS4_TILES; ++q)
            if (wanted(k, q)) {
                const int d = slot_d(q);
                const double* src = d == 0 ? a.LI + (size_t)k * T32 : lb_tile(a, k, d);
                bulk_g2s(smem_u32(slot + (size_t)q * T32), src, T32 * sizeof(double), bar);
            }
    };
    const bool producer = warp == 7 && lane == 0;
    if (producer) {
        for (int q = 0; q < S; ++q) { mbar_init(smem_u32(full + q), 1); mbar_init(smem_u32(empty + q), n_cons); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NP * NB; i += THREADS) {
        double v = qnan;
        if (i >= a.n) v = 0.0;
        else if (i >= ke32) v = top ? xm[i - ke32] : xm[t.Lm - 1 - (i - ke32)];
        xbuf[i] = v;
    }
    if (rank == 0) {
        for (int i = tid; i < ke32; i += THREADS) { ybuf[i] = __ldcg(a.g + i); pg[i] = qnan; }
        for (int i = tid; i < ke * WB * NB; i += THREADS) mail[i] = qnan;
    }
    cluster.sync();                                          // everybody armed before the first remote store
    if (prof) a.prof[57] = clock64();
    if (warp == 7) {
        if (producer)
            for (int k = kmax; k >= 0; --k) produce(k);
    } else if (rank == 0 && warp < 2) {
        // ---- the chain, on two warps that alternate panels: while one is on the chain the other loads its next
        // operands into registers.  Target k: x_k = L(k,k)^-T (pg_k - L(k+1,k)^T x_{k+1}), pg_k from warp 3.
        bool poisoned = false;
        double tq[NB], ti[NB];
        long long tacc[3] = {0, 0, 0}, c0 = 0, c1;
#define BS4PROF(slot_) do { if (prof) { c1 = clock64(); tacc[slot_] += c1 - c0; c0 = c1; } } while (0)
        for (int k = kmax; k >= 0; --k) {
            const int idx = kmax - k, stage = idx % S;
            mbar_wait(smem_u32(full + stage), (unsigned)((idx / S) & 1));
            const double* slot = ring + (size_t)stage * BS4_TILES * T32;
            const bool mine = k < ke && (k & 1) == warp;                       // I solve panel k
            const bool feeds_mine = k >= 1 && k - 1 < ke && ((k - 1) & 1) == warp && wanted(k, 1);   // tile (k, d=1) feeds my next target
            if (mine) {
#pragma unroll
                for (int r = 0; r < NB; ++r) ti[r] = slot[r * NB + lane];
            }
            if (mine) {
                if (prof) c0 = clock64();
                double c = 0.0;
                if (k + 1 <= kmax && wanted(k + 1, 1)) {                      // tq holds L(k+1,k), loaded one iteration ago
                    const double* xk1 = xbuf + NB * (k + 1);
                    (void)poll_value(xk1 + lane, poisoned);
                    __syncwarp();
                    BS4PROF(0);
                    double c0_ = 0.0, c1_ = 0.0, c2_ = 0.0, c3_ = 0.0;
#pragma unroll
                    for (int r = 0; r < NB; r += 4) {
                        c0_ = fma(tq[r], xk1[r], c0_);
                        c1_ = fma(tq[r + 1], xk1[r + 1], c1_);
                        c2_ = fma(tq[r + 2], xk1[r + 2], c2_);
                        c3_ = fma(tq[r + 3], xk1[r + 3], c3_);
                    }
                    c = (c0_ + c1_) + (c2_ + c3_);
                }
                const double sv = poll_value(pg + NB * k + lane, poisoned) - c;
                xs[lane] = sv;
                __syncwarp();
                if (prof && sv == 1.2345e300) c0 = 0;
                BS4PROF(1);
                double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
                for (int r = 0; r < NB; r += 4) {
                    x0 = fma(ti[r], xs[r], x0);
                    x1 = fma(ti[r + 1], xs[r + 1], x1);
                    x2 = fma(ti[r + 2], xs[r + 2], x2);
                    x3 = fma(ti[r + 3], xs[r + 3], x3);
                }
                const double x = (x0 + x1) + (x2 + x3);
                double* xk = xbuf + NB * k;
                *(volatile double*)(xk + lane) = x;
#pragma unroll
                for (int h = 1; h < BS4_C; ++h) *(volatile double*)(cluster.map_shared_rank(xk + lane, h)) = x;
                __syncwarp();
                if (prof && x == 1.2345e300) c0 = 0;
                BS4PROF(2);
            }
            if (feeds_mine) {                       // after my own step: tq of the step above is dead now
#pragma unroll
                for (int r = 0; r < NB; ++r) tq[r] = slot[T32 + r * NB + lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(empty + stage));
            if (prof && k == ke) a.prof[60] = clock64();
        }
        if (poisoned) s_poison = 1;
        if (prof) { a.prof[58] = clock64(); for (int q5 = 0; q5 < 3; ++q5) a.prof[64 + q5] = tacc[q5]; }
#undef BS4PROF
    } else if (rank == 0 && warp == 3) {
        // ---- everything of s_k that has slack: pg_k = y_k - sum_{d>=2} mail[k][d]
        bool poisoned = false;
        for (int k = ke - 1; k >= 0; --k) {
            const int nd = min(WB, NP - 1 - k);
            const double* mk = mail + (size_t)k * WB * NB + lane;
            double mv[BS4_MAXWB];
            long long spins = 0;
            while (true) {
                bool ok = true;
#pragma unroll
                for (int d = 1; d < BS4_MAXWB; ++d) {
                    mv[d] = d < nd ? *(const volatile double*)(mk + d * NB) : 0.0;
                    ok = ok && (mv[d] == mv[d]);
                }
                if (ok || poisoned) break;
                if (++spins > (1LL << 18)) poisoned = true;
            }
            double sv = ybuf[NB * k + lane];
#pragma unroll
            for (int d = 1; d < BS4_MAXWB; ++d) sv -= mv[d];
            *(volatile double*)(pg + NB * k + lane) = sv;
        }
    } else if (warp < n_cons) {
        // ---- one tile per warp and panel: leader warp 2 -> d = 2 (local mail), helper warps 0..2 -> their d (leader's mail)
        double* lead_mail = rank == 0 ? mail : cluster.map_shared_rank(mail, 0);
        bool poisoned = false;
        const int q = rank == 0 ? 2 : warp;
        for (int k = kmax; k >= 0; --k) {
            const int idx = kmax - k, stage = idx % S;
            mbar_wait(smem_u32(full + stage), (unsigned)((idx / S) & 1));
            const double* slot = ring + (size_t)stage * BS4_TILES * T32;
            const double* xk = xbuf + NB * k;
            if (wanted(k, q)) {
                double tq[NB];
#pragma unroll
                for (int r = 0; r < NB; ++r) tq[r] = slot[(size_t)q * T32 + r * NB + lane];
                (void)poll_value(xk + lane, poisoned);
                __syncwarp();
                const int d = slot_d(q);
                double c0_ = 0.0, c1_ = 0.0, c2_ = 0.0, c3_ = 0.0;
#pragma unroll
                for (int r = 0; r < NB; r += 4) {
                    c0_ = fma(tq[r], xk[r], c0_);
                    c1_ = fma(tq[r + 1], xk[r + 1], c1_);
                    c2_ = fma(tq[r + 2], xk[r + 2], c2_);
                    c3_ = fma(tq[r + 3], xk[r + 3], c3_);
                }
                *(volatile double*)(lead_mail + ((size_t)(k - d) * WB + (d - 1)) * NB + lane) = (c0_ + c1_) + (c2_ + c3_);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(empty + stage));
        }
    }
    __syncthreads();
    if (prof) a.prof[59] = clock64();
    if (rank == 0) {
        // a timed-out poll means x is garbage: no step, *info = 2 and the LM failure flag, so that the caller sees it
        const bool timed_out = s_poison != 0;
        if (timed_out && tid == 0) {
            *a.info = 2;
            if (step.failed) *step.failed = 1;
        }
        const bool do_step = step.beta && !timed_out && !(step.failed && *(const volatile int*)step.failed != 0);
        const int rows = top ? a.n : ke32;                   // top: rows [0, 32 m + Lm); bottom: rows [32 m + Lm, n), un-reversed
        for (int i = tid; i < rows; i += THREADS) {
            const int gi = top ? i : t.n - 1 - i;
            const double x = xbuf[i];
            gout[gi] = x;
            if (do_step) {
                const int p = gi / 7, c = gi - 7 * p;
                step.beta[7 * (step.pos_node ? step.pos_node[p] : p) + c] += x;
            }
        }
    }
    cluster.sync();                                          // nobody leaves while its shared memory may still be written
    if (prof) a.prof[61] = clock64();
}

// the step as its own launch, behind the solvers that do not fold it in
__global__ void band_step_kernel(const int* __restrict__ info, const double* __restrict__ x, int n, StepArgs step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = (step.failed && *step.failed != 0) || *info != 0;
    if (i < n && !bad) {
        const int p = i / 7, c = i - 7 * p;
        step.beta[7 * (step.pos_node ? step.pos_node[p] : p) + c] += x[i];
    }
    if (i == 0 && bad && step.failed) *step.failed = 1;
}

__global__ void __launch_bounds__(THREADS, 1) band_chol3_dual_kernel(Args3 a0, Args3 a1) {
    extern __shared__ double smem[];
    run_roles((int)blockIdx.x < a1.rank0 ? a0 : a1, smem);
}

__global__ void __launch_bounds__(THREADS, 1) band_chol3_kernel(Args3 a) {
    extern __shared__ double smem[];
    if (blockIdx.x == 0) role_P(a, smem);
    else if (blockIdx.x == 1) role_R(a, smem);
    else if (a.two_phase) role_U2(a, smem);
    else role_U(a, smem);
}

size_t smem_bytes3(int n) {
    const size_t NP = (size_t)(n + NB - 1) / NB;
    const size_t p_role = 2 * (size_t)T33 + 8 * (size_t)T36 + 2 * NB + 3 * 96;
    const size_t r_role = NP * NB + (size_t)T33 + NB + 8 * (size_t)T33;
    const size_t u_role = 5 * (size_t)T36;
    size_t m = p_role > r_role ? p_role : r_role;
    if (u_role > m) m = u_role;
    return m * sizeof(double);
}

long long ws_bytes3(int n, int bw) {
    const long long NP = (n + NB - 1) / NB, WB = (bw + NB - 1) / NB;
    return (NP * (WB > 0 ? WB : 1) + 3 * NP) * (long long)T32 * (long long)sizeof(double) + ((4 * NP * (long long)sizeof(int) + 255) & ~255LL) + 256 + 1024;
}

#ifdef SB_DEBUG_EXPORTS
int g_debug3 = 0;            // timing experiments only; the product build has no mutable globals
cudaEvent_t g_ev4[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // stage stamps of sb_band_solve4 (debug & 256)
inline void stamp4(int i, cudaStream_t st) {
    if (!(g_debug3 & 256)) return;
    if (!g_ev4[i]) cudaEventCreate(&g_ev4[i]);
    cudaEventRecord(g_ev4[i], st);
}
#else
constexpr int g_debug3 = 0;
inline void stamp4(int, cudaStream_t) {}
#endif

// Per-device launch configuration (SM count; largest dynamic shared memory opted in per kernel): cudaFuncSetAttribute and
// the co-residency bound are properties of the DEVICE the call runs on, so they are cached per device ordinal.
struct DevConf { int sms; size_t smem_single, smem_dual, smem_back2, smem_back3, smem_back4; };
DevConf* dev_conf() {
    static DevConf conf[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (conf[dev].sms == 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        conf[dev].sms = sms;
    }
    return &conf[dev];
}
template <typename K>
bool opt_in_smem(K kernel, size_t smem, size_t& configured) {
    if (smem <= configured) return true;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return false;
    configured = smem;
    return true;
}

}  // namespace

extern "C" {

#ifdef SB_DEBUG_EXPORTS
int sb_band3_debug(int flags) { g_debug3 = flags; return SB_OK; }

/* byte offset of the 32 cycle counters inside the workspace (debug flag 4) */
long long sb_band3_prof_offset(int n, int bw) { return ws_bytes3(n, bw) - 1024; }
#endif

int sb_band_max_bw(void) { return MAX_WB3 * NB; }

int sb_band_solve4_step_fx(const long long* fx_store0, const long long* fx_store1, const int* sel, int fx_shift,
                           int fx_gshift, int zero_other, double* AB, int ldab, int n, int bw, double* g, const double* u,
                           double* dinv, int* info, void* workspace, long long ws_bytes, int n_ctas, int* lm_failed,
                           double* beta, const int* pos_node, void* stream);

int sb_band3_update_role(int n, int bw, int n_ctas) {
    const int NP = (n + NB - 1) / NB;
    int WB = (bw + NB - 1) / NB;
    if (WB > NP - 1) WB = NP - 1 > 0 ? NP - 1 : 0;
    return (WB * (WB + 1) / 2 - 1 > n_ctas - 2 || (g_debug3 & 64)) ? 2 : 1;
}

int sb_band3_fits(int n, int bw) {
    return (n > 0 && bw >= 0 && (bw + NB - 1) / NB <= MAX_WB3 && smem_bytes3(n) <= 227 * 1024) ? 1 : 0;
}

long long sb_band3_workspace_bytes(int n, int bw) { return ws_bytes3(n, bw); }

int sb_band_solve3(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                   void* workspace, long long ws_bytes, int n_ctas, void* stream) {
    if (!AB || !g || !dinv || !info || !workspace || n <= 0 || bw < 0 || ldab < bw + 1) return SB_ERR_ARG;
    if (n_ctas < 3) return SB_ERR_ARG;
    if (!sb_band3_fits(n, bw)) return SB_ERR_ARG;
    if (ws_bytes < ws_bytes3(n, bw)) return SB_ERR_WORKSPACE;
    DevConf* dc = dev_conf();
    if (!dc) return SB_ERR_CUDA;
    const size_t smem = smem_bytes3(n);
    if (!opt_in_smem(band_chol3_kernel, smem, dc->smem_single)) return SB_ERR_CUDA;
    if (n_ctas > dc->sms) n_ctas = dc->sms;   // 1 CTA per SM (launch bounds): the cooperative launch checks co-residency
    Args3 a;
    a.AB = AB; a.ldab = ldab; a.n = n; a.bw = bw; a.g = g; a.u = u; a.dinv = dinv; a.info = info;
    a.NP = (n + NB - 1) / NB;
    a.WB = (bw + NB - 1) / NB;
    if (a.WB > a.NP - 1) a.WB = a.NP - 1 > 0 ? a.NP - 1 : 0;
    const long long tiles = (long long)a.NP * (a.WB > 0 ? a.WB : 1);
    a.LB = (double*)workspace;
    a.LI = a.LB + tiles * T32;
    a.HM = a.LI + (long long)a.NP * T32;
    a.flags = (int*)(a.HM + 2LL * a.NP * T32);
    a.two_phase = (a.WB * (a.WB + 1) / 2 - 1 > n_ctas - 2 || (g_debug3 & 64)) ? 1 : 0;
    a.ke = a.NP; a.rank0 = 0; a.ncta = n_ctas;
    a.prof = (long long*)((char*)workspace + ws_bytes3(n, bw) - 1024);
    a.debug = g_debug3;
    if (cudaMemsetAsync(a.flags, 0, 4 * (size_t)a.NP * sizeof(int), (cudaStream_t)stream) != cudaSuccess)
        return SB_ERR_CUDA;
    // the inverse tiles and the hot-tile mailbox are taken by value: all-ones bytes are a NaN
    if (cudaMemsetAsync(a.LI, 0xff, 3 * (size_t)a.NP * T32 * sizeof(double), (cudaStream_t)stream) != cudaSuccess)
        return SB_ERR_CUDA;
    void* kargs[] = {(void*)&a};
    if (cudaLaunchCooperativeKernel((const void*)band_chol3_kernel, dim3(n_ctas), dim3(THREADS), kargs, smem,
                                    (cudaStream_t)stream) != cudaSuccess)
        return SB_ERR_CUDA;
    SB_CHECK_LAUNCH();
    return SB_OK;
}


/* ---- two-sided solve ------------------------------------------------------------------------------------ */
static void fill_args3(Args3& a, double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                       void* ws, int ke, int rank0, int ncta) {
    a.AB = AB; a.ldab = ldab; a.n = n; a.bw = bw; a.g = g; a.u = u; a.dinv = dinv; a.info = info;
    a.NP = (n + NB - 1) / NB;
    a.WB = (bw + NB - 1) / NB;
    if (a.WB > a.NP - 1) a.WB = a.NP - 1 > 0 ? a.NP - 1 : 0;
    const long long tiles = (long long)a.NP * (a.WB > 0 ? a.WB : 1);
    a.LB = (double*)ws;
    a.LI = a.LB + tiles * T32;
    a.HM = a.LI + (long long)a.NP * T32;
    a.flags = (int*)(a.HM + 2LL * a.NP * T32);
    a.two_phase = (a.WB * (a.WB + 1) / 2 - 1 > ncta - 2 || (g_debug3 & 64)) ? 1 : 0;
    a.ke = ke < 0 ? a.NP : ke; a.rank0 = rank0; a.ncta = ncta;
    a.prof = (long long*)((char*)ws + ws_bytes3(n, bw) - 1024);
    a.debug = ((g_debug3 & 2048) && rank0 == 0 && ke >= 0) ? g_debug3 : (g_debug3 & ~4);   // 2048: role counters of the top instance
}

/* split: m panels eliminated from the top, m from the bottom, the middle Lm = n - 64 m >= bw rows; 0 = do not split */
static int two_sided_m(int n, int bw) {
    const int need = (bw > 64 ? bw : 64);
    int m = (n - need) / 64;
    while (m > 0 && n - 64 * m < need) --m;
    return m >= 4 ? m : 0;
}

static long long align256(long long x) { return (x + 255) & ~255LL; }

long long sb_band4_workspace_bytes(int n, int bw, int ldab) {
    const int m = two_sided_m(n, bw);
    if (m == 0) return ws_bytes3(n, bw);
    const int Lm = n - 64 * m, nA = 32 * m + Lm;
    return 2 * align256(ws_bytes3(nA, bw)) + align256(ws_bytes3(Lm, bw < Lm - 1 ? bw : Lm - 1)) +
           align256((long long)nA * ldab * 8) + align256((long long)nA * 8) + align256((long long)Lm * ldab * 8) +
           align256((long long)Lm * 8) + 2 * align256((long long)n * 8) + align256(12LL * ((n + NB - 1) / NB + 2) * 4) + 1024;
}

static int launch_step(const int* info, const double* x, int n, const StepArgs& step, cudaStream_t st) {
    if (!step.beta) return SB_OK;
    band_step_kernel<<<(n + 255) / 256, 256, 0, st>>>(info, x, n, step);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

static int solve4_impl(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                       void* workspace, long long ws_bytes, int n_ctas, void* stream, StepArgs step, FxSrc fx = FxSrc{}) {
    if (!AB || !g || !dinv || !info || !workspace || n <= 0 || bw < 0 || ldab < bw + 1) return SB_ERR_ARG;
    const int m = two_sided_m(n, bw);
    if (m == 0 || n_ctas < 8) {
        if (fx.s0) {
            const long long n_ab = (long long)n * ldab, n_tot = n_ab + n;
            long long blocks = (n_tot + 255) / 256;
            if (blocks > 1184) blocks = 1184;
            band_from_fixed3_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(fx, n_ab, n_tot, AB, g);
            SB_CHECK_LAUNCH();
        }
        const int rc = sb_band_solve3(AB, ldab, n, bw, g, u, dinv, info, workspace, ws_bytes, n_ctas, stream);
        return rc != SB_OK ? rc : launch_step(info, g, n, step, (cudaStream_t)stream);
    }
    if (ws_bytes < sb_band4_workspace_bytes(n, bw, ldab)) return SB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    DevConf* dc = dev_conf();
    if (!dc) return SB_ERR_CUDA;
    if (n_ctas > dc->sms) n_ctas = dc->sms;
    TwoSided t;
    t.n = n; t.bw = bw; t.ldab = ldab; t.m32 = 32 * m; t.Lm = n - 64 * m; t.nB = n - 32 * m;
    const int nA = t.m32 + t.Lm, bwm = bw < t.Lm - 1 ? bw : t.Lm - 1;
    if (!sb_band3_fits(nA, bw) || !sb_band3_fits(t.Lm, bwm)) return SB_ERR_ARG;
    char* w = (char*)workspace;
    void* wsA = w; w += align256(ws_bytes3(nA, bw));
    void* wsB = w; w += align256(ws_bytes3(nA, bw));
    void* wsM = w; w += align256(ws_bytes3(t.Lm, bwm));
    double* AB2 = (double*)w; w += align256((long long)nA * ldab * 8);
    double* g2 = (double*)w; w += align256((long long)nA * 8);
    double* ABm = (double*)w; w += align256((long long)t.Lm * ldab * 8);
    double* gm = (double*)w; w += align256((long long)t.Lm * 8);
    double* dinvB = (double*)w; w += align256((long long)n * 8);
    double* dinvM = (double*)w; w += align256((long long)n * 8);
    int* flags4 = (int*)w;
    const int cA = n_ctas / 2, cB = n_ctas - cA;
    Args3 aA, aB, aM;
    fill_args3(aA, AB, ldab, nA, bw, g, u, dinv, info, wsA, m, 0, cA);
    fill_args3(aB, AB2, ldab, t.nB, bw, g2, u, dinvB, info, wsB, m, cA, cB);
    const int cM = n_ctas < 96 ? n_ctas : 96;
    fill_args3(aM, ABm, ldab, t.Lm, bwm, gm, u, dinvM, info, wsM, -1, 0, cM);
    aA.flags = flags4; aB.flags = aA.flags + 4 * aA.NP; aM.flags = aB.flags + 4 * aB.NP;
    t.flags = flags4; t.nflags = 4 * (aA.NP + aB.NP + aM.NP);
    t.arm[0] = aA.LI; t.narm[0] = 3 * aA.NP * T32;
    t.arm[1] = aB.LI; t.narm[1] = 3 * aB.NP * T32;
    t.arm[2] = aM.LI; t.narm[2] = 3 * aM.NP * T32;
    size_t smem = smem_bytes3(nA);
    if (smem_bytes3(t.Lm) > smem) smem = smem_bytes3(t.Lm);
    if (!opt_in_smem(band_chol3_dual_kernel, smem, dc->smem_dual)) return SB_ERR_CUDA;
    if (!opt_in_smem(band_chol3_kernel, smem, dc->smem_single)) return SB_ERR_CUDA;
    const size_t smem_back = (size_t)aA.NP * NB * sizeof(double);
    if (!opt_in_smem(band_backsub2_kernel, smem_back, dc->smem_back2)) return SB_ERR_CUDA;
    // 1. reversed copy of the bottom part
    stamp4(0, st);
    {
        const long long cover = fx.s0 ? (long long)n * ldab + n : (long long)t.nB * ldab;     // >= nB * ldab either way
        band_reverse_kernel<<<(int)((cover + 255) / 256), 256, 0, st>>>(AB, g, AB2, g2, t, fx);
    }
    SB_CHECK_LAUNCH();
    // 2. both ends at once
    stamp4(1, st);
    {
        void* kargs[] = {(void*)&aA, (void*)&aB};
        if (cudaLaunchCooperativeKernel((const void*)band_chol3_dual_kernel, dim3(n_ctas), dim3(THREADS), kargs, smem, st) != cudaSuccess)
            return SB_ERR_CUDA;
    }
    // 3. middle system
    stamp4(2, st);
    band_combine_kernel<<<(t.Lm * ldab + 255) / 256, 256, 0, st>>>(AB, g, AB2, g2, ABm, gm, t);
    SB_CHECK_LAUNCH();
    stamp4(3, st);
    {
        void* kargs[] = {(void*)&aM};
        if (cudaLaunchCooperativeKernel((const void*)band_chol3_kernel, dim3(cM), dim3(THREADS), kargs, smem, st) != cudaSuccess)
            return SB_ERR_CUDA;
    }
    // 4. both back substitutions outwards from the middle
    stamp4(4, st);
    const size_t smem_b3 = smem_backsub3(aA.NP, aA.WB);
    const Bs4Layout L4 = bs4_layout(aA.NP, aA.WB, aA.ke);
    if (aA.WB <= BS4_MAXWB && aA.WB >= 1 && L4.stages >= 2 && !(g_debug3 & (128 | 512))) {
        if (!opt_in_smem(band_backsub4_kernel, (size_t)L4.total, dc->smem_back4)) return SB_ERR_CUDA;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * BS4_C);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = L4.total;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = BS4_C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, band_backsub4_kernel, aA, aB, (const double*)gm, g, t, L4, step) != cudaSuccess) return SB_ERR_CUDA;
        step.beta = nullptr;     // folded in
    } else if (smem_b3 <= 227 * 1024 && !(g_debug3 & 128)) {
        if (!opt_in_smem(band_backsub3_kernel, smem_b3, dc->smem_back3)) return SB_ERR_CUDA;
        band_backsub3_kernel<<<2, THREADS, smem_b3, st>>>(aA, aB, gm, g, t);
    } else {
        band_backsub2_kernel<<<2, THREADS, smem_back, st>>>(aA, aB, gm, g, t);
    }
    SB_CHECK_LAUNCH();
    if (launch_step(info, g, n, step, st) != SB_OK) return SB_ERR_CUDA;
    stamp4(5, st);
    return SB_OK;
}

int sb_band_solve4(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                   void* workspace, long long ws_bytes, int n_ctas, void* stream) {
    return solve4_impl(AB, ldab, n, bw, g, u, dinv, info, workspace, ws_bytes, n_ctas, stream, StepArgs{nullptr, nullptr, nullptr});
}

int sb_band_solve4_step(double* AB, int ldab, int n, int bw, double* g, const double* u, double* dinv, int* info,
                        void* workspace, long long ws_bytes, int n_ctas, int* lm_failed, double* beta,
                        const int* pos_node, void* stream) {
    if (!beta || !lm_failed) return SB_ERR_ARG;
    return solve4_impl(AB, ldab, n, bw, g, u, dinv, info, workspace, ws_bytes, n_ctas, stream, StepArgs{lm_failed, beta, pos_node});
}

int sb_band_solve4_step_fx(const long long* fx_store0, const long long* fx_store1, const int* sel, int fx_shift,
                           int fx_gshift, int zero_other, double* AB, int ldab, int n, int bw, double* g, const double* u,
                           double* dinv, int* info, void* workspace, long long ws_bytes, int n_ctas, int* lm_failed,
                           double* beta, const int* pos_node, void* stream) {
    if (!fx_store0 || !fx_store1 || !beta || !lm_failed || fx_shift < 0 || fx_shift > 60 || fx_gshift < 0 || fx_gshift > 60)
        return SB_ERR_ARG;
    FxSrc fx{fx_store0, fx_store1, sel, ldexp(1.0, -fx_shift), ldexp(1.0, -fx_gshift), zero_other};
    return solve4_impl(AB, ldab, n, bw, g, u, dinv, info, workspace, ws_bytes, n_ctas, stream,
                       StepArgs{lm_failed, beta, pos_node}, fx);
}

#ifdef SB_DEBUG_EXPORTS
/* timing experiments (sb_band3_debug flag 256): milliseconds of the five stages of the last sb_band_solve4 --
   reverse | both ends | combine + flag memset | middle | back substitution.  Synchronises. */
int sb_band4_stage_ms(float* out5) {
    if (!out5 || !g_ev4[5]) return SB_ERR_ARG;
    if (cudaEventSynchronize(g_ev4[5]) != cudaSuccess) return SB_ERR_CUDA;
    for (int i = 0; i < 5; ++i)
        if (cudaEventElapsedTime(out5 + i, g_ev4[i], g_ev4[i + 1]) != cudaSuccess) return SB_ERR_CUDA;
    return SB_OK;
}
#endif

}  // extern "C"
