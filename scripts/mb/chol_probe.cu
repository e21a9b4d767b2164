// micro-benchmark: (1) FP64 issue rate of independent DFMAs (one warp, two warps of one scheduler, four warps on four
// schedulers), (2) the pivot CTA's 32x32 Cholesky (warp_potrf_blocked of band_chol3.cu) in isolation, with the other warps
// idle or polling shared memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I python-super_b200/super_b200/csrc
// -I include scripts/mb/chol_probe.cu -o scripts/mb/chol_probe
#include <cstdio>
#include "../../python-super_b200/super_b200/csrc/band_chol3.cu"

__global__ void k_dfma_tp(double* out, long long* cyc, double seed, unsigned warp_mask, int nchain) {
    const int warp = threadIdx.x >> 5;
    if (!((warp_mask >> warp) & 1u)) return;
    double x[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) x[q] = seed + q * 1e-3 + threadIdx.x * 1e-9;
    const double y = 1.0000001;
    long long t0 = clock64();
    if (nchain == 16) {
#pragma unroll 1
        for (int r = 0; r < 64; ++r) {
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = fma(x[q], y, 1e-9);
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < 128; ++r) {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fma(x[q], y, 1e-9);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += x[q];
    out[threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

template <int SEED, int LEN = 1024>
__device__ __noinline__ unsigned big_code(unsigned x) {
#pragma unroll
    for (int q = 0; q < LEN; ++q) { x ^= x >> 7; x *= (2654435761u + 2u * (unsigned)(q * 31 + SEED)); }
    return x;
}
// mode 0: the other warps only deliver their barrier arrivals; 1: three of them poll a shared word in a tight loop while
// warp 0 factors; 2: the same with the backed-off poll (48 cycles between loads)
__global__ void __launch_bounds__(256, 1) k_potrf(double* out, long long* cyc, int mode, int reps) {
    extern __shared__ double smem[];
    double* D = smem;                 // T33
    double* Lcol = D + T33;           // T36
    double* Ls = Lcol + T36;          // T36
    double* dinvs = Ls + T36;         // NB
    volatile int* stop = (volatile int*)(dinvs + NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    long long ts[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int e = tid; e < NB * NB; e += 256) {
            const int r = e >> 5, c = e & 31;
            D[r * S33 + c] = (r == c) ? 40.0 + r : 1.0 / (1.0 + abs(r - c));
        }
        for (int e = tid; e < T36; e += 256) { Lcol[e] = qnan; Ls[e] = 0.0; }
        if (tid < NB) dinvs[tid] = qnan;
        if (tid == 0) *stop = 0;
        __syncthreads();
        if (warp == 0) {
            const long long t0 = clock64();
            warp_potrf_blocked(D, Lcol, dinvs, Ls, lane, ts, t0);
            const long long t1 = clock64();
            total += t1 - t0;
            if (lane == 0) *stop = 1;
        } else {
            if (warp == 2 || warp == 3 || warp == 5) { bar_arrive(4, 128); bar_arrive(5, 128); bar_arrive(6, 128); }
            // low nibble 15: 4 bits per warp (warp w at bits 4w): 1 tight shared poll, 2 DMMA chain, 3 integer ALU loop,
            // 4 shared store/load loop, 5 global load loop, 6 DFMA chains, 7 backed-off shared poll
            const int act = ((mode & 15) != 15) ? ((warp == 2 || warp == 3 || warp == 5) ? (mode == 1 ? 1 : mode == 2 ? 7 : 0) : 0)
                                        : ((mode >> (4 * warp)) & 15);
            if (act == 1) {
                while (!*stop) { double d = ((volatile double*)dinvs)[31]; if (d == 1.2345e300) break; }
            } else if (act == 7) {
                while (!*stop) {
                    const long long t0 = clock64();
                    while (clock64() - t0 < 48) {}
                    double d = ((volatile double*)dinvs)[31]; if (d == 1.2345e300) break;
                }
            } else if (act == 2) {
                double c0 = 0, c1 = 0, c2 = 0, c3 = 0; const double av = 1e-3 * lane, bv = 1e-3;
                while (!*stop) { dmma884(c0, c1, av, bv); dmma884(c2, c3, av, bv); dmma884(c0, c1, av, bv); dmma884(c2, c3, av, bv); }
                if (c0 + c2 == 1.2345e300) out[300] = c1 + c3;
            } else if (act == 3) {
                unsigned x = lane;
                while (!*stop) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) x = x * 1664525u + 1013904223u;
                }
                if (x == 0x12345678u) out[301] = x;
            } else if (act == 4) {
                volatile double* sc = (volatile double*)(Ls + T36 / 2);
                double v = lane;
                while (!*stop) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) { sc[lane + 32 * (q & 3)] = v; v += sc[(lane + 1) & 31]; }
                }
                if (v == 1.2345e300) out[302] = v;
            } else if (act == 5) {
                double v = 0;
                while (!*stop) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v += __ldcg(out + 512 + ((lane + 32 * q + (int)v) & 255));
                }
                if (v == 1.2345e300) out[303] = v;
            } else if (act == 8) {
                unsigned x = lane;
                while (!*stop) {
                    switch (warp) {
                        case 1: x = big_code<1>(x); break; case 2: x = big_code<2>(x); break; case 3: x = big_code<3>(x); break;
                        case 4: x = big_code<4>(x); break; case 5: x = big_code<5>(x); break; case 6: x = big_code<6>(x); break;
                        default: x = big_code<7>(x); break;
                    }
                }
                if (x == 0x12345678u) out[305] = x;
            } else if (act >= 10 && act <= 13) {
                unsigned x = lane;
                while (!*stop) {
#define BC(W, L) case W: x = big_code<W, L>(x); break;
                    if (act == 10) { switch (warp) { BC(1, 85) BC(2, 85) BC(3, 85) BC(4, 85) BC(5, 85) BC(6, 85) default: x = big_code<7, 85>(x); } }
                    else if (act == 11) { switch (warp) { BC(1, 170) BC(2, 170) BC(3, 170) BC(4, 170) BC(5, 170) BC(6, 170) default: x = big_code<7, 170>(x); } }
                    else if (act == 12) { switch (warp) { BC(1, 340) BC(2, 340) BC(3, 340) BC(4, 340) BC(5, 340) BC(6, 340) default: x = big_code<7, 340>(x); } }
                    else { switch (warp) { BC(1, 2048) BC(2, 2048) BC(3, 2048) BC(4, 2048) BC(5, 2048) BC(6, 2048) default: x = big_code<7, 2048>(x); } }
                }
                if (x == 0x12345678u) out[307] = x;
            } else if (act == 9) {
                // DMMA fed from shared memory like the helpers' products
                const int fr = lane >> 2, fc = lane & 3;
                double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                const double* Pa = Ls;
                while (!*stop) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        dmma884(c0, c1, ((const volatile double*)Pa)[(fr) * S36 + 4 * ks + fc], ((const volatile double*)Pa)[(8 + fr) * S36 + 4 * ks + fc]);
                        dmma884(c2, c3, ((const volatile double*)Pa)[(16 + fr) * S36 + 4 * ks + fc], ((const volatile double*)Pa)[(24 + fr) * S36 + 4 * ks + fc]);
                    }
                }
                if (c0 + c2 == 1.2345e300) out[306] = c1 + c3;
            } else if (act == 6) {
                double x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = 1.0 + q * 1e-3;
                while (!*stop) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = fma(x[q], 1.0000001, 1e-9);
                }
                if (x[0] + x[1] + x[2] + x[3] + x[4] + x[5] + x[6] + x[7] == 1.2345e300) out[304] = x[0];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        cyc[0] = total / reps;
        for (int q = 4; q < 7; ++q) cyc[q] = ts[q] / reps;
    }
    if (warp == 0) out[lane] = Lcol[31 * S36 + lane] + dinvs[lane];
}

int main() {
    double* dout; long long* dc; cudaMalloc(&dout, 16384); cudaMemset(dout, 0, 16384); cudaMalloc(&dc, 256);
    long long h[16];
    struct { unsigned mask; int nchain; const char* note; } tp[] = {
        {0x1, 16, "one warp, 16 independent chains"}, {0x1, 8, "one warp, 8 independent chains"},
        {0x11, 16, "warps 0 and 4 (same scheduler), 16 chains each"}, {0xf, 16, "warps 0-3 (four schedulers), 16 chains each"},
        {0xff, 16, "eight warps, 16 chains each"}};
    for (auto& c : tp) {
        for (int r = 0; r < 2; ++r) { k_dfma_tp<<<1, 256>>>(dout, dc, 1.25, c.mask, c.nchain); cudaDeviceSynchronize(); }
        cudaMemcpy(h, dc, 64, cudaMemcpyDeviceToHost);
        printf("DFMA issue: %-52s %.2f cycles per warp instruction (warp 0)\n", c.note, h[0] / 1024.0);
    }
    struct { unsigned mode; const char* note; } pm[] = {
        {0, "others idle"}, {1, "warps 2,3,5 poll shared (tight)"}, {2, "warps 2,3,5 poll shared (backed off)"},
        {0x00200220 | 15, "DMMA chains on warps 1,2,5 (other schedulers)"}, {0x00000020 | 15, "DMMA chain on warp 1"},
        {0x00020000 | 15, "DMMA chain on warp 4 (same scheduler)"}, {0x00030000 | 15, "integer loop on warp 4 (same scheduler)"},
        {0x00040000 | 15, "shared store/load loop on warp 4"}, {0x00050000 | 15, "global load loop on warp 4"},
        {0x55050000 | 15, "global load loops on warps 4,6,7"}, {0x00060000 | 15, "DFMA chains on warp 4 (same scheduler)"},
        {0x00600660 | 15, "DFMA chains on warps 1,2,5 (other schedulers)"}, {0x44040000 | 15, "shared loops on warps 4,6,7"},
        {0x00000040 | 15, "shared loop on warp 1"}, {0x44444440 | 15, "shared loops on warps 1-7"},
        {0x88888880 | 15, "big straight-line code (48 KB each) on warps 1-7"}, {0x80800880 | 15, "big code on warps 1,2,5,7"},
        {0x00080000 | 15, "big code on warp 4"}, {0xa0a00aa0u | 15, "4 KB loops on warps 1,2,5,7"}, {0xb0b00bb0u | 15, "8 KB loops on warps 1,2,5,7"},
        {0xc0c00cc0u | 15, "16 KB loops on warps 1,2,5,7"}, {0xd0d00dd0u | 15, "96 KB loops on warps 1,2,5,7"}, {0xaaaaaaa0u | 15, "4 KB loops on warps 1-7"},
        {0xccccccc0u | 15, "16 KB loops on warps 1-7"}, {0x00000080u | 15, "48 KB loop on warp 1"}, {0x000000c0u | 15, "16 KB loop on warp 1"}, {0x99999990 | 15, "shared-fed DMMA on warps 1-7"}, {0x90900990 | 15, "shared-fed DMMA on warps 1,2,5,7"}, {0x55242220 | 15, "mix: DMMA 1,2,3,5; shared 4; global 6,7"}};
    const size_t smem = (T33 + 2 * T36 + NB + 8) * sizeof(double);
    for (auto& c : pm) {
        for (int r = 0; r < 2; ++r) { k_potrf<<<1, 256, smem>>>(dout, dc, (int)c.mode, 8); cudaDeviceSynchronize(); }
        cudaMemcpy(h, dc, 64, cudaMemcpyDeviceToHost);
        printf("potrf 32x32, %-52s: %lld cycles [replica loads %lld, chain %lld, trailing dmma %lld]\n", c.note,
               h[0], h[4], h[5], h[6]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
