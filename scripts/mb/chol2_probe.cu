// micro-benchmark + check of the SPLIT 32x32 Cholesky of the pivot CTA: the chain warp factors only the 8x8 diagonal blocks
// (replica in registers) and, between two blocks, forms L(b+1,b) and the next diagonal block; three helper warps form the rows
// below and every other trailing update one block behind.  Against warp_potrf_blocked (chain warp carries all 32 rows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I python-super_b200/super_b200/csrc -I include scripts/mb/chol2_probe.cu -o scripts/mb/chol2_probe
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../python-super_b200/super_b200/csrc/band_chol3.cu"

__global__ void __launch_bounds__(256, 1) k_potrf2(const double* A, double* outL, double* outT, double* outdinv, long long* cyc, int reps) {
    extern __shared__ double smem[];
    double* D = smem;                 // T33
    double* Lcol = D + T33;           // T36
    double* Ls = Lcol + T36;          // T36
    double* dinvs = Ls + T36;         // NB
    volatile int* lready = (volatile int*)(dinvs + NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    long long ts[12] = {0};
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int e = tid; e < NB * NB; e += 256) D[(e >> 5) * S33 + (e & 31)] = A[e];
        for (int e = tid; e < T36; e += 256) {
            Lcol[e] = qnan;
            const int r = e / S36, c = e % S36;
            Ls[e] = (c < NB && (r >> 3) == (c >> 3)) ? qnan : 0.0;
        }
        if (tid < NB) dinvs[tid] = qnan;
        if (tid < 4) lready[tid] = 0;
        __syncthreads();
        const int h = (warp == 2) ? 0 : (warp == 3) ? 1 : (warp == 5) ? 2 : -1;
        if (warp == 0) {
            const long long t0 = clock64();
            warp_potrf_split(D, Lcol, dinvs, Ls, lready, rep + 1, lane, ts, t0);
            total += clock64() - t0;
        } else if (h >= 0) {
            bar_arrive(4, 128);                       // nothing in front of block 0 here (in the kernel: the rest of D -= L L^T)
            helper_split(D, Lcol, Ls, lready, rep + 1, 0, h, lane);
            bar_arrive(5, 128);
            helper_split(D, Lcol, Ls, lready, rep + 1, 1, h, lane);
            bar_arrive(6, 128);
        }
        __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += 256) { outL[e] = Lcol[(e & 31) * S36 + (e >> 5)]; outT[e] = Ls[(e >> 5) * S36 + (e & 31)]; }
    if (tid < NB) outdinv[tid] = dinvs[tid];
    if (tid == 0) { cyc[0] = total / reps; for (int q = 4; q < 7; ++q) cyc[q] = ts[q] / reps; }
}

int main() {
    const int n = 32;
    std::vector<double> A(n * n), L(n * n, 0.0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) A[r * n + c] = (r == c) ? 40.0 + r : 1.0 / (1.0 + abs(r - c)) + 0.01 * ((r * 7 + c * 3) % 5);
    for (int r = 0; r < n; ++r) for (int c = 0; c < r; ++c) A[c * n + r] = A[r * n + c];
    for (int j = 0; j < n; ++j) {
        double s = A[j * n + j];
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        L[j * n + j] = sqrt(s);
        for (int i = j + 1; i < n; ++i) {
            double t = A[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / L[j * n + j];
        }
    }
    double *dA, *dL, *dT, *dd; long long* dc;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dT, n * n * 8); cudaMalloc(&dd, n * 8); cudaMalloc(&dc, 256);
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    const size_t smem = (T33 + 2 * T36 + NB + 8) * sizeof(double);
    for (int r = 0; r < 2; ++r) { k_potrf2<<<1, 256, smem>>>(dA, dL, dT, dd, dc, 8); cudaDeviceSynchronize(); }
    std::vector<double> hL(n * n), hT(n * n), hd(n); long long h[16];
    cudaMemcpy(hL.data(), dL, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hT.data(), dT, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hd.data(), dd, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(h, dc, 64, cudaMemcpyDeviceToHost);
    double eL = 0, ed = 0, eT = 0;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
            if ((r >> 3) > (c >> 3)) eL = fmax(eL, fabs(hL[r * n + c] - L[r * n + c]));
    for (int c = 0; c < n; ++c) ed = fmax(ed, fabs(hd[c] - 1.0 / L[c * n + c]));
    for (int b = 0; b < 4; ++b)          // T_b L_bb = I
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                double s = 0;
                for (int k = 0; k < 8; ++k) s += hT[(8 * b + i) * n + 8 * b + k] * L[(8 * b + k) * n + 8 * b + j];
                eT = fmax(eT, fabs(s - (i == j ? 1.0 : 0.0)));
            }
    printf("split potrf 32x32: %lld cycles [replica loads %lld, chain %lld, transition %lld] | max err: L below the diagonal blocks %.2e, 1/L(c,c) %.2e, T_b L_bb - I %.2e\n",
           h[0], h[4], h[5], h[6], eL, ed, eT);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
