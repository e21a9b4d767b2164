// micro-benchmark: latency of dependent / interleaved DMMA m8n8k4 chains with different operand data
#include <cstdio>
#include <cuda_runtime.h>
#include <cmath>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a_, double b_) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a_), "d"(b_));
}
constexpr int S36 = 36;
__global__ void probe(const double* A, const double* B, double* out, long long* cyc, int mode) {
    __shared__ double As[32 * S36], Bs[32 * S36];
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) { As[(e >> 5) * S36 + (e & 31)] = A[e]; Bs[(e >> 5) * S36 + (e & 31)] = B[e]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fr = lane >> 2, fc = lane & 3;
    double acc0 = 0, acc1 = 0, tot = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < 100; ++rep) {
        if (mode == 0) {           // 20 serial DMMAs (one accumulator), triangular k-range
            for (int bj = 0; bj < 4; ++bj) {
                double c0 = 0, c1 = 0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) if (ks < 2 * bj + 2) dmma884(c0, c1, As[(8 * warp + fr) * S36 + 4 * ks + fc], Bs[(8 * bj + fr) * S36 + 4 * ks + fc]);
                tot += c0 + c1;
            }
        } else {                    // full 8-step chains, 3 blocks
            for (int bj = 0; bj < 3; ++bj) {
                double c0 = 0, c1 = 0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) dmma884(c0, c1, As[(8 * warp + fr) * S36 + 4 * ks + fc], Bs[(8 * bj + fr) * S36 + 4 * ks + fc]);
                tot += c0 + c1;
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = (t1 - t0) / 100;
    out[threadIdx.x] = tot + acc0 + acc1;
}
int main() {
    double hA[1024], hB[1024];
    double *dA, *dB, *dout; long long* dc;
    cudaMalloc(&dA, 8192); cudaMalloc(&dB, 8192); cudaMalloc(&dout, 8192); cudaMalloc(&dc, 64);
    const char* names[] = {"randn x randn", "randn x lower-tri(zeros above)", "randn x tiny(1e-200)", "randn x denormal(1e-310)", "randn x ones", "zeros x randn"};
    for (int data = 0; data < 6; ++data) {
        srand(1);
        for (int e = 0; e < 1024; ++e) {
            double r1 = (rand() / (double)RAND_MAX - 0.5) * 2, r2 = (rand() / (double)RAND_MAX - 0.5) * 2;
            hA[e] = r1; hB[e] = r2;
            int i = e >> 5, j = e & 31;
            if (data == 1 && j > i) hB[e] = 0.0;
            if (data == 2) hB[e] = r2 * 1e-200;
            if (data == 3) hB[e] = r2 * 1e-310;
            if (data == 4) hB[e] = 1.0;
            if (data == 5) hA[e] = 0.0;
        }
        cudaMemcpy(dA, hA, 8192, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 8192, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 2; ++mode)
            for (int nw = 1; nw <= 4; nw *= 4) {
                probe<<<1, 32 * nw>>>(dA, dB, dout, dc, mode); cudaDeviceSynchronize();
                probe<<<1, 32 * nw>>>(dA, dB, dout, dc, mode); cudaDeviceSynchronize();
                long long hc[8]; cudaMemcpy(hc, dc, 64, cudaMemcpyDeviceToHost);
                printf("%-34s mode %d (%s) warps %d: %lld cycles per %d DMMAs = %.1f cyc/DMMA\n", names[data], mode, mode ? "3x8 chains" : "2+4+6+8 chains", nw, hc[0], mode ? 24 : 20, hc[0] / (mode ? 24.0 : 20.0));
            }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
