// micro-benchmark: dependent-issue latency of the FP64 ops on the pivot chain
#include <cstdio>
#include <cuda_runtime.h>
#define CHAIN(name, init, body)                                                  \
    __global__ void name(double* out, long long* cyc, double seed) {             \
        double x = seed + threadIdx.x * 1e-9; init;                              \
        long long t0 = clock64();                                                \
        _Pragma("unroll 1") for (int r = 0; r < 64; ++r) {                       \
            _Pragma("unroll") for (int q = 0; q < 16; ++q) { body; }             \
        }                                                                        \
        long long t1 = clock64();                                                \
        out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;            \
    }
CHAIN(k_dfma, double y = 1.0000001, x = fma(x, y, 1e-9))
CHAIN(k_dmul, double y = 1.0000001, x = x * y)
CHAIN(k_dadd, double y = 1e-9, x = x + y)
CHAIN(k_rsq64h, , asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x)); x = x + 1.5)
CHAIN(k_rcp64h, , asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x)); x = x + 1.5)
CHAIN(k_rsqf, , x = (double)rsqrtf((float)x) + 1.5)
CHAIN(k_shfl, , x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31))
CHAIN(k_sel, , x = (threadIdx.x == 5) ? x * 1.0000001 : x)
CHAIN(k_ffma, float f = (float)x, f = fmaf(f, 1.0000001f, 1e-9f); x = f)
__global__ void k_smem(double* out, long long* cyc, double seed) {
    __shared__ double buf[64];
    buf[threadIdx.x] = seed; __syncwarp();
    double x = seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 64; ++r) {
#pragma unroll
        for (int q = 0; q < 16; ++q) { ((volatile double*)buf)[threadIdx.x] = x; __syncwarp(); x = ((volatile double*)buf)[(threadIdx.x + 1) & 31] + 1e-9; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* dout; long long* dc; cudaMalloc(&dout, 1024); cudaMalloc(&dc, 64);
#define RUN(k, note) { k<<<1, 32>>>(dout, dc, 1.25); cudaDeviceSynchronize(); k<<<1, 32>>>(dout, dc, 1.25); cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("%-28s %.1f cycles per dependent step %s\n", #k, h / 1024.0, note); }
    RUN(k_dfma, "") RUN(k_dmul, "") RUN(k_dadd, "") RUN(k_rsq64h, "(MUFU.RSQ64H + DADD)") RUN(k_rcp64h, "(MUFU.RCP64H + DADD)")
    RUN(k_rsqf, "(F2F + MUFU.RSQ + F2F + DADD)") RUN(k_shfl, "(64-bit shuffle)") RUN(k_sel, "") RUN(k_ffma, "(FFMA + F2F.F64.F32 ... per step incl. cvt)") RUN(k_smem, "(STS + syncwarp + LDS + DADD)")
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
