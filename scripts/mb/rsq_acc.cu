#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double* out) {
    double m0 = 0, m1 = 0, m2 = 0, mf = 0;
    for (int i = threadIdx.x; i < (1 << 20); i += blockDim.x) {
        double x = exp2((i % 41) - 20.0) * (1.0 + (i * 0.6180339887498949 - floor(i * 0.6180339887498949)));
        double r0; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
        double ex = 1.0 / sqrt(x);
        double e = fma(-x * r0, r0, 1.0); double r1 = fma(0.5 * r0, e, r0);
        double e2 = fma(-x * r1, r1, 1.0); double r2 = fma(0.5 * r1, e2, r1);
        double rf = (double)rsqrtf((float)x); double ef = fma(-x * rf, rf, 1.0); double rf1 = fma(0.5 * rf, ef, rf);
        m0 = fmax(m0, fabs(r0 - ex) / ex); m1 = fmax(m1, fabs(r1 - ex) / ex); m2 = fmax(m2, fabs(r2 - ex) / ex); mf = fmax(mf, fabs(rf1 - ex) / ex);
    }
    out[4 * threadIdx.x] = m0; out[4 * threadIdx.x + 1] = m1; out[4 * threadIdx.x + 2] = m2; out[4 * threadIdx.x + 3] = mf;
}
int main() {
    double* d; cudaMalloc(&d, 256 * 32); k<<<1, 256>>>(d); double h[1024]; cudaMemcpy(h, d, 8192, cudaMemcpyDeviceToHost);
    double m[4] = {0, 0, 0, 0}; for (int i = 0; i < 256; ++i) for (int q = 0; q < 4; ++q) m[q] = fmax(m[q], h[4 * i + q]);
    printf("rsqrt.approx.ftz.f64 seed max rel err %.3e | +1 Newton %.3e | +2 Newton %.3e | float seed +1 Newton %.3e\n", m[0], m[1], m[2], m[3]);
}
