// micro-benchmark: how fast can ONE CTA pull L2-resident data (the back substitution's wall)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_ldg(const double* __restrict__ src, size_t n, double* out, long long* cyc) {
    double acc = 0;
    long long t0 = clock64();
    for (size_t i = threadIdx.x; i < n; i += blockDim.x * 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (i + q * blockDim.x < n) ? __ldcg(src + i + q * blockDim.x) : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += v[q];
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ldg128(const double2* __restrict__ src, size_t n2, double* out, long long* cyc) {
    double acc = 0;
    long long t0 = clock64();
    for (size_t i = threadIdx.x; i < n2; i += blockDim.x * 8) {
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (i + q * blockDim.x < n2) ? __ldcg(src + i + q * blockDim.x) : make_double2(0, 0);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += v[q].x + v[q].y;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// cp.async.bulk (TMA 1-D) 8 KB chunks into a 4-slot ring, mbarrier completion
__global__ void k_bulk(const double* __restrict__ src, size_t n, double* out, long long* cyc) {
    extern __shared__ __align__(128) unsigned char sm[];
    double* ring = reinterpret_cast<double*>(sm);                 // 16 x 1024 doubles
    __shared__ uint64_t bar[16];
    const int NS = 16;
    const size_t chunks = n / 1024;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t c) {
        const int s = c % NS;
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
        unsigned d = (unsigned)__cvta_generic_to_shared(ring + (size_t)s * 1024);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8192;" ::"r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 8192, [%2];" ::"r"(d), "l"(src + c * 1024), "r"(b) : "memory");
    };
    double acc = 0;
    long long t0 = clock64();
    if (threadIdx.x == 0) for (size_t c = 0; c < NS && c < chunks; ++c) issue(c);
    for (size_t c = 0; c < chunks; ++c) {
        const int s = c % NS;
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
        unsigned ph = (c / NS) & 1, ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
        for (int e = threadIdx.x; e < 1024; e += blockDim.x) acc += ring[(size_t)s * 1024 + e];
        __syncthreads();
        if (threadIdx.x == 0 && c + NS < chunks) issue(c + NS);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    const size_t n = 704 * 1024;   // 5.5 MB of doubles
    double *d, *out; long long* dc;
    cudaMalloc(&d, n * 8); cudaMalloc(&out, 8192); cudaMalloc(&dc, 64);
    cudaMemset(d, 0, n * 8);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192);
    for (int rep = 0; rep < 2; ++rep) {
        long long h;
        k_ldg<<<1, 256>>>(d, n, out, dc); cudaDeviceSynchronize(); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
        printf("LDG.64   256 thr x 8 in flight: %lld cycles, %.1f B/cycle\n", h, n * 8.0 / h);
        k_ldg<<<1, 1024>>>(d, n, out, dc); cudaDeviceSynchronize(); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
        printf("LDG.64  1024 thr x 8 in flight: %lld cycles, %.1f B/cycle\n", h, n * 8.0 / h);
        k_ldg128<<<1, 256>>>((double2*)d, n / 2, out, dc); cudaDeviceSynchronize(); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
        printf("LDG.128  256 thr x 8 in flight: %lld cycles, %.1f B/cycle\n", h, n * 8.0 / h);
        k_ldg128<<<1, 1024>>>((double2*)d, n / 2, out, dc); cudaDeviceSynchronize(); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
        printf("LDG.128 1024 thr x 8 in flight: %lld cycles, %.1f B/cycle\n", h, n * 8.0 / h);
        k_bulk<<<1, 256, 16 * 8192>>>(d, n, out, dc); cudaDeviceSynchronize(); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
        printf("cp.async.bulk 8 KB x 16 in flight: %lld cycles, %.1f B/cycle  (%s)\n", h, n * 8.0 / h, cudaGetErrorString(cudaGetLastError()));
    }
}
