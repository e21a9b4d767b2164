// Per-frame producer: depth image -> candidate surfels, as dense per-pixel maps.
// Restates depth_preprocessing for --load_depth inputs
//   /root/reference/utils/data_loader.py:333-523  (getN :532-583,
//   BackprojectDepth /root/reference/depth/monodepth2/layers.py:139-167)
// in two stencil passes instead of ~150 ATen launches, and keeps everything on the device (the
// reference leaves index_map on the CPU, data_loader.py:464).
//
// float32 arithmetic follows the reference's op order and roundings instruction for instruction (points, normals,
// radii and confidences are compared BIT-EXACTLY against oracle/super_oracle.py preprocess in the tests): torch's CPU
// kernels' summation order and FMA placement as measured (oracle normals_8 docstring), the IEEE sqrt / divide, and the
// exponential DEFINED in the oracle (exp32_def: the reference's is MKL's vsExp, not reproducible from source).
#include "common.cuh"
#include "super_b200.h"

namespace {

struct PreArgs {
    const float* depth;          // (H,W) f32
    const float* color;          // (3,H,W) f32 planar
    const unsigned char* inval;  // (H,W) optional extra invalid mask (valid-mask / del_seg_classes / morphology)
    float ik[9];                 // inv_K[:3,:3] row-major, float32
    float fx;                    // K[0,0] float32
    float divterm;
    float depth_max;             // superv1: 1.5 ; superv2: +inf
    int zero_invalid;            // superv2: depth == 0 invalid ; superv1: depth <= 0 invalid
    int mask_rows;               // superv2: first int(0.1*W) ROWS invalid (reference quirk, Appendix B #10)
    int H, W;
};

// pass 1: back-projection + depth validity -> pcd (float4: x,y,z,depth; NaN where invalid)
__global__ void backproject_kernel(PreArgs a, float4* __restrict__ pcd) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.W) return;
    const int p = y * a.W + x;
    const float d = a.depth[p];
    bool bad = a.zero_invalid ? (d == 0.f) : (d <= 0.f);
    bad |= d > a.depth_max;
    bad |= y < a.mask_rows;
    if (a.inval) bad |= a.inval[p] != 0;
    const float fxp = (float)x, fyp = (float)y;
    float4 o;
    if (bad) {
        o.x = o.y = o.z = o.w = __int_as_float(0x7fc00000);
    } else {
        // matmul(inv_K[:3,:3], [x,y,1]) then depth * (.)   (layers.py:162-163).  The reference's CPU BLAS
        // evaluates each entry as t = a*x; t = fma(b,y,t); t = fma(c,1,t) (measured bit for bit, see oracle).
        const float cx_ = __fadd_rn(__fmaf_rn(a.ik[1], fyp, __fmul_rn(a.ik[0], fxp)), a.ik[2]);
        const float cy_ = __fadd_rn(__fmaf_rn(a.ik[4], fyp, __fmul_rn(a.ik[3], fxp)), a.ik[5]);
        const float cz_ = __fadd_rn(__fmaf_rn(a.ik[7], fyp, __fmul_rn(a.ik[6], fxp)), a.ik[8]);
        o.x = __fmul_rn(d, cx_); o.y = __fmul_rn(d, cy_); o.z = __fmul_rn(d, cz_); o.w = d;
    }
    pcd[p] = o;
}

// exp32_def of oracle/super_oracle.py, same operations in the same order: float64, no FMA contraction
__device__ __forceinline__ float exp32_def(float xf) {
    const double x = (double)xf;
    const double n = rint(__dmul_rn(x, 0x1.71547652b82fep+0));
    const double r = __dsub_rn(__dsub_rn(x, __dmul_rn(n, 0x1.62e42fee00000p-1)), __dmul_rn(n, 0x1.a39ef35793c76p-33));
    const double c[14] = {0x1.0000000000000p+0,  0x1.0000000000000p+0,  0x1.0000000000000p-1,  0x1.5555555555555p-3,
                          0x1.5555555555555p-5,  0x1.1111111111111p-7,  0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13,
                          0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19, 0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26,
                          0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33};
    double p = c[13];
#pragma unroll
    for (int k = 12; k >= 0; --k) p = __dadd_rn(__dmul_rn(p, r), c[k]);
    return __double2float_rn(ldexp(p, (int)n));
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 add3(F3 a, F3 b) { return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ F3 crossf(F3 a, F3 b) {   // torch.cross CPU pattern: fma(a1,b2,-(a2*b1))
    return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
              __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// pass 2: 8-neighbour colour-weighted normals, validity, radii, confidences
__global__ void normals_kernel(PreArgs a, const float4* __restrict__ pcd, float4* __restrict__ vmap,
                               float4* __restrict__ nmap, double* __restrict__ radii, float* __restrict__ confs,
                               int* __restrict__ valid_i32) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.W) return;
    const int W = a.W, H = a.H, P = W * H, p = y * W + x;
    const float nanf_ = __int_as_float(0x7fc00000);
    const float4 pc = pcd[p];
    const float c0 = a.color[p], c1 = a.color[P + p], c2 = a.color[2 * P + p];
    // ring order L, LU, U, RU, R, RD, D, DL   (data_loader.py:553-570)
    const int dxs[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int dys[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    F3 h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int xx = x + dxs[k], yy = y + dys[k];
        if (xx < 0 || xx >= W || yy < 0 || yy >= H) { h[k] = f3(nanf_, nanf_, nanf_); continue; }
        const int q = yy * W + xx;
        const float4 pa = pcd[q];
        const float m = __fdiv_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(a.color[q], c0)), fabsf(__fsub_rn(a.color[P + q], c1))),
                                            fabsf(__fsub_rn(a.color[2 * P + q], c2))), 3.f);
        const float wgt = exp32_def(-m);
        h[k] = f3(__fmul_rn(__fsub_rn(pa.x, pc.x), wgt), __fmul_rn(__fsub_rn(pa.y, pc.y), wgt),
                  __fmul_rn(__fsub_rn(pa.z, pc.z), wgt));
    }
    F3 t[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        F3 rest = h[i + 1];
#pragma unroll
        for (int j = i + 2; j < 8; ++j) rest = add3(rest, h[j]);
        t[i] = crossf(h[i], rest);
    }
    // torch's row sum over 7 contiguous floats (4 accumulators): (((((t0+t4)+t5)+t6)+t1)+t2)+t3
    F3 N = add3(add3(add3(add3(add3(add3(t[0], t[4]), t[5]), t[6]), t[1]), t[2]), t[3]);
    // F.normalize: vector_norm accumulates by FMA, z*z + (y*y + x*x); IEEE sqrt and divide
    const float nn = fmaxf(__fsqrt_rn(__fmaf_rn(N.z, N.z, __fmaf_rn(N.y, N.y, __fmul_rn(N.x, N.x)))), 1e-12f);
    N = f3(__fdiv_rn(N.x, nn), __fdiv_rn(N.y, nn), __fdiv_rn(N.z, nn));
    const bool valid = !(isnan(N.x) || isnan(N.y) || isnan(N.z) || isnan(pc.x) || isnan(pc.y) || isnan(pc.z));
    float4 v, n;
    if (valid) {
        v.x = pc.x; v.y = pc.y; v.z = pc.z; v.w = 1.f;
        n.x = N.x; n.y = N.y; n.z = N.z; n.w = 0.f;
        // radii = (-depth) / (sqrt(2) * fx * clamp(|n_z|, 0.26, 1))   (data_loader.py:444,468-469)
        const float c = __fmul_rn((float)1.4142135623730951, a.fx);
        radii[p] = (double)(-pc.w) / ((double)c * fmin(fmax(fabs((double)N.z), 0.26), 1.0));
    } else {
        v.x = v.y = v.z = v.w = 0.f;
        n.x = n.y = n.z = n.w = 0.f;
        radii[p] = 0.0;
    }
    vmap[p] = v;
    nmap[p] = n;
    // confs = exp(-((2x/W-1)^2 + (2y/H-1)^2) * divterm)   (data_loader.py:472-475), float32
    const float su = __fsub_rn(__fmul_rn(2.f, __fdiv_rn((float)x, (float)W)), 1.f);
    const float sv = __fsub_rn(__fmul_rn(2.f, __fdiv_rn((float)y, (float)H)), 1.f);
    const float dc2 = __fadd_rn(__fmul_rn(su, su), __fmul_rn(sv, sv));
    confs[p] = exp32_def(__fmul_rn(-dc2, a.divterm));
    if (valid_i32) valid_i32[p] = valid ? 1 : 0;
}

// seg = argmax_c scores (first maximum), seg_conf = softmax_c scores     (data_loader.py:229-231,457)
__global__ void seg_maps_kernel(const double* __restrict__ scores, int C, int P, int* __restrict__ seg,
                                double* __restrict__ conf) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double mx = -INFINITY;
    int am = 0;
    for (int c = 0; c < C; ++c) {
        const double v = scores[(size_t)c * P + p];
        if (v > mx) { mx = v; am = c; }
    }
    double sum = 0.0, e[8];
    for (int c = 0; c < C; ++c) { e[c] = exp(scores[(size_t)c * P + p] - mx); sum += e[c]; }
    for (int c = 0; c < C; ++c) conf[(size_t)p * C + c] = e[c] / sum;
    seg[p] = am;
}

}  // namespace

extern "C" {

int sb_seg_maps(const double* scores, int C, int H, int W, int* seg, double* seg_conf, void* stream) {
    if (!scores || !seg || !seg_conf || C < 1 || C > 8 || H <= 0 || W <= 0) return SB_ERR_ARG;
    const int P = H * W;
    seg_maps_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scores, C, P, seg, seg_conf);
    SB_CHECK_LAUNCH();
    return SB_OK;
}


int sb_preprocess(const float* depth, const float* color, const unsigned char* inval, const float* inv_K3x3,
                  float fx, float divterm, int superv2, int H, int W, float* pcd_scratch, float* vmap, float* nmap,
                  double* radii, float* confs, int* valid_i32, void* stream) {
    if (!depth || !color || !inv_K3x3 || !pcd_scratch || !vmap || !nmap || !radii || !confs) return SB_ERR_ARG;
    if (H <= 2 || W <= 2) return SB_ERR_ARG;
    PreArgs a;
    a.depth = depth; a.color = color; a.inval = inval;
    for (int i = 0; i < 9; ++i) a.ik[i] = inv_K3x3[i];
    a.fx = fx; a.divterm = divterm;
    a.depth_max = superv2 ? INFINITY : 1.5f;
    a.zero_invalid = superv2 ? 1 : 0;
    a.mask_rows = superv2 ? (int)(0.1 * W) : 0;
    a.H = H; a.W = W;
    dim3 block(128), grid((W + 127) / 128, H);
    cudaStream_t s = (cudaStream_t)stream;
    backproject_kernel<<<grid, block, 0, s>>>(a, reinterpret_cast<float4*>(pcd_scratch));
    SB_CHECK_LAUNCH();
    normals_kernel<<<grid, block, 0, s>>>(a, reinterpret_cast<const float4*>(pcd_scratch),
                                         reinterpret_cast<float4*>(vmap), reinterpret_cast<float4*>(nmap), radii, confs,
                                         valid_i32);
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // extern "C"
