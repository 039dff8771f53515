// Per-sample normalisations HighOrderMLP puts between layers, each as ONE fused
// forward and ONE fused backward kernel instead of the ~8 + ~10 elementwise / reduction launches
// autograd issues for the reference expressions:
//   max_abs     y = x / (max_i |x_i| + eps)                           utils.py:8-13,  layers.py:70-81
//   max_center  mid = (max+min)/2, mag = max - mid, y = (x-mid)/(mag+eps)   utils.py:20-28, layers.py:84-96
//   l2          y = x / (||x||_2 + eps)                               utils.py:57-58, layers.py:132-141
// HBM-bound: one warp per row, the row is read once from HBM (second sweep from L1/L2).  The
// forward keeps the reference's op order (true division, separately rounded adds), so max_abs /
// max_center outputs are bit-identical to torch's; the backward applies the closed forms autograd
// derives (the max / min gradients go to the FIRST index attaining them, like torch.max(dim)).
#include "kernels.cuh"

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// (value, first index) arg-max over the warp; ties -> smallest index
__device__ __forceinline__ void warp_argmax(float& v, int& idx) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, m);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, m);
        if (ov > v || (ov == v && oi < idx)) v = ov, idx = oi;
    }
}

// stats[row] = {den, aux, idx_a, idx_b}: max_abs {m+eps, -, argmax, -}; max_center {mag+eps, mid,
// argmax, argmin}; l2 {norm+eps, norm, -, -}
template <int KIND>
__global__ void __launch_bounds__(256) pw_row_norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              float4* __restrict__ stats, int64_t B, int W, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    const float* xr = x + row * W;
    float* yr = y + row * W;
    if (KIND == HOL_NORM_L2) {
        float ss = 0.0f;
        for (int c = lane; c < W; c += 32) ss = fmaf(xr[c], xr[c], ss);
        const float nrm = sqrtf(warp_sum(ss));
        const float den = __fadd_rn(nrm, eps);
        for (int c = lane; c < W; c += 32) yr[c] = __fdiv_rn(xr[c], den);
        if (lane == 0) stats[row] = make_float4(den, nrm, 0.0f, 0.0f);
        return;
    }
    // torch.max / torch.min propagate NaN: a NaN anywhere in the row makes the whole row NaN (and takes the
    // arg-max), so NaN is treated as larger than everything; the stored indices always lie inside the row
    // (an all-NaN row -- the normal state after a divergence -- must not send the backward out of bounds).
    float vmax = -INFINITY, vmin = -INFINITY;  // vmin tracks -x
    int imax = W, imin = W;
    bool nan_seen = false;
    for (int c = lane; c < W; c += 32) {
        const float v = xr[c];
        const float a = (KIND == HOL_NORM_MAX_ABS) ? fabsf(v) : v;
        if (v != v) nan_seen = true;
        if (a > vmax || c < imax && a == vmax || imax == W) vmax = a, imax = c;
        if (KIND == HOL_NORM_MAX_CENTER && (-v > vmin || c < imin && -v == vmin || imin == W)) vmin = -v, imin = c;
    }
    nan_seen = __any_sync(0xffffffffu, nan_seen);
    warp_argmax(vmax, imax);
    if (imax >= W) imax = 0;
    if (nan_seen) vmax = NAN;
    if (KIND == HOL_NORM_MAX_ABS) {
        const float den = __fadd_rn(vmax, eps);
        for (int c = lane; c < W; c += 32) yr[c] = __fdiv_rn(xr[c], den);
        if (lane == 0) stats[row] = make_float4(den, 0.0f, __int_as_float(imax), 0.0f);
    } else {
        warp_argmax(vmin, imin);
        if (imin >= W) imin = 0;
        const float mn = nan_seen ? NAN : -vmin;
        const float mid = __fmul_rn(0.5f, __fadd_rn(vmax, mn));
        const float den = __fadd_rn(__fsub_rn(vmax, mid), eps);
        for (int c = lane; c < W; c += 32) yr[c] = __fdiv_rn(__fsub_rn(xr[c], mid), den);
        if (lane == 0) stats[row] = make_float4(den, mid, __int_as_float(imax), __int_as_float(imin));
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) pw_row_norm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                              const float4* __restrict__ stats, float* __restrict__ dx,
                                                              int64_t B, int W) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    const float* xr = x + row * W;
    const float* gr = gy + row * W;
    float* dr = dx + row * W;
    const float4 st = stats[row];
    const float den = st.x;
    float sg = 0.0f, sgx = 0.0f;
    for (int c = lane; c < W; c += 32) {
        const float g = gr[c];
        sg += g;
        sgx = fmaf(g, (KIND == HOL_NORM_MAX_CENTER) ? xr[c] - st.y : xr[c], sgx);
    }
    sg = warp_sum(sg);
    sgx = warp_sum(sgx);
    const float inv = 1.0f / den;
    if (KIND == HOL_NORM_MAX_ABS) {
        const int im = min(max(__float_as_int(st.z), 0), W - 1);
        const float xm = xr[im];
        const float sgn = xm > 0.0f ? 1.0f : (xm < 0.0f ? -1.0f : 0.0f);
        const float corr = -sgx * inv * inv * sgn;
        for (int c = lane; c < W; c += 32) dr[c] = gr[c] * inv + (c == im ? corr : 0.0f);
    } else if (KIND == HOL_NORM_MAX_CENTER) {
        // y = (x - mid) / d, d = mag + eps:  dL/dmid = -sum(g)/d,  dL/dd = -sum(g (x - mid))/d^2;
        // d mid / d x_max = d mid / d x_min = 1/2,  d mag / d x_max = 1/2,  d mag / d x_min = -1/2
        const int ia = min(max(__float_as_int(st.z), 0), W - 1), ib = min(max(__float_as_int(st.w), 0), W - 1);
        const float dmid = -sg * inv, dd = -sgx * inv * inv;
        const float ca = 0.5f * (dmid + dd), cb = 0.5f * (dmid - dd);
        for (int c = lane; c < W; c += 32) dr[c] = gr[c] * inv + (c == ia ? ca : 0.0f) + (c == ib ? cb : 0.0f);
    } else {
        const float nrm = st.y;
        const float k = nrm > 0.0f ? -sgx * inv * inv / nrm : 0.0f;
        for (int c = lane; c < W; c += 32) dr[c] = fmaf(k, xr[c], gr[c] * inv);
    }
}

extern "C" {

int hol_row_norm_forward_f32(int32_t kind, const float* x, float* y, float* stats, int64_t B, int32_t W, float eps,
                             void* stream) {
    if (B < 0 || W <= 0 || kind < 0 || kind > 2) return HOL_EINVAL;
    if (B == 0) return HOL_OK;
    if (!x || !y || !stats) return HOL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((B + 7) / 8);
    float4* s4 = reinterpret_cast<float4*>(stats);
    if (kind == HOL_NORM_MAX_ABS) pw_row_norm_fwd_kernel<HOL_NORM_MAX_ABS><<<grid, 256, 0, st>>>(x, y, s4, B, W, eps);
    if (kind == HOL_NORM_MAX_CENTER) pw_row_norm_fwd_kernel<HOL_NORM_MAX_CENTER><<<grid, 256, 0, st>>>(x, y, s4, B, W, eps);
    if (kind == HOL_NORM_L2) pw_row_norm_fwd_kernel<HOL_NORM_L2><<<grid, 256, 0, st>>>(x, y, s4, B, W, eps);
    return (int)cudaGetLastError();
}

int hol_row_norm_backward_f32(int32_t kind, const float* gy, const float* x, const float* stats, float* dx, int64_t B,
                              int32_t W, void* stream) {
    if (B < 0 || W <= 0 || kind < 0 || kind > 2) return HOL_EINVAL;
    if (B == 0) return HOL_OK;
    if (!gy || !x || !stats || !dx) return HOL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((B + 7) / 8);
    const float4* s4 = reinterpret_cast<const float4*>(stats);
    if (kind == HOL_NORM_MAX_ABS) pw_row_norm_bwd_kernel<HOL_NORM_MAX_ABS><<<grid, 256, 0, st>>>(gy, x, s4, dx, B, W);
    if (kind == HOL_NORM_MAX_CENTER) pw_row_norm_bwd_kernel<HOL_NORM_MAX_CENTER><<<grid, 256, 0, st>>>(gy, x, s4, dx, B, W);
    if (kind == HOL_NORM_L2) pw_row_norm_bwd_kernel<HOL_NORM_L2><<<grid, 256, 0, st>>>(gy, x, s4, dx, B, W);
    return (int)cudaGetLastError();
}

}  // extern "C"
