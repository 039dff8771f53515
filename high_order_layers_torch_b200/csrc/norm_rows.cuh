// Row-wise pieces of the per-sample normalisations (one warp per row), shared by the standalone kernels
// (norm.cu) and by the kernels that fold a normalisation into a layer's last forward pass / first
// backward pass (tc_fixup.cuh: pw_tc_finalize_kernel; norm.cu: pw_row_norm_bwd_kernel with statistics).
//   max_abs     y = x / (max_i |x_i| + eps)                           utils.py:8-13,  layers.py:70-81
//   max_center  mid = (max+min)/2, mag = max - mid, y = (x-mid)/(mag+eps)   utils.py:20-28, layers.py:84-96
//   l2          y = x / (||x||_2 + eps)                               utils.py:57-58, layers.py:132-141
// The forward keeps the reference's op order (true division, separately rounded adds), so max_abs /
// max_center outputs are bit-identical to torch's; the backward applies the closed forms autograd
// derives (the max / min gradients go to the FIRST index attaining them, like torch.max(dim)).
#pragma once
#include <math.h>

#include "common.cuh"

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

// stats = {den, aux, idx_a, idx_b}: max_abs {m+eps, -, argmax, -}; max_center {mag+eps, mid,
// argmax, argmin}; l2 {norm+eps, norm, -, -}.  xr and yr may alias.
template <int KIND>
__device__ __forceinline__ void norm_row_forward(const float* xr, float* yr, float4* stat, int W, float eps, int lane) {
    if (KIND == HOL_NORM_L2) {
        float ss = 0.0f;
        for (int c = lane; c < W; c += 32) ss = fmaf(xr[c], xr[c], ss);
        const float nrm = sqrtf(warp_sum(ss));
        const float den = __fadd_rn(nrm, eps);
        for (int c = lane; c < W; c += 32) yr[c] = __fdiv_rn(xr[c], den);
        if (lane == 0) *stat = make_float4(den, nrm, 0.0f, 0.0f);
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
        if (lane == 0) *stat = make_float4(den, 0.0f, __int_as_float(imax), 0.0f);
    } else {
        warp_argmax(vmin, imin);
        if (imin >= W) imin = 0;
        const float mn = nan_seen ? NAN : -vmin;
        const float mid = __fmul_rn(0.5f, __fadd_rn(vmax, mn));
        const float den = __fadd_rn(__fsub_rn(vmax, mid), eps);
        for (int c = lane; c < W; c += 32) yr[c] = __fdiv_rn(__fsub_rn(xr[c], mid), den);
        if (lane == 0) *stat = make_float4(den, mid, __int_as_float(imax), __int_as_float(imin));
    }
}

// dx of one row; returns max |dx| over this lane's elements (finite values only)
template <int KIND>
__device__ __forceinline__ float norm_row_backward(const float* gr, const float* xr, float4 st, float* dr, int W, int lane) {
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
    float amax = 0.0f;
    auto put = [&](int c, float v) {
        dr[c] = v;
        const float a = fabsf(v);
        if (a < 3.0e38f) amax = fmaxf(amax, a);
    };
    if (KIND == HOL_NORM_MAX_ABS) {
        const int im = min(max(__float_as_int(st.z), 0), W - 1);
        const float xm = xr[im];
        const float sgn = xm > 0.0f ? 1.0f : (xm < 0.0f ? -1.0f : 0.0f);
        const float corr = -sgx * inv * inv * sgn;
        for (int c = lane; c < W; c += 32) put(c, gr[c] * inv + (c == im ? corr : 0.0f));
    } else if (KIND == HOL_NORM_MAX_CENTER) {
        // y = (x - mid) / d, d = mag + eps:  dL/dmid = -sum(g)/d,  dL/dd = -sum(g (x - mid))/d^2;
        // d mid / d x_max = d mid / d x_min = 1/2,  d mag / d x_max = 1/2,  d mag / d x_min = -1/2
        const int ia = min(max(__float_as_int(st.z), 0), W - 1), ib = min(max(__float_as_int(st.w), 0), W - 1);
        const float dmid = -sg * inv, dd = -sgx * inv * inv;
        const float ca = 0.5f * (dmid + dd), cb = 0.5f * (dmid - dd);
        for (int c = lane; c < W; c += 32) put(c, gr[c] * inv + (c == ia ? ca : 0.0f) + (c == ib ? cb : 0.0f));
    } else {
        const float nrm = st.y;
        const float k = nrm > 0.0f ? -sgx * inv * inv / nrm : 0.0f;
        for (int c = lane; c < W; c += 32) put(c, fmaf(k, xr[c], gr[c] * inv));
    }
    return amax;
}
