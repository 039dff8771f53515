// Per-sample normalisations HighOrderMLP puts between layers, each as ONE fused
// forward and ONE fused backward kernel instead of the ~8 + ~10 elementwise / reduction launches
// autograd issues for the reference expressions (row math: norm_rows.cuh).  HBM-bound: one warp per
// row, the row is read once from HBM (second sweep from L1/L2).  A normalisation that directly follows
// a piecewise layer is folded into that layer's passes instead (hol_pw_forward_norm_f32 /
// hol_pw_norm_backward_f32): the forward's last kernel writes the normalised row next to the raw one,
// and the backward kernel below doubles as the layer's max|gy| statistics pass.
#include "kernels.cuh"
#include "norm_rows.cuh"

template <int KIND>
__global__ void __launch_bounds__(256) pw_row_norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              float4* __restrict__ stats, int64_t B, int W, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    norm_row_forward<KIND>(x + row * W, y + row * W, stats + row, W, eps, lane);
}

// absmax != nullptr: also max |dx| over the whole tensor (bits of a non-negative float; the caller zeroed it)
template <int KIND>
__global__ void __launch_bounds__(256) pw_row_norm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                              const float4* __restrict__ stats, float* __restrict__ dx,
                                                              int64_t B, int W, unsigned* absmax) {
    __shared__ unsigned s_max[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float amax = 0.0f;
    // a few rows per warp: the grid stays within ~4 CTAs per SM, so the statistic costs a few hundred atomics
    for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < B; row += (int64_t)gridDim.x * 8)
        amax = fmaxf(amax, norm_row_backward<KIND>(gy + row * W, x + row * W, stats[row], dx + row * W, W, lane));
    if (absmax) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, m));
        if (lane == 0) s_max[wid] = amax > 0.0f ? __float_as_uint(amax) : 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned bits = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) bits = max(bits, s_max[k]);
            if (bits > *reinterpret_cast<volatile unsigned*>(absmax)) atomicMax(absmax, bits);
        }
    }
}

int launch_row_norm_forward(int kind, const float* x, float* y, float* stats, int64_t B, int W, float eps, cudaStream_t st) {
    if (B == 0) return HOL_OK;
    const unsigned grid = (unsigned)((B + 7) / 8);
    float4* s4 = reinterpret_cast<float4*>(stats);
    if (kind == HOL_NORM_MAX_ABS) pw_row_norm_fwd_kernel<HOL_NORM_MAX_ABS><<<grid, 256, 0, st>>>(x, y, s4, B, W, eps);
    if (kind == HOL_NORM_MAX_CENTER) pw_row_norm_fwd_kernel<HOL_NORM_MAX_CENTER><<<grid, 256, 0, st>>>(x, y, s4, B, W, eps);
    if (kind == HOL_NORM_L2) pw_row_norm_fwd_kernel<HOL_NORM_L2><<<grid, 256, 0, st>>>(x, y, s4, B, W, eps);
    return (int)cudaGetLastError();
}

int launch_row_norm_backward(int kind, const float* gy, const float* x, const float* stats, float* dx, int64_t B, int W,
                             unsigned* absmax, cudaStream_t st) {
    if (B == 0) return HOL_OK;
    int64_t blocks = (B + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const unsigned grid = (unsigned)blocks;
    const float4* s4 = reinterpret_cast<const float4*>(stats);
    if (kind == HOL_NORM_MAX_ABS) pw_row_norm_bwd_kernel<HOL_NORM_MAX_ABS><<<grid, 256, 0, st>>>(gy, x, s4, dx, B, W, absmax);
    if (kind == HOL_NORM_MAX_CENTER) pw_row_norm_bwd_kernel<HOL_NORM_MAX_CENTER><<<grid, 256, 0, st>>>(gy, x, s4, dx, B, W, absmax);
    if (kind == HOL_NORM_L2) pw_row_norm_bwd_kernel<HOL_NORM_L2><<<grid, 256, 0, st>>>(gy, x, s4, dx, B, W, absmax);
    return (int)cudaGetLastError();
}

extern "C" {

int hol_row_norm_forward_f32(int32_t kind, const float* x, float* y, float* stats, int64_t B, int32_t W, float eps,
                             void* stream) {
    if (B < 0 || W <= 0 || kind < 0 || kind > 2) return HOL_EINVAL;
    if (B == 0) return HOL_OK;
    if (!x || !y || !stats) return HOL_EINVAL;
    return launch_row_norm_forward(kind, x, y, stats, B, W, eps, (cudaStream_t)stream);
}

int hol_row_norm_backward_f32(int32_t kind, const float* gy, const float* x, const float* stats, float* dx, int64_t B,
                              int32_t W, void* stream) {
    if (B < 0 || W <= 0 || kind < 0 || kind > 2) return HOL_EINVAL;
    if (B == 0) return HOL_OK;
    if (!gy || !x || !stats || !dx) return HOL_EINVAL;
    return launch_row_norm_backward(kind, gy, x, stats, dx, B, W, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
