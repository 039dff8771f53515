// Shared definitions of the tensor-core path (see tc_gemm.cu for the design notes).
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_A_IMG (TC_BM * TC_BK * 2)  // bytes of one fp16 A image (hi or lo)
#define TC_MAXHALF 128                // register sums per epilogue thread (BN <= 256)

enum { EPI_Y = 0, EPI_DW = 1, EPI_DX = 2 };

struct TcScalars {      // device scalars written by pw_absmax_kernel (bits of non-negative floats)
    unsigned max_xl, max_w, max_gy;
};

// power-of-two scale that maps `bound` just below 2^14 (scale >= 1 is allowed: small values are
// scaled UP so that their lo parts stay out of the fp16 subnormals)
__host__ __device__ __forceinline__ float pow2_scale_for(float bound) {
    if (!(bound > 0.0f) || !(bound < 3.0e38f)) return 1.0f;
    int e;
    frexpf(bound, &e);  // bound = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, 14 - e);
}
// |phi_j(t)| <= max_j |C_j| * prod_m (|t| + |X_m|)   (conservative: phi_j skips one of the factors)
__device__ __forceinline__ float phi_scale(float max_xl, const BasisTab& tab, int n) {
    float cmax = 0.0f, prod = 1.0f;
    for (int j = 0; j < n; ++j) {
        cmax = fmaxf(cmax, fabsf(tab.C[j]));
        prod *= (max_xl + fabsf(tab.X[j]));
    }
    const float bound = (n == 1) ? 1.0f : cmax * prod;
    return fminf(pow2_scale_for(fmaxf(bound, 1.0f)), 1.0f);  // only ever scale phi down
}

__global__ void __launch_bounds__(256) pw_absmax_kernel(const float* __restrict__ v, int64_t count, unsigned* out) {
    float m = 0.0f;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const float a = fabsf(v[k]);
        if (a < 3.0e38f) m = fmaxf(m, a);
    }
    for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order like uints
}

__device__ __forceinline__ void split_store8(const float (&v)[8], float scale, uint4* hi, uint4* lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a = v[2 * e] * scale, b = v[2 * e + 1] * scale;
        const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        const __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
        h[e] = __halves2half2(ha, hb);
        l[e] = __halves2half2(la, lb);
    }
    *hi = *reinterpret_cast<uint4*>(h);
    *lo = *reinterpret_cast<uint4*>(l);
}

