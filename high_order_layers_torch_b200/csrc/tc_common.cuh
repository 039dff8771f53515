// Shared definitions of the tensor-core path (see tc_gemm.cu for the design notes).
#pragma once
#include <cuda_fp16.h>
#include <math.h>

#include "kernels.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_A_IMG (TC_BM * TC_BK * 2)  // bytes of one fp16 A image (hi or lo)
#define TC_MAXHALF 128                // register sums per epilogue thread (BN <= 256)

enum { EPI_Y = 0, EPI_DW = 1, EPI_DX = 2 };

// Device scalars of one layer call (workspace region tc_scalars, zeroed by ONE memset at the start
// of the forward): maxima as bits of non-negative floats (they order like unsigned integers).
struct TcScalars {
    unsigned max_xl;      // max |xl| over the elements that are NOT outliers (written by the prep kernels)
    unsigned max_w;       // max |w|             (pw_tc_wstats_kernel)
    unsigned max_gy;      // max |gy|            (pw_absmax_kernel, backward)
    unsigned n_outliers;  // elements flagged HOL_SEG_OUTLIER (prep kernels); 0 => the fix-up kernels return at once
};

// power-of-two scale that maps `bound` just below 2^14 (scale >= 1 is allowed: small values are
// scaled UP so that their lo parts stay out of the fp16 subnormals)
__host__ __device__ __forceinline__ float pow2_scale_for(float bound) {
    if (!(bound > 0.0f) || !(bound < 3.0e38f)) return 1.0f;
    int e;
    frexpf(bound, &e);  // bound = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, 14 - e);
}

// sup |phi_j(t)| over |t| <= T.  Between the Chebyshev-Lobatto nodes |phi_j| <= 1; outside [-1, 1] every
// |phi_j| grows monotonically, so the supremum sits at t = +-T:  |C_j| * prod_{m != j} |T -+ X_m|.
__host__ __device__ __forceinline__ float phi_bound(float T, const BasisTab& tab, int n) {
    float best = 1.0f;
    if (n == 1 || !(T > 1.0f)) return best;
    for (int j = 0; j < n; ++j) {
        float pp = fabsf(tab.C[j]), pm = pp;
        for (int m = 0; m < n; ++m) {
            if (m == j) continue;
            pp *= fabsf(T - tab.X[m]);
            pm *= fabsf(-T - tab.X[m]);
        }
        best = fmaxf(best, fmaxf(pp, pm));
    }
    return best * 1.0001f;
}

// The Phi tiles are scaled by a power of two in [2^-3, 1]: elements whose basis values would need a
// smaller scale (|xl| > xl_limit, tc_xl_limit) never enter the tiles -- they are flagged
// HOL_SEG_OUTLIER by the prep kernels and added in fp32 by the fix-up kernels -- so one wild input
// cannot push the in-range samples of the batch into the fp16 subnormals.
#define TC_PHI_MIN_SCALE_LOG2 (-3)
__host__ __device__ __forceinline__ float phi_scale(float max_xl, const BasisTab& tab, int n) {
    const float s = pow2_scale_for(phi_bound(max_xl, tab, n));
    return fminf(fmaxf(s, ldexpf(1.0f, TC_PHI_MIN_SCALE_LOG2)), 1.0f);  // only ever scale phi down
}
// largest |xl| whose basis values still fit the fp16 range at the minimum scale
static inline float tc_xl_limit(const BasisTab& tab, int n) {
    if (n == 1) return INFINITY;
    const float cap = ldexpf(1.0f, 14 - TC_PHI_MIN_SCALE_LOG2);  // bound * 2^-3 < 2^14
    float lo = 1.0f, hi = 1.0e6f;
    if (phi_bound(hi, tab, n) < cap) return hi;
    for (int it = 0; it < 60; ++it) {
        const float mid = 0.5f * (lo + hi);
        if (phi_bound(mid, tab, n) < cap) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) pw_absmax_kernel(const float* __restrict__ v, int64_t count, unsigned* out) {
    float m = 0.0f;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const float a = fabsf(v[k]);
        if (a < 3.0e38f) m = fmaxf(m, a);
    }
    for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));  // non-negative floats order like uints
}

// a = hi + lo in fp16 after the power-of-two pre-scale.  A non-zero value too small for even the lo
// part (|a * scale| < 2^-25) keeps the smallest fp16 subnormal in lo, so that products with it stay
// non-zero: dW must be exactly zero only where the reference's is (SparseLion keys on grad != 0).
__device__ __forceinline__ void split_store8(const float (&v)[8], float scale, uint4* hi, uint4* lo) {
    __half2 h[4], l[4];
    const __half tiny = __ushort_as_half((unsigned short)0x0001), ntiny = __ushort_as_half((unsigned short)0x8001);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a = v[2 * e] * scale, b = v[2 * e + 1] * scale;
        const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
        // |a| < 2^-25 rounds to hi = lo = 0: keep the sign and the smallest subnormal instead
        if (fabsf(a) <= 2.9802323e-8f && a != 0.0f) la = a > 0.0f ? tiny : ntiny;
        if (fabsf(b) <= 2.9802323e-8f && b != 0.0f) lb = b > 0.0f ? tiny : ntiny;
        h[e] = __halves2half2(ha, hb);
        l[e] = __halves2half2(la, lb);
    }
    *hi = *reinterpret_cast<uint4*>(h);
    *lo = *reinterpret_cast<uint4*>(l);
}
