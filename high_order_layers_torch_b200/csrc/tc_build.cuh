// Operand builders of the tensor-core path (HBM-bound): see tc_gemm.cu.
#pragma once
#include "tc_common.cuh"

// ---------------------------------------------------------------------------------------
// operand builders (HBM-bound).  Image layout of a tile with R rows: [k/8 (8)][row (R)][8 x fp16],
// hi image then lo image.
// ---------------------------------------------------------------------------------------

// forward A: Phi tiles [m-tile (128 samples)][k-chunk], k = i*Kin + v.
// One thread = one row x 32 dense columns (4 of the 8 k-groups of the chunk).  The n non-zero values
// of an input sit at a column offset that depends on the sample's segment; instead of selecting every
// one of the 32 columns out of the n basis values (n compares per column, the first version of this
// kernel was issue bound at half the HBM rate), the thread scatters its (hi, lo) pairs into a private,
// zeroed 32-word strip of shared memory -- shared memory takes run-time addresses, registers do not --
// and reads the strip back as eight 16-byte chunks, which a byte permute splits into the hi and the lo
// image.  HBM-bound: 16-byte coalesced stores of 2 x 32 KB per tile.
template <int N>
__global__ void __launch_bounds__(256) pw_tc_expand_kernel(const float* __restrict__ xl_t, const uint16_t* __restrict__ seg_t,
                                                           unsigned char* __restrict__ tiles, const TcScalars* sc,
                                                           int64_t Bp, int64_t row0, int I, int Kin, int nkc, int stride,
                                                           const __grid_constant__ BasisTab tab) {
    constexpr int PITCH = 36;  // words per strip: 32 columns + 4 (16-byte aligned, conflict-free LDS.128)
    __shared__ __align__(16) uint32_t strip[256 * PITCH];
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int64_t mt = blockIdx.x;
    const int kc = blockIdx.y;
    const int64_t b = row0 + mt * TC_BM + r;
    const float scale = phi_scale(__uint_as_float(sc->max_xl), tab, N);
    unsigned char* base = tiles + ((mt * nkc + kc) * 2) * (int64_t)TC_A_IMG;
    uint32_t* mine = strip + threadIdx.x * PITCH;
    uint4* mine4 = reinterpret_cast<uint4*>(mine);
#pragma unroll
    for (int q = 0; q < 8; ++q) mine4[q] = make_uint4(0u, 0u, 0u, 0u);
    const int c0 = kc * TC_BK + half * 32;  // first dense column of this thread's strip
    if (b < Bp) {
        // every input with a column in this strip (the first and the last may belong to the neighbours as well);
        // the loads of four inputs are issued together: the kernel is otherwise bound by their latency
        const int i_lo = c0 / Kin, i_hi = min(I - 1, (c0 + 31) / Kin);
        for (int i4 = i_lo; i4 <= i_hi; i4 += 4) {
            uint32_t sg4[4];
            float xl4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t at = (int64_t)min(i4 + u, i_hi) * Bp + b;
                sg4[u] = seg_t[at];
                xl4[u] = xl_t[at];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i4 + u;
                const uint32_t sg = sg4[u];
                if (i > i_hi || (sg & (HOL_SEG_INVALID | HOL_SEG_OUTLIER))) continue;
                float phi[N];
                lagrange_basis<N>(xl4[u], tab, phi);
                const int col = i * Kin + (int)(sg & HOL_SEG_MASK) * stride - c0;  // strip column of phi_0
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const float a = phi[j] * scale;
                    const __half ha = __float2half_rn(a);
                    __half la = __float2half_rn(a - __half2float(ha));
                    // |a| < 2^-25 rounds to hi = lo = 0: keep the sign and the smallest subnormal (split_store8)
                    if (fabsf(a) <= 2.9802323e-8f && a != 0.0f) la = __ushort_as_half((unsigned short)(a > 0.0f ? 0x0001 : 0x8001));
                    if ((unsigned)(col + j) < 32u)
                        mine[col + j] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(la) << 16);
                }
            }
        }
    }
    // strip word c = (hi | lo << 16) of column c0 + c  ->  hi image / lo image, 8 columns per 16-byte chunk
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 w0 = mine4[2 * q], w1 = mine4[2 * q + 1];
        uint4 hi, lo;
        hi.x = __byte_perm(w0.x, w0.y, 0x5410); lo.x = __byte_perm(w0.x, w0.y, 0x7632);
        hi.y = __byte_perm(w0.z, w0.w, 0x5410); lo.y = __byte_perm(w0.z, w0.w, 0x7632);
        hi.z = __byte_perm(w1.x, w1.y, 0x5410); lo.z = __byte_perm(w1.x, w1.y, 0x7632);
        hi.w = __byte_perm(w1.z, w1.w, 0x5410); lo.w = __byte_perm(w1.z, w1.w, 0x7632);
        const int kgl = half * 4 + q;
        *reinterpret_cast<uint4*>(base + kgl * (TC_BM * 16) + r * 16) = hi;
        *reinterpret_cast<uint4*>(base + TC_A_IMG + kgl * (TC_BM * 16) + r * 16) = lo;
    }
}

// dW A: Phi^T tiles [m-tile (128 rows k = i*Kin+v)][batch chunk (64 samples)]; one thread = one
// (tile row, 8 samples) -> one coalesced 16-byte store per image.  Row k of the tile is basis
// j = v - s*stride of the samples whose segment covers column v (zero otherwise), so each thread
// evaluates a single phi_j per active sample: C_j * prod_{m != j} (t - X_m).
template <int N>
__global__ void __launch_bounds__(256) pw_tc_expand_t_kernel(const float* __restrict__ xl_t,
                                                             const uint16_t* __restrict__ seg_t,
                                                             unsigned char* __restrict__ tiles, const TcScalars* sc,
                                                             int64_t Bp, int64_t row0, int I, int Kin, int nbc, int stride,
                                                             const __grid_constant__ BasisTab tab) {
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int mt = blockIdx.x, bc = blockIdx.y;
    const int R = mt * TC_BM + r;
    const int i = R / Kin, v = R % Kin;
    const float scale = phi_scale(__uint_as_float(sc->max_xl), tab, N);
    unsigned char* base = tiles + (((int64_t)mt * nbc + bc) * 2) * (int64_t)TC_A_IMG;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int bgl = half * 4 + q;
        const int64_t b0 = row0 + (int64_t)bc * TC_BK + bgl * 8;
        float val[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) val[e] = 0.0f;
        if (i < I && b0 < Bp) {  // Bp is a multiple of 32: the 8 samples are all inside or all outside
            const uint4 sraw = *reinterpret_cast<const uint4*>(seg_t + (int64_t)i * Bp + b0);
            const float4 x0 = *reinterpret_cast<const float4*>(xl_t + (int64_t)i * Bp + b0);
            const float4 x1 = *reinterpret_cast<const float4*>(xl_t + (int64_t)i * Bp + b0 + 4);
            const uint32_t sw[4] = {sraw.x, sraw.y, sraw.z, sraw.w};
            const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t sg = (sw[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                const int j = v - (int)(sg & HOL_SEG_MASK) * stride;
                if (!(sg & (HOL_SEG_INVALID | HOL_SEG_OUTLIER)) && j >= 0 && j < N) {
                    float prod = 1.0f, cj = 0.0f;
#pragma unroll
                    for (int m = 0; m < N; ++m) {
                        prod *= (m == j) ? 1.0f : (xs[e] - tab.X[m]);
                        cj = (m == j) ? tab.C[m] : cj;
                    }
                    val[e] = (N == 1) ? 1.0f : prod * cj;
                }
            }
        }
        uint4* hi = reinterpret_cast<uint4*>(base + bgl * (TC_BM * 16) + r * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + TC_A_IMG + bgl * (TC_BM * 16) + r * 16);
        split_store8(val, scale, hi, lo);
    }
}

// forward B: W tiles [n-tile (BN outputs)][k-chunk], k = i*Kin + v
__global__ void __launch_bounds__(256) pw_tc_pack_w_kernel(const float* __restrict__ w, unsigned char* __restrict__ tiles,
                                                           const TcScalars* sc, int I, int O, int nb, int Kin, int nkc,
                                                           int BN) {
    const int nt = blockIdx.x, kc = blockIdx.y;
    const float scale = pow2_scale_for(__uint_as_float(sc->max_w));
    const int img = BN * TC_BK * 2;
    unsigned char* base = tiles + ((int64_t)(nt * nkc + kc) * 2) * img;
    for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < BN * 8; idx += gridDim.z * blockDim.x) {
        const int row = idx % BN, kgl = idx / BN;
        const int o = nt * BN + row;
        const int k0 = kc * TC_BK + kgl * 8;
        int i = k0 / Kin, col = k0 % Kin;  // the 8 columns may run over into the next input(s)
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            v[e] = (o < O && i < I && col < nb) ? w[((int64_t)o * I + i) * nb + col] : 0.0f;
            if (++col == Kin) col = 0, ++i;
        }
        uint4* hi = reinterpret_cast<uint4*>(base + kgl * (BN * 16) + row * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + img + kgl * (BN * 16) + row * 16);
        split_store8(v, scale, hi, lo);
    }
}

// One pass over the weights in the forward: the mean of every link (o, i) over its nb nodes (one warp
// per link) and max |w|.  dX only needs the weights up to a constant per link -- sum_j phi'_j = 0 inside
// every segment -- and centred weights remove the cancellation that otherwise amplifies the rounding
// of T = gy W^T (same-sign weights: 1e-5 -> 1e-6); max |w| sets the power-of-two scale of the W tiles.
__global__ void __launch_bounds__(256) pw_tc_wstats_kernel(const float* __restrict__ w, float* __restrict__ wmean,
                                                           unsigned* max_w, int64_t links, int nb) {
    const int lane = threadIdx.x & 31;
    float amax = 0.0f;
    if (nb <= 32) {  // short links: one thread per link (the 32 links of a warp are one contiguous stretch of w)
        for (int64_t link = (int64_t)blockIdx.x * 256 + threadIdx.x; link < links; link += (int64_t)gridDim.x * 256) {
            float s = 0.0f;
            for (int v = 0; v < nb; ++v) {
                const float x = w[link * nb + v];
                s += x;
                const float a = fabsf(x);
                if (a < 3.0e38f) amax = fmaxf(amax, a);
            }
            if (wmean) wmean[link] = s / (float)nb;
        }
    } else {
        for (int64_t link = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); link < links; link += (int64_t)gridDim.x * 8) {
            float s = 0.0f;
            for (int v = lane; v < nb; v += 32) {
                const float x = w[link * nb + v];
                s += x;
                const float a = fabsf(x);
                if (a < 3.0e38f) amax = fmaxf(amax, a);
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            if (lane == 0 && wmean) wmean[link] = s / (float)nb;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, m));
    if (lane == 0 && amax > 0.0f) atomicMax(max_w, __float_as_uint(amax));
}

// dX B: (W - link mean)^T tiles [n-tile (BN rows k = i*Kin+v)][o-chunk (64 outputs)]
__global__ void __launch_bounds__(256) pw_tc_pack_wt_kernel(const float* __restrict__ w, const float* __restrict__ wmean,
                                                            unsigned char* __restrict__ tiles,
                                                            const TcScalars* sc, int I, int O, int nb, int Kin, int noc,
                                                            int BN) {
    const int nt = blockIdx.x, oc = blockIdx.y;
    const float scale = pow2_scale_for(__uint_as_float(sc->max_w));
    const int img = BN * TC_BK * 2;
    unsigned char* base = tiles + ((int64_t)(nt * noc + oc) * 2) * img;
    for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < BN * 8; idx += gridDim.z * blockDim.x) {
        const int row = idx % BN, ogl = idx / BN;
        const int R = nt * BN + row;
        const int i = R / Kin, col = R % Kin;
        const int o0 = oc * TC_BK + ogl * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
            v[e] = (o0 + e < O && i < I && col < nb)
                       ? w[((int64_t)(o0 + e) * I + i) * nb + col] - wmean[(int64_t)(o0 + e) * I + i]
                       : 0.0f;
        uint4* hi = reinterpret_cast<uint4*>(base + ogl * (BN * 16) + row * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + img + ogl * (BN * 16) + row * 16);
        split_store8(v, scale, hi, lo);
    }
}

// dX A: gy tiles [m-tile (128 samples)][o-chunk]
__global__ void __launch_bounds__(256) pw_tc_pack_gy_kernel(const float* __restrict__ gy, unsigned char* __restrict__ tiles,
                                                            const TcScalars* sc, int64_t B, int64_t row0, int O, int noc) {
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int64_t mt = blockIdx.x;
    const int oc = blockIdx.y;
    const int64_t b = row0 + mt * TC_BM + r;
    const float scale = pow2_scale_for(__uint_as_float(sc->max_gy));
    unsigned char* base = tiles + ((mt * noc + oc) * 2) * (int64_t)TC_A_IMG;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ogl = half * 4 + q;
        const int o0 = oc * TC_BK + ogl * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (b < B && o0 + e < O) ? gy[b * O + o0 + e] : 0.0f;
        uint4* hi = reinterpret_cast<uint4*>(base + ogl * (TC_BM * 16) + r * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + TC_A_IMG + ogl * (TC_BM * 16) + r * 16);
        split_store8(v, scale, hi, lo);
    }
}

// dW B: gy^T tiles [n-tile (BN outputs)][batch chunk (64 samples)]
__global__ void __launch_bounds__(256) pw_tc_pack_gyt_kernel(const float* __restrict__ gy, unsigned char* __restrict__ tiles,
                                                             const TcScalars* sc, int64_t B, int64_t row0, int O, int nbc,
                                                             int BN) {
    const int nt = blockIdx.x, bc = blockIdx.y;
    const float scale = pow2_scale_for(__uint_as_float(sc->max_gy));
    const int img = BN * TC_BK * 2;
    unsigned char* base = tiles + ((int64_t)(nt * nbc + bc) * 2) * img;
    for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < BN * 8; idx += gridDim.z * blockDim.x) {
        const int row = idx % BN, bgl = idx / BN;
        const int o = nt * BN + row;
        const int64_t b0 = row0 + (int64_t)bc * TC_BK + bgl * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (o < O && b0 + e < B) ? gy[(b0 + e) * O + o] : 0.0f;
        uint4* hi = reinterpret_cast<uint4*>(base + bgl * (BN * 16) + row * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + img + bgl * (BN * 16) + row * 16);
        split_store8(v, scale, hi, lo);
    }
}

