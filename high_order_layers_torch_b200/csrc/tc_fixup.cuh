// fp32 fix-up kernels of the tensor-core path for elements flagged HOL_SEG_OUTLIER.
//
// The split-fp16 operands carry basis values up to 2^17 (Phi scaled by at most 2^-3, tc_common.cuh).
// The reference defines the layer for any input -- outside [-1, 1] the end segments extrapolate and
// the polynomial grows like |t|^(n-1) -- and every sample is independent, so an input far outside
// the range must neither overflow the tiles nor force a scale that flushes the in-range samples of
// the batch to zero.  The prep kernels flag such elements (|xl| > tc_xl_limit(n): 15 for n = 5, 3.2
// for n = 8), the operand builders leave them out of the tiles, and the kernels below add their
// contribution in plain fp32 after the GEMM -- one warp owns a sample (y, dX) or an input (dW) and
// walks its outliers in a fixed order, so the results stay bit-reproducible.  With no outlier in the
// batch (TcScalars::n_outliers == 0, the normal case) each kernel returns at once.
#pragma once
#include "norm_rows.cuh"
#include "tc_common.cuh"

// run-time-n versions of lagrange_basis / lagrange_basis_derivative (common.cuh)
__device__ __forceinline__ void fix_basis(float t, const BasisTab& tab, int n, float (&phi)[HOL_MAX_N], bool derivative) {
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        float P = 1.0f, Ssum = 0.0f;
        for (int m = 0; m < n; ++m) {
            if (m == j) continue;
            const float d = t - tab.X[m];
            Ssum = fmaf(Ssum, d, P);
            P *= d;
        }
        phi[j] = (derivative ? Ssum : P) * tab.C[j];
    }
    if (n == 1) phi[0] = derivative ? 0.0f : 1.0f;
}

struct FixParams {
    const float* xl_t;
    const uint16_t* seg_t;
    const float* w;    // [O][I][nb]
    const float* gy;   // [B][O]      (dX, dW)
    float* out;        // y[B][O] (+=), dx[B][I] or dx_t[I][Bp] (=), dw[O][I][nb] (+=)
    const TcScalars* sc;
    int64_t B, Bp;
    int I, O, nb, stride, n, transposed;
    float length, Sf;
    BasisTab tab;
};

// y[b, :] += sum_j phi_j(xl) w[:, i, s*stride + j]   and   dx[b, i] = chain * sum_o gy[b,o] sum_j phi'_j w[o, i, ..]
// for every outlier (b, i): one warp per 32 consecutive samples, inputs in increasing order.
template <bool DX>
__global__ void __launch_bounds__(256) pw_tc_fix_rows_kernel(const __grid_constant__ FixParams p) {
    if (p.sc->n_outliers == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t b0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    if (b0 >= p.B) return;
    const int64_t b = b0 + lane;
    for (int i = 0; i < p.I; ++i) {
        const uint32_t sg = (b < p.B) ? p.seg_t[(int64_t)i * p.Bp + b] : 0u;
        unsigned mask = __ballot_sync(0xffffffffu, (sg & HOL_SEG_OUTLIER) != 0);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const int64_t bb = b0 + l;
            const uint32_t sgl = __shfl_sync(0xffffffffu, sg, l);
            const float xl = p.xl_t[(int64_t)i * p.Bp + bb];
            const int s = (int)(sgl & HOL_SEG_MASK);
            float phi[HOL_MAX_N];
            fix_basis(xl, p.tab, p.n, phi, DX);
            const int col = s * p.stride;
            if (!DX) {
                for (int o = lane; o < p.O; o += 32) {
                    const float* wl = p.w + ((int64_t)o * p.I + i) * p.nb + col;
                    float acc = 0.0f;
                    for (int j = 0; j < p.n; ++j) acc = fmaf(phi[j], wl[j], acc);
                    p.out[bb * p.O + o] += acc;
                }
            } else {
                float acc = 0.0f;
                for (int o = lane; o < p.O; o += 32) {
                    const float* wl = p.w + ((int64_t)o * p.I + i) * p.nb + col;
                    float t = 0.0f;
                    for (int j = 0; j < p.n; ++j) t = fmaf(phi[j], wl[j], t);
                    acc = fmaf(p.gy[bb * p.O + o], t, acc);
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                if (lane == 0) {
                    const float e0 = __fsub_rn(__fmul_rn(__fdiv_rn((float)s, p.Sf), 2.0f), 1.0f);
                    const float e1 = __fsub_rn(__fmul_rn(__fdiv_rn((float)(s + 1), p.Sf), 2.0f), 1.0f);
                    float chain = __fdiv_rn(p.length, __fsub_rn(e1, e0));
                    if (sgl & HOL_SEG_MIRROR) chain = -chain;
                    if (p.transposed)
                        p.out[(int64_t)i * p.Bp + bb] = acc * chain;
                    else
                        p.out[bb * p.I + i] = acc * chain;
                }
            }
        }
    }
}

// dw[:, i, s*stride + j] += phi_j(xl) gy[b, :] for every outlier (b, i): one warp per input, samples in
// increasing order, lanes over the outputs.
__global__ void __launch_bounds__(256) pw_tc_fix_dw_kernel(const __grid_constant__ FixParams p) {
    if (p.sc->n_outliers == 0) return;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= p.I) return;
    for (int64_t b0 = 0; b0 < p.B; b0 += 32) {
        const int64_t b = b0 + lane;
        const uint32_t sg = (b < p.B) ? p.seg_t[(int64_t)i * p.Bp + b] : 0u;
        unsigned mask = __ballot_sync(0xffffffffu, (sg & HOL_SEG_OUTLIER) != 0);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const int64_t bb = b0 + l;
            const uint32_t sgl = __shfl_sync(0xffffffffu, sg, l);
            const float xl = p.xl_t[(int64_t)i * p.Bp + bb];
            float phi[HOL_MAX_N];
            fix_basis(xl, p.tab, p.n, phi, false);
            const int col = (int)(sgl & HOL_SEG_MASK) * p.stride;
            for (int o = lane; o < p.O; o += 32) {
                const float g = p.gy[bb * p.O + o];
                float* dl = p.out + ((int64_t)o * p.I + i) * p.nb + col;
                for (int j = 0; j < p.n; ++j) dl[j] = fmaf(phi[j], g, dl[j]);
            }
        }
    }
}

// Last kernel of a tensor-core forward: ONE pass per output row (one warp per row) that
//   (1) sums the partial results of a forward split over K (ks > 1; fixed order),
//   (2) adds the fp32 contributions of the row's outlier elements (skipped when the batch has none),
//   (3) applies the per-sample normalisation that follows the layer in the network, if any
//       (max_abs / max_center / l2 as the forward's epilogue): the raw row goes to y_raw
//       (the normalisation's backward needs it), the normalised row to y_norm, its statistics to stats.
// It replaces the separate reduction, fix-up and normalisation launches (three passes over [B, O]).
struct FinalizeParams {
    FixParams fix;        // fix.out = y_raw [B][O]
    const float* parts;   // [ks][B][O] partial results, or nullptr: y_raw already holds the GEMM's result
    float* y_norm;        // nullptr: no normalisation
    float4* stats;
    int ks, norm_kind;
    float eps;
};

__global__ void __launch_bounds__(256) pw_tc_finalize_kernel(const __grid_constant__ FinalizeParams q) {
    const FixParams& p = q.fix;
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= p.B) return;
    float* yr = p.out + b * p.O;
    if (q.parts) {
        const int64_t stride = p.B * (int64_t)p.O;
        for (int o = lane; o < p.O; o += 32) {
            float acc = 0.0f;
            for (int k = 0; k < q.ks; ++k) acc += q.parts[(int64_t)k * stride + b * p.O + o];
            yr[o] = acc;
        }
    }
    if (p.sc->n_outliers != 0) {
        // lanes scan the row's inputs 32 at a time; an outlier is then handled by the whole warp over the outputs
        for (int i0 = 0; i0 < p.I; i0 += 32) {
            const int i = i0 + lane;
            const uint32_t sg = (i < p.I) ? p.seg_t[(int64_t)i * p.Bp + b] : 0u;
            unsigned mask = __ballot_sync(0xffffffffu, (sg & HOL_SEG_OUTLIER) != 0);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const int ii = i0 + l;
                const uint32_t sgl = __shfl_sync(0xffffffffu, sg, l);
                const float xl = p.xl_t[(int64_t)ii * p.Bp + b];
                float phi[HOL_MAX_N];
                fix_basis(xl, p.tab, p.n, phi, false);
                const int col = (int)(sgl & HOL_SEG_MASK) * p.stride;
                __syncwarp();
                for (int o = lane; o < p.O; o += 32) {
                    const float* wl = p.w + ((int64_t)o * p.I + ii) * p.nb + col;
                    float acc = 0.0f;
                    for (int j = 0; j < p.n; ++j) acc = fmaf(phi[j], wl[j], acc);
                    yr[o] += acc;
                }
            }
        }
    }
    if (q.y_norm) {
        __syncwarp();  // the row was written by this warp's lanes: visible to all of them from here on
        float* nr = q.y_norm + b * p.O;
        if (q.norm_kind == HOL_NORM_MAX_ABS) norm_row_forward<HOL_NORM_MAX_ABS>(yr, nr, q.stats + b, p.O, q.eps, lane);
        else if (q.norm_kind == HOL_NORM_MAX_CENTER) norm_row_forward<HOL_NORM_MAX_CENTER>(yr, nr, q.stats + b, p.O, q.eps, lane);
        else norm_row_forward<HOL_NORM_L2>(yr, nr, q.stats + b, p.O, q.eps, lane);
    }
}
