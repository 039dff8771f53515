// HBM-bound helper stages: local-coordinate prep (bit-exact segment index), weight packing,
// stable per-input bucket sort for dW, col2im for the conv input gradient.
#include <math.h>

#include "kernels.cuh"

// ---------------------------------------------------------------------------------------
// the reference's elementwise recipe, op for op (each fp32 op rounded separately, no FMA
// contraction: the *_rn intrinsics are never fused)
//   make_periodic ........ utils.py:74-79
//   segment index ........ PolynomialLayers.py:299-303
//   _eta, x_in ........... PolynomialLayers.py:329-344
// ---------------------------------------------------------------------------------------
struct PrepConst {
    float length, half, Sf, half_per, two_per, per;
    float xl_limit;    // tensor-core plans: |xl| above this is flagged HOL_SEG_OUTLIER (INFINITY otherwise)
    float inv_length;  // 1 / length when that is exact (length a power of two: 2.0 for every FC layer), else 0
    float inv_de;      // 1 / (eta(s+1) - eta(s)) when every segment has the same power-of-two width (S a power of
                       // two: eta is exact then), else 0: dividing by a power of two is exactly a multiplication
    int S, periodic;
};

static PrepConst make_prep_const(const PwShape& s) {
    PrepConst c;
    c.length = s.length;
    c.half = s.half;
    c.Sf = (float)s.S;
    c.half_per = s.half_per;
    c.two_per = s.two_per;
    c.per = s.periodicity;
    c.xl_limit = s.tc.enabled ? s.tc.xl_limit : INFINITY;
    {
        int e;
        const float m = frexpf(s.length, &e);  // a power of two divides exactly like its reciprocal multiplies
        c.inv_length = (m == 0.5f && e > -100 && e < 100) ? 1.0f / s.length : 0.0f;
    }
    c.inv_de = 0.0f;
    {
        // eta(k) with the reference's three separately rounded fp32 operations (volatile: no host-side contraction)
        auto eta_host = [&](int k) {
            volatile float q = (float)k / (float)s.S;
            volatile float m = q * 2.0f;
            volatile float e = m - 1.0f;
            return (float)e;
        };
        volatile float de0 = eta_host(1) - eta_host(0);
        bool same = true;
        for (int k = 1; k < s.S && same; ++k) {
            volatile float de = eta_host(k + 1) - eta_host(k);
            same = de == de0;
        }
        int e;
        const float m = frexpf(de0, &e);
        if (same && m == 0.5f && e > -100 && e < 100) c.inv_de = 1.0f / de0;
    }
    c.S = s.S;
    c.periodic = s.periodic;
    return c;
}

__device__ __forceinline__ float fold_periodic(float xv, const PrepConst& c, uint32_t& flags) {
    float xp = __fadd_rn(xv, c.half_per);
    float r = fmodf(xp, c.two_per);  // exact; torch.remainder = fmod + sign fix-up
    if (r != 0.0f && ((r < 0.0f) != (c.two_per < 0.0f))) r = __fadd_rn(r, c.two_per);
    if (r > c.per) {
        r = __fsub_rn(c.two_per, r);
        flags |= HOL_SEG_MIRROR;
    }
    return __fsub_rn(r, c.half_per);
}

__device__ __forceinline__ int segment_of(float xv, const PrepConst& c) {
    const float sh = __fadd_rn(xv, c.half);
    float v = __fmul_rn(c.inv_length != 0.0f ? __fmul_rn(sh, c.inv_length) : __fdiv_rn(sh, c.length), c.Sf);
    // .long() truncates toward zero, clamp(0, S-1) follows; NaN -> 0 like the reference on CPU
    return (v >= c.Sf) ? c.S - 1 : (v > 0.0f ? (int)v : 0);
}

// eta(k) = k / float(S) * 2 - 1, the reference's three separately rounded ops (PolynomialLayers.py:338-344)
__device__ __forceinline__ float eta_of(int k, const PrepConst& c) {
    return __fsub_rn(__fmul_rn(__fdiv_rn((float)k, c.Sf), 2.0f), 1.0f);
}

// `eta`: optional table eta[0..S] in shared memory (the same values, computed once per CTA instead of
// two divisions per element)
__device__ __forceinline__ void prep_elem(float xv, const PrepConst& c, float& xl, uint32_t& sg, const float* eta = nullptr) {
    uint32_t flags = 0;
    if (c.periodic) xv = fold_periodic(xv, c, flags);
    int s = segment_of(xv, c);
    float e0 = eta ? eta[s] : eta_of(s, c);
    float q;
    if (c.inv_de != 0.0f) {
        q = __fmul_rn(__fsub_rn(xv, e0), c.inv_de);
    } else {
        float e1 = eta ? eta[s + 1] : eta_of(s + 1, c);
        q = __fdiv_rn(__fsub_rn(xv, e0), __fsub_rn(e1, e0));
    }
    xl = __fsub_rn(__fmul_rn(c.length, q), c.half);
    if (fabsf(xl) > c.xl_limit) flags |= HOL_SEG_OUTLIER;  // NaN is not an outlier: it propagates like in the reference
    sg = (uint32_t)s | flags;
}

// Statistics the tensor-core passes need, gathered while the coordinates are in registers: the
// largest |xl| among the elements that stay in the GEMM, and how many were flagged as outliers.
// stats == nullptr on layers without a tensor-core pass.
__device__ __forceinline__ void prep_stats(float amax, unsigned nout, unsigned* stats) {
    if (!stats) return;  // uniform over the grid
    __shared__ unsigned s_max[8], s_out[8];  // CTAs of 256 threads; every thread of the CTA gets here
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
        nout += __shfl_xor_sync(0xffffffffu, nout, d);
    }
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) {
        s_max[tid >> 5] = amax > 0.0f ? __float_as_uint(amax) : 0u;
        s_out[tid >> 5] = nout;
    }
    __syncthreads();
    if (tid == 0) {
        // One atomic per CTA, and only from a CTA that raises the value: every warp of the grid hitting ONE
        // address was 0.5 ms of L2 serialisation at config 2 (half a million atomics), and still ~8 us on a
        // single-wave grid of a small layer, where looking first does not help (all look at once).
        unsigned bits = 0u, n = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) bits = max(bits, s_max[k]), n += s_out[k];
        if (bits > *reinterpret_cast<volatile unsigned*>(stats + 0)) atomicMax(stats + 0, bits);  // TcScalars::max_xl
        if (n) atomicAdd(stats + 3, n);                                                          // TcScalars::n_outliers
    }
}
__device__ __forceinline__ void prep_track(float xl, uint32_t sg, float& amax, unsigned& nout) {
    if (sg & HOL_SEG_OUTLIER) ++nout;
    else if (!(sg & HOL_SEG_INVALID)) {
        const float a = fabsf(xl);
        if (a < 3.0e38f) amax = fmaxf(amax, a);
    }
}

__global__ void __launch_bounds__(256) pw_segment_index_kernel(const float* __restrict__ x, int64_t* __restrict__ seg,
                                                               int64_t count, PrepConst c) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; k < count; k += step) {
        float xv = x[k];
        uint32_t flags = 0;
        if (c.periodic) xv = fold_periodic(xv, c, flags);
        seg[k] = (int64_t)segment_of(xv, c);
    }
}

// x[B, I] -> xl_t[I][Bp], seg_t[I][Bp] (transposed so that the contraction kernels, whose lanes are
// samples, read them coalesced).  64 x 64 tile through shared memory: 16 independent loads per thread, 256-byte
// bursts both ways (the first version moved 32 x 32 tiles with 4 elements per thread and ran at 1.5 TB/s).
#define PREP_T 64
__global__ void __launch_bounds__(256) pw_prep_fc_kernel(const float* __restrict__ x, float* __restrict__ xl_t,
                                                         uint16_t* __restrict__ seg_t, int64_t B, int64_t Bp, int I,
                                                         PrepConst c, unsigned* stats) {
    __shared__ float t_xl[PREP_T][PREP_T + 1];                     // [input][sample]
    __shared__ __align__(4) uint16_t t_sg[PREP_T][PREP_T + 2];
    extern __shared__ float eta_tab[];  // [S + 1]
    const int tid = threadIdx.x;
    for (int k = tid; k <= c.S; k += 256) eta_tab[k] = eta_of(k, c);
    const int64_t b0 = (int64_t)blockIdx.x * PREP_T;
    const int i0 = blockIdx.y * PREP_T;
    {
        const int ci = tid & (PREP_T - 1), rb = tid >> 6;  // 4 sample rows at a time, 64 consecutive inputs each
        const int i = i0 + ci;
        float xv[PREP_T / 4];
#pragma unroll
        for (int r = 0; r < PREP_T / 4; ++r) {
            const int64_t b = b0 + rb + 4 * r;
            xv[r] = (i < I && b < B) ? x[b * I + i] : 0.0f;
        }
        __syncthreads();  // eta table
        float amax = 0.0f;
        unsigned nout = 0;
#pragma unroll
        for (int r = 0; r < PREP_T / 4; ++r) {
            const int64_t b = b0 + rb + 4 * r;
            float xl = 0.0f;
            uint32_t sg = HOL_SEG_INVALID;
            if (i < I && b < B) prep_elem(xv[r], c, xl, sg, eta_tab);
            prep_track(xl, sg, amax, nout);
            t_xl[ci][rb + 4 * r] = xl;
            t_sg[ci][rb + 4 * r] = (uint16_t)sg;
        }
        prep_stats(amax, nout, stats);  // (contains a __syncthreads when stats != nullptr)
    }
    __syncthreads();
    {
        const int cb = tid & (PREP_T - 1), ri = tid >> 6;
        if (b0 + cb < Bp) {
#pragma unroll
            for (int r = 0; r < PREP_T / 4; ++r) {
                const int il = ri + 4 * r;
                if (i0 + il < I) xl_t[(int64_t)(i0 + il) * Bp + b0 + cb] = t_xl[il][cb];
            }
        }
        const int pb = tid & 31, rj = tid >> 5;  // two samples (one 32-bit word) per thread
        if (b0 + 2 * pb < Bp) {                  // Bp is even: the pair is inside or outside as a whole
#pragma unroll
            for (int r = 0; r < PREP_T / 8; ++r) {
                const int il = rj + 8 * r;
                if (i0 + il < I)
                    *reinterpret_cast<uint32_t*>(seg_t + (int64_t)(i0 + il) * Bp + b0 + 2 * pb) =
                        *reinterpret_cast<const uint32_t*>(&t_sg[il][2 * pb]);
            }
        }
    }
}

// Implicit im2col: row = (b, ho, wo), input i = (c, kh, kw).  Taps that fall into the zero
// padding are flagged invalid (the reference pads the expanded tensor, so they add nothing).
struct ConvGeom {  // rectangular geometry: 1-d convolutions are H = Kh = 1
    int C, H, W, Kh, Kw, sh, sw, ph, pw, dh, dw, Ho, Wo;
};

// Stage 1: the elementwise recipe once per pixel, image layout kept (NCHW).
__global__ void __launch_bounds__(256) pw_prep_image_kernel(const float* __restrict__ x, float* __restrict__ xl_img,
                                                            uint16_t* __restrict__ seg_img, int64_t count, PrepConst c,
                                                            unsigned* stats) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float xl = 0.0f, amax = 0.0f;
    uint32_t sg = HOL_SEG_INVALID;
    unsigned nout = 0;
    if (e < count) {
        prep_elem(x[e], c, xl, sg);
        xl_img[e] = xl;
        seg_img[e] = (uint16_t)sg;
    }
    prep_track(xl, sg, amax, nout);
    prep_stats(amax, nout, stats);
}

// Stage 2: unfold the prepared image into xl_t / seg_t [I][Bp].  One thread = one row (b, ho, wo),
// decoded once; the loop over the taps writes one coalesced line per tap and re-reads the
// window from L1.
__global__ void __launch_bounds__(256) pw_unfold_kernel(const float* __restrict__ xl_img, const uint16_t* __restrict__ seg_img,
                                                        float* __restrict__ xl_t, uint16_t* __restrict__ seg_t, int64_t rows,
                                                        int64_t Bp, ConvGeom g) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= Bp) return;
    const bool live = row < rows;
    const int wo = (int)(row % g.Wo);
    const int64_t t = row / g.Wo;
    const int ho = (int)(t % g.Ho);
    const int64_t b = t / g.Ho;
    const int h0 = ho * g.sh - g.ph, w0 = wo * g.sw - g.pw;
    int i = 0;
    for (int ch = 0; ch < g.C; ++ch) {
        const int64_t plane = (b * g.C + ch) * (int64_t)g.H * g.W;
        for (int kh = 0; kh < g.Kh; ++kh) {
            const int h = h0 + kh * g.dh;
            const bool hok = live && h >= 0 && h < g.H;
#pragma unroll 5
            for (int kw = 0; kw < g.Kw; ++kw, ++i) {
                const int w = w0 + kw * g.dw;
                float xl = 0.0f;
                uint16_t sg = (uint16_t)HOL_SEG_INVALID;
                if (hok && w >= 0 && w < g.W) {
                    const int64_t src = plane + (int64_t)h * g.W + w;
                    xl = __ldg(xl_img + src);
                    sg = __ldg(seg_img + src);
                }
                xl_t[(int64_t)i * Bp + row] = xl;
                seg_t[(int64_t)i * Bp + row] = sg;
            }
        }
    }
}

// dx[b,c,h,w] = sum over the (kh,kw) taps that read this pixel of dx_t[(c,kh,kw)][row].
__global__ void __launch_bounds__(256) pw_col2im_kernel(const float* __restrict__ dx_t, float* __restrict__ dx,
                                                        int64_t total, int64_t Bp, ConvGeom g) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int w = (int)(e % g.W);
    int64_t t = e / g.W;
    const int h = (int)(t % g.H);
    t /= g.H;
    const int ch = (int)(t % g.C);
    const int64_t b = t / g.C;
    float acc = 0.0f;
    for (int kh = 0; kh < g.Kh; ++kh) {
        const int hn = h + g.ph - kh * g.dh;
        if (hn < 0 || hn % g.sh) continue;
        const int ho = hn / g.sh;
        if (ho >= g.Ho) continue;
        for (int kw = 0; kw < g.Kw; ++kw) {
            const int wn = w + g.pw - kw * g.dw;
            if (wn < 0 || wn % g.sw) continue;
            const int wo = wn / g.sw;
            if (wo >= g.Wo) continue;
            const int i = (ch * g.Kh + kh) * g.Kw + kw;
            acc += dx_t[(int64_t)i * Bp + ((b * g.Ho + ho) * g.Wo + wo)];
        }
    }
    dx[e] = acc;
}

// ---------------------------------------------------------------------------------------
// conv-transpose (FunctionalConvolutionTranspose.py:75-160): the expansion followed by
// nn.ConvTranspose2d equals the piecewise layer applied per input pixel (rows = B*H*W, inputs = C,
// outputs = O*K*K) followed by a fold of the K*K taps onto the output grid
//   y[b, o, h*s - p + kh*d, w*s - p + kw*d] += z[(b, h, w)][o*K*K + kh*K + kw].
// fold gathers (one thread per output element, fixed order: no atomics), unfold is its adjoint.
// ---------------------------------------------------------------------------------------
struct FoldGeom {
    int O, H, W, Ho, Wo, K, stride, padding, dilation;
};

__global__ void __launch_bounds__(256) pw_fold_rows_kernel(const float* __restrict__ z, float* __restrict__ y, int64_t total,
                                                           FoldGeom g) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int wo = (int)(e % g.Wo);
    int64_t t = e / g.Wo;
    const int ho = (int)(t % g.Ho);
    t /= g.Ho;
    const int o = (int)(t % g.O);
    const int64_t b = t / g.O;
    const int64_t ld = (int64_t)g.O * g.K * g.K;
    float acc = 0.0f;
    for (int kh = 0; kh < g.K; ++kh) {
        const int hn = ho + g.padding - kh * g.dilation;
        if (hn < 0 || hn % g.stride) continue;
        const int h = hn / g.stride;
        if (h >= g.H) continue;
        for (int kw = 0; kw < g.K; ++kw) {
            const int wn = wo + g.padding - kw * g.dilation;
            if (wn < 0 || wn % g.stride) continue;
            const int w = wn / g.stride;
            if (w >= g.W) continue;
            acc += z[((b * g.H + h) * g.W + w) * ld + (o * g.K + kh) * g.K + kw];
        }
    }
    y[e] = acc;
}

__global__ void __launch_bounds__(256) pw_unfold_rows_kernel(const float* __restrict__ gy, float* __restrict__ gz,
                                                             int64_t total, FoldGeom g) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [rows][O*K*K]
    if (e >= total) return;
    const int kw = (int)(e % g.K);
    int64_t t = e / g.K;
    const int kh = (int)(t % g.K);
    t /= g.K;
    const int o = (int)(t % g.O);
    t /= g.O;
    const int w = (int)(t % g.W);
    t /= g.W;
    const int h = (int)(t % g.H);
    const int64_t b = t / g.H;
    const int ho = h * g.stride - g.padding + kh * g.dilation, wo = w * g.stride - g.padding + kw * g.dilation;
    float v = 0.0f;
    if (ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo) v = gy[((b * g.O + o) * g.Ho + ho) * (int64_t)g.Wo + wo];
    gz[e] = v;
}

extern "C" int hol_fold_rows_f32(const float* z_rows, float* y, int64_t Bimg, int32_t O, int32_t H, int32_t W, int32_t Ho,
                                 int32_t Wo, int32_t K, int32_t stride, int32_t padding, int32_t dilation, void* stream) {
    if (Bimg < 0 || O <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || K <= 0 || stride <= 0 || dilation <= 0 || padding < 0)
        return HOL_EINVAL;
    const int64_t total = Bimg * O * Ho * Wo;
    if (total == 0) return HOL_OK;
    if (!z_rows || !y) return HOL_EINVAL;
    FoldGeom g{O, H, W, Ho, Wo, K, stride, padding, dilation};
    pw_fold_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z_rows, y, total, g);
    return (int)cudaGetLastError();
}

extern "C" int hol_unfold_rows_f32(const float* gy, float* gz_rows, int64_t Bimg, int32_t O, int32_t H, int32_t W,
                                   int32_t Ho, int32_t Wo, int32_t K, int32_t stride, int32_t padding, int32_t dilation,
                                   void* stream) {
    if (Bimg < 0 || O <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || K <= 0 || stride <= 0 || dilation <= 0 || padding < 0)
        return HOL_EINVAL;
    const int64_t total = Bimg * H * W * O * K * K;
    if (total == 0) return HOL_OK;
    if (!gy || !gz_rows) return HOL_EINVAL;
    FoldGeom g{O, H, W, Ho, Wo, K, stride, padding, dilation};
    pw_unfold_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gy, gz_rows, total, g);
    return (int)cudaGetLastError();
}

// w[O][I][nb] -> wk[nOT][I][R][OTP].  Packed row r = (c, a) = (r / RS, r % RS) holds reference
// column v = a*stride + c; for continuous layers the shared end node of segment a-1 / a is the
// single row (0, a) and rows (c >= 1, a == S) are padding.  Pad rows/columns are zero.
__global__ void __launch_bounds__(256) pw_pack_kernel(const float* __restrict__ w, float* __restrict__ wk, int I, int O,
                                                      int nb, int stride, int S, int RS, int R, int OT, int OTP) {
    __shared__ float tile[64][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x, ot = blockIdx.y, r0 = blockIdx.z * 32;
    const int o0 = ot * OT;
    {
        const int r = r0 + tx;
        const int cg = r / RS, a = r % RS;
        const int v = a * stride + cg;
        const bool ok = r < R && v < nb && (a < S || cg == 0);
        for (int ol = ty; ol < OT; ol += 8) {
            float val = 0.0f;
            if (ok && o0 + ol < O) val = w[((int64_t)(o0 + ol) * I + i) * nb + v];
            tile[ol][tx] = val;
        }
    }
    __syncthreads();
    for (int rl = ty; rl < 32; rl += 8) {
        const int r = r0 + rl;
        if (r >= R) continue;
        float* dst = wk + (((int64_t)ot * I + i) * R + r) * OTP;
        for (int ol = tx; ol < OTP; ol += 32) dst[ol] = ol < OT ? tile[ol][rl] : 0.0f;
    }
}

// Stable counting sort of the samples of one input by segment, chunk by chunk (one CTA per
// (input, chunk of HOL_SORT_CHUNK samples)).  The order inside a bucket is the sample order, so
// the dW reduction order is fixed run to run.  order[i][b] holds sample ids, offs[i][chunk][s]
// the start of bucket s inside the chunk (absolute position in order[i]).
__global__ void __launch_bounds__(32 * HOL_WARPS_SORT)
    pw_bucket_sort_kernel(const uint16_t* __restrict__ seg_t, uint32_t* __restrict__ order, uint32_t* __restrict__ offs,
                          int64_t B, int64_t Bp, int Seff, int nchunks, int single_bucket) {
    extern __shared__ uint32_t cnt[];  // [HOL_WARPS_SORT][Seff]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = blockIdx.y;
    const int chunk = blockIdx.x;
    const int64_t c_lo = (int64_t)chunk * HOL_SORT_CHUNK;
    const int64_t c_hi = c_lo + HOL_SORT_CHUNK < B ? c_lo + HOL_SORT_CHUNK : B;
    const uint16_t* sg = seg_t + (int64_t)i * Bp;
    uint32_t* ord = order + (int64_t)i * Bp;
    uint32_t* myoffs = offs + ((int64_t)i * nchunks + chunk) * (Seff + 1);
    for (int k = tid; k < HOL_WARPS_SORT * Seff; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    const int64_t ngroups = (c_hi - c_lo + 31) / 32;
    const int64_t gper = (ngroups + HOL_WARPS_SORT - 1) / HOL_WARPS_SORT;
    const int64_t g_lo = warp * gper;
    const int64_t g_hi = g_lo + gper < ngroups ? g_lo + gper : ngroups;
    uint32_t* mycnt = cnt + warp * Seff;
    const uint32_t lt = (1u << lane) - 1u;
    for (int64_t gidx = g_lo; gidx < g_hi; ++gidx) {
        const int64_t b = c_lo + gidx * 32 + lane;
        const uint32_t v = b < c_hi ? sg[b] : HOL_SEG_INVALID;
        const bool valid = !(v & HOL_SEG_INVALID);
        const uint32_t key = single_bucket ? 0u : (v & HOL_SEG_MASK);
        const uint32_t mask = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
        if (valid && (mask & lt) == 0) mycnt[key] += __popc(mask);
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t carry = (uint32_t)c_lo;
        for (int s0 = 0; s0 < Seff; s0 += 32) {
            const int s = s0 + lane;
            uint32_t tot = 0;
            if (s < Seff)
                for (int w = 0; w < HOL_WARPS_SORT; ++w) tot += cnt[w * Seff + s];
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (s < Seff) {
                uint32_t off = carry + incl - tot;
                myoffs[s] = off;
                for (int w = 0; w < HOL_WARPS_SORT; ++w) {
                    const uint32_t t = cnt[w * Seff + s];
                    cnt[w * Seff + s] = off;
                    off += t;
                }
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) myoffs[Seff] = carry;
    }
    __syncthreads();
    for (int64_t gidx = g_lo; gidx < g_hi; ++gidx) {
        const int64_t b = c_lo + gidx * 32 + lane;
        const uint32_t v = b < c_hi ? sg[b] : HOL_SEG_INVALID;
        const bool valid = !(v & HOL_SEG_INVALID);
        const uint32_t key = single_bucket ? 0u : (v & HOL_SEG_MASK);
        const uint32_t mask = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
        const uint32_t cur = valid ? mycnt[key] : 0u;
        __syncwarp();
        if (valid && (mask & lt) == 0) mycnt[key] = cur + __popc(mask);
        __syncwarp();
        if (valid) ord[cur + __popc(mask & lt)] = (uint32_t)b;
    }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
int launch_segment_index(const float* x, int64_t* seg, int64_t count, const PwShape& s, cudaStream_t st) {
    if (count == 0) return HOL_OK;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pw_segment_index_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, seg, count, make_prep_const(s));
    return (int)cudaGetLastError();
}

int launch_prep_fc(const float* x, const PwShape& s, const PwWorkspace& ws, cudaStream_t st) {
    dim3 grid((unsigned)((s.Bp + PREP_T - 1) / PREP_T), (unsigned)((s.I + PREP_T - 1) / PREP_T));
    pw_prep_fc_kernel<<<grid, 256, (size_t)(s.S + 1) * sizeof(float), st>>>(
        x, ws.xl_t, ws.seg_t, s.B, s.Bp, s.I, make_prep_const(s), (unsigned*)ws.tc_scalars);
    return (int)cudaGetLastError();
}

static ConvGeom make_geom(const HolConvGeom& c) {
    return ConvGeom{c.C, c.H, c.W, c.Kh, c.Kw, c.sh, c.sw, c.ph, c.pw, c.dh, c.dw, c.Ho, c.Wo};
}

int launch_prep_conv(const float* x, const PwShape& s, const PwWorkspace& ws, int64_t Bimg, const HolConvGeom& cg,
                     cudaStream_t st) {
    const ConvGeom g = make_geom(cg);
    const int64_t count = Bimg * g.C * g.H * g.W;
    if (count > 0) {
        pw_prep_image_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(x, ws.xl_img, ws.seg_img, count,
                                                                              make_prep_const(s), (unsigned*)ws.tc_scalars);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    pw_unfold_kernel<<<(unsigned)((s.Bp + 255) / 256), 256, 0, st>>>(ws.xl_img, ws.seg_img, ws.xl_t, ws.seg_t, s.B, s.Bp, g);
    return (int)cudaGetLastError();
}

int launch_col2im(const float* dx_t, float* dx, int64_t Bimg, int64_t Bp, const HolConvGeom& cg, cudaStream_t st) {
    const ConvGeom g = make_geom(cg);
    const int64_t total = Bimg * g.C * g.H * g.W;
    if (total == 0) return HOL_OK;
    pw_col2im_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dx_t, dx, total, Bp, g);
    return (int)cudaGetLastError();
}

int launch_pack(const float* w, const PwShape& s, const PwWorkspace& ws, cudaStream_t st) {
    dim3 grid((unsigned)s.I, (unsigned)s.nOT, (unsigned)((s.R + 31) / 32));
    pw_pack_kernel<<<grid, dim3(32, 8), 0, st>>>(w, ws.wk, s.I, s.O, s.nb, s.stride, s.S, s.RS, s.R, s.OT, s.OTP);
    return (int)cudaGetLastError();
}

int launch_bucket_sort(const PwShape& s, const PwWorkspace& ws, cudaStream_t st) {
    const size_t smem = (size_t)HOL_WARPS_SORT * s.Seff * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        static std::atomic<uint64_t> attr_done{0};
        cudaError_t e = hol_ensure_max_smem(pw_bucket_sort_kernel, attr_done);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)s.nchunks, (unsigned)s.I);
    pw_bucket_sort_kernel<<<grid, 32 * HOL_WARPS_SORT, smem, st>>>(ws.seg_t, ws.order, ws.offs, s.B, s.Bp, s.Seff, s.nchunks,
                                                                   (!s.disc && s.n == 1) ? 1 : 0);
    return (int)cudaGetLastError();
}
