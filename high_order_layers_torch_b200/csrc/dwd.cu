// Weight gradient of NARROW layers (conv: O = 6, 16; classifier heads: O = 10; few segments):
//   dW[o,i,v] = sum_b coef_v(b,i) * gy[b,o],   coef_v = phi_j(xl(b,i)) if v == s(b,i)*stride + j else 0.
//
// With only nb*O <= ~128 gradient entries per input the whole [nb][OC] block of one input lives
// in the registers of every thread: lanes are consecutive samples (xl_t / seg_t are read fully
// coalesced, no bucket sort, no gather of gy rows), the one-hot over the segments is expanded
// into the nb coefficients with S predicated multiplies, and the block is updated with nb*OC/2
// packed FMAs per (sample, input).  That spends nb/n times the useful flops -- on the FMA pipe,
// which is idle in the bucketed kernel (dw.cu) for such layers: there a warp gathers one
// 24-64 byte gy row per sample for n*4 FMAs.  One CTA = (input, output group, sample range);
// the partial sums are combined in a fixed order (warp butterfly, 8 warps in order, then the
// sample ranges in order by pw_dw_reduce_kernel), so the result is bit-reproducible and weights
// no sample touched are exact zeros.
#include <stdlib.h>

#include "kernels.cuh"

struct DwdParams {
    const float* xl_t;
    const uint16_t* seg_t;
    const float* gy;
    float* out;  // dw, or the partial buffer [ksplit][O][I][nb]
    int64_t B, Bp, rows_per_split, part_stride;
    int I, O, nb, nog;
    BasisTab tab;
};

template <int N, int SMAX, bool DISC>
struct DwdShape {
    static constexpr int NB = DISC ? N * SMAX : (N - 1) * SMAX + 1;
};

template <int N, int SMAX, int OC, bool DISC>
__global__ void __launch_bounds__(256) pw_dw_dense_kernel(const __grid_constant__ DwdParams p) {
    constexpr int NB = DwdShape<N, SMAX, DISC>::NB;
    constexpr int ST = DISC ? N : N - 1;
    __shared__ float red[8][NB * OC];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = blockIdx.x % p.I;
    const int og = blockIdx.x / p.I;
    const int ks = blockIdx.y;
    const int o0 = og * OC;
    const int64_t r_lo = (int64_t)ks * p.rows_per_split;
    const int64_t r_hi = min(p.B, r_lo + p.rows_per_split);
    const float* xl = p.xl_t + (int64_t)i * p.Bp;
    const uint16_t* sgp = p.seg_t + (int64_t)i * p.Bp;
    const int vecw = (p.O % 4 == 0) ? 4 : ((p.O % 2 == 0) ? 2 : 1);  // o0 is a multiple of 4

    float2 acc[NB][OC / 2];
#pragma unroll
    for (int v = 0; v < NB; ++v)
#pragma unroll
        for (int q = 0; q < OC / 2; ++q) acc[v][q] = make_float2(0.0f, 0.0f);

#pragma unroll 2
    for (int64_t b = r_lo + tid; b < r_hi; b += 256) {
        const float t = __ldg(xl + b);
        const uint32_t sg = __ldg(sgp + b);
        float g[OC];
        const float* row = p.gy + b * p.O + o0;
        if (vecw == 4) {
#pragma unroll
            for (int q = 0; q < OC / 4; ++q) {
                float4 v4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (o0 + 4 * q < p.O) v4 = __ldg(reinterpret_cast<const float4*>(row + 4 * q));
                g[4 * q] = v4.x, g[4 * q + 1] = v4.y, g[4 * q + 2] = v4.z, g[4 * q + 3] = v4.w;
            }
        } else if (vecw == 2) {
#pragma unroll
            for (int q = 0; q < OC / 2; ++q) {
                float2 v2 = make_float2(0.0f, 0.0f);
                if (o0 + 2 * q < p.O) v2 = __ldg(reinterpret_cast<const float2*>(row + 2 * q));
                g[2 * q] = v2.x, g[2 * q + 1] = v2.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < OC; ++q) g[q] = (o0 + q < p.O) ? __ldg(row + q) : 0.0f;
        }
        float phi[N];
        lagrange_basis<N>(t, p.tab, phi);
        const int s = (sg & HOL_SEG_INVALID) ? -1 : (int)(sg & HOL_SEG_MASK);
        float coef[NB];
#pragma unroll
        for (int v = 0; v < NB; ++v) coef[v] = 0.0f;
#pragma unroll
        for (int ss = 0; ss < SMAX; ++ss) {
            // continuous n == 1 has one weight per link whatever the segment
            const float m = ((!DISC && N == 1) ? (s >= 0 && ss == 0) : (s == ss)) ? 1.0f : 0.0f;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const int v = (!DISC && N == 1) ? 0 : ss * ST + j;
                if (v < NB) coef[v] = fmaf(m, phi[j], coef[v]);  // end nodes shared by two segments: at most one is non-zero
            }
        }
#pragma unroll
        for (int v = 0; v < NB; ++v)
#pragma unroll
            for (int q = 0; q < OC / 2; ++q) acc[v][q] = ffma2(coef[v], make_float2(g[2 * q], g[2 * q + 1]), acc[v][q]);
    }

    // lanes (fixed butterfly) -> warps (in order) -> global
#pragma unroll
    for (int v = 0; v < NB; ++v)
#pragma unroll
        for (int q = 0; q < OC / 2; ++q) {
            float ax = acc[v][q].x, ay = acc[v][q].y;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, m);
                ay += __shfl_xor_sync(0xffffffffu, ay, m);
            }
            if (lane == 0) {
                red[warp][v * OC + 2 * q] = ax;
                red[warp][v * OC + 2 * q + 1] = ay;
            }
        }
    __syncthreads();
    if (tid < NB * OC) {
        const int v = tid / OC, o = o0 + tid % OC;
        if (v < p.nb && o < p.O) {
            float sum = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += red[w][tid];
            p.out[(int64_t)ks * p.part_stride + ((int64_t)o * p.I + i) * p.nb + v] = sum;
        }
    }
}

// Plan: which template handles (n, S, O) -- or none.  OC = outputs per thread, SMAX = compiled
// segment bound; NB(SMAX) * OC <= 128 accumulators.
bool dwd_plan(int n, int S, int O, int disc, int64_t B, int I, DwdPlan& d) {
    d.enabled = false;
    const char* off = getenv("HOL_DISABLE_DWD");
    if (off && off[0] == '1') return false;
    if (n > 5 || S > 8 || O >= 32 || B < 1) return false;
    const int smax = S <= 2 ? 2 : (S <= 4 ? 4 : 8);
    const int nbmax = (!disc && n == 1) ? 1 : (disc ? n * smax : (n - 1) * smax + 1);
    int oc = O <= 4 ? 4 : 8;
    if (nbmax * oc > 128) oc = 4;
    if (nbmax * oc > 128) return false;
    d.smax = smax;
    d.oc = oc;
    d.nog = (O + oc - 1) / oc;
    // sample ranges: enough CTAs for ~4 per SM, at least 8 rows per thread
    int64_t ks = (4 * 148 + (int64_t)I * d.nog - 1) / ((int64_t)I * d.nog);
    const int64_t ks_max = B / (256 * 8) > 1 ? B / (256 * 8) : 1;
    if (ks > ks_max) ks = ks_max;
    if (ks > 64) ks = 64;
    if (ks < 1) ks = 1;
    d.rows_per_split = (B + ks - 1) / ks;
    d.ksplit = (int)((B + d.rows_per_split - 1) / d.rows_per_split);
    d.enabled = true;
    return true;
}

template <int N, int SMAX, int OC, bool DISC>
static int launch_dwd_t(const PwShape& s, const PwWorkspace& ws, const float* gy, float* out, cudaStream_t st) {
    if constexpr (DwdShape<N, SMAX, DISC>::NB * OC > 128) {
        return HOL_EUNSUPPORTED;
    } else {
        DwdParams p;
        p.xl_t = ws.xl_t;
        p.seg_t = ws.seg_t;
        p.gy = gy;
        p.out = out;
        p.B = s.B;
        p.Bp = s.Bp;
        p.rows_per_split = s.dwd.rows_per_split;
        p.part_stride = (int64_t)s.O * s.I * s.nb;
        p.I = s.I;
        p.O = s.O;
        p.nb = s.nb;
        p.nog = s.dwd.nog;
        p.tab = s.tab;
        dim3 grid((unsigned)(s.I * s.dwd.nog), (unsigned)s.dwd.ksplit);
        pw_dw_dense_kernel<N, SMAX, OC, DISC><<<grid, 256, 0, st>>>(p);
        return (int)cudaGetLastError();
    }
}

template <int N, int SMAX, bool DISC>
static int launch_dwd_oc(const PwShape& s, const PwWorkspace& ws, const float* gy, float* out, cudaStream_t st) {
    return s.dwd.oc == 8 ? launch_dwd_t<N, SMAX, 8, DISC>(s, ws, gy, out, st)
                         : launch_dwd_t<N, SMAX, 4, DISC>(s, ws, gy, out, st);
}

template <int N, bool DISC>
static int launch_dwd_s(const PwShape& s, const PwWorkspace& ws, const float* gy, float* out, cudaStream_t st) {
    switch (s.dwd.smax) {
        case 2: return launch_dwd_oc<N, 2, DISC>(s, ws, gy, out, st);
        case 4: return launch_dwd_oc<N, 4, DISC>(s, ws, gy, out, st);
        case 8: return launch_dwd_oc<N, 8, DISC>(s, ws, gy, out, st);
    }
    return HOL_EUNSUPPORTED;
}

template <bool DISC>
static int launch_dwd_n(const PwShape& s, const PwWorkspace& ws, const float* gy, float* out, cudaStream_t st) {
    switch (s.n) {
        case 1: return launch_dwd_s<1, DISC>(s, ws, gy, out, st);
        case 2: return launch_dwd_s<2, DISC>(s, ws, gy, out, st);
        case 3: return launch_dwd_s<3, DISC>(s, ws, gy, out, st);
        case 4: return launch_dwd_s<4, DISC>(s, ws, gy, out, st);
        case 5: return launch_dwd_s<5, DISC>(s, ws, gy, out, st);
    }
    return HOL_EUNSUPPORTED;
}

int launch_backward_dw_dense(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dw, cudaStream_t st) {
    float* out = s.dwd.ksplit > 1 ? ws.dw_part : dw;
    const int rc = s.disc ? launch_dwd_n<true>(s, ws, gy, out, st) : launch_dwd_n<false>(s, ws, gy, out, st);
    if (rc != HOL_OK) return rc;
    if (s.dwd.ksplit > 1) return launch_dw_reduce(ws.dw_part, dw, (int64_t)s.O * s.I * s.nb, s.dwd.ksplit, st);
    return HOL_OK;
}
