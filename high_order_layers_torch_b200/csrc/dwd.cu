// Weight gradient of NARROW layers (conv: O = 6, 16; classifier heads: O = 10; few segments):
//   dW[o,i,v] = sum_b coef_v(b,i) * gy[b,o],   coef_v = phi_j(xl(b,i)) if v == s(b,i)*stride + j else 0.
//
// With only nb*O <= ~128 gradient entries per input the whole [nb][OC] block of one input lives
// in the registers of every thread: lanes are consecutive samples (xl_t / seg_t are read fully
// coalesced, no bucket sort, no gather of gy rows), the one-hot over the segments is expanded
// into the nb coefficients with S predicated multiplies, and the block is updated with nb*OC/2
// packed FMAs per (sample, input).  That spends nb/n times the useful flops -- on the FMA pipe,
// which is idle in the bucketed kernel (dw.cu) for such layers: there a warp gathers one
// 24-64 byte gy row per sample for n*4 FMAs.
//
// One CTA = (output group, input group, sample-range group).  The gy columns of a sample range
// are staged ONCE in shared memory and reused for every input of the group (the first version
// re-read gy from L2 per input and was L2-bandwidth bound); per input the CTA then streams only
// xl_t / seg_t (6 bytes per (sample, input)).  Partial sums are combined in a fixed order
// (transposing lane butterfly, 8 warps in order, sample ranges in order: single writer per
// partial, then pw_dw_reduce_kernel), so the result is bit-reproducible and weights no sample
// touched are exact zeros.
#include <stdlib.h>

#include "kernels.cuh"

struct DwdParams {
    const float* xl_t;
    const uint16_t* seg_t;
    const float* gy;
    float* out;  // dw, or the partial buffer [nkg][O][I][nb]
    int64_t B, Bp, part_stride;
    int I, O, nb, nog, R, P, G, nranges, nkg;
    BasisTab tab;
};

template <int N, int SMAX, bool DISC>
struct DwdShape {
    static constexpr int NB = DISC ? N * SMAX : (N - 1) * SMAX + 1;
};

// One step of the transposing butterfly: CNT values per lane -> CNT/2, lanes whose `off` bit is
// set keep the upper half of the index range.
template <int CNT>
__device__ __forceinline__ void dwd_fold(float* v, int lane, int off) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < CNT / 2; ++k) {
        const float send = up ? v[k] : v[k + CNT / 2];
        const float keep = up ? v[k + CNT / 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
}

template <int N, int SMAX, int OC, bool DISC>
__global__ void __launch_bounds__(256) pw_dw_dense_kernel(const __grid_constant__ DwdParams p) {
    constexpr int NB = DwdShape<N, SMAX, DISC>::NB;
    constexpr int ST = DISC ? N : N - 1;
    constexpr int NV = NB * OC, NVP = (NV + 31) / 32 * 32;
    extern __shared__ __align__(16) float gys[];  // [R][P] staged gy columns of this output group
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int og = blockIdx.x % p.nog, ig = blockIdx.x / p.nog;
    const int kg = blockIdx.y;
    const int o0 = og * OC;
    const int i_lo = ig * p.G, i_hi = min(p.I, i_lo + p.G);
    const int P = p.P;
    float* out = p.out + (int64_t)kg * p.part_stride;
    // after the butterfly this lane holds the sums of the NVP/32 consecutive entries starting here
    const int base = ((lane & 16) ? NVP / 2 : 0) + ((lane & 8) ? NVP / 4 : 0) + ((lane & 4) ? NVP / 8 : 0) +
                     ((lane & 2) ? NVP / 16 : 0) + ((lane & 1) ? NVP / 32 : 0);
    bool first = true;

    for (int range = kg; range < p.nranges; range += p.nkg) {
        const int64_t r_lo = (int64_t)range * p.R;
        const int nrows = (int)min((int64_t)p.R, p.B - r_lo);
        __syncthreads();
        for (int idx = tid; idx < nrows * OC; idx += 256) {
            const int r = idx / OC, c = idx % OC;
            gys[r * P + c] = (o0 + c < p.O) ? __ldg(p.gy + (r_lo + r) * p.O + o0 + c) : 0.0f;
        }
        __syncthreads();
        // every warp owns whole inputs: its lanes sweep the staged range, no cross-warp reduction
        for (int i = i_lo + warp; i < i_hi; i += 8) {
            const float* xl = p.xl_t + (int64_t)i * p.Bp + r_lo;
            const uint16_t* sgp = p.seg_t + (int64_t)i * p.Bp + r_lo;
            float2 acc[NB][OC / 2];
#pragma unroll
            for (int v = 0; v < NB; ++v)
#pragma unroll
                for (int q = 0; q < OC / 2; ++q) acc[v][q] = make_float2(0.0f, 0.0f);

#pragma unroll 4
            for (int rr = lane; rr < nrows; rr += 32) {
                const float t = __ldg(xl + rr);
                const uint32_t sg = __ldg(sgp + rr);
                float2 g[OC / 2];
#pragma unroll
                for (int q = 0; q < OC / 2; ++q) g[q] = *reinterpret_cast<const float2*>(gys + rr * P + 2 * q);
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
                        if (v < NB) coef[v] = fmaf(m, phi[j], coef[v]);  // shared end nodes: at most one term is non-zero
                    }
                }
#pragma unroll
                for (int v = 0; v < NB; ++v)
#pragma unroll
                    for (int q = 0; q < OC / 2; ++q) acc[v][q] = ffma2(coef[v], g[q], acc[v][q]);
            }

            // lanes: transposing butterfly (NVP values per lane -> NVP/32), fixed order
            float v[NVP];
#pragma unroll
            for (int k = 0; k < NVP; ++k) v[k] = 0.0f;
#pragma unroll
            for (int vv = 0; vv < NB; ++vv)
#pragma unroll
                for (int q = 0; q < OC / 2; ++q) {
                    v[vv * OC + 2 * q] = acc[vv][q].x;
                    v[vv * OC + 2 * q + 1] = acc[vv][q].y;
                }
            dwd_fold<NVP>(v, lane, 16);
            dwd_fold<NVP / 2>(v, lane, 8);
            dwd_fold<NVP / 4>(v, lane, 4);
            dwd_fold<NVP / 8>(v, lane, 2);
            dwd_fold<NVP / 16>(v, lane, 1);
#pragma unroll
            for (int k = 0; k < NVP / 32; ++k) {
                const int e = base + k;
                const int vv = e / OC, o = o0 + e % OC;
                if (e < NV && vv < p.nb && o < p.O) {
                    float* dst = out + ((int64_t)o * p.I + i) * p.nb + vv;
                    *dst = first ? v[k] : *dst + v[k];  // single writer per partial entry, ranges in order
                }
            }
        }
        first = false;
    }
}

// Plan: which template handles (n, S, O) -- or none.  OC = outputs per thread, SMAX = compiled
// segment bound; NB(SMAX) * OC <= 128 accumulators.
bool dwd_plan(int n, int S, int O, int disc, int64_t B, int I, DwdPlan& d) {
    d.enabled = false;
    if (hol_plan_overrides() & HOL_OVR_DISABLE_DWD) return false;
    if (n > 5 || S > 8 || O >= 32 || B < 1) return false;
    const int smax = S <= 2 ? 2 : (S <= 4 ? 4 : 8);
    const int nbmax = (!disc && n == 1) ? 1 : (disc ? n * smax : (n - 1) * smax + 1);
    // outputs per thread: 4, 6 or 8 -- the choice with the fewest padded columns that still fits
    // 128 accumulators (O = 6 -> 6, O = 10 -> 2 x 6, O = 16 -> 2 x 8)
    int oc = 0, best = 1 << 30;
    for (int c : {8, 6, 4}) {
        if (nbmax * c > 128) continue;
        const int slots = (O + c - 1) / c * c;
        if (slots < best) best = slots, oc = c;
    }
    if (!oc) return false;
    d.smax = smax;
    d.oc = oc;
    d.nog = (O + oc - 1) / oc;
    // staged gy pitch (floats): even, and P/2 odd so that the 64-bit reads of consecutive rows are
    // conflict free; ~96 KB of gy per CTA (two CTAs per SM)
    d.P = (oc / 2) % 2 ? oc : oc + 2;
    int64_t R = (96 * 1024) / (d.P * 4) / 256 * 256;
    if (R > (B + 255) / 256 * 256) R = (B + 255) / 256 * 256;
    for (;;) {
        d.R = (int)R;
        d.nranges = (int)((B + R - 1) / R);
        d.nkg = d.nranges < 64 ? d.nranges : 64;
        // input groups: enough CTAs for ~4 per SM, at least one input per warp where possible; gy
        // is staged once per group
        int64_t nig = (4 * 148 + (int64_t)d.nkg * d.nog - 1) / ((int64_t)d.nkg * d.nog);
        if (nig > (I + 7) / 8) nig = (I + 7) / 8;
        if (nig < 1) nig = 1;
        d.G = (int)((I + nig - 1) / nig);
        d.G = (d.G + 7) / 8 * 8;
        d.nig = (I + d.G - 1) / d.G;
        // small problems: shorter sample ranges until every SM has ~2 CTAs
        if ((int64_t)d.nkg * d.nog * d.nig >= 2 * 148 || R <= 1024 || d.nkg >= 64) break;
        R /= 2;
    }
    d.ksplit = d.nkg;
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
        p.part_stride = (int64_t)s.O * s.I * s.nb;
        p.I = s.I;
        p.O = s.O;
        p.nb = s.nb;
        p.nog = s.dwd.nog;
        p.R = s.dwd.R;
        p.P = s.dwd.P;
        p.G = s.dwd.G;
        p.nranges = s.dwd.nranges;
        p.nkg = s.dwd.nkg;
        p.tab = s.tab;
        const size_t smem = (size_t)p.R * p.P * sizeof(float);
        auto kern = pw_dw_dense_kernel<N, SMAX, OC, DISC>;
        static std::atomic<uint64_t> attr_done{0};
        cudaError_t e = hol_ensure_max_smem(kern, attr_done);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)(s.dwd.nig * s.dwd.nog), (unsigned)s.dwd.nkg);
        kern<<<grid, 256, smem, st>>>(p);
        return (int)cudaGetLastError();
    }
}

template <int N, int SMAX, bool DISC>
static int launch_dwd_oc(const PwShape& s, const PwWorkspace& ws, const float* gy, float* out, cudaStream_t st) {
    switch (s.dwd.oc) {
        case 8: return launch_dwd_t<N, SMAX, 8, DISC>(s, ws, gy, out, st);
        case 6: return launch_dwd_t<N, SMAX, 6, DISC>(s, ws, gy, out, st);
        case 4: return launch_dwd_t<N, SMAX, 4, DISC>(s, ws, gy, out, st);
    }
    return HOL_EUNSUPPORTED;
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
