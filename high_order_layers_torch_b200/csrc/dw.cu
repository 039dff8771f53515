// Weight gradient  dW[o,i,s*stride+j] = sum_{b : s(b,i)=s} gy[b,o] * phi_j(xl(b,i)).
//
// Segment-bucketed reduction, no global atomics on the hot path: a stable per-input counting
// sort (prep.cu) lists the samples of every (input, segment) bucket; one warp owns one
// (128-output tile, input, segment) and walks its bucket with the N x 4 partial sums of each
// lane in registers (one coalesced 512-byte gy row per sample, N packed FMAs per loaded
// float2).  Every dW element has a single writer and a fixed summation order, so the result is
// bit-reproducible; untouched weights are written as exact zeros.  The one exception is the end
// node shared by two neighbouring segments of a continuous layer: its two partial sums are
// combined with one atomicAdd each onto zeroed memory (two addends: order-independent).
#include "kernels.cuh"

struct DwParams {
    const float* xl_t;
    const uint32_t* order;
    const uint32_t* offs;
    const float* gy;
    float* dw;
    int64_t B, Bp, ntasks;
    int I, O, nb, stride, S, Seff, disc;
    BasisTab tab;
};

template <int N>
__global__ void __launch_bounds__(256) pw_dw_kernel(const __grid_constant__ DwParams p) {
    constexpr int PH = (N + 1 + 3) / 4 * 4;  // floats per staged sample: phi[N], sample id, pad
    __shared__ __align__(16) float sc[8][32][PH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * 8 + warp;
    if (task >= p.ntasks) return;
    const int64_t per_ot = (int64_t)p.I * p.Seff;
    const int ot = (int)(task / per_ot);
    const int rem = (int)(task % per_ot);
    const int i = rem / p.Seff;
    const int s = rem % p.Seff;
    const int o = ot * 128 + 4 * lane;
    const bool vec = (p.O & 3) == 0;
    const uint32_t start = p.offs[(int64_t)i * (p.Seff + 1) + s];
    const uint32_t end = p.offs[(int64_t)i * (p.Seff + 1) + s + 1];
    const uint32_t* ord = p.order + (int64_t)i * p.Bp;
    const float* xl = p.xl_t + (int64_t)i * p.Bp;
    float(*my)[PH] = sc[warp];

    float2 acc[N][2];
#pragma unroll
    for (int j = 0; j < N; ++j) acc[j][0] = acc[j][1] = make_float2(0.0f, 0.0f);

    for (uint32_t base = start; base < end; base += 32) {
        const int cnt = (int)min(32u, end - base);
        if (lane < cnt) {
            const uint32_t b = ord[base + lane];
            float phi[N];
            lagrange_basis<N>(xl[b], p.tab, phi);
#pragma unroll
            for (int j = 0; j < N; ++j) my[lane][j] = phi[j];
            my[lane][N] = __uint_as_float(b);
        }
        __syncwarp();
        for (int t0 = 0; t0 < cnt; t0 += 4) {
            float4 g[4];
            float ph[4][PH];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = min(t0 + u, cnt - 1);
#pragma unroll
                for (int q = 0; q < PH / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(&my[t][4 * q]);
                    ph[u][4 * q + 0] = v.x;
                    ph[u][4 * q + 1] = v.y;
                    ph[u][4 * q + 2] = v.z;
                    ph[u][4 * q + 3] = v.w;
                }
                const float* row = p.gy + (int64_t)__float_as_uint(ph[u][N]) * p.O + o;
                g[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (t0 + u < cnt) {
                    if (vec && o + 3 < p.O) {
                        g[u] = __ldg(reinterpret_cast<const float4*>(row));
                    } else {
                        if (o + 0 < p.O) g[u].x = __ldg(row + 0);
                        if (o + 1 < p.O) g[u].y = __ldg(row + 1);
                        if (o + 2 < p.O) g[u].z = __ldg(row + 2);
                        if (o + 3 < p.O) g[u].w = __ldg(row + 3);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    acc[j][0] = ffma2(ph[u][j], make_float2(g[u].x, g[u].y), acc[j][0]);
                    acc[j][1] = ffma2(ph[u][j], make_float2(g[u].z, g[u].w), acc[j][1]);
                }
            }
        }
        __syncwarp();
    }

    // write the n rows of this (input, segment); reference column v = s*stride + j
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int v = s * p.stride + j;
        const bool shared_node = !p.disc && N > 1 && ((j == 0 && s > 0) || (j == N - 1 && s < p.S - 1));
        const float vals[4] = {acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (o + c < p.O) {
                float* dst = p.dw + ((int64_t)(o + c) * p.I + i) * p.nb + v;
                if (shared_node)
                    atomicAdd(dst, vals[c]);
                else
                    *dst = vals[c];
            }
        }
    }
}

template <int N>
static int launch_dw_n(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dw, cudaStream_t st) {
    DwParams p;
    p.xl_t = ws.xl_t;
    p.order = ws.order;
    p.offs = ws.offs;
    p.gy = gy;
    p.dw = dw;
    p.B = s.B;
    p.Bp = s.Bp;
    p.I = s.I;
    p.O = s.O;
    p.nb = s.nb;
    p.stride = s.stride;
    p.S = s.S;
    p.Seff = s.Seff;
    p.disc = s.disc;
    p.tab = s.tab;
    const int64_t n128 = (s.O + 127) / 128;
    p.ntasks = n128 * s.I * s.Seff;
    const int64_t blocks = (p.ntasks + 7) / 8;
    if (blocks > 0x7fffffffLL) return HOL_EUNSUPPORTED;
    pw_dw_kernel<N><<<(unsigned)blocks, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

int launch_backward_dw(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dw, cudaStream_t st) {
    // shared end nodes of continuous layers are accumulated onto zeros
    if (!s.disc && s.n > 1 && s.S > 1) {
        cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)s.O * s.I * s.nb * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    switch (s.n) {
        case 1: return launch_dw_n<1>(s, ws, gy, dw, st);
        case 2: return launch_dw_n<2>(s, ws, gy, dw, st);
        case 3: return launch_dw_n<3>(s, ws, gy, dw, st);
        case 4: return launch_dw_n<4>(s, ws, gy, dw, st);
        case 5: return launch_dw_n<5>(s, ws, gy, dw, st);
        case 6: return launch_dw_n<6>(s, ws, gy, dw, st);
        case 7: return launch_dw_n<7>(s, ws, gy, dw, st);
        case 8: return launch_dw_n<8>(s, ws, gy, dw, st);
        case 9: return launch_dw_n<9>(s, ws, gy, dw, st);
        case 10: return launch_dw_n<10>(s, ws, gy, dw, st);
    }
    return HOL_EUNSUPPORTED;
}
