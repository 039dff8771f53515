// Weight gradient  dW[o,i,s*stride+j] = sum_{b : s(b,i)=s} gy[b,o] * phi_j(xl(b,i)).
//
// Segment-bucketed reduction, no global atomics on the hot path: a stable chunked counting sort
// (prep.cu) lists the samples of every (input, segment) bucket; one warp owns one
// (output tile, input, segment, chunk group) and walks its buckets with the N x 4 partial sums
// of each lane in registers (one coalesced gy row segment per sample, N packed FMAs per loaded
// float2).  Lanes are (samples in flight) x (4-output quads): wide layers use all 32 lanes for
// 128 outputs of one sample, narrow layers (conv: O = 6, 16) process 32/LO samples at once and
// combine them with a fixed shuffle tree.  Every dW element has a single writer and a fixed
// summation order, so the result is bit-reproducible; untouched weights are exact zeros.
// When a layer has too few (tile, input, segment) tasks to fill the GPU, the sample chunks are
// split into `ksplit` groups whose partial results are summed by a second kernel in a fixed
// order.  The one use of atomics is the end node shared by two neighbouring segments of a
// continuous layer: its two partial sums are combined with one atomicAdd each onto zeroed
// memory (two addends: order-independent).
#include "kernels.cuh"

struct DwParams {
    const float* xl_t;
    const uint32_t* order;
    const uint32_t* offs;
    const float* gy;
    float* out;  // dw, or the partial buffer [ksplit][O][I][nb]
    int64_t B, Bp, ntasks, part_stride;
    int I, O, nb, stride, S, Seff, disc, nchunks, ksplit, lo, n_ot;
    BasisTab tab;
};

template <int N>
__global__ void __launch_bounds__(256, (N <= 5 ? 3 : 2)) pw_dw_kernel(const __grid_constant__ DwParams p) {
    constexpr int PH = (N + 1 + 3) / 4 * 4;  // floats per staged sample: phi[N], sample id, pad
    __shared__ __align__(16) float sc[8][32][PH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t task = (int64_t)blockIdx.x * 8 + warp;
    if (task >= p.ntasks) return;
    // task -> (output tile, chunk group, input, segment); tiles outermost so that concurrently
    // running warps share the same gy columns in L2
    const int s = (int)(task % p.Seff);
    task /= p.Seff;
    const int i = (int)(task % p.I);
    task /= p.I;
    const int kg = (int)(task % p.ksplit);
    const int ot = (int)(task / p.ksplit);
    const int LO = p.lo, GS = 32 / LO;  // lanes per sample, samples in flight
    const int g = lane / LO, ol = lane % LO;
    const int o = ot * (4 * LO) + 4 * ol;
    const bool vec = (p.O & 3) == 0;
    const uint32_t* ord = p.order + (int64_t)i * p.Bp;
    const float* xl = p.xl_t + (int64_t)i * p.Bp;
    float(*my)[PH] = sc[warp];

    float2 acc[N][2];
#pragma unroll
    for (int j = 0; j < N; ++j) acc[j][0] = acc[j][1] = make_float2(0.0f, 0.0f);

    const int cpg = (p.nchunks + p.ksplit - 1) / p.ksplit;
    const int c_lo = kg * cpg, c_hi = min(p.nchunks, c_lo + cpg);
    for (int c = c_lo; c < c_hi; ++c) {
        const uint32_t* co = p.offs + ((int64_t)i * p.nchunks + c) * (p.Seff + 1);
        const uint32_t start = co[s], end = co[s + 1];
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
            for (int t0 = 0; t0 < cnt; t0 += 4 * GS) {
                float4 gq[4];
                float ph[4][PH];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tt = t0 + u * GS + g;
                    const int t = min(tt, cnt - 1);
#pragma unroll
                    for (int q = 0; q < PH / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(&my[t][4 * q]);
                        ph[u][4 * q + 0] = v.x;
                        ph[u][4 * q + 1] = v.y;
                        ph[u][4 * q + 2] = v.z;
                        ph[u][4 * q + 3] = v.w;
                    }
                    const float* row = p.gy + (int64_t)__float_as_uint(ph[u][N]) * p.O + o;
                    gq[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (tt < cnt) {
                        if (vec && o + 3 < p.O) {
                            gq[u] = __ldg(reinterpret_cast<const float4*>(row));
                        } else {
                            if (o + 0 < p.O) gq[u].x = __ldg(row + 0);
                            if (o + 1 < p.O) gq[u].y = __ldg(row + 1);
                            if (o + 2 < p.O) gq[u].z = __ldg(row + 2);
                            if (o + 3 < p.O) gq[u].w = __ldg(row + 3);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        acc[j][0] = ffma2(ph[u][j], make_float2(gq[u].x, gq[u].y), acc[j][0]);
                        acc[j][1] = ffma2(ph[u][j], make_float2(gq[u].z, gq[u].w), acc[j][1]);
                    }
                }
            }
            __syncwarp();
        }
    }

    // combine the GS samples-in-flight groups (fixed tree), then lanes of group 0 write
    for (int m = LO; m < 32; m <<= 1) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            acc[j][0].x += __shfl_xor_sync(0xffffffffu, acc[j][0].x, m);
            acc[j][0].y += __shfl_xor_sync(0xffffffffu, acc[j][0].y, m);
            acc[j][1].x += __shfl_xor_sync(0xffffffffu, acc[j][1].x, m);
            acc[j][1].y += __shfl_xor_sync(0xffffffffu, acc[j][1].y, m);
        }
    }
    if (g != 0) return;
    float* out = p.out + (int64_t)kg * p.part_stride;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int v = s * p.stride + j;  // reference column
        const bool shared_node = !p.disc && N > 1 && ((j == 0 && s > 0) || (j == N - 1 && s < p.S - 1));
        const float vals[4] = {acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (o + c < p.O) {
                float* dst = out + ((int64_t)(o + c) * p.I + i) * p.nb + v;
                if (shared_node)
                    atomicAdd(dst, vals[c]);
                else
                    *dst = vals[c];
            }
        }
    }
}

// dw[e] = sum_k part[k][e], k ascending (deterministic)
__global__ void __launch_bounds__(256) pw_dw_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                                           int64_t count, int ksplit) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    float acc = 0.0f;
    for (int k = 0; k < ksplit; ++k) acc += part[(int64_t)k * count + e];
    dw[e] = acc;
}

template <int N>
static int launch_dw_n(const PwShape& s, const PwWorkspace& ws, const float* gy, float* out, cudaStream_t st) {
    DwParams p;
    p.xl_t = ws.xl_t;
    p.order = ws.order;
    p.offs = ws.offs;
    p.gy = gy;
    p.out = out;
    p.B = s.B;
    p.Bp = s.Bp;
    p.I = s.I;
    p.O = s.O;
    p.nb = s.nb;
    p.stride = s.stride;
    p.S = s.S;
    p.Seff = s.Seff;
    p.disc = s.disc;
    p.nchunks = s.nchunks;
    p.ksplit = s.ksplit;
    p.lo = s.dw_lo;
    p.tab = s.tab;
    p.part_stride = (int64_t)s.O * s.I * s.nb;
    const int64_t otw = 4 * s.dw_lo;
    p.n_ot = (int)((s.O + otw - 1) / otw);
    p.ntasks = (int64_t)p.n_ot * s.ksplit * s.I * s.Seff;
    const int64_t blocks = (p.ntasks + 7) / 8;
    if (blocks > 0x7fffffffLL) return HOL_EUNSUPPORTED;
    pw_dw_kernel<N><<<(unsigned)blocks, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

int launch_backward_dw(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dw, cudaStream_t st) {
    const int64_t count = (int64_t)s.O * s.I * s.nb;
    float* out = s.ksplit > 1 ? ws.dw_part : dw;
    // shared end nodes of continuous layers are accumulated onto zeros
    if (!s.disc && s.n > 1 && s.S > 1) {
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)count * (s.ksplit > 1 ? s.ksplit : 1) * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = HOL_EUNSUPPORTED;
    switch (s.n) {
        case 1: rc = launch_dw_n<1>(s, ws, gy, out, st); break;
        case 2: rc = launch_dw_n<2>(s, ws, gy, out, st); break;
        case 3: rc = launch_dw_n<3>(s, ws, gy, out, st); break;
        case 4: rc = launch_dw_n<4>(s, ws, gy, out, st); break;
        case 5: rc = launch_dw_n<5>(s, ws, gy, out, st); break;
        case 6: rc = launch_dw_n<6>(s, ws, gy, out, st); break;
        case 7: rc = launch_dw_n<7>(s, ws, gy, out, st); break;
        case 8: rc = launch_dw_n<8>(s, ws, gy, out, st); break;
        case 9: rc = launch_dw_n<9>(s, ws, gy, out, st); break;
        case 10: rc = launch_dw_n<10>(s, ws, gy, out, st); break;
    }
    if (rc != HOL_OK) return rc;
    if (s.ksplit > 1) return launch_dw_reduce(ws.dw_part, dw, count, s.ksplit, st);
    return HOL_OK;
}

int launch_dw_reduce(const float* part, float* dw, int64_t count, int ksplit, cudaStream_t st) {
    pw_dw_reduce_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(part, dw, count, ksplit);
    return (int)cudaGetLastError();
}
