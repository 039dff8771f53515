// Input gradient
//   dX[b,i] = sigma * length/(eta(s+1)-eta(s)) * sum_j phi'_j(xl) * t_j,
//   t_j = sum_o gy[b,o] * Wk[i][row(j, s(b,i))][o]
// (the backward autograd derives for PolynomialLayers.py:329-335 + Basis.py:492-512).
//
// Same mapping as the forward: one thread = one sample; a CTA owns BT samples x IT inputs and
// streams every output tile of those IT inputs through shared memory (the same contiguous
// packed blocks the forward stages).  The IT*N dot products t_j live in registers as fp32x2
// pairs, so dX[b,i] has exactly one writer and a fixed summation order (no atomics).
#include "kernels.cuh"

struct DxParams {
    const float* xl_t;
    const uint16_t* seg_t;
    const float* wk;
    const float* gy;    // [B][O] (narrow tiles, OT <= 16)
    const float* gy_t;  // [nOT*OT][Bp] (OT >= 32): lanes (consecutive samples) read consecutive words
    float* dx;
    int64_t B, Bp;
    int I, O, R, nOT, stage_floats, srow, transposed, it_eff, nst;
    int jrow[HOL_MAX_N];
    float length, Sf;
    BasisTab tab;
};

// ROT (S > 8, OT >= 32): the packed rows have a pitch of OT floats and lane l walks the 16-byte
// column chunks of a tile in the rotated order ((qb + (l%8)/4) % (OT/16), (t + l%8) % 4), so the 8
// lanes of a quarter warp always read 8 different bank groups whatever their segments are (see
// fwd.cu); the matching gy chunk is loaded in the same order.
// IT = 2 (layers whose packed rows only let two inputs share a stage: many segments): 16 + 2 running sums leave
// room for 512-sample CTAs -- twice the warps per SM and half the weight traffic per sample.
template <int N, int OT, int IT, bool ROT>
__global__ void __launch_bounds__((IT <= 2 ? 512 : 256), (OT <= 16 && IT > 2 ? 2 : 1)) pw_dx_kernel(const __grid_constant__ DxParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float* slab0 = reinterpret_cast<float*>(smem_raw + 128);
    constexpr int OTP = ROT ? OT : OT + 4;
    const int rot = threadIdx.x & 7;
    const int tid = threadIdx.x;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + tid;
    const int in_floats = p.R * OTP;
    // A CTA walks the input groups blockIdx.y, blockIdx.y + gridDim.y, ... (it_eff <= IT inputs each,
    // as many as fit in shared memory) and, per group, every output tile: one continuous TMA
    // pipeline over the (group, tile) chunks, so narrow layers (a single tile) do not pay a
    // pipeline fill per 8 inputs.
    const int ngroups = (p.I + p.it_eff - 1) / p.it_eff;
    const int my_groups = ((int)blockIdx.y < ngroups) ? (ngroups - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
    const int total_chunks = my_groups * p.nOT;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        const int grp = blockIdx.y + (c / p.nOT) * gridDim.y, ot = c % p.nOT;
        const int gi0 = grp * p.it_eff;
        const uint32_t bytes = (uint32_t)min(p.it_eff, p.I - gi0) * (uint32_t)in_floats * 4u;
        const int stg = c % p.nst;
        uint64_t* bar = &bars[stg];
        mbar_expect_tx(bar, bytes);
        bulk_g2s(slab0 + (size_t)stg * p.stage_floats, p.wk + ((int64_t)ot * p.I + gi0) * in_floats, bytes, bar);
    };
    if (tid == 0 && total_chunks > 0) {
        issue(0);
        if (p.nst > 1 && total_chunks > 1) issue(1);
    }

    // Wide tiles read the TRANSPOSED upstream gradient (pw_transpose_gy_kernel): with the row-major gy every lane
    // of a warp-wide LDG.128 touches its own 128-byte line -- 32 LSU wavefronts per load, as many per block as the
    // weight reads themselves (config-5 sweep, n = 4, S = 64: dX 557 -> 398 ms).  Narrow tiles (OT <= 16: classifier
    // heads, conv layers) are not LSU-bound; they keep the four vector loads and skip the transpose launch.
    constexpr bool GT = OT >= 32;
    const bool rowok = b < p.B;
    const bool vec = (p.O & 3) == 0;
    const float* gyrow = GT ? nullptr : p.gy + (rowok ? b : 0) * p.O;
    const float* gyt = GT ? p.gy_t + b : nullptr;
    int rowoff[IT];
    float2 acc[IT][N];
    int i0 = 0, ni = 0;

    for (int c = 0; c < total_chunks; ++c) {
        const int ot = c % p.nOT;
        if (ot == 0) {
            i0 = (blockIdx.y + (c / p.nOT) * gridDim.y) * p.it_eff;
            ni = min(p.it_eff, p.I - i0);
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                uint32_t sg = 0;
                if (k < ni) sg = __ldg(p.seg_t + (int64_t)(i0 + k) * p.Bp + b);
                rowoff[k] = (k < ni ? k : 0) * in_floats + (int)(sg & HOL_SEG_MASK) * p.srow;
#pragma unroll
                for (int j = 0; j < N; ++j) acc[k][j] = make_float2(0.0f, 0.0f);
            }
        }
        mbar_wait(&bars[c % p.nst], (uint32_t)((c / p.nst) & 1));
        const float* slab = slab0 + (size_t)(c % p.nst) * p.stage_floats;
        constexpr int QW = OT < 16 ? OT : 16;  // output columns per register block
#pragma unroll 1
        for (int qb0 = 0; qb0 < OT / QW; ++qb0) {
            const int qb = ROT ? ((qb0 + (rot >> 2)) & (OT / QW - 1)) : qb0;
            // The thread's QW upstream gradients of this block, from the TRANSPOSED gy: a warp-wide 4-byte load is
            // one 128-byte line per group of lanes that works on the same block (two groups under ROT) -- the row-major
            // gy cost 32 lines per LDG.128 (one per lane), as many LSU wavefronts as the weight reads themselves.
            float4 g[QW / 4];
            if (!GT) {
#pragma unroll
                for (int t = 0; t < QW / 4; ++t) {
                    const int o = ot * OT + qb * QW + 4 * t;
                    g[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (rowok) {
                        if (vec && o + 3 < p.O) {
                            g[t] = __ldg(reinterpret_cast<const float4*>(gyrow + o));
                        } else {
                            if (o + 0 < p.O) g[t].x = __ldg(gyrow + o + 0);
                            if (o + 1 < p.O) g[t].y = __ldg(gyrow + o + 1);
                            if (o + 2 < p.O) g[t].z = __ldg(gyrow + o + 2);
                            if (o + 3 < p.O) g[t].w = __ldg(gyrow + o + 3);
                        }
                    }
                }
            } else {
                const float* src = gyt + (int64_t)(ot * OT + qb * QW) * p.Bp;
                float gn[QW];
#pragma unroll
                for (int e = 0; e < QW; ++e) gn[e] = __ldg(src + (int64_t)e * p.Bp);
#pragma unroll
                for (int t = 0; t < QW / 4; ++t) g[t] = make_float4(gn[4 * t], gn[4 * t + 1], gn[4 * t + 2], gn[4 * t + 3]);
            }
            if (ROT) {  // chunk t of this lane is output chunk (t + rot) & 3: rotate by rot & 3 in two select stages
                static_assert(!ROT || QW == 16, "the rotated order walks four 16-byte chunks per block");
                const bool r1 = (rot & 1) != 0, r2 = (rot & 2) != 0;
                float4 h[QW / 4];
#pragma unroll
                for (int t = 0; t < QW / 4; ++t) h[t] = r1 ? g[(t + 1) & (QW / 4 - 1)] : g[t];
#pragma unroll
                for (int t = 0; t < QW / 4; ++t) g[t] = r2 ? h[(t + 2) & (QW / 4 - 1)] : h[t];
            }
            const float* slabq = slab + qb * QW;
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                if (k >= ni) break;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const float4* wp = reinterpret_cast<const float4*>(slabq + rowoff[k] + p.jrow[j]);
#pragma unroll
                    for (int t = 0; t < QW / 4; ++t) {
                        const float4 w = wp[ROT ? ((t + rot) & 3) : t];
                        acc[k][j] = ffma2v(make_float2(g[t].x, g[t].y), make_float2(w.x, w.y), acc[k][j]);
                        acc[k][j] = ffma2v(make_float2(g[t].z, g[t].w), make_float2(w.z, w.w), acc[k][j]);
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && c + p.nst < total_chunks) issue(c + p.nst);
        if (ot != p.nOT - 1) continue;

#pragma unroll
        for (int k = 0; k < IT; ++k) {
            if (k >= ni) break;
            const int64_t idx = (int64_t)(i0 + k) * p.Bp + b;
            const float xl = __ldg(p.xl_t + idx);
            const uint32_t sg = __ldg(p.seg_t + idx);
            float dphi[N];
            lagrange_basis_derivative<N>(xl, p.tab, dphi);
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < N; ++j) t = fmaf(dphi[j], acc[k][j].x + acc[k][j].y, t);
            const int s = (int)(sg & HOL_SEG_MASK);
            const float e0 = __fsub_rn(__fmul_rn(__fdiv_rn((float)s, p.Sf), 2.0f), 1.0f);
            const float e1 = __fsub_rn(__fmul_rn(__fdiv_rn((float)(s + 1), p.Sf), 2.0f), 1.0f);
            float chain = __fdiv_rn(p.length, __fsub_rn(e1, e0));
            if (sg & HOL_SEG_MIRROR) chain = -chain;
            float out = (sg & HOL_SEG_INVALID) ? 0.0f : t * chain;
            if (p.transposed) {
                p.dx[idx] = out;
            } else if (rowok) {
                p.dx[b * p.I + i0 + k] = out;
            }
        }
    }
}

// gy[B][O] -> gy_t[Opad][Bp], zero beyond B and O (32 x 32 shared-memory tiles)
__global__ void __launch_bounds__(256) pw_transpose_gy_kernel(const float* __restrict__ gy, float* __restrict__ gy_t, int64_t B,
                                                              int64_t Bp, int O, int Opad) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const int o0 = blockIdx.y * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t b = b0 + r;
        const int o = o0 + tx;
        tile[r][tx] = (b < B && o < O) ? __ldg(gy + b * O + o) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int o = o0 + r;
        const int64_t b = b0 + tx;
        if (o < Opad && b < Bp) gy_t[(int64_t)o * Bp + b] = tile[tx][r];
    }
}

template <int N>
struct DxTile {
    static constexpr int IT = N <= 2 ? 16 : (N <= 5 ? 8 : 4);
};

template <int N, int OT, bool ROT>
static int launch_dx_t(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dx, int transposed,
                       cudaStream_t st) {
    constexpr int OTP = ROT ? OT : OT + 4;
    if (s.OTP != OTP) return HOL_EINVAL;
    constexpr int IT = DxTile<N>::IT;
    const size_t in_bytes = (size_t)s.R * OTP * 4;
    int it_eff = IT, nst = 2;
    while (it_eff > 1 && 128 + 2 * (size_t)it_eff * in_bytes > HOL_SMEM_BUDGET) --it_eff;
    if (128 + 2 * (size_t)it_eff * in_bytes > HOL_SMEM_BUDGET) nst = 1;
    const size_t smem = 128 + (size_t)nst * it_eff * in_bytes;
    if (smem > HOL_SMEM_BUDGET) return HOL_EUNSUPPORTED;
    DxParams p;
    p.xl_t = ws.xl_t;
    p.seg_t = ws.seg_t;
    p.wk = ws.wk;
    p.gy = gy;
    p.gy_t = ws.gy_t;
    p.dx = dx;
    p.B = s.B;
    p.Bp = s.Bp;
    p.I = s.I;
    p.O = s.O;
    p.R = s.R;
    p.nOT = s.nOT;
    p.stage_floats = it_eff * s.R * OTP;
    p.it_eff = it_eff;
    p.nst = nst;
    p.srow = s.srow;
    p.transposed = transposed;
    for (int j = 0; j < HOL_MAX_N; ++j) p.jrow[j] = s.jrow[j];
    p.length = s.length;
    p.Sf = (float)s.S;
    p.tab = s.tab;
    if (OT >= 32) {
        if (!ws.gy_t) return HOL_EINVAL;
        const int Opad = s.nOT * OT;
        dim3 tgrid((unsigned)((s.Bp + 31) / 32), (unsigned)((Opad + 31) / 32));
        pw_transpose_gy_kernel<<<tgrid, 256, 0, st>>>(gy, ws.gy_t, s.B, s.Bp, s.O, Opad);
        cudaError_t te = cudaGetLastError();
        if (te != cudaSuccess) return (int)te;
    }
    int BT;
    if (s.Bp % 512 != 0) {
        BT = (int)s.Bp;
    } else {
        BT = 256;
        if ((s.Bp / BT) * ((s.I + it_eff - 1) / it_eff) < 2 * 148) BT = 128;
    }
    if (IT > 2 && it_eff <= 2 && s.Bp % 512 == 0 && (s.Bp / 512) * ((s.I + it_eff - 1) / it_eff) >= 2 * 148) {
        auto kern2 = pw_dx_kernel<N, OT, 2, ROT>;
        static std::atomic<uint64_t> attr_done2{0};
        cudaError_t e2 = hol_ensure_max_smem(kern2, attr_done2);
        if (e2 != cudaSuccess) return (int)e2;
        dim3 grid2((unsigned)(s.Bp / 512), (unsigned)((s.I + it_eff - 1) / it_eff));
        kern2<<<grid2, 512, smem, st>>>(p);
        return (int)cudaGetLastError();
    }
    auto kern = pw_dx_kernel<N, OT, IT, ROT>;
    static std::atomic<uint64_t> attr_done{0};
    cudaError_t e = hol_ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return (int)e;
    // input groups are walked by the CTA; grid.y only has to fill the machine (~4 CTAs per SM)
    const int ngroups = (s.I + it_eff - 1) / it_eff;
    int64_t gy_ = (4 * 148 + s.Bp / BT - 1) / (s.Bp / BT);
    if (gy_ > ngroups || s.nOT > 1) gy_ = ngroups;  // wide layers: one group per CTA (measured faster at C2: 49.4 vs 56.5 ms)
    if (gy_ < 1) gy_ = 1;
    dim3 grid((unsigned)(s.Bp / BT), (unsigned)gy_);
    kern<<<grid, BT, smem, st>>>(p);
    return (int)cudaGetLastError();
}

template <int N>
static int launch_dx_n(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dx, int transposed,
                       cudaStream_t st) {
    if (s.rot) {
        switch (s.OT) {
            case 64: return launch_dx_t<N, 64, true>(s, ws, gy, dx, transposed, st);
            case 32: return launch_dx_t<N, 32, true>(s, ws, gy, dx, transposed, st);
        }
        return HOL_EUNSUPPORTED;
    }
    switch (s.OT) {
        case 64: return launch_dx_t<N, 64, false>(s, ws, gy, dx, transposed, st);
        case 32: return launch_dx_t<N, 32, false>(s, ws, gy, dx, transposed, st);
        case 16: return launch_dx_t<N, 16, false>(s, ws, gy, dx, transposed, st);
        case 8: return launch_dx_t<N, 8, false>(s, ws, gy, dx, transposed, st);
    }
    return HOL_EUNSUPPORTED;
}

int launch_backward_dx(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dx, int transposed,
                       cudaStream_t st) {
    switch (s.n) {
        case 2: return launch_dx_n<2>(s, ws, gy, dx, transposed, st);
        case 3: return launch_dx_n<3>(s, ws, gy, dx, transposed, st);
        case 4: return launch_dx_n<4>(s, ws, gy, dx, transposed, st);
        case 5: return launch_dx_n<5>(s, ws, gy, dx, transposed, st);
        case 6: return launch_dx_n<6>(s, ws, gy, dx, transposed, st);
        case 7: return launch_dx_n<7>(s, ws, gy, dx, transposed, st);
        case 8: return launch_dx_n<8>(s, ws, gy, dx, transposed, st);
        case 9: return launch_dx_n<9>(s, ws, gy, dx, transposed, st);
        case 10: return launch_dx_n<10>(s, ws, gy, dx, transposed, st);
    }
    return HOL_EUNSUPPORTED;  // n == 1 has dX == 0 and is handled by the caller (memset)
}
