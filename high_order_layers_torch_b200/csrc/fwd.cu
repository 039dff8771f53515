// Fused forward contraction  y[b,o] = sum_i sum_j phi_j(xl(b,i)) * Wk[i][row(j, s(b,i))][o].
//
// Mapping (sm_100a): one thread = one sample, OT output accumulators in registers (packed
// fp32x2 FMAs), one CTA = BT samples x one output tile.  The packed weights of IC inputs are
// staged in shared memory by the TMA engine (1-D bulk copies, double buffered, mbarrier
// completion).  The lanes of a warp read the SAME basis row j and column chunk of DIFFERENT
// segments: rows of neighbouring segments are adjacent with a pitch of OT+4 floats, so the (up
// to 8) distinct 16-byte chunks a warp touches per LDS.128 sit in distinct banks and samples
// that share a segment are served by the shared-memory broadcast.
//
// More than 8 segments cannot be separated by the pitch: rows 8 segments apart share their
// banks and a warp with random segments pays ~2.5 wavefronts too many per LDS.128 (measured:
// the forward ran at 6 TFLOP/s at S = 64 against 16 at S = 8).  For S > 8 (ROT) the row pitch is a
// multiple of 128 bytes and lane l reads the column chunks in the rotated order (q + l%8) mod
// (OT/4): the 8 lanes of every quarter warp then hit 8 different 16-byte bank groups whatever
// their segments are -- always the 4-wavefront minimum -- and the rotation is undone for free when
// the outputs are stored (slot q of lane l is output chunk (q + l%8) mod (OT/4)).
#include "kernels.cuh"

struct FwdParams {
    const float* xl_t;
    const uint16_t* seg_t;
    const float* wk;
    float* y;
    int64_t B, Bp;
    int I, O, R, IC, stage_floats, srow, nst;
    int jrow[HOL_MAX_N];
    BasisTab tab;
};

template <int N, int OT, bool ROT>
__device__ __forceinline__ void fwd_one(float xl, uint32_t sg, const float* slab_i, float2 (&acc)[OT / 2],
                                        const FwdParams& p, int rot) {
    float phi[N];
    lagrange_basis<N>(xl, p.tab, phi);
    const float keep = (sg & HOL_SEG_INVALID) ? 0.0f : 1.0f;
    const float* r0 = slab_i + (sg & HOL_SEG_MASK) * p.srow;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float ph = phi[j] * keep;
        const float4* rj = reinterpret_cast<const float4*>(r0 + p.jrow[j]);
#pragma unroll
        for (int q = 0; q < OT / 4; ++q) {
            const float4 w = rj[ROT ? ((q + rot) & (OT / 4 - 1)) : q];
            acc[2 * q] = ffma2(ph, make_float2(w.x, w.y), acc[2 * q]);
            acc[2 * q + 1] = ffma2(ph, make_float2(w.z, w.w), acc[2 * q + 1]);
        }
    }
}

// IS > 1 (narrow layers: a single tile of <= 16 outputs and too few samples to fill the GPU with one
// thread per sample): IS consecutive lanes share a sample, take the input groups of a chunk round
// robin and combine their partial sums with a butterfly at the end -- IS times the threads, 1/IS of
// the serial chain per thread (the 128 -> 10 classifier head of config 3 ran 40 us on 64 CTAs of 128
// threads, pure latency).
template <int N, int OT, bool ROT, int IS = 1>
__global__ void __launch_bounds__(512, 1) pw_fwd_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float* slab0 = reinterpret_cast<float*>(smem_raw + 128);
    constexpr int OTP = ROT ? OT : OT + 4;
    // coordinate prefetch depth: with 64 accumulators under the 128-register cap of a 512-thread CTA a depth of 4
    // spills inside the input loop for n <= 6 (n = 4, S = 64 of the config-5 sweep: 76 STL/LDL, 69 GB of
    // spill traffic per forward at batch 8192; depth 2: 371 -> 297 ms); the larger bases spill less at depth 4
    constexpr int G = (OT >= 64 && N <= 6) ? 2 : 4;
    const int tid = threadIdx.x;
    const int q = IS > 1 ? tid % IS : 0;                 // which share of the input groups
    const int rot = (tid / IS) & 7;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + tid) / IS;
    const int ot = blockIdx.y;
    const int in_floats = p.R * OTP;
    const float* wk_ot = p.wk + (int64_t)ot * p.I * in_floats;
    const int nchunks = (p.I + p.IC - 1) / p.IC;

    if (tid == 0) {
        for (int k = 0; k < p.nst; ++k) mbar_init(&bars[k], 1);  // nst <= 16: the barriers own the first 128 bytes
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        const int i0 = c * p.IC;
        const int ni = min(p.IC, p.I - i0);
        const uint32_t bytes = (uint32_t)ni * (uint32_t)in_floats * 4u;
        const int stg = c % p.nst;
        uint64_t* bar = &bars[stg];
        mbar_expect_tx(bar, bytes);
        bulk_g2s(slab0 + (size_t)stg * p.stage_floats, wk_ot + (int64_t)i0 * in_floats, bytes, bar);
    };
    if (tid == 0)
        for (int c = 0; c < p.nst && c < nchunks; ++c) issue(c);

    float2 acc[OT / 2];
#pragma unroll
    for (int k = 0; k < OT / 2; ++k) acc[k] = make_float2(0.0f, 0.0f);

    const float* xlp = p.xl_t + b;
    const uint16_t* sgp = p.seg_t + b;
    // coordinates of this thread's next input group, loaded one group ahead of use
    float xr[G];
    uint32_t sr[G];
    auto fetch = [&](int base) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const int ii = base + k;
            xr[k] = 0.0f;
            sr[k] = HOL_SEG_INVALID;
            if (ii < p.I) {
                xr[k] = __ldg(xlp + (int64_t)ii * p.Bp);
                sr[k] = __ldg(sgp + (int64_t)ii * p.Bp);
            }
        }
    };
    fetch(q * G);

    for (int c = 0; c < nchunks; ++c) {
        const int i0 = c * p.IC;
        const int ni = min(p.IC, p.I - i0);
        mbar_wait(&bars[c % p.nst], (uint32_t)((c / p.nst) & 1));
        const float* slab = slab0 + (size_t)(c % p.nst) * p.stage_floats;
        for (int g0 = q * G; g0 < ni; g0 += G * IS) {
            // the registers hold the coordinates of inputs i0 + g0 .. +G-1; after an input is consumed its slot is
            // refilled with the same slot of this thread's next group (possibly in the next chunk)
            const int nbase = (g0 + G * IS < ni) ? i0 + g0 + G * IS : i0 + ni + q * G;
#pragma unroll
            for (int k = 0; k < G; ++k) {
                if (g0 + k < ni) fwd_one<N, OT, ROT>(xr[k], sr[k], slab + (size_t)(g0 + k) * in_floats, acc, p, rot);
                const int ii = nbase + k;
                xr[k] = 0.0f;
                sr[k] = HOL_SEG_INVALID;
                // a group of the next chunk is only this thread's if it starts inside that chunk
                if (ii < p.I && (IS == 1 || nbase < i0 + ni || q * G < min(p.IC, p.I - (i0 + ni)))) {
                    xr[k] = __ldg(xlp + (int64_t)ii * p.Bp);
                    sr[k] = __ldg(sgp + (int64_t)ii * p.Bp);
                }
            }
        }
        if (IS > 1 && q * G >= ni) {
            // no group of this chunk was this thread's (short last chunk): its registers may hold the head of a
            // chunk it does not own either -- nothing to do, they are refetched when a group becomes due
            if (c + 1 < nchunks && q * G < min(p.IC, p.I - (i0 + ni))) fetch(i0 + ni + q * G);
        }
        __syncthreads();  // everyone is done with this stage before it is refilled
        if (tid == 0 && c + p.nst < nchunks) issue(c + p.nst);
    }

    if (IS > 1) {  // combine the IS partial sums of a sample (adjacent lanes)
#pragma unroll
        for (int d = 1; d < IS; d <<= 1) {
#pragma unroll
            for (int k = 0; k < OT / 2; ++k) {
                acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, d);
                acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, d);
            }
        }
    }
    if (b < p.B && q == 0) {
        float* yrow = p.y + b * p.O + (int64_t)ot * OT;
        const int o0 = ot * OT;
        const bool vec = (p.O & 3) == 0;
#pragma unroll
        for (int k = 0; k < OT / 4; ++k) {
            const int qc = ROT ? ((k + rot) & (OT / 4 - 1)) : k;  // output chunk held in slot k
            const int o = o0 + 4 * qc;
            if (vec && o + 3 < p.O) {
                *reinterpret_cast<float4*>(yrow + 4 * qc) =
                    make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
            } else {
                if (o + 0 < p.O) yrow[4 * qc + 0] = acc[2 * k].x;
                if (o + 1 < p.O) yrow[4 * qc + 1] = acc[2 * k].y;
                if (o + 2 < p.O) yrow[4 * qc + 2] = acc[2 * k + 1].x;
                if (o + 3 < p.O) yrow[4 * qc + 3] = acc[2 * k + 1].y;
            }
        }
    }
}

template <int N, int OT, bool ROT>
static int launch_fwd_t(const PwShape& s, const PwWorkspace& ws, float* y, cudaStream_t st) {
    constexpr int OTP = ROT ? OT : OT + 4;
    if (s.OTP != OTP) return HOL_EINVAL;
    FwdParams p;
    p.xl_t = ws.xl_t;
    p.seg_t = ws.seg_t;
    p.wk = ws.wk;
    p.y = y;
    p.B = s.B;
    p.Bp = s.Bp;
    p.I = s.I;
    p.O = s.O;
    p.R = s.R;
    p.srow = s.srow;
    for (int j = 0; j < HOL_MAX_N; ++j) p.jrow[j] = s.jrow[j];
    p.tab = s.tab;
    const size_t in_bytes = (size_t)s.R * OTP * 4;
    int nst = 2;
    int IC = (int)((HOL_SMEM_BUDGET - 128) / 2 / in_bytes);
    if (IC < 1) {
        nst = 1;
        IC = (int)((HOL_SMEM_BUDGET - 128) / in_bytes);
    }
    if (IC < 1) return HOL_EUNSUPPORTED;
    if (IC > s.I) IC = s.I;
    if (IC > 16) IC = 16;
    if (IC >= 4) IC &= ~3;
    p.IC = IC;
    p.stage_floats = IC * s.R * OTP;
    p.nst = nst;
    const size_t smem = 128 + (size_t)nst * p.stage_floats * 4;
    // block = samples per CTA: the largest tile that still gives >= 2 waves on 148 SMs
    int BT;
    if (s.Bp % 512 != 0) {
        BT = (int)s.Bp;  // small batch: Bp <= 256, multiple of 32
    } else {
        BT = 512;
        while (BT > 128 && (s.Bp / BT) * s.nOT < 2 * 148) BT >>= 1;
    }
    if constexpr (OT <= 16 && !ROT) {
        // A single narrow tile and at most one 128-sample CTA per SM (the 128 -> 10 classifier head of config 3 at
        // batch 8192: 64 CTAs).  The work is tiny, so what counts is (a) using every SM -- each (sample, input)
        // pair reads n*OTP words of shared memory, 268 MB in all, which on 64 SMs alone is 17 us -- and (b) not
        // paying one L2 round trip per chunk of weights: 32-sample CTAs with four lanes per sample, and every
        // chunk of the layer's weights (they fit shared memory) requested up front.
        if (s.nOT == 1 && BT == 128 && s.Bp % 128 == 0 && (s.Bp / BT) <= 148 && s.I >= 16) {
            auto kern4 = pw_fwd_kernel<N, OT, ROT, 4>;
            static std::atomic<uint64_t> attr_done4{0};
            cudaError_t e4 = hol_ensure_max_smem(kern4, attr_done4);
            if (e4 != cudaSuccess) return (int)e4;
            int bt4 = 128;
            while (bt4 > 32 && s.Bp / bt4 < 2 * 148) bt4 >>= 1;
            const int nchunks = (s.I + IC - 1) / IC;
            int nst4 = (int)((HOL_SMEM_BUDGET - 128) / ((size_t)p.stage_floats * 4));
            if (nst4 > nchunks) nst4 = nchunks;
            if (nst4 > 16) nst4 = 16;
            if (nst4 < nst) nst4 = nst;
            p.nst = nst4;
            const size_t smem4 = 128 + (size_t)nst4 * p.stage_floats * 4;
            dim3 grid4((unsigned)(s.Bp / bt4), 1u);
            kern4<<<grid4, bt4 * 4, smem4, st>>>(p);
            return (int)cudaGetLastError();
        }
    }
    auto kern = pw_fwd_kernel<N, OT, ROT>;
    static std::atomic<uint64_t> attr_done{0};
    cudaError_t e = hol_ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)(s.Bp / BT), (unsigned)s.nOT);
    kern<<<grid, BT, smem, st>>>(p);
    return (int)cudaGetLastError();
}

template <int N>
static int launch_fwd_n(const PwShape& s, const PwWorkspace& ws, float* y, cudaStream_t st) {
    if (s.rot) {  // more than 8 segments: lane-rotated column order (OT >= 32 only, see make_shape)
        switch (s.OT) {
            case 64: return launch_fwd_t<N, 64, true>(s, ws, y, st);
            case 32: return launch_fwd_t<N, 32, true>(s, ws, y, st);
        }
        return HOL_EUNSUPPORTED;
    }
    switch (s.OT) {
        case 64: return launch_fwd_t<N, 64, false>(s, ws, y, st);
        case 32: return launch_fwd_t<N, 32, false>(s, ws, y, st);
        case 16: return launch_fwd_t<N, 16, false>(s, ws, y, st);
        case 8: return launch_fwd_t<N, 8, false>(s, ws, y, st);
    }
    return HOL_EUNSUPPORTED;
}

int launch_forward(const PwShape& s, const PwWorkspace& ws, float* y, cudaStream_t st) {
    switch (s.n) {
        case 1: return launch_fwd_n<1>(s, ws, y, st);
        case 2: return launch_fwd_n<2>(s, ws, y, st);
        case 3: return launch_fwd_n<3>(s, ws, y, st);
        case 4: return launch_fwd_n<4>(s, ws, y, st);
        case 5: return launch_fwd_n<5>(s, ws, y, st);
        case 6: return launch_fwd_n<6>(s, ws, y, st);
        case 7: return launch_fwd_n<7>(s, ws, y, st);
        case 8: return launch_fwd_n<8>(s, ws, y, st);
        case 9: return launch_fwd_n<9>(s, ws, y, st);
        case 10: return launch_fwd_n<10>(s, ws, y, st);
    }
    return HOL_EUNSUPPORTED;
}
