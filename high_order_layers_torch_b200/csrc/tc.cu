// Tensor-core (tcgen05) path for layers with few segments: the segment-expanded dense contraction.
//
//   Phi[b, i*Kin + s(b,i)*stride + j] = phi_j(xl(b,i))        (n non-zeros out of Kin per (sample, input))
//   forward : y[b,o]        = sum_k Phi[b,k]  * W[o,k]         M = batch,  N = O,      K = I*Kin
//   dW      : dW[o,k]       = sum_b Phi[b,k]  * gy[b,o]        M = I*Kin,  N = O,      K = batch
//   dX      : T[b,k]        = sum_o gy[b,o]   * W[o,k]         M = batch,  N = I*Kin,  K = O
//             dX[b,i]       = chain(b,i) * sum_j phi'_j(xl) * T[b, i*Kin + s*stride + j]   (epilogue gather)
//
// The dense form spends Kin/n times the useful MACs -- but on the tensor pipe, which pays while
// Kin <= ~20 n (S <= 8 for the named configs): the gather kernels are pinned to one shared-memory
// word per FMA (25 % of the fp32 SIMT peak); the tensor pipe is ~20x that peak.
//
// fp32 accuracy from fp16 tensor cores: both operands are split a = hi + lo (two fp16 values, 22
// significant bits, power-of-two pre-scaling keeps them in the fp16 range) and the kernel issues
// hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator (dropped lo*lo term ~2^-22 relative).
// The tensor core accumulates with round-toward-zero: every chained MMA instruction loses on
// average half an ulp of the running sum, a bias that grows linearly with K (measured 3e-5 at
// K = 3136).  So a TMEM accumulator only ever holds <= 36 chained instructions (bias <= 2e-6); the
// epilogue warps drain it into round-to-nearest fp32 register sums while the MMA warp fills the
// other of the two TMEM accumulators.
//
// Operands are stored in global memory as ready-made shared-memory images (K-major, no-swizzle
// UMMA core-matrix layout: [k/8][row][8 x fp16]), so one 1-D TMA bulk copy per operand per
// pipeline stage fills the stage -- no tensor maps.  Warp roles: warp 0 = TMA producer, warp 1 =
// MMA issuer (one elected lane), warps 2-9 = epilogue (tcgen05.ld -> fp32 register sums -> global).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_A_IMG (TC_BM * TC_BK * 2)  // bytes of one fp16 A image (hi or lo)
#define TC_MAXHALF 128                // register sums per epilogue thread (BN <= 256)

enum { EPI_Y = 0, EPI_DW = 1, EPI_DX = 2 };

struct TcScalars {      // device scalars written by pw_absmax_kernel (bits of non-negative floats)
    unsigned max_xl, max_w, max_gy;
};

// power-of-two scale that maps `bound` just below 2^14 (scale >= 1 is allowed: small values are
// scaled UP so that their lo parts stay out of the fp16 subnormals)
__host__ __device__ __forceinline__ float pow2_scale_for(float bound) {
    if (!(bound > 0.0f) || !(bound < 3.0e38f)) return 1.0f;
    int e;
    frexpf(bound, &e);  // bound = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, 14 - e);
}
// |phi_j(t)| <= max_j |C_j| * prod_m (|t| + |X_m|)   (conservative: phi_j skips one of the factors)
__device__ __forceinline__ float phi_scale(float max_xl, const BasisTab& tab, int n) {
    float cmax = 0.0f, prod = 1.0f;
    for (int j = 0; j < n; ++j) {
        cmax = fmaxf(cmax, fabsf(tab.C[j]));
        prod *= (max_xl + fabsf(tab.X[j]));
    }
    const float bound = (n == 1) ? 1.0f : cmax * prod;
    return fminf(pow2_scale_for(fmaxf(bound, 1.0f)), 1.0f);  // only ever scale phi down
}

__global__ void __launch_bounds__(256) pw_absmax_kernel(const float* __restrict__ v, int64_t count, unsigned* out) {
    float m = 0.0f;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const float a = fabsf(v[k]);
        if (a < 3.0e38f) m = fmaxf(m, a);
    }
    for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order like uints
}

__device__ __forceinline__ void split_store8(const float (&v)[8], float scale, uint4* hi, uint4* lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a = v[2 * e] * scale, b = v[2 * e + 1] * scale;
        const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        const __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
        h[e] = __halves2half2(ha, hb);
        l[e] = __halves2half2(la, lb);
    }
    *hi = *reinterpret_cast<uint4*>(h);
    *lo = *reinterpret_cast<uint4*>(l);
}

// ---------------------------------------------------------------------------------------
// operand builders (HBM-bound).  Image layout of a tile with R rows: [k/8 (8)][row (R)][8 x fp16],
// hi image then lo image.
// ---------------------------------------------------------------------------------------

// forward A: Phi tiles [m-tile (128 samples)][k-chunk], k = i*Kin + v
template <int N>
__global__ void __launch_bounds__(256) pw_tc_expand_kernel(const float* __restrict__ xl_t, const uint16_t* __restrict__ seg_t,
                                                           unsigned char* __restrict__ tiles, const TcScalars* sc,
                                                           int64_t Bp, int64_t row0, int I, int Kin, int nkc, int stride,
                                                           const __grid_constant__ BasisTab tab) {
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int64_t mt = blockIdx.x;
    const int kc = blockIdx.y;
    const int64_t b = row0 + mt * TC_BM + r;
    const float scale = phi_scale(__uint_as_float(sc->max_xl), tab, N);
    unsigned char* base = tiles + ((mt * nkc + kc) * 2) * (int64_t)TC_A_IMG;
    int cur_i = -1;
    float phi[N];
    int a0 = 0;
    bool ok = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int kgl = half * 4 + q;
        const int k0 = kc * TC_BK + kgl * 8;
        const int i = k0 / Kin, v0 = k0 % Kin;
        if (i != cur_i) {
            cur_i = i;
            ok = false;
            if (i < I && b < Bp) {
                const uint32_t sg = seg_t[(int64_t)i * Bp + b];
                ok = !(sg & HOL_SEG_INVALID);
                a0 = (int)(sg & HOL_SEG_MASK) * stride;
                lagrange_basis<N>(xl_t[(int64_t)i * Bp + b], tab, phi);
            }
        }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = v0 + e - a0;
            float val = 0.0f;
#pragma unroll
            for (int jj = 0; jj < N; ++jj) val = (j == jj) ? phi[jj] : val;
            v[e] = ok ? val : 0.0f;
        }
        uint4* hi = reinterpret_cast<uint4*>(base + kgl * (TC_BM * 16) + r * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + TC_A_IMG + kgl * (TC_BM * 16) + r * 16);
        split_store8(v, scale, hi, lo);
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
                if (!(sg & HOL_SEG_INVALID) && j >= 0 && j < N) {
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
    for (int idx = threadIdx.x; idx < BN * 8; idx += blockDim.x) {
        const int row = idx % BN, kgl = idx / BN;
        const int o = nt * BN + row;
        const int k0 = kc * TC_BK + kgl * 8;
        const int i = k0 / Kin, v0 = k0 % Kin;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = v0 + e;
            v[e] = (o < O && i < I && col < nb) ? w[((int64_t)o * I + i) * nb + col] : 0.0f;
        }
        uint4* hi = reinterpret_cast<uint4*>(base + kgl * (BN * 16) + row * 16);
        uint4* lo = reinterpret_cast<uint4*>(base + img + kgl * (BN * 16) + row * 16);
        split_store8(v, scale, hi, lo);
    }
}

// dX B: W^T tiles [n-tile (BN rows k = i*Kin+v)][o-chunk (64 outputs)]
__global__ void __launch_bounds__(256) pw_tc_pack_wt_kernel(const float* __restrict__ w, unsigned char* __restrict__ tiles,
                                                            const TcScalars* sc, int I, int O, int nb, int Kin, int noc,
                                                            int BN) {
    const int nt = blockIdx.x, oc = blockIdx.y;
    const float scale = pow2_scale_for(__uint_as_float(sc->max_w));
    const int img = BN * TC_BK * 2;
    unsigned char* base = tiles + ((int64_t)(nt * noc + oc) * 2) * img;
    for (int idx = threadIdx.x; idx < BN * 8; idx += blockDim.x) {
        const int row = idx % BN, ogl = idx / BN;
        const int R = nt * BN + row;
        const int i = R / Kin, col = R % Kin;
        const int o0 = oc * TC_BK + ogl * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
            v[e] = (o0 + e < O && i < I && col < nb) ? w[((int64_t)(o0 + e) * I + i) * nb + col] : 0.0f;
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
    for (int idx = threadIdx.x; idx < BN * 8; idx += blockDim.x) {
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

// ---------------------------------------------------------------------------------------
// tcgen05 helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version=1 [46,48), layout NONE
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcGemmParams {
    const unsigned char* a_tiles;  // [m-tile][k-chunk][hi|lo] images of 128 rows
    const unsigned char* b_tiles;  // [n-tile][k-chunk][hi|lo] images of BN rows
    float* out;                    // EPI_Y: y[rows][ld]; EPI_DW: dw[O][I][nb]; EPI_DX: dx[B][I] or dx_t[I][Bp]
    const TcScalars* sc;
    const float* xl_t;             // EPI_DX
    const uint16_t* seg_t;         // EPI_DX
    int64_t row0, rows_valid, Bp;  // EPI_Y/DX: batch rows of this launch; EPI_DW: row0 > 0 => accumulate into out
    int BN, n_valid, ld, nkc, nst, n_basis, flush_chunks;
    int I, nb, Kin, stride, transposed;
    float length, Sf, inv_extra;
    BasisTab tab;
};

template <int EPI>
__global__ void __launch_bounds__(320, 1) pw_tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int BN = p.BN;
    const int B_IMG = BN * TC_BK * 2;
    const int STAGE_BYTES = 2 * TC_A_IMG + 2 * B_IMG;
    const int HALF = BN / 2;  // columns per epilogue thread
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [nst]
    uint64_t* empty = full + 8;                              // [nst]
    uint64_t* tmem_full = full + 16;                         // [2]
    uint64_t* tmem_empty = full + 18;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 24);
    unsigned char* stages = smem_raw + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = blockIdx.x;  // n-tile fastest: CTAs that share an A tile run together
    const int64_t mt = blockIdx.y;
    const int nst = p.nst;
    const int FG = p.flush_chunks;
    const int ngroups = (p.nkc + FG - 1) / FG;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;  // two accumulators; power of two >= 32

    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char* a_src = p.a_tiles + (mt * p.nkc) * (int64_t)(2 * TC_A_IMG);
            const unsigned char* b_src = p.b_tiles + ((int64_t)nt * p.nkc) * (int64_t)(2 * B_IMG);
            for (int kc = 0; kc < p.nkc; ++kc) {
                const int st = kc % nst;
                mbar_wait(&empty[st], (uint32_t)(((kc / nst) & 1) ^ 1));
                mbar_expect_tx(&full[st], (uint32_t)STAGE_BYTES);
                unsigned char* dst = stages + (size_t)st * STAGE_BYTES;
                bulk_g2s(dst, a_src + (int64_t)kc * (2 * TC_A_IMG), 2 * TC_A_IMG, &full[st]);
                bulk_g2s(dst + 2 * TC_A_IMG, b_src + (int64_t)kc * (2 * B_IMG), (uint32_t)(2 * B_IMG), &full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t LBO_A = TC_BM * 16, LBO_B = (uint32_t)BN * 16, SBO = 128;
            int g = 0;
            for (int kc = 0; kc < p.nkc; ++kc) {
                const int st = kc % nst;
                const int acc = g & 1;
                const bool first = (kc % FG) == 0;
                if (first) {
                    mbar_wait(&tmem_empty[acc], (uint32_t)(((g >> 1) & 1) ^ 1));  // drained by the epilogue
                    tc_fence_after();
                }
                mbar_wait(&full[st], (uint32_t)((kc / nst) & 1));
                tc_fence_after();
                const uint32_t sa = smem_u32(stages + (size_t)st * STAGE_BYTES);
                const uint32_t sb = sa + 2 * TC_A_IMG;
                const uint32_t td = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 16; ++kk) {
                    const uint64_t a_hi = umma_desc_kmajor_noswizzle(sa + kk * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t a_lo = umma_desc_kmajor_noswizzle(sa + TC_A_IMG + kk * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t b_hi = umma_desc_kmajor_noswizzle(sb + kk * 2 * LBO_B, LBO_B, SBO);
                    const uint64_t b_lo = umma_desc_kmajor_noswizzle(sb + B_IMG + kk * 2 * LBO_B, LBO_B, SBO);
                    umma_f16(td, a_hi, b_hi, idesc, (first && kk == 0) ? 0u : 1u);
                    umma_f16(td, a_hi, b_lo, idesc, 1u);
                    umma_f16(td, a_lo, b_hi, idesc, 1u);
                }
                umma_commit(&empty[st]);  // arrives when the MMAs above have finished reading this stage
                if ((kc % FG) == FG - 1 || kc == p.nkc - 1) {
                    umma_commit(&tmem_full[acc]);
                    ++g;
                }
            }
        }
    } else {
        // epilogue warps 2..9: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        float racc[TC_MAXHALF];
#pragma unroll
        for (int e = 0; e < TC_MAXHALF; ++e) racc[e] = 0.0f;
        for (int g = 0; g < ngroups; ++g) {
            const int acc = g & 1;
            mbar_wait(&tmem_full[acc], (uint32_t)((g >> 1) & 1));
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + h * HALF);
#pragma unroll
            for (int c0 = 0; c0 < TC_MAXHALF; c0 += 32) {
                if (c0 < HALF) {
                    uint32_t v[32];
                    tmem_ld32(t0 + (uint32_t)c0, v);
#pragma unroll
                    for (int e = 0; e < 32; ++e) racc[c0 + e] += __uint_as_float(v[e]);  // columns >= HALF: unused
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int rl = q * 32 + lane;             // row inside the 128-row tile
        const int64_t row = mt * TC_BM + rl;      // row inside this launch
        const int colbase = nt * BN + h * HALF;

        if (EPI == EPI_Y) {
            const float inv = 1.0f / (phi_scale(__uint_as_float(p.sc->max_xl), p.tab, p.n_basis) *
                                      pow2_scale_for(__uint_as_float(p.sc->max_w)));
            if (row < p.rows_valid) {
                float* yrow = p.out + (p.row0 + row) * (int64_t)p.ld + colbase;
                const bool vec = (p.ld & 3) == 0 && (HALF & 3) == 0;
#pragma unroll
                for (int e = 0; e < TC_MAXHALF; e += 4) {
                    if (e < HALF) {
                        const int col = colbase + e;
                        const float4 o4 = make_float4(racc[e] * inv, racc[e + 1] * inv, racc[e + 2] * inv, racc[e + 3] * inv);
                        if (vec && col + 3 < p.n_valid) {
                            *reinterpret_cast<float4*>(yrow + e) = o4;
                        } else {
                            if (col + 0 < p.n_valid) yrow[e + 0] = o4.x;
                            if (col + 1 < p.n_valid) yrow[e + 1] = o4.y;
                            if (col + 2 < p.n_valid) yrow[e + 2] = o4.z;
                            if (col + 3 < p.n_valid) yrow[e + 3] = o4.w;
                        }
                    }
                }
            }
        } else if (EPI == EPI_DW) {
            // row = dense column k = i*Kin + v, columns = outputs; write the reference layout dw[o][i][v]
            const float inv = 1.0f / (phi_scale(__uint_as_float(p.sc->max_xl), p.tab, p.n_basis) *
                                      pow2_scale_for(__uint_as_float(p.sc->max_gy)));
            const int i = (int)(row / p.Kin), v = (int)(row % p.Kin);
            if (i < p.I && v < p.nb) {
                const bool accumulate = p.row0 > 0;
#pragma unroll
                for (int e = 0; e < TC_MAXHALF; ++e) {
                    if (e < HALF) {
                        const int o = colbase + e;
                        if (o < p.n_valid) {
                            float* dst = p.out + ((int64_t)o * p.I + i) * p.nb + v;
                            *dst = accumulate ? *dst + racc[e] * inv : racc[e] * inv;
                        }
                    }
                }
            }
        } else {
            // EPI_DX: stage T[128][BN] in the (now idle) pipeline buffers, then every thread gathers the
            // n active columns of each of its inputs.  After the last tmem_full wait all MMAs have
            // completed, so the stages are no longer read.
            float* T = reinterpret_cast<float*>(stages);
            const int pitch = BN + 1;
#pragma unroll
            for (int e = 0; e < TC_MAXHALF; ++e)
                if (e < HALF) T[rl * pitch + h * HALF + e] = racc[e];
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps
            const float inv = p.inv_extra / (pow2_scale_for(__uint_as_float(p.sc->max_gy)) *
                                             pow2_scale_for(__uint_as_float(p.sc->max_w)));
            const int64_t b = p.row0 + row;
            const int ninp = HALF / p.Kin;
            const int n = p.n_basis;
            for (int ii = 0; ii < ninp; ++ii) {
                const int i = colbase / p.Kin + ii;
                if (i >= p.I || b >= p.Bp) continue;
                const int64_t idx = (int64_t)i * p.Bp + b;
                const uint32_t sg = p.seg_t[idx];
                const float t = p.xl_t[idx];
                const int s = (int)(sg & HOL_SEG_MASK);
                const float* Trow = T + rl * pitch + h * HALF + ii * p.Kin + s * p.stride;
                // d phi_j / dt by the product-rule recurrence (see lagrange_basis_derivative)
                float sum = 0.0f;
                for (int j = 0; j < n; ++j) {
                    float P = 1.0f, Ssum = 0.0f;
                    for (int m = 0; m < n; ++m) {
                        if (m == j) continue;
                        const float d = t - p.tab.X[m];
                        Ssum = fmaf(Ssum, d, P);
                        P *= d;
                    }
                    sum = fmaf(Ssum * p.tab.C[j], Trow[j], sum);
                }
                const float e0 = __fsub_rn(__fmul_rn(__fdiv_rn((float)s, p.Sf), 2.0f), 1.0f);
                const float e1 = __fsub_rn(__fmul_rn(__fdiv_rn((float)(s + 1), p.Sf), 2.0f), 1.0f);
                float chain = __fdiv_rn(p.length, __fsub_rn(e1, e0));
                if (sg & HOL_SEG_MIRROR) chain = -chain;
                const float outv = (sg & HOL_SEG_INVALID) ? 0.0f : sum * inv * chain;
                if (p.transposed)
                    p.out[idx] = outv;
                else if (row < p.rows_valid)
                    p.out[b * p.I + i] = outv;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// planning + launch
// ---------------------------------------------------------------------------------------
static int tc_bn_for(int O) { return O > 128 ? 256 : (O > 64 ? 128 : 64); }

bool tc_plan(const PwShape& s, TcPlan& t) {
    t.enabled = false;
    const char* off = getenv("HOL_DISABLE_TC");
    if (off && off[0] == '1') return false;
    t.Kin = (s.nb + 7) / 8 * 8;
    // the dense form costs 3*Kin tensor MACs per n useful fp32 FMAs
    if (t.Kin > 20 * s.n || t.Kin > 128) return false;
    if (s.O < 32 || s.I * (int64_t)t.Kin < 256) return false;  // too narrow to be worth a tile
    t.BN = tc_bn_for(s.O);
    t.nNT = (s.O + t.BN - 1) / t.BN;
    const int64_t K = (int64_t)s.I * t.Kin;
    t.nkc = (int)((K + TC_BK - 1) / TC_BK);
    // dX: N tile = whole inputs per epilogue half: a multiple of 2*Kin (and of 16), <= 256
    t.BNx = (256 / (2 * t.Kin)) * (2 * t.Kin);
    t.nNTx = (int)((K + t.BNx - 1) / t.BNx);
    t.noc = (s.O + TC_BK - 1) / TC_BK;
    // dW: M tiles over the dense columns
    t.nMTw = (int)((K + TC_BM - 1) / TC_BM);
    // rows per launch: bound the Phi tiles to ~8 GiB
    const int64_t bytes_per_row = (int64_t)t.nkc * TC_BK * 2 * 2;
    int64_t rows = (int64_t)(8ll << 30) / bytes_per_row / 512 * 512;
    if (rows < 512) return false;
    if (rows > s.Bp) rows = s.Bp;
    rows = (rows + TC_BM - 1) / TC_BM * TC_BM;
    t.rows_per_launch = rows;
    const size_t mtiles = (size_t)(rows / TC_BM);
    const size_t bchunks = (size_t)((rows + TC_BK - 1) / TC_BK);
    size_t a = mtiles * t.nkc * 2 * TC_A_IMG;                                     // Phi
    const size_t a_t = (size_t)t.nMTw * bchunks * 2 * TC_A_IMG;                   // Phi^T
    const size_t a_gy = mtiles * t.noc * 2 * TC_A_IMG;                            // gy
    if (a_t > a) a = a_t;
    if (a_gy > a) a = a_gy;
    size_t bb = (size_t)t.nNT * t.nkc * 2 * ((size_t)t.BN * TC_BK * 2);           // W
    const size_t b_gyt = (size_t)t.nNT * bchunks * 2 * ((size_t)t.BN * TC_BK * 2);  // gy^T
    const size_t b_wt = (size_t)t.nNTx * t.noc * 2 * ((size_t)t.BNx * TC_BK * 2);   // W^T
    if (b_gyt > bb) bb = b_gyt;
    if (b_wt > bb) bb = b_wt;
    t.a_bytes = a;
    t.b_bytes = bb;
    t.enabled = true;
    return true;
}

static void fill_common(TcGemmParams& p, const PwShape& s, const PwWorkspace& ws, const TcPlan& t) {
    p.a_tiles = (const unsigned char*)ws.tc_a;
    p.b_tiles = (const unsigned char*)ws.tc_b;
    p.sc = (const TcScalars*)ws.tc_scalars;
    p.xl_t = ws.xl_t;
    p.seg_t = ws.seg_t;
    p.Bp = s.Bp;
    p.n_basis = s.n;
    p.I = s.I;
    p.nb = s.nb;
    p.Kin = t.Kin;
    p.stride = s.stride;
    p.transposed = 0;
    p.length = s.length;
    p.Sf = (float)s.S;
    p.inv_extra = 1.0f;
    p.tab = s.tab;
}

template <int EPI>
static int launch_gemm(TcGemmParams p, int64_t mtiles, int ntiles, cudaStream_t st) {
    const int stage = 2 * TC_A_IMG + 2 * p.BN * TC_BK * 2;
    int nst = (HOL_SMEM_BUDGET - 1024) / stage;
    if (nst > 6) nst = 6;
    if (nst < 2) return HOL_EUNSUPPORTED;
    size_t smem = 1024 + (size_t)nst * stage;
    if (EPI == EPI_DX) {  // the epilogue stages a 128 x (BN+1) fp32 tile in the pipeline buffers
        const size_t need = 1024 + (size_t)TC_BM * (p.BN + 1) * 4;
        if (need > smem) smem = need;
        if (smem > HOL_SMEM_BUDGET) return HOL_EUNSUPPORTED;
    }
    p.nst = nst;
    p.flush_chunks = p.BN >= 128 ? 2 : 3;  // <= 36 chained MMA instructions per TMEM accumulator
    auto kern = pw_tc_gemm_kernel<EPI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ntiles, (unsigned)mtiles);
    kern<<<grid, 320, smem, st>>>(p);
    return (int)cudaGetLastError();
}

#define DISPATCH_N(n, CALL)                       \
    switch (n) {                                  \
        case 1: rc = CALL(1); break;              \
        case 2: rc = CALL(2); break;              \
        case 3: rc = CALL(3); break;              \
        case 4: rc = CALL(4); break;              \
        case 5: rc = CALL(5); break;              \
        case 6: rc = CALL(6); break;              \
        case 7: rc = CALL(7); break;              \
        case 8: rc = CALL(8); break;              \
        case 9: rc = CALL(9); break;              \
        case 10: rc = CALL(10); break;            \
        default: rc = HOL_EUNSUPPORTED;           \
    }

static int tc_absmax(const float* v, int64_t count, unsigned* out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return (int)e;
    pw_absmax_kernel<<<148 * 4, 256, 0, st>>>(v, count, out);
    return (int)cudaGetLastError();
}

int launch_forward_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, float* y, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = tc_absmax(ws.xl_t, (int64_t)s.I * s.Bp, &sc->max_xl, st);
    if (rc != HOL_OK) return rc;
    rc = tc_absmax(w, (int64_t)s.O * s.I * s.nb, &sc->max_w, st);
    if (rc != HOL_OK) return rc;
    dim3 pgrid((unsigned)t.nNT, (unsigned)t.nkc);
    pw_tc_pack_w_kernel<<<pgrid, 256, 0, st>>>(w, (unsigned char*)ws.tc_b, sc, s.I, s.O, s.nb, t.Kin, t.nkc, t.BN);
    rc = (int)cudaGetLastError();
    if (rc != HOL_OK) return rc;
    for (int64_t row0 = 0; row0 < s.Bp; row0 += t.rows_per_launch) {
        int64_t rows = s.Bp - row0 < t.rows_per_launch ? s.Bp - row0 : t.rows_per_launch;
        rows = (rows + TC_BM - 1) / TC_BM * TC_BM;
        dim3 egrid((unsigned)(rows / TC_BM), (unsigned)t.nkc);
#define EXPAND(NN) (pw_tc_expand_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, (unsigned char*)ws.tc_a, sc, s.Bp, row0, s.I, t.Kin, t.nkc, s.stride, s.tab), (int)cudaGetLastError())
        DISPATCH_N(s.n, EXPAND)
#undef EXPAND
        if (rc != HOL_OK) return rc;
        TcGemmParams p;
        fill_common(p, s, ws, t);
        p.out = y;
        p.row0 = row0;
        p.rows_valid = s.B - row0 < rows ? s.B - row0 : rows;
        p.BN = t.BN;
        p.n_valid = s.O;
        p.ld = s.O;
        p.nkc = t.nkc;
        rc = launch_gemm<EPI_Y>(p, rows / TC_BM, t.nNT, st);
        if (rc != HOL_OK) return rc;
    }
    return HOL_OK;
}

int launch_backward_dw_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* gy, float* dw,
                          cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = tc_absmax(ws.xl_t, (int64_t)s.I * s.Bp, &sc->max_xl, st);
    if (rc != HOL_OK) return rc;
    rc = tc_absmax(gy, s.B * (int64_t)s.O, &sc->max_gy, st);
    if (rc != HOL_OK) return rc;
    for (int64_t row0 = 0; row0 < s.Bp; row0 += t.rows_per_launch) {
        int64_t rows = s.Bp - row0 < t.rows_per_launch ? s.Bp - row0 : t.rows_per_launch;
        const int nbc = (int)((rows + TC_BK - 1) / TC_BK);
        dim3 egrid((unsigned)t.nMTw, (unsigned)nbc);
#define EXPANDT(NN) (pw_tc_expand_t_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, (unsigned char*)ws.tc_a, sc, s.Bp, row0, s.I, t.Kin, nbc, s.stride, s.tab), (int)cudaGetLastError())
        DISPATCH_N(s.n, EXPANDT)
#undef EXPANDT
        if (rc != HOL_OK) return rc;
        dim3 ggrid((unsigned)t.nNT, (unsigned)nbc);
        pw_tc_pack_gyt_kernel<<<ggrid, 256, 0, st>>>(gy, (unsigned char*)ws.tc_b, sc, s.B, row0, s.O, nbc, t.BN);
        rc = (int)cudaGetLastError();
        if (rc != HOL_OK) return rc;
        TcGemmParams p;
        fill_common(p, s, ws, t);
        p.out = dw;
        p.row0 = row0;  // > 0: accumulate onto the previous batch chunk's result
        p.rows_valid = 0;
        p.BN = t.BN;
        p.n_valid = s.O;
        p.ld = 0;
        p.nkc = nbc;
        rc = launch_gemm<EPI_DW>(p, t.nMTw, t.nNT, st);
        if (rc != HOL_OK) return rc;
    }
    return HOL_OK;
}

int launch_backward_dx_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, const float* gy,
                          float* dx, int transposed, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = tc_absmax(w, (int64_t)s.O * s.I * s.nb, &sc->max_w, st);
    if (rc != HOL_OK) return rc;
    rc = tc_absmax(gy, s.B * (int64_t)s.O, &sc->max_gy, st);
    if (rc != HOL_OK) return rc;
    dim3 pgrid((unsigned)t.nNTx, (unsigned)t.noc);
    pw_tc_pack_wt_kernel<<<pgrid, 256, 0, st>>>(w, (unsigned char*)ws.tc_b, sc, s.I, s.O, s.nb, t.Kin, t.noc, t.BNx);
    rc = (int)cudaGetLastError();
    if (rc != HOL_OK) return rc;
    for (int64_t row0 = 0; row0 < s.Bp; row0 += t.rows_per_launch) {
        int64_t rows = s.Bp - row0 < t.rows_per_launch ? s.Bp - row0 : t.rows_per_launch;
        rows = (rows + TC_BM - 1) / TC_BM * TC_BM;
        dim3 ggrid((unsigned)(rows / TC_BM), (unsigned)t.noc);
        pw_tc_pack_gy_kernel<<<ggrid, 256, 0, st>>>(gy, (unsigned char*)ws.tc_a, sc, s.B, row0, s.O, t.noc);
        rc = (int)cudaGetLastError();
        if (rc != HOL_OK) return rc;
        TcGemmParams p;
        fill_common(p, s, ws, t);
        p.out = dx;
        p.row0 = row0;
        p.rows_valid = s.B - row0 < rows ? s.B - row0 : rows;
        p.BN = t.BNx;
        p.n_valid = 0;
        p.ld = 0;
        p.nkc = t.noc;
        p.transposed = transposed;
        rc = launch_gemm<EPI_DX>(p, rows / TC_BM, t.nNTx, st);
        if (rc != HOL_OK) return rc;
    }
    return HOL_OK;
}
