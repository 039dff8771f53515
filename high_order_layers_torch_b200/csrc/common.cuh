// Shared definitions for the sm_100a piecewise-polynomial kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hol_b200.h"

// seg_t entries (uint16): segment index in the low 12 bits (S <= 4096) plus three flags.
#define HOL_SEG_MASK 0x0fffu
#define HOL_SEG_OUTLIER 0x2000u  // |xl| beyond the range the split-fp16 tensor-core operands can carry: the
                                 // tensor-core passes skip the element, an fp32 fix-up kernel adds it (tc_fixup.cu)
#define HOL_SEG_INVALID 0x4000u  // padding row / zero-padded conv tap: contributes nothing
#define HOL_SEG_MIRROR 0x8000u   // make_periodic took the mirrored branch: d fold / d x = -1

// diagnostic plan overrides (hol_set_plan_overrides): tests and A/B runs force a kernel family
#define HOL_OVR_DISABLE_TC HOL_PLAN_DISABLE_TENSOR_CORES
#define HOL_OVR_DISABLE_DWD HOL_PLAN_DISABLE_DENSE_DW
#define HOL_OVR_NO_SHARE_PHI HOL_PLAN_NO_SHARED_PHI
#define HOL_OVR_FLUSH_SHIFT HOL_PLAN_FLUSH_SHIFT
#define HOL_OVR_NO_CLUSTER HOL_PLAN_NO_CLUSTER

#define HOL_WARPS_SORT 8
#define HOL_SORT_CHUNK 8192  // samples per bucket-sort CTA / dW split unit

struct BasisTab {
    float X[HOL_MAX_N];  // interpolation nodes (reference fp32 table)
    float C[HOL_MAX_N];  // C[j] = 1 / prod_{m != j} (X_j - X_m)
};

// Plan of the tensor-core (tcgen05) passes for layers with few segments (tc_gemm.cu).
struct TcPlan {
    bool enabled;             // any pass on the tensor cores
    bool fwd, dx, dw;         // per pass (thresholds fitted to the c5 sweep, see tc_plan)
    bool share_phi;           // Phi tiles of the WHOLE batch are built once by the forward and read again by dW
                              // (MN-major A operand), so Phi^T is never built
    int nkc_pad;              // share_phi: k-chunks per batch tile in the Phi storage (nkc rounded up to even)
    size_t phi_bytes;         // share_phi: Phi storage for the whole batch
    int Kin;                  // dense columns per input in Phi / W / dW: exactly the weights per link (inputs are packed
                              // back to back, tiles may straddle them)
    int Kinx;                 // ... in the dX GEMM's N dimension: padded to 8 (its epilogue keeps whole inputs per thread)
    int BN, nNT;              // forward: output tile width and count
    int BNw, nNTw;            // dW: output tile width and count
    int nkc;                  // 64-wide K chunks
    int BNx, nNTx, noc;       // dX: N tile over the dense columns (multiple of 2*Kinx), its count, 64-wide output chunks
    int nMTw;                 // dW: 128-row M tiles over the dense columns
    int dw_ksplit;            // dW: the batch (K) is split over this many CTAs per tile (partials -> dw_part)
    int fwd_ksplit;           // forward: the dense columns (K) are split likewise when the batch gives few tiles (partials -> y_part)
    int flush_chunks;         // k-chunks accumulated in TMEM between drains
    int64_t rows_per_launch;  // batch rows expanded + multiplied per launch (bounds the Phi tiles)
    size_t a_bytes, b_bytes;  // Phi tile / W tile storage
    float xl_limit;           // |xl| above this marks the element HOL_SEG_OUTLIER (fp32 fix-up instead of the GEMM)
};

// Plan of the dense-register dW kernel for narrow layers (dwd.cu).
struct DwdPlan {
    bool enabled;
    int smax, oc, nog;   // compiled segment bound, outputs per thread, output groups
    int R, P;            // samples per staged gy range and its shared-memory pitch (floats)
    int G, nig;          // inputs per CTA, input groups
    int nranges, nkg;    // sample ranges, range groups (= partial results)
    int ksplit;          // == nkg: partials summed by pw_dw_reduce_kernel when > 1
};

// Problem description shared by host planning code and kernels.
struct PwShape {
    int64_t B;    // rows (samples; im2col rows for conv)
    int64_t Bp;   // rows padded so that every CTA tile is full
    int I, O, n, S, disc;
    int nb;       // weights per link in the reference layout
    int stride;   // reference column step per segment: n (disc) or n-1 (cont)
    int RS;       // rows per basis group in the packed layout: S (disc) or S+1 (cont)
    int R;        // packed rows per input
    int Seff;     // number of dW buckets per input (1 for continuous n == 1)
    int nchunks;  // ceil(B / HOL_SORT_CHUNK): the bucket sort and the dW split work chunk by chunk
    int ksplit;   // dW: number of chunk groups reduced separately (deterministic two-stage sum)
    int dw_lo;    // dW: lanes per sample (power of two; 4 outputs per lane)
    int OT, OTP;  // output tile and its padded pitch (floats) in the packed layout
    int rot;      // S > 8, OT >= 32: pitch == OT and the gather kernels read the column chunks lane-rotated
    int nOT;      // number of output tiles
    int srow;     // packed floats per segment step (0 for continuous n == 1)
    int jrow[HOL_MAX_N];  // packed float offset of basis j inside an input block
    float length, half, periodicity, half_per, two_per;
    int periodic;
    BasisTab tab;
    TcPlan tc;
    DwdPlan dwd;
};

struct PwWorkspace {
    float* xl_t;      // [I][Bp]  local coordinate, transposed
    uint16_t* seg_t;  // [I][Bp]  segment index + flags, transposed
    float* wk;        // [nOT][I][R][OTP] packed weights
    uint32_t* order;  // [I][Bp]  sample ids sorted by segment (stable)
    uint32_t* offs;   // [I][nchunks][Seff+1] bucket offsets into order
    float* dw_part;   // [ksplit][O][I][nb] partial weight gradients (ksplit > 1 only)
    float* y_part;    // [fwd_ksplit][B][O] partial outputs of a forward split over K (tensor-core path)
    float* dx_rows;   // conv only: [rows][I] unfolded input gradient
    float* xl_img;    // conv only: [Bimg][C][H][W] local coordinate per pixel (before the unfold)
    uint16_t* seg_img;  // conv only: segment + flags per pixel
    void* tc_phi;     // tensor-core path, share_phi: Phi tiles of the whole batch (forward -> dW)
    void* tc_a;       // tensor-core path: Phi / Phi^T / gy tiles (fp16 hi/lo shared-memory images)
    void* tc_b;       // tensor-core path: W tiles
    void* tc_scalars; // tensor-core path: device scalars (max |xl|, max |w|, max |gy|, outlier count)
    float* tc_wmean;  // tensor-core dX: mean weight of every link [O][I] (the W^T tiles hold w - mean)
    float* gy_t;      // gather dX: upstream gradient transposed, [nOT*OT][Bp] (zero beyond O and B)
    void* prof;       // HOST pointer, normally null: hol_pw_profile_f32 collects per-GEMM events here
    size_t total;
};

static __host__ __device__ __forceinline__ int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
    return __ffma2_rn(make_float2(a, a), b, c);
}
__device__ __forceinline__ float2 ffma2v(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// phi_j(t) = C_j * prod_{m != j} (t - X_m): prefix/suffix products, no divisions.
template <int N>
__device__ __forceinline__ void lagrange_basis(float t, const BasisTab& tab, float (&phi)[N]) {
    if (N == 1) {
        phi[0] = 1.0f;
        return;
    }
    float d[N];
#pragma unroll
    for (int m = 0; m < N; ++m) d[m] = t - tab.X[m];
    float run = 1.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        phi[j] = run;
        run *= d[j];
    }
    run = 1.0f;
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        phi[j] = phi[j] * (run * tab.C[j]);
        run *= d[j];
    }
}

// d phi_j / dt = C_j * sum_{k != j} prod_{m != j,k} (t - X_m), by the recurrence
// P_l = P_{l-1} a_l, S_l = S_{l-1} a_l + P_{l-1} over the list a = {d_m : m != j}.
template <int N>
__device__ __forceinline__ void lagrange_basis_derivative(float t, const BasisTab& tab, float (&dphi)[N]) {
    if (N == 1) {
        dphi[0] = 0.0f;
        return;
    }
    float d[N];
#pragma unroll
    for (int m = 0; m < N; ++m) d[m] = t - tab.X[m];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float P = 1.0f, Ssum = 0.0f;
#pragma unroll
        for (int m = 0; m < N; ++m) {
            if (m == j) continue;
            Ssum = fmaf(Ssum, d[m], P);
            P *= d[m];
        }
        dphi[j] = Ssum * tab.C[j];
    }
}

#endif  // __CUDACC__
