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
// K = 3136).  So a TMEM accumulator only ever holds <= 48 chained instructions (bias <= 2.3e-6); the
// epilogue warps drain it into round-to-nearest fp32 register sums while the MMA warp fills the
// other of the two TMEM accumulators.
//
// Operands are stored in global memory as ready-made shared-memory images (K-major, no-swizzle
// UMMA core-matrix layout: [k/8][row][8 x fp16]), so one 1-D TMA bulk copy per operand per
// pipeline stage fills the stage -- no tensor maps.  Warp roles: warp 0 = TMA producer, warp 1 =
// MMA issuer (one elected lane), warps 2-9 = epilogue (tcgen05.ld -> fp32 register sums -> global).
#include <string.h>

#include "tc_build.cuh"

#include "tc_kernel.cuh"

// ---------------------------------------------------------------------------------------
// planning + launch
// ---------------------------------------------------------------------------------------
static int tc_bn_for(int O) { return O > 128 ? 256 : (O > 64 ? 128 : 64); }

bool tc_plan(const PwShape& s, TcPlan& t) {
    t.enabled = t.fwd = t.dx = t.dw = false;
    const char* off = getenv("HOL_DISABLE_TC");
    if (off && off[0] == '1') return false;
    t.Kin = (s.nb + 7) / 8 * 8;
    if (s.O < 32 || s.I * (int64_t)t.Kin < 256) return false;  // too narrow to be worth a tile
    // The dense form costs 3*Kin tensor MACs per n useful fp32 FMAs, so its time grows with Kin (i.e.
    // with the segment count) while the gather kernels' does not.  Cross-over points measured on
    // B200 (c5 sweep, profiles/r1_sweep_c5.md): forward/dX gather run at ~7-16 TFLOP/s useful, the
    // bucketed dW gather at ~25-40, the tensor path at ~1.3 PFLOP/s executed.
    t.fwd = t.Kin <= 40 * s.n;
    t.dx = t.fwd && t.Kin <= 128 && s.n > 1;  // the dX epilogue keeps whole inputs per thread (Kin <= 128)
    t.dw = t.Kin <= 44 + 7 * s.n;
    if (!t.fwd && !t.dw) return false;
    t.BN = tc_bn_for(s.O);
    t.nNT = (s.O + t.BN - 1) / t.BN;
    const int64_t K = (int64_t)s.I * t.Kin;
    t.nkc = (int)((K + TC_BK - 1) / TC_BK);
    // dX: N tile = whole inputs per epilogue half: a multiple of 2*Kin (and of 16), <= 256
    t.BNx = t.Kin <= 128 ? (256 / (2 * t.Kin)) * (2 * t.Kin) : 256;
    t.nNTx = (int)((K + t.BNx - 1) / t.BNx);
    t.noc = (s.O + TC_BK - 1) / TC_BK;
    // dW: M tiles over the dense columns
    t.nMTw = (int)((K + TC_BM - 1) / TC_BM);
    // In-kernel generation of the Phi / Phi^T tiles (two generator warps, <= TC_GEN_PMAX (sample, input)
    // pairs per thread and k-chunk, Kin >= 16).  Measured on B200 at C2 it is generator-bound (forward
    // GEMM 36.9 ms against 11.6 ms with materialised tiles, dW 31.7 against 11.8), so it is opt-in
    // (HOL_TC_GEN=1) until the generator keeps up with the MMA warp; the default materialises the tiles.
    const char* gen = getenv("HOL_TC_GEN");
    t.gen = gen && gen[0] == '1' && t.Kin >= 16;
    // rows per launch: everything at once when the operand is generated in the kernel, else bound the
    // Phi tiles to ~8 GiB
    int64_t rows = s.Bp;
    if (!t.gen) {
        const int64_t bytes_per_row = (int64_t)t.nkc * TC_BK * 2 * 2;
        rows = (int64_t)(8ll << 30) / bytes_per_row / 512 * 512;
        if (rows < 512) return false;
        if (rows > s.Bp) rows = s.Bp;
    }
    // Shared Phi: when forward and dW both run on the tensor cores, the forward materialises the Phi
    // tiles of the WHOLE batch once and dW reads the same tiles as its MN-major A operand -- no Phi^T
    // builder, no batch chunking.  Bounded to 40 GiB of tiles; larger layers keep the chunked scheme.
    t.share_phi = false;
    t.nkc_pad = (t.nkc + 1) & ~1;
    t.phi_bytes = 0;
    {
        const char* off2 = getenv("HOL_TC_SHARE_PHI");
        const size_t all = (size_t)((s.Bp + TC_BM - 1) / TC_BM) * t.nkc_pad * 2 * TC_A_IMG;
        if (!(off2 && off2[0] == '0') && t.fwd && t.dw && !t.gen && all <= ((size_t)40 << 30)) {
            t.share_phi = true;
            t.phi_bytes = all;
            rows = s.Bp;
        }
    }
    rows = (rows + TC_BM - 1) / TC_BM * TC_BM;
    t.rows_per_launch = rows;
    const size_t mtiles = (size_t)(rows / TC_BM);
    const size_t bchunks = (size_t)((rows + TC_BK - 1) / TC_BK);
    size_t a = (t.gen || t.share_phi) ? 0 : mtiles * t.nkc * 2 * TC_A_IMG;                    // Phi
    const size_t a_t = (t.gen || t.share_phi) ? 0 : (size_t)t.nMTw * bchunks * 2 * TC_A_IMG;  // Phi^T
    const size_t a_gy = t.dx ? mtiles * t.noc * 2 * TC_A_IMG : 0;                 // gy
    if (a_t > a) a = a_t;
    if (a_gy > a) a = a_gy;
    size_t bb = (size_t)t.nNT * t.nkc * 2 * ((size_t)t.BN * TC_BK * 2);           // W
    const size_t b_gyt = (size_t)t.nNT * bchunks * 2 * ((size_t)t.BN * TC_BK * 2);  // gy^T
    const size_t b_wt = t.dx ? (size_t)t.nNTx * t.noc * 2 * ((size_t)t.BNx * TC_BK * 2) : 0;  // W^T
    if (b_gyt > bb) bb = b_gyt;
    if (b_wt > bb) bb = b_wt;
    t.a_bytes = a > 0 ? a : 256;
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
    p.tab = s.tab;
    p.a_nkc = 0;  // set by the launchers
}

// diagnostic timing of the GEMM launches (only while hol_pw_profile_f32 runs on this thread)
struct TcProf {
    int active, count;
    cudaEvent_t ev[48][2];
    int kind[48];
};
static thread_local TcProf g_prof = {0, 0, {}, {}};

void tc_prof_begin() {
    g_prof.active = 1;
    g_prof.count = 0;
}
// ms3[k] = summed duration of the GEMM launches with epilogue k (EPI_Y, EPI_DW, EPI_DX); call after a sync
void tc_prof_end(float* ms3) {
    ms3[0] = ms3[1] = ms3[2] = 0.0f;
    for (int k = 0; k < g_prof.count; ++k) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, g_prof.ev[k][0], g_prof.ev[k][1]) == cudaSuccess) ms3[g_prof.kind[k]] += ms;
        cudaEventDestroy(g_prof.ev[k][0]);
        cudaEventDestroy(g_prof.ev[k][1]);
    }
    g_prof.active = 0;
    g_prof.count = 0;
}

static int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

template <int EPI, int N, int AMODE>
static int launch_gemm(TcGemmParams p, int64_t mtiles, int ntiles_n, cudaStream_t st) {
    constexpr bool GEN = AMODE == A_GEN;
    if (p.a_nkc == 0) p.a_nkc = p.nkc;
    const int stage = 2 * TC_A_IMG + 2 * p.BN * TC_BK * 2;
    int nst = (HOL_SMEM_BUDGET - 1024) / stage;
    if (nst > 6) nst = 6;
    if (nst < 2) return HOL_EUNSUPPORTED;
    const size_t smem = 1024 + (size_t)nst * stage;
    p.nst = nst;
    // k-chunks accumulated in TMEM between drains: 4 chunks = 48 chained MMA instructions.  The tensor
    // core accumulates with round-toward-zero (bias ~ chain length); measured against the fp64 oracle
    // with same-sign operands (the worst case, tools/check_flush_error.py): y / dX / dW relative error
    // 1.4e-6 / 1.7e-6 / 5e-7 at 2 chunks, 1.3e-6 / 2.3e-6 / 9e-7 at 4, while draining half as often
    // takes the dX GEMM of C2 from 15.7 to 13.6 ms.
    p.flush_chunks = 4;
    if (const char* fc = getenv("HOL_TC_FLUSH")) {  // experiment knob: k-chunks per accumulator drain
        const int v = atoi(fc);
        if (v >= 1 && v <= 8) p.flush_chunks = v;
    }
    p.ntile_n = ntiles_n;
    p.ntile_m = mtiles;
    p.ntiles = mtiles * ntiles_n;
    p.group_m = ntiles_n >= 12 ? 12 : (148 + ntiles_n - 1) / ntiles_n;  // ~square block of co-running tiles
    if (p.group_m > mtiles) p.group_m = (int)mtiles;
    if (p.group_m < 1) p.group_m = 1;
    auto kern = pw_tc_gemm_kernel<EPI, N, AMODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t sms = sm_count();
    const unsigned grid = (unsigned)(p.ntiles < sms ? p.ntiles : sms);  // persistent: one CTA per SM
    const bool prof = g_prof.active && g_prof.count < 48;
    if (prof) {
        cudaEventCreate(&g_prof.ev[g_prof.count][0]);
        cudaEventCreate(&g_prof.ev[g_prof.count][1]);
        g_prof.kind[g_prof.count] = EPI;
        cudaEventRecord(g_prof.ev[g_prof.count][0], st);
    }
    kern<<<grid, GEN ? 384 : 320, smem, st>>>(p);
    if (prof) cudaEventRecord(g_prof.ev[g_prof.count++][1], st);
    return (int)cudaGetLastError();
}

// kernels of ours per call (bench.py's gpu_launches): kind 0 forward, 1 dX, 2 dW (prep not included)
int tc_launch_count(const PwShape& s, const TcPlan& t, int kind) {
    const int nl = (int)((s.Bp + t.rows_per_launch - 1) / t.rows_per_launch);
    const int ex = t.gen ? 0 : 1;          // Phi / Phi^T expansion kernel (absent when generated in the GEMM)
    if (kind == 0) return 2 + 1 + nl * (1 + ex);  // absmax x2, pack_w, ([expand,] gemm) per batch chunk
    if (kind == 1) return 2 + 2 + nl * 2;         // absmax x2, link mean, pack_wt, (pack_gy, gemm)
    if (t.share_phi) return 2 + 2;                // absmax x2, pack_gyt, gemm (the Phi tiles come from the forward)
    return 2 + nl * (2 + ex);                     // absmax x2, ([expand_t,] pack_gyt, gemm)
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
        // shared Phi: one launch builds the tiles of the whole batch (k-chunks padded to even with zero tiles)
        const int ekc = t.share_phi ? t.nkc_pad : t.nkc;
        unsigned char* phi = (unsigned char*)(t.share_phi ? ws.tc_phi : ws.tc_a);
        dim3 egrid((unsigned)(rows / TC_BM), (unsigned)ekc);
        rc = HOL_OK;
        if (!t.gen) {
#define EXPAND(NN) (pw_tc_expand_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, phi, sc, s.Bp, row0, s.I, t.Kin, ekc, s.stride, s.tab), (int)cudaGetLastError())
            DISPATCH_N(s.n, EXPAND)
#undef EXPAND
        }
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
        p.a_nkc = ekc;
        p.a_tiles = phi;
        rc = t.gen ? launch_gemm<EPI_Y, 0, A_GEN>(p, rows / TC_BM, t.nNT, st)
                   : launch_gemm<EPI_Y, 0, A_TILES>(p, rows / TC_BM, t.nNT, st);
        if (rc != HOL_OK) return rc;
    }
    return HOL_OK;
}

int launch_backward_dw_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* gy, float* dw,
                          bool phi_valid, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = tc_absmax(ws.xl_t, (int64_t)s.I * s.Bp, &sc->max_xl, st);
    if (rc != HOL_OK) return rc;
    rc = tc_absmax(gy, s.B * (int64_t)s.O, &sc->max_gy, st);
    if (rc != HOL_OK) return rc;
    if (t.share_phi) {
        const int64_t mtiles = (s.Bp + TC_BM - 1) / TC_BM;
        if (!phi_valid) {  // backward without the forward's workspace: build the Phi tiles first
            dim3 egrid((unsigned)mtiles, (unsigned)t.nkc_pad);
#define EXPAND(NN) (pw_tc_expand_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, (unsigned char*)ws.tc_phi, sc, s.Bp, (int64_t)0, s.I, t.Kin, t.nkc_pad, s.stride, s.tab), (int)cudaGetLastError())
            DISPATCH_N(s.n, EXPAND)
#undef EXPAND
            if (rc != HOL_OK) return rc;
        }
        const int nbc = (int)((s.Bp + TC_BK - 1) / TC_BK);
        dim3 ggrid((unsigned)t.nNT, (unsigned)nbc);
        pw_tc_pack_gyt_kernel<<<ggrid, 256, 0, st>>>(gy, (unsigned char*)ws.tc_b, sc, s.B, (int64_t)0, s.O, nbc, t.BN);
        rc = (int)cudaGetLastError();
        if (rc != HOL_OK) return rc;
        TcGemmParams p;
        fill_common(p, s, ws, t);
        p.a_tiles = (const unsigned char*)ws.tc_phi;
        p.a_nkc = t.nkc_pad;
        p.out = dw;
        p.row0 = 0;
        p.rows_valid = 0;
        p.BN = t.BN;
        p.n_valid = s.O;
        p.ld = 0;
        p.nkc = nbc;  // 64-sample stages: stage sc = samples 64*(sc & 1).. of batch tile sc >> 1
        return launch_gemm<EPI_DW, 0, A_PHI_MN>(p, t.nkc_pad / 2, t.nNT, st);
    }
    for (int64_t row0 = 0; row0 < s.Bp; row0 += t.rows_per_launch) {
        int64_t rows = s.Bp - row0 < t.rows_per_launch ? s.Bp - row0 : t.rows_per_launch;
        const int nbc = (int)((rows + TC_BK - 1) / TC_BK);
        dim3 egrid((unsigned)t.nMTw, (unsigned)nbc);
        rc = HOL_OK;
        if (!t.gen) {
#define EXPANDT(NN) (pw_tc_expand_t_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, (unsigned char*)ws.tc_a, sc, s.Bp, row0, s.I, t.Kin, nbc, s.stride, s.tab), (int)cudaGetLastError())
            DISPATCH_N(s.n, EXPANDT)
#undef EXPANDT
        }
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
        rc = t.gen ? launch_gemm<EPI_DW, 0, A_GEN>(p, t.nMTw, t.nNT, st) : launch_gemm<EPI_DW, 0, A_TILES>(p, t.nMTw, t.nNT, st);
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
    const int64_t links = (int64_t)s.O * s.I;
    pw_tc_link_mean_kernel<<<(unsigned)((links + 7) / 8), 256, 0, st>>>(w, ws.tc_wmean, links, s.nb);
    rc = (int)cudaGetLastError();
    if (rc != HOL_OK) return rc;
    dim3 pgrid((unsigned)t.nNTx, (unsigned)t.noc);
    pw_tc_pack_wt_kernel<<<pgrid, 256, 0, st>>>(w, ws.tc_wmean, (unsigned char*)ws.tc_b, sc, s.I, s.O, s.nb, t.Kin, t.noc,
                                                t.BNx);
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
#define GEMM_DX(NN) launch_gemm<EPI_DX, NN, A_TILES>(p, rows / TC_BM, t.nNTx, st)
        DISPATCH_N(s.n, GEMM_DX)
#undef GEMM_DX
        if (rc != HOL_OK) return rc;
    }
    return HOL_OK;
}
