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
// pipeline stage fills the stage -- no tensor maps.  Warp roles: warp 0 = producer, warp 1 =
// MMA issuer (one elected lane), warps 2-9 = epilogue (tcgen05.ld -> fp32 register sums -> global).
//
// Elements whose |xl| is beyond what the fp16 tiles can carry are left out of the tiles and added
// in fp32 by the fix-up kernels (tc_fixup.cuh), so samples stay independent of one another.
#include <string.h>

#include "tc_build.cuh"

#include "tc_fixup.cuh"
#include "tc_kernel.cuh"

// ---------------------------------------------------------------------------------------
// planning + launch
// ---------------------------------------------------------------------------------------
static int sm_count() {
    static std::atomic<int> cached{0};
    int n = cached.load(std::memory_order_relaxed);
    if (n > 0) return n;
    int dev = 0;
    n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cached.store(n, std::memory_order_relaxed);
    return n;
}

// Output tile width.  A 128 x BN x 16 MMA takes ~BN/2 cycles down to BN = 128; at BN = 64 the shared-
// memory read of the A tile binds (~48 instead of 32 cycles) and every Phi tile is fetched by twice as
// many CTAs (measured at the 784 -> 128 layer of config 3: the forward with two 64-wide n-tiles ran at
// the HBM rate of reading Phi twice, 128 us, against ~80 us of MMA time for one 128-wide tile on 64
// SMs).  The forward has no split over K, so the tile is only narrowed when the batch gives fewer
// than ~60 tiles (40 % of the SMs).
static int tc_bn_for(int O, int64_t mtiles, bool may_shrink) {
    int bn = O > 128 ? 256 : (O > 64 ? 128 : 64);
    if (!may_shrink) return bn;
    while (bn > 64 && mtiles * ((O + bn - 1) / bn) < 60 && (O + bn / 2 - 1) / (bn / 2) > (O + bn - 1) / bn) bn >>= 1;
    return bn;
}

// Split of the reduction dimension over several work items per tile: the persistent grid runs
// `slots` work items at a time (pairs of tiles when there are at least two m-tiles), so the time is
// ~ceil(items / slots) waves of 1/ks of the k-chunks each; a small charge per split stands for the
// partial results' reduction pass.  E.g. 98 dW tiles (49 pairs on 74 slots): ks = 3 -> 2 waves of 1/3.
static int pick_ksplit(int64_t mtiles, int64_t ntiles, int64_t kchunks, int min_chunks, int max_ks) {
    const int64_t units = (mtiles >= 2 ? (mtiles + 1) / 2 : mtiles) * ntiles;
    const int64_t slots = mtiles >= 2 ? sm_count() / 2 : sm_count();
    if (units >= 2 * slots) return 1;  // several waves anyway: the tail is not worth partial buffers of the full result
    int best = 1;
    double best_cost = 1e30;
    for (int ks = 1; ks <= max_ks; ++ks) {
        if (ks > 1 && kchunks / ks < min_chunks) break;
        const double cost = (double)((units * ks + slots - 1) / slots) / ks + 0.02 * (ks - 1);
        if (cost < best_cost - 1e-9) best_cost = cost, best = ks;
    }
    return best;
}

bool tc_plan(const PwShape& s, TcPlan& t) {
    memset(&t, 0, sizeof(t));
    const uint32_t ovr = hol_plan_overrides();
    if (ovr & HOL_OVR_DISABLE_TC) return false;
    // Dense columns per input.  Phi, the forward's W tiles and dW pack the inputs back to back (k = i*nb + v):
    // padding every input to a multiple of 8 columns would add MMA work and Phi bytes for nothing (nb = 13 at
    // config 3: 16 -> 13 is 19 % of both).  Only dX, whose epilogue wants whole inputs per thread, pads.
    t.Kin = s.nb;
    t.Kinx = (s.nb + 7) / 8 * 8;
    if (s.O < 32 || s.I * (int64_t)t.Kinx < 256) return false;  // too narrow to be worth a tile
    // The dense form costs 3*Kin tensor MACs per n useful fp32 FMAs, so its time grows with Kin (i.e.
    // with the segment count) while the gather kernels' does not.  Cross-over points measured on
    // B200 (c5 sweep, profiles/r1c_sweep_c5.jsonl): forward/dX gather run at ~7-16 TFLOP/s useful, the
    // bucketed dW gather at ~25-40, the tensor path at ~1.3 PFLOP/s executed.
    t.fwd = t.Kinx <= 40 * s.n;
    t.dx = t.fwd && t.Kinx <= 128 && s.n > 1;  // the dX epilogue keeps whole inputs per thread (Kinx <= 128)
    t.dw = t.Kinx <= 44 + 7 * s.n;
    if (!t.fwd && !t.dw) return false;
    const int64_t K = (int64_t)s.I * t.Kin;
    const int64_t mtiles_all = (s.Bp + TC_BM - 1) / TC_BM;
    t.nkc = (int)((K + TC_BK - 1) / TC_BK);
    t.BN = tc_bn_for(s.O, mtiles_all, true);
    t.nNT = (s.O + t.BN - 1) / t.BN;
    // dX: N tile = whole inputs per epilogue half: a multiple of 2*Kinx (and of 16), <= 256
    t.BNx = t.Kinx <= 128 ? (256 / (2 * t.Kinx)) * (2 * t.Kinx) : 256;
    t.nNTx = (int)(((int64_t)s.I * t.Kinx + t.BNx - 1) / t.BNx);
    t.noc = (s.O + TC_BK - 1) / TC_BK;
    // dW: M tiles over the dense columns; when they are too few to fill the GPU the batch (K) is split
    t.nMTw = (int)((K + TC_BM - 1) / TC_BM);
    t.BNw = tc_bn_for(s.O, t.nMTw, false);
    t.nNTw = (s.O + t.BNw - 1) / t.BNw;
    t.flush_chunks = 4;
    if (const int fc = (int)((ovr >> HOL_OVR_FLUSH_SHIFT) & 15u)) t.flush_chunks = fc > 8 ? 8 : fc;
    // rows per launch: bound the Phi tiles to ~8 GiB
    int64_t rows = s.Bp;
    {
        const int64_t bytes_per_row = (int64_t)t.nkc * TC_BK * 2 * 2;
        rows = (int64_t)(8ll << 30) / bytes_per_row / 512 * 512;
        if (rows < 512) return false;
        if (rows > s.Bp) rows = s.Bp;
    }
    // Shared Phi: when forward and dW both run on the tensor cores, the forward materialises the Phi
    // tiles of the WHOLE batch once and dW reads the same tiles as its MN-major A operand -- no Phi^T
    // builder, no batch chunking.  Bounded to 40 GiB of tiles; larger layers keep the chunked scheme.
    t.nkc_pad = (t.nkc + 1) & ~1;
    {
        const size_t all = (size_t)mtiles_all * t.nkc_pad * 2 * TC_A_IMG;
        if (!(ovr & HOL_OVR_NO_SHARE_PHI) && t.fwd && t.dw && all <= ((size_t)40 << 30)) {
            t.share_phi = true;
            t.phi_bytes = all;
            rows = s.Bp;
        }
    }
    rows = (rows + TC_BM - 1) / TC_BM * TC_BM;
    t.rows_per_launch = rows;
    t.dw_ksplit = 1;
    if (t.share_phi && t.dw)  // 64-sample stages, at least 8 per split
        t.dw_ksplit = pick_ksplit(t.nkc_pad / 2, t.nNTw, (s.Bp + TC_BK - 1) / TC_BK, 8, 16);
    // forward: no reduction dimension to spare SMs with either -- when the batch gives fewer tiles than ~half
    // the SMs (config 3: 64 tiles of 196 k-chunks) the k-chunks are split over 2-4 work items per tile; every
    // CTA is bound by its own shared-memory / L2 port, so the extra CTAs add bandwidth, not just issue slots
    t.fwd_ksplit = 1;
    if (t.fwd && rows >= s.Bp) t.fwd_ksplit = pick_ksplit(mtiles_all, t.nNT, t.nkc, 12, 4);
    const size_t mtiles = (size_t)(rows / TC_BM);
    const size_t bchunks = (size_t)((rows + TC_BK - 1) / TC_BK);
    size_t a = t.share_phi ? 0 : mtiles * t.nkc * 2 * TC_A_IMG;                    // Phi
    const size_t a_t = t.share_phi ? 0 : (size_t)t.nMTw * bchunks * 2 * TC_A_IMG;  // Phi^T
    const size_t a_gy = t.dx ? mtiles * t.noc * 2 * TC_A_IMG : 0;                  // gy
    if (a_t > a) a = a_t;
    if (a_gy > a) a = a_gy;
    size_t bb = (size_t)t.nNT * t.nkc * 2 * ((size_t)t.BN * TC_BK * 2);             // W
    const size_t b_gyt = (size_t)t.nNTw * bchunks * 2 * ((size_t)t.BNw * TC_BK * 2);  // gy^T
    const size_t b_wt = t.dx ? (size_t)t.nNTx * t.noc * 2 * ((size_t)t.BNx * TC_BK * 2) : 0;  // W^T
    if (b_gyt > bb) bb = b_gyt;
    if (b_wt > bb) bb = b_wt;
    t.a_bytes = a > 0 ? a : 256;
    t.b_bytes = bb;
    t.xl_limit = tc_xl_limit(s.tab, s.n);
    t.enabled = true;
    return true;
}

static void fill_common(TcGemmParams& p, const PwShape& s, const PwWorkspace& ws, const TcPlan& t) {
    memset(&p, 0, sizeof(p));
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
    p.nseg = s.Seff;
    p.transposed = 0;
    p.length = s.length;
    p.Sf = (float)s.S;
    p.tab = s.tab;
    p.flush_chunks = t.flush_chunks;
    p.ksplit = 1;
    p.a_nkc = 0;  // set by the launchers
}

// diagnostic timing of the GEMM launches: hol_pw_profile_f32 hands a TcProf through PwWorkspace::prof
void tc_prof_end(TcProf* prof, float* ms3) {
    ms3[0] = ms3[1] = ms3[2] = 0.0f;
    for (int k = 0; k < prof->count; ++k) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, prof->ev[k][0], prof->ev[k][1]) == cudaSuccess) ms3[prof->kind[k]] += ms;
        cudaEventDestroy(prof->ev[k][0]);
        cudaEventDestroy(prof->ev[k][1]);
    }
    prof->count = 0;
}

template <int EPI, int N, int AMODE, int CL>
static int launch_gemm_cl(TcGemmParams p, int64_t mtiles, int ntiles_n, TcProf* prof, size_t smem, cudaStream_t st) {
    static std::atomic<uint64_t> attr_done{0};
    p.ntile_n = ntiles_n;
    p.ntile_m = (mtiles + CL - 1) / CL;  // the rasterisation walks m-tiles or pairs of m-tiles
    p.mt_last = mtiles - 1;
    p.ntiles = p.ntile_m * ntiles_n;
    p.nitems = p.ntiles * p.ksplit;
    p.group_m = ntiles_n >= 12 ? 12 : (148 / CL + ntiles_n - 1) / ntiles_n;  // ~square block of co-running tiles
    if (p.group_m > p.ntile_m) p.group_m = (int)p.ntile_m;
    if (p.group_m < 1) p.group_m = 1;
    auto kern = pw_tc_gemm_kernel<EPI, N, AMODE, CL>;
    cudaError_t e = hol_ensure_max_smem(kern, attr_done);
    if (e != cudaSuccess) return (int)e;
    const int64_t slots = sm_count() / CL;  // persistent: one CTA per SM, CL CTAs per work item
    const unsigned grid = (unsigned)((p.nitems < slots ? p.nitems : slots) * CL);
    const bool rec = prof && prof->count < 48;
    if (rec) {
        cudaEventCreate(&prof->ev[prof->count][0]);
        cudaEventCreate(&prof->ev[prof->count][1]);
        prof->kind[prof->count] = EPI;
        cudaEventRecord(prof->ev[prof->count][0], st);
    }
    if (CL == 1) {
        kern<<<grid, TC_THREADS, smem, st>>>(p);
    } else {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = CL;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, p);
        if (e != cudaSuccess) return (int)e;
    }
    if (rec) cudaEventRecord(prof->ev[prof->count++][1], st);
    return (int)cudaGetLastError();
}

template <int EPI, int N, int AMODE>
static int launch_gemm(TcGemmParams p, int64_t mtiles, int ntiles_n, int ksplit, TcProf* prof, cudaStream_t st) {
    if (p.a_nkc == 0) p.a_nkc = p.nkc;
    const int stage = 2 * TC_A_IMG + 2 * p.BN * TC_BK * 2;
    int nst = (HOL_SMEM_BUDGET - 1024) / stage;
    if (nst > 6) nst = 6;
    if (nst < 2) return HOL_EUNSUPPORTED;
    const size_t smem = 1024 + (size_t)nst * stage;
    p.nst = nst;
    // flush_chunks = k-chunks accumulated in TMEM between drains: 4 chunks = 48 chained MMA instructions.
    // The tensor core accumulates with round-toward-zero (bias ~ chain length); measured against the fp64
    // oracle with same-sign operands (the worst case, tools/check_flush_error.py): y / dX / dW relative
    // error 1.4e-6 / 1.7e-6 / 5e-7 at 2 chunks, 1.3e-6 / 2.3e-6 / 9e-7 at 4, while draining half as often
    // takes the dX GEMM of C2 from 15.7 to 13.6 ms.
    p.ksplit = ksplit < 1 ? 1 : ksplit;
    p.kc_per_split = (p.nkc + p.ksplit - 1) / p.ksplit;
    p.ksplit = (p.nkc + p.kc_per_split - 1) / p.kc_per_split;  // no empty work items
    // Pairs of m-tiles as two-CTA clusters that multicast the B tile to each other (tc_kernel.cuh) whenever
    // there are enough pairs to fill the machine; small problems keep single CTAs.
    const int64_t pair_items = ((mtiles + 1) / 2) * ntiles_n * p.ksplit;
    const bool pairs = mtiles >= 2 && pair_items >= 32 && !(hol_plan_overrides() & HOL_OVR_NO_CLUSTER);
    return pairs ? launch_gemm_cl<EPI, N, AMODE, 2>(p, mtiles, ntiles_n, prof, smem, st)
                 : launch_gemm_cl<EPI, N, AMODE, 1>(p, mtiles, ntiles_n, prof, smem, st);
}

// kernels of ours per call (bench.py's gpu_launches): kind 0 forward, 1 dX, 2 dW, 3 the shared
// max|gy| pass of a backward with any tensor-core part (prep not included)
int tc_launch_count(const PwShape& s, const TcPlan& t, int kind, bool prepared) {
    const int nl = (int)((s.Bp + t.rows_per_launch - 1) / t.rows_per_launch);
    if (kind == 0) return 2 + nl * 2 + 1;                          // wstats, pack_w, (expand, gemm) per batch chunk, finalize
    if (kind == 1) return (prepared ? 0 : 1) + 1 + nl * 2 + 1;     // [wstats], pack_wt, (pack_gy, gemm), fix_dx
    if (kind == 3) return 1;
    if (t.share_phi) return (prepared && t.fwd ? 0 : 1) + 2 + (t.dw_ksplit > 1 ? 1 : 0) + 1;  // [expand], pack_gyt, gemm, [reduce], fix_dw
    return nl * 3 + 1;                                             // (expand_t, pack_gyt, gemm), fix_dw
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

static void fill_fix(FixParams& f, const PwShape& s, const PwWorkspace& ws, const float* w, const float* gy, float* out,
                     int transposed) {
    f.xl_t = ws.xl_t;
    f.seg_t = ws.seg_t;
    f.w = w;
    f.gy = gy;
    f.out = out;
    f.sc = (const TcScalars*)ws.tc_scalars;
    f.B = s.B;
    f.Bp = s.Bp;
    f.I = s.I;
    f.O = s.O;
    f.nb = s.nb;
    f.stride = s.stride;
    f.n = s.n;
    f.transposed = transposed;
    f.length = s.length;
    f.Sf = (float)s.S;
    f.tab = s.tab;
}

// max |w| and the link means (one pass over the weights; the forward leaves both in the workspace)
int launch_tc_wstats(const PwShape& s, const PwWorkspace& ws, const float* w, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    const int64_t links = (int64_t)s.O * s.I;
    int64_t blocks = s.nb <= 32 ? (links + 255) / 256 : (links + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pw_tc_wstats_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, ws.tc_wmean, &sc->max_w, links, s.nb);
    return (int)cudaGetLastError();
}

int launch_tc_gy_absmax(const PwShape& s, const PwWorkspace& ws, const float* gy, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    const int64_t count = s.B * (int64_t)s.O;
    int64_t blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    pw_absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(gy, count, &sc->max_gy);
    return (int)cudaGetLastError();
}

static int launch_expand(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, unsigned char* phi, int64_t row0,
                         int64_t rows, int ekc, cudaStream_t st) {
    const TcScalars* sc = (const TcScalars*)ws.tc_scalars;
    dim3 egrid((unsigned)(rows / TC_BM), (unsigned)ekc);
    int rc = HOL_OK;
#define EXPAND(NN) (pw_tc_expand_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, phi, sc, s.Bp, row0, s.I, t.Kin, ekc, s.stride, s.tab), (int)cudaGetLastError())
    DISPATCH_N(s.n, EXPAND)
#undef EXPAND
    return rc;
}

// norm.kind >= 0: the per-sample normalisation that follows the layer is applied by the forward's last
// kernel (y = raw output, norm.y_norm / norm.stats = normalised output and its statistics)
int launch_forward_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, float* y,
                      const PwNormArgs& norm, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = launch_tc_wstats(s, ws, w, st);
    if (rc != HOL_OK) return rc;
    dim3 pgrid((unsigned)t.nNT, (unsigned)t.nkc, (unsigned)(t.BN * 8 / 256));
    pw_tc_pack_w_kernel<<<pgrid, 256, 0, st>>>(w, (unsigned char*)ws.tc_b, sc, s.I, s.O, s.nb, t.Kin, t.nkc, t.BN);
    rc = (int)cudaGetLastError();
    if (rc != HOL_OK) return rc;
    for (int64_t row0 = 0; row0 < s.Bp; row0 += t.rows_per_launch) {
        int64_t rows = s.Bp - row0 < t.rows_per_launch ? s.Bp - row0 : t.rows_per_launch;
        rows = (rows + TC_BM - 1) / TC_BM * TC_BM;
        // shared Phi: one launch builds the tiles of the whole batch (k-chunks padded to even with zero tiles)
        const int ekc = t.share_phi ? t.nkc_pad : t.nkc;
        unsigned char* phi = (unsigned char*)(t.share_phi ? ws.tc_phi : ws.tc_a);
        rc = launch_expand(s, ws, t, phi, row0, rows, ekc, st);
        if (rc != HOL_OK) return rc;
        TcGemmParams p;
        fill_common(p, s, ws, t);
        const bool split = t.fwd_ksplit > 1;  // only planned when one launch covers the batch
        p.out = split ? ws.y_part : y;
        p.part_stride = s.B * (int64_t)s.O;
        p.row0 = row0;
        p.rows_valid = s.B - row0 < rows ? s.B - row0 : rows;
        p.BN = t.BN;
        p.n_valid = s.O;
        p.ld = s.O;
        p.nkc = t.nkc;
        p.a_nkc = ekc;
        p.a_tiles = phi;
        rc = launch_gemm<EPI_Y, 0, A_TILES>(p, rows / TC_BM, t.nNT, t.fwd_ksplit, (TcProf*)ws.prof, st);
        if (rc != HOL_OK) return rc;
    }
    // one pass over the rows: sum of the k-split partials, outlier contributions, optional normalisation
    FinalizeParams q;
    fill_fix(q.fix, s, ws, w, nullptr, y, 0);
    q.parts = t.fwd_ksplit > 1 ? ws.y_part : nullptr;
    {
        const int kcs = (t.nkc + t.fwd_ksplit - 1) / t.fwd_ksplit;
        q.ks = (t.nkc + kcs - 1) / kcs;
    }
    q.y_norm = norm.kind >= 0 ? norm.y_norm : nullptr;
    q.stats = reinterpret_cast<float4*>(norm.stats);
    q.norm_kind = norm.kind;
    q.eps = norm.eps;
    pw_tc_finalize_kernel<<<(unsigned)((s.B + 7) / 8), 256, 0, st>>>(q);
    return (int)cudaGetLastError();
}

// `gy` statistics (TcScalars::max_gy) must have been taken by launch_tc_gy_absmax before this call
int launch_backward_dw_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* gy, float* dw,
                          bool phi_valid, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = HOL_OK;
    FixParams f;
    fill_fix(f, s, ws, nullptr, gy, dw, 0);
    if (t.share_phi) {
        const int64_t mtiles = (s.Bp + TC_BM - 1) / TC_BM;
        if (!phi_valid) {  // backward without the forward's workspace: build the Phi tiles first
            rc = launch_expand(s, ws, t, (unsigned char*)ws.tc_phi, 0, mtiles * TC_BM, t.nkc_pad, st);
            if (rc != HOL_OK) return rc;
        }
        const int nbc = (int)((s.Bp + TC_BK - 1) / TC_BK);
        dim3 ggrid((unsigned)t.nNTw, (unsigned)nbc, (unsigned)(t.BNw * 8 / 256));
        pw_tc_pack_gyt_kernel<<<ggrid, 256, 0, st>>>(gy, (unsigned char*)ws.tc_b, sc, s.B, (int64_t)0, s.O, nbc, t.BNw);
        rc = (int)cudaGetLastError();
        if (rc != HOL_OK) return rc;
        TcGemmParams p;
        fill_common(p, s, ws, t);
        p.a_tiles = (const unsigned char*)ws.tc_phi;
        p.a_nkc = t.nkc_pad;
        const bool split = t.dw_ksplit > 1;
        p.out = split ? ws.dw_part : dw;
        p.part_stride = (int64_t)s.O * s.I * s.nb;
        p.row0 = 0;
        p.rows_valid = 0;
        p.BN = t.BNw;
        p.n_valid = s.O;
        p.ld = 0;
        p.nkc = nbc;  // 64-sample stages: stage sc = samples 64*(sc & 1).. of batch tile sc >> 1
        rc = launch_gemm<EPI_DW, 0, A_PHI_MN>(p, t.nkc_pad / 2, t.nNTw, t.dw_ksplit, (TcProf*)ws.prof, st);
        if (rc != HOL_OK) return rc;
        if (split) {
            const int ks = (nbc + ((nbc + t.dw_ksplit - 1) / t.dw_ksplit) - 1) / ((nbc + t.dw_ksplit - 1) / t.dw_ksplit);
            rc = launch_dw_reduce(ws.dw_part, dw, (int64_t)s.O * s.I * s.nb, ks, st);
            if (rc != HOL_OK) return rc;
        }
    } else {
        for (int64_t row0 = 0; row0 < s.Bp; row0 += t.rows_per_launch) {
            int64_t rows = s.Bp - row0 < t.rows_per_launch ? s.Bp - row0 : t.rows_per_launch;
            const int nbc = (int)((rows + TC_BK - 1) / TC_BK);
            dim3 egrid((unsigned)t.nMTw, (unsigned)nbc);
            rc = HOL_OK;
#define EXPANDT(NN) (pw_tc_expand_t_kernel<NN><<<egrid, 256, 0, st>>>(ws.xl_t, ws.seg_t, (unsigned char*)ws.tc_a, sc, s.Bp, row0, s.I, t.Kin, nbc, s.stride, s.tab), (int)cudaGetLastError())
            DISPATCH_N(s.n, EXPANDT)
#undef EXPANDT
            if (rc != HOL_OK) return rc;
            dim3 ggrid((unsigned)t.nNTw, (unsigned)nbc, (unsigned)(t.BNw * 8 / 256));
            pw_tc_pack_gyt_kernel<<<ggrid, 256, 0, st>>>(gy, (unsigned char*)ws.tc_b, sc, s.B, row0, s.O, nbc, t.BNw);
            rc = (int)cudaGetLastError();
            if (rc != HOL_OK) return rc;
            TcGemmParams p;
            fill_common(p, s, ws, t);
            p.out = dw;
            p.row0 = row0;  // > 0: accumulate onto the previous batch chunk's result
            p.rows_valid = 0;
            p.BN = t.BNw;
            p.n_valid = s.O;
            p.ld = 0;
            p.nkc = nbc;
            rc = launch_gemm<EPI_DW, 0, A_TILES>(p, t.nMTw, t.nNTw, 1, (TcProf*)ws.prof, st);
            if (rc != HOL_OK) return rc;
        }
    }
    pw_tc_fix_dw_kernel<<<(unsigned)((s.I + 7) / 8), 256, 0, st>>>(f);
    return (int)cudaGetLastError();
}

// `wstats_valid`: TcScalars::max_w and the link means in ws.tc_wmean are those of `w` (left by the forward)
int launch_backward_dx_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, const float* gy,
                          float* dx, int transposed, bool wstats_valid, cudaStream_t st) {
    TcScalars* sc = (TcScalars*)ws.tc_scalars;
    int rc = HOL_OK;
    if (!wstats_valid) {
        rc = launch_tc_wstats(s, ws, w, st);
        if (rc != HOL_OK) return rc;
    }
    dim3 pgrid((unsigned)t.nNTx, (unsigned)t.noc, (unsigned)((t.BNx * 8 + 255) / 256));
    pw_tc_pack_wt_kernel<<<pgrid, 256, 0, st>>>(w, ws.tc_wmean, (unsigned char*)ws.tc_b, sc, s.I, s.O, s.nb, t.Kinx, t.noc,
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
        p.Kin = t.Kinx;
        p.transposed = transposed;
#define GEMM_DX(NN) launch_gemm<EPI_DX, NN, A_TILES>(p, rows / TC_BM, t.nNTx, 1, (TcProf*)ws.prof, st)
        DISPATCH_N(s.n, GEMM_DX)
#undef GEMM_DX
        if (rc != HOL_OK) return rc;
    }
    FixParams f;
    fill_fix(f, s, ws, w, gy, dx, transposed);
    pw_tc_fix_rows_kernel<true><<<(unsigned)((s.B + 255) / 256), 256, 0, st>>>(f);
    return (int)cudaGetLastError();
}
