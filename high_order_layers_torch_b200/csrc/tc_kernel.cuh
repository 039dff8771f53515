#pragma once
// Persistent split-fp16 tcgen05 GEMM with three epilogues (design notes: tc_gemm.cu).
#include "tc_build.cuh"

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
// One lane of a converged warp (cute::elect_one_sync).  ptxas knows that a branch on elect.sync has
// exactly one active thread and issues the warp-uniform instructions inside it (UTCHMMA, UBLKCP,
// UTCBAR) directly; behind a plain `lane == 0` test it wraps EVERY one of them in an
// ELECT / BRA.U.ANY loop, which made the MMA-issuing thread the bottleneck of all three GEMMs
// (ncu: 81 % of its samples in those wrappers, ~2100 cycles of issue per 1440-cycle k-chunk).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 1-D bulk copy global -> the same shared-memory offset of every CTA in `mask` of the cluster; each
// destination CTA's mbarrier (same offset) receives complete_tx for the bytes that landed in it
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct TcGemmParams {
    const unsigned char* a_tiles;  // [m-tile][k-chunk][hi|lo] images of 128 rows
    const unsigned char* b_tiles;  // [n-tile][k-chunk][hi|lo] images of BN rows
    float* out;                    // EPI_Y: y[rows][ld]; EPI_DW: dw[O][I][nb] (or its partials); EPI_DX: dx[B][I] or dx_t[I][Bp]
    const TcScalars* sc;
    const float* xl_t;             // EPI_DX epilogue
    const uint16_t* seg_t;
    int64_t row0, rows_valid, Bp;  // EPI_Y/DX: batch rows of this launch; EPI_DW: row0 > 0 => accumulate into out
    int64_t ntiles;                // m-tiles * n-tiles
    int64_t nitems;                // ntiles * ksplit
    int64_t ntile_m;               // m-tiles (CL = 1) or pairs of m-tiles (CL = 2) the rasterisation walks
    int64_t mt_last;               // last valid m-tile (a pair may end on a tile that does not exist: loads are
                                   // clamped to this tile and its epilogue stores nothing)
    int64_t part_stride;           // ksplit > 1 (EPI_DW, EPI_Y): floats between the partial results of two k-splits
    int ntile_n, group_m;          // rasterisation: groups of group_m m-tiles, m fastest inside a group
    int ksplit, kc_per_split;      // the k-chunks of a tile are split over ksplit work items
    int BN, n_valid, ld, nkc, nst, n_basis, flush_chunks;
    int a_nkc;                     // k-chunks per m-tile in the A tile storage (>= nkc: shared Phi pads to even)
    int I, nb, Kin, stride, nseg, transposed;
    float length, Sf;
    BasisTab tab;
};

// tile index -> (m-tile, n-tile): m-tiles are taken in groups of group_m, inside a group the m-tile
// varies fastest, so the ~148 CTAs running together cover a group_m x (148/group_m) block of tiles
// and share both their A tiles and their B tiles in L2.
__device__ __forceinline__ void tile_coords(const TcGemmParams& p, int64_t tile, int64_t& mt, int& nt) {
    const int64_t per_group = (int64_t)p.group_m * p.ntile_n;
    const int64_t group = tile / per_group;
    const int64_t first_m = group * p.group_m;
    const int64_t gsize = p.ntile_m - first_m < p.group_m ? p.ntile_m - first_m : p.group_m;
    const int64_t local = tile - group * per_group;
    mt = first_m + local % gsize;
    nt = (int)(local / gsize);
}
// work item -> (m-tile or m-tile pair, n-tile, k-chunk range)
__device__ __forceinline__ void item_coords(const TcGemmParams& p, int64_t item, int64_t& mt, int& nt, int& ks, int& kc_lo,
                                            int& kc_hi) {
    ks = (int)(item % p.ksplit);
    tile_coords(p, item / p.ksplit, mt, nt);
    kc_lo = ks * p.kc_per_split;
    kc_hi = min(p.nkc, kc_lo + p.kc_per_split);
}

// r[j] = racc[base + j], j < N, for a run-time `base`: a switch over the 128 possible bases whose
// cases index the register sums with compile-time constants.  (Indexing racc with a run-time value
// would move the whole array to local memory -- which is what the first version of the dX epilogue
// did: 15.7 GB of L1->L2 write-through traffic per pass at C2.)  `base` differs between the lanes of a
// warp (it depends on the sample's segment), so a warp walks one case per distinct segment.  The
// moves are volatile asm so that the compiler keeps the branches instead of evaluating all 128
// cases and selecting.
#define TC_DX_CASE(k)                                                                                  \
    case (k): {                                                                                        \
        _Pragma("unroll") for (int j = 0; j < N; ++j)                                                  \
            asm volatile("mov.b32 %0, %1;" : "=f"(r[j]) : "f"(racc[((k) + j) < TC_MAXHALF ? ((k) + j) : 0])); \
        break;                                                                                         \
    }
#define TC_DX_CASE8(k) TC_DX_CASE(k) TC_DX_CASE(k + 1) TC_DX_CASE(k + 2) TC_DX_CASE(k + 3) TC_DX_CASE(k + 4) TC_DX_CASE(k + 5) TC_DX_CASE(k + 6) TC_DX_CASE(k + 7)
#define TC_DX_CASE32(k) TC_DX_CASE8(k) TC_DX_CASE8(k + 8) TC_DX_CASE8(k + 16) TC_DX_CASE8(k + 24)
template <int N>
__device__ __forceinline__ void dx_pick(const float (&racc)[TC_MAXHALF], int base, float (&r)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = 0.0f;
    switch (base) {
        TC_DX_CASE32(0)
        TC_DX_CASE32(32)
        TC_DX_CASE32(64)
        TC_DX_CASE32(96)
    }
}

// Persistent: grid = min(#items, #SMs); every role walks the same work-item sequence
// item = blockIdx.x, blockIdx.x + gridDim.x, ...  The smem ring (full/empty) and the two TMEM
// accumulators (tmem_full/tmem_empty) run continuously across items, so the epilogue of tile t
// overlaps the main loop of tile t+1.
// Warps: 0 = producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle (warpgroup 0, trimmed to 56
// registers per thread with setmaxnreg), 4..11 = epilogue (warpgroups 1-2, raised to 224 registers:
// 128 running sums + the gather state without spills).  N = basis size for the EPI_DX gather.
// AMODE: where the A operand of a stage comes from
//   A_TILES  one bulk copy of a ready-made K-major tile image (Phi, Phi^T or gy tiles)
//   A_PHI_MN EPI_DW only: the FORWARD's Phi tiles [128 samples][64 dense columns] read as the
//            MN-major operand Phi^T: a stage = 128 dense columns (two forward k-chunks) x 64
//            samples (half a batch tile), gathered by the producer warp with 16-byte cp.async copies
//            into the canonical MN-major no-swizzle layout
//            [image][column group of 8][64 samples][8 x fp16]  (SBO = 1024 B, LBO = 128 B)
// CL = 2: the kernel runs as clusters of two CTAs that own two NEIGHBOURING m-tiles of the same n-tile.
// Both need the same B tile: each CTA fetches one of its two fp16 images (hi / lo) and multicasts it
// into both shared memories, so a k-chunk costs every CTA 32 KB (A) + 32 KB (half of B) of L2 reads
// instead of 32 + 64 -- these GEMMs run at the L2 -> SM bandwidth cap (~42 B/clk/SM: 96 KB per 1536-cycle
// MMA chunk is more than the cap delivers), so the bytes per chunk are what sets their speed.  A stage
// may only be refilled when BOTH CTAs' MMAs are done with it (the peer writes into it), hence the
// `empty` barriers count two commits, multicast by each MMA warp to both CTAs.
enum { A_TILES = 0, A_PHI_MN = 2 };
#define TC_THREADS 384

template <int EPI, int N, int AMODE, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1) pw_tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
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
    // EPI_DX: d xl / d x per segment (the reference's length / (eta(s+1) - eta(s)), same fp32 op order), once
    // per CTA: three IEEE divisions per (sample, input) in the epilogue's dependent chain otherwise
    float* chain_tab = reinterpret_cast<float*>(smem_raw + 256);  // [nseg <= 128]
    unsigned char* stages = smem_raw + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;  // this CTA's m-tile inside the pair
    const int64_t cid = blockIdx.x / CL, ncl = gridDim.x / CL;  // cluster id / number of clusters
    const int nst = p.nst;
    const int FG = p.flush_chunks;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;  // two accumulators; power of two >= 32

    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) {
            // A_PHI_MN: the bulk copy of B (expect_tx arrival of lane 0) + one asynchronous arrival per
            // producer lane when its cp.async copies of the stage have landed
            mbar_init(&full[s], AMODE == A_PHI_MN ? 33 : 1);
            mbar_init(&empty[s], CL);  // one tcgen05.commit per CTA of the cluster
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (EPI == EPI_DX && threadIdx.x >= 128 && (int)threadIdx.x - 128 < p.nseg && threadIdx.x < 128 + 160) {
        const int s = (int)threadIdx.x - 128;
        const float e0 = __fsub_rn(__fmul_rn(__fdiv_rn((float)s, p.Sf), 2.0f), 1.0f);
        const float e1 = __fsub_rn(__fmul_rn(__fdiv_rn((float)(s + 1), p.Sf), 2.0f), 1.0f);
        chain_tab[s] = __fdiv_rn(p.length, __fsub_rn(e1, e0));
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Register re-allocation between the warpgroups (the launch gives every thread 168).  All four warps
    // of a warpgroup execute the SAME setmaxnreg instruction (.aligned), and the roles are dispatched
    // INSIDE the two branches: ptxas budgets a region by the setmaxnreg that dominates it, so a join of
    // the two settings before the role dispatch would cap every role at the smaller one.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && AMODE == A_PHI_MN) {
        // The whole producer warp gathers the A half of a stage with 16-byte cp.async copies (2048
        // per stage, 64 per lane, 512 contiguous bytes per warp instruction): the pieces that are
        // contiguous in the forward's tile storage are only 1 KB long, and 32 bulk copies of 1 KB
        // (14.6 ms at C2) or four tiled TMA boxes with 16-byte rows (17.2 ms) cost more than the
        // MMAs of the stage.  The copies of a stage complete asynchronously on its `full` barrier
        // (cp.async.mbarrier.arrive.noinc), so the warp runs ahead over the whole ring instead of
        // waiting for every stage to land before issuing the next one.
        int it = 0;
        for (int64_t item = cid; item < p.nitems; item += ncl) {
            int nt, ks, kc_lo, kc_hi;
            int64_t mt;
            item_coords(p, item, mt, nt, ks, kc_lo, kc_hi);
            mt = min(mt * CL + (int64_t)crank, p.mt_last);  // this CTA's m-tile (clamped: a pair may end on a missing tile)
            const unsigned char* a_mt = p.a_tiles + (2 * mt) * (int64_t)(2 * TC_A_IMG) + lane * 16;
            const int64_t a_bt = (int64_t)p.a_nkc * (2 * TC_A_IMG);  // bytes per batch tile
            const unsigned char* b_src = p.b_tiles + ((int64_t)nt * p.nkc) * (int64_t)(2 * B_IMG);
            for (int sc = kc_lo; sc < kc_hi; ++sc, ++it) {
                const int st = it % nst;
                mbar_wait(&empty[st], (uint32_t)(((it / nst) & 1) ^ 1));
                unsigned char* dst = stages + (size_t)st * STAGE_BYTES;
                if (lane == 0) {
                    mbar_expect_tx(&full[st], (uint32_t)(2 * B_IMG));
                    if (CL == 2)  // this CTA's image of the B tile (hi: rank 0, lo: rank 1) goes to both CTAs
                        bulk_g2s_multicast(dst + 2 * TC_A_IMG + crank * B_IMG, b_src + (int64_t)sc * (2 * B_IMG) + crank * B_IMG,
                                           (uint32_t)B_IMG, &full[st], (uint16_t)3);
                    else
                        bulk_g2s(dst + 2 * TC_A_IMG, b_src + (int64_t)sc * (2 * B_IMG), (uint32_t)(2 * B_IMG), &full[st]);
                }
                // samples 64*(sc & 1) .. +63 of batch tile sc >> 1
                const unsigned char* src = a_mt + (int64_t)(sc >> 1) * a_bt + (sc & 1) * 1024;
                const uint32_t d0 = smem_u32(dst) + lane * 16;
#pragma unroll
                for (int q = 0; q < 64; ++q) {
                    // unit = (image, column group g = 8*c2 + kg, 32-row half)
                    const int img = q >> 5, g = (q & 31) >> 1, c2 = g >> 3, kg = g & 7, rr = (q & 1) * 32;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + img * TC_A_IMG + g * 1024 + rr * 16),
                                 "l"(src + (c2 * 2 + img) * TC_A_IMG + kg * (TC_BM * 16) + rr * 16)
                                 : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[st])) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");  // nothing in flight when the CTA exits
        if (CL == 2)  // tail: every commit the peer's MMA warp multicasts to this CTA's barriers has landed
            for (int k = 0; k < nst; ++k, ++it) mbar_wait(&empty[it % nst], (uint32_t)(((it / nst) & 1) ^ 1));
    } else if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int64_t item = cid; item < p.nitems; item += ncl) {
                int nt, ks, kc_lo, kc_hi;
                int64_t mt;
                item_coords(p, item, mt, nt, ks, kc_lo, kc_hi);
                mt = min(mt * CL + (int64_t)crank, p.mt_last);
                const unsigned char* a_src = p.a_tiles + (mt * p.a_nkc) * (int64_t)(2 * TC_A_IMG);
                const unsigned char* b_src = p.b_tiles + ((int64_t)nt * p.nkc) * (int64_t)(2 * B_IMG);
                for (int kc = kc_lo; kc < kc_hi; ++kc, ++it) {
                    const int st = it % nst;
                    mbar_wait(&empty[st], (uint32_t)(((it / nst) & 1) ^ 1));
                    unsigned char* dst = stages + (size_t)st * STAGE_BYTES;
                    mbar_expect_tx(&full[st], (uint32_t)STAGE_BYTES);
                    bulk_g2s(dst, a_src + (int64_t)kc * (2 * TC_A_IMG), 2 * TC_A_IMG, &full[st]);
                    if (CL == 2)
                        bulk_g2s_multicast(dst + 2 * TC_A_IMG + crank * B_IMG, b_src + (int64_t)kc * (2 * B_IMG) + crank * B_IMG,
                                           (uint32_t)B_IMG, &full[st], (uint16_t)3);
                    else
                        bulk_g2s(dst + 2 * TC_A_IMG, b_src + (int64_t)kc * (2 * B_IMG), (uint32_t)(2 * B_IMG), &full[st]);
                }
            }
            if (CL == 2)  // tail: every commit the peer's MMA warp multicasts to this CTA's barriers has landed
                for (int k = 0; k < nst; ++k, ++it) mbar_wait(&empty[it % nst], (uint32_t)(((it / nst) & 1) ^ 1));
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
            // (A_PHI_MN: a_major bit 15 = MN-major)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24) |
                                   (AMODE == A_PHI_MN ? (1u << 15) : 0u);
            // LBO = byte stride between core matrices along K, SBO = along M / N
            const uint32_t LBO_A = AMODE == A_PHI_MN ? 128u : TC_BM * 16, SBO_A = AMODE == A_PHI_MN ? 1024u : 128u;
            const uint32_t LBO_B = (uint32_t)BN * 16, SBO = 128;
            int it = 0, g = 0;
            for (int64_t item = cid; item < p.nitems; item += ncl) {
                const int ks = (int)(item % p.ksplit);
                const int kc_lo = ks * p.kc_per_split, kc_hi = min(p.nkc, kc_lo + p.kc_per_split);
                for (int kc = kc_lo; kc < kc_hi; ++kc, ++it) {
                    const int st = it % nst;
                    const int acc = g & 1;
                    const int rel = kc - kc_lo;
                    const bool first = (rel % FG) == 0;
                    if (first) {
                        mbar_wait(&tmem_empty[acc], (uint32_t)(((g >> 1) & 1) ^ 1));  // drained by the epilogue
                        tc_fence_after();
                    }
                    mbar_wait(&full[st], (uint32_t)((it / nst) & 1));
                    if (AMODE == A_PHI_MN) fence_proxy_async_smem();  // the A half was written by cp.async (generic proxy)
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stages + (size_t)st * STAGE_BYTES);
                    const uint32_t sb = sa + 2 * TC_A_IMG;
                    const uint32_t td = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 16; ++kk) {
                        const uint64_t a_hi = umma_desc_kmajor_noswizzle(sa + kk * 2 * LBO_A, LBO_A, SBO_A);
                        const uint64_t a_lo = umma_desc_kmajor_noswizzle(sa + TC_A_IMG + kk * 2 * LBO_A, LBO_A, SBO_A);
                        const uint64_t b_hi = umma_desc_kmajor_noswizzle(sb + kk * 2 * LBO_B, LBO_B, SBO);
                        const uint64_t b_lo = umma_desc_kmajor_noswizzle(sb + B_IMG + kk * 2 * LBO_B, LBO_B, SBO);
                        umma_f16(td, a_hi, b_hi, idesc, (first && kk == 0) ? 0u : 1u);
                        umma_f16(td, a_hi, b_lo, idesc, 1u);
                        umma_f16(td, a_lo, b_hi, idesc, 1u);
                    }
                    // arrives when the MMAs above have finished reading this stage -- on the peer's barrier too:
                    // the peer's producer writes its B image into this CTA's stage
                    if (CL == 2) umma_commit_multicast(&empty[st], (uint16_t)3);
                    else umma_commit(&empty[st]);
                    if ((rel % FG) == FG - 1 || kc == kc_hi - 1) {
                        umma_commit(&tmem_full[acc]);
                        ++g;
                    }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // epilogue warps 4..11: TMEM lane quadrant = warp % 4, column half = (warp - 4) / 4
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int rl = q * 32 + lane;  // row inside the 128-row tile
        int g = 0;
        // scale factors (device scalars written before this launch)
        float inv;
        if (EPI == EPI_Y)
            inv = 1.0f / (phi_scale(__uint_as_float(p.sc->max_xl), p.tab, p.n_basis) * pow2_scale_for(__uint_as_float(p.sc->max_w)));
        else if (EPI == EPI_DW)
            inv = 1.0f / (phi_scale(__uint_as_float(p.sc->max_xl), p.tab, p.n_basis) * pow2_scale_for(__uint_as_float(p.sc->max_gy)));
        else
            inv = 1.0f / (pow2_scale_for(__uint_as_float(p.sc->max_gy)) * pow2_scale_for(__uint_as_float(p.sc->max_w)));

        for (int64_t item = cid; item < p.nitems; item += ncl) {
            int nt, ks, kc_lo, kc_hi;
            int64_t mt;
            item_coords(p, item, mt, nt, ks, kc_lo, kc_hi);
            mt = mt * CL + (int64_t)crank;
            const bool tile_valid = mt <= p.mt_last;  // the second tile of the last pair may not exist
            const int ngroups = (kc_hi - kc_lo + FG - 1) / FG;
            float racc[TC_MAXHALF];
#pragma unroll
            for (int e = 0; e < TC_MAXHALF; ++e) racc[e] = 0.0f;
            // EPI_DX: fetch the coordinates of the first two inputs of this thread's columns now, so the
            // gather after the last drain does not wait on global memory while the MMA warp needs
            // the next drains
            uint32_t cur_sg[2] = {HOL_SEG_INVALID, HOL_SEG_INVALID};
            float cur_xl[2] = {0.0f, 0.0f};
            if (EPI == EPI_DX) {
                const int64_t b = p.row0 + mt * TC_BM + rl;
                const int i0 = (nt * BN + h * HALF) / p.Kin;
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (tile_valid && b < p.Bp && i0 + u < p.I && (u + 1) * p.Kin <= HALF) {
                        cur_sg[u] = p.seg_t[(int64_t)(i0 + u) * p.Bp + b];
                        cur_xl[u] = p.xl_t[(int64_t)(i0 + u) * p.Bp + b];
                    }
            }
            for (int gg = 0; gg < ngroups; ++gg, ++g) {
                const int acc = g & 1;
                mbar_wait(&tmem_full[acc], (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + h * HALF);
                if (EPI == EPI_DX) {  // 16 columns per load: the gather state leaves no room for 32 more registers
#pragma unroll
                    for (int c0 = 0; c0 < TC_MAXHALF; c0 += 16) {
                        if (c0 < HALF) {
                            uint32_t v[16];
                            tmem_ld16(t0 + (uint32_t)c0, v);
#pragma unroll
                            for (int e = 0; e < 16; ++e) racc[c0 + e] += __uint_as_float(v[e]);
                        }
                    }
                } else {
#pragma unroll
                    for (int c0 = 0; c0 < TC_MAXHALF; c0 += 32) {
                        if (c0 < HALF) {
                            uint32_t v[32];
                            tmem_ld32(t0 + (uint32_t)c0, v);
#pragma unroll
                            for (int e = 0; e < 32; ++e) racc[c0 + e] += __uint_as_float(v[e]);  // columns >= HALF: unused
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            const int64_t row = mt * TC_BM + rl;  // row inside this launch
            const int colbase = nt * BN + h * HALF;

            if (!tile_valid) {
                // nothing to store
            } else if (EPI == EPI_Y) {
                if (row < p.rows_valid) {
                    // (a forward split over K writes this split's partial result: summed by pw_dw_reduce_kernel)
                    float* yrow = p.out + (int64_t)ks * p.part_stride + (p.row0 + row) * (int64_t)p.ld + colbase;
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
                // (or, when the batch is split over several work items, this split's partial result)
                const int i = (int)(row / p.Kin), v = (int)(row % p.Kin);
                if (i < p.I && v < p.nb) {
                    const bool accumulate = p.row0 > 0;
                    float* outp = p.out + (int64_t)ks * p.part_stride;
#pragma unroll
                    for (int e = 0; e < TC_MAXHALF; ++e) {
                        if (e < HALF) {
                            const int o = colbase + e;
                            if (o < p.n_valid) {
                                float* dst = outp + ((int64_t)o * p.I + i) * p.nb + v;
                                *dst = accumulate ? *dst + racc[e] * inv : racc[e] * inv;
                            }
                        }
                    }
                }
            } else {
                // EPI_DX: this thread holds T[b][k] for HALF/Kin whole inputs.  The n active columns of an
                // input depend on the sample's segment: dx_pick moves them out of the register sums
                // through a switch, so every register index stays static.
                constexpr int NB = N > 0 ? N : 1;
                const int64_t b = p.row0 + row;
                const int ninp = HALF / p.Kin;
                const int ibase = colbase / p.Kin;
                if (b < p.Bp) {
                    // Two inputs per trip, the coordinates of the next two requested before this pair is worked
                    // on: with one input at a time the trip is one dependent chain (global load -> basis
                    // derivative -> pick -> dot) and eight warps per SM cannot hide it -- at K = O = 128 the
                    // epilogue, not the MMAs, set the time of the launch (config 3 layer 2: 27 k cycles per tile).
#pragma unroll 1
                    for (int ii = 0; ii < ninp; ii += 2) {
                        if (ibase + ii >= p.I) break;
                        uint32_t nxt_sg[2] = {HOL_SEG_INVALID, HOL_SEG_INVALID};
                        float nxt_xl[2] = {0.0f, 0.0f};
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int in = ii + 2 + u;
                            if (in < ninp && ibase + in < p.I) {
                                nxt_sg[u] = p.seg_t[(int64_t)(ibase + in) * p.Bp + b];
                                nxt_xl[u] = p.xl_t[(int64_t)(ibase + in) * p.Bp + b];
                            }
                        }
                        float dphi[2][NB], tv[2][NB];
#pragma unroll
                        for (int u = 0; u < 2; ++u) lagrange_basis_derivative<NB>(cur_xl[u], p.tab, dphi[u]);
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int s = (int)(cur_sg[u] & HOL_SEG_MASK);
                            // (an input that does not exist carries HOL_SEG_INVALID with segment bits 0: base stays in range)
                            dx_pick<NB>(racc, (ii + u < ninp ? ii + u : 0) * p.Kin + s * p.stride, tv[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = ibase + ii + u;
                            if (ii + u < ninp && i < p.I) {
                                const uint32_t sg = cur_sg[u];
                                float sum = 0.0f;
#pragma unroll
                                for (int j = 0; j < NB; ++j) sum = fmaf(dphi[u][j], tv[u][j], sum);
                                float chain = chain_tab[min((int)(sg & HOL_SEG_MASK), p.nseg - 1)];
                                if (sg & HOL_SEG_MIRROR) chain = -chain;
                                // outliers are written by the fp32 fix-up kernel after this launch
                                const float outv = (sg & (HOL_SEG_INVALID | HOL_SEG_OUTLIER)) ? 0.0f : sum * inv * chain;
                                if (p.transposed)
                                    p.out[(int64_t)i * p.Bp + b] = outv;
                                else if (row < p.rows_valid)
                                    p.out[b * p.I + i] = outv;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u) cur_sg[u] = nxt_sg[u], cur_xl[u] = nxt_xl[u];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it / arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}
