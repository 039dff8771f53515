// Data-parallel exchange step: sum (or mean) all-reduce of a slice of the flat gradient buffer over
// NVLink peer memory -- our own kernels instead of a library collective.
//
// Every rank's flat buffer lives in symmetric memory (the same allocation mapped into every process of
// the node, peer pointers handed in by the host), so a rank reads its peers' gradients with plain loads
// over NVLink 5 / NVSwitch and writes results with plain stores.  Two-shot, owner-computes:
//   barrier-in   every rank's dW kernels for this slice have finished (flag exchange, release/acquire.sys)
//   reduce       rank r owns the r-th 1/N of the slice: it loads that part from all N buffers, adds them in
//                RANK ORDER (bit-reproducible, identical on every rank), scales, and stores the result into
//                all N buffers.  Only rank r ever touches part r of any buffer, so nothing races.
//   barrier-out  every rank's stores have landed; the slice is complete everywhere.
// Traffic per rank: (N-1)/N of the slice in, the same out -- 9 MB for the 5.2 MB first-layer gradient of
// the MLP at N = 8, ~13 us at 700 GB/s per direction; at that size the exchange is latency-bound (two
// flag round trips over NVLink + three launches): measured 31 / 42 us at 2 / 4 GPUs against 38 / 42 us for
// NCCL's all-reduce, and inside the config-3 step 0.796 ms at 8 GPUs against 0.818 ms with NCCL.
// Three tiny launches on the caller's stream (stream order is the grid-wide sync between the phases);
// the flags carry a device-side epoch, so the sequence can be captured in a CUDA graph and replayed.
// A rank that never shows up must not hang the others forever: the flag waits give up after ~4 s and
// raise a sticky error word the host checks (hol_p2p_error).
#include "kernels.cuh"

namespace {

struct P2pFlags {          // one per rank, in symmetric memory, zero-initialised
    unsigned in[16];       // in[p]  = epoch of the last barrier-in  rank p has announced to this rank
    unsigned out[16];      // out[p] = ... barrier-out
    unsigned epoch;        // this rank's call counter (device side: valid under graph replay)
    unsigned error;        // sticky: a wait timed out
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// phase 0: barrier-in (advance == 1, starts a new epoch); phase 1: barrier-out
__global__ void __launch_bounds__(32) pw_p2p_barrier_kernel(P2pFlags* const* flags, int rank, int world, int phase) {
    __shared__ unsigned e;
    P2pFlags* mine = flags[rank];
    if (threadIdx.x == 0) {
        e = mine->epoch + (phase == 0 ? 1u : 0u);
        mine->epoch = e;
    }
    __syncwarp();
    __threadfence_system();
    const int p = threadIdx.x;
    if (p < world) {
        unsigned* theirs = phase == 0 ? &flags[p]->in[rank] : &flags[p]->out[rank];
        st_release_sys(theirs, e);
        const unsigned* slot = phase == 0 ? &mine->in[p] : &mine->out[p];
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(slot) - e) < 0) {
            if (clock64() - t0 > 8000000000ll) {  // ~4 s: a peer is gone
                mine->error = 1u;
                break;
            }
        }
    }
}

__global__ void __launch_bounds__(256) pw_p2p_reduce_kernel(float* const* bufs, int rank, int world, int64_t offset,
                                                            int64_t count, float scale) {
    // this rank's part of the slice, in float4 units where alignment allows
    const int64_t per = (count + world - 1) / world;
    const int64_t lo = offset + (int64_t)rank * per;
    int64_t hi = lo + per;
    if (hi > offset + count) hi = offset + count;
    if (lo >= hi) return;
    const int64_t head = (4 - (lo & 3)) & 3;  // scalars up to the next 16-byte boundary (buffers are 16-byte aligned)
    const int64_t v_lo = lo + head < hi ? lo + head : hi;
    const int64_t nvec = (hi - v_lo) / 4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    // All `world` peer loads of an element are issued before the first add (a load over NVLink takes ~2 us;
    // a load-add chain over the peers would serialise them).  The sum itself still runs in rank order.
    for (int64_t k = tid; k < nvec; k += nthr) {
        const int64_t e = v_lo + 4 * k;
        float4 v[16];
#pragma unroll
        for (int p = 0; p < 16; ++p)
            if (p < world) v[p] = __ldcg(reinterpret_cast<const float4*>(bufs[p] + e));
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < 16; ++p)
            if (p < world) s.x += v[p].x, s.y += v[p].y, s.z += v[p].z, s.w += v[p].w;
        s.x *= scale, s.y *= scale, s.z *= scale, s.w *= scale;
#pragma unroll
        for (int p = 0; p < 16; ++p)
            if (p < world) *reinterpret_cast<float4*>(bufs[p] + e) = s;
    }
    // ragged ends
    const int64_t tail0 = v_lo + 4 * nvec;
    const int64_t nscal = (v_lo - lo) + (hi - tail0);
    for (int64_t k = tid; k < nscal; k += nthr) {
        const int64_t e = k < (v_lo - lo) ? lo + k : tail0 + (k - (v_lo - lo));
        float s = __ldcg(bufs[0] + e);
        for (int p = 1; p < world; ++p) s += __ldcg(bufs[p] + e);
        s *= scale;
        for (int p = 0; p < world; ++p) bufs[p][e] = s;
    }
    __threadfence_system();
}

}  // namespace

extern "C" {

size_t hol_p2p_flags_bytes(void) { return sizeof(P2pFlags); }

int hol_p2p_allreduce_f32(void* const* bufs_dev, void* const* flags_dev, int32_t rank, int32_t world, int64_t offset,
                          int64_t count, float scale, void* stream) {
    if (!bufs_dev || !flags_dev || world < 1 || world > 16 || rank < 0 || rank >= world || offset < 0 || count < 0)
        return HOL_EINVAL;
    if (count == 0 || world == 1) return HOL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    P2pFlags* const* flags = reinterpret_cast<P2pFlags* const*>(flags_dev);
    float* const* bufs = reinterpret_cast<float* const*>(bufs_dev);
    pw_p2p_barrier_kernel<<<1, 32, 0, st>>>(flags, rank, world, 0);
    const int64_t per = (count + world - 1) / world;
    int64_t blocks = (per / 4 + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    pw_p2p_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(bufs, rank, world, offset, count, scale);
    pw_p2p_barrier_kernel<<<1, 32, 0, st>>>(flags, rank, world, 1);
    return (int)cudaGetLastError();
}

}  // extern "C"
