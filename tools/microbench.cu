// Micro-benchmarks that back two design assumptions of the contraction kernels:
//  (1) an LDS.128 whose 32 lanes read k distinct, bank-disjoint 16-byte chunks costs ~max(1,k/8)
//      shared-memory wavefronts (broadcast serves lanes that share a segment row);
//  (2) FFMA2 (fma.rn.f32x2) issues two FMAs per lane per instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lds_bcast(const int* __restrict__ rowsel, float* out, long long* cyc, int iters, int pitch4) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 64 * pitch4; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int row = rowsel[threadIdx.x & 31];
    const float4* base = sm + row * pitch4;
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float4 v = base[q];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        base = sm + ((row + (it & 1)) % 64) * pitch4;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <bool PACKED>
__global__ void fma_rate(float* out, long long* cyc, int iters, float a) {
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x + i, i);
    float2 bb = make_float2(a, a * 0.5f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (PACKED) acc[i] = __ffma2_rn(make_float2(a, a), bb, acc[i]);
            else { acc[i].x = fmaf(a, bb.x, acc[i].x); acc[i].y = fmaf(a, bb.y, acc[i].y); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out; long long* cyc; int* rowsel;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096 * 8); cudaMalloc(&rowsel, 128);
    const int pitch4 = 17;  // 68 floats: the OT=64 slab pitch
    const int iters = 2000;
    cudaFuncSetAttribute(lds_bcast, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * pitch4 * 16);
    printf("LDS.128 broadcast test: 512 threads/CTA, 148 CTAs, 16 LDS.128 per iteration\n");
    for (int k : {1, 2, 4, 8, 16, 32}) {
        int h[32];
        for (int l = 0; l < 32; ++l) h[l] = (l * 7) % k;  // k distinct consecutive rows, shuffled over lanes
        cudaMemcpy(rowsel, h, sizeof(h), cudaMemcpyHostToDevice);
        lds_bcast<<<148, 512, 64 * pitch4 * 16>>>(rowsel, out, cyc, iters, pitch4);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // 16 warps x 16 LDS.128 per iteration per SM
        printf("  distinct rows %2d : %.2f cycles per warp-LDS.128 (SM-wide)\n", k, (double)c / iters / (16.0 * 16.0));
    }
    for (int packed = 0; packed < 2; ++packed) {
        if (packed) fma_rate<true><<<148, 512>>>(out, cyc, iters, 1.0001f);
        else fma_rate<false><<<148, 512>>>(out, cyc, iters, 1.0001f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // per iteration per SM: 16 warps x 32 lanes x 32 FMAs
        printf("%s: %.1f FMA/clk/SM (peak 128)\n", packed ? "FFMA2" : "FFMA ", 16.0 * 32 * 32 * iters / (double)c);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
