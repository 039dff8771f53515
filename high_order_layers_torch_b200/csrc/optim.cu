// SparseLion step (reference sparse_optimizers/sparse_lion.py:13-92) as ONE fused
// elementwise kernel per parameter instead of ~10 masked-indexing launches (each of which builds
// an index list of the non-zero gradient entries):
//   touched = grad != 0                     (the layers' dW is exactly zero for untouched segments)
//   update  = sign(beta1 * exp_avg + (1 - beta1) * grad)
//   p       = p - lr * update
//   exp_avg = beta2 * exp_avg + (1 - beta2) * grad
// all only where `touched`.  The reference's weight-decay line multiplies a COPY
// (`p.data[mask].mul_(...)`: advanced indexing copies), i.e. it has no effect; that behaviour is
// kept (wd is accepted and ignored) so that the two implementations stay weight-for-weight equal.
// HBM-bound: reads p, grad, exp_avg once, writes p and exp_avg only where touched.
#include "kernels.cuh"

__global__ void __launch_bounds__(256) pw_sparse_lion_kernel(float* __restrict__ p, const float* __restrict__ grad,
                                                             float* __restrict__ exp_avg, int64_t count, float lr,
                                                             float beta1, float omb1, float beta2, float omb2) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const float g = grad[k];
        if (g == 0.0f) continue;  // NaN != 0: updated like the reference
        const float m = exp_avg[k];
        const float c = fmaf(omb1, g, __fmul_rn(m, beta1));
        const float u = c > 0.0f ? 1.0f : (c < 0.0f ? -1.0f : (c == 0.0f ? 0.0f : c));
        p[k] = fmaf(-lr, u, p[k]);
        exp_avg[k] = fmaf(omb2, g, __fmul_rn(m, beta2));
    }
}

extern "C" int hol_sparse_lion_step_f32(float* p, const float* grad, float* exp_avg, int64_t count, float lr, float wd,
                                        float beta1, float beta2, void* stream) {
    (void)wd;  // no effect in the reference (see above)
    if (count < 0) return HOL_EINVAL;
    if (count == 0) return HOL_OK;
    if (!p || !grad || !exp_avg) return HOL_EINVAL;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const float omb1 = (float)(1.0 - (double)beta1), omb2 = (float)(1.0 - (double)beta2);
    pw_sparse_lion_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, grad, exp_avg, count, lr, beta1, omb1,
                                                                              beta2, omb2);
    return (int)cudaGetLastError();
}
