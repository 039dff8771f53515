/*
 * hol_b200.h -- C ABI of the B200-native piecewise Lagrange-polynomial layer kernels.
 *
 * Drop-in boundary for the hot path of jloveric/high-order-layers-torch v2.6.0.  The
 * reference is pure Python; the "FFI" these entry points replace is the chain of ATen
 * calls issued by (paths relative to the reference checkout)
 *
 *   Piecewise.forward / PiecewiseDiscontinuous.forward
 *       high_order_layers_torch/PolynomialLayers.py:294-336, :639-676
 *   Basis.interpolate                   high_order_layers_torch/Basis.py:492-512
 *   LagrangeBasis.__call__              high_order_layers_torch/LagrangePolynomial.py:205-211
 *   make_periodic                       high_order_layers_torch/utils.py:74-79
 *   PiecewiseExpand.__call__ + Expansion2d + nn.Conv2d
 *       Basis.py:67-127, FunctionalConvolution.py:123-139, :441-447
 *
 * and their autograd-generated backward.  The binding a maintainer adds on the reference
 * side (ctypes + torch.library) is shown in INTEGRATION.md and shipped in
 * high_order_layers_torch_b200/_lib.py and ops.py.
 *
 * Conventions (all entry points)
 *   - return 0 on success, a HOL_E* code (< 0) or a cudaError_t (> 0) otherwise;
 *     hol_error_string() explains either;
 *   - never throw, never allocate device memory, never synchronise, never touch the
 *     default stream: every launch goes to `stream` (a cudaStream_t passed as void*);
 *   - every buffer is a caller-owned DEVICE pointer, contiguous row-major, 16-byte aligned
 *     (torch's caching allocator gives 512-byte alignment);
 *   - re-entrant (the node table is an argument); safe to call from several host threads on
 *     different streams.  The only process-wide state is the diagnostic plan-override mask of
 *     hol_set_plan_overrides (0 unless a test sets it) and write-once caches (SM count, "shared
 *     memory opt-in done" bits); no environment variable is read;
 *   - fp32 arithmetic, fp32 accumulation, no fast-math; segment indices follow the
 *     reference's three separately rounded fp32 ops + truncation + clamp bit for bit
 *     (CPU semantics: true divisions), finite inputs only (NaN -> segment 0; +-inf is
 *     undefined behaviour in the reference).
 *
 * Shapes: x[B, I], w[O, I, nb], y / gy[B, O], dx[B, I], dw[O, I, nb] with
 *   nb = (n-1)*S + 1 (continuous: neighbouring segments share their end node) or
 *   nb = n*S         (discontinuous),   1 <= n <= HOL_MAX_N, 1 <= S <= HOL_MAX_SEGMENTS.
 * `nodes` are the n fp32 interpolation nodes built on the host with the reference's own
 * expression (LagrangePolynomial.py:9-22, :191-195); they stay a HOST pointer.
 * `periodicity` is NaN for "None".
 */
#ifndef HOL_B200_H
#define HOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOL_ABI_VERSION 1
#define HOL_MAX_N 10
#define HOL_MAX_SEGMENTS 4096

#define HOL_NORM_MAX_ABS 0     /* per-sample normalisation kinds (utils.py:8-58) */
#define HOL_NORM_MAX_CENTER 1
#define HOL_NORM_L2 2

#define HOL_OK 0
#define HOL_EINVAL (-1)        /* bad argument (null pointer, non-positive size, ...)      */
#define HOL_EUNSUPPORTED (-2)  /* valid request outside the compiled envelope (n, S, smem) */
#define HOL_EWORKSPACE (-3)    /* workspace too small (see hol_pw_workspace_bytes)          */

/* flags for hol_pw_backward_f32 */
#define HOL_WS_PREPARED 1u     /* `ws` still holds the forward's prepared state for the very
                                  same (x, w, shape): skip the prep + pack stages           */
#define HOL_COUNT_FUSED_NORM 4u /* hol_pw_launch_count only: the layer runs with a folded normalisation */
#define HOL_GY_STATS_VALID 2u  /* `ws` still holds max|gy| of this very gy, left by a previous
                                  backward call of the same step (a backward split into a dw
                                  call and a dx call takes the statistics once)              */

int hol_abi_version(void);
const char* hol_error_string(int code);

/* DIAGNOSTIC: force kernel families for the calls that follow (tests, A/B measurements); returns the
 * previous mask.  0 = the planner's own choice.  Must not change between a forward and the backward
 * that reuses its workspace (the workspace layout depends on the plan). */
#define HOL_PLAN_DISABLE_TENSOR_CORES 1u /* gather kernels everywhere                                */
#define HOL_PLAN_DISABLE_DENSE_DW 2u     /* bucketed dW instead of the dense-register dW              */
#define HOL_PLAN_NO_SHARED_PHI 4u        /* chunked Phi / Phi^T tiles instead of the shared Phi tiles */
#define HOL_PLAN_NO_CLUSTER 16u          /* single-CTA GEMM tiles instead of two-CTA multicast clusters */
#define HOL_PLAN_FLUSH_SHIFT 8           /* bits 8..11: k-chunks per TMEM drain (0 = default 4)       */
uint32_t hol_set_plan_overrides(uint32_t mask);

/* Bytes of scratch the forward/backward pair needs for one layer call (same for both). */
size_t hol_pw_workspace_bytes(int64_t B, int32_t I, int32_t O, int32_t n, int32_t S,
                              int32_t discontinuous);

/* Piecewise.which_segment (PolynomialLayers.py:259-275): seg[k] = clamp(trunc(((x+half)/
 * length)*S), 0, S-1) as int64, after the optional mirror-periodic fold.  Elementwise. */
int hol_pw_segment_index_i64(const float* x, int64_t* seg, int64_t count, int32_t S,
                             float length, float periodicity, void* stream);

/* Piecewise.forward: y = sum_i sum_j phi_j(x_in(b,i)) * w[o, i, s(b,i)*stride + j].
 * Leaves the prepared state (transposed local coordinates, packed weights) in `ws`. */
int hol_pw_forward_f32(const float* x, const float* w, const float* nodes_host, float* y,
                       int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, float length,
                       int32_t discontinuous, float periodicity, void* ws, size_t ws_bytes,
                       void* stream);

/* Backward of the above.  dx and/or dw may be NULL (gradient not needed).  dw is written
 * densely (exact zeros where no sample touched a weight), dx includes the chain factors
 * length/(eta(s+1)-eta(s)) and the +-1 of the periodic fold; dx == 0 for n == 1.  dw is computed
 * before dx, so a caller may split the call in two (dw only, then dx only, both with
 * HOL_WS_PREPARED) and start the data-parallel all-reduce of dw while dx is still running.
 * B == 0: dw (if requested) is zero-filled and nothing else is touched (pointers may be NULL). */
int hol_pw_backward_f32(const float* gy, const float* x, const float* w,
                        const float* nodes_host, float* dx, float* dw, int64_t B, int32_t I,
                        int32_t O, int32_t n, int32_t S, float length, int32_t discontinuous,
                        float periodicity, void* ws, size_t ws_bytes, uint32_t flags,
                        void* stream);

/* ---- a layer with the per-sample normalisation that follows it --------------------------
 * HighOrderMLP puts max_abs / max_center / l2 normalisation behind every layer but the last
 * (networks.py:235-270, layers.py:43-145).  Forward: as hol_pw_forward_f32, and the forward's LAST kernel
 * (the pass that sums the k-split partial results and adds outlier contributions on the tensor-core
 * path) also writes the normalised rows y_norm[B, O] and their statistics stats[B][4]; y_raw keeps the
 * layer's own output (the normalisation's backward needs it).  norm_kind: HOL_NORM_*.
 * Backward: hol_pw_norm_backward_f32 turns the gradient w.r.t. y_norm into gy_raw[B, O] and, in the same
 * sweep, takes the max|gy_raw| statistic of the layer's tensor-core passes into `ws`; follow it with
 * hol_pw_backward_f32(gy_raw, ..., flags = HOL_WS_PREPARED | HOL_GY_STATS_VALID). */
int hol_pw_forward_norm_f32(const float* x, const float* w, const float* nodes_host, float* y_raw, float* y_norm,
                            float* stats, int32_t norm_kind, float eps, int64_t B, int32_t I, int32_t O,
                            int32_t n, int32_t S, float length, int32_t discontinuous, float periodicity,
                            void* ws, size_t ws_bytes, void* stream);
int hol_pw_norm_backward_f32(int32_t norm_kind, const float* gy_norm, const float* y_raw, const float* stats,
                             float* gy_raw, int64_t B, int32_t I, int32_t O, int32_t n, int32_t S,
                             int32_t discontinuous, void* ws, size_t ws_bytes, void* stream);

/* ---- unfolded PiecewisePolynomialConvolution2d (FunctionalConvolution.py:441-447) -------
 * The convolution is the layer above applied to im2col rows: rows = Bimg*Ho*Wo,
 * I = C*K*K (column c*K*K + kh*K + kw), w_fc[O, I, nb] = conv.weight[O, C*nb, K, K] viewed
 * as [O, C, nb, K, K] and permuted to [O, C, K, K, nb].  The im2col matrix is never
 * materialised: the prep stage reads x[Bimg, C, H, W] directly; zero-padded taps contribute
 * nothing (the reference pads the *expanded* tensor).  y_rows[rows, O] / gy_rows[rows, O]
 * are row-major over (b, ho, wo); dx is NCHW like x. */
size_t hol_pw_conv2d_workspace_bytes(int64_t Bimg, int32_t C, int32_t H, int32_t W, int32_t O,
                                     int32_t K, int32_t stride, int32_t padding,
                                     int32_t dilation, int32_t n, int32_t S,
                                     int32_t discontinuous);

int hol_pw_conv2d_forward_f32(const float* x, const float* w_fc, const float* nodes_host,
                              float* y_rows, int64_t Bimg, int32_t C, int32_t H, int32_t W,
                              int32_t O, int32_t K, int32_t stride, int32_t padding,
                              int32_t dilation, int32_t n, int32_t S, float length,
                              int32_t discontinuous, float periodicity, void* ws,
                              size_t ws_bytes, void* stream);

int hol_pw_conv2d_backward_f32(const float* gy_rows, const float* x, const float* w_fc,
                               const float* nodes_host, float* dx, float* dw_fc, int64_t Bimg,
                               int32_t C, int32_t H, int32_t W, int32_t O, int32_t K,
                               int32_t stride, int32_t padding, int32_t dilation, int32_t n,
                               int32_t S, float length, int32_t discontinuous,
                               float periodicity, void* ws, size_t ws_bytes, uint32_t flags,
                               void* stream);

/* ---- rectangular / 1-d geometry (FunctionalConvolution.py:53-103 conv_wrapper kwargs; the 1-d layers
 * :493-519, :622-650 are H = Kh = 1) -----------------------------------------------------------------
 * geom = {C, H, W, Kh, Kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w} (HOST pointer, 11 values).
 * I = C*Kh*Kw with column c*Kh*Kw + kh*Kw + kw; everything else as the square entry points above, which
 * forward to these. */
size_t hol_pw_conv_workspace_bytes(int64_t Bimg, const int32_t* geom, int32_t O, int32_t n, int32_t S,
                                   int32_t discontinuous);
int hol_pw_conv_forward_f32(const float* x, const float* w_fc, const float* nodes_host, float* y_rows, int64_t Bimg,
                            const int32_t* geom, int32_t O, int32_t n, int32_t S, float length,
                            int32_t discontinuous, float periodicity, void* ws, size_t ws_bytes, void* stream);
int hol_pw_conv_backward_f32(const float* gy_rows, const float* x, const float* w_fc, const float* nodes_host,
                             float* dx, float* dw_fc, int64_t Bimg, const int32_t* geom, int32_t O, int32_t n,
                             int32_t S, float length, int32_t discontinuous, float periodicity, void* ws,
                             size_t ws_bytes, uint32_t flags, void* stream);

/* ---- conv-transpose (FunctionalConvolutionTranspose.py:75-160) --------------------------------------
 * Expansion + nn.ConvTranspose2d == the piecewise layer per input pixel (hol_pw_conv_forward_f32 with
 * a 1x1 kernel: rows = Bimg*H*W, I = C, O' = O*K*K, w_fc[O*K*K, C, nb]) followed by a fold of the K*K
 * taps onto the output grid: y[b, o, h*s - p + kh*d, w*s - p + kw*d] += z[(b,h,w)][o*K*K + kh*K + kw].
 * fold writes y[Bimg, O, Ho, Wo] (every element, gathered in a fixed order); unfold is its adjoint and
 * writes gz_rows[Bimg*H*W, O*K*K] from gy[Bimg, O, Ho, Wo]. */
int hol_fold_rows_f32(const float* z_rows, float* y, int64_t Bimg, int32_t O, int32_t H, int32_t W, int32_t Ho,
                      int32_t Wo, int32_t K, int32_t stride, int32_t padding, int32_t dilation, void* stream);
int hol_unfold_rows_f32(const float* gy, float* gz_rows, int64_t Bimg, int32_t O, int32_t H, int32_t W, int32_t Ho,
                        int32_t Wo, int32_t K, int32_t stride, int32_t padding, int32_t dilation, void* stream);

/* DIAGNOSTIC entry (the only one that synchronises): runs forward + backward once on `stream`
 * with CUDA events between the kernels and returns their durations in milliseconds in
 * ms_out_host[9] = {prep, pack, forward, dx, bucket_sort, dw, gemm_fwd, gemm_dx, gemm_dw}: the
 * first six bracket the stages (for a tensor-core layer "forward"/"dx"/"dw" include the operand
 * builders and the outlier fix-up), the last three are the tcgen05 GEMM launches alone (0 on the
 * gather path).  Used by bench.py for the per-kernel roofline numbers; not part of the drop-in path. */
int hol_pw_profile_f32(const float* gy, const float* x, const float* w, const float* nodes_host,
                       float* y, float* dx, float* dw, int64_t B, int32_t I, int32_t O, int32_t n,
                       int32_t S, float length, int32_t discontinuous, float periodicity, void* ws,
                       size_t ws_bytes, float* ms_out_host, void* stream);

/* Number of kernels of this library one call launches for the given shape (what bench.py
 * reports as gpu_launches).  kind: 0 forward, 1 backward (want_dx / want_dw select the parts). */
int hol_pw_launch_count(int kind, int64_t B, int32_t I, int32_t O, int32_t n, int32_t S,
                        int32_t discontinuous, int32_t want_dx, int32_t want_dw, uint32_t flags);

/* Which kernels the planner picks for a layer shape (diagnostic, used by bench.py): bit 0 / 1 / 2 =
 * forward / dX / dW on the tcgen05 path, bit 3 = dense-register dW (narrow layers), bit 4 = the
 * forward's Phi tiles are shared with dW, bit 5 = lane-rotated gather kernels (S > 8), bits 8..15 /
 * 16..23 = k-split of the tensor-core dW / forward GEMM; -1 = bad shape. */
int hol_pw_plan_flags(int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, int32_t discontinuous);

/* ---- per-sample normalisations between layers ------------------------------
 * One fused kernel each way for the reference's max_abs_normalization (utils.py:8-13, used by
 * layers.py:70-81 MaxAbsNormalization), max_center_normalization (utils.py:20-28, layers.py:84-96)
 * and l2_normalization (utils.py:57-58, layers.py:132-141) over dim 1 of x[B, W].
 * kind: 0 max_abs, 1 max_center, 2 l2.  stats[B][4] (fp32) is written by the forward and read by
 * the backward (denominator, centre / norm, arg-max / arg-min as int bits). */
int hol_row_norm_forward_f32(int32_t kind, const float* x, float* y, float* stats, int64_t B,
                             int32_t W, float eps, void* stream);
int hol_row_norm_backward_f32(int32_t kind, const float* gy, const float* x, const float* stats,
                              float* dx, int64_t B, int32_t W, void* stream);

/* ---- SparseLion step (reference sparse_optimizers/sparse_lion.py:13-92) --------
 * One fused elementwise kernel: where grad != 0, update = sign(beta1*exp_avg + (1-beta1)*grad),
 * p -= lr*update, exp_avg = beta2*exp_avg + (1-beta2)*grad.  `wd` is accepted and, exactly like
 * the reference (whose decay line acts on a copy), has no effect.  In place on p and exp_avg. */
int hol_sparse_lion_step_f32(float* p, const float* grad, float* exp_avg, int64_t count, float lr,
                             float wd, float beta1, float beta2, void* stream);

/* ---- data-parallel exchange step over NVLink peer memory ---------------------------------
 * Sum (scale = 1) or mean (scale = 1/world) all-reduce of `count` floats at `offset` of the flat gradient
 * buffer, by our own kernels: every rank's buffer lives in symmetric memory (one allocation per rank,
 * mapped into every process of the node).  bufs_dev / flags_dev are DEVICE arrays of `world` device
 * pointers -- rank p's buffer and rank p's zero-initialised flag block of hol_p2p_flags_bytes() bytes --
 * in the same order on every rank.  Every rank of the group must make the same calls in the same order.
 * Deterministic (rank-order sum), graph-capturable (device-side epochs), never blocks the host; a peer
 * that never arrives makes the flag waits give up after ~4 s and sets the error word
 * (uint32 index 33: 16 + 16 flag words, epoch, error) of the caller's flag block. */
size_t hol_p2p_flags_bytes(void);
int hol_p2p_allreduce_f32(void* const* bufs_dev, void* const* flags_dev, int32_t rank, int32_t world, int64_t offset,
                          int64_t count, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HOL_B200_H */
