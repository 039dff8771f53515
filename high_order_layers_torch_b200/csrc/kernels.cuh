// Launcher declarations shared by the translation units (one .cu per kernel family so the
// template instantiations compile in parallel).
#pragma once
#include <atomic>

#include "common.cuh"

int launch_segment_index(const float* x, int64_t* seg, int64_t count, const PwShape& s, cudaStream_t st);
int launch_prep_fc(const float* x, const PwShape& s, const PwWorkspace& ws, cudaStream_t st);
// convolution geometry (rectangular; 1-d convolutions are H = Kh = 1); Ho / Wo derived by the ABI layer
struct HolConvGeom {
    int C, H, W, Kh, Kw, sh, sw, ph, pw, dh, dw, Ho, Wo;
};
int launch_prep_conv(const float* x, const PwShape& s, const PwWorkspace& ws, int64_t Bimg, const HolConvGeom& g,
                     cudaStream_t st);
int launch_col2im(const float* dx_t, float* dx, int64_t Bimg, int64_t Bp, const HolConvGeom& g, cudaStream_t st);
int launch_pack(const float* w, const PwShape& s, const PwWorkspace& ws, cudaStream_t st);
int launch_bucket_sort(const PwShape& s, const PwWorkspace& ws, cudaStream_t st);

// contraction kernels
int launch_forward(const PwShape& s, const PwWorkspace& ws, float* y, cudaStream_t st);
// dx_transposed: write dx_t[I][Bp] (conv path, consumed by col2im) instead of dx[B][I]
int launch_backward_dx(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dx, int dx_transposed,
                       cudaStream_t st);
int launch_backward_dw(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dw, cudaStream_t st);
int launch_dw_reduce(const float* part, float* dw, int64_t count, int ksplit, cudaStream_t st);

// dense-register dW for narrow layers (dwd.cu)
bool dwd_plan(int n, int S, int O, int disc, int64_t B, int I, DwdPlan& d);
int launch_backward_dw_dense(const PwShape& s, const PwWorkspace& ws, const float* gy, float* dw, cudaStream_t st);

// tensor-core passes (tc_gemm.cu)
bool tc_plan(const PwShape& s, TcPlan& t);
// a per-sample normalisation folded into a layer's passes (kind < 0: none)
struct PwNormArgs {
    int kind;
    float eps;
    float* y_norm;  // [B][O]
    float* stats;   // [B][4]
};
int launch_forward_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, float* y,
                      const PwNormArgs& norm, cudaStream_t st);
// max |gy| -> TcScalars::max_gy; once per backward call before the dW / dX passes below
int launch_tc_gy_absmax(const PwShape& s, const PwWorkspace& ws, const float* gy, cudaStream_t st);
// phi_valid: ws.tc_phi still holds the forward's Phi tiles (shared-Phi path)
int launch_backward_dw_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* gy, float* dw,
                          bool phi_valid, cudaStream_t st);
// wstats_valid: max |w| and the link means in the workspace are the forward's
int launch_backward_dx_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, const float* gy,
                          float* dx, int transposed, bool wstats_valid, cudaStream_t st);

// per-sample normalisations (norm.cu); absmax != nullptr: the backward also takes max |dx| (a layer's gy statistics)
int launch_row_norm_forward(int kind, const float* x, float* y, float* stats, int64_t B, int W, float eps, cudaStream_t st);
int launch_row_norm_backward(int kind, const float* gy, const float* x, const float* stats, float* dx, int64_t B, int W,
                             unsigned* absmax, cudaStream_t st);

// per-GEMM event pairs collected while hol_pw_profile_f32 runs (PwWorkspace::prof)
struct TcProf {
    int count;
    cudaEvent_t ev[48][2];
    int kind[48];
};
void tc_prof_end(TcProf* prof, float* ms3);
int tc_launch_count(const PwShape& s, const TcPlan& t, int kind, bool prepared);

// diagnostic plan-override mask (hol_set_plan_overrides, HOL_OVR_*)
uint32_t hol_plan_overrides();

// usable dynamic shared memory per CTA on sm_100 (227 KB) minus a safety margin
#define HOL_SMEM_BUDGET (220 * 1024)
#define HOL_SMEM_MAX (227 * 1024)  // opt-in maximum (cudaFuncAttributeMaxDynamicSharedMemorySize), set once per kernel

// Opt in to the full 227 KB of dynamic shared memory ONCE per kernel and device (not per launch).
template <typename K>
static inline cudaError_t hol_ensure_max_smem(K kern, std::atomic<uint64_t>& done) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HOL_SMEM_MAX);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
