// Launcher declarations shared by the translation units (one .cu per kernel family so the
// template instantiations compile in parallel).
#pragma once
#include "common.cuh"

int launch_segment_index(const float* x, int64_t* seg, int64_t count, const PwShape& s, cudaStream_t st);
int launch_prep_fc(const float* x, const PwShape& s, const PwWorkspace& ws, cudaStream_t st);
int launch_prep_conv(const float* x, const PwShape& s, const PwWorkspace& ws, int64_t Bimg, int C, int H, int W, int K,
                     int stride, int padding, int dilation, int Ho, int Wo, cudaStream_t st);
int launch_col2im(const float* dx_t, float* dx, int64_t Bimg, int64_t Bp, int C, int H, int W, int K, int stride,
                  int padding, int dilation, int Ho, int Wo, cudaStream_t st);
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

// tensor-core forward (tc.cu)
bool tc_plan(const PwShape& s, TcPlan& t);
int launch_forward_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, float* y, cudaStream_t st);
// phi_valid: ws.tc_phi still holds the forward's Phi tiles (shared-Phi path)
int launch_backward_dw_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* gy, float* dw,
                          bool phi_valid, cudaStream_t st);
int launch_backward_dx_tc(const PwShape& s, const PwWorkspace& ws, const TcPlan& t, const float* w, const float* gy,
                          float* dx, int transposed, cudaStream_t st);

void tc_prof_begin();
void tc_prof_end(float* ms3);
int tc_launch_count(const PwShape& s, const TcPlan& t, int kind);

// usable dynamic shared memory per CTA on sm_100 (227 KB) minus a safety margin
#define HOL_SMEM_BUDGET (220 * 1024)
