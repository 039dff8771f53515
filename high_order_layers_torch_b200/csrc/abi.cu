// C ABI (include/hol_b200.h): argument validation, planning, workspace carving, launches.
// Never allocates, never synchronises; everything goes to the caller's stream.
#include <math.h>
#include <string.h>

#include <atomic>

#include "kernels.cuh"

// The one piece of process-wide state: a diagnostic mask that forces a kernel family (tests, A/B
// runs).  Read once per call when the plan is made; 0 (the default) = the planner's own choice.
static std::atomic<uint32_t> g_plan_overrides{0};
uint32_t hol_plan_overrides() { return g_plan_overrides.load(std::memory_order_relaxed); }

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Fill the plan for one layer call.  Returns HOL_OK / HOL_EINVAL / HOL_EUNSUPPORTED.
int make_shape(PwShape& s, int64_t B, int I, int O, int n, int S, int disc, float length, float periodicity,
               const float* nodes_host) {
    if (B < 0 || I <= 0 || O <= 0 || n <= 0 || S <= 0) return HOL_EINVAL;
    if (n > HOL_MAX_N || S > HOL_MAX_SEGMENTS) return HOL_EUNSUPPORTED;
    memset(&s, 0, sizeof(s));
    s.B = B;
    s.Bp = B <= 256 ? round_up64(B > 0 ? B : 1, 32) : round_up64(B, 512);
    s.I = I;
    s.O = O;
    s.n = n;
    s.S = S;
    s.disc = disc ? 1 : 0;
    s.stride = disc ? n : n - 1;
    s.nb = disc ? n * S : (n - 1) * S + 1;
    const bool single = !disc && n == 1;  // one weight per link whatever the segment
    s.RS = single ? 1 : (disc ? S : S + 1);
    s.R = single ? 1 : (disc ? n * S : (n - 1) * (S + 1));
    s.Seff = single ? 1 : S;
    s.nchunks = (int)((B + HOL_SORT_CHUNK - 1) / HOL_SORT_CHUNK);
    if (s.nchunks < 1) s.nchunks = 1;
    // dW work decomposition: lanes per sample (4 outputs each) and split over sample chunks so
    // that narrow/few-bucket layers still fill the machine (>= ~4 warps per SM scheduler)
    s.dw_lo = 32;
    while (s.dw_lo > 1 && 4 * (s.dw_lo / 2) >= O) s.dw_lo >>= 1;
    {
        const int64_t otw = 4 * s.dw_lo;
        const int64_t tasks = ((O + otw - 1) / otw) * (int64_t)I * s.Seff;
        int64_t ks = (148 * 32 + tasks - 1) / tasks;
        if (ks > s.nchunks) ks = s.nchunks;
        if (ks > 64) ks = 64;
        if (ks < 1) ks = 1;
        s.ksplit = (int)ks;
    }
    // output tile: wide layers use 64 accumulators per thread; the packed rows of ONE input
    // (R x (OT+4) floats) must fit the shared-memory budget at least once
    int OT = O >= 48 ? 64 : (O > 16 ? 32 : (O > 8 ? 16 : 8));
    while (OT > 8 && (size_t)s.R * (OT + 4) * 4 + 128 > HOL_SMEM_BUDGET) OT >>= 1;
    if ((size_t)s.R * (OT + 4) * 4 + 128 > HOL_SMEM_BUDGET) return HOL_EUNSUPPORTED;
    s.OT = OT;
    // up to 8 segments the +4 pitch keeps the segment rows in distinct banks; beyond that the pitch
    // is the tile itself (a multiple of 128 bytes) and the lanes rotate their column order (fwd.cu)
    s.rot = (S > 8 && OT >= 32) ? 1 : 0;
    s.OTP = s.rot ? OT : OT + 4;
    s.nOT = (O + OT - 1) / OT;
    s.srow = single ? 0 : s.OTP;
    for (int j = 0; j < HOL_MAX_N; ++j) s.jrow[j] = 0;
    for (int j = 0; j < n; ++j) {
        int row0;
        if (single)
            row0 = 0;
        else if (disc)
            row0 = j * S;
        else
            row0 = (j < n - 1) ? j * (S + 1) : 1;  // last node of segment s == first node of s+1
        s.jrow[j] = row0 * s.OTP;
    }
    s.length = length;
    s.half = (float)(0.5 * (double)length);
    s.periodic = !isnan(periodicity);
    s.periodicity = s.periodic ? periodicity : 0.0f;
    s.half_per = (float)(0.5 * (double)s.periodicity);
    s.two_per = (float)(2.0 * (double)s.periodicity);
    if (nodes_host) {
        for (int j = 0; j < n; ++j) s.tab.X[j] = nodes_host[j];
        for (int j = 0; j < n; ++j) {
            double prod = 1.0;
            for (int m = 0; m < n; ++m)
                if (m != j) prod *= (double)(float)(nodes_host[j] - nodes_host[m]);  // fp32 denominators like the reference
            s.tab.C[j] = (float)(1.0 / prod);
        }
    }
    tc_plan(s, s.tc);
    dwd_plan(n, S, O, s.disc, B, I, s.dwd);
    return HOL_OK;
}

size_t carve(const PwShape& s, void* base, PwWorkspace& ws, int64_t conv_pixels) {
    const bool conv = conv_pixels >= 0;
    size_t off = 0;
    char* p = (char*)base;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return p ? p + o : (char*)nullptr;
    };
    ws.xl_t = (float*)take((size_t)s.I * s.Bp * sizeof(float));
    ws.seg_t = (uint16_t*)take((size_t)s.I * s.Bp * sizeof(uint16_t));
    ws.wk = (float*)take((size_t)s.nOT * s.I * s.R * s.OTP * sizeof(float));
    // the bucket lists only exist for the bucketed dW kernel
    const bool bucketed = !s.dwd.enabled;
    ws.order = bucketed ? (uint32_t*)take((size_t)s.I * s.Bp * sizeof(uint32_t)) : nullptr;
    ws.offs = bucketed ? (uint32_t*)take((size_t)s.I * s.nchunks * (s.Seff + 1) * sizeof(uint32_t)) : nullptr;
    int parts = bucketed ? s.ksplit : s.dwd.ksplit;
    if (s.tc.dw) parts = s.tc.dw_ksplit;  // tensor-core dW: batch split over several CTAs per tile
    ws.dw_part = parts > 1 ? (float*)take((size_t)parts * s.O * s.I * s.nb * sizeof(float)) : nullptr;
    ws.y_part = (s.tc.fwd && s.tc.fwd_ksplit > 1) ? (float*)take((size_t)s.tc.fwd_ksplit * s.B * s.O * sizeof(float)) : nullptr;
    ws.dx_rows = conv ? (float*)take((size_t)s.I * s.Bp * sizeof(float)) : nullptr;
    ws.xl_img = conv ? (float*)take((size_t)conv_pixels * sizeof(float)) : nullptr;
    ws.seg_img = conv ? (uint16_t*)take((size_t)conv_pixels * sizeof(uint16_t)) : nullptr;
    ws.tc_phi = (s.tc.enabled && s.tc.share_phi) ? (void*)take(s.tc.phi_bytes) : nullptr;
    ws.tc_a = s.tc.enabled ? (void*)take(s.tc.a_bytes) : nullptr;
    ws.tc_b = s.tc.enabled ? (void*)take(s.tc.b_bytes) : nullptr;
    ws.tc_scalars = s.tc.enabled ? (void*)take(256) : nullptr;
    ws.tc_wmean = (s.tc.enabled && s.tc.dx) ? (float*)take((size_t)s.O * s.I * sizeof(float)) : nullptr;
    ws.gy_t = (s.n > 1 && s.OT >= 32 && !(s.tc.enabled && s.tc.dx)) ? (float*)take((size_t)s.nOT * s.OT * s.Bp * sizeof(float)) : nullptr;
    ws.prof = nullptr;
    ws.total = off;
    return off;
}

int conv_out(int H, int K, int stride, int padding, int dilation) {
    return (H + 2 * padding - dilation * (K - 1) - 1) / stride + 1;
}

#define HOL_TRY(expr)               \
    do {                            \
        int _e = (expr);            \
        if (_e != HOL_OK) return _e; \
    } while (0)

// One memset per forward zeroes the device scalars of the tensor-core passes (maxima, outlier count)
int reset_scalars(const PwShape& s, const PwWorkspace& ws, cudaStream_t st) {
    if (!s.tc.enabled) return HOL_OK;
    return (int)cudaMemsetAsync(ws.tc_scalars, 0, 256, st);
}

// `packed`: the gather-path weight packing (wk) in the workspace is valid
// `prepared`: the workspace still holds the forward's state (Phi tiles on the shared-Phi path, max |w|
// and the link means).
// dW runs first: under data parallelism its all-reduce can then start while dX is still computing.
int backward_common(const PwShape& s, const PwWorkspace& ws, const float* gy, const float* w, float* dx,
                    int dx_transposed, float* dw, bool packed, bool prepared, bool gy_stats_valid, cudaStream_t st) {
    // max|gy| is taken whenever the plan has a tensor-core backward pass at all, not only when THIS call runs
    // one: a backward split into a dW call and a dX call (either order) hands HOL_GY_STATS_VALID to the second.
    const bool tc_any = (s.n > 1 && s.tc.dx) || s.tc.dw;
    if (tc_any && (dx || dw) && !gy_stats_valid) HOL_TRY(launch_tc_gy_absmax(s, ws, gy, st));
    if (dw) {
        if (s.tc.dw) {
            HOL_TRY(launch_backward_dw_tc(s, ws, s.tc, gy, dw, prepared && s.tc.fwd, st));
        } else if (s.dwd.enabled) {
            HOL_TRY(launch_backward_dw_dense(s, ws, gy, dw, st));
        } else {
            HOL_TRY(launch_bucket_sort(s, ws, st));
            HOL_TRY(launch_backward_dw(s, ws, gy, dw, st));
        }
    }
    if (dx) {
        if (s.n == 1) {
            const size_t bytes = dx_transposed ? (size_t)s.I * s.Bp * 4 : (size_t)s.B * s.I * 4;
            cudaError_t e = cudaMemsetAsync(dx, 0, bytes, st);
            if (e != cudaSuccess) return (int)e;
        } else if (s.tc.dx) {
            HOL_TRY(launch_backward_dx_tc(s, ws, s.tc, w, gy, dx, dx_transposed, prepared && s.tc.fwd, st));
        } else {
            if (!packed) HOL_TRY(launch_pack(w, s, ws, st));
            HOL_TRY(launch_backward_dx(s, ws, gy, dx, dx_transposed, st));
        }
    }
    return HOL_OK;
}

}  // namespace

extern "C" {

int hol_abi_version(void) { return HOL_ABI_VERSION; }

uint32_t hol_set_plan_overrides(uint32_t mask) { return g_plan_overrides.exchange(mask, std::memory_order_relaxed); }

const char* hol_error_string(int code) {
    switch (code) {
        case HOL_OK: return "ok";
        case HOL_EINVAL: return "hol: invalid argument";
        case HOL_EUNSUPPORTED: return "hol: shape outside the compiled envelope (n <= 10, segments <= 4096, rows must fit shared memory)";
        case HOL_EWORKSPACE: return "hol: workspace too small (see hol_pw_workspace_bytes)";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "hol: unknown error";
}

size_t hol_pw_workspace_bytes(int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, int32_t discontinuous) {
    PwShape s;
    if (make_shape(s, B, I, O, n, S, discontinuous, 2.0f, NAN, nullptr) != HOL_OK) return 0;
    PwWorkspace ws;
    return carve(s, nullptr, ws, -1);
}

int hol_pw_segment_index_i64(const float* x, int64_t* seg, int64_t count, int32_t S, float length, float periodicity,
                             void* stream) {
    if (count < 0 || S <= 0 || (count > 0 && (!x || !seg))) return HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, 1, 1, 1, 1, S, 1, length, periodicity, nullptr));
    return launch_segment_index(x, seg, count, s, (cudaStream_t)stream);
}

static int forward_fc(const float* x, const float* w, const float* nodes_host, float* y, const PwNormArgs& norm, int64_t B,
                      int32_t I, int32_t O, int32_t n, int32_t S, float length, int32_t discontinuous, float periodicity,
                      void* ws_ptr, size_t ws_bytes, void* stream) {
    if (!x || !w || !nodes_host || !y || !ws_ptr) return B == 0 ? HOL_OK : HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, B, I, O, n, S, discontinuous, length, periodicity, nodes_host));
    if (B == 0) return HOL_OK;
    PwWorkspace ws;
    if (carve(s, ws_ptr, ws, -1) > ws_bytes) return HOL_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    HOL_TRY(reset_scalars(s, ws, st));
    HOL_TRY(launch_prep_fc(x, s, ws, st));
    if (s.tc.fwd) return launch_forward_tc(s, ws, s.tc, w, y, norm, st);
    HOL_TRY(launch_pack(w, s, ws, st));
    HOL_TRY(launch_forward(s, ws, y, st));
    // gather path: the normalisation runs as its own kernel right behind the forward, inside the same call
    if (norm.kind >= 0) return launch_row_norm_forward(norm.kind, y, norm.y_norm, norm.stats, B, O, norm.eps, st);
    return HOL_OK;
}

int hol_pw_forward_f32(const float* x, const float* w, const float* nodes_host, float* y, int64_t B, int32_t I,
                       int32_t O, int32_t n, int32_t S, float length, int32_t discontinuous, float periodicity,
                       void* ws_ptr, size_t ws_bytes, void* stream) {
    const PwNormArgs none{-1, 0.0f, nullptr, nullptr};
    return forward_fc(x, w, nodes_host, y, none, B, I, O, n, S, length, discontinuous, periodicity, ws_ptr, ws_bytes, stream);
}

int hol_pw_forward_norm_f32(const float* x, const float* w, const float* nodes_host, float* y_raw, float* y_norm,
                            float* stats, int32_t norm_kind, float eps, int64_t B, int32_t I, int32_t O, int32_t n,
                            int32_t S, float length, int32_t discontinuous, float periodicity, void* ws_ptr,
                            size_t ws_bytes, void* stream) {
    if (norm_kind < 0 || norm_kind > 2) return HOL_EINVAL;
    if (B > 0 && (!y_norm || !stats)) return HOL_EINVAL;
    const PwNormArgs norm{norm_kind, eps, y_norm, stats};
    return forward_fc(x, w, nodes_host, y_raw, norm, B, I, O, n, S, length, discontinuous, periodicity, ws_ptr, ws_bytes,
                      stream);
}

int hol_pw_norm_backward_f32(int32_t norm_kind, const float* gy_norm, const float* y_raw, const float* stats,
                             float* gy_raw, int64_t B, int32_t I, int32_t O, int32_t n, int32_t S,
                             int32_t discontinuous, void* ws_ptr, size_t ws_bytes, void* stream) {
    if (norm_kind < 0 || norm_kind > 2 || B < 0) return HOL_EINVAL;
    if (B == 0) return HOL_OK;
    if (!gy_norm || !y_raw || !stats || !gy_raw || !ws_ptr) return HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, B, I, O, n, S, discontinuous, 2.0f, NAN, nullptr));
    PwWorkspace ws;
    if (carve(s, ws_ptr, ws, -1) > ws_bytes) return HOL_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    // the same sweep takes max|gy_raw| for the tensor-core backward passes of the layer (HOL_GY_STATS_VALID)
    unsigned* absmax = nullptr;
    if (s.tc.enabled && (s.tc.dx || s.tc.dw)) {
        absmax = reinterpret_cast<unsigned*>(ws.tc_scalars) + 2;  // TcScalars::max_gy
        cudaError_t e = cudaMemsetAsync(absmax, 0, sizeof(unsigned), st);
        if (e != cudaSuccess) return (int)e;
    }
    return launch_row_norm_backward(norm_kind, gy_norm, y_raw, stats, gy_raw, B, O, absmax, st);
}

int hol_pw_backward_f32(const float* gy, const float* x, const float* w, const float* nodes_host, float* dx, float* dw,
                        int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, float length, int32_t discontinuous,
                        float periodicity, void* ws_ptr, size_t ws_bytes, uint32_t flags, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {  // empty batch (its tensors have null data pointers): dX is empty, dW is all zeros
        PwShape s0;
        HOL_TRY(make_shape(s0, 0, I, O, n, S, discontinuous, length, periodicity, nullptr));
        if (dw) {
            cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)O * I * s0.nb * 4, st);
            if (e != cudaSuccess) return (int)e;
        }
        return HOL_OK;
    }
    if (!gy || !x || !w || !nodes_host || !ws_ptr) return HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, B, I, O, n, S, discontinuous, length, periodicity, nodes_host));
    PwWorkspace ws;
    if (carve(s, ws_ptr, ws, -1) > ws_bytes) return HOL_EWORKSPACE;
    if (!(flags & HOL_WS_PREPARED)) {
        HOL_TRY(reset_scalars(s, ws, st));
        HOL_TRY(launch_prep_fc(x, s, ws, st));
    }
    return backward_common(s, ws, gy, w, dx, 0, dw, (flags & HOL_WS_PREPARED) != 0 && !s.tc.fwd,
                           (flags & HOL_WS_PREPARED) != 0, (flags & HOL_GY_STATS_VALID) != 0, st);
}

// ---- convolutions: rectangular geometry; the square 2-d entry points forward to it -----------------
static bool conv_geom(HolConvGeom& g, const int32_t* q) {
    if (!q) return false;
    g.C = q[0], g.H = q[1], g.W = q[2], g.Kh = q[3], g.Kw = q[4], g.sh = q[5], g.sw = q[6], g.ph = q[7], g.pw = q[8];
    g.dh = q[9], g.dw = q[10];
    if (g.C <= 0 || g.H <= 0 || g.W <= 0 || g.Kh <= 0 || g.Kw <= 0 || g.sh <= 0 || g.sw <= 0 || g.dh <= 0 || g.dw <= 0 ||
        g.ph < 0 || g.pw < 0)
        return false;
    g.Ho = conv_out(g.H, g.Kh, g.sh, g.ph, g.dh);
    g.Wo = conv_out(g.W, g.Kw, g.sw, g.pw, g.dw);
    return g.Ho > 0 && g.Wo > 0;
}

size_t hol_pw_conv_workspace_bytes(int64_t Bimg, const int32_t* geom, int32_t O, int32_t n, int32_t S,
                                   int32_t discontinuous) {
    HolConvGeom g;
    if (Bimg < 0 || !conv_geom(g, geom)) return 0;
    PwShape s;
    if (make_shape(s, Bimg * g.Ho * g.Wo, g.C * g.Kh * g.Kw, O, n, S, discontinuous, 2.0f, NAN, nullptr) != HOL_OK) return 0;
    PwWorkspace ws;
    return carve(s, nullptr, ws, Bimg * g.C * g.H * g.W);
}

int hol_pw_conv_forward_f32(const float* x, const float* w_fc, const float* nodes_host, float* y_rows, int64_t Bimg,
                            const int32_t* geom, int32_t O, int32_t n, int32_t S, float length, int32_t discontinuous,
                            float periodicity, void* ws_ptr, size_t ws_bytes, void* stream) {
    HolConvGeom g;
    if (Bimg < 0 || !conv_geom(g, geom)) return HOL_EINVAL;
    if (Bimg == 0) return HOL_OK;
    if (!x || !w_fc || !nodes_host || !y_rows || !ws_ptr) return HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, Bimg * g.Ho * g.Wo, g.C * g.Kh * g.Kw, O, n, S, discontinuous, length, periodicity, nodes_host));
    PwWorkspace ws;
    if (carve(s, ws_ptr, ws, Bimg * g.C * g.H * g.W) > ws_bytes) return HOL_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    HOL_TRY(reset_scalars(s, ws, st));
    HOL_TRY(launch_prep_conv(x, s, ws, Bimg, g, st));
    if (s.tc.fwd) return launch_forward_tc(s, ws, s.tc, w_fc, y_rows, PwNormArgs{-1, 0.0f, nullptr, nullptr}, st);
    HOL_TRY(launch_pack(w_fc, s, ws, st));
    return launch_forward(s, ws, y_rows, st);
}

int hol_pw_conv_backward_f32(const float* gy_rows, const float* x, const float* w_fc, const float* nodes_host, float* dx,
                             float* dw_fc, int64_t Bimg, const int32_t* geom, int32_t O, int32_t n, int32_t S,
                             float length, int32_t discontinuous, float periodicity, void* ws_ptr, size_t ws_bytes,
                             uint32_t flags, void* stream) {
    HolConvGeom g;
    if (Bimg < 0 || !conv_geom(g, geom)) return HOL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (Bimg == 0) {  // empty batch: dX is empty, dW is all zeros
        PwShape s0;
        HOL_TRY(make_shape(s0, 0, g.C * g.Kh * g.Kw, O, n, S, discontinuous, length, periodicity, nullptr));
        if (dw_fc) {
            cudaError_t e = cudaMemsetAsync(dw_fc, 0, (size_t)O * s0.I * s0.nb * 4, st);
            if (e != cudaSuccess) return (int)e;
        }
        return HOL_OK;
    }
    if (!gy_rows || !x || !w_fc || !nodes_host || !ws_ptr) return HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, Bimg * g.Ho * g.Wo, g.C * g.Kh * g.Kw, O, n, S, discontinuous, length, periodicity, nodes_host));
    PwWorkspace ws;
    if (carve(s, ws_ptr, ws, Bimg * g.C * g.H * g.W) > ws_bytes) return HOL_EWORKSPACE;
    if (!(flags & HOL_WS_PREPARED)) {
        HOL_TRY(reset_scalars(s, ws, st));
        HOL_TRY(launch_prep_conv(x, s, ws, Bimg, g, st));
    }
    HOL_TRY(backward_common(s, ws, gy_rows, w_fc, dx ? ws.dx_rows : nullptr, 1, dw_fc,
                            (flags & HOL_WS_PREPARED) != 0 && !s.tc.fwd, (flags & HOL_WS_PREPARED) != 0,
                            (flags & HOL_GY_STATS_VALID) != 0, st));
    if (dx) HOL_TRY(launch_col2im(ws.dx_rows, dx, Bimg, s.Bp, g, st));
    return HOL_OK;
}

#define HOL_SQUARE_GEOM(q) const int32_t q[11] = {C, H, W, K, K, stride, stride, padding, padding, dilation, dilation}

size_t hol_pw_conv2d_workspace_bytes(int64_t Bimg, int32_t C, int32_t H, int32_t W, int32_t O, int32_t K,
                                     int32_t stride, int32_t padding, int32_t dilation, int32_t n, int32_t S,
                                     int32_t discontinuous) {
    HOL_SQUARE_GEOM(q);
    return hol_pw_conv_workspace_bytes(Bimg, q, O, n, S, discontinuous);
}

int hol_pw_conv2d_forward_f32(const float* x, const float* w_fc, const float* nodes_host, float* y_rows, int64_t Bimg,
                              int32_t C, int32_t H, int32_t W, int32_t O, int32_t K, int32_t stride, int32_t padding,
                              int32_t dilation, int32_t n, int32_t S, float length, int32_t discontinuous,
                              float periodicity, void* ws_ptr, size_t ws_bytes, void* stream) {
    HOL_SQUARE_GEOM(q);
    return hol_pw_conv_forward_f32(x, w_fc, nodes_host, y_rows, Bimg, q, O, n, S, length, discontinuous, periodicity, ws_ptr,
                                   ws_bytes, stream);
}

int hol_pw_conv2d_backward_f32(const float* gy_rows, const float* x, const float* w_fc, const float* nodes_host,
                               float* dx, float* dw_fc, int64_t Bimg, int32_t C, int32_t H, int32_t W, int32_t O,
                               int32_t K, int32_t stride, int32_t padding, int32_t dilation, int32_t n, int32_t S,
                               float length, int32_t discontinuous, float periodicity, void* ws_ptr, size_t ws_bytes,
                               uint32_t flags, void* stream) {
    HOL_SQUARE_GEOM(q);
    return hol_pw_conv_backward_f32(gy_rows, x, w_fc, nodes_host, dx, dw_fc, Bimg, q, O, n, S, length, discontinuous,
                                    periodicity, ws_ptr, ws_bytes, flags, stream);
}

int hol_pw_profile_f32(const float* gy, const float* x, const float* w, const float* nodes_host, float* y, float* dx,
                       float* dw, int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, float length,
                       int32_t discontinuous, float periodicity, void* ws_ptr, size_t ws_bytes, float* ms_out_host,
                       void* stream) {
    if (!gy || !x || !w || !nodes_host || !y || !dx || !dw || !ws_ptr || !ms_out_host || B <= 0) return HOL_EINVAL;
    PwShape s;
    HOL_TRY(make_shape(s, B, I, O, n, S, discontinuous, length, periodicity, nodes_host));
    PwWorkspace ws;
    if (carve(s, ws_ptr, ws, -1) > ws_bytes) return HOL_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[7];
    for (int k = 0; k < 7; ++k)
        if (cudaEventCreate(&ev[k]) != cudaSuccess) return (int)cudaGetLastError();
    int rc = HOL_OK;
    auto mark = [&](int k) { cudaEventRecord(ev[k], st); };
    TcProf prof;
    prof.count = 0;
    ws.prof = &prof;  // the GEMM launchers record an event pair per launch
    mark(0);
    rc = reset_scalars(s, ws, st);
    if (rc == HOL_OK) rc = launch_prep_fc(x, s, ws, st);
    mark(1);
    if (rc == HOL_OK && !(s.tc.fwd && (s.tc.dx || n == 1))) rc = launch_pack(w, s, ws, st);
    mark(2);
    if (rc == HOL_OK)
        rc = s.tc.fwd ? launch_forward_tc(s, ws, s.tc, w, y, PwNormArgs{-1, 0.0f, nullptr, nullptr}, st)
                      : launch_forward(s, ws, y, st);
    mark(3);
    if (rc == HOL_OK && ((s.tc.dx && n > 1) || s.tc.dw)) rc = launch_tc_gy_absmax(s, ws, gy, st);
    if (rc == HOL_OK) {
        if (n == 1)
            cudaMemsetAsync(dx, 0, (size_t)B * I * 4, st);
        else
            rc = s.tc.dx ? launch_backward_dx_tc(s, ws, s.tc, w, gy, dx, 0, s.tc.fwd, st)
                         : launch_backward_dx(s, ws, gy, dx, 0, st);
    }
    mark(4);
    if (rc == HOL_OK && !s.tc.dw && !s.dwd.enabled) rc = launch_bucket_sort(s, ws, st);
    mark(5);
    if (rc == HOL_OK)
        rc = s.tc.dw ? launch_backward_dw_tc(s, ws, s.tc, gy, dw, s.tc.fwd, st)
                     : (s.dwd.enabled ? launch_backward_dw_dense(s, ws, gy, dw, st) : launch_backward_dw(s, ws, gy, dw, st));
    mark(6);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == HOL_OK && e != cudaSuccess) rc = (int)e;
    for (int k = 0; k < 6; ++k) {
        float ms = 0.0f;
        if (rc == HOL_OK) cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
        ms_out_host[k] = ms;
    }
    for (int k = 0; k < 7; ++k) cudaEventDestroy(ev[k]);
    float g3[3];
    tc_prof_end(&prof, g3);
    ms_out_host[6] = g3[0];  // tensor-core GEMM inside "forward"
    ms_out_host[7] = g3[2];  // ... inside "dx"
    ms_out_host[8] = g3[1];  // ... inside "dw"
    return rc;
}

int hol_pw_plan_flags(int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, int32_t discontinuous) {
    PwShape s;
    if (make_shape(s, B, I, O, n, S, discontinuous, 2.0f, NAN, nullptr) != HOL_OK) return -1;
    return (s.tc.fwd ? 1 : 0) | (s.tc.dx ? 2 : 0) | (s.tc.dw ? 4 : 0) | (s.dwd.enabled ? 8 : 0) |
           ((s.tc.enabled && s.tc.share_phi) ? 16 : 0) | (s.rot ? 32 : 0) |
           ((s.tc.dw ? (s.tc.dw_ksplit & 255) : 0) << 8) | ((s.tc.fwd ? (s.tc.fwd_ksplit & 255) : 0) << 16);
}

int hol_pw_launch_count(int kind, int64_t B, int32_t I, int32_t O, int32_t n, int32_t S, int32_t discontinuous,
                        int32_t want_dx, int32_t want_dw, uint32_t flags) {
    PwShape s;
    if (make_shape(s, B, I, O, n, S, discontinuous, 2.0f, NAN, nullptr) != HOL_OK) return -1;
    const bool prepared = (flags & HOL_WS_PREPARED) != 0;
    const bool fused_norm = (flags & HOL_COUNT_FUSED_NORM) != 0;
    // prep + (pack, forward [, normalisation kernel]); on the tensor-core path the normalisation rides in the finalize pass
    if (kind == 0) return 1 + (s.tc.fwd ? tc_launch_count(s, s.tc, 0, false) : 2 + (fused_norm ? 1 : 0));
    int c = prepared ? 0 : 1;  // prep
    if (fused_norm) c += 1;    // normalisation backward (takes max|gy| along the way: HOL_GY_STATS_VALID)
    const bool tc_any = (n > 1 && s.tc.dx) || s.tc.dw;
    if (tc_any && (want_dx || want_dw) && !(flags & HOL_GY_STATS_VALID)) c += tc_launch_count(s, s.tc, 3, prepared);  // max |gy|
    if (want_dx && n > 1) c += s.tc.dx ? tc_launch_count(s, s.tc, 1, prepared) : ((prepared && !s.tc.fwd) ? 1 : 2) + (s.OT >= 32 ? 1 : 0);  // [pack], [gy transpose: wide tiles], dX
    if (want_dw)
        c += s.tc.dw ? tc_launch_count(s, s.tc, 2, prepared)
                     : (s.dwd.enabled ? 1 + (s.dwd.ksplit > 1 ? 1 : 0) : 2 + (s.ksplit > 1 ? 1 : 0));
    return c;
}

}  // extern "C"
