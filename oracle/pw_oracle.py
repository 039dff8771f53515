"""CPU oracle for the piecewise Lagrange-polynomial hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
It may be imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- always as the
checker or the timed baseline, never as a compute path of the product.

It is an independent numpy restatement of the algorithm of
jloveric/high-order-layers-torch v2.6.0 (paths relative to the reference
checkout):

* segment index ............ high_order_layers_torch/PolynomialLayers.py:299-303 (= :259-275, :643-647)
* local coordinate / _eta .. PolynomialLayers.py:329-344 (disc :671-674, :699-701)
* weight row selection ..... PolynomialLayers.py:307-326 (disc :650-669)
* Chebyshev-Lobatto nodes .. LagrangePolynomial.py:9-22
* Lagrange basis ........... LagrangePolynomial.py:190-227
* contraction .............. Basis.py:492-512  (einsum "ijk,jkli->jl")
* make_periodic ............ utils.py:74-79
* conv expansion ........... Basis.py:38-135, FunctionalConvolution.py:107-139, :392-447
* max-abs normalisation .... utils.py:8-13, layers.py:70-81

Parity pin: the reference stores no golden vectors, so the oracle is pinned
against outputs of the reference itself, generated in the build container by
``oracle/gen_golden.py`` (imports /root/reference) and committed under
``tests/golden/``; ``tests/test_oracle.py`` checks every function here against
those fixtures.

Precision conventions
---------------------
``*_f32`` helpers reproduce the reference's fp32 operation ORDER (each numpy
float32 op is individually rounded, like the ATen elementwise kernels), which
is what makes the segment index bit-exact.  The contraction itself is offered
in fp64 (``dtype=np.float64``, the error yardstick) or fp32.

One known, documented deviation: the node table.  The reference evaluates
``-cos(pi*k/(n-1))`` with torch's fp32 ``cos`` (Sleef), numpy's fp32 ``cos``
differs from it by at most 1 ulp for some n >= 9.  Every function therefore
accepts an explicit ``nodes`` array; the golden fixtures carry the reference's
own tables and the product builds its table with the identical torch
expression on the host.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# node table and basis  (LagrangePolynomial.py:9-22, 190-227)
# --------------------------------------------------------------------------
def lobatto_nodes_f32(n: int, length: float = 2.0) -> np.ndarray:
    """X = (length/2) * (-cos(pi*arange(n)/(n-1))) in fp32; n == 1 -> [0]."""
    if n == 1:
        return np.zeros(1, dtype=F32)
    arg = (F32(np.pi) * np.arange(n).astype(F32)) / F32(n - 1)
    return (F32(length / 2.0) * (-np.cos(arg))).astype(F32)


def lagrange_basis(t: np.ndarray, nodes: np.ndarray, dtype=np.float64) -> np.ndarray:
    """phi[j, ...] = prod_{m != j} (t - X_m) / (X_j - X_m), m ascending.

    n == 1 is the degenerate constant basis (LagrangeBasis1): phi_0 == 1.
    The denominators are kept as divisors, exactly like the reference.
    """
    n = int(nodes.shape[0])
    t = np.asarray(t, dtype=dtype)
    X = np.asarray(nodes, dtype=F32)
    out = np.ones((n,) + t.shape, dtype=dtype)
    if n == 1:
        return out
    for j in range(n):
        for m in range(n):
            if m == j:
                continue
            # the reference's denominator table is fp32: X[j]-X[m] rounded to fp32
            denom = dtype(F32(X[j] - X[m]))
            out[j] = out[j] * ((t - dtype(X[m])) / denom)
    return out


def lagrange_basis_derivative(t: np.ndarray, nodes: np.ndarray, dtype=np.float64) -> np.ndarray:
    """dphi_j/dt = sum_{k != j} 1/(X_j-X_k) prod_{m != j,k} (t-X_m)/(X_j-X_m).

    Product-rule form: exact at the nodes (what autograd of prod/where gives).
    """
    n = int(nodes.shape[0])
    t = np.asarray(t, dtype=dtype)
    X = np.asarray(nodes, dtype=F32)
    out = np.zeros((n,) + t.shape, dtype=dtype)
    if n == 1:
        return out
    for j in range(n):
        for k in range(n):
            if k == j:
                continue
            term = np.full(t.shape, 1.0, dtype=dtype) / dtype(F32(X[j] - X[k]))
            for m in range(n):
                if m == j or m == k:
                    continue
                term = term * ((t - dtype(X[m])) / dtype(F32(X[j] - X[m])))
            out[j] += term
    return out


# --------------------------------------------------------------------------
# elementwise fp32 pre-ops
# --------------------------------------------------------------------------
def make_periodic_f32(x: np.ndarray, periodicity: float):
    """utils.py:74-79.  Returns (folded x, d folded / d x  in {+1,-1})."""
    x = np.asarray(x, dtype=F32)
    p = F32(periodicity)
    two_p = F32(2 * periodicity)
    xp = x + F32(0.5 * periodicity)
    # torch.remainder: fmod, then shift by the divisor when signs differ
    r = np.fmod(xp, two_p)
    fix = (r != 0) & ((r < 0) != (two_p < 0))
    r = np.where(fix, r + two_p, r).astype(F32)
    mirrored = r > p
    r = np.where(mirrored, two_p - r, r).astype(F32)
    out = (r - F32(0.5 * periodicity)).astype(F32)
    sign = np.where(mirrored, F32(-1.0), F32(1.0)).astype(F32)
    return out, sign


def segment_index(x: np.ndarray, segments: int, length: float = 2.0,
                  periodicity: float | None = None) -> np.ndarray:
    """clamp(trunc(((x+half)/length)*S), 0, S-1) as int64; three rounded fp32 ops."""
    x = np.asarray(x, dtype=F32)
    if periodicity is not None:
        x, _ = make_periodic_f32(x, periodicity)
    half = F32(0.5 * length)
    v = ((x + half) / F32(length)) * F32(segments)
    with np.errstate(invalid="ignore"):
        idx = np.where(np.isnan(v), 0.0, np.trunc(v))
        idx = np.clip(idx, -1.0, float(segments)).astype(np.int64)  # clip first: no int overflow
    return np.clip(idx, 0, segments - 1)


def eta_f32(index: np.ndarray, segments: int) -> np.ndarray:
    """_eta: index/float(S)*2 - 1 (int64 -> fp32 promotion, then fp32 ops)."""
    return ((np.asarray(index).astype(F32) / F32(segments)) * F32(2.0) - F32(1.0)).astype(F32)


def local_coordinate_f32(x: np.ndarray, seg: np.ndarray, segments: int, length: float = 2.0):
    """x_in = length*((x - eta(s)) / (eta(s+1)-eta(s))) - half; returns (x_in, d x_in / d x)."""
    x = np.asarray(x, dtype=F32)
    e0 = eta_f32(seg, segments)
    e1 = eta_f32(seg + 1, segments)
    de = (e1 - e0).astype(F32)
    xin = (F32(length) * ((x - e0) / de) - F32(0.5 * length)).astype(F32)
    return xin, (F32(length) / de).astype(F32)


def weight_stride(n: int, discontinuous: bool) -> int:
    return n if discontinuous else n - 1


def weights_per_link(n: int, segments: int, discontinuous: bool) -> int:
    return n * segments if discontinuous else (n - 1) * segments + 1


# --------------------------------------------------------------------------
# fully connected layer: forward and backward
# --------------------------------------------------------------------------
def _prepare(x, n, segments, length, discontinuous, periodicity, nodes):
    x = np.asarray(x, dtype=F32)
    sign = np.ones_like(x)
    if periodicity is not None:
        x, sign = make_periodic_f32(x, periodicity)
    seg = segment_index(x, segments, length, None)
    xin, dxin = local_coordinate_f32(x, seg, segments, length)
    if nodes is None:
        nodes = lobatto_nodes_f32(n)
    rows = seg * weight_stride(n, discontinuous)
    return x, sign, seg, xin, dxin, np.asarray(nodes, dtype=F32), rows


def pw_forward(x, w, n: int, segments: int, length: float = 2.0, discontinuous: bool = False,
               periodicity: float | None = None, nodes=None, dtype=np.float64) -> np.ndarray:
    """y[b,o] = sum_i sum_j phi_j(x_in(b,i)) * w[o, i, s(b,i)*stride + j].

    x: [B, I] fp32, w: [O, I, nb].  Loops over inputs so it never builds the
    reference's [B, I, O, n] gathered tensor.
    """
    _, _, _, xin, _, nodes, rows = _prepare(x, n, segments, length, discontinuous, periodicity, nodes)
    w = np.asarray(w)
    B, I = xin.shape
    O = w.shape[0]
    phi = lagrange_basis(xin, nodes, dtype=dtype)            # [n, B, I]
    y = np.zeros((B, O), dtype=dtype)
    jj = np.arange(n)
    for i in range(I):
        wi = w[:, i, :].astype(dtype)                        # [O, nb]
        g = wi[:, rows[:, i, None] + jj[None, :]]            # [O, B, n]
        y += np.einsum("jb,obj->bo", phi[:, :, i], g)
    return y


def pw_forward_reference_style(x, w, n: int, segments: int, length: float = 2.0,
                               discontinuous: bool = False, periodicity: float | None = None,
                               nodes=None) -> np.ndarray:
    """fp32 forward that materialises the gathered [B, I, O, n] weights like the reference
    (PolynomialLayers.py:307-335).  Small shapes only; used as the bounded CPU-port baseline."""
    _, _, _, xin, _, nodes, rows = _prepare(x, n, segments, length, discontinuous, periodicity, nodes)
    w = np.asarray(w, dtype=F32)
    B, I = xin.shape
    cols = rows[:, :, None] + np.arange(n)[None, None, :]    # [B, I, n]
    ii = np.arange(I)[None, :, None]
    gathered = np.transpose(w[:, ii, cols], (1, 2, 0, 3))    # [B, I, O, n]
    phi = lagrange_basis(xin, nodes, dtype=F32)              # [n, B, I]
    return np.einsum("ijk,jkli->jl", phi, gathered).astype(F32)


def pw_backward(gy, x, w, n: int, segments: int, length: float = 2.0, discontinuous: bool = False,
                periodicity: float | None = None, nodes=None, dtype=np.float64):
    """Explicit gradients of pw_forward (what autograd derives for the reference graph).

    dW[o,i,s*stride+j] += gy[b,o] * phi_j(x_in(b,i))                        (dense, zeros elsewhere)
    dX[b,i] = sigma * (length/(eta(s+1)-eta(s))) * sum_o gy[b,o] sum_j phi'_j(x_in) w[o,i,s*stride+j]
    (.long() cuts the graph, so the segment index carries no gradient.)
    """
    _, sign, _, xin, dxin, nodes, rows = _prepare(x, n, segments, length, discontinuous, periodicity, nodes)
    w = np.asarray(w)
    gy = np.asarray(gy).astype(dtype)
    B, I = xin.shape
    O = w.shape[0]
    phi = lagrange_basis(xin, nodes, dtype=dtype)
    dphi = lagrange_basis_derivative(xin, nodes, dtype=dtype)
    dx = np.zeros((B, I), dtype=dtype)
    dw = np.zeros(w.shape, dtype=dtype)
    jj = np.arange(n)
    for i in range(I):
        cols = rows[:, i, None] + jj[None, :]                # [B, n]
        wi = w[:, i, :].astype(dtype)
        g = wi[:, cols]                                      # [O, B, n]
        t = np.einsum("bo,obj->bj", gy, g)                   # [B, n]
        dx[:, i] = np.einsum("jb,bj->b", dphi[:, :, i], t)
        contrib = np.einsum("bo,jb->obj", gy, phi[:, :, i])  # [O, B, n]
        for o in range(O):
            np.add.at(dw[o, i], cols, contrib[o])
    dx = dx * dxin.astype(dtype) * sign.astype(dtype)
    return dx, dw


# --------------------------------------------------------------------------
# normalisation between layers (utils.py:8-13) and an MLP built from the above
# --------------------------------------------------------------------------
def max_abs_normalization(x, eps: float = 1e-6, dtype=np.float64):
    x = np.asarray(x).astype(dtype)
    m = np.max(np.abs(x), axis=1, keepdims=True)
    return x / (m + dtype(eps))


def max_abs_normalization_backward(g, x, eps: float = 1e-6, dtype=np.float64):
    """Gradient of x/(max|x|+eps); the max routes its gradient to the first arg-max."""
    x = np.asarray(x).astype(dtype)
    g = np.asarray(g).astype(dtype)
    a = np.abs(x)
    k = np.argmax(a, axis=1)
    d = a[np.arange(x.shape[0]), k][:, None] + dtype(eps)
    dx = g / d
    coef = -np.sum(g * x, axis=1) / (d[:, 0] ** 2)
    dx[np.arange(x.shape[0]), k] += coef * np.sign(x[np.arange(x.shape[0]), k])
    return dx


def max_center_normalization(x, eps: float = 1e-6, dtype=np.float64):
    """utils.py:20-28: midrange = (max+min)/2, mag = max - midrange, (x - midrange)/(mag + eps)."""
    x = np.asarray(x).astype(dtype)
    mx, mn = np.max(x, axis=1, keepdims=True), np.min(x, axis=1, keepdims=True)
    mid = dtype(0.5) * (mx + mn)
    return (x - mid) / ((mx - mid) + dtype(eps))


def max_center_normalization_backward(g, x, eps: float = 1e-6, dtype=np.float64):
    """Gradient of the above; max / min route their gradients to the first arg-max / arg-min."""
    x = np.asarray(x).astype(dtype)
    g = np.asarray(g).astype(dtype)
    rows = np.arange(x.shape[0])
    ia, ib = np.argmax(x, axis=1), np.argmin(x, axis=1)
    mx, mn = x[rows, ia][:, None], x[rows, ib][:, None]
    mid = dtype(0.5) * (mx + mn)
    d = (mx - mid) + dtype(eps)
    dx = g / d
    dmid = -np.sum(g, axis=1) / d[:, 0]
    dd = -np.sum(g * (x - mid), axis=1) / d[:, 0] ** 2
    np.add.at(dx, (rows, ia), dtype(0.5) * (dmid + dd))
    np.add.at(dx, (rows, ib), dtype(0.5) * (dmid - dd))
    return dx


def l2_normalization(x, eps: float = 1e-6, dtype=np.float64):
    """utils.py:57-58"""
    x = np.asarray(x).astype(dtype)
    return x / (np.sqrt(np.sum(x * x, axis=1, keepdims=True)) + dtype(eps))


def l2_normalization_backward(g, x, eps: float = 1e-6, dtype=np.float64):
    x = np.asarray(x).astype(dtype)
    g = np.asarray(g).astype(dtype)
    nrm = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
    d = nrm + dtype(eps)
    k = np.where(nrm > 0, -np.sum(g * x, axis=1, keepdims=True) / (d * d * np.where(nrm > 0, nrm, 1)), 0)
    return g / d + k * x


ROW_NORMS = {
    "max_abs": (max_abs_normalization, max_abs_normalization_backward),
    "max_center": (max_center_normalization, max_center_normalization_backward),
    "l2": (l2_normalization, l2_normalization_backward),
}


def mlp_forward_backward(x, weights, gy, n, segments, discontinuous=False, periodicity=None,
                         normalize=True, nodes=None, dtype=np.float64):
    """HighOrderMLP (networks.py:222-288) = layer, [max_abs, layer]*; returns (y, dx, [dW...])."""
    acts = [np.asarray(x, dtype=F32)]
    pre = []
    h = acts[0]
    for li, w in enumerate(weights):
        if li > 0 and normalize:
            pre.append(h)
            h = max_abs_normalization(h, dtype=dtype).astype(F32)
            acts.append(h)
        elif li > 0:
            pre.append(None)
            acts.append(h)
        h = pw_forward(h, w, n, segments, 2.0, discontinuous, periodicity, nodes, dtype).astype(F32)
    y = h
    g = np.asarray(gy).astype(dtype)
    dws = [None] * len(weights)
    for li in range(len(weights) - 1, -1, -1):
        dx, dws[li] = pw_backward(g, acts[li], weights[li], n, segments, 2.0, discontinuous,
                                  periodicity, nodes, dtype)
        g = dx
        if li > 0 and normalize:
            g = max_abs_normalization_backward(g, pre[li - 1], dtype=dtype)
    return y, g, dws


# --------------------------------------------------------------------------
# convolution: expansion (Basis.py:67-127 + FunctionalConvolution.py:123-139) and conv2d
# --------------------------------------------------------------------------
def expand2d(x, n: int, segments: int, length: float = 2.0, discontinuous: bool = False,
             nodes=None, dtype=np.float64) -> np.ndarray:
    """[B, C, H, W] -> [B, C*V, H, W] with channel index c*V + v, V = weights_per_link.

    Entry v = s*stride + j holds phi_j(x_in); every other entry is zero.  For conv the
    reference builds its basis with the layer's ``length`` (LagrangePolynomial.py:242-246)."""
    x = np.asarray(x, dtype=F32)
    B, C, H, W = x.shape
    V = weights_per_link(n, segments, discontinuous)
    seg = segment_index(x, segments, length, None)
    xin, _ = local_coordinate_f32(x, seg, segments, length)
    if nodes is None:
        nodes = lobatto_nodes_f32(n, length)
    phi = lagrange_basis(xin, np.asarray(nodes, dtype=F32), dtype=dtype)     # [n, B, C, H, W]
    out = np.zeros((B, C, V, H, W), dtype=dtype)
    base = seg * weight_stride(n, discontinuous)
    bb, cc, hh, ww = np.meshgrid(np.arange(B), np.arange(C), np.arange(H), np.arange(W), indexing="ij")
    for j in range(n):
        # plain assignment like the reference's index_put (later j overwrite nothing: distinct v)
        out[bb, cc, base + j, hh, ww] = phi[j]
    return out.reshape(B, C * V, H, W)


def _hw(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) else (int(v), int(v))


def conv2d_nchw(x, weight, stride=1, padding=0, dilation=1, dtype=np.float64) -> np.ndarray:
    """Plain cross-correlation (what nn.Conv2d computes), zero padding, no bias, groups=1.
    stride / padding / dilation: ints or (h, w) pairs."""
    x = np.asarray(x).astype(dtype)
    weight = np.asarray(weight).astype(dtype)
    B, C, H, W = x.shape
    O, C2, KH, KW = weight.shape
    assert C == C2
    (sh, sw), (ph, pw), (dh, dw) = _hw(stride), _hw(padding), _hw(dilation)
    if ph or pw:
        x = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)))
    Ho = (H + 2 * ph - dh * (KH - 1) - 1) // sh + 1
    Wo = (W + 2 * pw - dw * (KW - 1) - 1) // sw + 1
    y = np.zeros((B, O, Ho, Wo), dtype=dtype)
    for kh in range(KH):
        for kw in range(KW):
            patch = x[:, :, kh * dh: kh * dh + sh * (Ho - 1) + 1: sh, kw * dw: kw * dw + sw * (Wo - 1) + 1: sw]
            y += np.einsum("bchw,oc->bohw", patch, weight[:, :, kh, kw])
    return y


def conv_transpose2d_nchw(x, weight, stride=1, padding=0, output_padding=0, dilation=1, dtype=np.float64) -> np.ndarray:
    """What nn.ConvTranspose2d computes (no bias, groups=1): weight [C, O, K, K],
    y[b, o, h*s - p + kh*d, w*s - p + kw*d] += x[b, c, h, w] * weight[c, o, kh, kw]."""
    x = np.asarray(x).astype(dtype)
    weight = np.asarray(weight).astype(dtype)
    B, C, H, W = x.shape
    C2, O, KH, KW = weight.shape
    assert C == C2
    Ho = (H - 1) * stride - 2 * padding + dilation * (KH - 1) + output_padding + 1
    Wo = (W - 1) * stride - 2 * padding + dilation * (KW - 1) + output_padding + 1
    full = np.zeros((B, O, Ho + 2 * padding, Wo + 2 * padding), dtype=dtype)
    for kh in range(KH):
        for kw in range(KW):
            contrib = np.einsum("bchw,co->bohw", x, weight[:, :, kh, kw])
            full[:, :, kh * dilation: kh * dilation + stride * (H - 1) + 1: stride,
                 kw * dilation: kw * dilation + stride * (W - 1) + 1: stride] += contrib
    return full[:, :, padding: padding + Ho, padding: padding + Wo]


def pw_conv_transpose2d_forward(x, weight, n: int, segments: int, length: float = 2.0, periodicity: float | None = None,
                                stride=1, padding=0, output_padding=0, dilation=1, rescale: float = 1.0, nodes=None,
                                dtype=np.float64) -> np.ndarray:
    """PiecewisePolynomialConvolutionTranspose.forward (FunctionalConvolutionTranspose.py:127-134)."""
    x = np.asarray(x, dtype=F32)
    if periodicity is not None:
        x, _ = make_periodic_f32(x, periodicity)
    ex = expand2d(x, n, segments, length, False, nodes, dtype)
    return conv_transpose2d_nchw(ex, weight, stride, padding, output_padding, dilation, dtype) * dtype(rescale)


def pw_conv1d_forward(x, weight, n: int, segments: int, length: float = 2.0, discontinuous: bool = False,
                      periodicity: float | None = None, stride=1, padding=0, dilation=1, rescale: float = 1.0,
                      nodes=None, dtype=np.float64) -> np.ndarray:
    """PiecewisePolynomialConvolution1d / Discontinuous1d (FunctionalConvolution.py:493-519, :622-650):
    Expansion1d + nn.Conv1d == the 2-d path with H = Kh = 1.  x [B, C, L], weight [O, C*V, K]."""
    x = np.asarray(x, dtype=F32)
    y = pw_conv2d_forward(x[:, :, None, :], np.asarray(weight)[:, :, None, :], n, segments, length, discontinuous,
                          periodicity, (1, stride), (0, padding), (1, dilation), rescale, nodes, dtype)
    return y[:, :, 0, :]


def pw_conv2d_forward(x, weight, n: int, segments: int, length: float = 2.0,
                      discontinuous: bool = False, periodicity: float | None = None,
                      stride=1, padding=0, dilation=1, rescale: float = 1.0, nodes=None,
                      dtype=np.float64) -> np.ndarray:
    """PiecewisePolynomialConvolution.forward (FunctionalConvolution.py:441-447)."""
    x = np.asarray(x, dtype=F32)
    if periodicity is not None:
        x, _ = make_periodic_f32(x, periodicity)
    ex = expand2d(x, n, segments, length, discontinuous, nodes, dtype)
    return conv2d_nchw(ex, weight, stride, padding, dilation, dtype) * dtype(rescale)


def unfold_nchw(x, k: int, stride=1, padding=0, dilation=1) -> np.ndarray:
    """im2col: [B,C,H,W] -> [B*Ho*Wo, C*k*k], column index c*k*k + kh*k + kw (F.unfold order)."""
    x = np.asarray(x)
    B, C, H, W = x.shape
    if padding:
        x = np.pad(x, ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    Ho = (H + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    cols = np.zeros((B, Ho, Wo, C, k, k), dtype=x.dtype)
    for kh in range(k):
        for kw in range(k):
            cols[:, :, :, :, kh, kw] = np.transpose(
                x[:, :, kh * dilation: kh * dilation + stride * (Ho - 1) + 1: stride,
                  kw * dilation: kw * dilation + stride * (Wo - 1) + 1: stride], (0, 2, 3, 1))
    return cols.reshape(B * Ho * Wo, C * k * k), (B, Ho, Wo)


def conv_weight_as_fc(weight, in_channels: int, n: int, segments: int, discontinuous: bool = False):
    """conv.weight [O, C*V, K, K] -> fc w [O, C*K*K, V]  (the unfold identity)."""
    weight = np.asarray(weight)
    O, CV, K, _ = weight.shape
    V = weights_per_link(n, segments, discontinuous)
    assert CV == in_channels * V
    return weight.reshape(O, in_channels, V, K, K).transpose(0, 1, 3, 4, 2).reshape(O, in_channels * K * K, V)
