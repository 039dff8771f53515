"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]

The reference (jloveric/high-order-layers-torch v2.6.0) is pure Python/torch, so it is
imported as-is (read-only, no bytecode written) and evaluated on CPU in fp32; its own
autograd provides the gradients.  ``networks.py`` imports ``pytorch_lightning`` only for a
base class, so a 3-line in-memory stub stands in for it.  Everything written here is data
(inputs, weights, outputs, gradients, index tensors); no reference source is copied.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True


def _import_reference(ref_root: str):
    import torch

    stub = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):  # base class only
        pass

    stub.LightningModule = LightningModule
    sys.modules.setdefault("pytorch_lightning", stub)
    sys.path.insert(0, ref_root)
    import high_order_layers_torch.layers as layers  # noqa
    import high_order_layers_torch.networks as networks  # noqa
    import high_order_layers_torch.PolynomialLayers as pl  # noqa
    import high_order_layers_torch.LagrangePolynomial as lp  # noqa
    import high_order_layers_torch.FunctionalConvolution as fcv  # noqa
    import high_order_layers_torch.utils as utils  # noqa

    return types.SimpleNamespace(layers=layers, networks=networks, pl=pl, lp=lp, fcv=fcv, utils=utils)


def special_inputs(S: int, length: float, rng: np.random.Generator, count: int) -> np.ndarray:
    """Segment boundaries, their one-ulp neighbours, +-half, values outside the domain, random fill."""
    half = 0.5 * length
    vals = []
    for k in range(S + 1):
        b = np.float32(-half + k * length / S)
        vals += [b, np.nextafter(b, np.float32(10)), np.nextafter(b, np.float32(-10))]
    vals += [np.float32(v) for v in (-half, half, -1.25 * half, 1.25 * half, 0.0, -2.0 * half, 1.5 * half)]
    vals = np.array(vals, dtype=np.float32)
    fill = rng.uniform(-1.25 * half, 1.25 * half, size=max(count - vals.size, 0)).astype(np.float32)
    out = np.concatenate([vals, fill])[:count] if count >= vals.size else vals
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)

    import torch

    torch.set_default_device("cpu")
    torch.set_num_threads(1)
    R = _import_reference(args.ref)
    rng = np.random.default_rng(20261019)

    # ---------------------------------------------------------------- nodes
    nodes = {}
    for n in range(1, 17):
        basis = R.lp.get_lagrange_basis(n)
        nodes[f"X_{n}"] = basis.X.numpy().astype(np.float32)
        if n > 1:
            nodes[f"denom_{n}"] = basis.denominators.numpy().astype(np.float32)
    for n in (2, 3, 5):
        nodes[f"X_{n}_len4"] = R.lp.get_lagrange_basis(n, length=4.0).X.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(args.out, "nodes.npz"), **nodes)

    # ------------------------------------------------------- segment index
    seg = {}
    cases = []
    for S in (1, 2, 3, 4, 5, 7, 8, 10, 32, 64):
        for length in (2.0, 1.0, 3.0):
            for per in (None, 2.0):
                name = f"S{S}_L{length}_P{per}"
                xs = special_inputs(S, length, rng, 4 * S + 64)
                if per is not None:
                    xs = np.concatenate([xs, rng.uniform(-7, 7, size=64).astype(np.float32),
                                         np.array([-1.0, 1.0, 3.0, -3.0, 5.0, 1.0000001, 0.99999994], np.float32)])
                layer = R.pl.PiecewisePolynomial(n=3, in_features=1, out_features=1, segments=S,
                                                 length=length, periodicity=per)
                xt = torch.from_numpy(xs).reshape(-1, 1)
                idx = layer.which_segment(xt).numpy().reshape(-1)
                seg[name + "_x"] = xs
                seg[name + "_idx"] = idx.astype(np.int64)
                xin = xt if per is None else R.utils.make_periodic(xt, per)
                seg[name + "_xlocal"] = layer.x_local(xin, torch.from_numpy(idx).reshape(-1, 1)).numpy().reshape(-1)
                if per is not None:
                    seg[name + "_xper"] = xin.numpy().reshape(-1)
                cases.append(name)
    # NaN behaviour of the reference on CPU (documented: index 0)
    nanl = R.pl.PiecewisePolynomial(n=2, in_features=1, out_features=1, segments=4)
    seg["nan_idx"] = nanl.which_segment(torch.tensor([[float("nan")]])).numpy().reshape(-1).astype(np.int64)
    seg["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(args.out, "segment_index.npz"), **seg)

    # ------------------------------------------------------------ FC layers
    fc = {}
    fc_cases = []

    def run_fc(name, kind, n, S, I, O, B, length=2.0, per=None, special=True, scale=1.25):
        cls = R.pl.PiecewiseDiscontinuousPolynomial if kind == "disc" else R.pl.PiecewisePolynomial
        layer = cls(n=n, in_features=I, out_features=O, segments=S, length=length, periodicity=per)
        with torch.no_grad():
            layer.w.copy_(torch.from_numpy(rng.uniform(-1, 1, size=tuple(layer.w.shape)).astype(np.float32)))
        half = 0.5 * length
        x = rng.uniform(-scale * half, scale * half, size=(B, I)).astype(np.float32)
        if per is not None:
            x = rng.uniform(-3 * per, 3 * per, size=(B, I)).astype(np.float32)
        if special:
            sp = special_inputs(S, length, rng, 0)
            flat = x.reshape(-1)
            k = min(sp.size, flat.size // 2)
            pos = rng.choice(flat.size, size=k, replace=False)
            flat[pos] = sp[:k]
            if n > 1 and per is None and length == 2.0:
                # land exactly on interpolation nodes of some segment
                X = nodes[f"X_{n}"]
                for q in range(min(n, flat.size // 4)):
                    s = int(rng.integers(0, S))
                    flat[int(rng.integers(0, flat.size))] = np.float32(-1 + (s + (X[q] + 1) / 2) * 2.0 / S)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = layer(xt)
        gy = rng.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        fc[name + "_x"] = x
        fc[name + "_w"] = layer.w.detach().numpy().copy()
        fc[name + "_gy"] = gy
        fc[name + "_y"] = y.detach().numpy().copy()
        fc[name + "_dx"] = (np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy())  # n=1: x is disconnected
        fc[name + "_dw"] = layer.w.grad.numpy().copy()
        fc[name + "_seg"] = layer.which_segment(torch.from_numpy(x)).numpy().astype(np.int64)
        fc[name + "_meta"] = np.array([n, S, I, O, B, 1 if kind == "disc" else 0], dtype=np.int64)
        fc[name + "_fmeta"] = np.array([length, np.nan if per is None else per], dtype=np.float64)
        fc_cases.append(name)

    for kind in ("cont", "disc"):
        for n in (1, 2, 3, 4, 5, 8, 10):
            for S in (1, 2, 3, 4, 8):
                run_fc(f"{kind}_n{n}_S{S}", kind, n, S, I=5, O=3, B=9)
        run_fc(f"{kind}_medium", kind, 4, 4, I=37, O=19, B=33)
        run_fc(f"{kind}_S64", kind, 3, 64, I=6, O=4, B=40)
        run_fc(f"{kind}_S32_n8", kind, 8, 32, I=3, O=7, B=17)
        run_fc(f"{kind}_per2", kind, 4, 4, I=6, O=5, B=21, per=2.0)
        run_fc(f"{kind}_per2_S3", kind, 3, 3, I=4, O=2, B=30, per=2.0)
        run_fc(f"{kind}_len4", kind, 3, 5, I=4, O=3, B=15, length=4.0)
        run_fc(f"{kind}_len1", kind, 4, 3, I=4, O=3, B=15, length=1.0)
        run_fc(f"{kind}_wide", kind, 5, 8, I=16, O=130, B=12)
        run_fc(f"{kind}_1x1", kind, 3, 2, I=1, O=1, B=1)
    fc["cases"] = np.array(fc_cases)
    np.savez_compressed(os.path.join(args.out, "fc_layers.npz"), **fc)

    # -------------------------------------- the reference's own value tests
    pins = {}
    # tests/test_layers.py:88-95  (5-node Lagrange reproduces the identity at 0.5)
    poly = R.lp.LagrangePoly(5)
    w = R.lp.chebyshevLobatto(5).reshape(1, 1, 1, 5)
    pins["identity_ans"] = poly.interpolate(torch.tensor([[0.5]]), w).numpy()
    # tests/test_layers.py:98-105  (n = 1 constant)
    pins["constant_ans"] = R.lp.LagrangePoly(1).interpolate(torch.tensor([[0.5]]),
                                                           torch.tensor([0.25]).reshape(1, 1, 1, 1)).numpy()
    # tests/test_layers.py:108-131 / tests/test_sparse_optimizer.py:24-58
    torch.manual_seed(7)
    layer = R.pl.PiecewiseDiscontinuousPolynomial(n=2, in_features=1, out_features=1, segments=10)
    x = torch.ones(1, 1) * 0.5
    out = layer(x)
    loss = torch.nn.functional.mse_loss(out, torch.ones(1, 1) * 0.2)
    loss.backward()
    pins["act_w"] = layer.w.detach().numpy().copy()
    pins["act_out"] = out.detach().numpy().copy()
    pins["act_dw"] = layer.w.grad.numpy().copy()
    # tests/test_refinement.py:21-44  (p-refinement keeps the function: weights + outputs)
    ref_cases = []
    for kind, cls in (("cont", R.pl.PiecewisePolynomial), ("disc", R.pl.PiecewiseDiscontinuousPolynomial)):
        for (n_in, n_out, I, O, S) in [(1, 4, 3, 2, 5), (3, 5, 3, 2, 5), (5, 5, 2, 3, 2), (7, 5, 3, 2, 5)]:
            li = cls(n=n_in, in_features=I, out_features=O, segments=S)
            lo = cls(n=n_out, in_features=I, out_features=O, segments=S)
            R.pl.interpolate_polynomial_layer(layer_in=li, layer_out=lo)
            xx = torch.rand(2, I)
            nm = f"interp_{kind}_{n_in}_{n_out}_{I}_{O}_{S}"
            pins[nm + "_x"] = xx.numpy().copy()
            pins[nm + "_w_in"] = li.w.detach().numpy().copy()
            pins[nm + "_w_out"] = lo.w.detach().numpy().copy()
            pins[nm + "_y_in"] = li(xx).detach().numpy().copy()
            pins[nm + "_y_out"] = lo(xx).detach().numpy().copy()
            pins[nm + "_meta"] = np.array([n_in, n_out, I, O, S, 1 if kind == "disc" else 0])
            ref_cases.append(nm)
    pins["interp_cases"] = np.array(ref_cases)
    # constructor RNG consumption: same seed => same initial weights
    for kind, cls in (("cont", R.pl.PiecewisePolynomial), ("disc", R.pl.PiecewiseDiscontinuousPolynomial)):
        torch.manual_seed(1234)
        l1 = cls(n=4, in_features=6, out_features=5, segments=3, weight_magnitude=2.0)
        pins[f"init_{kind}_w"] = l1.w.detach().numpy().copy()
    torch.manual_seed(1234)
    c1 = R.fcv.PiecewisePolynomialConvolution2d(n=3, segments=4, in_channels=2, out_channels=3, kernel_size=3)
    pins["init_conv_w"] = c1.conv.weight.detach().numpy().copy()
    np.savez_compressed(os.path.join(args.out, "reference_pins.npz"), **pins)

    # ------------------------------------------------------------------ conv
    cv = {}
    cv_cases = []

    def run_conv(name, n, S, C, O, K, H, W, B, length=2.0, per=None, rescale=False, **kw):
        layer = R.fcv.PiecewisePolynomialConvolution2d(n=n, segments=S, in_channels=C, out_channels=O,
                                                      kernel_size=K, length=length, periodicity=per,
                                                      rescale_output=rescale, **kw)
        with torch.no_grad():
            layer.conv.weight.copy_(torch.from_numpy(
                rng.uniform(-1, 1, size=tuple(layer.conv.weight.shape)).astype(np.float32)))
        half = 0.5 * length
        x = rng.uniform(-1.2 * half, 1.2 * half, size=(B, C, H, W)).astype(np.float32)
        if per is not None:
            x = rng.uniform(-3 * per, 3 * per, size=(B, C, H, W)).astype(np.float32)
        sp = special_inputs(S, length, rng, 0)
        flat = x.reshape(-1)
        k = min(sp.size, flat.size // 3)
        flat[rng.choice(flat.size, size=k, replace=False)] = sp[:k]
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = layer(xt)
        gy = rng.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        cv[name + "_x"] = x
        cv[name + "_w"] = layer.conv.weight.detach().numpy().copy()
        cv[name + "_gy"] = gy
        cv[name + "_y"] = y.detach().numpy().copy()
        cv[name + "_dx"] = (np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy())
        cv[name + "_dw"] = layer.conv.weight.grad.numpy().copy()
        cv[name + "_expand"] = layer.poly(torch.from_numpy(x) if per is None
                                          else R.utils.make_periodic(torch.from_numpy(x), per)).numpy().copy()
        cv[name + "_meta"] = np.array([n, S, C, O, K, H, W, B, kw.get("stride", 1), kw.get("padding", 0),
                                       kw.get("dilation", 1), 1 if rescale else 0], dtype=np.int64)
        cv[name + "_fmeta"] = np.array([length, np.nan if per is None else per], dtype=np.float64)
        cv_cases.append(name)

    run_conv("sizes_test", 3, 4, 2, 2, 4, 5, 5, 1)                  # tests/test_convolution.py:68-94 shape
    run_conv("lenet1", 3, 4, 3, 6, 5, 12, 12, 2)
    run_conv("lenet2", 3, 4, 6, 16, 5, 9, 9, 2)
    run_conv("stride2_pad1", 4, 3, 2, 3, 3, 8, 7, 2, stride=2, padding=1)
    run_conv("dil2", 2, 5, 2, 2, 3, 9, 9, 1, dilation=2)
    run_conv("rescale", 3, 2, 2, 4, 3, 6, 6, 2, rescale=True)
    run_conv("per2", 3, 4, 2, 2, 3, 6, 6, 2, per=2.0)
    run_conv("len4", 3, 3, 2, 2, 3, 6, 6, 1, length=4.0)
    run_conv("n5_S1", 5, 1, 1, 2, 2, 4, 4, 1)
    cv["cases"] = np.array(cv_cases)
    np.savez_compressed(os.path.join(args.out, "conv2d.npz"), **cv)

    # ------------------------------------------------------------------- MLP
    mlp = {}
    mlp_cases = []

    def run_mlp(name, layer_type, n, in_w, out_w, hl, hw, S, B, norm=True, per=None):
        net = R.networks.HighOrderMLP(layer_type=layer_type, n=n, in_width=in_w, out_width=out_w,
                                      hidden_layers=hl, hidden_width=hw, in_segments=S, out_segments=S,
                                      hidden_segments=S, periodicity=per,
                                      normalization=R.layers.MaxAbsNormalization if norm else None)
        mods = [m for m in net.model if hasattr(m, "w")]
        with torch.no_grad():
            for m in mods:
                m.w.copy_(torch.from_numpy(rng.uniform(-1, 1, size=tuple(m.w.shape)).astype(np.float32)))
        x = rng.uniform(-1, 1, size=(B, in_w)).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = net(xt)
        gy = rng.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        mlp[name + "_x"] = x
        mlp[name + "_gy"] = gy
        mlp[name + "_y"] = y.detach().numpy().copy()
        mlp[name + "_dx"] = xt.grad.numpy().copy()
        for k, m in enumerate(mods):
            mlp[name + f"_w{k}"] = m.w.detach().numpy().copy()
            mlp[name + f"_dw{k}"] = m.w.grad.numpy().copy()
        mlp[name + "_meta"] = np.array([n, in_w, out_w, hl, hw, S, B, len(mods), 1 if norm else 0,
                                        1 if layer_type == "discontinuous" else 0], dtype=np.int64)
        mlp[name + "_keys"] = np.array(sorted(net.state_dict().keys()))
        mlp_cases.append(name)

    run_mlp("c1_function", "continuous", 3, 1, 1, 1, 10, 2, 32)     # BASELINE config 1 (function_example)
    run_mlp("c1_xor", "continuous", 3, 2, 1, 1, 10, 2, 32)          # BASELINE config 1 (xor)
    run_mlp("c3_small", "continuous", 4, 20, 10, 1, 16, 4, 8)       # config-3 topology, reduced widths
    run_mlp("disc_nonorm", "discontinuous", 3, 4, 3, 2, 6, 3, 8, norm=False)
    mlp["cases"] = np.array(mlp_cases)
    np.savez_compressed(os.path.join(args.out, "mlp.npz"), **mlp)

    # -------------------------------------------- SURVEY 8f-3 variants: discontinuous conv2d, Polynomial
    var = {}
    var_conv, var_poly = [], []
    for (name, n, S, C, O, K, H, W, B, kw) in [("dconv_a", 3, 4, 2, 3, 3, 7, 7, 2, {}),
                                               ("dconv_pad", 2, 5, 3, 2, 3, 6, 8, 2, dict(stride=2, padding=1)),
                                               ("dconv_n1", 1, 3, 2, 2, 2, 5, 5, 1, {})]:
        layer = R.fcv.PiecewiseDiscontinuousPolynomialConvolution2d(n=n, segments=S, in_channels=C, out_channels=O,
                                                                   kernel_size=K, **kw)
        with torch.no_grad():
            layer.conv.weight.copy_(torch.from_numpy(
                rng.uniform(-1, 1, size=tuple(layer.conv.weight.shape)).astype(np.float32)))
        x = rng.uniform(-1.2, 1.2, size=(B, C, H, W)).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = layer(xt)
        gy = rng.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        var[name + "_x"], var[name + "_w"], var[name + "_gy"] = x, layer.conv.weight.detach().numpy().copy(), gy
        var[name + "_y"] = y.detach().numpy().copy()
        var[name + "_dx"] = np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy()
        var[name + "_dw"] = layer.conv.weight.grad.numpy().copy()
        var[name + "_meta"] = np.array([n, S, C, O, K, H, W, B, kw.get("stride", 1), kw.get("padding", 0)], dtype=np.int64)
        var_conv.append(name)
    torch.manual_seed(1234)
    dc = R.fcv.PiecewiseDiscontinuousPolynomialConvolution2d(n=3, segments=4, in_channels=2, out_channels=3, kernel_size=3)
    var["init_dconv_w"] = dc.conv.weight.detach().numpy().copy()
    for (name, n, I, O, B, per) in [("poly_n3", 3, 5, 4, 9, None), ("poly_n5", 5, 7, 3, 11, None),
                                    ("poly_n1", 1, 3, 2, 4, None), ("poly_n8_per", 8, 4, 6, 10, 2.0)]:
        layer = R.pl.Polynomial(n=n, in_features=I, out_features=O, periodicity=per)
        with torch.no_grad():
            layer.w.copy_(torch.from_numpy(rng.uniform(-1, 1, size=tuple(layer.w.shape)).astype(np.float32)))
        x = rng.uniform(-1, 1, size=(B, I)).astype(np.float32)
        if per is not None:
            x = rng.uniform(-5, 5, size=(B, I)).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = layer(xt)
        gy = rng.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        var[name + "_x"], var[name + "_w"], var[name + "_gy"] = x, layer.w.detach().numpy().copy(), gy
        var[name + "_y"] = y.detach().numpy().copy()
        var[name + "_dx"] = np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy()
        var[name + "_dw"] = layer.w.grad.numpy().copy()
        var[name + "_meta"] = np.array([n, I, O, B], dtype=np.int64)
        var[name + "_fmeta"] = np.array([np.nan if per is None else per], dtype=np.float64)
        var[name + "_keys"] = np.array(sorted(layer.state_dict().keys()))
        var_poly.append(name)
    var["conv_cases"] = np.array(var_conv)
    var["poly_cases"] = np.array(var_poly)
    np.savez_compressed(os.path.join(args.out, "variants.npz"), **var)

    # -------------------------------------------- per-sample normalisations (utils.py:8-58)
    nr = {}
    nr_cases = []
    fns = {"max_abs": R.utils.max_abs_normalization, "max_center": R.utils.max_center_normalization,
           "l2": R.utils.l2_normalization}
    for kind, fn in fns.items():
        for (B, W) in [(1, 1), (3, 10), (64, 128), (17, 1000), (5, 33)]:
            x = rng.uniform(-2, 2, size=(B, W)).astype(np.float32)
            if B >= 3:
                x[1] = 0.0                       # an all-zero sample
                x[2, : max(W // 2, 1)] = x[2, 0]  # ties for the maximum
            xt = torch.from_numpy(x.copy()).requires_grad_(True)
            y = fn(xt)
            gy = rng.standard_normal(size=(B, W)).astype(np.float32)
            y.backward(torch.from_numpy(gy))
            name = f"{kind}_{B}x{W}"
            nr[name + "_x"], nr[name + "_gy"] = x, gy
            nr[name + "_y"], nr[name + "_dx"] = y.detach().numpy().copy(), xt.grad.numpy().copy()
            nr_cases.append(name)
    nr["cases"] = np.array(nr_cases)
    np.savez_compressed(os.path.join(args.out, "row_norms.npz"), **nr)

    # -------------------------------------------- p / h refinement, linear init, smoothing
    import random as pyrandom
    rfn = {}
    interp_cases, refine_cases, init_cases = [], [], []
    kinds = {"cont": R.pl.PiecewisePolynomial, "disc": R.pl.PiecewiseDiscontinuousPolynomial}
    for kind, cls in kinds.items():
        for (n_in, n_out, I, O, S) in [(1, 4, 3, 2, 5), (3, 5, 3, 2, 5), (5, 5, 2, 3, 2), (7, 5, 3, 2, 5), (4, 1, 2, 2, 3)]:
            li = cls(n=n_in, in_features=I, out_features=O, segments=S)
            lo = cls(n=n_out, in_features=I, out_features=O, segments=S)
            with torch.no_grad():
                li.w.copy_(torch.from_numpy(rng.uniform(-1, 1, size=tuple(li.w.shape)).astype(np.float32)))
            R.pl.interpolate_polynomial_layer(layer_in=li, layer_out=lo)
            name = f"interp_{kind}_{n_in}_{n_out}_S{S}"
            rfn[name + "_w_in"], rfn[name + "_w_out"] = li.w.detach().numpy().copy(), lo.w.detach().numpy().copy()
            rfn[name + "_meta"] = np.array([n_in, n_out, I, O, S, int(kind == "disc")], dtype=np.int64)
            interp_cases.append(name)
    for (n_in, n_out, S_in, S_out) in [(3, 3, 3, 9), (5, 5, 5, 5), (2, 2, 2, 12), (3, 4, 2, 6), (4, 2, 5, 3), (2, 1, 3, 4)]:
        li = R.pl.PiecewisePolynomial(n=n_in, in_features=3, out_features=2, segments=S_in)
        lo = R.pl.PiecewisePolynomial(n=n_out, in_features=3, out_features=2, segments=S_out)
        with torch.no_grad():
            li.w.copy_(torch.from_numpy(rng.uniform(-1, 1, size=tuple(li.w.shape)).astype(np.float32)))
        R.pl.refine_polynomial_layer(layer_in=li, layer_out=lo)
        name = f"refine_{n_in}_{n_out}_S{S_in}_{S_out}"
        rfn[name + "_w_in"], rfn[name + "_w_out"] = li.w.detach().numpy().copy(), lo.w.detach().numpy().copy()
        rfn[name + "_meta"] = np.array([n_in, n_out, S_in, S_out], dtype=np.int64)
        refine_cases.append(name)
    for kind, cls in kinds.items():
        layer = cls(n=4, in_features=3, out_features=2, segments=5)
        pyrandom.seed(7)
        R.pl.initialize_polynomial_layer(layer, max_slope=1.5, max_offset=0.3)
        rfn[f"init_{kind}_w"] = layer.w.detach().numpy().copy()
        init_cases.append(f"init_{kind}")
    sm = R.pl.PiecewiseDiscontinuousPolynomial(n=3, in_features=3, out_features=2, segments=5)
    with torch.no_grad():
        sm.w.copy_(torch.from_numpy(rng.uniform(-1, 1, size=tuple(sm.w.shape)).astype(np.float32)))
    rfn["smooth_w_in"] = sm.w.detach().numpy().copy()
    R.pl.smooth_discontinuous_layer(sm, 0.5)
    rfn["smooth_w_out"] = sm.w.detach().numpy().copy()
    rfn["interp_cases"], rfn["refine_cases"], rfn["init_cases"] = (np.array(interp_cases), np.array(refine_cases),
                                                                  np.array(init_cases))
    np.savez_compressed(os.path.join(args.out, "refinement.npz"), **rfn)

    # -------------------------------------------- SparseLion (sparse_optimizers/sparse_lion.py)
    from high_order_layers_torch.sparse_optimizers import SparseLion as RefSparseLion
    sl = {}
    p0 = rng.standard_normal(size=(4, 5, 12)).astype(np.float32)
    par = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = RefSparseLion([par], lr=1e-2, betas=(0.9, 0.99), weight_decay=0.1)
    grads = []
    for k in range(4):
        g = (rng.standard_normal(size=p0.shape) * (rng.uniform(size=p0.shape) < 0.3)).astype(np.float32)
        grads.append(g)
        par.grad = torch.from_numpy(g.copy())
        opt.step()
    sl["p0"], sl["grads"] = p0, np.stack(grads)
    sl["p_final"], sl["exp_avg_final"] = par.detach().numpy().copy(), opt.state[par]["exp_avg"].numpy().copy()
    sl["hyper"] = np.array([1e-2, 0.9, 0.99, 0.1])
    np.savez_compressed(os.path.join(args.out, "sparse_lion.npz"), **sl)

    # -------------------------------------------- 1-d convolutions (FunctionalConvolution.py:493-650)
    c1 = {}
    c1_cases = []
    kinds1 = {"cont": R.fcv.PiecewisePolynomialConvolution1d, "disc": R.fcv.PiecewiseDiscontinuousPolynomialConvolution1d}
    for kind, cls in kinds1.items():
        for (n, S, C, O, K, L, B, stride, dil, resc) in [(3, 4, 2, 3, 3, 11, 2, 1, 1, False), (2, 5, 3, 2, 4, 17, 2, 2, 1, True),
                                                         (4, 2, 1, 2, 3, 20, 3, 1, 2, False), (1, 3, 2, 2, 2, 7, 1, 1, 1, False)]:
            layer = cls(n=n, segments=S, in_channels=C, out_channels=O, kernel_size=K, stride=stride, dilation=dil,
                        rescale_output=resc)
            with torch.no_grad():
                layer.conv.weight.copy_(torch.from_numpy(
                    rng.uniform(-1, 1, size=tuple(layer.conv.weight.shape)).astype(np.float32)))
            x = rng.uniform(-1.2, 1.2, size=(B, C, L)).astype(np.float32)
            xt = torch.from_numpy(x.copy()).requires_grad_(True)
            y = layer(xt)
            gy = rng.standard_normal(size=tuple(y.shape)).astype(np.float32)
            y.backward(torch.from_numpy(gy))
            name = f"{kind}_n{n}_S{S}_K{K}"
            c1[name + "_x"], c1[name + "_w"], c1[name + "_gy"] = x, layer.conv.weight.detach().numpy().copy(), gy
            c1[name + "_y"] = y.detach().numpy().copy()
            c1[name + "_dx"] = np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy()
            c1[name + "_dw"] = layer.conv.weight.grad.numpy().copy()
            c1[name + "_meta"] = np.array([n, S, C, O, K, L, B, stride, dil, int(resc), int(kind == "disc")], dtype=np.int64)
            c1_cases.append(name)
    torch.manual_seed(1234)
    c1["init_cont_w"] = R.fcv.PiecewisePolynomialConvolution1d(n=3, segments=4, in_channels=2, out_channels=3,
                                                                kernel_size=3).conv.weight.detach().numpy().copy()
    c1["cases"] = np.array(c1_cases)
    np.savez_compressed(os.path.join(args.out, "conv1d.npz"), **c1)

    # -------------------------------------------- round 2 (own RNG: the files above stay bit-identical)
    # padded 1-d convolutions, paired 2-d geometry, conv-transpose (FunctionalConvolutionTranspose.py:75-160)
    import high_order_layers_torch.FunctionalConvolutionTranspose as fct
    rng2 = np.random.default_rng(20260101)
    v2 = {}
    names = []
    for kind, cls in kinds1.items():
        for (n, S, C, O, K, L, B, stride, pad, dil, resc) in [(3, 4, 2, 3, 3, 11, 2, 1, 1, 1, False),
                                                              (2, 5, 3, 2, 4, 17, 2, 2, 3, 1, True),
                                                              (4, 2, 1, 2, 3, 9, 3, 1, 2, 2, False)]:
            layer = cls(n=n, segments=S, in_channels=C, out_channels=O, kernel_size=K, stride=stride, padding=pad,
                        dilation=dil, rescale_output=resc)
            with torch.no_grad():
                layer.conv.weight.copy_(torch.from_numpy(
                    rng2.uniform(-1, 1, size=tuple(layer.conv.weight.shape)).astype(np.float32)))
            x = rng2.uniform(-1.2, 1.2, size=(B, C, L)).astype(np.float32)
            xt = torch.from_numpy(x.copy()).requires_grad_(True)
            y = layer(xt)
            gy = rng2.standard_normal(size=tuple(y.shape)).astype(np.float32)
            y.backward(torch.from_numpy(gy))
            name = f"conv1d_{kind}_n{n}_S{S}_K{K}_p{pad}"
            v2[name + "_x"], v2[name + "_w"], v2[name + "_gy"] = x, layer.conv.weight.detach().numpy().copy(), gy
            v2[name + "_y"] = y.detach().numpy().copy()
            v2[name + "_dx"] = np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy()
            v2[name + "_dw"] = layer.conv.weight.grad.numpy().copy()
            v2[name + "_meta"] = np.array([n, S, C, O, K, L, B, stride, pad, dil, int(resc), int(kind == "disc")], dtype=np.int64)
            names.append(name)
    v2["conv1d_cases"] = np.array(names)
    names = []
    for (n, S, C, O, K, H, W, B, stride, pad, dil) in [(3, 4, 2, 3, 3, 9, 7, 2, (2, 1), (1, 2), (1, 1)),
                                                     (2, 3, 1, 2, 2, 6, 8, 1, (1, 2), (0, 1), (2, 1))]:
        layer = R.fcv.PiecewisePolynomialConvolution2d(n=n, segments=S, in_channels=C, out_channels=O, kernel_size=K,
                                                       stride=stride, padding=pad, dilation=dil)
        with torch.no_grad():
            layer.conv.weight.copy_(torch.from_numpy(
                rng2.uniform(-1, 1, size=tuple(layer.conv.weight.shape)).astype(np.float32)))
        x = rng2.uniform(-1.2, 1.2, size=(B, C, H, W)).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = layer(xt)
        gy = rng2.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        name = f"conv2d_n{n}_S{S}_K{K}_s{stride[0]}{stride[1]}"
        v2[name + "_x"], v2[name + "_w"], v2[name + "_gy"] = x, layer.conv.weight.detach().numpy().copy(), gy
        v2[name + "_y"], v2[name + "_dx"] = y.detach().numpy().copy(), xt.grad.numpy().copy()
        v2[name + "_dw"] = layer.conv.weight.grad.numpy().copy()
        v2[name + "_meta"] = np.array([n, S, C, O, K, H, W, B, stride[0], stride[1], pad[0], pad[1], dil[0], dil[1]],
                                      dtype=np.int64)
        names.append(name)
    v2["conv2d_cases"] = np.array(names)
    names = []
    for (n, S, C, O, K, H, W, B, stride, pad, opad, dil, resc, per) in [
            (3, 4, 2, 3, 3, 5, 6, 2, 1, 0, 0, 1, False, None), (2, 3, 3, 2, 4, 4, 4, 2, 2, 1, 1, 1, True, None),
            (4, 2, 1, 2, 3, 6, 5, 1, 2, 0, 0, 2, False, 2.0), (1, 3, 2, 2, 2, 3, 3, 2, 1, 0, 0, 1, False, None)]:
        layer = fct.PiecewisePolynomialConvolutionTranspose2d(n=n, segments=S, in_channels=C, out_channels=O, kernel_size=K,
                                                              stride=stride, padding=pad, output_padding=opad, dilation=dil,
                                                              rescale_output=resc, periodicity=per)
        with torch.no_grad():
            layer.conv.weight.copy_(torch.from_numpy(
                rng2.uniform(-1, 1, size=tuple(layer.conv.weight.shape)).astype(np.float32)))
        x = rng2.uniform(-1.2, 1.2, size=(B, C, H, W)).astype(np.float32)
        if per is not None:
            x = rng2.uniform(-5, 5, size=(B, C, H, W)).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = layer(xt)
        gy = rng2.standard_normal(size=tuple(y.shape)).astype(np.float32)
        y.backward(torch.from_numpy(gy))
        name = f"convT_n{n}_S{S}_K{K}_s{stride}"
        v2[name + "_x"], v2[name + "_w"], v2[name + "_gy"] = x, layer.conv.weight.detach().numpy().copy(), gy
        v2[name + "_y"] = y.detach().numpy().copy()
        v2[name + "_dx"] = np.zeros_like(x) if xt.grad is None else xt.grad.numpy().copy()
        v2[name + "_dw"] = layer.conv.weight.grad.numpy().copy()
        v2[name + "_meta"] = np.array([n, S, C, O, K, H, W, B, stride, pad, opad, dil, int(resc)], dtype=np.int64)
        v2[name + "_fmeta"] = np.array([np.nan if per is None else per], dtype=np.float64)
        names.append(name)
    v2["convT_cases"] = np.array(names)
    torch.manual_seed(1234)
    v2["init_convT_w"] = fct.PiecewisePolynomialConvolutionTranspose2d(
        n=3, segments=4, in_channels=2, out_channels=3, kernel_size=3).conv.weight.detach().numpy().copy()
    np.savez_compressed(os.path.join(args.out, "variants2.npz"), **v2)

    total = sum(os.path.getsize(os.path.join(args.out, f)) for f in os.listdir(args.out))
    print(f"wrote {len(os.listdir(args.out))} files, {total/1e3:.1f} kB, to {os.path.abspath(args.out)}")


if __name__ == "__main__":
    main()
