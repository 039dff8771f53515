"""CPU-side checks: module API mirrors the reference (constructor arguments, parameter names and
shapes, RNG consumption, registry errors), the C-ABI library loads and exports every declared
symbol, and the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import high_order_layers_torch_b200 as hol
from high_order_layers_torch_b200 import _build, _lib, ops
from high_order_layers_torch_b200.LagrangePolynomial import LagrangePoly, chebyshevLobatto
from tests.conftest import ROOT, load_golden


def test_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hol_b200.h")).read()
    declared = set(re.findall(r"\b(hol_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_build.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.hol_abi_version() == 1
    assert loaded.hol_error_string(-2).startswith(b"hol:")
    assert loaded.hol_error_string(0) == b"ok"


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    c2 = lib.hol_pw_workspace_bytes(65536, 1024, 1024, 5, 8, 1)
    assert 5e8 < c2 < 2e10      # tensor-core path: + the Phi tiles of the whole batch (10.7 GB) and the operand tiles
    prev = ops.set_plan_overrides(_lib.HOL_PLAN_DISABLE_TENSOR_CORES)
    try:
        assert 5e8 < lib.hol_pw_workspace_bytes(65536, 1024, 1024, 5, 8, 1) < 2e9   # gather path only
    finally:
        assert ops.set_plan_overrides(prev) == _lib.HOL_PLAN_DISABLE_TENSOR_CORES
    assert lib.hol_pw_workspace_bytes(8, 4, 4, 11, 2, 0) == 0          # n above the compiled envelope
    assert lib.hol_pw_workspace_bytes(8, 4, 4, 3, 5000, 0) == 0        # too many segments
    assert lib.hol_pw_conv2d_workspace_bytes(4, 3, 32, 32, 6, 5, 1, 0, 1, 3, 4, 0) > 0
    assert lib.hol_pw_conv2d_workspace_bytes(4, 3, 4, 4, 6, 5, 1, 0, 1, 3, 4, 0) == 0   # kernel larger than image
    assert lib.hol_pw_launch_count(0, 64, 8, 4, 3, 2, 0, 1, 1, 0) == 3       # prep, pack, forward (gather path)
    assert lib.hol_pw_launch_count(1, 64, 8, 4, 3, 2, 0, 1, 1, 1) == 2       # dX, dense-register dW (narrow layer)
    prev = ops.set_plan_overrides(_lib.HOL_PLAN_DISABLE_DENSE_DW)
    try:
        assert lib.hol_pw_launch_count(1, 64, 8, 4, 3, 2, 0, 1, 1, 1) == 3   # dX, bucket sort, bucketed dW
    finally:
        ops.set_plan_overrides(prev)
    assert lib.hol_pw_launch_count(0, 65536, 1024, 1024, 5, 8, 1, 1, 1, 0) > 3   # tensor-core path


def test_node_table_is_the_reference_table():
    g = load_golden("nodes.npz")
    for n in range(1, 17):
        assert np.array_equal(chebyshevLobatto(n).numpy().view(np.int32), g[f"X_{n}"].view(np.int32)), n
        assert np.array_equal(LagrangePoly(n).basis.X.numpy(), g[f"X_{n}"])
    for n in (2, 3, 5):
        assert np.array_equal(ops.lobatto_nodes(n, 4.0).numpy(), g[f"X_{n}_len4"])
    for n in range(2, 11):
        assert np.array_equal(LagrangePoly(n).basis.denominators.numpy(), g[f"denom_{n}"])


def test_same_seed_same_initial_weights():
    p = load_golden("reference_pins.npz")
    torch.manual_seed(1234)
    a = hol.PiecewisePolynomial(n=4, in_features=6, out_features=5, segments=3, weight_magnitude=2.0)
    assert a.w.shape == (5, 6, 10) and a.w.requires_grad
    assert np.array_equal(a.w.detach().numpy(), p["init_cont_w"])
    torch.manual_seed(1234)
    b = hol.PiecewiseDiscontinuousPolynomial(n=4, in_features=6, out_features=5, segments=3, weight_magnitude=2.0)
    assert b.w.shape == (5, 6, 12)
    assert np.array_equal(b.w.detach().numpy(), p["init_disc_w"])
    torch.manual_seed(1234)
    c = hol.PiecewisePolynomialConvolution2d(n=3, segments=4, in_channels=2, out_channels=3, kernel_size=3)
    assert c.conv.weight.shape == (3, 18, 3, 3)
    assert np.array_equal(c.conv.weight.detach().numpy(), p["init_conv_w"])


def test_factories_and_errors():
    layer = hol.high_order_fc_layers("continuous", n=3, in_features=2, out_features=4, segments=5,
                                     rescale_output=False, scale=2.0, initialization="constant_random")
    assert isinstance(layer, hol.PiecewisePolynomial) and layer.w.shape == (4, 2, 11)
    layer = hol.high_order_fc_layers("discontinuous", n=3, in_features=2, out_features=4, segments=5)
    assert isinstance(layer, hol.PiecewiseDiscontinuousPolynomial) and layer.w.shape == (4, 2, 15)
    assert layer.n == 3 and layer._segments == 5 and layer._length == 2.0 and layer._half == 1.0
    conv = hol.high_order_convolution_layers("continuous2d", n=3, segments=4, in_channels=2, out_channels=2,
                                             kernel_size=4, stride=1, device="cpu")
    assert conv.conv.weight.shape == (2, 18, 4, 4)
    with pytest.raises(ValueError, match="not recognized"):
        hol.high_order_fc_layers("fourier", n=3, in_features=2, out_features=4, segments=5)
    with pytest.raises(ValueError, match="not recognized"):
        hol.high_order_convolution_layers("fourier2d", n=3, in_channels=2, out_channels=2, kernel_size=3)
    with pytest.raises(ValueError):
        hol.PiecewisePolynomialConvolution2d(n=3, segments=2, in_channels=2, out_channels=2, kernel_size=3, groups=2)


def test_high_order_mlp_structure_matches_reference():
    g = load_golden("mlp.npz")
    net = hol.HighOrderMLP(layer_type="continuous", n=3, in_width=1, out_width=1, hidden_layers=1, hidden_width=10,
                           in_segments=2, out_segments=2, hidden_segments=2, normalization=hol.MaxAbsNormalization)
    assert sorted(net.state_dict().keys()) == [str(k) for k in g["c1_function_keys"]]
    shapes = [tuple(m.w.shape) for m in net.model if hasattr(m, "w")]
    assert shapes == [g[f"c1_function_w{k}"].shape for k in range(3)]
    assert net.input_layer is net.model[0] and net.output_layer is net.model[-1]


def test_no_cpu_fallback():
    layer = hol.PiecewisePolynomial(n=3, in_features=2, out_features=2, segments=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros(4, 2))
    conv = hol.PiecewisePolynomialConvolution2d(n=3, segments=2, in_channels=1, out_channels=1, kernel_size=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(torch.zeros(1, 1, 4, 4))
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.piecewise_polynomial(torch.zeros(4, 2), layer.w, 3, 2, 2.0, False, None)


def test_cpu_helpers_match_reference_segment_index():
    g = load_golden("segment_index.npz")
    for name in g["cases"]:
        name = str(name)
        S = int(name.split("_")[0][1:])
        length = float(name.split("_")[1][1:])
        per = name.split("_")[2][1:]
        per = None if per == "None" else float(per)
        layer = hol.PiecewisePolynomial(n=3, in_features=1, out_features=1, segments=S, length=length,
                                        periodicity=per)
        x = torch.from_numpy(g[name + "_x"]).reshape(-1, 1)
        idx = layer.which_segment(x)
        assert idx.dtype == torch.int64
        assert np.array_equal(idx.numpy().reshape(-1), g[name + "_idx"]), name


def test_fake_tensor_shapes():
    # register_fake: shape/dtype propagation without a device
    x = torch.empty(7, 5, device="meta")
    w = torch.empty(3, 5, 9, device="meta")
    y = ops.piecewise_polynomial(x, w, 3, 4, 2.0, False, None)
    assert y.shape == (7, 3) and y.dtype == torch.float32
    xi = torch.empty(2, 3, 8, 8, device="meta")
    wf = torch.empty(4, 27, 9, device="meta")
    yr, ws = ops.pw_conv2d_rows(xi, wf, 3, 4, 2.0, False, None, [3, 3], [1, 1], [0, 0], [1, 1])
    assert yr.shape == (2 * 6 * 6, 4) and ws.dtype == torch.uint8
    assert ws.numel() == _lib.load().hol_pw_conv2d_workspace_bytes(2, 3, 8, 8, 4, 3, 1, 0, 1, 3, 4, 0)
    # autograd through the fake ops (shape propagation of the gradients)
    xm = torch.empty(7, 5, device="meta", requires_grad=True)
    wm = torch.empty(3, 5, 9, device="meta", requires_grad=True)
    ym = ops.piecewise_polynomial(xm, wm, 3, 4, 2.0, False, None)
    assert ym.requires_grad


def test_conv_transpose_host_side():
    g = load_golden("variants2.npz")
    torch.manual_seed(1234)
    ct = hol.high_order_convolution_transpose_layers("continuous2d", n=3, segments=4, in_channels=2, out_channels=3,
                                                     kernel_size=3)
    assert ct.conv.weight.shape == (18, 3, 3, 3)
    assert np.array_equal(ct.conv.weight.detach().numpy(), g["init_convT_w"])
    with pytest.raises(ValueError, match="not recognized"):
        hol.high_order_convolution_transpose_layers("fourier2d", n=3, in_channels=2, out_channels=2, kernel_size=3)
    with pytest.raises(RuntimeError, match="CUDA"):
        ct(torch.zeros(1, 2, 4, 4))


def test_variant_layers_host_side():
    g = load_golden("variants.npz")
    poly = hol.high_order_fc_layers("polynomial", n=3, in_features=5, out_features=4)
    assert sorted(poly.state_dict().keys()) == [str(k) for k in g["poly_n3_keys"]]
    assert poly.w.shape == (4, 5, 3)
    with pytest.raises(ValueError):
        hol.Polynomial(n=3, in_features=2, out_features=2, length=4.0)
    torch.manual_seed(1234)
    dc = hol.high_order_convolution_layers("discontinuous2d", n=3, segments=4, in_channels=2, out_channels=3,
                                           kernel_size=3)
    assert dc.conv.weight.shape == (3, 24, 3, 3)
    assert np.array_equal(dc.conv.weight.detach().numpy(), g["init_dconv_w"])


def test_refinement_utilities_match_reference_fixtures():
    """the batched p / h refinement, linear initialisation and smoothing reproduce the
    reference's loop implementations (fixture refinement.npz, PolynomialLayers.py:757-1023)."""
    import random

    import numpy as np
    import torch

    import high_order_layers_torch_b200 as hol
    from tests.conftest import load_golden

    g = load_golden("refinement.npz")
    for name in g["interp_cases"]:
        name = str(name)
        n_in, n_out, I, O, S, disc = (int(v) for v in g[name + "_meta"])
        cls = hol.PiecewiseDiscontinuousPolynomial if disc else hol.PiecewisePolynomial
        li, lo = cls(n=n_in, in_features=I, out_features=O, segments=S), cls(n=n_out, in_features=I, out_features=O, segments=S)
        with torch.no_grad():
            li.w.copy_(torch.from_numpy(g[name + "_w_in"]))
        li.interpolate(lo)
        assert np.abs(lo.w.detach().numpy() - g[name + "_w_out"]).max() < 1e-6, name
    for name in g["refine_cases"]:
        name = str(name)
        n_in, n_out, S_in, S_out = (int(v) for v in g[name + "_meta"])
        li = hol.PiecewisePolynomial(n=n_in, in_features=3, out_features=2, segments=S_in)
        lo = hol.PiecewisePolynomial(n=n_out, in_features=3, out_features=2, segments=S_out)
        with torch.no_grad():
            li.w.copy_(torch.from_numpy(g[name + "_w_in"]))
        li.refine(lo)
        assert np.abs(lo.w.detach().numpy() - g[name + "_w_out"]).max() < 1e-6, name
    for kind, cls in (("cont", hol.PiecewisePolynomial), ("disc", hol.PiecewiseDiscontinuousPolynomial)):
        layer = cls(n=4, in_features=3, out_features=2, segments=5)
        random.seed(7)
        hol.initialize_polynomial_layer(layer, max_slope=1.5, max_offset=0.3)
        assert np.abs(layer.w.detach().numpy() - g[f"init_{kind}_w"]).max() < 1e-6, kind
    sm = hol.PiecewiseDiscontinuousPolynomial(n=3, in_features=3, out_features=2, segments=5)
    with torch.no_grad():
        sm.w.copy_(torch.from_numpy(g["smooth_w_in"]))
    hol.smooth_discontinuous_layer(sm, 0.5)
    assert np.abs(sm.w.detach().numpy() - g["smooth_w_out"]).max() < 1e-6
    import pytest
    with pytest.raises(ValueError):
        sm.refine(sm)
    with pytest.raises(ValueError):
        hol.interpolate_polynomial_layer(hol.PiecewisePolynomial(n=2, in_features=1, out_features=1, segments=2),
                                         hol.PiecewisePolynomial(n=2, in_features=1, out_features=1, segments=3))


def _run_sparse_lion(device):
    import numpy as np
    import torch

    from high_order_layers_torch_b200.sparse_optimizers import SparseLion
    from tests.conftest import load_golden

    g = load_golden("sparse_lion.npz")
    lr, b1, b2, wd = (float(v) for v in g["hyper"])
    par = torch.nn.Parameter(torch.from_numpy(g["p0"].copy()).to(device))
    opt = SparseLion([par], lr=lr, betas=(b1, b2), weight_decay=wd)
    for grad in g["grads"]:
        par.grad = torch.from_numpy(grad.copy()).to(device)
        opt.step()
    assert np.abs(par.detach().cpu().numpy() - g["p_final"]).max() < 1e-6
    assert np.abs(opt.state[par]["exp_avg"].cpu().numpy() - g["exp_avg_final"]).max() < 1e-6
    untouched = np.all(g["grads"] == 0, axis=0)
    assert np.array_equal(par.detach().cpu().numpy()[untouched], g["p0"][untouched])   # never decayed, never moved


def test_no_torch_fallbacks_in_optimizer_and_normalisations():
    """north_star: no CPU fallback anywhere -- SparseLion and the normalisation helpers refuse CPU tensors
    like the layers do (the fixture comparison of SparseLion runs on the GPU: test_sparse_lion_fused_kernel)."""
    from high_order_layers_torch_b200 import utils as U
    from high_order_layers_torch_b200.sparse_optimizers import SparseLion
    par = torch.nn.Parameter(torch.ones(3, 4))
    par.grad = torch.ones(3, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        SparseLion([par], lr=1e-2).step()
    for fn in (U.max_abs_normalization, U.max_center_normalization, U.l2_normalization, U.max_abs_normalization_nd,
               U.max_abs_normalization_last, U.max_center_normalization_nd):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn(torch.ones(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        hol.MaxAbsNormalization()(torch.ones(2, 3))


def test_conv1d_unfold_mapping_and_init_match_reference():
    """the 1-d convolution is the 2-d path with H = Kh = 1: the weight mapping conv.weight
    [O, C*V, K] -> w_fc [O, C*K, V] and the tap order c*K + k are checked on CPU with the oracle standing
    in for the CUDA op (fixtures from FunctionalConvolution.py:493-650, incl. zero padding), plus the
    initial weights for a seed."""
    from oracle import pw_oracle as orc
    for fname, key, padded in (("conv1d.npz", "cases", False), ("variants2.npz", "conv1d_cases", True)):
        g = load_golden(fname)
        for name in g[key]:
            name = str(name)
            meta = [int(v) for v in g[name + "_meta"]]
            if padded:
                n, S, C, O, K, L, B, stride, pad, dil, resc, disc = meta
            else:
                (n, S, C, O, K, L, B, stride, dil, resc, disc), pad = meta, 0
            nb = n * S if disc else (n - 1) * S + 1
            x = g[name + "_x"]
            xp = np.pad(x, ((0, 0), (0, 0), (pad, pad)))
            Lo = (L + 2 * pad - dil * (K - 1) - 1) // stride + 1
            taps = np.arange(Lo)[:, None] * stride + np.arange(K)[None, :] * dil                  # [Lo, K]
            rows = xp[:, :, taps].transpose(0, 2, 1, 3).reshape(B * Lo, C * K)                     # column c*K + k
            valid = ((taps >= pad) & (taps < pad + L))                                             # zero-padded taps add nothing
            w_fc = torch.from_numpy(g[name + "_w"]).unsqueeze(2).view(O, C, nb, 1, K).permute(0, 1, 3, 4, 2) \
                .reshape(O, C * K, nb).numpy()
            nodes = ops.lobatto_nodes(n).numpy()
            # the oracle has no invalid-tap flag: evaluate tap by tap and mask
            y_rows = np.zeros((B * Lo, O))
            for k in range(K):
                cols = [c * K + k for c in range(C)]
                part = orc.pw_forward(rows[:, cols], w_fc[:, cols], n, S, 2.0, bool(disc), None, nodes=nodes)
                y_rows += part * np.tile(valid[:, k], B)[:, None]
            y = y_rows.reshape(B, Lo, O).transpose(0, 2, 1) * (1.0 / (C * K * K) if resc else 1.0)
            ref = g[name + "_y"]
            assert y.shape == ref.shape, name
            assert np.abs(y - ref).max() <= 2e-6 * max(np.abs(ref).max(), 1e-30), name
    g = load_golden("conv1d.npz")
    torch.manual_seed(1234)
    layer = hol.high_order_convolution_layers("continuous1d", n=3, segments=4, in_channels=2, out_channels=3,
                                              kernel_size=3)
    assert np.array_equal(layer.conv.weight.detach().numpy(), g["init_cont_w"])
    padded = hol.PiecewisePolynomialConvolution1d(n=3, segments=4, in_channels=2, out_channels=3, kernel_size=3, padding=1)
    assert padded.conv.padding == (1,)
    with pytest.raises(ValueError):
        hol.PiecewisePolynomialConvolution1d(n=3, segments=4, in_channels=2, out_channels=3, kernel_size=3, groups=2)


def test_planner_sweep_is_consistent():
    """The host-side planner (make_shape / tc_plan / dwd_plan behind the ABI queries) over a grid of
    shapes: every supported shape gets a non-zero workspace that fits the device, a positive
    launch count and a consistent set of plan flags; nothing crashes or loops (no GPU involved)."""
    lib = _lib.load()
    shapes = [(B, I, O, n, S, d)
              for B in (1, 33, 256, 257, 8192, 65536, 802816)
              for (I, O) in ((1, 1), (3, 2), (75, 6), (150, 16), (128, 10), (784, 128), (1024, 1024), (33, 65))
              for (n, S) in ((1, 1), (2, 1), (3, 4), (5, 8), (4, 32), (10, 3), (2, 64))
              for d in (0, 1)]
    for (B, I, O, n, S, d) in shapes:
        ws = lib.hol_pw_workspace_bytes(B, I, O, n, S, d)
        assert ws > 0, (B, I, O, n, S, d)
        # (not monotonic in B: above 40 GiB of Phi tiles the tensor-core path switches to batch chunks)
        assert lib.hol_pw_workspace_bytes(2 * B, I, O, n, S, d) > 0
        assert ws < 1.2e11, (B, I, O, n, S, d, ws)      # every planned shape fits well inside 180 GB of HBM
        f = ops.plan_flags(B, I, O, n, S, bool(d))
        assert not (f["dense_dw"] and f["tc_dw"])
        assert not f["shared_phi"] or (f["tc_forward"] and f["tc_dw"])
        if O < 32:
            assert not (f["tc_forward"] or f["tc_dx"] or f["tc_dw"])
            smax = 2 if S <= 2 else (4 if S <= 4 else 8)
            nbmax = 1 if (not d and n == 1) else (n * smax if d else (n - 1) * smax + 1)
            assert f["dense_dw"] == (n <= 5 and S <= 8 and nbmax * 4 <= 128)   # 128 accumulators per thread
        if f["rotated_gather"]:
            assert S > 8
        for kind, dx, dw in ((0, 1, 1), (1, 1, 1), (1, 0, 1), (1, 1, 0)):
            c = lib.hol_pw_launch_count(kind, B, I, O, n, S, d, dx, dw, 1)
            # dX of n == 1 is a memset, not a kernel -- but a plan with a tensor-core dW takes max|gy| in whichever
            # half of a split backward comes first (the other half is told HOL_GY_STATS_VALID)
            expect_some = kind == 0 or dw or (dx and n > 1) or (dx and f["tc_dw"])
            assert (0 < c < 512) if expect_some else c == 0, (kind, B, I, O, n, S, d, c)
    # conv geometry
    for (Bi, C, H, W, O, K, st, pad, dil) in ((4, 3, 32, 32, 6, 5, 1, 0, 1), (2, 6, 14, 14, 16, 5, 1, 2, 1),
                                            (1, 1, 7, 9, 2, 3, 2, 1, 2)):
        assert lib.hol_pw_conv2d_workspace_bytes(Bi, C, H, W, O, K, st, pad, dil, 3, 4, 0) > 0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver times next to ours) runs without a GPU: the unmodified
    reference from baseline/_ref when it is installed (kind "reference"), the numpy oracle port otherwise, and
    prints ONE JSON line with the same metric / unit / config as our arm plus the keys the contract names."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fwd+bwd samples/sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("c3:")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    have_ref = os.path.isfile(os.path.join(root, "baseline", "_ref", "high_order_layers_torch", "PolynomialLayers.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")


def test_our_arm_refuses_to_run_without_a_gpu():
    """There is no CPU path: without a CUDA device `bench.py` (our arm) exits with an error instead of timing
    anything, and a layer called on CPU tensors raises."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=root)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    layer = hol.PiecewisePolynomial(n=3, in_features=4, out_features=2, segments=2)
    with pytest.raises((ValueError, RuntimeError)):
        layer(torch.zeros(5, 4))
