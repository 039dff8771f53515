"""Pin the numpy oracle against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pw_oracle as orc
from tests.conftest import load_golden, rel_err

TOL = 1e-5  # north_star: fp32 outputs and gradients within 1e-5 relative
MLP_TOL = 1e-4


def test_node_tables_within_one_ulp_of_reference():
    g = load_golden("nodes.npz")
    for n in range(1, 17):
        ref = g[f"X_{n}"]
        mine = orc.lobatto_nodes_f32(n)
        ulp = np.abs(ref.view(np.int32).astype(np.int64) - mine.view(np.int32).astype(np.int64))
        # signed-zero / tiny middle node: compare values there
        assert np.all((ulp <= 1) | (np.abs(ref - mine) < 1e-12)), (n, ref, mine)
    # documented fp32 artefacts of the reference table
    assert g["X_5"][2] == np.float32(4.371139e-08)
    assert g["X_1"][0] == 0.0
    for n in (2, 3, 5):
        assert np.array_equal(g[f"X_{n}_len4"], (np.float32(2.0) * g[f"X_{n}"]).astype(np.float32))


def test_denominator_table_matches_reference():
    g = load_golden("nodes.npz")
    for n in range(2, 17):
        X = g[f"X_{n}"]
        D = g[f"denom_{n}"]
        for j in range(n):
            for m in range(n):
                want = np.float32(1.0) if m == j else np.float32(X[j] - X[m])
                assert D[j, m] == want


def test_segment_index_bit_exact():
    g = load_golden("segment_index.npz")
    for name in g["cases"]:
        name = str(name)
        S = int(name.split("_")[0][1:])
        length = float(name.split("_")[1][1:])
        per = name.split("_")[2][1:]
        per = None if per == "None" else float(per)
        x = g[name + "_x"]
        idx = orc.segment_index(x, S, length, per)
        assert idx.dtype == np.int64
        assert np.array_equal(idx, g[name + "_idx"]), name
        xin = x
        if per is not None:
            xin, _ = orc.make_periodic_f32(x, per)
            assert np.array_equal(xin.view(np.int32), g[name + "_xper"].view(np.int32)), name
        xl, _ = orc.local_coordinate_f32(xin, idx, S, length)
        assert np.array_equal(xl.view(np.int32), g[name + "_xlocal"].view(np.int32)), name
    assert np.array_equal(orc.segment_index(np.array([np.nan], np.float32), 4), g["nan_idx"])


def test_survey_boundary_examples():
    # SURVEY "Hard parts": S=4 measured examples
    x = np.array([-0.5, -0.50000006, 0.4999999, 0.5, 1.0, 1.5, -2.0], dtype=np.float32)
    assert orc.segment_index(x, 4).tolist() == [1, 0, 2, 3, 3, 3, 0]


def _fc_args(g, name):
    n, S, I, O, B, disc = (int(v) for v in g[name + "_meta"])
    length, per = (float(v) for v in g[name + "_fmeta"])
    per = None if np.isnan(per) else per
    return n, S, bool(disc), length, per


def test_fc_forward_backward_matches_reference():
    g = load_golden("fc_layers.npz")
    nodes = load_golden("nodes.npz")
    for name in g["cases"]:
        name = str(name)
        n, S, disc, length, per = _fc_args(g, name)
        X = nodes[f"X_{n}"]
        x, w, gy = g[name + "_x"], g[name + "_w"], g[name + "_gy"]
        assert np.array_equal(orc.segment_index(x, S, length, per), g[name + "_seg"]), name
        y = orc.pw_forward(x, w, n, S, length, disc, per, nodes=X)
        dx, dw = orc.pw_backward(gy, x, w, n, S, length, disc, per, nodes=X)
        assert rel_err(y, g[name + "_y"]) < TOL, name
        assert rel_err(dw, g[name + "_dw"]) < TOL, name
        if n == 1:
            assert not dx.any()
        else:
            assert rel_err(dx, g[name + "_dx"]) < TOL, name
        # dense dW with exact zeros wherever the reference has them (tests/test_layers.py:126-130)
        assert np.array_equal(dw == 0, g[name + "_dw"] == 0), name
        y32 = orc.pw_forward_reference_style(x, w, n, S, length, disc, per, nodes=X)
        assert rel_err(y32, g[name + "_y"]) < TOL, name


def test_reference_value_pins():
    p = load_golden("reference_pins.npz")
    nodes = load_golden("nodes.npz")
    # tests/test_layers.py:88-95
    w = nodes["X_5"].reshape(1, 1, 5)
    y = orc.pw_forward(np.array([[0.5]], np.float32), w, 5, 1, discontinuous=True, nodes=nodes["X_5"])
    assert abs(y[0, 0] - 0.5) < 1e-6 and abs(p["identity_ans"].item() - 0.5) < 1e-6
    # tests/test_layers.py:98-105
    y = orc.pw_forward(np.array([[0.5]], np.float32), np.full((1, 1, 1), 0.25, np.float32), 1, 1,
                       discontinuous=True)
    assert abs(y[0, 0] - 0.25) < 1e-6 and abs(p["constant_ans"].item() - 0.25) < 1e-6
    # tests/test_layers.py:108-131: exactly n=2 non-zero dW entries, shape [1,1,20]
    x = np.full((1, 1), 0.5, np.float32)
    y = orc.pw_forward(x, p["act_w"], 2, 10, discontinuous=True)
    assert rel_err(y, p["act_out"]) < TOL
    gy = 2.0 * (y - 0.2)
    _, dw = orc.pw_backward(gy, x, p["act_w"], 2, 10, discontinuous=True)
    assert dw.shape == (1, 1, 20) and np.count_nonzero(dw) == 2
    assert rel_err(dw, p["act_dw"]) < TOL
    # tests/test_refinement.py:21-44: p-refined weights give the same function
    for name in p["interp_cases"]:
        name = str(name)
        n_in, n_out, I, O, S, disc = (int(v) for v in p[name + "_meta"])
        x = p[name + "_x"]
        yi = orc.pw_forward(x, p[name + "_w_in"], n_in, S, discontinuous=bool(disc), nodes=nodes[f"X_{n_in}"])
        yo = orc.pw_forward(x, p[name + "_w_out"], n_out, S, discontinuous=bool(disc), nodes=nodes[f"X_{n_out}"])
        assert rel_err(yi, p[name + "_y_in"]) < TOL and rel_err(yo, p[name + "_y_out"]) < TOL
        if n_in <= n_out:
            assert np.allclose(yi, yo, rtol=1e-5, atol=1e-6)


def test_conv_matches_reference_and_unfold_identity():
    g = load_golden("conv2d.npz")
    nodes = load_golden("nodes.npz")
    for name in g["cases"]:
        name = str(name)
        n, S, C, O, K, H, W, B, stride, padding, dilation, rescale = (int(v) for v in g[name + "_meta"])
        length, per = (float(v) for v in g[name + "_fmeta"])
        per = None if np.isnan(per) else per
        X = (np.float32(length / 2.0) * nodes[f"X_{n}"]).astype(np.float32)
        x, w = g[name + "_x"], g[name + "_w"]
        xp = x if per is None else orc.make_periodic_f32(x, per)[0]
        ex = orc.expand2d(xp, n, S, length, nodes=X)
        assert rel_err(ex, g[name + "_expand"]) < TOL, name
        assert np.array_equal(ex == 0, g[name + "_expand"] == 0), name
        scale = 1.0 / (C * K * K) if rescale else 1.0
        y = orc.pw_conv2d_forward(x, w, n, S, length, False, per, stride, padding, dilation, scale, nodes=X)
        assert rel_err(y, g[name + "_y"]) < TOL, name
        if padding == 0:
            # conv == FC layer applied to unfolded patches
            cols, (b, ho, wo) = orc.unfold_nchw(x, K, stride, padding, dilation)
            wfc = orc.conv_weight_as_fc(w, C, n, S)
            yfc = orc.pw_forward(cols, wfc, n, S, length, False, per, nodes=X) * scale
            yfc = yfc.reshape(b, ho, wo, O).transpose(0, 3, 1, 2)
            assert rel_err(yfc, g[name + "_y"]) < TOL, name


def test_mlp_matches_reference():
    g = load_golden("mlp.npz")
    nodes = load_golden("nodes.npz")
    for name in g["cases"]:
        name = str(name)
        n, in_w, out_w, hl, hw, S, B, nl, norm, disc = (int(v) for v in g[name + "_meta"])
        ws = [g[name + f"_w{k}"] for k in range(nl)]
        y, dx, dws = orc.mlp_forward_backward(g[name + "_x"], ws, g[name + "_gy"], n, S, bool(disc),
                                              normalize=bool(norm), nodes=nodes[f"X_{n}"])
        # network level: fp32 rounding of one layer is amplified by the slope of the next
        # (S*|w| per link), so the bar is the reference's own MLP-level rtol of 1e-4
        # (tests/test_refinement.py:55-92); the 1e-5 bar is per layer (tests above).
        assert rel_err(y, g[name + "_y"]) < MLP_TOL, name
        assert rel_err(dx, g[name + "_dx"]) < MLP_TOL, name
        for k in range(nl):
            assert rel_err(dws[k], g[name + f"_dw{k}"]) < MLP_TOL, (name, k)
        # HighOrderMLP registers first/last layer twice (checkpoint keys)
        keys = set(str(k) for k in g[name + "_keys"])
        assert {"input_layer.w", "output_layer.w", "model.0.w"} <= keys


def test_variants_discontinuous_conv_and_polynomial():
    """the discontinuous 2-d convolution and the single-segment Polynomial layer."""
    g = load_golden("variants.npz")
    nodes = load_golden("nodes.npz")
    for name in g["conv_cases"]:
        name = str(name)
        n, S, C, O, K, H, W, B, stride, padding = (int(v) for v in g[name + "_meta"])
        y = orc.pw_conv2d_forward(g[name + "_x"], g[name + "_w"], n, S, 2.0, True, None, stride, padding, 1, 1.0,
                                  nodes=nodes[f"X_{n}"])
        assert rel_err(y, g[name + "_y"]) < TOL, name
    for name in g["poly_cases"]:
        name = str(name)
        n, I, O, B = (int(v) for v in g[name + "_meta"])
        per = float(g[name + "_fmeta"][0])
        per = None if np.isnan(per) else per
        x, w, gy = g[name + "_x"], g[name + "_w"], g[name + "_gy"]
        y = orc.pw_forward(x, w, n, 1, 2.0, True, per, nodes=nodes[f"X_{n}"])
        dx, dw = orc.pw_backward(gy, x, w, n, 1, 2.0, True, per, nodes=nodes[f"X_{n}"])
        assert rel_err(y, g[name + "_y"]) < TOL, name
        assert rel_err(dw, g[name + "_dw"]) < TOL, name
        if n > 1:
            assert rel_err(dx, g[name + "_dx"]) < TOL, name


def test_row_norms_match_reference():
    """the oracle's normalisations and their closed-form gradients against the
    reference's utils.py:8-58 + autograd (fixture row_norms.npz, incl. all-zero rows and ties)."""
    g = load_golden("row_norms.npz")
    for name in g["cases"]:
        name = str(name)
        fwd, bwd = orc.ROW_NORMS[name.rsplit("_", 1)[0]]
        x, gy = g[name + "_x"], g[name + "_gy"]
        assert rel_err(fwd(x), g[name + "_y"]) < 1e-6, name
        # width-1 rows: the gradient is a cancellation residue (~eps), compare absolutely
        assert np.abs(bwd(gy, x) - g[name + "_dx"]).max() < 1e-6 * max(np.abs(g[name + "_dx"]).max(), 1.0), name


def test_dx_is_invariant_under_a_constant_per_link():
    """The tcgen05 dX GEMM runs on link-centred weights (w - mean over the link's nodes, csrc/tc_build.cuh):
    inside every segment sum_j phi'_j = 0 (up to the rounding of the fp32 node table), so a constant per
    (output, input) link does not change dX -- continuous and discontinuous layers, any n > 1 and S."""
    rng = np.random.default_rng(3)
    for disc in (False, True):
        for n, S in ((2, 3), (3, 4), (5, 8), (8, 2)):
            I, O, B = 6, 5, 40
            nb = orc.weights_per_link(n, S, disc)
            x = rng.uniform(-1.2, 1.2, size=(B, I)).astype(np.float32)
            w = rng.uniform(-1, 1, size=(O, I, nb))
            gy = rng.standard_normal(size=(B, O))
            shift = rng.uniform(-5, 5, size=(O, I, 1))
            nodes = orc.lobatto_nodes_f32(n)
            dx0, dw0 = orc.pw_backward(gy, x, w, n, S, 2.0, disc, None, nodes=nodes)
            dx1, dw1 = orc.pw_backward(gy, x, w + shift, n, S, 2.0, disc, None, nodes=nodes)
            # not exactly: the reference's fp32 denominators make sum_j phi_j = 1 + O(1e-7), so a shift of
            # 5x the weight range moves dX by ~1e-7 relative -- the size of the artefact the centring drops
            assert np.abs(dx1 - dx0).max() < 1e-6 * np.abs(dx0).max(), (disc, n, S)
            assert np.array_equal(dw0, dw1)   # dW does not depend on w at all


def test_round2_variants_against_reference_fixtures():
    """variants2.npz (padded 1-d convolutions, paired 2-d geometry, conv-transpose): the oracle's
    restatements reproduce the reference's outputs."""
    g = load_golden("variants2.npz")
    nodes = load_golden("nodes.npz")
    for name in g["conv1d_cases"]:
        name = str(name)
        n, S, C, O, K, L, B, stride, pad, dil, resc, disc = (int(v) for v in g[name + "_meta"])
        y = orc.pw_conv1d_forward(g[name + "_x"], g[name + "_w"], n, S, 2.0, bool(disc), None, stride, pad, dil,
                                  1.0 / (C * K * K) if resc else 1.0, nodes=nodes[f"X_{n}"])
        assert y.shape == g[name + "_y"].shape and rel_err(y, g[name + "_y"]) < 2e-6, name
    for name in g["conv2d_cases"]:
        name = str(name)
        n, S, C, O, K, H, W, B, s0, s1, p0, p1, d0, d1 = (int(v) for v in g[name + "_meta"])
        y = orc.pw_conv2d_forward(g[name + "_x"], g[name + "_w"], n, S, 2.0, False, None, (s0, s1), (p0, p1), (d0, d1),
                                  nodes=nodes[f"X_{n}"])
        assert y.shape == g[name + "_y"].shape and rel_err(y, g[name + "_y"]) < 2e-6, name
    for name in g["convT_cases"]:
        name = str(name)
        n, S, C, O, K, H, W, B, stride, pad, opad, dil, resc = (int(v) for v in g[name + "_meta"])
        per = float(g[name + "_fmeta"][0])
        per = None if np.isnan(per) else per
        y = orc.pw_conv_transpose2d_forward(g[name + "_x"], g[name + "_w"], n, S, 2.0, per, stride, pad, opad, dil,
                                            1.0 / (C * K * K) if resc else 1.0, nodes=nodes[f"X_{n}"])
        assert y.shape == g[name + "_y"].shape and rel_err(y, g[name + "_y"]) < 2e-6, name
