"""Pin the CPU oracle (numpy and C restatements) against the reference's own outputs.

The golden vectors in tests/golden/*.npz were produced by the unmodified reference
core/corr.py (see tests/golden/make_golden.py).  Tolerances: the reference lookup goes
through F.grid_sample's normalise/denormalise round trip (core/utils/utils.py:84-99), so
lookup values differ from the direct sampler formula by ~1e-5 of max.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_max

VOL_TOL = 2e-6      # fp32 GEMM re-association
LOOKUP_TOL = 5e-5   # grid_sample coordinate round trip
GRAD_TOL = 5e-5


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_volume_and_pyramid(oracle, name, impl):
    g = load_golden(name)
    L = g["dims"]["L"]
    vol = (oracle.numpy_corr_volume if impl == "numpy" else oracle.c_corr_volume)(g["f1"], g["f2"])
    assert vol.shape == g["vol"].shape
    assert rel_max(vol, g["vol"]) < VOL_TOL
    # pyramid from the reference's own level 0: must be bit-exact ((a+b)/2 in fp32)
    levels = (oracle.numpy_pyramid if impl == "numpy" else oracle.c_pyramid)(g["lvl0"].copy(), L)
    for i in range(L):
        assert levels[i].shape == g[f"lvl{i}"].shape
        np.testing.assert_array_equal(levels[i], g[f"lvl{i}"])
    assert [lv.shape[-1] for lv in levels] == oracle.level_widths(g["dims"]["W2"], L)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_lookup(oracle, name, impl):
    g = load_golden(name)
    d = g["dims"]
    levels = [np.ascontiguousarray(g[f"lvl{i}"]) for i in range(d["L"])]
    fn = oracle.numpy_lookup if impl == "numpy" else oracle.c_lookup
    for t in range(d["T"]):
        out = fn(levels, np.ascontiguousarray(g["coords"][t]), d["r"])
        assert out.shape == g["out"][t].shape
        assert rel_max(out, g["out"][t]) < LOOKUP_TOL
    if "out_alt" in g:   # the "alt" block computes the same quantity (core/corr.py:64-107)
        out = fn(levels, np.ascontiguousarray(g["coords"][0]), d["r"])
        assert rel_max(out, g["out_alt"]) < 2e-4


@pytest.mark.parametrize("name", golden_names())
def test_numpy_and_c_agree(oracle, name):
    g = load_golden(name)
    d = g["dims"]
    levels = [np.ascontiguousarray(g[f"lvl{i}"]) for i in range(d["L"])]
    for t in range(d["T"]):
        c = np.ascontiguousarray(g["coords"][t])
        a = oracle.numpy_lookup(levels, c, d["r"])
        b = oracle.c_lookup(levels, c, d["r"])
        assert rel_max(a, b) < 1e-6


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_lookup_backward(oracle, name, impl):
    g = load_golden(name)
    d = g["dims"]
    widths = oracle.level_widths(d["W2"], d["L"])
    rd = 2 * d["r"] + 1
    glv = []
    for i in range(d["L"]):
        acc = np.zeros((d["B"], d["H"], d["W1"], widths[i]), dtype=np.float32)
        for t in range(d["T"]):
            grad = np.ascontiguousarray(g["gout"][t][:, i * rd:(i + 1) * rd])
            coords = np.ascontiguousarray(g["coords"][t])
            if impl == "numpy":
                acc += oracle.numpy_lookup_level_bwd(coords[:, 0] / np.float32(2 ** i), grad,
                                                     widths[i], d["r"])
            else:
                oracle.c_lookup_level_bwd(coords, grad, widths[i], d["r"], 0.5 ** i, into=acc)
        assert rel_max(acc, g[f"glvl{i}"]) < GRAD_TOL
        glv.append(acc)
    # fold + volume backward reproduces the reference's autograd w.r.t. the feature maps
    fold = (oracle.numpy_fold_pyramid_grad if impl == "numpy" else oracle.c_fold_pyramid_grad)(glv)
    bwd = oracle.numpy_corr_volume_bwd if impl == "numpy" else oracle.c_corr_volume_bwd
    df1, df2 = bwd(g["f1"], g["f2"], np.ascontiguousarray(fold))
    assert rel_max(df1, g["df1"]) < GRAD_TOL
    assert rel_max(df2, g["df2"]) < GRAD_TOL


def test_c_backward_overwrite_mode(oracle):
    g = load_golden("odd37")
    d = g["dims"]
    rd = 2 * d["r"] + 1
    grad = np.ascontiguousarray(g["gout"][0][:, :rd])
    coords = np.ascontiguousarray(g["coords"][0])
    fresh = oracle.c_lookup_level_bwd(coords, grad, d["W2"], d["r"])
    ref = oracle.numpy_lookup_level_bwd(coords[:, 0], grad, d["W2"], d["r"])
    assert rel_max(fresh, ref) < 1e-6


def test_whole_path_baseline_matches_pieces(oracle):
    g = load_golden("odd37")
    d = g["dims"]
    base = oracle.CorrPathBaseline(d["B"], d["C"], d["H"], d["W1"], d["W2"], d["L"], d["r"], d["T"])
    out = base.run(g["f1"], g["f2"], np.ascontiguousarray(g["coords"]))
    assert rel_max(out, g["out"]) < LOOKUP_TOL


def test_nonfinite_coords_give_zero(oracle):
    vol = np.ones((1, 1, 4, 8), dtype=np.float32)
    coords = np.array([np.nan, np.inf, -np.inf, 3.0e9], dtype=np.float32).reshape(1, 1, 1, 4)
    for fn in (oracle.numpy_lookup, oracle.c_lookup):
        out = fn([vol], coords, 2)
        assert np.all(out == 0)


@pytest.mark.parametrize("name", __import__("conftest").convc1_names())
def test_oracle_conv1x1_matches_reference_motion_encoder(oracle, name):
    """The restated consumer layer (relu(convc1(corr)), core/update.py:70,77) against the
    reference's own BasicMotionEncoder output stored by make_golden.py."""
    from conftest import load_convc1
    g = load_golden(name)
    c = load_convc1(name)
    assert c["weight"].shape == (64, g["dims"]["L"] * (2 * g["dims"]["r"] + 1))
    for t in range(g["dims"]["T"]):
        got = oracle.numpy_conv1x1(g["out"][t], c["weight"], c["bias"], relu=True)
        assert rel_max(got, c["cor"][t]) < 2e-6
        assert (c["cor"][t] == 0).mean() > 0.2          # the ReLU does clip in these fixtures
