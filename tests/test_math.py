"""The arithmetic specification (fb_math.h, evaluated through the C oracle) against float64
references and the known answers of the reference's float32 phi."""
import numpy as np
import pytest


def ulp_err(y, ref):
    ref32 = ref.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(y.astype(np.float64) - ref) / ulp


@pytest.fixture(scope="module")
def rng():
    return np.random.default_rng(0)


def test_exp_log_accuracy(oracle, rng):
    x = np.concatenate([rng.uniform(-87, 88, 500000), rng.uniform(-1, 1, 200000)]).astype(np.float32)
    assert ulp_err(oracle.math_fn("expf", x), np.exp(x.astype(np.float64))).max() < 1.0
    assert oracle.math_fn("expf", np.array([0.0, -0.0], np.float32)).tolist() == [1.0, 1.0]
    tiny = oracle.math_fn("expf", np.array([-87.0, -87.5, -104.0, -1e30], np.float32))     # clamped at exp(-87)
    assert np.all(tiny == tiny[0]) and 1.1e-38 < tiny[0] < 2e-38
    x = np.exp(rng.uniform(-80, 80, 500000)).astype(np.float32)
    assert ulp_err(oracle.math_fn("logf", x), np.log(x.astype(np.float64))).max() < 1.0
    assert oracle.math_fn("logf", np.array([1.0], np.float32))[0] == 0.0


def test_softplus_logaddexp_tanh_accuracy(oracle, rng):
    x = rng.uniform(-110, 110, 500000).astype(np.float32)
    xd = x.astype(np.float64)
    ref = np.where(xd > 13.942385, xd, np.where(xd < -13.942385, np.exp(xd), np.log1p(np.exp(xd))))
    y = oracle.math_fn("softplusf", x)
    m = ref > 1e-37
    assert ulp_err(y[m], ref[m]).max() < 30           # worst case only for |result| ~ 1e-6 (crude 1/u correction)
    big = ref > 1e-4
    assert ulp_err(y[big], ref[big]).max() < 2.5
    a, b = rng.uniform(-60, 60, 300000).astype(np.float32), rng.uniform(-60, 60, 300000).astype(np.float32)
    ref = np.logaddexp(a.astype(np.float64), b.astype(np.float64))
    y = oracle.math_fn("logaddexpf", a, b)
    assert np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1.0)) < 2.5e-7
    x = rng.uniform(-9, 9, 500000).astype(np.float32)
    assert ulp_err(oracle.math_fn("tanhf", x), np.tanh(x.astype(np.float64))).max() < 6
    assert oracle.math_fn("tanhf", np.array([0.0, 1e-5, 30.0, -30.0], np.float32)).tolist() == \
        pytest.approx([0.0, 1e-5, 1.0, -1.0], rel=1e-6)


def test_phi_known_answers(oracle):
    """phi(8.5e-8)=16.635532, phi(16.635532)=0, phi(1e-3)=7.6008792,
    phi(1)=0.7719368 in float32; clipping at both ends (decoding_q.py:372)."""
    x = np.array([8.5e-8, 16.635532, 1e-3, 1.0, 0.0, 1e-9, 50.0], np.float32)
    y4 = oracle.math_fn("phi4f", x)
    assert y4[0] == np.float32(16.635532) and y4[1] == 0.0
    assert abs(float(y4[2]) - 7.6008792) < 1e-6 and abs(float(y4[3]) - 0.7719368) < 1e-7
    assert y4[4] == y4[0] and y4[5] == y4[0] and y4[6] == 0.0
    y2 = oracle.math_fn("phi2f", x)
    assert y2[0] == np.float32(16.635532) and y2[1] == 0.0 and abs(float(y2[3]) - 0.7719368) < 1e-7


def test_phi_against_float32_numpy(oracle, rng):
    """phi evaluated with numpy's float32 exp/log/log1p in the reference's order of operations
    agrees except for the cancellation noise the reference formula itself carries."""
    from oracle import np_oracle as N
    x = np.exp(rng.uniform(np.log(1e-4), np.log(12.0), 300000)).astype(np.float32)
    a, b = oracle.math_fn("phi4f", x), N.phi4(x)
    # absolute error bounded by a few ulps of the two terms being subtracted (~x + |log x|) plus the
    # one-ulp ambiguity of exp(x) ~ 1 amplified by log(exp(x) - 1): 2^-23 / x
    scale = np.spacing(np.maximum(x, np.abs(np.log(x))).astype(np.float32))
    assert np.all(np.abs(a - b) <= 4 * scale + 1.5 * 2.0 ** -23 / x)
    assert np.mean(a == b) > 0.5


def test_saturation_identities_used_by_the_kernels(oracle, rng):
    """The CUDA kernels skip the polynomial evaluation when every lane of a warp is saturated
    (fbgnn_kernels.cuh: phi_sat, logaddexp_sat).  These are the identities that makes exact."""
    hi = np.concatenate([[16.635532], rng.uniform(16.635532, 200, 100000), [1e30]]).astype(np.float32)
    lo = np.concatenate([[8.5e-8, 0.0, -0.0], -rng.uniform(0, 200, 100000), rng.uniform(0, 8.5e-8, 1000)]).astype(np.float32)
    for fn in ("phi4f", "phi2f"):
        assert np.all(oracle.math_fn(fn, hi).view(np.uint32) == 0)                          # +0.0
        assert np.all(oracle.math_fn(fn, lo).view(np.uint32) == np.float32(16.635532).view(np.uint32))
    a = rng.uniform(-120, 120, 1000000).astype(np.float32)
    b = (a + rng.uniform(-60, 0, 1000000)).astype(np.float32)
    mx, mn = np.maximum(a, b), np.minimum(a, b)
    sat = (mn - mx).astype(np.float32) < np.float32(-17.5)
    r = oracle.math_fn("logaddexpf", a, b)
    assert sat.sum() > 100000
    assert np.array_equal(r[sat].view(np.uint32), (np.float32(0) + mx[sat]).view(np.uint32))
    assert np.array_equal(oracle.math_fn("logaddexpf", b, a)[sat].view(np.uint32), r[sat].view(np.uint32))
