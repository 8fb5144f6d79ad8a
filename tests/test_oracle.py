"""Pins the CPU oracle: Philox known answers, the reference's deterministic notebook outputs,
cross-check against the independent numpy restatement, and published logical error rates."""
import numpy as np
import pytest


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kat:
        assert oracle.philox(ctr, key).tolist() == out


def test_pauli_statistics_and_thresholds(oracle):
    """pauli.py:100-108 with px=pz=2p/3, py=p/3: [0,p/3) -> Y?  no: u<px is X-type, u in [px-py, px+pz-py) Z-type."""
    p = 0.12
    nx, nz = oracle.pauli(1, 0, 4000, 500, p)
    assert abs(nx.mean() - 2 * p / 3) < 2e-3 and abs(nz.mean() - 2 * p / 3) < 2e-3
    assert abs((nx & nz).mean() - p / 3) < 2e-3           # Y errors
    assert abs((nx | nz).mean() - p) < 2e-3
    # frame-id addressing: a shard starting at frame 1000 equals rows 1000.. of the full run
    a, _ = oracle.pauli(1, 0, 1200, 64, p)
    b, _ = oracle.pauli(1, 1000, 200, 64, p)
    assert np.array_equal(a[1000:], b)
    from oracle import np_oracle as N
    u = np.array([0.0, 0.0399, 0.04, 0.0799, 0.08, 0.1199, 0.12, 0.5], np.float32)
    x, z = N.pauli_from_uniform(u, p)
    assert x.tolist() == [1, 1, 1, 1, 0, 0, 0, 0] and z.tolist() == [0, 0, 1, 1, 1, 1, 0, 0]


def test_first_stage_marginal_extrema_kat(oracle, c1270):
    """examples/n1270.ipynb cell 12: after 64 iterations (f=1.0, p0=0.05) the marginals saturate at
    log(57) +- 3*phi_max and +- 6*phi_max:  max [53.9496498 103.856247 53.9496498],
    min [-45.8635445 -95.7701416 -45.8635445]."""
    g = oracle.CodeGraph(c1270)
    B = 200
    nx, nz = oracle.pauli(3, 0, B, c1270.N, 0.045)       # the notebook's data set has weight 10..80 errors
    sx = (c1270.hx @ nz.T.astype(np.int64)) & 1
    sz = (c1270.hz @ nx.T.astype(np.int64)) & 1
    prior = oracle.prior_llr(0.05)
    assert prior == np.float32(4.0430512)
    r = oracle.bp4(g, float(prior), sx, sz, 64)
    mx = [float(r[k].max()) for k in ("Lx", "Ly", "Lz")]
    mn = [float(r[k].min()) for k in ("Lx", "Ly", "Lz")]
    assert mx == pytest.approx([53.9496498, 103.856247, 53.9496498], abs=2e-5)
    assert mn == pytest.approx([-45.8635445, -95.7701416, -45.8635445], abs=2e-5)
    assert mx[0] == float(np.float32(prior + np.float32(3) * np.float32(16.635532)))


_CANCEL = 4e-6         # absolute noise of a check's phi sum in units of exp(|m|): ~2 float32 quanta at x >= 16 (1.9e-6 each)


def _tol(ref, cn_type, got=None):
    """Float32 noise model of one check-node update.  A message of magnitude |m| leaves the phi
    formula as -log(T/2) with T ~ 2 exp(-|m|) a sum of phi values that are themselves differences of
    two ~equal float32 numbers (quantum 2^-23 .. 2^-20): the absolute error of the message grows
    like exp(|m|) until it saturates at the clip (phi_max = 16.64).  Min-sum has no such term."""
    ar = np.abs(ref) if got is None else np.maximum(np.abs(ref), np.abs(got))
    tol = 2e-5 * ar + 1e-5
    if cn_type != "minsum":
        tol = tol + np.minimum(_CANCEL * np.exp(np.minimum(ar, 20.0)), 4.0)
    return tol


def _stack(r):
    return np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)


@pytest.mark.parametrize("name,cn_type", [("c882", "boxplus-phi"), ("c882", "minsum"), ("c882", "boxplus"),
                                          ("c1270", "boxplus-phi"), ("c1270", "minsum")])
def test_c_oracle_vs_numpy_oracle_teacher_forced(oracle, codes, c1270, name, cn_type):
    """One BP4 iteration from the C oracle's own state, recomputed with numpy's libm and numpy's
    reductions.  The two float32 evaluations agree to ~1e-6 except where the reference formula
    phi(x) = softplus(x) - log(exp(x)-1) cancels: there the error is bounded in
    absolute terms by the quantisation of the two ~equal terms."""
    from oracle import np_oracle as N
    code = c1270 if name == "c1270" else codes[name]
    g = oracle.CodeGraph(code)
    X, Z = N.Side(code.hx), N.Side(code.hz)
    assert np.array_equal(X.cn_of_edge, g.X.vn_cn) and np.array_equal(Z.cn_of_edge, g.Z.vn_cn)
    B = 48
    nx, nz = oracle.pauli(5, 0, B, code.N, 0.1)
    sx = (code.hx @ nz.T.astype(np.int64)) & 1
    sz = (code.hz @ nx.T.astype(np.int64)) & 1
    llr = np.full((B, 3, code.N), oracle.prior_llr(0.05), np.float32)
    ssx, ssz = (1 - 2 * sx).astype(np.float32), (1 - 2 * sz).astype(np.float32)
    for k in (0, 1, 4, 15):
        a = oracle.bp4(g, llr, sx, sz, k, 0.9, cn_type, want_msgs=True)
        b = oracle.bp4(g, llr, sx, sz, k + 1, 0.9, cn_type, want_msgs=True)
        mx, mz, _, _ = N.bp4_iteration(X, Z, a["msg_x"].T.copy(), a["msg_z"].T.copy(), llr.transpose(1, 2, 0),
                                       ssx, ssz, 0.9, cn_type)
        for got, ref in ((mx.T, b["msg_x"]), (mz.T, b["msg_z"])):
            err = np.abs(got - ref)
            assert np.all(err <= _tol(ref, cn_type, got)), (k, float(err.max()))
            weak = np.abs(ref) < 6.0        # away from the saturation regime the agreement is ~1 ulp
            if weak.any():
                assert np.median(err[weak] / np.maximum(np.abs(ref[weak]), 1e-6)) < 1e-5
    # epilogue: marginals, decisions, soft syndromes from the same messages
    a = oracle.bp4(g, llr, sx, sz, 32, want_msgs=True)
    n_ = N.bp4(X, Z, llr, sx, sz, 0, init=(a["msg_x"].T.copy(), a["msg_z"].T.copy()))
    for key in ("Lx", "Ly", "Lz"):
        assert np.allclose(n_[key], a[key], rtol=1e-5, atol=1e-4)
    assert np.array_equal(n_["x_hat"], a["x_hat"]) and np.array_equal(n_["z_hat"], a["z_hat"])
    for key in ("x_logit", "z_logit"):
        assert np.all(np.abs(n_[key] - a[key]) <= _tol(a[key], "boxplus-phi", n_[key]))


@pytest.mark.parametrize("name", ["c882", "c1270"])
def test_c_oracle_vs_numpy_oracle_gnn_and_bp2(oracle, codes, c1270, weights, name):
    from oracle import np_oracle as N
    code = c1270 if name == "c1270" else codes[name]
    g = oracle.CodeGraph(code)
    X, Z = N.Side(code.hx), N.Side(code.hz)
    B = 32
    nx, nz = oracle.pauli(6, 0, B, code.N, 0.1)
    sx = (code.hx @ nz.T.astype(np.int64)) & 1
    sz = (code.hz @ nx.T.astype(np.int64)) & 1
    r = oracle.bp4(g, float(oracle.prior_llr(0.05)), sx, sz, 16)
    for red in ("mean", "sum", "max", "min"):
        oc = oracle.gnn(g, oracle.Gnn(weights[name], "tanh", red), _stack(r), r["z_logit"], r["x_logit"], sx, sz)
        on = N.gnn(X, Z, weights[name], _stack(r), r["z_logit"], r["x_logit"], sx, sz, reduce_op=red)
        assert np.allclose(oc, on, rtol=1e-5, atol=2e-6), red
    # binary decoder, one iteration at a time is not needed: it has no cancellation in the VN update
    noise = oracle.bsc(2, 0, B, code.N, 0.03)
    synd = (code.hx @ noise.T.astype(np.int64)) & 1
    llr = np.full((B, code.N), -np.log((1 - 0.2) / 0.2), np.float32)
    for it in (1, 2):
        sc, hc = oracle.bp2(code.hx, llr, synd, it)
        sn, hn = N.bp2(X, llr, synd, it)
        assert np.all(np.abs(sc - sn) <= 3 * _tol(sc, "boxplus-phi", sn))


def test_pipeline_flags_equal_dense_reference_formulation(oracle, codes, weights):
    """any(hx_perp . x_diff) == any(hz . x_diff) or any(lz . x_diff), checked on the
    oracle's own residual errors; skip_inactive is result-identical to full work."""
    code = codes["c882"]
    g = oracle.CodeGraph(code)
    G = oracle.Gnn(weights["c882"])
    r = oracle.pipeline(g, [32, 8, 8], [G, G], 0.13, seed=1, B=160, want_diff=True)
    xd, zd = r["x_diff"].astype(np.int64), r["z_diff"].astype(np.int64)
    flagged = np.any((xd @ code.hz.T) & 1, 1) | np.any((zd @ code.hx.T) & 1, 1)
    block = np.any((xd @ code.hx_perp.T) & 1, 1) | np.any((zd @ code.hz_perp.T) & 1, 1)
    assert np.array_equal(r["flags"] & 1, flagged.astype(np.uint8))
    assert np.array_equal((r["flags"] >> 1) & 1, block.astype(np.uint8))
    assert r["counters"].tolist() == [160, int(flagged.sum()), int(block.sum()), int(((r["flags"] >> 2) > 0).sum())]
    r2 = oracle.pipeline(g, [32, 8, 8], [G, G], 0.13, seed=1, B=160, skip_inactive=True, want_diff=True)
    assert np.array_equal(r["flags"], r2["flags"]) and np.array_equal(r["x_diff"], r2["x_diff"])
    # sharding by global frame id: two halves == the whole
    a = oracle.pipeline(g, [32, 8, 8], [G, G], 0.13, seed=1, first_frame=0, B=80, skip_inactive=True)
    b = oracle.pipeline(g, [32, 8, 8], [G, G], 0.13, seed=1, first_frame=80, B=80, skip_inactive=True)
    assert np.array_equal(np.concatenate([a["flags"], b["flags"]]), r["flags"])


def test_numpy_pipeline_matches_c_pipeline_statistically(oracle, codes, weights):
    """The numpy restatement of the whole Sandwich model (dense hx_perp products, masked scatter)
    on the same noise: identical where trajectories agree; overall counts within sampling noise."""
    from oracle import np_oracle as N
    code = codes["c882"]
    g = oracle.CodeGraph(code)
    B, p = 96, 0.12
    nx, nz = oracle.pauli(8, 0, B, code.N, p)
    rc = oracle.pipeline(g, [32, 16], [oracle.Gnn(weights["c882"])], p, B=B, noise=(nx, nz))
    s_hat, ls_hat = N.pipeline(code, N.Side(code.hx), N.Side(code.hz), [32, 16], [weights["c882"]], nx, nz,
                               oracle.prior_llr(0.05))
    fl = np.any(s_hat, 1).astype(np.uint8) | (np.any(ls_hat, 1).astype(np.uint8) << 1)
    agree = np.mean(fl == (rc["flags"] & 3))
    assert agree > 0.9, agree
    assert abs(int(np.any(ls_hat, 1).sum()) - int(rc["counters"][2])) <= 6


def _cp_interval(k, n, alpha=1e-3):
    from scipy.stats import beta
    lo = 0.0 if k == 0 else beta.ppf(alpha / 2, k, n - k + 1)
    hi = 1.0 if k == n else beta.ppf(1 - alpha / 2, k + 1, n - k)
    return lo, hi


def test_published_logical_error_rates(oracle, codes, weights):
    """examples/n882.ipynb cell 2: (64,G,16,G,16,G,16), f=1.0, p0=0.05 -> 396/5000 block errors at
    p=0.12; examples/QLDPC.ipynb cell 12: plain BP4 64 it., f=0.8, p0=0.3 -> BLER 1.318e-1 at p=0.09.
    The oracle's counts must be statistically compatible (two-sample binomial, 3.3 sigma)."""
    code = codes["c882"]
    g = oracle.CodeGraph(code)
    G = oracle.Gnn(weights["c882"])
    for cfg, (k_pub, n_pub), B in (
            (dict(num_iters=[64, 16, 16, 16], gnns=[G, G, G], p=0.12, p0=0.05), (396, 5000), 700),
            (dict(num_iters=[64], gnns=[], p=0.09, p0=0.3, factors=[0.8]), (1318, 10000), 600)):
        r = oracle.pipeline(g, seed=42, B=B, skip_inactive=True, **cfg)
        k = int(r["counters"][2])
        p_pool = (k + k_pub) / (B + n_pub)
        sigma = np.sqrt(p_pool * (1 - p_pool) * (1 / B + 1 / n_pub))
        assert abs(k / B - k_pub / n_pub) < 3.3 * sigma, (k, B, k_pub, n_pub)
        lo, hi = _cp_interval(k, B)
        assert lo < k_pub / n_pub < hi
