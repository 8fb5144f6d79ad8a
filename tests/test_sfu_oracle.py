"""CPU side of the SFU arithmetic: the committed MUFU tables, the accuracy of the functions built on them, the
reference's known answers in this arithmetic, and the cross-check of the C oracle (SFU) against the independent
numpy oracle (numpy's libm) one teacher-forced step at a time."""
import numpy as np
import pytest


@pytest.fixture()
def sfu_oracle(oracle):
    oracle.set_math("sfu")
    yield oracle
    oracle.set_math("exact")


def test_tables_load_and_are_sane():
    from oracle import sfu_tables as T
    assert T.available()
    ex2, lg2, lg2b, rcp = T.tables()
    assert (ex2.size, lg2.size, lg2b.size, rcp.size) == (T.EX2_COUNT, T.LG2_COUNT, T.LG2B_COUNT, T.RCP_COUNT)
    assert np.all(np.diff(ex2) >= 0) and np.all(np.diff(lg2) >= 0) and np.all(np.diff(rcp) <= 0)     # monotone
    assert np.all(np.diff(lg2b) >= 0) and lg2b.max() < 0.0
    assert rcp[0] == 1.0 and rcp[-1] == 0.5 and lg2[-1] == 1.0
    assert np.max(np.abs(rcp * T.rcp_inputs().astype(np.float64) - 1)) < 2.0 ** -22.5
    assert ex2[0x3f800000 - T.EX2_BASE] == 2.0 and lg2[0x3f800000 - T.LG2_BASE] == 0.0
    assert abs(float(lg2b[2]) + 23.0) < 4e-6                     # lg2(2^-23): what bounds phi by 24 ln 2
    # within a few ulp / 2^-22 of the true functions (PTX: ex2.approx 2 ulp, lg2.approx 2^-22 absolute on (0.5, 2),
    # 2^-22 relative elsewhere)
    u = T.ex2_inputs().astype(np.float64)
    assert np.max(np.abs(ex2 / np.exp2(u) - 1)) < 2.0 ** -21.5
    assert np.max(np.abs(lg2 - np.log2(T.lg2_inputs().astype(np.float64)))) < 2.0 ** -22
    tb = np.log2(T.lg2b_inputs().astype(np.float64))[1:]
    assert np.max(np.abs(lg2b[1:] - tb) / np.maximum(np.abs(tb), 1.0)) < 2.0 ** -21.5


def test_sfu_functions_accuracy_and_known_answers(sfu_oracle):
    O = sfu_oracle
    rng = np.random.default_rng(0)
    # exp: the exponent x log2(e) is formed in one FMA without a two-constant reduction, so the error grows with |x|:
    # a few ulp over the range the decoders use (|x| <= 20), 2e-6 relative at the ends of the float32 range
    x = rng.uniform(-20, 20, 500000).astype(np.float32)
    ref = np.exp(x.astype(np.float64))
    ulp = np.abs(O.math_fn("sfu_expf", x) - ref) / np.spacing(ref.astype(np.float32))
    assert ulp.max() < 6.0
    x = rng.uniform(-86, 88, 500000).astype(np.float32)
    ref = np.exp(x.astype(np.float64))
    assert np.max(np.abs(O.math_fn("sfu_expf", x) / ref - 1)) < 2.5e-6
    a = np.exp(rng.uniform(-80, 80, 500000)).astype(np.float32)
    err = np.abs(O.math_fn("sfu_logf", a) - np.log(a.astype(np.float64)))
    assert np.all(err < 2.0 * np.spacing(np.abs(np.log(a.astype(np.float64))).astype(np.float32)) + 2.5e-7)
    # the reference's phi known answers hold in this arithmetic too
    phi = lambda v: float(O.math_fn("m_phi4f", np.array([v], np.float32))[0])
    assert phi(8.5e-8) == float(np.float32(16.635532)) and phi(1e-9) == float(np.float32(16.635532))
    assert phi(16.635532) == 0.0 and phi(20.0) == 0.0
    # phi(1e-3): 1 - exp(-x) = 1e-3 carries the float32 rounding of exp (6e-8), i.e. 6e-5 relative -- as the reference's
    # exp(x) - 1 does (its own known answer 7.6008792 is 2.3e-5 off the true 7.6009025)
    assert phi(1e-3) == pytest.approx(7.6009025, abs=1.2e-4) and phi(1.0) == pytest.approx(0.7719368, abs=3e-7)
    xs = np.exp(rng.uniform(np.log(8.5e-8), np.log(16.635532), 500000)).astype(np.float32)
    ps = O.math_fn("m_phi4f", xs)
    assert ps.min() >= 0.0 and ps.max() <= np.float32(16.635532) + np.float32(0.7)
    t = rng.uniform(-12, 12, 300000).astype(np.float32)
    th = O.math_fn("m_tanhf", t)
    assert np.max(np.abs(th - np.tanh(t.astype(np.float64)))) < 4e-7 and np.all(np.abs(th) <= 1.0)
    assert O.math_fn("m_tanhf", np.array([0.0, 30.0, -30.0], np.float32)).tolist() == [0.0, 1.0, -1.0]
    sp = O.math_fn("m_softplusf", np.array([-100, -20, -1, 0, 1, 13.9, 14, 50], np.float32))
    assert np.allclose(sp, np.logaddexp(0, np.array([-100, -20, -1, 0, 1, 13.9, 14, 50], np.float64)), rtol=3e-7, atol=2e-7)


def test_first_stage_marginal_extrema_kat_sfu(sfu_oracle, c1270):
    """examples/n1270.ipynb cell 12 in SFU arithmetic: the saturated marginals are the same known answers."""
    import test_oracle as T
    T.test_first_stage_marginal_extrema_kat(sfu_oracle, c1270)


@pytest.mark.parametrize("name", ["c882", "c1270"])
def test_sfu_c_oracle_vs_numpy_oracle_teacher_forced(sfu_oracle, codes, c1270, name, monkeypatch):
    """One BP4 iteration from the SFU oracle's own state recomputed with numpy's libm: the float32 noise model of the
    exact arithmetic (tests/test_oracle.py::_tol) with the cancellation term doubled -- the SFU form rounds
    softplus(x) and log(expm1(x)) at the same magnitude as the reference but through different operations, so the
    two evaluations differ by up to one quantum per term more often than two polynomial evaluations do."""
    import test_oracle as T
    monkeypatch.setattr(T, "_CANCEL", 8e-6)
    T.test_c_oracle_vs_numpy_oracle_teacher_forced(sfu_oracle, codes, c1270, name, "boxplus-phi")


def test_sfu_oracle_reproduces_published_error_rates(sfu_oracle, codes, weights):
    import test_oracle as T
    T.test_published_logical_error_rates(sfu_oracle, codes, weights)
