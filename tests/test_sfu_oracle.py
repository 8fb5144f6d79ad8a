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
    ex2, lg2, rcp = T.tables()
    assert ex2.size == T.EX2_COUNT and lg2.size == T.LG2_COUNT and rcp.size == T.RCP_COUNT
    assert np.all(np.diff(ex2) >= 0) and np.all(np.diff(lg2) >= 0) and np.all(np.diff(rcp) <= 0)     # monotone
    assert rcp[0] == 1.0 and rcp[-1] == 0.5 and lg2[-1] == 1.0
    assert np.max(np.abs(rcp * T.rcp_inputs().astype(np.float64) - 1)) < 2.0 ** -22.5
    assert ex2[0x3fc00000 - T.EX2_BASE] == 1.0 and lg2[0x3f800000 - T.LG2_BASE] == 0.0
    # within a few ulp / 2^-22 of the true functions (PTX: ex2.approx 2 ulp, lg2.approx 2^-22 absolute on (0.5, 2))
    w = T.ex2_inputs().astype(np.float64)
    assert np.max(np.abs(ex2 / np.exp2(w) - 1)) < 2.0 ** -21.5
    assert np.max(np.abs(lg2 - np.log2(T.lg2_inputs().astype(np.float64)))) < 2.0 ** -22


def test_sfu_functions_accuracy_and_known_answers(sfu_oracle):
    O = sfu_oracle
    rng = np.random.default_rng(0)
    x = rng.uniform(-87, 88, 500000).astype(np.float32)
    ref = np.exp(x.astype(np.float64))
    ulp = np.abs(O.math_fn("sfu_expf", x) - ref) / np.spacing(ref.astype(np.float32))
    assert ulp.max() < 3.0
    a = np.exp(rng.uniform(-80, 80, 500000)).astype(np.float32)
    err = np.abs(O.math_fn("sfu_logf", a) - np.log(a.astype(np.float64)))
    assert np.all(err < 2.0 * np.spacing(np.abs(np.log(a.astype(np.float64))).astype(np.float32)) + 2.5e-7)
    # the reference's phi known answers hold in this arithmetic too
    phi = lambda v: float(O.math_fn("m_phi4f", np.array([v], np.float32))[0])
    assert phi(8.5e-8) == float(np.float32(16.635532)) and phi(1e-9) == float(np.float32(16.635532))
    assert phi(16.635532) == 0.0 and phi(20.0) == 0.0
    assert phi(1e-3) == pytest.approx(7.6008792, abs=2e-6) and phi(1.0) == pytest.approx(0.7719368, abs=2e-7)
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
def test_sfu_c_oracle_vs_numpy_oracle_teacher_forced(sfu_oracle, codes, c1270, name):
    """One BP4 iteration from the SFU oracle's own state recomputed with numpy's libm: same float32 noise model as
    for the exact arithmetic (tests/test_oracle.py::_tol)."""
    import test_oracle as T
    T.test_c_oracle_vs_numpy_oracle_teacher_forced(sfu_oracle, codes, c1270, name, "boxplus-phi")


def test_sfu_oracle_reproduces_published_error_rates(sfu_oracle, codes, weights):
    import test_oracle as T
    T.test_published_logical_error_rates(sfu_oracle, codes, weights)
