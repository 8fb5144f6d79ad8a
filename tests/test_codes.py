"""Code construction (host side): known answers from examples/QLDPC.ipynb cells 3 and 5, and the
bit-packed GF(2) elimination against the literal restatement of the reference's loops."""
import os

import numpy as np
import pytest


def test_steane_known_answer():
    """QLDPC.ipynb cell 3 prints Hx, Hz, Lx, Lz of css_code(hamming_code(3), hamming_code(3))."""
    import fbgnn as F
    c = F.css_code(F.hamming_code(3), F.hamming_code(3), name="Steane_n7_k1_d3")
    H = np.array([[0, 0, 0, 1, 1, 1, 1], [0, 1, 1, 0, 0, 1, 1], [1, 0, 1, 0, 1, 0, 1]])
    assert np.array_equal(c.hx, H) and np.array_equal(c.hz, H)
    assert np.array_equal(c.lx, [[1, 1, 1, 0, 0, 0, 0]]) and np.array_equal(c.lz, [[1, 1, 1, 0, 0, 0, 0]])
    assert (c.N, c.K, c.name) == (7, 1, "Steane_n7_k1_d3")


def _alist(name):
    import fbgnn as F
    return F.readAlist(os.path.join(os.path.dirname(F.__file__), "codes_q", name))


CODES = {
    # name: (constructor, expected N, expected K)  -- the [[n,k]] in the notebook's variable names
    "Surface_n1201_k1": (lambda F: F.create_surface_codes(25), 1201, 1),
    "Rotated_Surface_n121_k1": (lambda F: F.create_rotated_surface_codes(11), 121, 1),
    "Toric_n100_k2": (lambda F: F.create_checkerboard_toric_codes(10), 100, 2),
    "GB_n254_k28": (lambda F: F.create_generalized_bicycle_codes(127, [0, 15, 20, 28, 66], [0, 58, 59, 100, 121]), 254, 28),
    "GB_n126_k28": (lambda F: F.create_generalized_bicycle_codes(63, [0, 1, 14, 16, 22], [0, 3, 13, 20, 42]), 126, 28),
    "GB_n48_k6": (lambda F: F.create_generalized_bicycle_codes(24, [0, 2, 8, 15], [0, 2, 12, 17]), 48, 6),
    "GB_n48_k6_oc": (lambda F: (lambda m: F.css_code(hx=m[:1000], hz=m[1000:], name_prefix="GB"))(_alist("GB_48_6_H_2000.alist")), 48, 6),
    "GB_n46_k2": (lambda F: F.create_generalized_bicycle_codes(23, [0, 5, 8, 12], [0, 1, 5, 7]), 46, 2),
    "GB_n46_k2_oc": (lambda F: (lambda m: F.css_code(hx=m[:400], hz=m[400:], name_prefix="GB"))(_alist("GB_46_2_H_800.alist")), 46, 2),
    "GB_n180_k10": (lambda F: F.create_generalized_bicycle_codes(90, [0, 28, 80, 89], [0, 2, 21, 25]), 180, 10),
    "GB_n900_k50": (lambda F: F.create_generalized_bicycle_codes(450, [0, 97, 372, 425], [0, 50, 265, 390]), 900, 50),
    "HP_n1922_k50": (lambda F: (lambda h: F.hypergraph_product(h, h))(F.create_circulant_matrix(31, [0, 2, 5])), 1922, 50),
    "GHP_n882_k24": (lambda F: F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6]), 882, 24),
    "GHP_n882_k48": (lambda F: F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 0, 27, 18, 0]), [0, 1, 6]), 882, 48),
    "IBM_n72_k12": (lambda F: F.create_bivariate_QC_codes(6, 6, [3], [1, 2], [1, 2], [3]), 72, 12),
    "IBM_n90_k8": (lambda F: F.create_bivariate_QC_codes(15, 3, [9], [1, 2], [2, 7], [0]), 90, 8),
    "IBM_n108_k8": (lambda F: F.create_bivariate_QC_codes(9, 6, [3], [1, 2], [1, 2], [3]), 108, 8),
    "IBM_n144_k12": (lambda F: F.create_bivariate_QC_codes(12, 6, [3], [1, 2], [1, 2], [3]), 144, 12),
    "IBM_n288_k12": (lambda F: F.create_bivariate_QC_codes(12, 12, [3], [2, 7], [1, 2], [3]), 288, 12),
    "IBM_n360_k12": (lambda F: F.create_bivariate_QC_codes(30, 6, [9], [1, 2], [25, 26], [3]), 360, 12),
    "IBM_n756_k16": (lambda F: F.create_bivariate_QC_codes(21, 18, [3], [10, 17], [3, 19], [5]), 756, 16),
}


@pytest.mark.parametrize("name", sorted(CODES))
def test_code_parameters(name):
    import fbgnn as F
    make, n, k = CODES[name]
    c = make(F)
    assert (c.N, c.K) == (n, k)
    # CSS condition and the defining properties of the derived matrices
    assert not np.any(c.hx @ c.hz.T % 2)
    assert not np.any(c.hx @ c.hx_perp.T % 2) and not np.any(c.hz @ c.hz_perp.T % 2)
    assert c.hx_perp.shape[0] == n - c.rank_hx and c.hz_perp.shape[0] == n - c.rank_hz
    assert c.lx.shape == (k, n) and c.lz.shape == (k, n)
    assert not np.any(c.hz @ c.lx.T % 2) and not np.any(c.hx @ c.lz.T % 2)
    assert F.rank(np.vstack([c.hx_basis, c.lx])) == c.rank_hx + k      # logicals independent of stabilisers
    assert F.rank(c.lx @ c.lz.T % 2) == k                               # and pairwise non-degenerate


def test_c1270_parameters(c1270):
    """n1270.py:37; SURVEY.md section 8 sizes: 635x1270, (3,6)-regular, rank 621, perps 649x1270."""
    c = c1270
    assert (c.N, c.K, c.name) == (1270, 28, "GHP_n1270_k28")
    assert c.hx.shape == (635, 1270) and c.rank_hx == 621 and c.rank_hz == 621
    assert set(c.hx.sum(0)) == {3} and set(c.hx.sum(1)) == {6} and set(c.hz.sum(0)) == {3} and set(c.hz.sum(1)) == {6}
    assert c.hx_perp.shape == (649, 1270) and (int(c.hx_perp.sum()), int(c.hz_perp.sum())) == (200862, 73404)
    assert (c.L, c.Q) == (3, 6)


@pytest.mark.parametrize("name", ["GB_n48_k6", "GHP_n882_k24", "Toric_n100_k2", "GB_n46_k2_oc"])
def test_gf2_matches_literal_reference_restatement(name):
    """Same kernel basis, pivots and logical operators as the reference's row loops produce."""
    import fbgnn as F
    from oracle.codes_ref import css_ref, row_echelon_ref
    c = CODES[name][0](F)
    r = css_ref(c.hx, c.hz)
    for key in ("hx_perp", "hz_perp", "hx_basis", "hz_basis", "lx", "lz"):
        assert np.array_equal(r[key], getattr(c, key)), key
    assert r["pivot_hx"] == c.pivot_hx and r["pivot_hz"] == c.pivot_hz and r["K"] == c.K
    rng = np.random.default_rng(0)
    m = (rng.random((17, 29)) < 0.3).astype(int)
    for reduced in (False, True):
        a, b = F.row_echelon(m, reduced=reduced), row_echelon_ref(m, reduced=reduced)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3] == b[3]
        assert np.array_equal(a[2] @ m % 2, a[0])
    sq = np.triu(np.ones((9, 9), int))
    assert np.array_equal(F.inverse(sq) @ sq % 2, np.eye(9, dtype=int))
    with pytest.raises(ValueError):
        F.inverse(np.zeros((3, 3), int))


def test_small_helpers():
    import fbgnn as F
    assert F.int2bin(5, 4) == [0, 1, 0, 1] and F.int2bin(12, 3) == [1, 0, 0] and F.int2bin(3, 0) == []
    assert F.int_mod_2(np.array([0, 1, 2, 3, -1])).tolist() == [0, 1, 0, 1, 1]
    assert F.create_circulant_matrix(4, [1]).tolist() == [[0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]
    A = F.create_cyclic_permuting_matrix(3, [5, 7])
    assert A.tolist() == [[5, -1, 7], [7, 5, -1], [-1, 7, 5]]
    assert F.rep_code(3).tolist() == [[1, 1, 0], [0, 1, 1]]
    with pytest.raises(AssertionError):
        F.create_rotated_surface_codes(4)
    with pytest.raises(AssertionError):
        F.css_code(np.zeros((2, 3), int), np.zeros((2, 4), int))
