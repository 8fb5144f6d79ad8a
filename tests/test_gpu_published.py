"""The reference's published plain-BP results (stored outputs of examples/QLDPC.ipynb) reproduced on
the GPU: flagged and block-error counts for six codes under quaternary BP (cell 12), the irregular
rotated surface code (cell 9) and binary BP on a BSC (cell 7).  Two-sample binomial tests at 3.9 sigma
(about 30 comparisons in this file)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compatible(k, n, k_pub, n_pub, z=3.9):
    p_pool = (k + k_pub) / (n + n_pub)
    sigma = np.sqrt(max(p_pool * (1 - p_pool), 1e-12) * (1 / n + 1 / n_pub))
    return abs(k / n - k_pub / n_pub) < z * sigma + 1e-12


def _alist(F, name):
    return F.readAlist(os.path.join(os.path.dirname(F.__file__), "codes_q", name))


def _code(F, name, codes, c1270):
    if name == "GB_n254_k28":
        return F.create_generalized_bicycle_codes(127, [0, 15, 20, 28, 66], [0, 58, 59, 100, 121])
    if name == "GB_n126_k28":
        return F.create_generalized_bicycle_codes(63, [0, 1, 14, 16, 22], [0, 3, 13, 20, 42], name="GB_n126_k28_d8")
    if name == "GB_n48_k6_oc":
        m = _alist(F, "GB_48_6_H_2000.alist")
        return F.css_code(hx=m[:1000], hz=m[1000:], name_prefix="GB")
    if name == "GB_n46_k2_oc":
        m = _alist(F, "GB_46_2_H_800.alist")
        return F.css_code(hx=m[:400], hz=m[400:], name_prefix="GB")
    if name == "GHP_n882_k24":
        return codes["c882"]
    if name == "GHP_n1270_k28":
        return c1270
    if name == "rsurf3":
        return codes["rsurf3"]
    raise KeyError(name)


# (code, num_iter, factor, p0, [(p, flagged_pub, block_pub, frames_pub, frames_here), ...])  -- QLDPC.ipynb cells 9, 12
PUBLISHED_BP4 = [
    ("GB_n254_k28", 64, 0.625, 0.1, [(0.08, 1295, 1295, 10000, 40000), (0.06, 182, 182, 20000, 80000)]),
    ("GB_n126_k28", 64, 0.8, 0.1, [(0.06, 533, 553, 10000, 40000), (0.04, 103, 113, 30000, 120000)]),
    ("GB_n48_k6_oc", 6, 1.0, 0.3, [(0.08, 288, 738, 10000, 40000), (0.05, 37, 124, 10000, 40000)]),
    ("GB_n46_k2_oc", 6, 1.0, 0.3, [(0.08, 183, 366, 10000, 40000), (0.05, 40, 122, 30000, 90000)]),
    ("GHP_n882_k24", 64, 0.8, 0.3, [(0.09, 1318, 1318, 10000, 40000), (0.08, 147, 147, 10000, 40000),
                                    (0.05, 101, 101, 690000, 690000)]),
    ("GHP_n1270_k28", 64, 0.8, 0.3, [(0.09, 1027, 1027, 10000, 40000), (0.08, 130, 130, 20000, 80000),
                                     (0.04, 102, 102, 470000, 470000)]),
    ("rsurf3", 60, 0.8, 0.05, [(0.10, 2747, 3078, 10000, 40000), (0.05, 1349, 1460, 10000, 40000)]),
]


@pytest.mark.parametrize("name,num_iter,factor,p0,points", PUBLISHED_BP4, ids=[r[0] for r in PUBLISHED_BP4])
def test_published_plain_bp4(codes, c1270, name, num_iter, factor, p0, points):
    import fbgnn as F
    code = _code(F, name, codes, c1270)
    dec = F.QLDPCBPDecoder(code=code, num_iter=num_iter, normalization_factor=factor, cn_type="boxplus-phi", stage_one=True)
    for i, (p, fl_pub, blk_pub, n_pub, frames) in enumerate(points):
        model = F.Sandwich_BP_GNN_Evaluation_Model(code, [dec], [], num_layers=1, p0=p0, seed=500 + i)
        c = np.zeros(4, np.int64)
        step = 10000
        for _ in range(frames // step):
            c += model.run(step, p, want_flags=False, want_diff=False, want_counters=True)["counters"]
        assert _compatible(int(c[1]), frames, fl_pub, n_pub), (name, p, "flagged", int(c[1]), frames, fl_pub, n_pub)
        assert _compatible(int(c[2]), frames, blk_pub, n_pub), (name, p, "block", int(c[2]), frames, blk_pub, n_pub)


def test_published_binary_bp_on_bsc(codes):
    """QLDPC.ipynb cell 7: [[882,24]] hx, 64 iterations, p0 = 0.2, logical_pcm = hz_perp; the physical
    rates are arange(0.01, 0.101, 0.01)[::-1] * 2/3."""
    import fbgnn as F
    code = codes["c882"]
    dec = F.LDPCBPDecoder(code.hx, is_syndrome=True, num_iter=64)
    pts = [(0.08 * 2 / 3, 3055, 10000, 40000), (0.07 * 2 / 3, 798, 10000, 40000), (0.06 * 2 / 3, 202, 20000, 80000),
           (0.05 * 2 / 3, 103, 160000, 320000)]
    for i, (p, k_pub, n_pub, frames) in enumerate(pts):
        model = F.BP_BSC_Model(pcm=code.hx, decoder=dec, logical_pcm=code.hz_perp, p0=0.2, seed=900 + i)
        c = np.zeros(4, np.int64)
        for _ in range(frames // 40000):
            c += model.run(40000, p, want_counters=True)["counters"]
        assert _compatible(int(c[2]), frames, k_pub, n_pub), (p, int(c[2]), frames, k_pub, n_pub)
        assert c[1] <= c[2]
