import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def codes():
    """The codes of BASELINE.json's configs plus small / irregular ones (QLDPC.ipynb cells 3, 5)."""
    import fbgnn as F
    c = {}
    c["steane"] = F.css_code(F.hamming_code(3), F.hamming_code(3), name="Steane_n7_k1_d3")
    c["rsurf3"] = F.create_rotated_surface_codes(3)
    c["toric4"] = F.create_checkerboard_toric_codes(4)
    c["gb48"] = F.create_generalized_bicycle_codes(24, [0, 2, 8, 15], [0, 2, 12, 17], name="GB_n48_k6_d8")
    c["c882"] = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
    return c


@pytest.fixture(scope="session")
def c1270():
    import fbgnn as F
    return F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                                [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7],
                                 name="GHP_n1270_k28")


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def weights():
    import fbgnn as F
    d = F.WEIGHTS_DIR
    return {
        "c882": F.read_weights(os.path.join(d, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy")),
        "c882_coarse": F.read_weights(os.path.join(d, "feedback_GNN_n882_k24_wt_4_40_iter_16_16.npy")),
        "c1270": F.read_weights(os.path.join(d, "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy")),
        "c1270_coarse": F.read_weights(os.path.join(d, "feedback_GNN_n1270_k28_wt_10_60_iter_16_16.npy")),
    }
