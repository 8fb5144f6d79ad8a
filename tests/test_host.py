"""Host-side logic that needs no GPU: weight files, the C-ABI library and its header, the
Monte-Carlo harness, input validation, and the refusal to compute without a CUDA device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shipped_weights_load_without_tensorflow(weights):
    """gnn.py:755-791; the 4 files hold 12 float32 arrays, 3 923 parameters (Feedback_GNN.ipynb cell 6)."""
    shapes = [(40, 3), (3,), (4, 40), (40,), (40, 20), (20,), (4, 40), (40,), (40, 20), (20,), (43, 40), (40,)]
    for key, w in weights.items():
        assert [a.shape for a in w] == shapes, key
        assert all(a.dtype == np.float32 for a in w) and sum(a.size for a in w) == 3923
        assert all(np.isfinite(a).all() for a in w)


def test_save_load_weights_roundtrip(tmp_path, codes, weights):
    import pickle
    import fbgnn as F
    G = F.Feedback_GNN(code=codes["steane"], num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
    w0 = G.get_weights()
    assert np.all(w0[0] == 0) and np.all(w0[1] == 1) and np.all(w0[3] == 1)      # Keras initialisers
    G.set_weights(weights["c882"])
    path = tmp_path / "w.npy"
    F.save_weights(G, str(path))
    with open(path, "rb") as f:                       # a plain pickle the reference's load_weights can read
        raw = pickle.load(f)
    assert all(isinstance(a, np.ndarray) for a in raw)
    G2 = F.Feedback_GNN(code=codes["steane"], num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=True)
    F.load_weights(G2, str(path))
    assert all(np.array_equal(a, b) for a, b in zip(G2.get_weights(), weights["c882"]))
    with pytest.raises(ValueError):
        G2.set_weights(weights["c882"][:5])
    G3 = F.Feedback_GNN(code=codes["steane"], num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, use_bias=False)
    assert len(G3.get_weights()) == 6
    G4 = F.Feedback_GNN(code=codes["steane"], num_msg_dims=20, num_hidden_units=40, num_mlp_layers=3, use_bias=True)
    assert [w.shape for w in G4.get_weights()] == [(40, 3), (3,)] + [(4, 40), (40,), (40, 40), (40,), (40, 20), (20,)] * 2 + \
        [(43, 40), (40,), (40, 40), (40,)]
    assert [w.shape for w in F.Feedback_GNN(codes["steane"], 20, 40, 1, use_bias=True).get_weights()] == \
        [(43, 3), (3,), (4, 20), (20,), (4, 20), (20,)]


def test_c_abi_library_exports_every_declared_symbol():
    """libfbgnn.so loads without a GPU and exports every function include/fbgnn.h declares."""
    import ctypes
    from fbgnn import _ffi
    hdr = open(os.path.join(ROOT, "include", "fbgnn.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fbgnn_[a-z0-9_]+)\s*\(", body)))
    assert len(declared) >= 30
    lib = _ffi.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fbgnn.h but not exported"
    assert sorted(_ffi.EXPORTED_SYMBOLS) == declared, "ctypes binding and header disagree"
    assert lib.fbgnn_version() == 200
    # torch-free, plain C ABI: the library must not depend on torch / python
    out = os.popen(f"ldd {_ffi._LIB_PATH}").read()
    assert "torch" not in out and "python" not in out


def test_no_cpu_fallback_without_a_gpu(codes):
    """Without a CUDA device every compute path raises; nothing is computed on the host."""
    import fbgnn as F
    from fbgnn import _ffi
    try:
        n = F.device_count()
    except F.FbgnnError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    code = codes["steane"]
    dec = F.QLDPCBPDecoder(code, num_iter=2, cn_type="boxplus-phi", stage_one=True)
    llr = np.zeros((1, 3, 7), np.float32)
    s = np.zeros((3, 1), np.uint8)
    with pytest.raises(F.FbgnnError):
        dec((llr, s, s))
    with pytest.raises(F.FbgnnError):
        F.LDPCBPDecoder(code.hx, is_syndrome=True, num_iter=1)((np.zeros((1, 7), np.float32), s))
    model = F.Sandwich_BP_GNN_Evaluation_Model(code, [dec], [], num_layers=1)
    with pytest.raises(F.FbgnnError):
        model(4, 0.05)


def test_constructor_validation(codes):
    import fbgnn as F
    code = codes["steane"]
    with pytest.raises(ValueError):
        F.QLDPCBPDecoder(code, cn_type="nope")
    with pytest.raises(ValueError):
        F.LDPCBPDecoder(code.hx, cn_type="nope")
    with pytest.raises(AssertionError):
        F.LDPCBPDecoder(code.hx, num_iter=-1)
    with pytest.raises(TypeError):
        F.LDPCBPDecoder([[1, 0]])
    F.QLDPCBPDecoder(code, trainable=True)                   # soft syndromes over the dense hx_perp / hz_perp rows
    F.QLDPCBPDecoder(code, trainable=True, stage_one=True)   # the stage_one return comes first (decoding_q.py:792-793)
    F.QLDPCBPDecoder(code, trainable=True, stage_two=True)
    d = F.QLDPCBPDecoder(code)                              # reference defaults, decoding_q.py:18-22
    assert (d.cn_type, d.num_iter, d.normalization_factor) == ("boxplus", 32, 0.625)
    d2 = F.LDPCBPDecoder(code.hx)                           # decoding.py:264-268
    assert (d2.cn_type, d2.num_iter, d2.normalization_factor) == ("boxplus-phi", 32, 1.0)
    with pytest.raises(ValueError):
        F.Sandwich_BP_GNN_Evaluation_Model(code, [d], [], num_layers=2)
    with pytest.raises(TypeError):
        F.Sandwich_BP_GNN_Evaluation_Model(code, [object()], [], num_layers=1)
    m = F.Sandwich_BP_GNN_Evaluation_Model(code, [d], [], num_layers=1)
    assert m.prior(0.3) == np.float32(4.0430512)           # p0 defaults to 0.05, not p (feedback_gnn.py:265)
    assert F.Sandwich_BP_GNN_Evaluation_Model(code, [d], [], num_layers=1, p0=None).prior(0.1) == \
        np.float32(np.log(27.0))
    assert np.allclose(F.pauli_thresholds(0.12), [0.08, 0.04, 0.12], atol=1e-7)


class _FakeIndicator:
    pass


def test_sim_ber_and_plotber_with_a_fake_model(capsys):
    """sim_ber's qldpc branch: counting, target-error stopping, early stop and the progress table
    (misc.py:557-575, 647-654, 710-716); PlotBER stores flagged and BLER curves at indices 0, 1
    (n1270.py:82 reads _bers[1])."""
    import fbgnn as F
    from fbgnn.feedback_gnn import ErrorIndicator
    calls = []

    def model(batch_size, ebno_db):
        calls.append(float(ebno_db))
        rng = np.random.default_rng(len(calls))
        fl = (rng.random(batch_size) < ebno_db).astype(np.uint8)
        blk = fl | (rng.random(batch_size) < ebno_db / 2).astype(np.uint8)
        return (ErrorIndicator(lambda: fl, 10, lambda: np.repeat(fl[:, None], 10, 1)),
                ErrorIndicator(lambda: blk, 12, lambda: np.repeat(blk[:, None], 12, 1)))

    plot = F.PlotBER()
    ber, bler = plot.simulate(model, ebno_dbs=[0.2, 0.1, 0.0, 0.3], batch_size=500, num_target_block_errors=100,
                              legend="fake", max_mc_iter=50, early_stop=True, add_bler=True, show_fig=False,
                              qldpc=True, forward_keyboard_interrupt=False)
    out = capsys.readouterr().out
    assert "Flagged" in out and "flag errors" in out and "reached target block errors" in out
    assert "Simulation stopped as no error occurred" in out
    assert 0.15 < ber[0] < 0.25 and bler[0] >= ber[0] and bler[2] == 0 and bler[3] == 0   # 4th point not simulated
    assert calls.count(0.0) == 50 and 0.3 not in [round(c, 3) for c in calls]
    assert len(plot._bers) == 2 and plot._is_bler == [False, True] and plot._legends[1] == "fake (BLER)"
    assert np.array_equal(plot._bers[1], bler)
    # dense matrices behave like the reference's tensors
    s_hat, ls_hat = model(64, 0.5)
    assert F.count_block_errors(np.zeros((64, 10)), np.asarray(s_hat)) == F.count_block_errors(None, s_hat)
    assert s_hat.shape == (64, 10)


def test_osd_models_drive_plotber_like_the_reference(capsys):
    """OSD.ipynb cells 2-3 call PlotBER.simulate(osd_model, ..., qldpc=False): the models return
    (zeros_like(ls_hat), ls_hat).  The BSC-based model keeps only per-frame flags on the device, so its indicators
    have no dense form; sim_ber must still count one error per failing frame instead of raising."""
    import fbgnn as F
    from fbgnn.bp_osd import _indicators

    class Flags:                      # stands in for the device flags array
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    def model(batch_size, ebno_db):
        rng = np.random.default_rng(int(ebno_db * 1000))
        return _indicators(Flags(((rng.random(batch_size) < ebno_db).astype(np.uint8)) << 1), 24)

    zeros, ls_hat = model(100, 0.3)
    assert not ls_hat.has_dense() and ls_hat.shape == (100, 24)
    with pytest.raises(F.FbgnnError):
        np.asarray(ls_hat)
    plot = F.PlotBER()
    ber, bler = plot.simulate(model, ebno_dbs=[0.3, 0.1], batch_size=400, num_target_block_errors=50, legend="osd",
                              max_mc_iter=10, early_stop=True, add_bler=True, show_fig=False, qldpc=False,
                              forward_keyboard_interrupt=False)
    assert 0.2 < bler[0] < 0.4 and 0.05 < bler[1] < 0.15
    assert np.allclose(ber, bler / 24)                # one count per failing frame over B x 24 entries
    assert F.count_errors(None, ls_hat) == ls_hat.count_nonzero_rows()
    # with the residual errors kept (BP4_OSD_Model) the dense matrix exists and is what the reference returns
    dense = np.zeros((100, 24), np.int64)
    dense[ls_hat.frame_flags() > 0, 3] = 1
    z2, l2 = _indicators(Flags(ls_hat.frame_flags() << 1), 24, lambda: dense)
    assert l2.has_dense() and np.array_equal(np.asarray(l2), dense) and not np.asarray(z2).any()


def test_weights_unpickler_refuses_foreign_globals(tmp_path):
    import pickle
    import fbgnn as F

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))

    path = tmp_path / "evil.npy"
    with open(path, "wb") as f:
        pickle.dump([Evil()], f)
    with pytest.raises(pickle.UnpicklingError):
        F.read_weights(str(path))


def test_model_run_validates_given_noise_shapes(codes):
    """run(noise=...) must reject arrays that are not [batch_size, n] before they reach the kernels."""
    import fbgnn as F
    src = open(os.path.join(ROOT, "feedback-gnn_b200", "fbgnn", "feedback_gnn.py")).read()
    assert src.count("must have shape [{B},") >= 3          # run(), run_bits() and BP_BSC_Model.run()


def test_shard_ranges_cover_exactly():
    from fbgnn.distributed import shard_range
    for total in (0, 1, 7, 100, 10 ** 8):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == total
            assert all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1


def test_metrics_helpers():
    """compute_bler / compute_ber / count_errors / count_block_errors of sionna/utils/metrics.py:98-223."""
    import fbgnn as F
    s = np.array([[0, 0, 1], [0, 0, 0], [1, 1, 0], [0, 0, 0]])
    z = np.zeros_like(s)
    assert F.compute_bler(z, s) == 0.5 and F.count_block_errors(z, s) == 2
    assert F.count_errors(z, s) == 3 and abs(F.compute_ber(z, s) - 0.25) < 1e-12
