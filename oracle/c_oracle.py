"""ORACLE (test infrastructure only): ctypes front end of ``oracle/fbgnn_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  It builds the Tanner-graph tables itself (independently of the product's
``fbgnn._graph``) and keeps the reference's tensor layouts at its surface so the parity
tests read like calls into the reference:

* ``bp4(...)``       ~ ``QLDPCBPDecoder(...)((llr_ch, syndrome_x, syndrome_z))``  decoding_q.py:661
* ``bp2(...)``       ~ ``LDPCBPDecoder(..., is_syndrome=True)((llr, syndrome))``   decoding.py:875
* ``gnn(...)``       ~ ``Feedback_GNN(...)((h_vn, logit_hx, logit_hz, sx, sz))``    feedback_gnn.py:161
* ``pipeline(...)``  ~ ``Sandwich_BP_GNN_Evaluation_Model(...)(batch_size, p)``     feedback_gnn.py:293
* ``bsc_pipeline(...)`` ~ ``BP_BSC_Model(...)(batch_size, p)``                      feedback_gnn.py:207
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfbgnn_oracle.so")
_lib = None

CN_TYPES = {"boxplus-phi": 0, "boxplus": 1, "minsum": 2}
ACTS = {"tanh": 0, "relu": 1, None: 2, "linear": 2}
REDUCE = {"mean": 0, "sum": 1, "max": 2, "min": 3}


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "fbgnn_oracle.c")
    hdr = os.path.join(_HERE, "..", "feedback-gnn_b200", "csrc", "fb_math.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


class _Side(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("E", C.c_int32),
                ("vn_ptr", C.c_void_p), ("vn_cn", C.c_void_p), ("cn_ptr", C.c_void_p),
                ("cn_edge", C.c_void_p), ("cn_vn", C.c_void_p)]


class _Rows(C.Structure):
    _fields_ = [("m", C.c_int32), ("ptr", C.c_void_p), ("col", C.c_void_p)]


class _Gnn(C.Structure):
    _fields_ = [("H", C.c_int32), ("M", C.c_int32), ("act", C.c_int32), ("reduce", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("W0", "b0", "W1x", "b1x", "W2x", "b2x",
                                          "W1z", "b1z", "W2z", "b2z", "W3", "b3")] + [("gemm", C.c_int32)]


class _PipeCfg(C.Structure):
    _fields_ = [("num_stages", C.c_int32), ("num_iter", C.c_void_p), ("factor", C.c_void_p),
                ("cn_type", C.c_void_p), ("gnn", C.c_void_p), ("prior", C.c_float),
                ("fixed_weight", C.c_int32), ("osd0", C.c_int32), ("basis_x", C.c_void_p),
                ("pivot_x", C.c_void_p), ("basis_z", C.c_void_p), ("pivot_z", C.c_void_p),
                ("early_stop", C.c_int32), ("skip_inactive", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ arithmetic selection --
_sfu_keep = None


def set_math(mode):
    """"exact" (software exp / log, default) or "sfu" (MUFU.EX2 / MUFU.LG2 through the tables measured on a
    B200, oracle/sfu_tables.py): the arithmetic of every oracle call from now on (fb_math.h)."""
    global _sfu_keep
    L = lib()
    if mode == "sfu":
        if _sfu_keep is None:
            from . import sfu_tables
            _sfu_keep = sfu_tables.tables()
            L.orc_set_sfu_tables(_p(_sfu_keep[0]), _p(_sfu_keep[1]), _p(_sfu_keep[2]), _p(_sfu_keep[3]))
    elif mode != "exact":
        raise ValueError("math mode must be 'exact' or 'sfu'")
    rc = L.orc_set_math(1 if mode == "sfu" else 0)
    assert rc == 0, rc


def get_math():
    return ["exact", "sfu"][lib().orc_get_math()]


class math:
    """``with oracle.math("sfu"): ...``"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_math()
        set_math(self.mode)

    def __exit__(self, *exc):
        set_math(self.prev)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class Side:
    """Edge tables of one parity-check matrix: VN order = sorted by (vn, cn) (the order
    decoding_q.py:63-71 intends), CN order = sorted by (cn, vn) (decoding_q.py:79)."""

    def __init__(self, pcm):
        pcm = np.asarray(pcm)
        self.m, self.n = pcm.shape
        cn, vn = np.nonzero(pcm)                    # row-major: sorted by (cn, vn)
        order = np.lexsort((cn, vn))                # VN order
        self.vn_cn = _i32(cn[order])
        vn_sorted = vn[order]
        self.E = int(len(cn))
        self.vn_ptr = _i32(np.concatenate([[0], np.cumsum(np.bincount(vn_sorted, minlength=self.n))]))
        self.cn_ptr = _i32(np.concatenate([[0], np.cumsum(np.bincount(cn, minlength=self.m))]))
        inv = np.empty(self.E, dtype=np.int64)
        inv[order] = np.arange(self.E)              # row-major edge k sits at VN position inv[k]
        self.cn_edge = _i32(inv)
        self.cn_vn = _i32(vn)
        self.c = _Side(self.n, self.m, self.E, _p(self.vn_ptr), _p(self.vn_cn), _p(self.cn_ptr),
                       _p(self.cn_edge), _p(self.cn_vn))


class Rows:
    def __init__(self, mat):
        mat = np.asarray(mat)
        self.m = mat.shape[0]
        r, c = np.nonzero(mat)
        self.ptr = _i32(np.concatenate([[0], np.cumsum(np.bincount(r, minlength=self.m))]))
        self.col = _i32(c)
        self.c = _Rows(self.m, _p(self.ptr), _p(self.col))


class Gnn:
    """Weights in Keras ``get_weights()`` order."""

    def __init__(self, weights, activation="tanh", reduce_op="mean", use_bias=True, gemm="fma"):
        """gemm="tf32x3": the tensor-core form of the layer (csrc/fbgnn_gnn_tc.cuh), its tcgen05.mma steps emulated
        exactly by csrc/fb_umma.h."""
        w = [_f32(a) for a in weights]
        if use_bias:
            (W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3) = w
        else:
            (W0, W1x, W2x, W1z, W2z, W3) = w
            b0 = b1x = b2x = b1z = b2z = b3 = None
        self.keep = [W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3]
        H, M = W2x.shape
        assert W1x.shape == (4, H) and W3.shape == (2 * M + 3, H) and W0.shape == (H, 3)
        assert gemm in ("fma", "tf32x3")
        if gemm == "tf32x3":
            assert activation == "tanh" and reduce_op in ("mean", "sum") and H % 8 == 0 and (2 * M) % 8 == 0
        self.c = _Gnn(H, M, ACTS[activation], REDUCE[reduce_op], *[_p(a) for a in self.keep], 1 if gemm == "tf32x3" else 0)


class _GnnD(C.Structure):
    _fields_ = [("H", C.c_int32), ("M", C.c_int32), ("L", C.c_int32), ("act", C.c_int32), ("reduce", C.c_int32),
                ("use_bias", C.c_int32), ("w", C.c_void_p)]


def pack_deep_gnn(weights, H, M, L, use_bias=True):
    """Keras get_weights() order of Feedback_GNN with L-layer MLPs -> the packed layer sequence of orc_gnn_deep /
    fbgnn_gnn_create_deep: [W row-major, bias (zeros if absent)] for _llr_inv_embed, msg_x (L), msg_z (L), embed (L-1)."""
    ws = [np.asarray(w, np.float32) for w in weights]
    step = 2 if use_bias else 1
    n_layers = 1 + 2 * L + (L - 1)
    assert len(ws) == n_layers * step, (len(ws), n_layers, step)
    out = []
    for i in range(n_layers):
        W = ws[i * step]
        out.append(W.reshape(-1))
        out.append(ws[i * step + 1].reshape(-1) if use_bias else np.zeros(W.shape[1], np.float32))
    return np.ascontiguousarray(np.concatenate(out), np.float32)


class GnnDeep:
    def __init__(self, weights, H, M, L, activation="tanh", reduce_op="mean", use_bias=True):
        assert max(H, M) <= 128
        self.w = pack_deep_gnn(weights, H, M, L, use_bias)
        self.c = _GnnD(H, M, L, ACTS[activation], REDUCE[reduce_op], int(use_bias), _p(self.w))


def gnn_deep(g, G, h_vn, logit_hx, logit_hz, syndrome_x, syndrome_z):
    h_vn = _f32(h_vn)
    B = h_vn.shape[0]
    lhx, lhz = _f32(logit_hx), _f32(logit_hz)
    sx, sz = _u8(syndrome_x), _u8(syndrome_z)
    out = np.empty((B, g.n, 3), np.float32)
    lib().orc_gnn_deep(C.byref(g.X.c), C.byref(g.Z.c), C.byref(G.c), C.c_int64(B), _p(h_vn), _p(lhx), _p(lhz), _p(sx),
                       _p(sz), _p(out))
    return out


def pauli_thresholds(p):
    """float32 thresholds of Pauli.call as called by the model (feedback_gnn.py:298, pauli.py:100-107)."""
    px, py, pz = np.float32(2 * p / 3), np.float32(p / 3), np.float32(2 * p / 3)
    return np.array([px, px - py, (px + pz) - py], dtype=np.float32)


def prior_llr(p0):
    """tf.math.log(3.*(1.-p0)/p0) (feedback_gnn.py:312): the argument is a Python double,
    cast to float32, then a float32 log (taken here as the correctly rounded one)."""
    x = np.float32(3. * (1. - float(p0)) / float(p0))
    return np.float32(np.log(np.float64(x)))


class CodeGraph:
    def __init__(self, code):
        self.code = code
        self.X = Side(code.hx)
        self.Z = Side(code.hz)
        self.lx = Rows(code.lx)
        self.lz = Rows(code.lz)
        self.n = code.hx.shape[1]


def pauli(seed, first_frame, B, n, p):
    nx = np.empty((B, n), np.uint8)
    nz = np.empty((B, n), np.uint8)
    thr = pauli_thresholds(p)
    lib().orc_pauli(C.c_uint64(seed), C.c_uint64(first_frame), C.c_int64(B), C.c_int(n), _p(thr),
                    _p(nx), _p(nz))
    return nx, nz


def pauli_wt(seed, first_frame, B, n, wt):
    """Pauli(wt=True): exactly ``wt`` erroneous qubits per frame (pauli.py:80-96)."""
    nx = np.empty((B, n), np.uint8)
    nz = np.empty((B, n), np.uint8)
    lib().orc_pauli_wt(C.c_uint64(seed), C.c_uint64(first_frame), C.c_int64(B), C.c_int(n), C.c_int(int(wt)),
                       _p(nx), _p(nz))
    return nx, nz


def bsc(seed, first_frame, B, n, p):
    noise = np.empty((B, n), np.uint8)
    lib().orc_bsc(C.c_uint64(seed), C.c_uint64(first_frame), C.c_int64(B), C.c_int(n),
                  C.c_float(np.float32(p)), _p(noise))
    return noise


def bp4(g, llr, syndrome_x, syndrome_z, num_iter, factor=1.0, cn_type="boxplus-phi",
        rows_x=None, rows_z=None, want_msgs=False, want_iter_logits=False, early_stop=False):
    """llr [B,3,n] f32 (or a float: constant prior); syndromes [m,B] 0/1.
    rows_x / rows_z: matrices whose rows define x_logit / z_logit (default: hz / hx, the
    stage_one choice of decoding_q.py:35-37).  Returns a dict."""
    sx, sz = _u8(syndrome_x), _u8(syndrome_z)
    B = sx.shape[1]
    n = g.n
    rx = Rows(g.code.hz) if rows_x is None else Rows(rows_x)
    rz = Rows(g.code.hx) if rows_z is None else Rows(rows_z)
    if np.isscalar(llr):
        llr_arr, prior = None, float(llr)
    else:
        llr_arr, prior = _f32(llr), 0.0
        assert llr_arr.shape == (B, 3, n)
    out = dict(Lx=np.empty((B, n), np.float32), Ly=np.empty((B, n), np.float32),
               Lz=np.empty((B, n), np.float32), x_hat=np.empty((B, n), np.uint8),
               z_hat=np.empty((B, n), np.uint8), x_logit=np.empty((rx.m, B), np.float32),
               z_logit=np.empty((rz.m, B), np.float32))
    if want_msgs:
        out["msg_x"] = np.empty((B, g.X.E), np.float32)
        out["msg_z"] = np.empty((B, g.Z.E), np.float32)
    if want_iter_logits:                # stage_two / trainable output, decoding_q.py:730,743-746,779-781
        assert rx.m == rz.m
        out["llr_hat"] = np.empty((2 * num_iter + 2, rx.m, B), np.float32)
    if early_stop:                      # opt-in: stop a frame once its decision reproduces the syndrome (not the reference)
        assert num_iter <= 255
        out["iters"] = np.empty(B, np.uint8)
    lib().orc_bp4(C.byref(g.X.c), C.byref(g.Z.c), C.byref(rx.c), C.byref(rz.c),
                  C.c_int(CN_TYPES[cn_type]), C.c_int(num_iter), C.c_float(factor), C.c_int64(B),
                  _p(llr_arr), C.c_float(prior), _p(sx), _p(sz), _p(out["Lx"]), _p(out["Ly"]),
                  _p(out["Lz"]), _p(out["x_hat"]), _p(out["z_hat"]), _p(out["x_logit"]),
                  _p(out["z_logit"]), _p(out.get("msg_x")), _p(out.get("msg_z")), _p(out.get("llr_hat")),
                  _p(out.get("iters")))
    return out


def bp2(pcm_or_side, llr, syndrome, num_iter, factor=1.0, cn_type="boxplus-phi", edge_weights=None, msg_in=None,
        want_msgs=False):
    """llr [B,n] logits; syndrome [m,B] or None.  Returns (soft [B,n], hard [B,n] u8) (+ final c2v messages [B,E], VN
    order, with want_msgs).  edge_weights [E] (VN order): the trainable decoder's weights on the v2c messages;
    msg_in [B,E]: the stateful decoder's incoming message state."""
    S = pcm_or_side if isinstance(pcm_or_side, Side) else Side(pcm_or_side)
    llr = _f32(llr)
    B = llr.shape[0]
    s = None if syndrome is None else _u8(syndrome)
    soft = np.empty((B, S.n), np.float32)
    hard = np.empty((B, S.n), np.uint8)
    ew = None if edge_weights is None else _f32(edge_weights)
    mi = None if msg_in is None else _f32(msg_in)
    mo = np.empty((B, S.E), np.float32) if want_msgs else None
    lib().orc_bp2(C.byref(S.c), C.c_int(CN_TYPES[cn_type]), C.c_int(num_iter), C.c_float(factor),
                  C.c_int64(B), _p(llr), _p(s), _p(soft), _p(hard), _p(ew), _p(mi), _p(mo))
    return (soft, hard, mo) if want_msgs else (soft, hard)


def gnn(g, G, h_vn, logit_hx, logit_hz, syndrome_x, syndrome_z):
    h_vn = _f32(h_vn)
    B = h_vn.shape[0]
    lhx, lhz = _f32(logit_hx), _f32(logit_hz)
    sx, sz = _u8(syndrome_x), _u8(syndrome_z)
    out = np.empty((B, g.n, 3), np.float32)
    lib().orc_gnn(C.byref(g.X.c), C.byref(g.Z.c), C.byref(G.c), C.c_int64(B), _p(h_vn), _p(lhx),
                  _p(lhz), _p(sx), _p(sz), _p(out))
    return out


def pipeline(g, num_iters, gnns, p, p0=0.05, factors=None, cn_types=None, seed=0, first_frame=0,
             B=1, noise=None, skip_inactive=False, want_diff=False, wt=0, osd0=False, early_stop=False):
    """Sandwich model: decoders[i] has num_iters[i] iterations; gnns[i] is feedbacks[i].
    Returns dict(flags [B] u8, counters [4] i64, x_diff, z_diff)."""
    S = len(num_iters)
    assert len(gnns) == S - 1
    ni = _i32(num_iters)
    fa = _f32([1.0] * S if factors is None else factors)
    ct = _i32([CN_TYPES[c] for c in (cn_types or ["boxplus-phi"] * S)])
    garr = (C.c_void_p * max(S - 1, 1))(*[C.addressof(G.c) for G in gnns])
    code = g.code
    bx, bz = Rows(np.asarray(code.hx)[code.pivot_hx]), Rows(np.asarray(code.hz)[code.pivot_hz])
    pvx, pvz = _i32(code.pivot_hx), _i32(code.pivot_hz)
    cfg = _PipeCfg(S, _p(ni), _p(fa), _p(ct), C.cast(garr, C.c_void_p),
                   C.c_float(prior_llr(p if p0 is None else p0)), int(wt), int(osd0),
                   C.cast(C.pointer(bx.c), C.c_void_p), _p(pvx), C.cast(C.pointer(bz.c), C.c_void_p), _p(pvz),
                   int(early_stop), int(skip_inactive))
    thr = pauli_thresholds(p)
    flags = np.empty(B, np.uint8)
    counters = np.zeros(4, np.int64)
    nx = nz = None
    if noise is not None:
        nx, nz = _u8(noise[0]), _u8(noise[1])
        assert nx.shape == (B, g.n)
    xd = np.empty((B, g.n), np.uint8) if want_diff else None
    zd = np.empty((B, g.n), np.uint8) if want_diff else None
    lib().orc_pipeline(C.byref(g.X.c), C.byref(g.Z.c), C.byref(g.lx.c), C.byref(g.lz.c),
                       C.byref(cfg), _p(thr), C.c_uint64(seed), C.c_uint64(first_frame),
                       C.c_int64(B), _p(nx), _p(nz), _p(flags), _p(counters), _p(xd), _p(zd))
    return dict(flags=flags, counters=counters, x_diff=xd, z_diff=zd)


def bsc_pipeline(pcm, logical_pcm, num_iter, p, p0=None, factor=1.0, cn_type="boxplus-phi",
                 seed=0, first_frame=0, B=1, noise=None, osd_basis=None, osd_pivot=None):
    S = Side(pcm)
    L = None if logical_pcm is None else Rows(logical_pcm)
    p0 = np.float32(p if p0 is None else p0)
    llr_const = np.float32(-np.log((np.float32(1.0) - p0) / p0))
    flags = np.empty(B, np.uint8)
    counters = np.zeros(4, np.int64)
    nz = None if noise is None else _u8(noise)
    basis = None if osd_basis is None else Rows(osd_basis)          # kept alive across the call
    pivot = None if osd_pivot is None else _i32(osd_pivot)
    lib().orc_bsc_pipeline(C.byref(S.c), None if L is None else C.byref(L.c),
                           C.c_int(CN_TYPES[cn_type]), C.c_int(num_iter), C.c_float(factor),
                           C.c_float(llr_const), C.c_float(np.float32(p)), C.c_uint64(seed),
                           C.c_uint64(first_frame), C.c_int64(B), _p(nz), _p(flags), _p(counters),
                           None if basis is None else C.byref(basis.c), _p(pivot))
    return dict(flags=flags, counters=counters)


def osd0(basis, llr, s):
    """OSD0_Decoder.call (bp_osd.py:51-77): basis [rank,n] full-rank rows, llr [B,n], s [rank,B] -> e_hat [B,n]."""
    R = Rows(basis)
    llr = _f32(llr)
    B, n = llr.shape
    s = _u8(s)
    assert s.shape == (R.m, B)
    out = np.empty((B, n), np.uint8)
    lib().orc_osd0(C.byref(R.c), C.c_int(n), C.c_int64(B), _p(llr), _p(s), _p(out))
    return out


class _Gbp(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("d", "H", "M", "act", "reduce", "use_bias", "num_iter")] + \
               [("Winv", C.c_void_p), ("binv", C.c_void_p)] + \
               [(k, C.c_void_p * 4) for k in ("cmx", "cmz", "cex", "cez", "vmx", "vmz", "ve")]


GBP_KEYS = ("cmx", "cmz", "cex", "cez", "vmx", "vmz", "ve")


def gnn_bp4_random_weights(d=20, M=20, H=40, seed=4, use_bias=True):
    """Glorot-uniform kernels, ones biases, non-zero llr_inv_embed kernel (the
    reference ships no weights for GNN_BP4)."""
    rng = np.random.default_rng(seed)

    def glorot(a, b):
        lim = np.sqrt(6.0 / (a + b))
        return rng.uniform(-lim, lim, (a, b)).astype(np.float32)

    def mlp(k_in, k_out):
        return (glorot(k_in, H), np.ones(H, np.float32) if use_bias else None, glorot(H, k_out),
                np.ones(k_out, np.float32) if use_bias else None)

    W = {"Winv": glorot(d, 3), "binv": np.ones(3, np.float32) if use_bias else None}
    for k in ("cmx", "cmz", "vmx", "vmz"):
        W[k] = mlp(2 * d, M)
    for k in ("cex", "cez"):
        W[k] = mlp(M + d + 1, d)
    W["ve"] = mlp(2 * M + d, d)
    return W


def gnn_bp4(g, W, synd_x, synd_z, num_iter, activation="tanh", reduce_op="mean"):
    """GNN_BP4(...)((syndrome_x [B,m_x], syndrome_z [B,m_z])) -> dict(x_logit [it, m_z+k, B],
    z_logit [it, m_x+k, B], x_hat [n,B], z_hat [n,B]) (gnn.py:379-420)."""
    sx, sz = _u8(synd_x), _u8(synd_z)
    B = sx.shape[0]
    d, H = W["cmx"][0].shape[0] // 2, W["cmx"][0].shape[1]
    M = W["cmx"][2].shape[1]
    G = _Gbp(d, H, M, ACTS[activation], REDUCE[reduce_op], int(W["binv"] is not None), num_iter)
    keep = []

    def ptr(a):
        if a is None:
            return None
        a = _f32(a)
        keep.append(a)
        return a.ctypes.data

    G.Winv, G.binv = ptr(W["Winv"]), ptr(W["binv"])
    for k in GBP_KEYS:
        setattr(G, k, (C.c_void_p * 4)(*[ptr(a) for a in W[k]]))
    rx, rz = g.Z.m + g.lz.m, g.X.m + g.lx.m
    out = dict(x_logit=np.empty((num_iter, rx, B), np.float32), z_logit=np.empty((num_iter, rz, B), np.float32),
               x_hat=np.empty((g.n, B), np.uint8), z_hat=np.empty((g.n, B), np.uint8))
    lib().orc_gnn_bp4(C.byref(G), C.byref(g.X.c), C.byref(g.Z.c), C.byref(g.lx.c), C.byref(g.lz.c), C.c_int64(B),
                      _p(sx), _p(sz), _p(out["x_logit"]), _p(out["z_logit"]), _p(out["x_hat"]), _p(out["z_hat"]))
    return out


def philox(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def math_fn(name, *args):
    args = [_f32(a) for a in args]
    y = np.empty_like(args[0])
    getattr(lib(), "orc_" + name)(*[_p(a) for a in args], _p(y), C.c_int64(y.size))
    return y


def num_threads():
    return lib().orc_num_threads()
