"""ORACLE (test infrastructure only): numpy float32 restatement of the reference's TF graph.

Independent of ``fb_math.h``: every transcendental here is numpy's own float32
``exp/log/log1p/tanh/arctanh``, the dense layers go through ``@`` (BLAS), the reductions
through ``np.add.reduceat`` -- i.e. a different libm and different summation orders than the
C oracle and the CUDA kernels.  It exists to cross-check the C oracle (and with it the
arithmetic specification) at a tolerance, teacher-forced one step at a time, and to pin the
notebook known answers.  Layouts are the reference's (batch last inside the decoder).

Follows: decoding_q.py:227-275 (_vn_update), 365-431 (_phi, _cn_update_phi), 313-363
(_cn_update_tanh), 539-644 (_cn_update_minsum), 433-471 (cal_logit), 661-797 (call);
decoding.py:511-535, 625-690, 875-1048; feedback_gnn.py:161-188, 293-361; gnn.py:31-69;
pauli.py:98-108.
"""
import numpy as np

F = np.float32
THR = F(13.942385)          # -(log(eps_f32) + 2): Eigen's softplus threshold


def softplus(x):
    """tf.math.softplus (Eigen functor)."""
    x = x.astype(F)
    with np.errstate(over="ignore", under="ignore"):
        e = np.exp(np.minimum(x, F(20)))
        r = np.where(x > THR, x, np.where(x < -THR, e, np.log1p(e)))
    return r.astype(F)


def logsumexp2(a, b):
    """tf.reduce_logsumexp over stack([a, b], -1): max-shifted, log (not log1p)."""
    mx = np.maximum(a, b)
    with np.errstate(under="ignore"):
        s = np.exp(a - mx) + np.exp(b - mx)
    return (np.log(s) + mx).astype(F)


def phi4(x):
    """decoding_q.py:365-373."""
    x = np.clip(x.astype(F), F(8.5e-8), F(16.635532))
    return (softplus(x) - np.log(np.exp(x) - F(1))).astype(F)


def phi2(x):
    """decoding.py:625-633."""
    x = np.clip(x.astype(F), F(8.5e-8), F(16.635532))
    e = np.exp(x)
    return (np.log(e + F(1)) - np.log(e - F(1))).astype(F)


class Side:
    """Edge tables of one pcm; VN order sorted by (vn, cn), CN order sorted by (cn, vn)."""

    def __init__(self, pcm):
        pcm = np.asarray(pcm)
        self.m, self.n = pcm.shape
        cn, vn = np.nonzero(pcm)
        order = np.lexsort((cn, vn))
        self.vn_of_edge = vn[order]                 # VN order
        self.cn_of_edge = cn[order]
        self.E = len(cn)
        self.ind_cn = np.lexsort((self.vn_of_edge, self.cn_of_edge))   # VN order -> CN order
        self.ind_cn_inv = np.argsort(self.ind_cn)
        self.cn_of_edge_c = self.cn_of_edge[self.ind_cn]
        self.vn_starts = np.searchsorted(self.vn_of_edge, np.arange(self.n))
        self.cn_starts = np.searchsorted(self.cn_of_edge_c, np.arange(self.m))
        self.vn_deg = np.bincount(self.vn_of_edge, minlength=self.n)
        self.cn_deg = np.bincount(self.cn_of_edge_c, minlength=self.m)

    def vn_sum(self, msg):            # [E,B] -> [n,B]
        out = np.add.reduceat(msg, self.vn_starts, axis=0)
        out[self.vn_deg == 0] = 0
        return out.astype(F)

    def cn_reduce(self, ufunc, msg_c):   # CN-ordered [E,B] -> [m,B]
        return ufunc.reduceat(msg_c, self.cn_starts, axis=0)


def cn_update(S, v2c, synd_sign, cn_type="boxplus-phi", phi=phi4):
    """v2c [E,B] in VN order, synd_sign [m,B] in {+1,-1} -> c2v [E,B] in VN order (no factor)."""
    m = v2c[S.ind_cn]
    rows = S.cn_of_edge_c
    if cn_type == "boxplus-phi":
        sign = np.where(m < 0, F(-1), F(1))
        node = S.cn_reduce(np.multiply, sign) * synd_sign
        a = phi(np.abs(m))
        T = S.cn_reduce(np.add, a).astype(F)
        out = sign * node[rows] * phi(T[rows] - a)
    elif cn_type == "boxplus":
        t = np.tanh((m / F(2)).astype(F))
        t = np.where(t == 0, F(1e-12), t)
        P = S.cn_reduce(np.multiply, t) * synd_sign
        out = (F(1) / t) * P[rows]
        out = np.where(np.abs(out) < 1e-7, F(0), out)
        out = np.clip(out, -F(1 - 1e-7), F(1 - 1e-7))
        out = F(2) * np.arctanh(out)
    elif cn_type == "minsum":
        LARGE = F(10000.)
        m = np.clip(m, F(-20), F(20))
        sign = np.where(m < 0, F(-1), F(1))
        node = S.cn_reduce(np.multiply, sign) * synd_sign
        a = np.abs(m)
        mn = S.cn_reduce(np.minimum, a)
        d = a - mn[rows]
        d = np.where(d == 0, LARGE, d)
        mn2 = S.cn_reduce(np.minimum, d) + mn
        node_sum = S.cn_reduce(np.add, d) - (2 * LARGE - 1)
        dm = F(0.5) * (1 - np.sign(node_sum))
        mne = (1 - dm) * mn + dm * mn2
        out = sign * node[rows] * np.where(d == LARGE, mne[rows], mn[rows])
    else:
        raise ValueError(cn_type)
    return out.astype(F)[S.ind_cn_inv]


def bp4_marginals(X, Z, mx, mz, llr):
    Sx, Sz = X.vn_sum(mx), Z.vn_sum(mz)
    Ly = ((Sz + Sx) + llr[1]).astype(F)
    return (Sz + llr[0]).astype(F), Ly, (Sx + llr[2]).astype(F)


def bp4_iteration(X, Z, mx, mz, llr, ssx, ssz, factor=1.0, cn_type="boxplus-phi"):
    """One flooding iteration. mx,mz: c2v [E,B]; llr: (llrx,llry,llrz) each [n,B];
    ssx/ssz: syndrome signs [m,B].  Returns new c2v messages (and the v2c messages)."""
    Lx, Ly, Lz = bp4_marginals(X, Z, mx, mz, llr)
    vx, vz = X.vn_of_edge, Z.vn_of_edge
    v2c_x = softplus(-Lx)[vx] - logsumexp2(-(Lz[vx] - mx), -(Ly[vx] - mx))
    v2c_z = softplus(-Lz)[vz] - logsumexp2(-(Lx[vz] - mz), -(Ly[vz] - mz))
    nmx = cn_update(X, v2c_x.astype(F), ssx, cn_type) * F(factor)
    nmz = cn_update(Z, v2c_z.astype(F), ssz, cn_type) * F(factor)
    return nmx.astype(F), nmz.astype(F), v2c_x.astype(F), v2c_z.astype(F)


def soft_syndrome(rows_mat_side, l):
    """_cn_update_phi_loss over the rows of a matrix given as a Side; l [n,B] -> [m,B]."""
    S = rows_mat_side
    vals = l[S.vn_of_edge[S.ind_cn]]
    sign = np.where(vals < 0, F(-1), F(1))
    node = S.cn_reduce(np.multiply, sign)
    T = S.cn_reduce(np.add, phi4(np.abs(vals))).astype(F)
    return (node * phi4(T)).astype(F)


def cal_logit(rows_x, rows_z, Lx, Ly, Lz):
    llr_z = softplus(-Lx) - logsumexp2(-Lz, -Ly)
    llr_x = softplus(-Lz) - logsumexp2(-Lx, -Ly)
    return soft_syndrome(rows_x, llr_x.astype(F)), soft_syndrome(rows_z, llr_z.astype(F))


def decide(Lx, Ly, Lz):
    d = np.argmin(np.stack([np.zeros_like(Lx), Lx, Lz, Ly], 0), axis=0)
    return (d & 1).astype(np.uint8), (d >> 1).astype(np.uint8)


def bp4(X, Z, llr, synd_x, synd_z, num_iter, factor=1.0, cn_type="boxplus-phi",
        rows_x=None, rows_z=None, init=None):
    """llr [B,3,n]; synd [m,B].  Returns dict with the reference's stage_one outputs
    (batch first for Lx,Ly,Lz,x_hat,z_hat; [m,B] for the logits) and final messages."""
    llr = np.asarray(llr, F).transpose(1, 2, 0)
    ssx = (1 - 2 * np.asarray(synd_x).astype(np.int32)).astype(F)
    ssz = (1 - 2 * np.asarray(synd_z).astype(np.int32)).astype(F)
    B = ssx.shape[1]
    mx = np.zeros((X.E, B), F) if init is None else init[0].astype(F)
    mz = np.zeros((Z.E, B), F) if init is None else init[1].astype(F)
    for _ in range(num_iter):
        mx, mz, _, _ = bp4_iteration(X, Z, mx, mz, llr, ssx, ssz, factor, cn_type)
    Lx, Ly, Lz = bp4_marginals(X, Z, mx, mz, llr)
    xl, zl = cal_logit(rows_x if rows_x is not None else Z, rows_z if rows_z is not None else X,
                       Lx, Ly, Lz)
    xh, zh = decide(Lx, Ly, Lz)
    return dict(Lx=Lx.T.copy(), Ly=Ly.T.copy(), Lz=Lz.T.copy(), x_hat=xh.T.copy(),
                z_hat=zh.T.copy(), x_logit=xl, z_logit=zl, msg_x=mx, msg_z=mz)


def bp2_iteration(S, msg, llr, ss, factor=1.0, cn_type="boxplus-phi"):
    x = (S.vn_sum(msg) + llr).astype(F)
    v2c = (x[S.vn_of_edge] - msg).astype(F)
    return (cn_update(S, v2c, ss, cn_type, phi=phi2) * F(factor)).astype(F)


def bp2(S, logits, synd, num_iter, factor=1.0, cn_type="boxplus-phi"):
    """logits [B,n]; synd [m,B] or None. Returns (soft [B,n] logits, hard [B,n])."""
    llr = -np.clip(np.asarray(logits, F), F(-20), F(20)).T
    B = llr.shape[1]
    ss = np.ones((S.m, B), F) if synd is None else (1 - 2 * np.asarray(synd).astype(np.int32)).astype(F)
    msg = np.zeros((S.E, B), F)
    for _ in range(num_iter):
        msg = bp2_iteration(S, msg, llr, ss, factor, cn_type)
    x = -(llr + S.vn_sum(msg)).astype(F)
    return x.T.copy(), (x.T > 0).astype(np.uint8)


def _act(name):
    return {"tanh": np.tanh, "relu": lambda v: np.maximum(v, 0), None: lambda v: v}[name]


def gnn(X, Z, weights, h_vn, logit_hx, logit_hz, synd_x, synd_z, activation="tanh",
        reduce_op="mean"):
    """Feedback_GNN.call (feedback_gnn.py:161-188). h_vn [B,n,3] -> [B,n,3]."""
    W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3 = [np.asarray(w, F) for w in weights]
    act = _act(activation)
    h_vn = np.asarray(h_vn, F)
    out_m = []
    for S, logit, synd, W1, b1, W2, b2 in ((X, logit_hx, synd_x, W1x, b1x, W2x, b2x),
                                           (Z, logit_hz, synd_z, W1z, b1z, W2z, b2z)):
        h_cn = (np.asarray(logit, F) * (1 - 2 * np.asarray(synd).astype(np.int32)).astype(F)).T  # [B,m]
        feat = np.concatenate([h_cn[:, S.cn_of_edge, None], h_vn[:, S.vn_of_edge, :]], -1)
        msg = (act((feat @ W1 + b1).astype(F)).astype(F) @ W2 + b2).astype(F)    # [B,E,M]
        if reduce_op in ("mean", "sum"):
            red = np.add.reduceat(msg, S.vn_starts, axis=1)
            if reduce_op == "mean":
                red = red / S.vn_deg[None, :, None].astype(F)
        elif reduce_op == "max":
            red = np.maximum.reduceat(msg, S.vn_starts, axis=1)
        else:
            red = np.minimum.reduceat(msg, S.vn_starts, axis=1)
        out_m.append(red.astype(F))
    inp = np.concatenate(out_m + [h_vn], -1)
    return ((act((inp @ W3 + b3).astype(F)).astype(F) @ W0) + b0).astype(F)


def pauli_from_uniform(u, p):
    """Pauli.call non-wt branch on given uniforms (pauli.py:98-108)."""
    px, py, pz = F(2 * p / 3), F(p / 3), F(2 * p / 3)
    u = np.asarray(u, F)
    return (u < px).astype(np.uint8), ((u >= (px - py)) & (u < ((px + pz) - py))).astype(np.uint8)


def pipeline(code, X, Z, num_iters, weights_list, noise_x, noise_z, prior, factors=None):
    """Sandwich_BP_GNN_Evaluation_Model.call on given noise [B,n]; returns (s_hat, ls_hat)
    exactly as the reference builds them (dense hx_perp/hz_perp products)."""
    nx, nz = np.asarray(noise_x).astype(np.int64), np.asarray(noise_z).astype(np.int64)
    B, n = nx.shape
    sx = (code.hx @ nz.T) & 1
    sz = (code.hz @ nx.T) & 1
    llr = np.full((B, 3, n), prior, F)
    factors = factors or [1.0] * len(num_iters)
    r = bp4(X, Z, llr, sx, sz, num_iters[0], factors[0])
    xh, zh = r["x_hat"].copy(), r["z_hat"].copy()
    errors = np.ones(B, bool)
    for i in range(1, len(num_iters)):
        s1 = (code.hz @ xh.T.astype(np.int64)) & 1
        s2 = (code.hx @ zh.T.astype(np.int64)) & 1
        new_errors = np.any(s1 != sz, 0) | np.any(s2 != sx, 0)
        errors &= new_errors
        h_vn = np.stack([r["Lx"], r["Ly"], r["Lz"]], -1)
        new_llr = gnn(X, Z, weights_list[i - 1], h_vn, r["z_logit"], r["x_logit"], sx, sz)
        r = bp4(X, Z, new_llr.transpose(0, 2, 1), sx, sz, num_iters[i], factors[i])
        xh[errors] = r["x_hat"][errors]
        zh[errors] = r["z_hat"][errors]
    xd = (xh.astype(np.int64) ^ nx).T
    zd = (zh.astype(np.int64) ^ nz).T
    s_hat = np.concatenate([(code.hz @ xd) & 1, (code.hx @ zd) & 1], 0).T
    ls_hat = np.concatenate([(code.hx_perp @ xd) & 1, (code.hz_perp @ zd) & 1], 0).T
    return s_hat, ls_hat


# ------------------------------------------------------------------ GNN_BP4 (gnn.py:71-751) ----
def _mlp2(x, w, act):
    """MLP with one hidden layer: w = (W1, b1, W2, b2) (gnn.py:31-69)."""
    h = act((x @ w[0] + (w[1] if w[1] is not None else 0)).astype(F)).astype(F)
    return (h @ w[2] + (w[3] if w[3] is not None else 0)).astype(F)


def _reduce(msg, starts, deg, reduce_op):
    if reduce_op in ("mean", "sum"):
        red = np.add.reduceat(msg, starts, axis=1)
        if reduce_op == "mean":
            red = red / deg[None, :, None].astype(F)
    elif reduce_op == "max":
        red = np.maximum.reduceat(msg, starts, axis=1)
    else:
        red = np.minimum.reduceat(msg, starts, axis=1)
    return red.astype(F)


def _soft_rows(rows_side, l):
    S = rows_side
    vals = l[S.vn_of_edge[S.ind_cn]]
    sign = np.where(vals < 0, F(-1), F(1))
    node = S.cn_reduce(np.multiply, sign)
    T = S.cn_reduce(np.add, phi2(np.abs(vals))).astype(F)
    return (node * phi2(T)).astype(F)


def gnn_bp4(X, Z, LX, LZ, W, synd_x, synd_z, num_iter, activation="tanh", reduce_op="mean"):
    """GNN_BP4.call.  X, Z, LX, LZ: Side objects of hx, hz, lx, lz; W: dict with keys Winv, binv and the
    4-tuples cmx, cmz, cex, cez, vmx, vmz, ve; synd_* [B, m] 0/1.  Returns (list of (x_logit, z_logit)
    with shapes [m_z+k, B], [m_x+k, B]), x_hat [n,B], z_hat [n,B]."""
    act = _act(activation)
    sx, sz = np.asarray(synd_x), np.asarray(synd_z)
    B, n, d = sx.shape[0], X.n, W["Winv"].shape[0]
    ssx, ssz = (1 - 2 * sx.astype(np.int32)).astype(F), (1 - 2 * sz.astype(np.int32)).astype(F)
    h_vn = np.ones((B, n, d), F)
    h_cn = {0: np.zeros((B, X.m, d), F), 1: np.zeros((B, Z.m, d), F)}
    sides = {0: X, 1: Z}

    def cn_update(logits):
        for s, S in sides.items():
            vn_c, cn_c = S.vn_of_edge[S.ind_cn], S.cn_of_edge_c             # edges in CN order
            feat = np.concatenate([h_vn[:, vn_c, :], h_cn[s][:, cn_c, :]], -1)
            msg = _mlp2(feat, W["cmx" if s == 0 else "cmz"], act)
            m = _reduce(msg, S.cn_starts, S.cn_deg, reduce_op)
            inp = np.concatenate([m, h_cn[s], logits[s][:, :, None]], -1)
            h_cn[s] = _mlp2(inp, W["cex" if s == 0 else "cez"], act)

    cn_update({0: np.zeros((B, X.m), F), 1: np.zeros((B, Z.m), F)})
    out = []
    for it in range(num_iter):
        ms = []
        for s, S in sides.items():
            feat = np.concatenate([h_cn[s][:, S.cn_of_edge, :], h_vn[:, S.vn_of_edge, :]], -1)   # VN order
            msg = _mlp2(feat, W["vmx" if s == 0 else "vmz"], act)
            msg = msg * (ssx if s == 0 else ssz)[:, S.cn_of_edge, None]
            ms.append(_reduce(msg.astype(F), S.vn_starts, S.vn_deg, reduce_op))
        h_vn = _mlp2(np.concatenate(ms + [h_vn], -1), W["ve"], act)
        llr = (h_vn @ W["Winv"] + (W["binv"] if W["binv"] is not None else 0)).astype(F)       # [B,n,3]
        Lx, Ly, Lz = llr[..., 0].T, llr[..., 1].T, llr[..., 2].T                              # [n,B]
        llr_z = (softplus(-Lx) - logsumexp2(-Lz, -Ly)).astype(F)
        llr_x = (softplus(-Lz) - logsumexp2(-Lx, -Ly)).astype(F)
        hz_logit, lz_logit = _soft_rows(Z, llr_x), _soft_rows(LZ, llr_x)
        hx_logit, lx_logit = _soft_rows(X, llr_z), _soft_rows(LX, llr_z)
        out.append((np.concatenate([hz_logit, lz_logit], 0), np.concatenate([hx_logit, lx_logit], 0)))
        if it == num_iter - 1:
            break
        cn_update({0: (hx_logit.T * ssx).astype(F), 1: (hz_logit.T * ssz).astype(F)})
    xh, zh = decide(Lx, Ly, Lz)
    return out, xh, zh
