"""Training of the feedback GNN -- the two-stage models of the reference and the pieces of its training
notebooks, without TensorFlow.

* ``First_Stage_BP_Model`` / ``Second_Stage_GNN_BP_Model`` (``sionna/fec/ldpc/feedback_gnn.py:364-460``):
  same constructors and call contracts.  The reference takes the gradient with ``tf.GradientTape``
  around the second model's call; here ``model.gradients()`` returns d loss / d weights of the last
  call (aligned with ``feedback.get_weights()``), computed by the hand-written reverse sweep behind
  ``fbgnn_second_stage_grad`` (``csrc/fbgnn_train.cuh``).
* ``Adam`` + ``clip_by_value`` + ``train_step``: the update rule of ``examples/Feedback_GNN.ipynb`` cells 2 / 8
  (Adam, learning rate 2e-4, element-wise gradient clipping at +-10).
* ``BP4_Error_Model`` / ``Feedback_GNN_Error_Model`` (``examples/Generate_dataset.ipynb`` cell 1): return the
  error strings the decoder failed on -- the training-set generators.
"""
import ctypes as C

import numpy as np

from . import _ffi
from .decoding_q import QLDPCBPDecoder, _to_u8
from .feedback_gnn import Feedback_GNN, Sandwich_BP_GNN_Evaluation_Model
from .pauli import Pauli

# blocks of the packed gradient returned by fbgnn_second_stage_grad: (rows incl. the bias row, columns)
_GRAD_BLOCKS = ((41, 3), (5, 40), (41, 20), (5, 40), (41, 20), (44, 40))


_graph_cache = {}


def _graphs(code, ctx):
    key = (id(code), id(ctx))
    ent = _graph_cache.get(key)
    if ent is None or ent[0] is not code:
        ent = (code, _ffi.Graph(code.hx, ctx), _ffi.Graph(code.hz, ctx))
        _graph_cache[key] = ent
    return ent[1], ent[2]


def _syndromes(code, ctx, noise_x, noise_z):
    """syndrome_x = hx . noise_z, syndrome_z = hz . noise_x (mod 2) as DeviceArrays [m,B] (feedback_gnn.py:383-384)."""
    dev = _ffi.device_code(code, ctx)
    gx, gz = _graphs(code, dev.ctx)
    nx = dev.ctx.asarray(_to_u8(np.asarray(noise_x)), np.uint8)
    nz = dev.ctx.asarray(_to_u8(np.asarray(noise_z)), np.uint8)
    B = nx.shape[0]
    sx, sz = dev.ctx.empty((dev.mx, B), np.uint8), dev.ctx.empty((dev.mz, B), np.uint8)
    _ffi.call("fbgnn_syndrome", gx.handle, B, nz.t2(), sx.t2())
    _ffi.call("fbgnn_syndrome", gz.handle, B, nx.t2(), sz.t2())
    return dev, sx, sz


class First_Stage_BP_Model:
    """First block of BP runs during training; nothing trainable (feedback_gnn.py:364-392).
    ``model(noise_x, noise_z) -> (h_vn [B,n,3], logit_hx_perp [m_z,B], logit_hz_perp [m_x,B])``."""

    def __init__(self, code, decoder, p0=0.05, ctx=None):
        if not isinstance(decoder, QLDPCBPDecoder) or not decoder._stage_one:
            raise TypeError("decoder must be an fbgnn QLDPCBPDecoder with stage_one=True")
        self.code, self.hx, self.hz = code, code.hx, code.hz
        self.decoder, self.p0, self._ctx = decoder, p0, ctx

    def __call__(self, noise_x, noise_z):
        dev, sx, sz = _syndromes(self.code, self._ctx, noise_x, noise_z)
        prior = np.float32(np.log(np.float64(np.float32(3. * (1. - self.p0) / self.p0))))
        Lx, Ly, Lz, _, _, xl, zl = self.decoder.decode_device(None, sx, sz, want_logits=True, prior=float(prior))
        h_vn = np.stack([Lx.numpy(), Ly.numpy(), Lz.numpy()], axis=-1)
        return h_vn, xl.numpy(), zl.numpy()

    call = __call__


class Second_Stage_GNN_BP_Model:
    """Feedback GNN followed by the stage-two BP decoder with the multi-loss of its soft syndromes
    (feedback_gnn.py:395-460).  ``model(noise_x, noise_z, h_vn, logit_hx_perp, logit_hz_perp) ->
    (s_hat [B, m_x+m_z], ls_hat [B, rows(hx_perp)+rows(hz_perp)], loss)``; ``model.gradients()`` afterwards."""

    def __init__(self, code, feedback, decoder, num_iter=16, trainable=True, loss_from=8, ctx=None):
        if not isinstance(feedback, Feedback_GNN):
            raise TypeError("feedback must be an fbgnn Feedback_GNN layer")
        if not isinstance(decoder, QLDPCBPDecoder) or not decoder._stage_two or decoder._stage_one:
            raise TypeError("decoder must be an fbgnn QLDPCBPDecoder with stage_two=True")
        if decoder.cn_type != "boxplus-phi":
            raise NotImplementedError("the gradient is provided for cn_type='boxplus-phi' (the one every script uses)")
        self.code = code
        self.k, self.n = code.K, code.N
        self.hx, self.hz, self.lx, self.lz = code.hx, code.hz, code.lx, code.lz
        self.hx_perp, self.hz_perp = code.hx_perp, code.hz_perp
        self.code_name = code.name
        self.num_checks = code.hx.shape[0] + code.hz.shape[0]
        self.feedback, self.decoder = feedback, decoder
        self.num_iter = int(np.asarray(num_iter))
        if self.num_iter != decoder.num_iter:
            raise ValueError("num_iter must equal the decoder's num_iter")
        self.loss_from = int(loss_from)
        self.trainable = bool(trainable)
        self._ctx = ctx
        self._grads = None
        # residual-syndrome products through float32 BLAS (exact: entries <= n < 2^24); numpy's integer matmul is slow
        self._f32 = [np.ascontiguousarray(m, np.float32) for m in (code.hz, code.hx, code.hx_perp, code.hz_perp)]

    @property
    def trainable_variables(self):
        return self.feedback.get_weights() if self.trainable else []

    trainable_weights = trainable_variables

    def __call__(self, noise_x, noise_z, h_vn, logit_hx_perp, logit_hz_perp):
        noise_x, noise_z = np.asarray(noise_x).astype(bool), np.asarray(noise_z).astype(bool)
        dev, sx, sz = _syndromes(self.code, self._ctx, noise_x, noise_z)
        ctx = dev.ctx
        B = noise_x.shape[0]
        h = ctx.asarray(h_vn, np.float32)
        # the GNN pairs logit_hz_perp (soft syndromes over hx rows) with hx and logit_hx_perp with hz (:425)
        lhx, lhz = ctx.asarray(logit_hz_perp, np.float32), ctx.asarray(logit_hx_perp, np.float32)
        if h.shape != (B, dev.n, 3) or lhx.shape != (dev.mx, B) or lhz.shape != (dev.mz, B):
            raise ValueError("h_vn / logit shapes do not match the code and the batch")
        loss = C.c_double()
        grads = np.zeros(sum(r * c for r, c in _GRAD_BLOCKS), np.float32)
        _ffi.call("fbgnn_second_stage_grad", dev.handle, self.feedback.device_handle(ctx), self.num_iter,
                  self.decoder.normalization_factor, self.loss_from, B, h.t3(), lhx.t2(), lhz.t2(), sx.t2(), sz.t2(),
                  1 if self.trainable else 0, C.byref(loss), grads.ctypes.data_as(C.POINTER(C.c_float)))
        self._grads = self._unpack(grads) if self.trainable else None
        # decisions of the same forward pass (feedback_gnn.py:425-426,433-452)
        new_llr = self.feedback((h, lhx, lhz, sx, sz))
        _, x_hat, z_hat = self.decoder((new_llr.transpose((0, 2, 1)), sx, sz))
        x_diff = np.logical_xor(noise_x.T, x_hat.numpy().astype(bool).T).astype(np.float32)    # [n, B]
        z_diff = np.logical_xor(noise_z.T, z_hat.numpy().astype(bool).T).astype(np.float32)
        hz, hx, hxp, hzp = self._f32
        mod2 = lambda a: a.astype(np.int64) & 1
        s_hat = np.concatenate([mod2(hz @ x_diff), mod2(hx @ z_diff)], axis=0).T
        ls_hat = np.concatenate([mod2(hxp @ x_diff), mod2(hzp @ z_diff)], axis=0).T
        return s_hat, ls_hat, float(loss.value)

    call = __call__

    def _unpack(self, flat):
        out, off = [], 0
        blocks = []
        for r, c in _GRAD_BLOCKS:
            blk = flat[off:off + r * c].reshape(r, c)
            off += r * c
            blocks.append((blk[:-1].copy(), blk[-1].copy()))
        # Keras order: [W0, b0, W1x, b1x, W2x, b2x, W1z, b1z, W2z, b2z, W3, b3]
        for W, b in blocks:
            out += [W, b]
        return out

    def gradients(self):
        """d loss / d weights of the last call, one array per entry of ``feedback.get_weights()``."""
        if self._grads is None:
            raise RuntimeError("call the model first (with trainable=True)")
        return [g.copy() for g in self._grads]


def clip_by_value(grads, lo, hi):
    return [np.clip(g, lo, hi) for g in grads]


class CosineDecay:
    """tf.keras.optimizers.schedules.CosineDecay(initial_learning_rate, decay_steps, alpha=0): the schedule the
    training notebook suggests for multi-epoch runs.  Callable with the 0-based step."""

    def __init__(self, initial_learning_rate, decay_steps, alpha=0.0):
        self.initial_learning_rate, self.decay_steps, self.alpha = float(initial_learning_rate), int(decay_steps), float(alpha)

    def __call__(self, step):
        t = min(max(step, 0), self.decay_steps) / max(self.decay_steps, 1)
        return self.initial_learning_rate * ((1.0 - self.alpha) * 0.5 * (1.0 + np.cos(np.pi * t)) + self.alpha)


class Adam:
    """tf.keras.optimizers.Adam (defaults beta_1=0.9, beta_2=0.999, epsilon=1e-7) on a layer's weight list."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._m = self._v = None

    def apply_gradients(self, grads, layer):
        weights = layer.get_weights()
        if self._m is None:
            self._m = [np.zeros_like(w, dtype=np.float64) for w in weights]
            self._v = [np.zeros_like(w, dtype=np.float64) for w in weights]
        self.iterations += 1
        t = self.iterations
        lr = self.learning_rate(t - 1) if callable(self.learning_rate) else self.learning_rate
        alpha = lr * np.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        new = []
        for w, g, m, v in zip(weights, grads, self._m, self._v):
            g = np.asarray(g, np.float64)
            m += (1.0 - self.beta_1) * (g - m)
            v += (1.0 - self.beta_2) * (g * g - v)
            new.append((w - alpha * m / (np.sqrt(v) + self.epsilon)).astype(np.float32))
        layer.set_weights(new)


def train_step(model_stage_one, model_stage_two, optimizer, noise_x, noise_z, clip_value_grad=10.0):
    """One iteration of the loop of examples/Feedback_GNN.ipynb cell 2.  Returns (loss, bler, flagged_bler)."""
    h_vn, logit_hx_perp, logit_hz_perp = model_stage_one(noise_x, noise_z)
    s_hat, b_hat, loss = model_stage_two(noise_x, noise_z, h_vn, logit_hx_perp, logit_hz_perp)
    grads = clip_by_value(model_stage_two.gradients(), -clip_value_grad, clip_value_grad)
    optimizer.apply_gradients(grads, model_stage_two.feedback)
    return loss, float(np.mean(np.any(b_hat, axis=1))), float(np.mean(np.any(s_hat, axis=1)))


class _ErrorModel:
    """Shared part of the dataset generators: run the fused pipeline on freshly sampled noise and return the
    error strings whose final correction still mismatches the syndrome."""

    def __init__(self, code, decoders, feedbacks, wt, p0, seed, ctx):
        self.code, self.n, self.wt = code, code.N, wt
        self.channel = Pauli(wt=wt, seed=seed, ctx=ctx)
        # skip_inactive: frames the first stage decodes do not run the later stages (result-identical)
        self._pipe = Sandwich_BP_GNN_Evaluation_Model(code, decoders, feedbacks, num_layers=len(decoders), wt=wt,
                                                      p0=p0, skip_inactive=True, ctx=ctx)

    def __call__(self, batch_size, ebno_db):
        B, p = int(np.asarray(batch_size)), float(np.asarray(ebno_db))
        if self.wt:
            nx, nz = self.channel.sample_device_wt(B, self.n, int(round(p)))
        else:
            from .pauli import pauli_thresholds
            nx, nz = self.channel.sample_device(B, self.n, pauli_thresholds(p))
        res = self._pipe.run(B, p, noise=(nx, nz), want_diff=False)
        err = (res["flags"].numpy() & 1).astype(bool)
        return nx.numpy().astype(bool)[err], nz.numpy().astype(bool)[err]

    call = __call__


class BP4_Error_Model(_ErrorModel):
    """Error strings plain BP4 fails to decode (Generate_dataset.ipynb cell 1, ``BP4_Error_Model``)."""

    def __init__(self, code, decoder, num_iter=32, trainable=False, loss_type="boxplus-phi", wt=False, seed=0, ctx=None):
        d = QLDPCBPDecoder(code, num_iter=decoder.num_iter, normalization_factor=decoder.normalization_factor,
                           cn_type=decoder.cn_type, stage_one=True, ctx=ctx)
        super().__init__(code, [d], [], wt, 0.05, seed, ctx)


class Feedback_GNN_Error_Model(_ErrorModel):
    """Error strings (BP, feedback GNN, BP) fails to decode (Generate_dataset.ipynb cell 1)."""

    def __init__(self, code, decoder1, feedback, decoder2, wt=False, p0=0.05, seed=0, ctx=None):
        mk = lambda d: QLDPCBPDecoder(code, num_iter=d.num_iter, normalization_factor=d.normalization_factor,
                                      cn_type=d.cn_type, stage_one=True, ctx=ctx)
        super().__init__(code, [mk(decoder1), mk(decoder2)], [feedback], wt, p0, seed, ctx)
