"""Weight (de)serialisation for the feedback GNN -- ``save_weights`` / ``load_weights`` of the
reference (``sionna/fec/ldpc/gnn.py:755-791``) without TensorFlow.

The reference pickles ``system.get_weights()``.  The four files it ships under ``weights/``
were written from a list of ``tf.Tensor`` -- the pickle stream calls
``tensorflow.python.framework.ops.convert_to_tensor(ndarray)`` and refers to numpy's old
``numpy.core.multiarray`` module path.  ``read_weights`` maps both to plain
numpy so the shipped files load as a list of 12 float32 arrays; ``save_weights`` writes a
pickle of plain ndarrays, which the reference's ``load_weights`` (``pickle.load`` +
``set_weights``) reads unchanged.
"""
import io
import os
import pickle

import numpy as np

WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


# everything a weights file may name: numpy's array reconstruction helpers (old and new module paths) and the
# TensorFlow tensor constructor the shipped files were written with.  Anything else is refused, so loading a
# weights file cannot run arbitrary code (plain pickle.load, as the reference uses, can).
_ALLOWED_NUMPY = {("multiarray", "_reconstruct"), ("multiarray", "scalar"), ("numeric", "_frombuffer"),
                  ("", "ndarray"), ("", "dtype")}


class _WeightsUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("tensorflow"):
            if name == "convert_to_tensor":
                return lambda x, *a, **k: np.asarray(x)
            raise pickle.UnpicklingError(f"unsupported TensorFlow object {module}.{name} in weights file")
        for root in ("numpy.core", "numpy._core", "numpy"):
            if module == root or module.startswith(root + "."):
                sub = module[len(root):].lstrip(".")
                if (sub, name) not in _ALLOWED_NUMPY:
                    break
                target = "numpy._core" + ("." + sub if sub else "") if root != "numpy" else module
                if sub == "":
                    return getattr(np, name)
                try:
                    return super().find_class(target, name)
                except (ImportError, AttributeError):
                    return super().find_class("numpy.core." + sub, name)
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a weights file")


def read_weights(model_path):
    """Return the list of float32 arrays stored in a reference weights file."""
    with open(model_path, "rb") as f:
        data = f.read()
    weights = _WeightsUnpickler(io.BytesIO(data)).load()
    return [np.ascontiguousarray(np.asarray(w), dtype=np.float32) for w in weights]


def save_weights(system, model_path):
    """Save ``system.get_weights()`` to ``model_path`` (gnn.py:757-772)."""
    weights = [np.asarray(w) for w in system.get_weights()]
    with open(model_path, "wb") as f:
        pickle.dump(weights, f)


def load_weights(system, model_path):
    """Load the weights stored at ``model_path`` into ``system`` (gnn.py:774-791)."""
    system.set_weights(read_weights(model_path))


# how the per-node matrix products run: FP32 FMAs (bit-exact with the oracle) or tcgen05 tensor cores with the
# 3-product TF32 split (float32 accuracy; needs reduce_op mean / sum and tanh)
GEMM_MODES = {"fma": 0, "tf32x3": 1}


class GNN_BP4:
    """The full GNN message-passing decoder -- ``GNN_BP4`` of the reference (``gnn.py:71-420`` with
    ``UpdateCNEmbeddings`` ``:423-610`` and ``UpdateVNEmbeddings`` ``:613-751``).

    ``GNN_BP4(code, num_embed_dims, num_msg_dims, num_hidden_units, num_mlp_layers, num_iter, ...)``
    called with ``(syndrome_x [B,m_x], syndrome_z [B,m_z])`` (batch first) returns
    ``(llr_hat, x_hat, z_hat)``: ``llr_hat`` is a list with one ``(x_perp_logit [m_z+k,B],
    z_perp_logit [m_x+k,B])`` pair per iteration, ``x_hat`` / ``z_hat`` are ``[n,B]``.

    As shipped the reference's ``call`` cannot run (it unpacks five values from ``cal_logit``, which
    returns four -- SURVEY.md F9); this layer implements the evident intent (the fifth value is unused).
    The reference ships no weights for it: ``get_weights`` / ``set_weights`` use the Keras order
    ``[_llr_inv_embed (W,b), update_h_cn: msg_mlp_x, msg_mlp_z, embed_mlp_x, embed_mlp_z, update_h_vn:
    msg_mlp_x, msg_mlp_z, embed_mlp]`` with ``(W1, b1, W2, b2)`` per MLP.  This build provides
    20 / 40 / 20 embedding / hidden / message dims with 2-layer MLPs and no node/edge attributes.

    ``gemm="tf32x3"`` (extension) evaluates the per-node matrix products on the tensor cores.
    """

    def __init__(self, code, num_embed_dims, num_msg_dims, num_hidden_units, num_mlp_layers, num_iter,
                 reduce_op="mean", activation="tanh", clip_llr_to=None, use_attributes=False,
                 node_attribute_dims=0, msg_attribute_dims=0, use_bias=False, input_embed=False,
                 loss_type="boxplus-phi", ctx=None, gemm="fma", chunk_frames=0):
        if int(num_mlp_layers) != 2 or use_attributes:
            raise NotImplementedError("this build provides 2-layer MLPs without node/edge attributes")
        if loss_type != "boxplus-phi":
            raise NotImplementedError("only loss_type='boxplus-phi' (soft syndromes) is provided")
        if reduce_op not in ("mean", "sum", "max", "min"):
            raise ValueError("unknown reduce operation")
        if gemm not in GEMM_MODES:
            raise ValueError("gemm must be 'fma' or 'tf32x3'")
        self._gemm = gemm
        self._chunk_frames = int(chunk_frames)          # extension: frames per pass over the batch (0 = automatic)
        self._code = code
        self._d, self._M, self._H = int(num_embed_dims), int(num_msg_dims), int(num_hidden_units)
        self._num_iter = int(num_iter)
        self._reduce_op, self._activation, self._use_bias = reduce_op, activation, bool(use_bias)
        self._ctx = ctx
        self._weights = None
        self._handle = None

    @property
    def num_iter(self):
        return self._num_iter

    @num_iter.setter
    def num_iter(self, value):
        self._num_iter = int(value)

    def _shapes(self):
        d, M, H = self._d, self._M, self._H
        mlp = lambda k_in, k_out: [(k_in, H), (H,), (H, k_out), (k_out,)]
        s = [(d, 3), (3,)]
        s += mlp(2 * d, M) + mlp(2 * d, M) + mlp(M + d + 1, d) + mlp(M + d + 1, d)      # update_h_cn
        s += mlp(2 * d, M) + mlp(2 * d, M) + mlp(2 * M + d, d)                         # update_h_vn
        return s if self._use_bias else s[0::2]

    def build(self, input_shape=None):
        """Keras initialisers: Dense kernels Glorot-uniform, MLP biases ones (gnn.py:52-58); _llr_inv_embed kernel
        zeros, bias ones (gnn.py:254-255)."""
        if self._weights is not None:
            return
        rng = np.random.default_rng(0)
        w = []
        for i, shp in enumerate(self._shapes()):
            if len(shp) == 1:
                w.append(np.ones(shp, np.float32))
            elif i == 0:
                w.append(np.zeros(shp, np.float32))
            else:
                lim = np.sqrt(6.0 / (shp[0] + shp[1]))
                w.append(rng.uniform(-lim, lim, shp).astype(np.float32))
        self._weights = w

    def get_weights(self):
        self.build()
        return [a.copy() for a in self._weights]

    def set_weights(self, weights):
        shapes = self._shapes()
        weights = [np.ascontiguousarray(np.asarray(w), dtype=np.float32) for w in weights]
        if len(weights) != len(shapes):
            raise ValueError(f"expected {len(shapes)} weight arrays, got {len(weights)}")
        for w, s in zip(weights, shapes):
            if w.shape != s:
                raise ValueError(f"Layer weight shape {s} not compatible with provided weight shape {w.shape}")
        self._weights = weights
        self._drop()

    def _drop(self):
        if getattr(self, "_handle", None) is not None:
            try:
                from . import _ffi
                _ffi.lib().fbgnn_gbp_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._drop()

    def _device_handle(self, ctx):
        import ctypes as C
        from . import _ffi
        from .feedback_gnn import ACTS, REDUCE
        self.build()
        if self._handle is None:
            full = list(self._weights)
            if not self._use_bias:                       # interleave missing biases as NULL
                full = [x for w in full for x in (w, None)]
            fp = C.POINTER(C.c_float)
            arr = (fp * 30)(*[None if a is None else a.ctypes.data_as(fp) for a in full])
            h = C.c_void_p()
            _ffi.call("fbgnn_gbp_create", ctx.handle, self._d, self._H, self._M, ACTS[self._activation],
                      REDUCE[self._reduce_op], arr, C.byref(h))
            self._handle = h
            _ffi.call("fbgnn_gbp_set_gemm", h, GEMM_MODES[self._gemm])
            _ffi.call("fbgnn_gbp_set_chunk", h, self._chunk_frames)
        return self._handle

    def __call__(self, inputs):
        from . import _ffi
        from .decoding_q import _is_device, _to_u8
        syndrome_x, syndrome_z = inputs
        dev = _ffi.device_code(self._code, self._ctx)
        ctx = dev.ctx
        on_device = _is_device(syndrome_x) or _is_device(syndrome_z)
        sx = ctx.asarray(_to_u8(syndrome_x), np.uint8)
        sz = ctx.asarray(_to_u8(syndrome_z), np.uint8)
        B = sx.shape[0]
        if sx.shape != (B, dev.mx) or sz.shape != (B, dev.mz) or not (sx.is_contiguous() and sz.is_contiguous()):
            raise ValueError(f"syndromes must be contiguous [B,{dev.mx}] and [B,{dev.mz}] (batch first)")
        kx, kz = np.asarray(self._code.lx).shape[0], np.asarray(self._code.lz).shape[0]
        it = self._num_iter
        xl = ctx.empty((it, B, dev.mz + kz), np.float32).transpose((0, 2, 1))
        zl = ctx.empty((it, B, dev.mx + kx), np.float32).transpose((0, 2, 1))
        xh = ctx.empty((B, dev.n), np.uint8).T
        zh = ctx.empty((B, dev.n), np.uint8).T
        _ffi.call("fbgnn_gbp_decode", dev.handle, self._device_handle(ctx), it, B, sx.t2(), sz.t2(), xl.t3(), zl.t3(),
                  xh.t2(), zh.t2())
        if on_device:
            return (xl, zl), xh, zh
        xl_h, zl_h = xl.numpy(), zl.numpy()
        return ([(xl_h[i], zl_h[i]) for i in range(it)], xh.numpy().astype(np.int64), zh.numpy().astype(np.float64))

    call = __call__
