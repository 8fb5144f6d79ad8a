"""Quaternary (BP4) syndrome decoder for CSS codes -- the ``QLDPCBPDecoder`` layer.

Same constructor arguments and call contract as the reference's Keras layer
(``sionna/fec/ldpc/decoding_q.py:14-111`` ``__init__``, ``:661-797`` ``call``); the decoding
itself runs in the CUDA kernel ``k_bp4`` behind ``fbgnn_bp4_decode`` (messages resident in
shared memory for all ``num_iter`` iterations).

    decoder = QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0,
                             cn_type="boxplus-phi", stage_one=True)
    llrx, llry, llrz, x_hat, z_hat, x_logit, z_logit = decoder((llr_ch, syndrome_x, syndrome_z))

Inputs: ``llr_ch`` float32 ``[B,3,n]`` (x, y, z priors, ``log p_I/p_P``), ``syndrome_x``
``[m_x,B]`` and ``syndrome_z`` ``[m_z,B]`` with 0/1 entries.  Each may be a numpy array (copied
to the GPU) or a DLPack-capable CUDA tensor / ``DeviceArray`` (used in place).  If any input
is a device tensor the outputs are ``DeviceArray`` objects, otherwise numpy arrays with the
reference's dtypes (``x_hat`` int64, ``z_hat`` float64).
"""
import numpy as np

from . import _ffi

CN_TYPES = {"boxplus-phi": 0, "boxplus": 1, "minsum": 2}


def _is_device(x):
    return isinstance(x, _ffi.DeviceArray) or (
        not isinstance(x, np.ndarray) and hasattr(x, "__dlpack_device__") and x.__dlpack_device__()[0] == 2)


class QLDPCBPDecoder:
    def __init__(self,
                 code,
                 trainable=False,
                 cn_type='boxplus',
                 hard_out=True,
                 track_exit=False,
                 num_iter=32,
                 normalization_factor=0.625,
                 output_dtype=np.float32,
                 loss_type='boxplus-phi',
                 stage_one=False,
                 stage_two=False,
                 ctx=None,
                 early_stop=False,
                 **kwargs):
        if cn_type not in CN_TYPES:
            raise ValueError('Unknown node type.')
        # In the reference `trainable` creates no variables (every weight is commented out, decoding_q.py:114-135,
        # 240-242, 748-749); it only switches the per-iteration soft syndromes on (:743, :794).  Their rows are those of
        # hx_perp / hz_perp -- dense kernel bases -- unless stage_one / stage_two select hz / hx (:32-37); with
        # stage_one=True the flag has no effect at all (the stage_one return comes first, :792-793).
        self._perp_rows = bool(trainable) and not (stage_one or stage_two)
        if self._perp_rows and np.asarray(code.hx_perp).shape[0] != np.asarray(code.hz_perp).shape[0]:
            raise ValueError("trainable mode stacks x and z soft syndromes: hx_perp and hz_perp need equal row counts")
        # extension (not in the reference): leave a frame's iteration loop once its decision reproduces the
        # syndrome; the call then also returns the iterations used per frame
        self._early_stop = bool(early_stop)
        self._code = code
        self._cn_type = cn_type
        self._hard_out = hard_out
        self._track_exit = track_exit
        self._num_iter = int(num_iter)
        self._normalization_factor = float(normalization_factor)
        self._output_dtype = output_dtype
        self._loss_type = loss_type
        self._stage_one = stage_one
        self._stage_two = stage_two
        self._trainable = trainable
        self._num_cns_x = code.hx.shape[0]
        self._num_cns_z = code.hz.shape[0]
        self._num_vns = code.hx.shape[1]
        self._ctx = ctx
        self._dev = None

    # properties the evaluation model reads to fuse the pipeline
    @property
    def num_iter(self):
        return self._num_iter

    @property
    def normalization_factor(self):
        return self._normalization_factor

    @property
    def cn_type(self):
        return self._cn_type

    @property
    def code(self):
        return self._code

    def _device(self):
        if self._dev is None:
            self._dev = _ffi.device_code(self._code, self._ctx)
        return self._dev

    def __call__(self, inputs):
        llr_ch, syndrome_x, syndrome_z = inputs
        dev = self._device()
        ctx = dev.ctx
        n, mx, mz = dev.n, dev.mx, dev.mz
        on_device = any(_is_device(t) for t in (llr_ch, syndrome_x, syndrome_z))
        if not _is_device(llr_ch):
            llr_ch = np.asarray(llr_ch)
            if llr_ch.dtype != np.float32:
                raise TypeError('Invalid input dtype.')
        llr = ctx.asarray(llr_ch, np.float32)
        if llr.ndim != 3 or llr.shape[1] != 3 or llr.shape[2] != n:
            raise ValueError('Last dimension must be of length n.')
        B = llr.shape[0]
        sx = ctx.asarray(_to_u8(syndrome_x), np.uint8)
        sz = ctx.asarray(_to_u8(syndrome_z), np.uint8)
        if sx.shape != (mx, B) or sz.shape != (mz, B):
            raise ValueError(f"syndromes must have shapes [{mx},{B}] and [{mz},{B}]")
        if (self._stage_two or self._trainable) and not self._stage_one:
            # (llr_hat [2*num_iter+2, m, B], x_hat, z_hat): decoding_q.py:794-795
            out = self.decode_device(llr, sx, sz, want_logits=False, want_iter_logits=True)
            xh, zh, llr_hat = out[3], out[4], out[-1]
            if on_device:
                return llr_hat, xh, zh
            return llr_hat.numpy(), xh.numpy().astype(np.int64), zh.numpy().astype(np.float64)
        out = self.decode_device(llr, sx, sz, want_logits=self._stage_one)
        Lx, Ly, Lz, xh, zh, xl, zl = out[:7]
        self.last_iterations = out[-1].numpy() if self._early_stop else None
        if self._stage_one:
            if on_device:
                return Lx, Ly, Lz, xh, zh, xl, zl
            return (Lx.numpy(), Ly.numpy(), Lz.numpy(), xh.numpy().astype(np.int64),
                    zh.numpy().astype(np.float64), xl.numpy(), zl.numpy())
        if on_device:
            return xh, zh
        return xh.numpy().astype(np.int64), zh.numpy().astype(np.float64)

    call = __call__

    def decode_device(self, llr, sx, sz, want_logits=True, want_msgs=False, prior=None, want_iter_logits=False):
        """Device-level entry: ``llr`` DeviceArray [B,3,n] (or None with a scalar ``prior``),
        syndromes DeviceArray [m,B].  Returns DeviceArrays (Lx, Ly, Lz [B,n] f32, x_hat, z_hat
        [B,n] u8, x_logit [m_z,B], z_logit [m_x,B] f32 or None) (+ msg_x, msg_z [B,E])."""
        dev = self._device()
        ctx = dev.ctx
        n, mx, mz = dev.n, dev.mx, dev.mz
        B = sx.shape[1]
        Lx, Ly, Lz = (ctx.empty((B, n), np.float32) for _ in range(3))
        xh, zh = ctx.empty((B, n), np.uint8), ctx.empty((B, n), np.uint8)
        xl = zl = None
        rows = None
        if self._perp_rows:
            rows = self._device_rows()
            mz, mx = rows[0].m, rows[1].m                # x_logit over hx_perp rows, z_logit over hz_perp rows
        if want_logits:
            # frame-major storage, exposed in the reference's [m, B] orientation
            xl = ctx.empty((B, mz), np.float32).T
            zl = ctx.empty((B, mx), np.float32).T
        msgs = ()
        if want_msgs:
            msgs = (ctx.empty((B, dev.Ex), np.float32), ctx.empty((B, dev.Ez), np.float32))
        il = ()
        if want_iter_logits:
            if mx != mz:
                raise ValueError("per-iteration soft syndromes need x and z row sets with the same number of rows")
            il = (ctx.empty((2 * self._num_iter + 2, B, mx), np.float32).transpose((0, 2, 1)),)
        t2 = lambda a: a.t2() if a is not None else _ffi.NULL2
        iters = (ctx.empty((B,), np.uint8),) if self._early_stop else ()
        opts = _ffi.Bp4Opts(iters[0].ptr if iters else None, rows[0].handle if rows else None,
                            rows[1].handle if rows else None)
        import ctypes as C
        _ffi.call("fbgnn_bp4_decode_ex", dev.handle, CN_TYPES[self._cn_type], self._num_iter,
                  self._normalization_factor, B,
                  llr.t3() if llr is not None else _ffi.NULL3, float(prior or 0.0),
                  sx.t2(), sz.t2(), Lx.t2(), Ly.t2(), Lz.t2(), xh.t2(), zh.t2(), t2(xl), t2(zl),
                  msgs[0].t2() if want_msgs else _ffi.NULL2, msgs[1].t2() if want_msgs else _ffi.NULL2,
                  il[0].t3() if want_iter_logits else _ffi.NULL3, C.byref(opts))
        return (Lx, Ly, Lz, xh, zh, xl, zl) + msgs + il + iters

    def _device_rows(self):
        """(rows of hx_perp, rows of hz_perp) on the device, one copy per decoder."""
        if getattr(self, "_rows", None) is None:
            ctx = self._device().ctx
            self._rows = (_ffi.Rows(self._code.hx_perp, ctx), _ffi.Rows(self._code.hz_perp, ctx))
        return self._rows


def _to_u8(s):
    if _is_device(s):
        return s
    s = np.asarray(s)
    return s.astype(np.uint8) if s.dtype != np.uint8 else s
