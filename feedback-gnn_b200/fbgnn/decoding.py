"""Binary syndrome BP decoder -- the ``LDPCBPDecoder`` layer.

Constructor arguments and call contract follow the reference's Keras layer
(``sionna/fec/ldpc/decoding.py:261-370`` ``__init__``, ``:875-1048`` ``call``) for the
configuration the quantum path uses: flooding BP on an arbitrary parity-check matrix,
optionally syndrome-based (``is_syndrome=True``: inputs ``(llr, syndrome)``), with
``cn_type`` in {"boxplus-phi", "boxplus", "minsum"} and ``normalization_factor``.
The decoding runs in the CUDA kernel ``k_bp2`` behind ``fbgnn_bp2_decode``.

Edge order: variable-node sorted, as the reference intends (its ``sp.sparse.find`` relied on
column-major output, which scipy >= 1.11 no longer gives -- SURVEY.md F8).

Not provided (outside the quantum hot path): ``trainable`` edge weights, ``stateful``
message passing between calls and ``track_exit`` EXIT-chart tracking.
"""
import numpy as np

from . import _ffi
from .decoding_q import CN_TYPES, _is_device, _to_u8


class LDPCBPDecoder:
    def __init__(self,
                 pcm,
                 trainable=False,
                 cn_type='boxplus-phi',
                 hard_out=True,
                 track_exit=False,
                 num_iter=32,
                 normalization_factor=1.0,
                 stateful=False,
                 is_syndrome=False,
                 output_dtype=np.float32,
                 ctx=None,
                 **kwargs):
        assert isinstance(trainable, bool), 'trainable must be bool.'
        assert isinstance(hard_out, bool), 'hard_out must be bool.'
        assert isinstance(track_exit, bool), 'track_exit must be bool.'
        assert isinstance(cn_type, str), 'cn_type must be str.'
        assert isinstance(num_iter, (int, np.integer)), 'num_iter must be int.'
        assert num_iter >= 0, 'num_iter cannot be negative.'
        assert isinstance(stateful, bool), 'stateful must be bool.'
        if cn_type not in CN_TYPES:
            raise ValueError('Unknown node type.')
        if trainable or stateful or track_exit:
            raise NotImplementedError("trainable / stateful / track_exit are outside the quantum hot path "
                                      "of this build")
        if not (isinstance(pcm, np.ndarray) or hasattr(pcm, "toarray")):
            raise TypeError("Unsupported dtype of pcm.")
        self._pcm = pcm
        self._cn_type = cn_type
        self._hard_out = hard_out
        self._num_iter = int(num_iter)
        self._is_syndrome = is_syndrome
        self._normalization_factor = float(normalization_factor)
        self._output_dtype = np.dtype(output_dtype)
        self._num_cns, self._num_vns = pcm.shape
        self._ctx = ctx
        self._graph = None

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
    def pcm(self):
        return self._pcm

    def graph(self):
        if self._graph is None:
            self._graph = _ffi.Graph(self._pcm, self._ctx)
        return self._graph

    def __call__(self, inputs):
        syndrome = None
        if self._is_syndrome:
            llr_ch, syndrome = inputs
        else:
            llr_ch = inputs
        g = self.graph()
        ctx = g.ctx
        on_device = _is_device(llr_ch) or (syndrome is not None and _is_device(syndrome))
        if not _is_device(llr_ch):
            llr_ch = np.asarray(llr_ch)
            if llr_ch.dtype != np.float32:
                raise TypeError('Invalid input dtype.')
            if llr_ch.shape[-1] != g.n:
                raise ValueError('Last dimension must be of length n.')
            lead_shape = llr_ch.shape[:-1]
            llr_ch = llr_ch.reshape(-1, g.n)
        else:
            lead_shape = None
        llr = ctx.asarray(llr_ch, np.float32)
        if llr.ndim != 2 or llr.shape[1] != g.n:
            raise ValueError('Last dimension must be of length n.')
        B = llr.shape[0]
        synd = None
        if syndrome is not None:
            synd = ctx.asarray(_to_u8(syndrome), np.uint8)
            if synd.shape != (g.m, B):
                raise ValueError(f"syndrome must have shape [{g.m},{B}]")
        soft = ctx.empty((B, g.n), np.float32)
        hard = ctx.empty((B, g.n), np.uint8)
        _ffi.call("fbgnn_bp2_decode", g.handle, CN_TYPES[self._cn_type], self._num_iter,
                  self._normalization_factor, B, llr.t2(), synd.t2() if synd is not None else _ffi.NULL2,
                  soft.t2(), hard.t2())
        if on_device:
            return hard if self._hard_out else soft
        out = hard.numpy().astype(self._output_dtype) if self._hard_out else soft.numpy().astype(self._output_dtype)
        return out.reshape(lead_shape + (g.n,))

    call = __call__
