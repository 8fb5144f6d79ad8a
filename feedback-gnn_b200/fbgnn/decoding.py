"""Binary syndrome BP decoder -- the ``LDPCBPDecoder`` layer.

Constructor arguments and call contract follow the reference's Keras layer
(``sionna/fec/ldpc/decoding.py:261-370`` ``__init__``, ``:875-1048`` ``call``) for the
configuration the quantum path uses: flooding BP on an arbitrary parity-check matrix,
optionally syndrome-based (``is_syndrome=True``: inputs ``(llr, syndrome)``), with
``cn_type`` in {"boxplus-phi", "boxplus", "minsum"} and ``normalization_factor``.
The decoding runs in the CUDA kernel ``k_bp2`` behind ``fbgnn_bp2_decode``.

Edge order: variable-node sorted, as the reference intends (its ``sp.sparse.find`` relied on
column-major output, which scipy >= 1.11 no longer gives -- SURVEY.md F8).

``trainable`` (per-edge weights), ``stateful`` (message state in / out) and ``track_exit`` (EXIT trajectory
``ie_v`` / ``ie_c``, ``decoding.py:955-1000``) are provided; the latter steps the kernel one iteration at a time and
evaluates the mutual-information estimate ``llr2mi`` (``fec/utils.py:151-218``) on the host.
"""
import numpy as np

from . import _ffi
from .decoding_q import CN_TYPES, _is_device, _to_u8


def _llr2mi(llr):
    """``sionna.fec.utils.llr2mi`` (``fec/utils.py:203-218``) for an all-zero codeword: 1 - mean(log2(1 + exp(llr)))
    with the LLRs clipped to +-20, float32."""
    z = np.clip(np.asarray(llr, np.float32), np.float32(-20.0), np.float32(20.0))
    x = np.log(np.float32(1.0) + np.exp(z)) / np.float32(np.log(2.0))
    return np.float32(1.0) - np.mean(x, dtype=np.float32)


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
        self._track_exit = track_exit
        self._ie_c = 0
        self._ie_v = 0
        if stateful and is_syndrome:
            raise ValueError("the reference takes either (llr, msg_vn) or (llr, syndrome), decoding.py:901-909")
        self._stateful = stateful
        self._has_weights = trainable
        self._edge_weights = None         # ones(num_edges) once the graph exists (decoding.py:361-366)
        self._edge_weights_dev = None
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

    @property
    def ie_c(self):
        "Extrinsic mutual information at check node."
        return self._ie_c

    @property
    def ie_v(self):
        "Extrinsic mutual information at variable node."
        return self._ie_v

    def graph(self):
        if self._graph is None:
            self._graph = _ffi.Graph(self._pcm, self._ctx)
        return self._graph

    # -- trainable=True: one weight per edge on the variable-to-check messages, initialised to one -----------------
    @property
    def has_weights(self):
        return self._has_weights

    @property
    def edge_weights(self):
        """[num_edges] float32, edges sorted by (variable, check) -- the order of the reference's message vector."""
        if not self._has_weights:
            return []
        if self._edge_weights is None:
            self._edge_weights = np.ones(self.graph().E, np.float32)
        return self._edge_weights

    def get_weights(self):
        return [self.edge_weights.copy()] if self._has_weights else []

    def set_weights(self, weights):
        if not self._has_weights:
            if len(weights):
                raise ValueError("the decoder has no weights (trainable=False)")
            return
        w = np.ascontiguousarray(np.asarray(weights[0]), np.float32)
        if w.shape != (self.graph().E,):
            raise ValueError(f"Layer weight shape {(self.graph().E,)} not compatible with provided weight shape {w.shape}")
        self._edge_weights, self._edge_weights_dev = w, None

    def __call__(self, inputs):
        syndrome = msg_vn = None
        if self._stateful:
            llr_ch, msg_vn = inputs
        elif self._is_syndrome:
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
        ew = None
        if self._has_weights:
            if self._edge_weights_dev is None:
                self._edge_weights_dev = ctx.asarray(self.edge_weights, np.float32)
            ew = self._edge_weights_dev.ptr
        # stateful: the message state is the flat [num_edges, B] tensor of the reference's ragged msg_vn
        m_in = m_out = None
        if self._stateful:
            m_out = ctx.empty((B, g.E), np.float32).T
            if msg_vn is not None:
                m_in = ctx.asarray(msg_vn, np.float32)
                if m_in.shape != (g.E, B):
                    raise ValueError(f"msg_vn must have shape [{g.E},{B}]")
        def run(num_iter, state_in, state_out):
            _ffi.call("fbgnn_bp2_decode_ex", g.handle, CN_TYPES[self._cn_type], num_iter,
                      self._normalization_factor, B, llr.t2(), synd.t2() if synd is not None else _ffi.NULL2,
                      soft.t2(), hard.t2(), ew,
                      state_in.T.t2() if state_in is not None else _ffi.NULL2,
                      state_out.T.t2() if state_out is not None else _ffi.NULL2)

        if self._track_exit:
            # EXIT trajectory (decoding.py:955-1000): one kernel call per iteration through the message-state interface;
            # slot it = 1 .. num_iter of ie_v / ie_c holds the mutual-information estimate of the variable-node / check-node
            # output messages of that iteration (slot 0 stays 0, as in the reference)
            ie_v = np.zeros(self._num_iter + 1, np.float32)
            ie_c = np.zeros(self._num_iter + 1, np.float32)
            vn_of_edge = g.edge_vn()
            llr_true = -llr.numpy().T                                     # [n, B], the reference's internal sign
            state = m_in if m_in is not None else ctx.asarray(np.zeros((g.E, B), np.float32))
            bufs = [ctx.empty((B, g.E), np.float32).T, ctx.empty((B, g.E), np.float32).T]
            if self._num_iter == 0:
                run(0, state, bufs[0])
                state = bufs[0]
            for it in range(1, self._num_iter + 1):
                c2v = state.numpy()                                       # [E, B], edges sorted by variable
                tot = llr_true.copy()
                np.add.at(tot, vn_of_edge, c2v)
                ie_v[it] = _llr2mi(-(tot[vn_of_edge] - c2v))
                nxt = bufs[it % 2]
                run(1, state, nxt)
                ie_c[it] = _llr2mi(-nxt.numpy())
                state = nxt
            if self._stateful:
                m_out = state
            self._ie_v, self._ie_c = ie_v, ie_c
        else:
            run(self._num_iter, m_in, m_out)
        if on_device:
            res = hard if self._hard_out else soft
            return (res, m_out) if self._stateful else res
        out = hard.numpy().astype(self._output_dtype) if self._hard_out else soft.numpy().astype(self._output_dtype)
        out = out.reshape(lead_shape + (g.n,))
        return (out, m_out.numpy()) if self._stateful else out

    call = __call__
