"""Pauli noise channel -- the ``Pauli`` layer of the reference (``sionna/channel/pauli.py:60-117``).

``wt=False``: iid depolarising-type noise from one uniform per qubit, ``noise_x = u < px``,
``noise_z = (u >= px - py) and (u < px + pz - py)`` (pauli.py:98-108).  ``wt=True``: exactly ``wt``
erroneous qubits per frame, each X / Y / Z with probability 1/3 (pauli.py:80-96).  Both are sampled
on the GPU with Philox4x32-10 keyed by ``seed`` and counted by the global frame id, so a
Monte-Carlo run gives the same frames however it is sharded.
"""
import ctypes as C

import numpy as np

from . import _ffi


def pauli_thresholds(p=None, px=None, py=None, pz=None):
    """float32 thresholds {px, px - py, (px + pz) - py}.  With only ``p`` given: the
    depolarising split px = pz = 2p/3, py = p/3 of feedback_gnn.py:298."""
    if px is None:
        px, py, pz = 2 * p / 3, p / 3, 2 * p / 3
    px, py, pz = np.float32(px), np.float32(py), np.float32(pz)
    return np.array([px, px - py, (px + pz) - py], dtype=np.float32)


class Pauli:
    def __init__(self, dtype=np.uint8, wt=False, seed=0, first_frame=0, ctx=None, **kwargs):
        self._wt = wt
        self._dtype = np.dtype(dtype)
        self.seed = int(seed)
        self.next_frame = int(first_frame)
        self._ctx = ctx

    def sample_device(self, B, n, thr):
        ctx = self._ctx or _ffi.default_context()
        nx, nz = ctx.empty((B, n), np.uint8), ctx.empty((B, n), np.uint8)
        thr = np.ascontiguousarray(thr, np.float32)
        _ffi.call("fbgnn_pauli_sample", ctx.handle, n, B, thr.ctypes.data_as(C.POINTER(C.c_float)),
                  self.seed, self.next_frame, nx.t2(), nz.t2())
        self.next_frame += B
        return nx, nz

    def sample_device_wt(self, B, n, wt):
        ctx = self._ctx or _ffi.default_context()
        nx, nz = ctx.empty((B, n), np.uint8), ctx.empty((B, n), np.uint8)
        _ffi.call("fbgnn_pauli_sample_wt", ctx.handle, n, B, int(wt), self.seed, self.next_frame, nx.t2(), nz.t2())
        self.next_frame += B
        return nx, nz

    def __call__(self, inputs):
        """``[cx, cz, px, py, pz]`` (or ``[cx, cz, wt]`` with ``wt=True``) -> ``(noise_x, noise_z)`` bool
        ``[B,n]`` when ``cz`` is None, else ``(y_x, y_z, noise_x, noise_z)`` (pauli.py:78-117)."""
        if self._wt:
            cx, cz, wt = inputs
            cx = np.asarray(cx)
            B, n = cx.shape
            nx, nz = self.sample_device_wt(B, n, int(np.asarray(wt)))
        else:
            cx, cz, px, py, pz = inputs
            cx = np.asarray(cx)
            B, n = cx.shape
            nx, nz = self.sample_device(B, n, pauli_thresholds(px=float(px), py=float(py), pz=float(pz)))
        nx, nz = nx.numpy().astype(bool), nz.numpy().astype(bool)
        if cx is not None and cz is not None:
            return np.logical_xor(cx.astype(bool), nx), np.logical_xor(np.asarray(cz).astype(bool), nz), nx, nz
        return nx, nz

    call = __call__
