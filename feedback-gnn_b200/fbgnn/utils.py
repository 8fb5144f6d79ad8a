"""Monte-Carlo harness: ``count_block_errors``, ``sim_ber`` and ``PlotBER``.

Same call contracts as the reference (``sionna/utils/metrics.py:194-223``,
``sionna/utils/misc.py:403-768`` with its ``qldpc=True`` branch, ``sionna/utils/plotting.py:148-447``)
for the quantum evaluation path: a model is called as ``mc_fun(batch_size=..., ebno_db=...)`` and
returns ``(s_hat, ls_hat)``; "Flagged" counts frames with a non-zero residual syndrome, "BLER"
frames with a non-zero ``ls_hat``.  Plotting itself is not part of this package.
"""
import time

import numpy as np

from .feedback_gnn import ErrorIndicator


def count_block_errors(b, b_hat):
    """Number of rows in which ``b`` and ``b_hat`` differ (metrics.py:194-223)."""
    if isinstance(b_hat, ErrorIndicator) and (b is None or _is_zero(b)):
        return b_hat.count_nonzero_rows()
    b = np.asarray(b)
    b_hat = np.asarray(b_hat)
    return int(np.sum(np.any(b != b_hat, axis=-1)))


class BinarySource:
    """Uniform random bits of a given shape (sionna/utils/misc.py ``BinarySource``; the quantum models hold one
    but never call it).  ``source([batch, n]) -> float32 array``."""

    def __init__(self, dtype=np.float32, seed=None, **kwargs):
        self._dtype = dtype
        self._rng = np.random.default_rng(seed)

    def __call__(self, inputs):
        return self._rng.integers(0, 2, size=tuple(int(x) for x in inputs)).astype(self._dtype)


def compute_bler(b, b_hat):
    """Fraction of rows in which ``b`` and ``b_hat`` differ (metrics.py:142-170)."""
    if isinstance(b_hat, ErrorIndicator) and (b is None or _is_zero(b)):
        return b_hat.count_nonzero_rows() / max(b_hat.shape[0], 1)
    b, b_hat = np.asarray(b), np.asarray(b_hat)
    return float(np.mean(np.any(b != b_hat, axis=-1).astype(np.float64)))


def compute_ber(b, b_hat):
    """Fraction of differing entries (metrics.py:98-118)."""
    if isinstance(b_hat, ErrorIndicator) and not b_hat.has_dense() and (b is None or _is_zero(b)):
        return b_hat.count_nonzero_rows() / max(b_hat.shape[0] * b_hat.shape[1], 1)     # one count per frame in error
    return float(np.mean((np.asarray(b) != np.asarray(b_hat)).astype(np.float64)))


def count_errors(b, b_hat):
    """Number of differing entries (metrics.py:172-192)."""
    if isinstance(b_hat, ErrorIndicator) and not b_hat.has_dense() and (b is None or _is_zero(b)):
        return b_hat.count_nonzero_rows()
    return int(np.sum(np.asarray(b) != np.asarray(b_hat)))


class _Zeros:
    """Stand-in for ``tf.zeros_like(x)`` of a lazy indicator."""

    def __init__(self, shape):
        self.shape = shape


def _is_zero(b):
    return isinstance(b, _Zeros) or (isinstance(b, (int, float)) and b == 0)


def zeros_like(x):
    if isinstance(x, ErrorIndicator):
        return _Zeros(x.shape)
    return np.zeros_like(np.asarray(x))


def hard_decisions(llr):
    return (np.asarray(llr) > 0).astype(np.asarray(llr).dtype)


def sim_ber(mc_fun, ebno_dbs, batch_size, max_mc_iter, soft_estimates=False, num_target_bit_errors=None,
            num_target_block_errors=None, early_stop=True, graph_mode=None, verbose=True,
            forward_keyboard_interrupt=True, qldpc=False, dtype=None):
    """Simulate until a target number of errors is reached; returns ``(ber, bler)`` arrays with one
    entry per point of ``ebno_dbs`` (misc.py:403-768).  With ``qldpc=True`` the points are physical
    error rates p, ``ber`` is the flagged rate and ``bler`` the logical error rate."""
    assert isinstance(early_stop, bool), "early_stop must be bool."
    assert isinstance(soft_estimates, bool), "soft_estimates must be bool."
    assert isinstance(verbose, bool), "verbose must be bool."
    if graph_mode not in (None, "default", "graph", "xla"):
        raise TypeError("Unknown graph_mode selected.")
    ebno_dbs = np.atleast_1d(np.asarray(ebno_dbs, dtype=np.float32))
    batch_size = int(np.asarray(batch_size))
    max_mc_iter = int(np.asarray(max_mc_iter))
    num_points = len(ebno_dbs)
    bit_errors = np.zeros(num_points, np.int64)
    block_errors = np.zeros(num_points, np.int64)
    nb_bits = np.zeros(num_points, np.int64)
    nb_blocks = np.zeros(num_points, np.int64)
    status = np.zeros(num_points)
    runtime = np.zeros(num_points)
    status_levels = ["not simulated", "reached max iter       ", "no errors - early stop",
                     "reached target bit errors", "reached target block errors"]
    if qldpc:
        header_text = ["p", "Flagged", "BLER", "flag errors", "block errors", "num blocks", "runtime [s]", "status"]
        fmt = "{: >9} |{: >11} |{: >11} |{: >12} |{: >13} |{: >12} |{: >12} |{: >10}"
    else:
        header_text = ["EbNo [dB]", "BER", "BLER", "bit errors", "num bits", "block errors", "num blocks",
                       "runtime [s]", "status"]
        fmt = "{: >9} |{: >11} |{: >11} |{: >12} |{: >12} |{: >13} |{: >12} |{: >12} |{: >10}"

    def _print_progress(is_final, rt, idx_snr, idx_it, header=None):
        end_str = "\n" if is_final else "\r"
        if header is not None:
            row_text, end_str = header, "\n"
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                ber_np = np.nan_to_num(bit_errors[idx_snr] / nb_bits[idx_snr])
                bler_np = np.nan_to_num(block_errors[idx_snr] / nb_blocks[idx_snr])
            status_txt = (f"iter: {idx_it:.0f}/{max_mc_iter:.0f}" if status[idx_snr] == 0
                          else status_levels[int(status[idx_snr])])
            row_text = [str(np.round(ebno_dbs[idx_snr], 3)), f"{ber_np:.4e}", f"{bler_np:.4e}",
                        bit_errors[idx_snr], nb_bits[idx_snr], block_errors[idx_snr], nb_blocks[idx_snr],
                        np.round(rt, 1), status_txt]
            if qldpc:
                row_text.pop(4)
        print(fmt.format(*row_text), end=end_str)

    i = 0
    try:
        for i in range(num_points):
            runtime[i] = time.perf_counter()
            iter_count = -1
            for ii in range(max_mc_iter):
                iter_count += 1
                outputs = mc_fun(batch_size=batch_size, ebno_db=ebno_dbs[i])
                if qldpc:
                    s_hat, l_hat = outputs[0], outputs[1]
                    bit_e = count_block_errors(zeros_like(s_hat), s_hat)       # flagged errors
                    block_e = count_block_errors(zeros_like(l_hat), l_hat)
                    bit_n = s_hat.shape[0]
                    block_n = l_hat.shape[0]
                elif isinstance(outputs[1], ErrorIndicator) and not outputs[1].has_dense():
                    # OSD / BSC models driven with qldpc=False (OSD.ipynb cells 2-3): the device keeps one
                    # flag per frame, not the dense ls_hat, so a frame in error counts as one "bit" error
                    l_hat = outputs[1]
                    bit_e = block_e = l_hat.count_nonzero_rows()
                    bit_n = l_hat.shape[0] * l_hat.shape[1]
                    block_n = l_hat.shape[0]
                else:
                    b = outputs[0]
                    b = np.zeros(b.shape, np.int64) if isinstance(b, ErrorIndicator) and not b.has_dense() else np.asarray(b)
                    b_hat = np.asarray(outputs[1])
                    if soft_estimates:
                        b_hat = hard_decisions(b_hat)
                    bit_e = int(np.sum(b != b_hat))
                    block_e = count_block_errors(b, b_hat)
                    bit_n = b.size
                    block_n = b[..., -1].size
                bit_errors[i] += bit_e
                block_errors[i] += block_e
                nb_bits[i] += bit_n
                nb_blocks[i] += block_n
                if verbose:
                    if i == 0 and iter_count == 0:
                        _print_progress(True, 0, 0, 0, header=header_text)
                        print('-' * 135)
                    _print_progress(False, time.perf_counter() - runtime[i], i, ii)
                if num_target_bit_errors is not None and bit_errors[i] >= num_target_bit_errors:
                    status[i] = 3
                    runtime[i] = time.perf_counter() - runtime[i]
                    break
                if num_target_block_errors is not None and block_errors[i] >= num_target_block_errors:
                    runtime[i] = time.perf_counter() - runtime[i]
                    status[i] = 4
                    break
                if iter_count == max_mc_iter - 1:
                    runtime[i] = time.perf_counter() - runtime[i]
                    status[i] = 1
            if verbose:
                _print_progress(True, runtime[i], i, iter_count)
            if early_stop and block_errors[i] == 0:
                status[i] = 2
                if verbose:
                    print(f"\nSimulation stopped as no error occurred @ EbNo = {ebno_dbs[i]:.1f} dB.\n")
                break
    except KeyboardInterrupt as e:
        if forward_keyboard_interrupt:
            raise e
        print(f"\nSimulation stopped by the user @ EbNo = {ebno_dbs[i]} dB")
        for idx in range(i + 1, num_points):
            bit_errors[idx] = -1
            block_errors[idx] = -1
            nb_bits[idx] = 1
            nb_blocks[idx] = 1
    with np.errstate(divide="ignore", invalid="ignore"):
        ber = np.nan_to_num(bit_errors.astype(np.float64) / nb_bits.astype(np.float64))
        bler = np.nan_to_num(block_errors.astype(np.float64) / nb_blocks.astype(np.float64))
    sim_ber.last = dict(bit_errors=bit_errors, block_errors=block_errors, nb_bits=nb_bits,
                        nb_blocks=nb_blocks, runtime=runtime, status=status)
    return ber, bler


class PlotBER:
    """Stores simulated curves (plotting.py:148-447); ``simulate`` forwards to ``sim_ber``.
    Figures are not drawn by this package (``show_fig`` is accepted and ignored)."""

    def __init__(self, title="Bit/Block Error Rate"):
        assert isinstance(title, str), "title must be str."
        self._title = title
        self._bers = []
        self._snrs = []
        self._legends = []
        self._is_bler = []

    @property
    def title(self):
        return self._title

    @property
    def ber(self):
        return self._bers

    @property
    def snr(self):
        return self._snrs

    @property
    def legend(self):
        return self._legends

    @property
    def is_bler(self):
        return self._is_bler

    def simulate(self, mc_fun, ebno_dbs, batch_size, max_mc_iter, legend="", add_ber=True, add_bler=False,
                 soft_estimates=False, num_target_bit_errors=None, num_target_block_errors=None,
                 early_stop=True, graph_mode=None, add_results=True, forward_keyboard_interrupt=True,
                 show_fig=True, qldpc=False, verbose=True):
        ber, bler = sim_ber(mc_fun, ebno_dbs, batch_size, soft_estimates=soft_estimates,
                            max_mc_iter=max_mc_iter, num_target_bit_errors=num_target_bit_errors,
                            num_target_block_errors=num_target_block_errors, early_stop=early_stop,
                            graph_mode=graph_mode, verbose=verbose, qldpc=qldpc,
                            forward_keyboard_interrupt=forward_keyboard_interrupt)
        if add_ber:
            self._bers += [ber]
            self._snrs += [ebno_dbs]
            self._legends += [legend]
            self._is_bler += [False]
        if add_bler:
            self._bers += [bler]
            self._snrs += [ebno_dbs]
            self._legends += [legend + " (BLER)"]
            self._is_bler += [True]
        if add_results is False:
            if add_bler:
                self.remove(-1)
            if add_ber:
                self.remove(-1)
        return ber, bler

    def add(self, ebno_db, ber, is_bler=False, legend=""):
        self._bers += [ber]
        self._snrs += [ebno_db]
        self._legends += [legend]
        self._is_bler += [is_bler]

    def reset(self):
        self._bers, self._snrs, self._legends, self._is_bler = [], [], [], []

    def remove(self, idx=-1):
        del self._bers[idx]
        del self._snrs[idx]
        del self._legends[idx]
        del self._is_bler[idx]
