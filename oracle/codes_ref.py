"""ORACLE (test infrastructure only): slow, literal restatement of the reference's GF(2)
elimination and CSS-code bookkeeping, used to check ``fbgnn.gf2`` / ``fbgnn.codes_q``.

Follows ``sionna/fec/utils.py:1022-1146`` (``row_echelon``, ``kernel``) and
``sionna/fec/ldpc/codes_q.py:10-78`` (``css_code.__init__``, ``compute_logicals``) step by
step with dense boolean matrices and Python loops -- no bit packing, no vectorised row
updates -- so that it shares no code path with the product implementation.
"""
import numpy as np


def row_echelon_ref(mat, reduced=False):
    """utils.py:1022-1087."""
    mat = np.array(mat).astype(bool)
    m, n = mat.shape
    transform = np.identity(m).astype(bool)
    pivot_row = 0
    pivot_cols = []
    for col in range(n):
        if not mat[pivot_row, col]:
            swap = pivot_row + int(np.argmax(mat[pivot_row:m, col]))
            if mat[swap, col]:
                mat[[swap, pivot_row]] = mat[[pivot_row, swap]]
                transform[[swap, pivot_row]] = transform[[pivot_row, swap]]
        if mat[pivot_row, col]:
            rng = range(pivot_row + 1, m) if not reduced else [k for k in range(m) if k != pivot_row]
            for r in rng:
                if mat[r, col]:
                    mat[r] ^= mat[pivot_row]
                    transform[r] ^= transform[pivot_row]
            pivot_row += 1
            pivot_cols.append(col)
        if pivot_row >= m:
            break
    return [mat.astype(int), pivot_row, transform.astype(int), pivot_cols]


def kernel_ref(mat):
    """utils.py:1104-1146."""
    t = np.asarray(mat).T
    m, _ = t.shape
    _, rank, transform, pivot_cols = row_echelon_ref(t)
    return transform[rank:m], rank, pivot_cols


def css_ref(hx, hz):
    """codes_q.py:35-40,63-78: returns a dict with the derived matrices."""
    hx = np.asarray(hx)
    hz = np.asarray(hz)
    hx_perp, rank_hx, pivot_hx = kernel_ref(hx)
    hz_perp, rank_hz, pivot_hz = kernel_ref(hz)
    hx_basis = hx[pivot_hx]
    hz_basis = hz[pivot_hz]

    def compute_lz(ker_hx, im_hzT):
        log_stack = np.vstack([im_hzT, ker_hx])
        pivots = row_echelon_ref(log_stack.T)[3]
        idx = [i for i in range(im_hzT.shape[0], log_stack.shape[0]) if i in pivots]
        return log_stack[idx]

    return dict(hx_perp=hx_perp, hz_perp=hz_perp, rank_hx=rank_hx, rank_hz=rank_hz,
                pivot_hx=pivot_hx, pivot_hz=pivot_hz, hx_basis=hx_basis, hz_basis=hz_basis,
                lx=compute_lz(hz_perp, hx_basis), lz=compute_lz(hx_perp, hz_basis),
                N=hx.shape[1], K=hx.shape[1] - rank_hx - rank_hz)
