"""GF(2) linear algebra used by the code constructors (host side, runs once per code).

Same functions, argument meaning and return values as the reference's
``sionna/fec/utils.py:1022-1228`` (``row_echelon``, ``rank``, ``kernel``, ``row_basis``,
``compute_code_distance``, ``inverse``), ``int2bin`` (``utils.py:714-741``) and
``int_mod_2`` (``utils.py:1565-1582``).  The elimination visits columns and picks pivots in
exactly the reference's order (first row at or below the pivot row holding a one, no
column swaps), so ``kernel`` returns the same basis and the logical operators derived
from it are identical.  Rows are bit-packed and the row updates are vectorised, which
turns the reference's ~2 s Python row loops for [[1270,28]] into tens of milliseconds.
"""
import numpy as np


def int2bin(num, len_):
    """``num`` as a list of ``len_`` bits, most significant first (utils.py:714)."""
    assert num >= 0, "Input integer should be non-negative"
    assert len_ >= 0, "width should be non-negative"
    bin_ = format(num, f'0{len_}b')
    return [int(x) for x in bin_[-len_:]] if len_ else []


def int_mod_2(x):
    """``x mod 2`` for integer (or integer-valued) arrays, keeping the dtype (utils.py:1565)."""
    x = np.asarray(x)
    return (x.astype(np.int32) & 1).astype(x.dtype)


def _pack(mat_bool):
    return np.packbits(mat_bool, axis=1)


def _unpack(packed, ncols):
    return np.unpackbits(packed, axis=1, count=ncols)


def row_echelon(mat, reduced=False, want_transform=True):
    """Row echelon form over GF(2) without column swaps (utils.py:1022-1087).

    Returns ``[row_ech_form, rank, transform, pivot_cols]`` with
    ``transform @ mat = row_ech_form (mod 2)``.  ``want_transform=False`` skips the
    transform (returned as ``None``) when only rank / pivots are needed.
    """
    mat = np.asarray(mat)
    m, n = mat.shape
    a = mat.astype(bool)
    if want_transform:
        a = np.concatenate([a, np.identity(m, dtype=bool)], axis=1)
    p = _pack(a)                      # [m, ceil(width/8)] uint8, MSB-first bit order
    pivot_row = 0
    pivot_cols = []
    for col in range(n):
        if pivot_row >= m:
            break
        byte, bit = col >> 3, 7 - (col & 7)
        colbits = (p[:, byte] >> bit) & 1
        if not colbits[pivot_row]:
            below = np.flatnonzero(colbits[pivot_row:])
            if below.size == 0:
                continue
            swap = pivot_row + int(below[0])
            p[[swap, pivot_row]] = p[[pivot_row, swap]]
            colbits[[swap, pivot_row]] = colbits[[pivot_row, swap]]
        colbits[pivot_row] = 0
        if not reduced:
            colbits[:pivot_row] = 0
        rows = np.flatnonzero(colbits)
        if rows.size:
            p[rows] ^= p[pivot_row]
        pivot_row += 1
        pivot_cols.append(col)
    rank_ = pivot_row
    width = n + m if want_transform else n
    full = _unpack(p, width).astype(int)
    row_ech_form = full[:, :n]
    transform = full[:, n:] if want_transform else None
    return [row_ech_form, rank_, transform, pivot_cols]


def rank(mat):
    """Rank over GF(2) (utils.py:1089-1102)."""
    return row_echelon(mat, want_transform=False)[1]


def kernel(mat):
    """Basis of ``{x : mat @ x = 0}`` plus rank and pivot columns of ``mat.T`` (utils.py:1104-1146)."""
    transpose = np.asarray(mat).T
    m, _ = transpose.shape
    _, rank_, transform, pivot_cols = row_echelon(transpose)
    ker = transform[rank_:m]
    return ker, rank_, pivot_cols


def row_basis(mat):
    """A basis of the row space made of rows of ``mat`` (utils.py:1148-1161)."""
    mat = np.asarray(mat)
    return mat[row_echelon(mat.T, want_transform=False)[3]]


def compute_code_distance(mat, is_pcm=True, is_basis=False):
    """Minimum row weight of a (basis of the) generator matrix (utils.py:1163-1194)."""
    gen = mat
    if is_pcm:
        gen = kernel(mat)[0]
    if len(gen) == 0:
        return np.inf
    cw = gen
    if not is_basis:
        cw = row_basis(gen)
    return np.min(np.sum(cw, axis=1))


def inverse(mat):
    """(Left) inverse of a full-column-rank binary matrix (utils.py:1196-1228)."""
    mat = np.asarray(mat)
    m, n = mat.shape
    reduced_row_ech, rank_, transform, _ = row_echelon(mat, reduced=True)
    if m == n and rank_ == m:
        return transform
    elif m > rank_ and n == rank_:
        return reduced_row_ech.T @ transform % 2
    else:
        raise ValueError("This matrix is not invertible. Please provide either a full-rank square"
                         " matrix or a rectangular matrix with full column rank.")
