"""CSS code construction for the decoders (host side, runs once per code).

Call surface of the reference's ``sionna/fec/ldpc/codes_q.py`` -- ``css_code`` with the attributes
``hx hz lx lz hx_perp hz_perp hx_basis hz_basis pivot_hx pivot_hz rank_hx rank_hz N K D L Q name``
(``codes_q.py:12-49``) and the constructors the scripts and notebooks call (``:84-280``) -- built
differently: every constructor assembles the **edge list** of the Tanner graph directly from the code's
algebraic structure (circulant shifts, Kronecker factors, lattice plaquettes) as integer index arithmetic
on whole arrays, and only then materialises the dense 0/1 matrices the reference API exposes.  Nothing is
filled entry by entry.  Two things ride along for the device side:

* ``code.csr_x`` / ``code.csr_z``: int32 CSR of ``hx`` / ``hz`` (what ``fbgnn_code_create`` takes);
* ``code.qc``: for quasi-cyclic constructions, the lifted description ``{"l", "x": [(block_row, block_col,
  shift)...], "z": [...]}`` -- every non-zero ``l x l`` block of ``hx`` / ``hz`` is the circulant
  permutation ``P^shift`` (row ``r`` has its one in column ``(r - shift) mod l``).  It is host-side metadata (tests rebuild
  the dense matrices from it); the kernels index through the CSR, see DESIGN.md on SURVEY.md H4.

The GF(2) algebra (kernel, pivots, logical operators) is ``fbgnn.gf2``: bit-packed, same pivot order as
the reference, hence identical ``hx_perp`` / ``lx`` / ``lz``.
"""
from functools import reduce

import numpy as np

from .gf2 import compute_code_distance, int2bin, inverse, kernel, rank, row_echelon  # noqa: F401  (re-exported)


# ------------------------------------------------------------------ edge-list algebra ----
class EdgeMatrix:
    """Binary matrix as a set of (row, col) index pairs.  All constructors below compose these."""

    __slots__ = ("shape", "rows", "cols")

    def __init__(self, shape, rows, cols):
        self.shape = (int(shape[0]), int(shape[1]))
        self.rows = np.asarray(rows, dtype=np.int64).ravel()
        self.cols = np.asarray(cols, dtype=np.int64).ravel()

    # -- sources ---------------------------------------------------------------------------
    @staticmethod
    def from_dense(mat):
        mat = np.asarray(mat)
        r, c = np.nonzero(mat)
        return EdgeMatrix(mat.shape, r, c)

    @staticmethod
    def eye(n):
        i = np.arange(n)
        return EdgeMatrix((n, n), i, i)

    @staticmethod
    def zeros(m, n):
        return EdgeMatrix((m, n), [], [])

    @staticmethod
    def circulant(l, powers):
        """Sum of cyclic shifts: column ``i`` carries ones in rows ``(i + c) mod l`` for ``c`` in ``powers``."""
        i = np.arange(l)[:, None]
        c = np.asarray(list(powers), dtype=np.int64)[None, :]
        return EdgeMatrix((l, l), (i + c) % l, np.broadcast_to(i, (l, c.shape[1])))

    # -- algebra ---------------------------------------------------------------------------
    @property
    def T(self):
        return EdgeMatrix(self.shape[::-1], self.cols, self.rows)

    def kron(self, other):
        """Kronecker product: entry (a, b) of self times entry (c, d) of other lands at (a*m2 + c, b*n2 + d)."""
        m2, n2 = other.shape
        rows = (self.rows[:, None] * m2 + other.rows[None, :])
        cols = (self.cols[:, None] * n2 + other.cols[None, :])
        return EdgeMatrix((self.shape[0] * m2, self.shape[1] * n2), rows, cols)

    @staticmethod
    def hstack(parts):
        offs = np.cumsum([0] + [p.shape[1] for p in parts])
        m = parts[0].shape[0]
        assert all(p.shape[0] == m for p in parts)
        return EdgeMatrix((m, offs[-1]), np.concatenate([p.rows for p in parts]),
                          np.concatenate([p.cols + o for p, o in zip(parts, offs)]))

    @staticmethod
    def block(grid):
        """2-D grid of equally sized blocks (``None`` = zero block)."""
        bm, bn = next(b.shape for row in grid for b in row if b is not None)
        rows, cols = [], []
        for i, row in enumerate(grid):
            for j, b in enumerate(row):
                if b is not None:
                    rows.append(b.rows + i * bm)
                    cols.append(b.cols + j * bn)
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
        return EdgeMatrix((len(grid) * bm, len(grid[0]) * bn), cat(rows), cat(cols))

    # -- sinks -----------------------------------------------------------------------------
    def toarray(self, dtype=int):
        """Dense matrix; coinciding pairs add up (as a sum of permutation matrices would)."""
        out = np.zeros(self.shape, dtype=dtype)
        np.add.at(out, (self.rows, self.cols), 1)
        return out

    def csr(self):
        return csr_of(self.toarray() & 1)


def csr_of(mat):
    """(indptr, indices) int32 CSR of a 0/1 matrix, column indices increasing within a row."""
    mat = np.asarray(mat)
    r, c = np.nonzero(mat)
    indptr = np.zeros(mat.shape[0] + 1, np.int32)
    np.cumsum(np.bincount(r, minlength=mat.shape[0]), out=indptr[1:])
    return indptr, np.ascontiguousarray(c, dtype=np.int32)


def _dense(x):
    if isinstance(x, EdgeMatrix):
        return x.toarray()
    if hasattr(x, "toarray"):
        return np.asarray(x.toarray())
    return np.asarray(x)


# ------------------------------------------------------------------ the code object ------
class css_code:
    """CSS code given by two parity-check matrices with ``hx @ hz.T = 0 (mod 2)``.

    ``lz`` spans ``ker(hx) / rowspace(hz)`` and ``lx`` spans ``ker(hz) / rowspace(hx)``; ``h*_perp`` are the
    kernel bases the elimination yields; ``h*_basis = h*[pivot_h*]`` are row bases.  ``D`` is the smallest row
    weight among the kernel bases unless ``code_distance`` is given (an upper-bound style figure the reference
    prints, not the true distance)."""

    def __init__(self, hx=np.array([[]]), hz=np.array([[]]), code_distance=np.nan, name=None, name_prefix="",
                 check_css=False, qc=None):
        self.hx, self.hz = _dense(hx), _dense(hz)
        n_x, n_z = self.hx.shape[1], self.hz.shape[1]
        assert n_x == n_z, "hx and hz should have equal number of columns!"
        assert n_x != 0, "number of variable nodes should not be zero!"
        if check_css:
            assert not np.any(self.hx @ self.hz.T % 2), "CSS constraint not satisfied"
        self.N = n_x
        self.qc = qc
        self.csr_x, self.csr_z = csr_of(self.hx), csr_of(self.hz)

        # one elimination per side gives the kernel, the rank and a set of independent rows
        self.hx_perp, self.rank_hx, self.pivot_hx = kernel(self.hx)
        self.hz_perp, self.rank_hz, self.pivot_hz = kernel(self.hz)
        self.hx_basis, self.hz_basis = self.hx[self.pivot_hx], self.hz[self.pivot_hz]
        self.K = self.N - self.rank_hx - self.rank_hz

        self.lx = self.lz = np.array([[]])
        self.L = self.Q = np.nan
        self.compute_ldpc_params()
        self.compute_logicals()
        self.D = code_distance
        if code_distance is np.nan:
            self.D = np.min([compute_code_distance(perp, is_pcm=False, is_basis=True)
                             for perp in (self.hx_perp, self.hz_perp)])
        self.name = name if name is not None else f"{name_prefix}_n{self.N}_k{self.K}"

    def compute_ldpc_params(self):
        """``L`` / ``Q``: largest column / row weight over both matrices."""
        self.L = np.max([h.sum(axis=0).max() for h in (self.hx, self.hz)]).astype(int)
        self.Q = np.max([h.sum(axis=1).max() for h in (self.hx, self.hz)]).astype(int)

    @staticmethod
    def _quotient_basis(kernel_rows, stabiliser_rows):
        """Rows of ``kernel_rows`` that extend ``stabiliser_rows`` to a basis of their joint span: greedy, in
        order -- exactly the rows an elimination of the stacked matrix's transpose marks as pivots."""
        stacked = np.vstack([stabiliser_rows, kernel_rows])
        independent = row_echelon(stacked.T, want_transform=False)[3]
        first = stabiliser_rows.shape[0]
        return stacked[[i for i in independent if i >= first]]

    def compute_logicals(self):
        self.lx = self._quotient_basis(self.hz_perp, self.hx_basis)
        self.lz = self._quotient_basis(self.hx_perp, self.hz_basis)
        return self.lx, self.lz

    def canonical_logicals(self):
        """Re-mix ``lx`` so that ``lx @ lz.T = I``."""
        self.lx = inverse(self.lx @ self.lz.T % 2) @ self.lx % 2


# ------------------------------------------------------------------ classical ingredients -
def create_circulant_matrix(l, pows):
    return EdgeMatrix.circulant(l, pows).toarray() & 1


def hamming_code(rank):
    """``rank x (2^rank - 1)`` matrix whose columns count 1, 2, ... in binary, most significant bit on top."""
    rank = int(rank)
    values = np.arange(1, 2 ** rank)
    return ((values[None, :] >> np.arange(rank - 1, -1, -1)[:, None]) & 1).astype(int)


def rep_code(d):
    """``(d-1) x d`` repetition-code checks: row ``i`` couples bits ``i`` and ``i+1``."""
    i = np.arange(d - 1)
    return EdgeMatrix((d - 1, d), np.r_[i, i], np.r_[i, i + 1]).toarray()


# ------------------------------------------------------------------ product constructions -
def create_generalized_bicycle_codes(l, a, b, name=None):
    """``hx = [A | B]``, ``hz = [B^T | A^T]`` with circulants ``A``, ``B``."""
    A, B = EdgeMatrix.circulant(l, a), EdgeMatrix.circulant(l, b)
    # P^s has its one of row r in column (r - s) mod l; the transpose is P^{-s}
    qc = dict(l=l, x=[(0, 0, s % l) for s in a] + [(0, 1, s % l) for s in b],
              z=[(0, 0, (-s) % l) for s in b] + [(0, 1, (-s) % l) for s in a])
    return css_code(EdgeMatrix.hstack([A, B]), EdgeMatrix.hstack([B.T, A.T]), name=name, name_prefix="GB", qc=qc)


def hypergraph_product(h1, h2, name=None):
    """``hx = [h1 (x) I | I (x) h2^T]``, ``hz = [I (x) h2 | h1^T (x) I]``."""
    h1, h2 = EdgeMatrix.from_dense(_dense(h1)), EdgeMatrix.from_dense(_dense(h2))
    (m1, n1), (m2, n2) = h1.shape, h2.shape
    eye = EdgeMatrix.eye
    hx = EdgeMatrix.hstack([h1.kron(eye(n2)), eye(m1).kron(h2.T)])
    hz = EdgeMatrix.hstack([eye(n1).kron(h2), h1.T.kron(eye(m2))])
    return css_code(hx, hz, name=name, name_prefix="HP")


def create_surface_codes(n):
    """The [[n^2 + (n-1)^2, 1, n]] surface code as the product of two repetition codes."""
    h = rep_code(n)
    return hypergraph_product(h, h, f"Surface_n{n ** 2 + (n - 1) ** 2}_k{1}_d{n}")


def create_QC_GHP_codes(l, a, b, name=None):
    """Quasi-cyclic generalised hypergraph product: ``a`` is an ``m x n`` array of shifts (negative = zero block),
    ``A`` its lift by ``l x l`` circulant permutations, ``B`` the circulant of ``b``;
    ``hx = [A | I_m (x) B]``, ``hz = [I_n (x) B^T | A^T]``."""
    a = np.asarray(a)
    m, n = a.shape
    A = EdgeMatrix.block([[EdgeMatrix.circulant(l, [s]) if s >= 0 else None for s in row] for row in a])
    B = EdgeMatrix.circulant(l, b)
    hx = EdgeMatrix.hstack([A, EdgeMatrix.eye(m).kron(B)])
    hz = EdgeMatrix.hstack([EdgeMatrix.eye(n).kron(B.T), A.T])
    nz = [(i, j, int(a[i, j])) for i in range(m) for j in range(n) if a[i, j] >= 0]
    qc = dict(l=l,
              x=[(i, j, s % l) for i, j, s in nz] + [(i, n + i, s % l) for i in range(m) for s in b],
              z=[(j, j, (-s) % l) for j in range(n) for s in b] + [(j, n + i, (-s) % l) for i, j, s in nz])
    return css_code(hx, hz, name=name, name_prefix="GHP", qc=qc)


def create_cyclic_permuting_matrix(n, shifts):
    """``n x n`` array of shifts: ``shifts[i]`` sits on the ``i``-th sub-diagonal (cyclically), ``-1`` elsewhere."""
    A = np.full((n, n), -1, dtype=int)
    j = np.arange(n)
    for i, s in enumerate(shifts):
        A[j, (j - i) % n] = s
    return A


def create_bivariate_QC_codes(l, m, A_x_pows, A_y_pows, B_x_pows, B_y_pows, name=None):
    """Bivariate bicycle codes: ``x = S_l (x) I_m``, ``y = I_l (x) S_m`` with ``S`` the cyclic shift
    ``circulant([-1])``; ``A`` / ``B`` sums of monomials; ``hx = [A | B]``, ``hz = [B^T | A^T]``."""
    def monomial(px, py):                     # x^px y^py = S_l^px (x) S_m^py, S^p = circulant([-p])
        return EdgeMatrix.circulant(l, [-px]).kron(EdgeMatrix.circulant(m, [-py]))

    def poly(x_pows, y_pows):
        terms = [monomial(p, 0) for p in x_pows] + [monomial(0, p) for p in y_pows]
        return reduce(lambda u, v: EdgeMatrix(u.shape, np.r_[u.rows, v.rows], np.r_[u.cols, v.cols]), terms)

    A, B = poly(A_x_pows, A_y_pows), poly(B_x_pows, B_y_pows)
    hx = np.hstack([A.toarray(), B.toarray()])
    hz = np.hstack([B.T.toarray(), A.T.toarray()])
    return css_code(hx, hz, name=name, name_prefix="IBM")


# ------------------------------------------------------------------ lattice codes ---------
def _plaquettes(n, cells, periodic):
    """Weight-4 checks on the ``n x n`` qubit lattice, one per cell ``(i, j)``: the cell's four corners."""
    cells = np.asarray(cells, dtype=np.int64).reshape(-1, 2)
    i, j = cells[:, 0], cells[:, 1]
    i1, j1 = ((i + 1) % n, (j + 1) % n) if periodic else (i + 1, j + 1)
    cols = np.stack([i * n + j, i1 * n + j1, i1 * n + j, i * n + j1], axis=1)
    rows = np.repeat(np.arange(len(cells)), 4)
    return rows, cols.ravel()


def _checks(n_qubits, *groups):
    """Stack groups of (rows, cols) check descriptions into one matrix."""
    rows, cols, base = [], [], 0
    for r, c in groups:
        r = np.asarray(r, dtype=np.int64)
        rows.append(r + base)
        cols.append(np.asarray(c, dtype=np.int64))
        base += (int(r.max()) + 1) if r.size else 0
    return EdgeMatrix((base, n_qubits), np.concatenate(rows), np.concatenate(cols)).toarray()


def _pairs(first, second):
    k = len(first)
    return np.repeat(np.arange(k), 2), np.stack([first, second], axis=1).ravel()


def create_rotated_surface_codes(n, name=None):
    """Rotated surface code on ``n x n`` qubits (``n`` odd): bulk plaquettes alternate Z / X in a checkerboard,
    weight-2 X checks close the top (even columns) and bottom (odd columns) edges, weight-2 Z checks the right
    (even rows) and left (odd rows) edges."""
    assert n % 2 == 1, "n should be odd"
    ii, jj = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel()], axis=1)
    even = (cells.sum(axis=1) % 2) == 0
    t = np.arange(n - 1)
    top_or_bottom = np.where(t % 2 == 0, t, (n - 1) * n + t)             # first qubit of the horizontal pair
    right_or_left = np.where(t % 2 == 0, t * n + (n - 1), t * n)         # first qubit of the vertical pair
    hx = _checks(n * n, _plaquettes(n, cells[~even], False), _pairs(top_or_bottom, top_or_bottom + 1))
    hz = _checks(n * n, _plaquettes(n, cells[even], False), _pairs(right_or_left, right_or_left + n))
    return css_code(hx, hz, name=name, name_prefix="Rotated_Surface")


def create_checkerboard_toric_codes(n, name=None):
    """Toric code on the ``n x n`` periodic lattice (``n`` even), Z checks on even cells, X checks on odd ones."""
    assert n % 2 == 0, "n should be even"
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel()], axis=1)
    even = (cells.sum(axis=1) % 2) == 0
    hx = _checks(n * n, _plaquettes(n, cells[~even], True))
    hz = _checks(n * n, _plaquettes(n, cells[even], True))
    return css_code(hx, hz, name=name, name_prefix="Toric")


# ------------------------------------------------------------------ A-list files ----------
def alistToNumpy(lines):
    """A-list (MacKay) rows of integers -> 0/1 matrix.  Line 0 is ``n_cols n_rows``; per-column / per-row degree
    lines may or may not be present; then one line per column with the 1-based row indices, zero padded."""
    n_cols, n_rows = lines[0]
    has_degree_lines = len(lines[2]) == n_cols and len(lines[3]) == n_rows
    column_lists = lines[4 if has_degree_lines else 2:][:n_cols]
    cols = np.repeat(np.arange(n_cols), [len(c) for c in column_lists])
    rows = np.fromiter((r for c in column_lists for r in c), dtype=np.int64, count=len(cols))
    keep = rows != 0
    matrix = np.zeros((n_rows, n_cols), dtype=float)
    matrix[rows[keep] - 1, cols[keep]] = 1
    return matrix


def readAlist(directory):
    """Parity-check matrix stored as an A-list text file (used for the over-complete check matrices)."""
    with open(directory, "r") as f:
        lines = [[int(tok) for tok in line.split()] for line in f]
    return alistToNumpy(lines).astype(int)
