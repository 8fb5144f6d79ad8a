"""CSS quantum LDPC code construction (host side, numpy).

Same public surface as the reference's ``sionna/fec/ldpc/codes_q.py``: the ``css_code``
class with attributes ``hx, hz, lx, lz, hx_perp, hz_perp, hx_basis, hz_basis, pivot_hx,
pivot_hz, rank_hx, rank_hz, N, K, D, L, Q, name`` (codes_q.py:8-49) and the constructors
``create_circulant_matrix`` (84-89), ``create_generalized_bicycle_codes`` (92-97),
``hypergraph_product`` (100-125), ``hamming_code`` (127-133), ``rep_code`` (135-140),
``create_surface_codes`` (142-145), ``create_rotated_surface_codes`` (152-186),
``create_checkerboard_toric_codes`` (188-206), ``create_QC_GHP_codes`` (208-227),
``create_cyclic_permuting_matrix`` (229-234), ``create_bivariate_QC_codes`` (236-247),
``readAlist`` / ``alistToNumpy`` (250-280).
"""
import numpy as np

from .gf2 import row_echelon, rank, kernel, compute_code_distance, inverse, int2bin


class css_code():
    """A CSS code given by its X and Z parity-check matrices (codes_q.py:8-82)."""

    def __init__(self, hx=np.array([[]]), hz=np.array([[]]), code_distance=np.nan, name=None,
                 name_prefix="", check_css=False):
        self.hx = hx
        self.hz = hz
        self.lx = np.array([[]])
        self.lz = np.array([[]])
        self.N = np.nan
        self.K = np.nan
        self.D = code_distance
        self.L = np.nan
        self.Q = np.nan

        _, nx = self.hx.shape
        _, nz = self.hz.shape
        assert nx == nz, "hx and hz should have equal number of columns!"
        assert nx != 0, "number of variable nodes should not be zero!"
        if check_css:
            assert not np.any(hx @ hz.T % 2), "CSS constraint not satisfied"

        self.N = nx
        self.hx_perp, self.rank_hx, self.pivot_hx = kernel(hx)
        self.hz_perp, self.rank_hz, self.pivot_hz = kernel(hz)
        self.hx_basis = self.hx[self.pivot_hx]
        self.hz_basis = self.hz[self.pivot_hz]
        self.K = self.N - self.rank_hx - self.rank_hz

        self.compute_ldpc_params()
        self.compute_logicals()
        if code_distance is np.nan:
            dx = compute_code_distance(self.hx_perp, is_pcm=False, is_basis=True)
            dz = compute_code_distance(self.hz_perp, is_pcm=False, is_basis=True)
            self.D = np.min([dx, dz])   # distance of the stabilizers, not of the code

        self.name = f"{name_prefix}_n{self.N}_k{self.K}" if name is None else name

    def compute_ldpc_params(self):
        hx_l = np.max(np.sum(self.hx, axis=0))
        hz_l = np.max(np.sum(self.hz, axis=0))
        self.L = np.max([hx_l, hz_l]).astype(int)
        hx_q = np.max(np.sum(self.hx, axis=1))
        hz_q = np.max(np.sum(self.hz, axis=1))
        self.Q = np.max([hx_q, hz_q]).astype(int)

    def compute_logicals(self):
        def compute_lz(ker_hx, im_hzT):
            # vectors of ker(hx) that are not in the row space of hz
            log_stack = np.vstack([im_hzT, ker_hx])
            pivots = set(row_echelon(log_stack.T, want_transform=False)[3])
            idx = [i for i in range(im_hzT.shape[0], log_stack.shape[0]) if i in pivots]
            return log_stack[idx]

        self.lx = compute_lz(self.hz_perp, self.hx_basis)
        self.lz = compute_lz(self.hx_perp, self.hz_basis)
        return self.lx, self.lz

    def canonical_logicals(self):
        temp = inverse(self.lx @ self.lz.T % 2)
        self.lx = temp @ self.lx % 2


def create_circulant_matrix(l, pows):
    h = np.zeros((l, l), dtype=int)
    cols = np.arange(l)
    for c in pows:
        h[(cols + c) % l, cols] = 1
    return h


def create_generalized_bicycle_codes(l, a, b, name=None):
    A = create_circulant_matrix(l, a)
    B = create_circulant_matrix(l, b)
    hx = np.hstack((A, B))
    hz = np.hstack((B.T, A.T))
    return css_code(hx, hz, name=name, name_prefix="GB")


def hypergraph_product(h1, h2, name=None):
    h1 = np.asarray(h1).astype(int)
    h2 = np.asarray(h2).astype(int)
    m1, n1 = h1.shape
    m2, n2 = h2.shape
    hx = np.hstack([np.kron(h1, np.identity(n2, dtype=int)),
                    np.kron(np.identity(m1, dtype=int), h2.T)])
    hz = np.hstack([np.kron(np.identity(n1, dtype=int), h2),
                    np.kron(h1.T, np.identity(m2, dtype=int))])
    return css_code(hx, hz, name=name, name_prefix="HP")


def hamming_code(rank):
    rank = int(rank)
    num_rows = (2 ** rank) - 1
    pcm = np.zeros((num_rows, rank), dtype=int)
    for i in range(num_rows):
        pcm[i] = int2bin(i + 1, rank)
    return pcm.T


def rep_code(d):
    pcm = np.zeros((d - 1, d), dtype=int)
    idx = np.arange(d - 1)
    pcm[idx, idx] = 1
    pcm[idx, idx + 1] = 1
    return pcm


def create_surface_codes(n):
    h = rep_code(n)
    return hypergraph_product(h, h, f"Surface_n{n**2 + (n-1)**2}_k{1}_d{n}")


def set_pcm_row(n, pcm, row_idx, i, j):
    i1, j1 = (i + 1) % n, (j + 1) % n
    pcm[row_idx][i * n + j] = pcm[row_idx][i1 * n + j1] = 1
    pcm[row_idx][i1 * n + j] = pcm[row_idx][i * n + j1] = 1


def create_rotated_surface_codes(n, name=None):
    assert n % 2 == 1, "n should be odd"
    n2 = n * n
    m = (n2 - 1) // 2
    hx = np.zeros((m, n2), dtype=int)
    hz = np.zeros((m, n2), dtype=int)
    x_idx = 0
    z_idx = 0
    for i in range(n - 1):
        for j in range(n - 1):
            if (i + j) % 2 == 0:
                set_pcm_row(n, hz, z_idx, i, j)
                z_idx += 1
            else:
                set_pcm_row(n, hx, x_idx, i, j)
                x_idx += 1
    for j in range(n - 1):          # weight-2 X checks on the upper / lower edge
        if j % 2 == 0:
            hx[x_idx][j] = hx[x_idx][j + 1] = 1
        else:
            hx[x_idx][(n - 1) * n + j] = hx[x_idx][(n - 1) * n + (j + 1)] = 1
        x_idx += 1
    for i in range(n - 1):          # weight-2 Z checks on the right / left edge
        if i % 2 == 0:
            hz[z_idx][i * n + (n - 1)] = hz[z_idx][(i + 1) * n + (n - 1)] = 1
        else:
            hz[z_idx][i * n] = hz[z_idx][(i + 1) * n] = 1
        z_idx += 1
    return css_code(hx, hz, name=name, name_prefix="Rotated_Surface")


def create_checkerboard_toric_codes(n, name=None):
    assert n % 2 == 0, "n should be even"
    n2 = n * n
    m = n2 // 2
    hx = np.zeros((m, n2), dtype=int)
    hz = np.zeros((m, n2), dtype=int)
    x_idx = 0
    z_idx = 0
    for i in range(n):
        for j in range(n):
            if (i + j) % 2 == 0:
                set_pcm_row(n, hz, z_idx, i, j)
                z_idx += 1
            else:
                set_pcm_row(n, hx, x_idx, i, j)
                x_idx += 1
    return css_code(hx, hz, name=name, name_prefix="Toric")


def create_QC_GHP_codes(l, a, b, name=None):
    """Quasi-cyclic generalized hypergraph product code (codes_q.py:208-227): ``a`` holds
    circulant shifts (-1 = zero block), ``b`` the exponents of the second circulant."""
    a = np.asarray(a)
    m, n = a.shape
    A = np.zeros((m * l, n * l), dtype=int)
    for i in range(m):
        for j in range(n):
            if a[i, j] >= 0:
                A[i * l:(i + 1) * l, j * l:(j + 1) * l] = create_circulant_matrix(l, [a[i, j]])
    temp_b = create_circulant_matrix(l, b)
    B = np.kron(np.identity(m, dtype=int), temp_b)
    hx = np.hstack((A, B))
    B_T = np.kron(np.identity(n, dtype=int), temp_b.T)
    hz = np.hstack((B_T, A.T))
    return css_code(hx, hz, name=name, name_prefix="GHP")


def create_cyclic_permuting_matrix(n, shifts):
    A = np.full((n, n), -1, dtype=int)
    for i, s in enumerate(shifts):
        for j in range(n):
            A[j, (j - i) % n] = s
    return A


def create_bivariate_QC_codes(l, m, A_x_pows, A_y_pows, B_x_pows, B_y_pows, name=None):
    """IBM's bivariate bicycle codes (codes_q.py:236-247): x = S_l (x) I_m, y = I_l (x) S_m."""
    S_l = create_circulant_matrix(l, [-1])
    S_m = create_circulant_matrix(m, [-1])
    x = np.kron(S_l, np.identity(m, dtype=int))
    y = np.kron(np.identity(l, dtype=int), S_m)

    def mpow(mat, p):
        return np.linalg.matrix_power(mat, int(p))

    A = sum([mpow(x, p) for p in A_x_pows] + [mpow(y, p) for p in A_y_pows])
    B = sum([mpow(x, p) for p in B_x_pows] + [mpow(y, p) for p in B_y_pows])
    hx = np.hstack((A, B))
    hz = np.hstack((B.T, A.T))
    return css_code(hx, hz, name=name, name_prefix="IBM")


def readAlist(directory):
    """Read a parity-check matrix in A-list format (codes_q.py:250-265); returns a 0/1 int array."""
    alist_raw = []
    with open(directory, "r") as f:
        for line in f.readlines():
            line = line.rstrip().split(" ")
            alist_raw.append(list(map(int, line)))
    return alistToNumpy(alist_raw).astype(int)


def alistToNumpy(lines):
    nCols, nRows = lines[0]
    if len(lines[2]) == nCols and len(lines[3]) == nRows:
        startIndex = 4
    else:
        startIndex = 2
    matrix = np.zeros((nRows, nCols), dtype=float)
    for col, nonzeros in enumerate(lines[startIndex:startIndex + nCols]):
        for rowIndex in nonzeros:
            if rowIndex != 0:
                matrix[rowIndex - 1, col] = 1
    return matrix
