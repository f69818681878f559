"""Oracle restatement of the reference preconditioners (TEST INFRASTRUCTURE ONLY).

* IC   -- PySolvers/Linear/ICPreconditioner.py:34-63
* ILUT -- PySolvers/Linear/ILUTPreconditioner.py:37-78
* level sets of a sparse triangular factor -- contract of SURVEY.md section 8e
  (the reference has no level analysis; this numpy version defines it)

The arithmetic is SuperLU's (scipy.sparse.linalg.spilu / spsolve_triangular /
SuperLU.solve, scipy 1.18.1 in this image), exactly as the reference calls it.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def ic_factor(A, drop_tol=0.001, fill_factor=15):
    """Return (L, Lt) as CSR exactly as ICRightPreconditioner.__init__ builds
    them (ICPreconditioner.py:45-56)."""
    ilu = spla.spilu(A.tocsc(), drop_tol=drop_tol, fill_factor=fill_factor,
                     diag_pivot_thresh=0.0, options={'ColPerm': 'NATURAL'})
    n = A.shape[0]
    scale = np.reciprocal(np.sqrt(ilu.U.diagonal()))
    Dinv = sp.dia_matrix((scale, [0]), shape=(n, n))
    Lt = Dinv * ilu.U
    del ilu
    L = Lt.transpose()
    return L.tocsr(), Lt.tocsr()


def ic_apply(L, Lt, v):
    # ICPreconditioner.py:58-63
    u = spla.spsolve_triangular(L, v, lower=True)
    return spla.spsolve_triangular(Lt, u, lower=False)


def ilut_factor(A, drop_tol=0.001, fill_factor=15):
    # ILUTPreconditioner.py:51-53 (SuperLU default COLAMD column permutation)
    return spla.spilu(A.tocsc(), drop_tol=drop_tol, fill_factor=fill_factor,
                      diag_pivot_thresh=0.0)


def ilut_apply(ilu, v):
    # ILUTPreconditioner.py:66-67, 77-78
    return ilu.solve(v)


def ilut_apply_explicit(ilu, v):
    """x = Pc . U^-1 . L^-1 . Pr . v written out with the factors SuperLU
    exposes: Pr[perm_r[i], i] = 1 and
    Pc[i, perm_c[i]] = 1.  Used to pin the permutation convention the device
    path uploads."""
    n = ilu.shape[0]
    w = np.empty(n)
    w[ilu.perm_r] = v                     # (Pr v)[perm_r[i]] = v[i]
    y = spla.spsolve_triangular(ilu.L.tocsr(), w, lower=True,
                                unit_diagonal=True)
    z = spla.spsolve_triangular(ilu.U.tocsr(), y, lower=False)
    return z[ilu.perm_c]                  # (Pc z)[i] = z[perm_c[i]]


def level_sets(T, lower=True):
    """Dependency levels of a sparse triangular matrix in CSR.

    level(i) = 1 + max(level(j)) over the strictly lower (upper) stored
    entries j of row i, or 0 when the row has none; rows are listed level-major
    and ascending inside a level.  Returns (level, level_ptr, level_rows) as
    int32 arrays.
    """
    T = sp.csr_matrix(T)
    n = T.shape[0]
    indptr, indices = T.indptr, T.indices
    level = np.zeros(n, dtype=np.int32)
    rows = range(n) if lower else range(n - 1, -1, -1)
    for i in rows:
        cols = indices[indptr[i]:indptr[i + 1]]
        deps = cols[cols < i] if lower else cols[cols > i]
        if deps.size:
            level[i] = level[deps].max() + 1
    nlev = int(level.max()) + 1 if n else 0
    counts = np.bincount(level, minlength=nlev)
    level_ptr = np.zeros(nlev + 1, dtype=np.int32)
    np.cumsum(counts, out=level_ptr[1:])
    level_rows = np.argsort(level, kind='stable').astype(np.int32)
    return level, level_ptr, level_rows


def trsv_rowwise(T, b, lower=True, unit_diagonal=False):
    """Row-oriented substitution, the summation order of the device kernels:
    the dependencies of a row are subtracted in the order in which they become
    available -- ascending dependency LEVEL (level_sets), ties in stored
    column order -- product rounded first, and the result is multiplied by the
    rounded reciprocal of the diagonal last.

    Why this order: a row's newest dependency is then its last operand, so a
    wavefront solver has one multiply-subtract left when it arrives (in stored
    order an IC row of the 2-D Laplacian still has ~8 operands to go).  The
    reference imposes no order of its own here: scipy's spsolve_triangular
    (ICPreconditioner.py:61,63) hands the factor to SuperLU's column-oriented
    gstrs.  The reciprocal IS the reference's arithmetic: spsolve_triangular
    forms ``invdiag = 1/diag`` once and finishes with ``x = y * invdiag``; it
    never divides on the data path.  Pure Python: small cases only."""
    T = sp.csr_matrix(T)
    n = T.shape[0]
    level, _, _ = level_sets(T, lower=lower)
    x = np.zeros(n)
    rows = range(n) if lower else range(n - 1, -1, -1)
    for i in rows:
        acc = b[i]
        d = 1.0
        deps = []
        for jj in range(T.indptr[i], T.indptr[i + 1]):
            j = T.indices[jj]
            if j == i:
                d = T.data[jj]
            elif (j < i) == lower:
                deps.append((level[j], jj))
        for _, jj in sorted(deps):
            acc = acc - T.data[jj] * x[T.indices[jj]]
        x[i] = acc if unit_diagonal else acc * (1.0 / d)
    return x
