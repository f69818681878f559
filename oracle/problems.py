"""Literal (slow) problem generators for checking the vectorised ones
(TEST INFRASTRUCTURE ONLY).

* ``fd_laplacian_2d_loop`` restates examples/FDLaplacian2D.py:5-23: h =
  |b-a|/(m+1), row k = m*iy + ix, diagonal -4/h^2, neighbours 1/h^2 inserted in
  the order [k, k-m, k+m, k-1, k+1] (dok insertion order survives tocsr()).
* ``fd_laplacian_3d_loop`` is the builder-defined 7-point extension of SURVEY.md
  section 8d (config 4): row k = m^2*iz + m*iy + ix, diagonal +6/h^2,
  neighbours -1/h^2 in the order [k, k-m^2, k+m^2, k-m, k+m, k-1, k+1].
* ``bratu_F`` / ``bratu_J`` restate examples/FDBratu2D.py:20-29.
"""
import numpy as np
import scipy.sparse as sp


def fd_laplacian_2d_loop(a, b, m):
    h = np.abs(b - a) / np.double(m + 1)
    A = sp.dok_matrix((m * m, m * m))
    for ix in range(m):
        for iy in range(m):
            k = m * iy + ix
            A[k, k] = -4.0 / h / h
            if iy > 0:
                A[k, k - m] = 1.0 / h / h
            if iy < m - 1:
                A[k, k + m] = 1.0 / h / h
            if ix > 0:
                A[k, k - 1] = 1.0 / h / h
            if ix < m - 1:
                A[k, k + 1] = 1.0 / h / h
    return A.tocsr()


def fd_laplacian_3d_loop(a, b, m):
    h = np.abs(b - a) / np.double(m + 1)
    n = m * m * m
    A = sp.dok_matrix((n, n))
    for iz in range(m):
        for iy in range(m):
            for ix in range(m):
                k = m * m * iz + m * iy + ix
                A[k, k] = 6.0 / h / h
                if iz > 0:
                    A[k, k - m * m] = -1.0 / h / h
                if iz < m - 1:
                    A[k, k + m * m] = -1.0 / h / h
                if iy > 0:
                    A[k, k - m] = -1.0 / h / h
                if iy < m - 1:
                    A[k, k + m] = -1.0 / h / h
                if ix > 0:
                    A[k, k - 1] = -1.0 / h / h
                if ix < m - 1:
                    A[k, k + 1] = -1.0 / h / h
    return A.tocsr()


def bratu_F(A, alpha, u):
    return A * u - alpha * np.exp(-u)


def bratu_J(A, alpha, u):
    J = A.copy()
    g = alpha * np.exp(-u)
    d = J.diagonal()
    J.setdiag(d + g)
    return J
