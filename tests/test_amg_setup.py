"""The O(nnz) smoothed-aggregation setup must reproduce the hierarchy the
reference's own (quadratic, pure-Python) coarsening builds -- aggregates, P, R
and the Galerkin operators -- bit for bit.  Fixtures:
hierarchies produced by the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_csr
from pysolvers_b200.Linear import amg_setup
from pysolvers_b200.problems import fd_laplacian_2d


def _same_csr(M, G, what):
    M = sp.csr_matrix(M)
    assert M.shape == G.shape, what
    assert np.array_equal(M.indptr, G.indptr), what
    assert np.array_equal(M.indices, G.indices), what      # stored order included
    assert np.array_equal(M.data, G.data), what


@pytest.mark.parametrize('m,nlev', [(16, 2), (32, 2), (32, 3)])
def test_hierarchy_bit_identical_to_reference(golden, m, nlev):
    A = -fd_laplacian_2d(0.0, 1.0, m)
    ops, ups, downs = amg_setup.build_hierarchy(A, num_levels=nlev)
    tag = 'amg/m%d_L%d' % (m, nlev)
    for k in range(nlev):
        _same_csr(ops[k], golden_csr(golden, '%s/A%d' % (tag, k)), 'A%d' % k)
    for k in range(nlev - 1):
        _same_csr(ups[k], golden_csr(golden, '%s/P%d' % (tag, k)), 'P%d' % k)
        _same_csr(downs[k], golden_csr(golden, '%s/R%d' % (tag, k)), 'R%d' % k)


@pytest.mark.parametrize('m,nlev', [(16, 2), (32, 2), (32, 3)])
def test_aggregates_match_reference(golden, m, nlev):
    A = -fd_laplacian_2d(0.0, 1.0, m)
    agg_of, n_agg, roots, _ = amg_setup.build_aggregates(A, lvl=nlev - 1)
    tag = 'amg/m%d_L%d' % (m, nlev)
    sizes = golden[tag + '/agg_sizes']
    flat = golden[tag + '/agg_flat']
    assert n_agg == len(sizes)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    for j in range(n_agg):
        members = flat[ptr[j]:ptr[j + 1]]
        assert np.array_equal(np.sort(np.flatnonzero(agg_of == j)), members), j


@pytest.mark.parametrize('tag,nlev', [('amg/dh7_L2', 2), ('amg/dh9_L3', 3), ('amg/bratuJ_m20_L2', 2),
                                      ('amg/rand300_L2', 2)])
def test_hierarchy_irregular_matrices(golden, tag, nlev):
    """FE matrices / random graphs with weak couplings exercise the filtered-matrix
    lumping, the phase-2 tie breaking and the set-aliasing quirk."""
    A = golden_csr(golden, tag + '/Afine')
    ops, ups, downs = amg_setup.build_hierarchy(A, num_levels=nlev)
    for k in range(nlev):
        _same_csr(ops[k], golden_csr(golden, '%s/A%d' % (tag, k)), 'A%d' % k)
    for k in range(nlev - 1):
        _same_csr(ups[k], golden_csr(golden, '%s/P%d' % (tag, k)), 'P%d' % k)
        _same_csr(downs[k], golden_csr(golden, '%s/R%d' % (tag, k)), 'R%d' % k)


def test_restriction_fast_path_equals_literal():
    """restriction_of's transpose short cut gives the arrays of the reference's lil route."""
    import scipy.sparse as sp
    from pysolvers_b200.Linear import amg_setup
    from pysolvers_b200.problems import FDBratu2D
    for m in (9, 40):
        J = sp.csr_matrix(FDBratu2D(m, 0.5).evalJ(np.ones(m * m)))
        P, _ = amg_setup.sa_coarsen(J, lvl=1)
        fast = amg_setup.restriction_of(P, True)
        lit = amg_setup._restriction_literal(P, not amg_setup._normalisation_is_noop())
        assert np.array_equal(fast.indptr, lit.indptr) and np.array_equal(fast.indices, lit.indices)
        assert np.array_equal(fast.data, lit.data)
