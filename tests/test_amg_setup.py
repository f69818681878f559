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


@pytest.mark.parametrize('m,nlev', [(128, 2), (128, 3), (256, 2)])
def test_hierarchy_at_size_digests(golden_large, m, nlev):
    """SURVEY.md section 8f-1 asks for m <= 256: phase-2 tie breaking, ``agg_idx = -1`` and the set
    aliasing are size-dependent paths.  The reference's hierarchies at these sizes are pinned
    by SHA-256 digests of every array (tests/golden/make_golden_large.py)."""
    import hashlib
    from conftest import assert_csr_digest
    A = -fd_laplacian_2d(0.0, 1.0, m)
    ops, ups, downs = amg_setup.build_hierarchy(A, num_levels=nlev)
    tag = 'amg/m%d_L%d' % (m, nlev)
    for k in range(nlev):
        assert_csr_digest(ops[k], golden_large, '%s/A%d' % (tag, k))
    for k in range(nlev - 1):
        assert_csr_digest(ups[k], golden_large, '%s/P%d' % (tag, k))
        assert_csr_digest(downs[k], golden_large, '%s/R%d' % (tag, k))
    agg_of, n_agg, roots, _ = amg_setup.build_aggregates(A, lvl=nlev - 1)
    assert n_agg == int(golden_large[tag + '/n_agg'])
    order = np.argsort(agg_of, kind='stable')                 # members ascending inside each aggregate
    sizes = np.bincount(agg_of, minlength=n_agg).astype(np.int64)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()
    assert sha(sizes) == str(golden_large[tag + '/sha_agg_sizes'])
    assert sha(order) == str(golden_large[tag + '/sha_agg_flat'])


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


@pytest.mark.parametrize('s_min', [2, 8])
def test_supernodal_collapse_is_the_same_solve(s_min):
    """Linear/supernodes.py on the host: L = L~ blockdiag(D), U = blockdiag(D) U~ with identity
    diagonal blocks in L~ / U~ and unchanged panel structure; solving with the collapsed factors plus
    the block-diagonal stage equals the original triangular solves to rounding, and the collapsed
    factors have fewer dependency levels."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import precond
    from pysolvers_b200.Linear import supernodes as SN
    rng = np.random.default_rng(3)
    A = sp.csc_matrix(-fd_laplacian_2d(0.0, 1.0, 36))
    lu = spla.splu(A, permc_spec='MMD_AT_PLUS_A')
    n = A.shape[0]

    def block_stage(blocks, x, transpose):
        row0, c_lo, c_hi, off, vals = SN.pack_blocks(blocks, n, transpose)
        y = x.copy()
        for r in np.flatnonzero(c_hi > c_lo):
            y[r] = vals[off[r]:off[r] + c_hi[r] - c_lo[r]] @ x[row0[r] + c_lo[r]:row0[r] + c_hi[r]]
        return y

    L = sp.csc_matrix(lu.L)
    L.sort_indices()
    first, size = SN.find_supernodes(L.indptr, L.indices, n)
    assert first[0] == 0 and size.sum() == n and size.max() >= 8
    ptr2, idx2, dat2, blocks = SN.collapse(L.indptr, L.indices, L.data, n, s_min=s_min)
    Lt = sp.csc_matrix((dat2, idx2, ptr2), shape=(n, n)).tocsr()
    assert blocks and Lt.nnz < L.nnz and sp.triu(Lt, 1).nnz == 0
    lv0 = len(precond.level_sets(sp.csr_matrix(L), lower=True)[1]) - 1
    lv1 = len(precond.level_sets(Lt, lower=True)[1]) - 1
    assert lv1 < lv0
    w = rng.standard_normal(n)
    ref = spla.spsolve_triangular(sp.csr_matrix(L), w, lower=True, unit_diagonal=True)
    got = block_stage(blocks, spla.spsolve_triangular(Lt, w, lower=True, unit_diagonal=True), True)
    assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)

    U = sp.csr_matrix(lu.U)
    U.sort_indices()
    ptr2, idx2, dat2, blocks = SN.collapse(U.indptr, U.indices, U.data, n, s_min=s_min)
    Ut = sp.csr_matrix((dat2, idx2, ptr2), shape=(n, n))
    assert blocks and sp.tril(Ut, -1).nnz == 0
    assert len(precond.level_sets(Ut, lower=False)[1]) < len(precond.level_sets(U, lower=False)[1])
    ref = spla.spsolve_triangular(U, w, lower=False)
    got = spla.spsolve_triangular(Ut, block_stage(blocks, w, False), lower=False)
    assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)


def test_phase1_native_equals_python_sweep():
    """psb_sa_phase1 (host code in the library) and its pure-Python restatement give the same
    aggregates and roots."""
    import scipy.sparse as sp
    rng = np.random.default_rng(12)
    n = 3000
    G = sp.random(n, n, density=3.0 / n, random_state=rng, format='csr')
    G = ((G + G.T) > 0).astype(np.float64).tocsr()
    G.setdiag(0)
    G.eliminate_zeros()
    s_ptr = G.indptr.astype(np.int64)
    s_cols = G.indices.astype(np.int64)
    size = np.diff(s_ptr) + 1
    iso = np.flatnonzero(size == 1)
    outs = []
    for fn in (amg_setup._phase1_native, amg_setup._phase1_python):
        agg = np.full(n, -1, dtype=np.int64)
        agg[iso] = np.arange(iso.size)
        roots = np.empty(n, dtype=np.int64)
        roots[:iso.size] = iso
        k = fn(n, s_ptr, s_cols, agg, roots, iso.size)
        assert k is not None
        outs.append((k, agg.copy(), roots[:k].copy()))
    assert outs[0][0] == outs[1][0] and outs[0][0] > iso.size
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
