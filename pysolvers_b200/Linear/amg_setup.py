"""Smoothed-aggregation hierarchy SETUP on the host, O(nnz).

north_star: the AMG solve phase runs on the device "using the reference's own
coarsening and hierarchy".  The reference's coarsening
(PySolvers/Linear/SmoothedAggregation.py:41-229) is pure Python and its phase 2
is O(n^2) -- ~20 h for the 2048^2 Bratu grid -- so this module restates it with
numpy in O(nnz) while reproducing its result bit for bit (checked against
hierarchies produced by the reference itself, tests/test_amg_setup.py),
including the behaviours that change numbers:

* strength test |a_ij| >= tol*sqrt(a_ii*a_jj) with tol = 0.08*0.5**(lvl-1)
  (SmoothedAggregation.py:53,62-63; SA_coarsen does not forward ``tol``, :218);
* phase 1 sweeps nodes in index order and takes a whole strong neighbourhood
  when none of it is aggregated yet (:84-89); isolated nodes first (:72-76);
* phase 2 attaches every remaining node to the aggregate -- among those that
  intersect its strong neighbourhood in the phase-1 snapshot -- holding the
  largest |a_ik| over ALL its members, first aggregate winning ties, and to the
  LAST aggregate when none intersects (agg_idx = -1, :104-127);
* the aggregates ARE the neighbourhood sets of their root nodes (:75,88), so
  phase-2 additions also enlarge the root's neighbourhood used by the filtered
  matrix (:165,178);
* filtered matrix: weak entries are lumped onto the diagonal one by one in
  stored order and kept as explicit zeros (:176-180);
* prolongator smoothing P = (I - omega D^-1 A_f) P_hat with omega = 2/3 and the
  UNFILTERED diagonal (:190-203), evaluated by the same scipy sparse product;
* restriction = transpose with the reference's row "normalisation", written
  with the same scipy calls so that it behaves exactly as the reference does
  under the installed scipy (a silent no-op on scipy 1.18);
  Galerkin product A_c = R (A P) (MLHierarchy.py:50-54).
"""
import warnings

import numpy as np
import scipy.sparse as sp


def _row_ids(A):
    return np.repeat(np.arange(A.shape[0], dtype=np.int64), np.diff(A.indptr))


def strong_mask(A, tol):
    """Boolean mask over the stored entries: j in N_i (the diagonal included
    through the set {i})."""
    d = A.diagonal()
    rows = _row_ids(A)
    cols = A.indices.astype(np.int64)
    with np.errstate(invalid='ignore'):
        thresh = tol * np.sqrt(d[rows] * d[cols])
        strong = np.abs(A.data) >= thresh          # NaN threshold -> False, as in the loop
    return strong | (rows == cols), rows, cols


def build_aggregates(A, lvl=1, tol=None):
    """Returns (agg_of, n_agg, root_of_agg): ``agg_of[i]`` is the aggregate of
    node i, ``root_of_agg[j]`` the node whose neighbourhood set the reference
    uses as aggregate j."""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    if tol is None:
        tol = 0.08 * (0.5) ** (lvl - 1)
    strong, rows, cols = strong_mask(A, tol)
    indptr = A.indptr.astype(np.int64)
    # strong neighbourhoods as CSR lists (ragged); membership is what matters
    s_rows, s_cols = rows[strong], cols[strong]
    s_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(s_ptr, s_rows + 1, 1)
    np.cumsum(s_ptr, out=s_ptr)
    # |N_i| counts distinct members; i itself is always one of them
    has_diag = np.zeros(n, dtype=bool)
    has_diag[rows[rows == cols]] = True
    size = np.diff(s_ptr) + (~has_diag)

    # isolated nodes first (SmoothedAggregation.py:72-76), then phase 1 (:84-89): a node whose
    # whole strong neighbourhood is still free founds an aggregate.  Sequential by nature: the sweep
    # runs in the library's host helper psb_sa_phase1 (a Python loop costs ~1 us per node, seconds
    # at 4 M nodes); the pure-Python version below is kept as its restatement and fallback.
    iso = np.flatnonzero(size == 1)
    agg_of = np.full(n, -1, dtype=np.int64)
    agg_of[iso] = np.arange(iso.size, dtype=np.int64)
    roots_arr = np.empty(n, dtype=np.int64)
    roots_arr[:iso.size] = iso
    n_roots = _phase1_native(n, s_ptr, s_cols, agg_of, roots_arr, iso.size)
    if n_roots is None:
        n_roots = _phase1_python(n, s_ptr, s_cols, agg_of, roots_arr, iso.size)
    roots = roots_arr[:n_roots]
    snapshot = agg_of.copy()
    n_agg = len(roots)
    # phase 2 (:104-127), all remaining nodes at once against the snapshot
    rem = np.flatnonzero(snapshot < 0)
    if rem.size:
        in_rem = np.zeros(n, dtype=bool)
        in_rem[rem] = True
        sel = in_rem[rows]
        e_rows, e_cols = rows[sel], cols[sel]
        e_abs = np.abs(A.data[sel])
        e_agg = snapshot[e_cols]
        e_strong = strong[sel]
        ok = e_agg >= 0
        # aggregates that intersect N_i: those holding a strong neighbour
        key_strong = np.unique(e_rows[ok & e_strong] * np.int64(n_agg + 1) + e_agg[ok & e_strong])
        key_all = e_rows[ok] * np.int64(n_agg + 1) + e_agg[ok]
        val_all = e_abs[ok]
        keep = np.isin(key_all, key_strong)
        key_all, val_all = key_all[keep], val_all[keep]
        # per (node, aggregate): largest |a_ik| over all members k of the aggregate
        order = np.lexsort((-val_all, key_all))
        key_s, val_s = key_all[order], val_all[order]
        first = np.ones(key_s.size, dtype=bool)
        first[1:] = key_s[1:] != key_s[:-1]
        key_u, val_u = key_s[first], val_s[first]
        node_u, agg_u = key_u // (n_agg + 1), key_u % (n_agg + 1)
        # per node: the largest value, lowest aggregate index among ties; it must be > 0
        order2 = np.lexsort((agg_u, -val_u, node_u))
        node_o, agg_o, val_o = node_u[order2], agg_u[order2], val_u[order2]
        firstn = np.ones(node_o.size, dtype=bool)
        firstn[1:] = node_o[1:] != node_o[:-1]
        choice = np.full(n, n_agg - 1, dtype=np.int64)         # agg_idx_of_max = -1 -> last
        good = firstn & (val_o > 0.0)
        choice[node_o[good]] = agg_o[good]
        agg_of[rem] = choice[rem]
    return agg_of, n_agg, np.asarray(roots, dtype=np.int64), (strong, rows, cols, snapshot)


def _phase1_python(n, s_ptr, s_cols, agg_of, roots, n_roots):
    """The phase-1 sweep in plain Python containers (restatement of psb_sa_phase1)."""
    agg_list = agg_of.tolist()
    ptr_l = s_ptr.tolist()
    col_l = s_cols.tolist()
    for i in range(n):
        if agg_list[i] >= 0:
            continue
        a, b = ptr_l[i], ptr_l[i + 1]
        ok = True
        for k in range(a, b):
            if agg_list[col_l[k]] >= 0:
                ok = False
                break
        if ok:
            for k in range(a, b):
                agg_list[col_l[k]] = n_roots
            agg_list[i] = n_roots
            roots[n_roots] = i
            n_roots += 1
    agg_of[:] = np.asarray(agg_list, dtype=np.int64)
    return n_roots


def _phase1_native(n, s_ptr, s_cols, agg_of, roots, n_roots):
    """The same sweep in libpysolv_b200 (host code, no device needed); None if the library is
    not available."""
    import ctypes as C
    try:
        from .. import _native as nat
        lib = nat.lib()
    except Exception:
        return None
    s_ptr = np.ascontiguousarray(s_ptr, dtype=np.int64)
    s_cols = np.ascontiguousarray(s_cols, dtype=np.int64)
    cnt = C.c_int64(int(n_roots))
    nat.check(lib.psb_sa_phase1(n, s_ptr.ctypes.data_as(C.c_void_p), s_cols.ctypes.data_as(C.c_void_p),
                                agg_of.ctypes.data_as(C.c_void_p), roots.ctypes.data_as(C.c_void_p),
                                C.byref(cnt)), 'psb_sa_phase1')
    return int(cnt.value)


def filtered_matrix(A, agg_of, roots, aux):
    """A with weak entries lumped onto the diagonal, explicit zeros kept
    (SmoothedAggregation.py:156-183)."""
    strong, rows, cols, snapshot = aux
    n = A.shape[0]
    Af = A.copy()
    # j in neighbourhoods[i]  <=>  strong, or i is the root of j's aggregate and j was
    # added to that aggregate (phase 1 members are strong neighbours of the root anyway)
    is_root_of = np.full(n, -1, dtype=np.int64)
    is_root_of[roots] = np.arange(roots.size)
    in_set = strong | ((is_root_of[rows] >= 0) & (is_root_of[rows] == agg_of[cols]))
    weak = np.flatnonzero(~in_set)
    # diagonal position of every row (first stored entry with j == i)
    diag_pos = np.full(n, -1, dtype=np.int64)
    dpos = np.flatnonzero(rows == cols)
    diag_pos[rows[dpos][::-1]] = dpos[::-1]
    data = Af.data
    # sequential lumping in stored order: ((d - a1) - a2) - ...
    np.subtract.at(data, diag_pos[rows[weak]], data[weak].copy())
    data[weak] = 0
    return Af


def smooth_prolongator(Phat, A, Af, omega=(2 / 3)):
    """(I - omega D^-1 A_f) P_hat with the unfiltered diagonal (:185-205)."""
    S = omega * Af
    d = A.diagonal()
    rows = _row_ids(S)
    S.data /= d[rows]
    on_diag = rows == S.indices
    S.data[on_diag] = 1 - S.data[on_diag]
    S.data[~on_diag] = -S.data[~on_diag]
    return S.dot(Phat)


def sa_coarsen(A, lvl=1):
    """Prolongator of one coarsening step (SA_coarsen, :208-229)."""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    agg_of, n_agg, roots, aux = build_aggregates(A, lvl=lvl)
    Phat = sp.csr_matrix((np.ones(n), agg_of, np.arange(n + 1)), shape=(n, n_agg))
    Af = filtered_matrix(A, agg_of, roots, aux)
    P = smooth_prolongator(Phat, A, Af)
    return P.tocsr(), agg_of


_NORMALIZE_IS_NOOP = None


def _normalisation_is_noop():
    """Does the reference's row "normalisation" (MLHierarchy.py:71-75) change anything under
    the installed scipy?  Probed once with the very same calls on a 2 x 3 example: on scipy
    1.18 ``row /= nrm`` on a ``lil.getrowview`` rebinds the view instead of mutating the
    parent, i.e. it is a silent no-op."""
    global _NORMALIZE_IS_NOOP
    if _NORMALIZE_IS_NOOP is None:
        probe = sp.csr_matrix(np.array([[1.0, 3.0, 0.0], [0.0, 2.0, 6.0]]))
        _NORMALIZE_IS_NOOP = bool(np.array_equal(_restriction_literal(probe.T.tocsr(), True).toarray(),
                                                 probe.toarray()))
    return _NORMALIZE_IS_NOOP


def _restriction_literal(I_up, normalize):
    I_down = I_up.transpose(copy=True).tolil()
    if normalize:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for r in range(I_down.shape[0]):
                row = I_down.getrowview(r)
                nrm = row.sum()
                row /= nrm
    return I_down.tocsr()


def restriction_of(I_up, normalize=True):
    """Transpose plus the reference's row normalisation, through the same scipy calls
    (MLHierarchy.py:60-78) so that it tracks the installed scipy.  When the probe shows the
    normalisation loop to be a no-op (scipy 1.18) the O(rows) Python loop is skipped -- the
    result is the same matrix."""
    if normalize and _normalisation_is_noop():
        normalize = False
    if not normalize and not (I_up.data == 0.0).any():
        # without the row loop the lil round trip is just a transpose with sorted indices: the
        # same arrays bit for bit (tests/test_amg_setup.py), 20 x faster at a million rows
        R = sp.csr_matrix(I_up.transpose(copy=True))
        R.sort_indices()
        return R
    return _restriction_literal(I_up, normalize)


def build_hierarchy(A_fine, num_levels=2, normalize=True):
    """ops[k], ups[k] (k -> k+1), downs[k] (k+1 -> k); level 0 is the coarsest."""
    ops = [None] * num_levels
    ups = [None] * num_levels
    downs = [None] * num_levels
    ops[num_levels - 1] = A_fine
    for lev in reversed(range(num_levels - 1)):
        P, _ = sa_coarsen(ops[lev + 1], lvl=lev + 1)
        ups[lev] = P
        downs[lev] = restriction_of(P, normalize)
        ops[lev] = downs[lev] * (ops[lev + 1] * ups[lev])
    return ops, ups, downs
