"""Oracle for the row partition and halo lists of the multi-GPU path
(TEST INFRASTRUCTURE ONLY).

The reference has no partitioning code; the
contract is the one SURVEY.md section 8e writes down, restated here with plain
numpy so the product's partitioner can be compared bit-exactly:

* contiguous block rows with the ``np.array_split`` boundaries;
* ``recv`` list of a rank = sorted unique global column ids outside its row
  range, grouped by owner; local column numbering = owned columns first
  (global - lo), then halo columns in sorted-global order;
* ``send`` list owner->rank = the same ids seen from the owner, as local
  offsets.
"""
import numpy as np
import scipy.sparse as sp


def row_starts(n, nranks):
    parts = np.array_split(np.arange(n), nranks)
    starts = np.zeros(nranks + 1, dtype=np.int64)
    for r, p in enumerate(parts):
        starts[r + 1] = starts[r] + len(p)
    return starts


def partition(A, nranks):
    """Return a list (one dict per rank) with the local CSR blocks and halo
    lists.  Python loops: small and medium cases only."""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    starts = row_starts(n, nranks)
    out = []
    for r in range(nranks):
        lo, hi = int(starts[r]), int(starts[r + 1])
        blk = A[lo:hi, :]
        cols = blk.indices.astype(np.int64)
        off = (cols < lo) | (cols >= hi)
        recv = np.unique(cols[off])
        owner = np.searchsorted(starts, recv, side='right') - 1
        nloc = hi - lo
        local_cols = np.empty_like(cols)
        local_cols[~off] = cols[~off] - lo
        local_cols[off] = nloc + np.searchsorted(recv, cols[off])
        out.append(dict(lo=lo, hi=hi, indptr=blk.indptr.astype(np.int32),
                        indices=local_cols.astype(np.int32),
                        data=blk.data.copy(), recv=recv, recv_owner=owner))
    for r in range(nranks):
        send = {}
        for q in range(nranks):
            if q == r:
                continue
            ids = out[q]['recv'][out[q]['recv_owner'] == r]
            if ids.size:
                send[q] = (ids - out[r]['lo']).astype(np.int32)
        out[r]['send'] = send
    return out


def partition_rect(M, nranks):
    """The same contract for a RECTANGULAR operator (restriction / prolongation blocks of the
    row-partitioned V-cycle): rows split like ``row_starts(n_rows)``, the INPUT vector like
    ``row_starts(n_cols)``; a rank's halo = sorted unique column ids outside its own input
    range, local numbering = owned input entries first, then halo in sorted-global order."""
    M = sp.csr_matrix(M)
    rs, cs = row_starts(M.shape[0], nranks), row_starts(M.shape[1], nranks)
    out = []
    for r in range(nranks):
        lo, hi, clo, chi = int(rs[r]), int(rs[r + 1]), int(cs[r]), int(cs[r + 1])
        blk = M[lo:hi, :]
        cols = blk.indices.astype(np.int64)
        off = (cols < clo) | (cols >= chi)
        recv = np.unique(cols[off])
        owner = np.searchsorted(cs, recv, side='right') - 1
        local_cols = np.empty_like(cols)
        local_cols[~off] = cols[~off] - clo
        local_cols[off] = (chi - clo) + np.searchsorted(recv, cols[off])
        out.append(dict(lo=lo, hi=hi, clo=clo, chi=chi, indptr=blk.indptr.astype(np.int32),
                        indices=local_cols.astype(np.int32), recv=recv, recv_owner=owner))
    for r in range(nranks):
        send = {}
        for q in range(nranks):
            if q == r:
                continue
            ids = out[q]['recv'][out[q]['recv_owner'] == r]
            if ids.size:
                send[q] = (ids - out[r]['clo']).astype(np.int32)
        out[r]['send'] = send
    return out


def interior_boundary_rows(indptr, indices, nloc):
    """Rows with no halo column (interior) / at least one (boundary)."""
    has_halo = np.zeros(len(indptr) - 1, dtype=bool)
    for i in range(len(indptr) - 1):
        has_halo[i] = np.any(indices[indptr[i]:indptr[i + 1]] >= nloc)
    return np.flatnonzero(~has_halo).astype(np.int32), \
        np.flatnonzero(has_halo).astype(np.int32)
