"""Supernodal collapse of a sparse triangular factor (host side, numpy; setup only).

A supernode of a factor is a run of consecutive lines -- columns of a lower factor L, rows of an
upper factor U -- whose off-diagonal structures are nested: line j+1 has the structure of line j
minus its first entry.  Inside a supernode the factor is a DENSE triangle D followed by a panel that
is the same set of positions for every line.  For a sparse triangular solve every row of that
triangle is a dependency level of its own; with D inverted once,

    L = L~ . blockdiag(D)        L y = w   <=>   z = L~^-1 w ,  y_T = D_T^-1 z_T   (afterwards)
    U = blockdiag(D) . U~        U x = y   <=>   y'_T = D_T^-1 y_T (first) ,  x = U~^-1 y'

L~ and U~ have IDENTITY diagonal blocks (the rows of a supernode no longer depend on each other)
and the same panel structure as before (panel x dense block: no fill), so the solve on them has
as many levels as the supernodal elimination tree is high, not as the factor has rows on a path.
The diagonal blocks of the factors this is used for (coarse-level LU of the AMG hierarchy, ILUT of
the DH matrices) are well conditioned (cond 20 - 60 measured for the largest ones).

Everything here works on "line-major" arrays: CSC of L or CSR of U, indices sorted inside a line,
the diagonal stored explicitly as the first entry of its line.  In those coordinates the diagonal
block of a supernode is upper triangular in both cases (Dline[c, c+t] = entry t of line c), and
with G = inv(Dline):  new panel of line c = sum_k G[c, k] * (panel of line k);  the block-diagonal
stage multiplies by G^T (L, afterwards) resp. G (U, first).
"""
import numpy as np


def find_supernodes(ptr, idx, n):
    """first line and size of every supernode of a line-major triangular factor with sorted
    indices and an explicit diagonal first in each line.  O(nnz), vectorised."""
    ptr = np.asarray(ptr, dtype=np.int64)
    ln = np.diff(ptr)
    cont = np.zeros(n, dtype=bool)                    # cont[j]: line j+1 continues the supernode of line j
    if n > 1:
        j = np.flatnonzero((ln[1:] == ln[:-1] - 1) & (ln[:-1] >= 2))
        seg = ln[j] - 1                               # entries of line j after its diagonal
        tot = int(seg.sum())
        if tot:
            first = np.cumsum(seg) - seg
            rel = np.arange(tot, dtype=np.int64) - np.repeat(first, seg)
            a = np.repeat(ptr[j] + 1, seg) + rel      # line j without its diagonal ...
            b = np.repeat(ptr[j + 1], seg) + rel      # ... against the whole of line j+1
            neq = (idx[a] != idx[b]).astype(np.int64)
            bad = np.add.reduceat(neq, first)
            cont[j] = bad == 0
    start = np.ones(n, dtype=bool)
    start[1:] = ~cont[:-1]
    first_line = np.flatnonzero(start)
    size = np.diff(np.append(first_line, n))
    return first_line, size


S_VECTORISED = 32      # supernodes up to this size are processed per size class, larger ones one by one


def _collapse_one(ptr, data, keep, r0, s):
    """One (large) supernode with dense matrix products: returns G."""
    D = np.zeros((s, s))
    p = int(ptr[r0 + 1] - ptr[r0]) - s
    P = np.empty((s, p))
    for c in range(s):
        base = int(ptr[r0 + c])
        D[c, c:] = data[base:base + s - c]
        P[c] = data[base + s - c:base + s - c + p]
    G = np.triu(np.linalg.inv(D))
    if p:
        P = G @ P
    for c in range(s):
        base = int(ptr[r0 + c])
        data[base] = 1.0
        keep[base + 1:base + s - c] = False
        data[base + s - c:base + s - c + p] = P[c]
    return G


def collapse(ptr, idx, data, n, s_min=16, s_max=512):
    """Collapse the supernodes of s_min .. s_max lines.  Returns (ptr2, idx2, data2, blocks):
    the transformed factor in the same line-major form (diagonal blocks of the collapsed
    supernodes = identity, their panels multiplied by G) and ``blocks`` = list of
    (first_line, G) for the block-diagonal stage, G dense s x s upper triangular."""
    ptr = np.asarray(ptr, dtype=np.int64)
    idx = np.asarray(idx)
    data = np.array(data, dtype=np.float64, copy=True)
    first_line, size = find_supernodes(ptr, idx, n)
    keep = np.ones(idx.shape[0], dtype=bool)
    blocks = []
    for m in np.flatnonzero((size > S_VECTORISED) & (size >= s_min) & (size <= s_max)):
        blocks.append((int(first_line[m]), _collapse_one(ptr, data, keep, int(first_line[m]), int(size[m]))))
    for s in np.unique(size[(size >= s_min) & (size <= min(s_max, S_VECTORISED))]):
        s = int(s)
        r0 = first_line[size == s]                                   # first lines of the supernodes of this size
        cnt = r0.shape[0]
        # --- diagonal blocks: Dline[m, c, c + t] = entry t of line r0[m] + c ----------------------
        D = np.zeros((cnt, s, s))
        for c in range(s):
            base = ptr[r0 + c]
            for t in range(s - c):
                D[:, c, c + t] = data[base + t]
        G = np.linalg.inv(D)
        for c in range(s):                                            # exact zeros below the diagonal
            G[:, c, :c] = 0.0
        # --- panels: p entries per line, the same positions for the s lines of a supernode --------
        p = (ptr[r0 + 1] - ptr[r0]) - s                               # panel length (line 0 has s block entries)
        tot = int(p.sum())
        if tot:
            seg0 = np.cumsum(p) - p
            rel = np.arange(tot, dtype=np.int64) - np.repeat(seg0, p)
            which = np.repeat(np.arange(cnt), p)                      # supernode of every panel position
            pos = [np.repeat(ptr[r0 + k] + (s - k), p) + rel for k in range(s)]   # panel entries of line k
            old = [data[pos[k]] for k in range(s)]
            for c in range(s):
                acc = np.zeros(tot)
                for k in range(c, s):                                 # G is upper triangular
                    acc += G[which, c, k] * old[k]
                data[pos[c]] = acc
        # --- the block itself becomes the identity ----------------------------------------------
        for c in range(s):
            base = ptr[r0 + c]
            data[base] = 1.0
            for t in range(1, s - c):
                keep[base + t] = False
        blocks.extend((int(r0[m]), G[m]) for m in range(cnt))
    # --- drop the entries of the collapsed triangles -------------------------------------------------
    line_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
    cnt_new = np.bincount(line_of[keep], minlength=n)
    ptr2 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt_new, out=ptr2[1:])
    return ptr2, idx[keep], data[keep], blocks


def pack_blocks(blocks, n, transpose):
    """Per-row data of the block-diagonal stage y = B x (rows outside a collapsed supernode copy):
    row r of a block starting at line r0 reads x[r0 + c] for c in [c_lo, c_hi) with the weights
    vals[off + c - c_lo].  ``transpose``: B = G^T (lower factor) else B = G (upper factor).
    Returns int32 arrays (row0, c_lo, c_hi, off) of length n and the float64 value array."""
    row0 = np.arange(n, dtype=np.int32)
    c_lo = np.zeros(n, dtype=np.int32)
    c_hi = np.zeros(n, dtype=np.int32)                # c_hi == c_lo: identity row
    off = np.zeros(n, dtype=np.int64)
    vals = []
    pos = 0
    for r0, G in blocks:
        s = G.shape[0]
        B = G.T if transpose else G
        for r in range(s):
            lo, hi = (0, r + 1) if transpose else (r, s)          # G^T lower, G upper triangular
            row0[r0 + r] = r0
            c_lo[r0 + r] = lo
            c_hi[r0 + r] = hi
            off[r0 + r] = pos
            vals.append(B[r, lo:hi])
            pos += hi - lo
    v = np.concatenate(vals) if vals else np.zeros(0)
    return row0, c_lo, c_hi, off, np.ascontiguousarray(v, dtype=np.float64)
