"""Structured orthogonal ("butterfly") multiply (CPU oracle, numpy).

Restates reference method.py:16-78.  A butterfly matrix of size n is the triple
Bpp = ([B0, B1], p_in, p_out) with (p1, p2) = butterfly_factors(n),
B0 of shape (n/p1, p1, p1) or (p1, p1), B1 of shape (n/p2, p2, p2) or (p2, p2).
Acting on a column vector v (method.py:46-67):

    u = v[p_in];  T = u.view(p1, p2)
    T[:, b] = B0[b] @ T[:, b]   for every column b      ("COL" pass)
    T[a, :] = B1[a] @ T[a, :]   for every row a         ("ROW" pass)
    out = T.view(n)[p_out]

The kernels run the two passes in a per-side *layout* chosen so that the larger
block is contiguous, with p_out folded into the packed weight's K (or N) order
and p_in applied as a gather (V side) or scatter (U side).  `side_plan` builds
those index maps; it is the checker for quip_b200/incoherence.py.

Test infrastructure only (see oracle/__init__.py).
"""
import math
import numpy as np


def prime_factors(n):
    n, d, out = int(n), 2, []
    while d * d <= n:
        while n % d == 0:
            out.append(d)
            n //= d
        d += 1 if d == 2 else 2
    if n > 1:
        out.append(n)
    return out


def butterfly_factors(n):
    """method.py:16-18: product of the even-indexed / odd-indexed sorted prime factors."""
    pf = prime_factors(n)
    return math.prod(pf[0::2]), math.prod(pf[1::2])


def _blk(B, p):
    B = np.asarray(B)
    return B.reshape(-1, p, p)        # (m, p, p); m == 1 for the no-block variant


def mul_butterfly(Bpp, x):
    """method.py:46-67 on a column vector (n,) or matrix (n, q)."""
    (B, p_in, p_out) = Bpp
    x = np.asarray(x)
    one_d = x.ndim == 1
    if one_d:
        x = x[:, None]
    n, q = x.shape
    p1, p2 = butterfly_factors(n)
    B0, B1 = _blk(B[0], p1), _blk(B[1], p2)
    T = x[np.asarray(p_in)].reshape(p1, p2, q)
    T = np.einsum('bij,jbq->ibq', np.broadcast_to(B0, (p2, p1, p1)), T)
    T = np.einsum('aij,ajq->aiq', np.broadcast_to(B1, (p1, p2, p2)), T)
    out = T.reshape(n, q)[np.asarray(p_out)]
    return out[:, 0] if one_d else out


def dense(Bpp, n, dtype=np.float32):
    """method.py:71-78: the dense orthogonal matrix."""
    return mul_butterfly(Bpp, np.eye(n, dtype=dtype))


def rows_times_Vt(X, Bpp):
    """X @ V^T for row-batched X (M, n): butterfly applied to every row."""
    return mul_butterfly(Bpp, np.asarray(X).T).T


def rows_times_U(Z, Bpp):
    """Z @ U for row-batched Z (M, n) (== U^T applied to every row); SURVEY App. B."""
    (B, p_in, p_out) = Bpp
    Z = np.asarray(Z)
    M, n = Z.shape
    p1, p2 = butterfly_factors(n)
    B0, B1 = _blk(B[0], p1), _blk(B[1], p2)
    W = np.empty_like(Z)
    W[:, np.asarray(p_out)] = Z
    T = W.reshape(M, p1, p2)
    T = np.einsum('aji,maj->mai', np.broadcast_to(B1, (p1, p2, p2)), T)
    T = np.einsum('bji,mjb->mib', np.broadcast_to(B0, (p2, p1, p1)), T)
    out = np.empty_like(Z)
    out[:, np.asarray(p_in)] = T.reshape(M, n)
    return out


# --------------------------------------------------------------------------
# kernel-side plan: layout choice, folded permutations, pass descriptors
# --------------------------------------------------------------------------
def side_plan(Bpp, n, side):
    """Index maps + pass list for one side ('V' acts on K, 'U' on N).

    Returns a dict with
      layout   'A' (index a*p2+b) if p2 >= p1 else 'B' (index b*p1+a)
      io_idx   gather (V) / scatter (U) index: position l of the layout buffer
               corresponds to feature io_idx[l] of x (V) or y (U)
      order    weight order: layout position l holds original column (V) / row (U)
               order[l] of the reference's Q
      passes   [(F, p, nblk, strided)] in execution order; F (nb, p, p) float32 with
               nb in {nblk, 1}; out_i = sum_j F[blk][i][j] in_j; element j of block blk
               sits at blk*p + j (contiguous) or j*nblk + blk (strided)
    """
    (B, p_in, p_out) = Bpp
    p_in, p_out = np.asarray(p_in), np.asarray(p_out)
    p1, p2 = butterfly_factors(n)
    B0, B1 = _blk(B[0], p1).astype(np.float32), _blk(B[1], p2).astype(np.float32)
    i = np.arange(n)
    a, b = i // p2, i % p2
    layout = 'A' if p2 >= p1 else 'B'
    l_of_i = i if layout == 'A' else b * p1 + a
    inv_pout = np.argsort(p_out)
    io_idx = np.empty(n, np.int64)
    io_idx[l_of_i] = p_in
    order = np.empty(n, np.int64)
    order[l_of_i] = inv_pout
    col = dict(p=p1, nblk=p2, strided=(layout == 'A'))
    row = dict(p=p2, nblk=p1, strided=(layout == 'B'))
    if side == 'V':
        passes = [(B0, col), (B1, row)]
    else:
        passes = [(np.ascontiguousarray(B1.transpose(0, 2, 1)), row),
                  (np.ascontiguousarray(B0.transpose(0, 2, 1)), col)]
    return dict(layout=layout, io_idx=io_idx, order=order, p1=p1, p2=p2,
                passes=[(F, d['p'], d['nblk'], d['strided']) for F, d in passes])


def apply_pass(X, F, p, nblk, strided):
    """One block-diagonal pass on row-batched X (M, n) in the side's layout."""
    X = np.asarray(X)
    M, n = X.shape
    Fb = np.broadcast_to(F, (nblk, p, p))
    if strided:
        T = X.reshape(M, p, nblk)
        return np.einsum('bij,mjb->mib', Fb, T).reshape(M, n)
    T = X.reshape(M, nblk, p)
    return np.einsum('bij,mbj->mbi', Fb, T).reshape(M, n)
