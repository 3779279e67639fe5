"""Oracle: dense tensor primitives (contract / permute / add / norm).  TEST INFRASTRUCTURE.

Storage convention (reference ``src/tensor/cudense.jl:252-254``): a tensor is a flat
**column-major** vector over its index order.  Here a tensor is a NumPy ndarray whose
*logical* shape is the index extents; ``flat(T)`` / ``unflat(v, dims)`` convert to and
from the flat column-major buffer that crosses the C ABI.
"""
import numpy as np


def flat(T):
    """ndarray -> flat column-major vector (what ``data(store(T))`` holds)."""
    return np.ascontiguousarray(np.asarray(T).ravel(order="F"))


def unflat(v, dims):
    """flat column-major vector -> ndarray with logical shape ``dims``."""
    return np.asarray(v).reshape(tuple(dims), order="F")


def output_labels(la, lb):
    """Labels of A*B: A's free labels in A's order, then B's free labels in B's order.

    [EXT] NDTensors ``contraction_output``; the reference numbers modes
    contracted -> A-only -> B-only (``src/tensor/cudense.jl:258-283``) which yields the
    same output order.
    """
    la, lb = list(la), list(lb)
    return tuple([l for l in la if l not in lb] + [l for l in lb if l not in la])


def contract(A, la, B, lb, conj_a=False, conj_b=False):
    """C = sum over shared labels of A*B.  Returns (C, lc).

    Follows ``contract!!`` -> ``_contract!`` (``src/tensor/cudense.jl:83-110,238-331``):
    scalar*tensor, outer product and general contraction are all the same einsum here.
    """
    la, lb = list(la), list(lb)
    if len(set(la)) != len(la) or len(set(lb)) != len(lb):
        raise ValueError("repeated label inside one tensor")
    A = np.conj(A) if conj_a else np.asarray(A)
    B = np.conj(B) if conj_b else np.asarray(B)
    shared = [l for l in la if l in lb]
    for l in shared:
        if A.shape[la.index(l)] != B.shape[lb.index(l)]:
            raise ValueError("dimension mismatch on contracted label %r" % (l,))
    axa = [la.index(l) for l in shared]
    axb = [lb.index(l) for l in shared]
    C = np.tensordot(A, B, axes=(axa, axb))
    return C, output_labels(la, lb)


def contract_into(C, lc, A, la, B, lb, alpha=1.0, beta=0.0, conj_a=False, conj_b=False):
    """C <- alpha * contract(A,B) permuted to lc + beta * C  (``src/tensor/dense.jl:1-48``)."""
    R, lr = contract(A, la, B, lb, conj_a, conj_b)
    perm = [list(lr).index(l) for l in lc]
    R = np.transpose(R, perm) if R.ndim else R
    if beta == 0:
        return alpha * R
    return alpha * R + beta * np.asarray(C)


def permute(A, la, lc):
    """B <- A with modes reordered to lc  (``permute!`` ``src/tensor/cudense.jl:447-478``)."""
    perm = [list(la).index(l) for l in lc]
    return np.transpose(np.asarray(A), perm)


def axpby(alpha, A, la, beta, B, lb):
    """B <- alpha*perm(A) + beta*B   (``+``/``-``: ``src/tensor/cudense.jl:333-445``)."""
    return alpha * permute(A, la, lb) + beta * np.asarray(B)


def norm(A):
    """Frobenius norm (``src/tensor/cudense.jl:27``)."""
    return float(np.linalg.norm(np.asarray(A).ravel()))


def dot(A, B):
    """<A|B> = sum conj(A)*B  ([EXT] ITensors ``dot`` = scalar(dag(A)*B))."""
    return np.vdot(np.asarray(A).ravel(), np.asarray(B).ravel())


def rel_err(X, Y):
    """Relative Frobenius error ||X-Y|| / ||Y||  (the north-star contraction bar: 1e-12)."""
    X = np.asarray(X)
    Y = np.asarray(Y)
    d = np.linalg.norm((X - Y).ravel())
    n = np.linalg.norm(Y.ravel())
    return float(d / n) if n > 0 else float(d)
