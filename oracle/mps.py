"""Oracle: MPS helpers (random isometric MPS, orthogonalize, inner, expectation).

TEST INFRASTRUCTURE.  MPS site tensor layout: ``A[l, s, r]`` (left bond, site, right bond).
[EXT] ITensors ``orthogonalize!`` / ``inner``; pinned by the reference's
``test/test_cumps.jl:71-101,138-149,200-229``.
"""
import numpy as np

from . import linalg


def bond_dims(N, d, chi):
    """min(d^k, d^(N-k), chi) -- the synthetic MPS shape of SURVEY.md section 8(d)."""
    out = []
    for k in range(N + 1):
        e = min(k, N - k)
        out.append(int(min(chi, d ** e if e < 40 else chi)))
    return out


def random_mps(N, d, chi, rng, dtype=np.float64, dims=None):
    """Random MPS whose site tensors are isometries (QR of Gaussian), orthogonality centre
    at site 0 (i.e. every tensor right-orthogonal, then site 0 normalised)."""
    D = dims if dims is not None else bond_dims(N, d, chi)
    psi = []
    for j in range(N):
        l, r = D[j], D[j + 1]
        G = rng.standard_normal((d * r, l))
        if np.issubdtype(dtype, np.complexfloating):
            G = (G + 1j * rng.standard_normal((d * r, l))) / np.sqrt(2)
        if d * r >= l:
            Q, _ = np.linalg.qr(G)           # (d r) x l, orthonormal columns
        else:
            Q = G / np.linalg.norm(G)
        A = Q.T.reshape(l, d, r, order="F")  # A[l,(s r)] has orthonormal rows
        psi.append(np.ascontiguousarray(A.astype(dtype)))
    psi[0] = psi[0] / np.linalg.norm(psi[0])
    return psi


def product_mps(N, d, states, dtype=np.float64):
    psi = []
    for j in range(N):
        A = np.zeros((1, d, 1), dtype=dtype)
        A[0, states[j], 0] = 1.0
        psi.append(A)
    return psi


def orthogonalize(psi, j):
    """[EXT] ``orthogonalize!(psi, j)``: QR sweeps from both ends towards site j."""
    psi = [A.copy() for A in psi]
    N = len(psi)
    for b in range(0, j):
        l, d, r = psi[b].shape
        Q, R = linalg.qr(psi[b].reshape(l * d, r, order="F"))
        k = Q.shape[1]
        psi[b] = Q.reshape(l, d, k, order="F")
        psi[b + 1] = np.tensordot(R, psi[b + 1], axes=(1, 0))
    for b in range(N - 1, j, -1):
        l, d, r = psi[b].shape
        Q, R = linalg.qr(psi[b].reshape(l, d * r, order="F").T)
        k = Q.shape[1]
        psi[b] = Q.T.reshape(k, d, r, order="F")
        psi[b - 1] = np.tensordot(psi[b - 1], R.T, axes=(2, 0))
    return psi


def inner(phi, psi):
    """<phi|psi>  (``test/test_cumps.jl:71-101``)."""
    E = np.ones((1, 1), dtype=np.result_type(phi[0], psi[0]))
    for A, B in zip(phi, psi):
        T = np.tensordot(E, B, axes=(1, 0))             # (l', s, r)
        E = np.tensordot(np.conj(A), T, axes=([0, 1], [0, 1]))
    return E[0, 0]


def expect_mpo(psi, Ws):
    """<psi|H|psi>  (``inner(phi,K,psi)``: ``test/test_cumpo.jl:42-91``)."""
    E = np.ones((1, 1, 1), dtype=np.result_type(psi[0], Ws[0]))  # (l, l', a)
    for A, W in zip(psi, Ws):
        T = np.tensordot(E, A, axes=(0, 0))             # (l', a, s, r)
        T = np.tensordot(T, W, axes=([1, 2], [0, 1]))   # (l', r, s', b)
        E = np.tensordot(T, np.conj(A), axes=([0, 2], [0, 1]))  # (r, b, r')
        E = np.transpose(E, (0, 2, 1))
    return E[0, 0, 0]


def left_orthogonality_error(A):
    l, d, r = A.shape
    M = A.reshape(l * d, r, order="F")
    return float(np.linalg.norm(M.conj().T @ M - np.eye(r)))


def right_orthogonality_error(A):
    l, d, r = A.shape
    M = A.reshape(l, d * r, order="F")
    return float(np.linalg.norm(M @ M.conj().T - np.eye(l)))


def to_dense(psi):
    T = psi[0]
    for A in psi[1:]:
        T = np.tensordot(T, A, axes=(T.ndim - 1, 0))
    return T.reshape(T.shape[1:-1])


# ---- MPS / MPO algebra ([EXT] ITensors `+`, `truncate!`, `contract(::MPO, ::MPS)`; pinned by the reference's
#      consistency tests test/test_cumpo.jl:42-173 and test/test_cumps.jl:196-246)
def add(psi, phi):
    """|psi> + |phi> as an MPS (direct sum of the bond spaces, exact)."""
    N = len(psi)
    out = []
    for j in range(N):
        A, B = psi[j], phi[j]
        l1, d, r1 = A.shape
        l2, _, r2 = B.shape
        if j == 0:
            C = np.concatenate([A, B], axis=2)
        elif j == N - 1:
            C = np.concatenate([A, B], axis=0)
        else:
            C = np.zeros((l1 + l2, d, r1 + r2), dtype=np.result_type(A, B))
            C[:l1, :, :r1] = A
            C[l1:, :, r1:] = B
        out.append(C)
    return out


def truncate(psi, maxdim=None, cutoff=None):
    """[EXT] truncate!(psi): orthogonalise to the last site, then SVD-split every bond on the way back.
    Returns a right-canonical MPS (centre at site 0)."""
    psi = orthogonalize(psi, len(psi) - 1)
    N = len(psi)
    C = psi[N - 1]
    for j in range(N - 1, 0, -1):
        l, d, r = C.shape
        Lm, Rm, _ = linalg.factorize(C.reshape(l, d * r, order="F"), ortho="right", which_decomp="svd", maxdim=maxdim,
                                     cutoff=cutoff)
        k = Rm.shape[0]
        psi[j] = Rm.reshape(k, d, r, order="F")
        C = np.tensordot(psi[j - 1], Lm, axes=(2, 0))
    psi[0] = C
    return psi


def contract_mpo_mps(Ws, psi, maxdim=None, cutoff=None):
    """H|psi> as an MPS: site-wise product B[(l a), s', (r b)] = sum_s W[a,s,s',b] A[l,s,r], then truncate."""
    out = []
    for W, A in zip(Ws, psi):
        T = np.einsum("asub,lsr->laurb", W, A)
        l, a, u, r, b = T.shape
        out.append(T.reshape(l * a, u, r * b, order="F"))
    if maxdim is None and cutoff is None:
        return out
    return truncate(out, maxdim=maxdim, cutoff=cutoff)


def contract_mpo_mpo(Ks, Ls, maxdim=None, cutoff=None):
    """K*L (L acts first) as an MPO: M[(aL aK), s, s'', (bL bK)] = sum_{s'} L[aL,s,s',bL] K[aK,s',s'',bK], then the MPS
    truncation on the fused site index -- [EXT] ``contract(::MPO, ::MPO)``, pinned by test/test_cumpo.jl:145-173."""
    out = []
    for Wk, Wl in zip(Ks, Ls):
        T = np.einsum("astb,ktuc->aksubc", Wl, Wk)
        a, k, sd, u, b, c = T.shape
        out.append(T.reshape(a * k, sd, u, b * c, order="F"))
    if maxdim is None and cutoff is None:
        return out
    sd = [(W.shape[1], W.shape[2]) for W in out]
    mps = truncate([W.reshape(W.shape[0], W.shape[1] * W.shape[2], W.shape[3], order="F") for W in out], maxdim=maxdim,
                   cutoff=cutoff)
    return [t.reshape(t.shape[0], d1, d2, t.shape[2], order="F") for t, (d1, d2) in zip(mps, sd)]


def add_mpo(Ks, Ls):
    """[EXT] add(::MPO, ::MPO): direct sum on the fused site index (test/test_cumpo.jl:133-143)."""
    sd = [(W.shape[1], W.shape[2]) for W in Ks]
    f = lambda Ws: [W.reshape(W.shape[0], W.shape[1] * W.shape[2], W.shape[3], order="F") for W in Ws]
    out = add(f(Ks), f(Ls))
    return [t.reshape(t.shape[0], d1, d2, t.shape[2], order="F") for t, (d1, d2) in zip(out, sd)]


def mpo_to_dense(Ws):
    """dense operator O[(s_1..s_N), (s'_1..s'_N)] of an MPO W[a,s,s',b] (small N only)"""
    T = Ws[0]                                            # (a, s, s', b)
    for W in Ws[1:]:
        T = np.tensordot(T, W, axes=(T.ndim - 1, 0))
    T = T.reshape(T.shape[1:-1])                         # (s1, s1', s2, s2', ...)
    n = T.ndim // 2
    T = np.transpose(T, list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))
    d = int(np.prod(T.shape[:n]))
    return T.reshape(d, -1)
