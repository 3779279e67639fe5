"""Oracle: gate application on an MPS (TEBD step).  TEST INFRASTRUCTURE.

[EXT] ITensors 0.2 ``apply(gates, psi; cutoff, maxdim)`` (= ``product``), reached from the
reference at ``examples/gate_evolution.jl:46``.  Per gate: move the orthogonality centre to
the first site the gate acts on (QR sweeps), build theta, contract the gate, split by
``factorize`` (left-orthogonal, centre moves to the last site of the gate).
Gate layouts: one-site ``G[s', s]``; two-site ``G[s1', s2', s1, s2]`` on sites (n, n+1).
"""
import numpy as np

from . import linalg


def move_center(psi, frm, to):
    """QR-shift the orthogonality centre from site ``frm`` to site ``to`` (in place)."""
    while frm < to:
        l, d, r = psi[frm].shape
        Q, R = linalg.qr(psi[frm].reshape(l * d, r, order="F"))
        psi[frm] = Q.reshape(l, d, Q.shape[1], order="F")
        psi[frm + 1] = np.tensordot(R, psi[frm + 1], axes=(1, 0))
        frm += 1
    while frm > to:
        l, d, r = psi[frm].shape
        Q, R = linalg.qr(psi[frm].reshape(l, d * r, order="F").T)
        psi[frm] = Q.T.reshape(Q.shape[1], d, r, order="F")
        psi[frm - 1] = np.tensordot(psi[frm - 1], R.T, axes=(2, 0))
        frm -= 1
    return to


def apply_gate(psi, center, G, n, maxdim=None, cutoff=None, mindim=1):
    """Apply one gate at site n (one-site) or bond (n, n+1) (two-site).  Returns new centre."""
    center = move_center(psi, center, n)
    if G.ndim == 2:
        psi[n] = np.transpose(np.tensordot(G, psi[n], axes=(1, 1)), (1, 0, 2))
        return center
    theta = np.tensordot(psi[n], psi[n + 1], axes=(2, 0))           # (l,s1,s2,r)
    theta = np.tensordot(theta, G, axes=([1, 2], [2, 3]))           # (l,r,s1',s2')
    theta = np.transpose(theta, (0, 2, 3, 1))
    l, d1, d2, r = theta.shape
    M = theta.reshape(l * d1, d2 * r, order="F")
    Lm, Rm, spec = linalg.factorize(M, ortho="left", maxdim=maxdim, mindim=mindim, cutoff=cutoff)
    k = Lm.shape[1]
    psi[n] = Lm.reshape(l, d1, k, order="F")
    psi[n + 1] = Rm.reshape(k, d2, r, order="F")
    return n + 1


def apply(gates, psi, center=0, maxdim=None, cutoff=None):
    """gates: list of (G, n).  Sequential application, exactly as ITensors ``apply``."""
    psi = [A.astype(np.result_type(A, *[g for g, _ in gates])) for A in psi]
    for G, n in gates:
        center = apply_gate(psi, center, G, n, maxdim=maxdim, cutoff=cutoff)
    return psi, center


def tebd_layer_gates(N, G, parity):
    """Even (parity 0: bonds 0,2,4..) or odd (1,3,5..) layer of a uniform two-site gate."""
    return [(G, n) for n in range(parity, N - 1, 2)]


# ---------------------------------------------------------------------------------------------
# B form (right-canonical tensors in the Schmidt bases + Schmidt values of every bond).
# This is what makes the gates of one even/odd layer INDEPENDENT: a gate on
# (n, n+1) needs only B[n], B[n+1] and the Schmidt values of the bond to its left.  Update without
# dividing by Schmidt values (Hastings, J. Math. Phys. 50, 095207 (2009)); B[0..N-1] is at every time a
# right-canonical MPS of the state, i.e. directly comparable with `apply`'s output.
# ---------------------------------------------------------------------------------------------
def canonical_bform(psi):
    """psi (any gauge) -> (Bs, lams): lams[j] = Schmidt values of the bond left of site j (lams[0] = lams[N] = [1])."""
    psi = [np.array(A) for A in psi]
    N = len(psi)
    for b in range(N - 1):                                   # left-canonicalise
        l, d, r = psi[b].shape
        Q, R = linalg.qr(psi[b].reshape(l * d, r, order="F"))
        psi[b] = Q.reshape(l, d, Q.shape[1], order="F")
        psi[b + 1] = np.tensordot(R, psi[b + 1], axes=(1, 0))
    Bs = [None] * N
    lams = [None] * (N + 1)
    lams[0] = np.ones(1)
    lams[N] = np.ones(1)
    C = psi[N - 1]
    for j in range(N - 1, 0, -1):                            # SVD sweep back: exact Schmidt decompositions
        l, d, r = C.shape
        U, S, Vh = np.linalg.svd(C.reshape(l, d * r, order="F"), full_matrices=False)
        Bs[j] = Vh.reshape(len(S), d, r, order="F")
        lams[j] = S / np.linalg.norm(S)
        C = np.tensordot(psi[j - 1], U * S[None, :], axes=(2, 0))
    Bs[0] = C / np.linalg.norm(C)
    return Bs, lams


def apply_gate_bform(Bs, lams, G, n, maxdim=None, cutoff=None, mindim=1):
    """Two-site gate G[s1',s2',s1,s2] on (n, n+1), in place.  Returns the truncation error."""
    B1, B2, lamL = Bs[n], Bs[n + 1], lams[n]
    tt = np.tensordot(B1, B2, axes=(2, 0))                           # (l,s1,s2,r), no Schmidt weights
    tt = np.transpose(np.tensordot(tt, G, axes=([1, 2], [2, 3])), (0, 2, 3, 1))
    l, d1, d2, r = tt.shape
    Mt = tt.reshape(l * d1, d2 * r, order="F")
    M = (lamL[:, None, None, None] * tt).reshape(l * d1, d2 * r, order="F")
    U, S, V, spec = linalg.svd(M, maxdim=maxdim, mindim=mindim, cutoff=cutoff)      # M ~ U diag(S) V^T
    k = len(S)
    nrm = np.linalg.norm(S)
    Bs[n + 1] = V.T.reshape(k, d2, r, order="F")
    Bs[n] = ((Mt @ V.conj()) / nrm).reshape(l, d1, k, order="F")
    lams[n + 1] = S / nrm
    return spec.truncerr


def tebd_layer_bform(Bs, lams, G, parity, maxdim=None, cutoff=None):
    """All gates of one layer (bonds parity, parity+2, ...): order is irrelevant, they share nothing."""
    N = len(Bs)
    errs = [apply_gate_bform(Bs, lams, G, n, maxdim=maxdim, cutoff=cutoff) for n in range(parity, N - 1, 2)]
    return max(errs) if errs else 0.0
