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
