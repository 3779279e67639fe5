"""Oracle: site operators, MPO builders and exact diagonalisation.  TEST INFRASTRUCTURE.

MPO tensor layout used everywhere in this repo: ``W[a, s, s', b]`` -- ``a``/``b`` left/right
MPO bond, ``s`` the ket (unprimed, contracted with the MPS) and ``s'`` the primed (output)
site index; the matrix element is ``<s'|O|s>``.  Edge tensors have a = 1 (first site) and
b = 1 (last site).  The models are the reference's workloads: S=1 Heisenberg
(``examples/dmrg.jl:7-15``, ``test/dmrg.jl:5-16``), TFIM (``test/dmrg.jl:58-70``) and the
S=1/2 Heisenberg chain of BASELINE.json config C3.  Operators follow [EXT] ITensors'
"S=1/2"/"S=1" site types (basis ordered Up..Dn).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def spin_ops(S):
    """Sz, S+, S-, Sx, Id for spin S (S = 0.5 or 1) in the basis m = S, S-1, ..., -S."""
    d = int(round(2 * S + 1))
    m = S - np.arange(d)
    Sz = np.diag(m)
    Sp = np.zeros((d, d))
    for i in range(1, d):
        Sp[i - 1, i] = np.sqrt(S * (S + 1) - m[i] * (m[i] + 1))
    Sm = Sp.T.copy()
    return dict(Sz=Sz, Sp=Sp, Sm=Sm, Sx=0.5 * (Sp + Sm), Id=np.eye(d), d=d)


def _place(W, a, b, O):
    W[a, :, :, b] += O.T  # W[a,s,s',b] = <s'|O|s>


def _finish(Wbulk, N):
    """bulk W (a = w-1 is the start row, b = 0 the end column) -> list of N site tensors."""
    out = []
    for j in range(N):
        W = Wbulk
        if j == 0:
            W = W[-1:, :, :, :]
        if j == N - 1:
            W = W[:, :, :, :1]
        out.append(np.ascontiguousarray(W))
    return out


def heisenberg_mpo(N, S=0.5):
    """H = sum_j Sz Sz + 1/2 (S+ S- + S- S+),  MPO bond w = 5."""
    o = spin_ops(S)
    d = o["d"]
    W = np.zeros((5, d, d, 5))
    _place(W, 0, 0, o["Id"])
    _place(W, 1, 0, o["Sm"])
    _place(W, 2, 0, o["Sp"])
    _place(W, 3, 0, o["Sz"])
    _place(W, 4, 1, 0.5 * o["Sp"])
    _place(W, 4, 2, 0.5 * o["Sm"])
    _place(W, 4, 3, o["Sz"])
    _place(W, 4, 4, o["Id"])
    return _finish(W, N)


def tfim_mpo(N, J=-1.0, h=-0.5):
    """H = J sum Sz Sz + h sum Sx  (``test/dmrg.jl:63-67``: J=-1, h=-0.5),  w = 3."""
    o = spin_ops(0.5)
    W = np.zeros((3, 2, 2, 3))
    _place(W, 0, 0, o["Id"])
    _place(W, 1, 0, o["Sz"])
    _place(W, 2, 0, h * o["Sx"])
    _place(W, 2, 1, J * o["Sz"])
    _place(W, 2, 2, o["Id"])
    return _finish(W, N)


def tfim_exact_energy(N):
    """``test/dmrg.jl:79``: critical TFIM, open boundaries."""
    return 0.25 - 0.25 / np.sin(np.pi / (2 * (2 * N + 1)))


def heisenberg_bond_gate(tau, S=0.5, imaginary_time=True):
    """Two-site gate exp(-tau h) (real) or exp(-i tau h) (complex), h = S.S on a bond.
    Returned as G[s1', s2', s1, s2]  (config C4; entry point ``examples/gate_evolution.jl:46``)."""
    o = spin_ops(S)
    d = o["d"]
    h = np.kron(o["Sz"], o["Sz"]) + 0.5 * (np.kron(o["Sp"], o["Sm"]) + np.kron(o["Sm"], o["Sp"]))
    w, v = np.linalg.eigh(h)
    if imaginary_time:
        G = (v * np.exp(-tau * w)[None, :]) @ v.T
    else:
        G = (v * np.exp(-1j * tau * w)[None, :]) @ v.conj().T
    return G.reshape(d, d, d, d)


def mpo_to_sparse(Ws):
    """Contract an MPO into a sparse matrix (small N only) -- for ED cross-checks."""
    d = Ws[0].shape[1]
    N = len(Ws)
    # blocks[a] = operator accumulated so far with open right MPO bond a
    blocks = {0: sp.identity(1, format="csr")}
    for W in Ws:
        new = {}
        for a, Oa in blocks.items():
            for b in range(W.shape[3]):
                M = W[a, :, :, b].T  # <s'|O|s>
                if np.any(M):
                    t = sp.kron(Oa, sp.csr_matrix(M), format="csr")
                    new[b] = new[b] + t if b in new else t
        blocks = new
    assert list(blocks.keys()) == [0]
    H = blocks[0]
    assert H.shape == (d ** N, d ** N)
    return H


def ed_ground_energy(Ws):
    H = mpo_to_sparse(Ws)
    if H.shape[0] <= 512:
        return float(np.linalg.eigvalsh(H.toarray())[0])
    w = spla.eigsh(H, k=1, which="SA", tol=1e-12)[0]
    return float(w[0])
