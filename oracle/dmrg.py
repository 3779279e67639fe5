"""Oracle: ProjMPO (H_eff), Lanczos / Davidson eigensolvers and two-site DMRG.

TEST INFRASTRUCTURE.  Everything here restates [EXT] code the reference *reaches* but does
not contain (ITensors.jl 0.2 ``src/mps/projmpo.jl``, ``src/mps/dmrg.jl``,
``src/iterativesolvers.jl``; KrylovKit ``eigsolve``), anchored on the reference's call
sites ``examples/dmrg.jl:25``, ``test/dmrg.jl:27,75``,
``test/test_cuiterativesolvers.jl:21,25``.  Each pairwise contraction is one
``np.tensordot`` in the *same order* ITensors issues them, so this file is also the
"reference CPU path" timed by ``bench.py --impl reference``.

Layouts: MPS ``A[l,s,r]``; MPO ``W[a,s,s',b]``; environments ``L[l,l',a]``, ``R[r,r',c]``
(unprimed = ket side, primed = bra/output side); two-site tensor ``phi[l,s1,s2,r]``.
"""
import numpy as np

from . import linalg


# --------------------------------------------------------------------------- ProjMPO
def heff_apply(L, W1, W2, R, phi):
    """H_eff*phi = noprime((((phi*L)*W_b)*W_{b+1})*R)  -- [EXT] ``product(::ProjMPO, v)``.

    Four pairwise contractions:
      1. T1[s1,s2,r,l',a]   = sum_l  phi[l,s1,s2,r] L[l,l',a]
      2. T2[s2,r,l',s1',b]  = sum_{a,s1} T1 W1[a,s1,s1',b]
      3. T3[r,l',s1',s2',c] = sum_{b,s2} T2 W2[b,s2,s2',c]
      4. out[l',s1',s2',r'] = sum_{r,c}  T3 R[r,r',c]
    """
    T = np.tensordot(phi, L, axes=(0, 0))
    T = np.tensordot(T, W1, axes=([4, 0], [0, 1]))
    T = np.tensordot(T, W2, axes=([4, 0], [0, 1]))
    T = np.tensordot(T, R, axes=([0, 4], [0, 2]))
    return T


def heff_flops(chiL, chiR, d, w1, w2=None, w3=None):
    """Real multiply-add flops of one heff_apply (2*M*N*K per pairwise contraction).
    For uniform w this is SURVEY.md's F = 2d^2w(chiL^2 chiR + chiL chiR^2) + 4d^3w^2 chiL chiR."""
    w2 = w1 if w2 is None else w2
    w3 = w1 if w3 is None else w3
    f1 = 2.0 * (d * d * chiR) * (chiL * w1) * chiL
    f2 = 2.0 * (d * chiR * chiL) * (d * w2) * (w1 * d)
    f3 = 2.0 * (chiR * chiL * d) * (d * w3) * (w2 * d)
    f4 = 2.0 * (chiL * d * d) * chiR * (chiR * w3)
    return f1 + f2 + f3 + f4


def env_left_update(L, A, W):
    """L_{j+1}[r,r',b] = ((L*A)*W)*conj(A')   -- [EXT] ``makeL!``."""
    T = np.tensordot(L, A, axes=(0, 0))                    # (l',a,s,r)
    T = np.tensordot(T, W, axes=([1, 2], [0, 1]))          # (l',r,s',b)
    T = np.tensordot(T, np.conj(A), axes=([0, 2], [0, 1]))  # (r,b,r')
    return np.ascontiguousarray(np.transpose(T, (0, 2, 1)))


def env_right_update(R, A, W):
    """R_{j-1}[l,l',a] = ((R*A)*W)*conj(A')   -- [EXT] ``makeR!``."""
    T = np.tensordot(R, A, axes=(0, 2))                    # (r',c,l,s)
    T = np.tensordot(T, W, axes=([1, 3], [3, 1]))          # (r',l,a,s')
    T = np.tensordot(T, np.conj(A), axes=([0, 3], [2, 1]))  # (l,a,l')
    return np.ascontiguousarray(np.transpose(T, (0, 2, 1)))


def noise_term(L, W1, W2, R, phi, ortho):
    """[EXT] ``noiseterm(P::ProjMPO, phi, ortho)`` -> density-matrix perturbation over the
    kept side's (bond, site) pair, as a matrix matching ``factorize``'s A2."""
    if ortho == "left":
        AL = np.tensordot(L, W1, axes=(2, 0))              # (l,l',s1,s1',b)
        nt = np.tensordot(AL, phi, axes=([0, 2], [0, 1]))   # (l',s1',b,s2,r)
        m = nt.shape[0] * nt.shape[1]
        M = nt.reshape(m, -1, order="F")
    else:
        AR = np.tensordot(W2, R, axes=(3, 2))              # (b,s2,s2',r,r')
        nt = np.tensordot(phi, AR, axes=([2, 3], [1, 3]))   # (l,s1,b,s2',r')
        nt = np.transpose(nt, (3, 4, 0, 1, 2))              # (s2',r',l,s1,b)
        m = nt.shape[0] * nt.shape[1]
        M = nt.reshape(m, -1, order="F")
    return M @ M.conj().T


# --------------------------------------------------------------------------- eigensolvers
def lanczos(matvec, v0, krylovdim=3, maxiter=1, tol=1e-14):
    """[EXT] ``KrylovKit.eigsolve(A, x0, 1, :SR; ishermitian=true, krylovdim, maxiter, tol)``.

    One Krylov cycle per iteration (ITensors 0.2 ``dmrg`` uses krylovdim=3, maxiter=1 -> 3
    matvecs, no restart).  Full re-orthogonalisation (modified Gram-Schmidt) each step.
    Returns (lambda, x, n_matvec) with ||x|| = 1.
    """
    x = v0
    nmv = 0
    lam = None
    for _ in range(maxiter):
        V = []
        alphas, betas = [], []
        nrm = np.linalg.norm(x.ravel())
        v = x / nrm
        beta = 0.0
        while True:
            V.append(v)
            w = matvec(v)
            nmv += 1
            a = np.vdot(v.ravel(), w.ravel()).real
            alphas.append(a)
            w = w - a * v
            if len(V) > 1:
                w = w - betas[-1] * V[-2]
            for u in V:  # re-orthogonalise
                w = w - np.vdot(u.ravel(), w.ravel()) * u
            beta = np.linalg.norm(w.ravel())
            if len(V) == krylovdim or beta <= tol:
                break
            betas.append(beta)
            v = w / beta
        k = len(V)
        T = np.diag(alphas) + np.diag(betas[: k - 1], 1) + np.diag(betas[: k - 1], -1)
        ev, U = np.linalg.eigh(T)
        lam = float(ev[0])
        y = U[:, 0]
        x = sum(y[i] * V[i] for i in range(k))
        if abs(beta * y[-1]) < tol:
            break
    return lam, x / np.linalg.norm(x.ravel()), nmv


def davidson(matvec, v0, maxiter=2, miniter=1, errgoal=1e-14, approx0=1e-12):
    """[EXT] ITensors ``davidson(A, phi0; maxiter=2)``; pinned by the residual test
    ``test/test_cuiterativesolvers.jl:22,26``.  No preconditioner is available for an
    implicit H_eff, so the correction vector is the (orthogonalised) residual."""
    phi = v0 / np.linalg.norm(v0.ravel())
    V = [phi]
    AV = [matvec(phi)]
    lam = np.vdot(V[0].ravel(), AV[0].ravel()).real
    q = AV[0] - lam * V[0]
    M = np.array([[lam]], dtype=complex)
    last = lam
    for ni in range(1, maxiter + 1):
        qn = np.linalg.norm(q.ravel())
        if qn < max(approx0, errgoal * 1e-3) and ni > miniter:
            break
        for _ in range(2):
            for u in V:
                q = q - np.vdot(u.ravel(), q.ravel()) * u
        qn = np.linalg.norm(q.ravel())
        if qn < 1e-10:
            break
        q = q / qn
        V.append(q)
        AV.append(matvec(q))
        k = len(V)
        Mn = np.zeros((k, k), dtype=complex)
        Mn[: k - 1, : k - 1] = M
        for i in range(k):
            Mn[i, k - 1] = np.vdot(V[i].ravel(), AV[k - 1].ravel())
            Mn[k - 1, i] = np.conj(Mn[i, k - 1])
        M = Mn
        ev, U = np.linalg.eigh(M)
        lam = float(ev[0])
        y = U[:, 0]
        if not np.iscomplexobj(V[0]) and not np.iscomplexobj(AV[0]):
            y = y.real if np.allclose(y.imag, 0) else y
        phi = sum(y[i] * V[i] for i in range(k))
        q = sum(y[i] * AV[i] for i in range(k)) - lam * phi
        if abs(lam - last) < errgoal and ni >= miniter and np.linalg.norm(q.ravel()) < np.sqrt(errgoal):
            break
        last = lam
    return lam, phi / np.linalg.norm(phi.ravel())


# --------------------------------------------------------------------------- DMRG
class Sweeps:
    """[EXT] ITensors ``Sweeps``: per-sweep maxdim/mindim/cutoff/noise, last value repeats
    (``examples/dmrg.jl:20-24``)."""

    def __init__(self, nsweep, maxdim=(1,), mindim=(1,), cutoff=(0.0,), noise=(0.0,)):
        self.nsweep = nsweep
        ext = lambda v: [list(np.atleast_1d(v))[min(i, len(np.atleast_1d(v)) - 1)] for i in range(nsweep)]
        self.maxdim = [int(x) for x in ext(maxdim)]
        self.mindim = [int(x) for x in ext(mindim)]
        self.cutoff = [float(x) for x in ext(cutoff)]
        self.noise = [float(x) for x in ext(noise)]


def build_right_envs(psi, Ws, upto=1):
    """R[j] = environment to the right of site j (covers sites j+1..N-1), for j >= upto."""
    N = len(psi)
    dt = np.result_type(psi[0], Ws[0])
    Rs = [None] * N
    Rs[N - 1] = np.ones((1, 1, 1), dtype=dt)
    for j in range(N - 1, upto, -1):
        Rs[j - 1] = env_right_update(Rs[j], psi[j], Ws[j])
    return Rs


def replacebond(phi, ortho, maxdim, mindim, cutoff, drho=None, which_decomp=None, normalize=True):
    """[EXT] ``replacebond!`` -> ``factorize`` (pinned by ``test/test_cumps.jl:153-179``)."""
    l, d1, d2, r = phi.shape
    M = phi.reshape(l * d1, d2 * r, order="F")
    Lm, Rm, spec = linalg.factorize(M, ortho=ortho, which_decomp=which_decomp, maxdim=maxdim,
                                    mindim=mindim, cutoff=cutoff, eigen_perturbation=drho)
    k = Lm.shape[1]
    A = Lm.reshape(l, d1, k, order="F")
    B = Rm.reshape(k, d2, r, order="F")
    if normalize:
        if ortho == "left":
            B = B / np.linalg.norm(B.ravel())
        else:
            A = A / np.linalg.norm(A.ravel())
    return A, B, spec


def bond_step(Lenv, W1, W2, Renv, A1, A2, ortho, maxdim, mindim=1, cutoff=0.0, noise=0.0,
              krylovdim=3, maxiter=1, which_decomp=None, eigsolver="lanczos"):
    """One two-site DMRG bond update.  Returns (energy, A1', A2', spec, n_matvec)."""
    phi = np.tensordot(A1, A2, axes=(2, 0))
    mv = lambda v: heff_apply(Lenv, W1, W2, Renv, v)
    if eigsolver == "lanczos":
        energy, phi, nmv = lanczos(mv, phi, krylovdim=krylovdim, maxiter=maxiter)
    else:
        energy, phi = davidson(mv, phi, maxiter=maxiter)
        nmv = maxiter + 1
    drho = None
    if noise > 0.0:
        drho = noise * noise_term(Lenv, W1, W2, Renv, phi, ortho)
    A, B, spec = replacebond(phi, ortho, maxdim, mindim, cutoff, drho=drho, which_decomp=which_decomp)
    return energy, A, B, spec, nmv


def dmrg(Ws, psi0, sweeps, krylovdim=3, maxiter=1, which_decomp=None, eigsolver="lanczos",
         observer=None):
    """[EXT] ITensors 0.2 ``dmrg(H, psi0, sweeps)``.  psi0 must have its orthogonality centre
    at site 0 (``orthogonalize!(psi,1)``).  Returns (energy, psi, energies_per_sweep)."""
    N = len(psi0)
    psi = [A.copy() for A in psi0]
    dt = np.result_type(psi[0], Ws[0])
    Rs = build_right_envs(psi, Ws, upto=1)
    Ls = [None] * N
    Ls[0] = np.ones((1, 1, 1), dtype=dt)
    energy = None
    history = []
    for sw in range(sweeps.nsweep):
        kw = dict(maxdim=sweeps.maxdim[sw], mindim=sweeps.mindim[sw], cutoff=sweeps.cutoff[sw],
                  noise=sweeps.noise[sw], krylovdim=krylovdim, maxiter=maxiter,
                  which_decomp=which_decomp, eigsolver=eigsolver)
        for b in range(0, N - 1):          # left-to-right half sweep, ortho = "left"
            energy, A, B, spec, _ = bond_step(Ls[b], Ws[b], Ws[b + 1], Rs[b + 1], psi[b], psi[b + 1],
                                              "left", **kw)
            psi[b], psi[b + 1] = A, B
            Ls[b + 1] = env_left_update(Ls[b], psi[b], Ws[b])
            if observer:
                observer(sw, b, "left", energy, spec)
        for b in range(N - 2, -1, -1):     # right-to-left half sweep, ortho = "right"
            energy, A, B, spec, _ = bond_step(Ls[b], Ws[b], Ws[b + 1], Rs[b + 1], psi[b], psi[b + 1],
                                              "right", **kw)
            psi[b], psi[b + 1] = A, B
            Rs[b] = env_right_update(Rs[b + 1], psi[b + 1], Ws[b + 1])
            if observer:
                observer(sw, b, "right", energy, spec)
        history.append(energy)
    return energy, psi, history
