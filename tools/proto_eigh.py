"""NumPy prototype of the fast Hermitian eigensolver that csrc/tridiag.cu + csrc/stedc.cu implement:
   1. blocked Householder tridiagonalisation (latrd/sytrd structure, full symmetric storage),
   2. Cuppen divide & conquer on the real tridiagonal (deflation, secular equation with
      origin shift, Gu-Eisenstat/Loewner z-hat),
   3. compact-WY back-transformation.
Design aid only (not shipped, not imported by the package): it fixes formulas, conjugation
conventions and tolerances on the CPU, where they can be debugged without a GPU round trip.
The structure (per-column K1/K2 steps, per-merge m1..m6 steps) mirrors the kernels 1:1."""
import numpy as np

EPS = np.finfo(np.float64).eps


# ----------------------------------------------------------------------------- stage 1
def tridiagonalize(A, nb=64):
    """A Hermitian (full storage).  Returns d, e, Vst (reflectors in columns, explicit 1 at the pivot,
    zeros above), tau.  A = Q T Q^H, Q = H_0 H_1 ... H_{n-2}, H_i = I - tau_i v_i v_i^H."""
    A = A.copy()
    n = A.shape[0]
    cplx = np.iscomplexobj(A)
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(n, dtype=A.dtype)
    Vst = np.zeros((n, n), dtype=A.dtype)
    p0 = 0
    while p0 < n - 1:
        jb = min(nb, n - 1 - p0)
        V = np.zeros((n, jb), dtype=A.dtype)
        W = np.zeros((n, jb), dtype=A.dtype)
        for j in range(jb):
            i = p0 + j
            # K1: column update (rows i..n-1) with the panel so far
            a = A[i:, i].copy()
            if j > 0:
                a -= V[i:, :j] @ np.conj(W[i, :j]) + W[i:, :j] @ np.conj(V[i, :j])
            d[i] = a[0].real
            x = a[1:]
            alpha = x[0]
            xn2 = np.sum(np.abs(x[1:]) ** 2)
            if xn2 == 0.0 and (not cplx or alpha.imag == 0.0):
                t = 0.0
                beta = alpha.real
                v = np.zeros_like(x)
                v[0] = 1.0
            else:
                nrm = np.sqrt(abs(alpha) ** 2 + xn2)
                beta = -nrm if alpha.real >= 0 else nrm
                t = (beta - alpha) / beta if not cplx else complex((beta - alpha.real) / beta, -alpha.imag / beta)
                v = x / (alpha - beta)
                v[0] = 1.0
            e[i] = beta
            tau[i] = t
            V[i + 1:, j] = v
            Vst[i + 1:, i] = v
            # K2: column-dots with v over rows T = i+1..n-1:  y_r = A[T,r]^H v,  p = W^H v, q = V^H v
            At = A[i + 1:, i + 1:]
            y = At.conj().T @ v                  # = A_trail v for Hermitian A_trail
            p = W[i + 1:, :j].conj().T @ v
            q = V[i + 1:, :j].conj().T @ v
            yhv = np.vdot(y, v)                  # y^H v
            # w = tau (y - V p - W q) + alpha2 v ;  alpha2 = -1/2 tau (w'^H v)
            wp_h_v = np.conj(t) * (yhv - np.vdot(p, q) - np.vdot(q, p))
            alpha2 = -0.5 * t * wp_h_v
            w = t * (y - V[i + 1:, :j] @ p - W[i + 1:, :j] @ q) + alpha2 * v
            W[i + 1:, j] = w
        # trailing update
        lo = p0 + jb
        A[lo:, lo:] -= V[lo:, :] @ W[lo:, :].conj().T + W[lo:, :] @ V[lo:, :].conj().T
        p0 += jb
    d[n - 1] = A[n - 1, n - 1].real
    return d, e, Vst, tau


def back_transform(Vst, tau, X, nb=64):
    """X <- Q X with Q = H_0 ... H_{n-2} (panels applied last to first, compact WY)."""
    n = Vst.shape[0]
    X = X.astype(Vst.dtype).copy()
    starts = list(range(0, n - 1, nb))
    for p0 in reversed(starts):
        jb = min(nb, n - 1 - p0)
        V = Vst[:, p0:p0 + jb]
        G = V.conj().T @ V
        T = np.zeros((jb, jb), dtype=Vst.dtype)
        for c in range(jb):
            T[c, c] = tau[p0 + c]
            if c > 0:
                T[:c, c] = -tau[p0 + c] * (T[:c, :c] @ G[:c, c])
        X -= V @ (T @ (V.conj().T @ X))
    return X


# ----------------------------------------------------------------------------- stage 2
def secular_roots(dk, zk, rho, maxit=80):
    """Roots of 1 + rho sum z_j^2/(d_j - lam) = 0, d ascending strictly, rho > 0, all z != 0.
    Returns (origin index o_i, mu_i): lam_i = d[o_i] + mu_i.  One 'thread' per root (vectorised)."""
    k = len(dk)
    z2 = zk * zk
    org = np.zeros(k, dtype=np.int64)
    mu = np.zeros(k)
    nits = 0
    for i in range(k):
        last = (i == k - 1)
        if not last:
            gap = dk[i + 1] - dk[i]
            mid = 0.5 * gap
            # f at midpoint with origin d_i
            dl = (dk - dk[i]) - mid
            fm = 1.0 + rho * np.sum(z2 / dl)
            if fm > 0:
                o = i; lo, hi = 0.0, mid
            else:
                o = i + 1; lo, hi = -mid, 0.0
        else:
            o = i; lo, hi = 0.0, rho * np.sum(z2)
            gap = hi
        delta = dk - dk[o]
        x = 0.5 * (lo + hi) if not last else hi * 0.5
        if last:
            # f(hi) >= 0 always; start from the one-pole estimate
            pass
        ip, iq = (i, i + 1) if not last else (i - 1, i)
        for it in range(maxit):
            nits += 1
            D = delta - x
            t = z2 / D
            if not last:
                psi = rho * np.sum(t[:i + 1]); dpsi = rho * np.sum(t[:i + 1] / D[:i + 1])
                phi = rho * np.sum(t[i + 1:]); dphi = rho * np.sum(t[i + 1:] / D[i + 1:])
            else:
                psi = rho * np.sum(t); dpsi = rho * np.sum(t / D)
                phi = 0.0; dphi = 0.0
            f = 1.0 + psi + phi
            err = 8.0 * EPS * (1.0 + abs(psi) + abs(phi)) * 1.0 + EPS * abs(x) * (dpsi + dphi)
            if f > 0: hi = min(hi, x)
            else: lo = max(lo, x)
            if abs(f) <= err or hi - lo <= 2 * EPS * max(abs(lo), abs(hi)):
                break
            # two-pole rational model step
            if not last:
                Di, Dj = D[i], D[i + 1]
                a = dpsi * Di * Di; s = psi - dpsi * Di
                b = dphi * Dj * Dj; tt = phi - dphi * Dj
                c = 1.0 + s + tt
                # c (Di-eta)(Dj-eta) + a (Dj-eta) + b (Di-eta) = 0
                qa = c
                qb = -(c * (Di + Dj) + a + b)
                qc = Di * Dj * f
            else:
                Di = D[i]
                a = dpsi * Di * Di; s = psi - dpsi * Di
                c = 1.0 + s
                # c (Di - eta) + a = 0
                qa = 0.0; qb = -c; qc = c * Di + a
            eta = None
            if qa == 0.0:
                if qb != 0.0: eta = -qc / qb
            else:
                disc = qb * qb - 4 * qa * qc
                if disc >= 0:
                    sq = np.sqrt(disc)
                    # both roots; choose the one keeping x+eta inside (lo,hi)
                    qq = -0.5 * (qb + (sq if qb >= 0 else -sq))
                    cands = []
                    if qq != 0: cands.append(qc / qq)
                    cands.append(qq / qa)
                    cands = [cc for cc in cands if lo < x + cc < hi]
                    if cands:
                        eta = min(cands, key=abs)
            xn = x + eta if eta is not None else None
            if xn is None or not (lo < xn < hi) or not np.isfinite(xn):
                xn = 0.5 * (lo + hi)
            if xn == x:
                break
            x = xn
        org[i] = o
        mu[i] = x
    return org, mu, nits


def merge(d1, Q1, d2, Q2, rho_e):
    """One Cuppen merge.  T = [[T1, rho_e e e^T],[.., T2]] where the halves were solved for
    T1' = T1 - |rho_e| e_last e_last^T, T2' = T2 - |rho_e| e_1 e_1^T."""
    n1, n2 = len(d1), len(d2)
    N = n1 + n2
    sgn = 1.0 if rho_e >= 0 else -1.0
    z = np.concatenate([Q1[-1, :], sgn * Q2[0, :]]) / np.sqrt(2.0)
    rho = 2.0 * abs(rho_e)
    d = np.concatenate([d1, d2])
    Q = np.zeros((N, N))
    Q[:n1, :n1] = Q1
    Q[n1:, n1:] = Q2
    order = np.argsort(d, kind="stable")
    ds = d[order].copy(); zs = z[order].copy(); col = order.copy()
    tol = 8.0 * EPS * max(np.max(np.abs(ds)), np.max(np.abs(zs)))
    if rho * np.max(np.abs(zs)) <= tol:
        return ds, Q[:, col], dict(k=0, N=N)
    defl = np.zeros(N, dtype=bool)
    pj = -1
    rots = []
    for t in range(N):
        if rho * abs(zs[t]) <= tol:
            defl[t] = True
            continue
        if pj < 0:
            pj = t
            continue
        s = zs[pj]; c = zs[t]
        tau = np.hypot(c, s)
        tt = ds[t] - ds[pj]
        c /= tau; s = -s / tau
        if abs(tt * c * s) <= tol:
            zs[t] = tau; zs[pj] = 0.0
            rots.append((col[pj], col[t], c, s))
            qp = Q[:, col[pj]].copy(); qn = Q[:, col[t]].copy()
            Q[:, col[pj]] = c * qp + s * qn
            Q[:, col[t]] = -s * qp + c * qn
            tnew = ds[pj] * c * c + ds[t] * s * s
            ds[t] = ds[pj] * s * s + ds[t] * c * c
            ds[pj] = tnew
            defl[pj] = True
            pj = t
        else:
            pj = t
    nd = np.where(~defl)[0]
    k = len(nd)
    dk = ds[nd]; zk = zs[nd]
    # after rotations dk may be slightly out of order?  (LAPACK keeps order because |t c s| <= tol moves are tiny)
    org, mu, nits = secular_roots(dk, zk, rho)
    # Loewner z-hat:  zhat_j^2 = prod_i (lam_i - d_j) / (rho * prod_{i != j} (d_i - d_j))
    lam_minus_d = (dk[org][:, None] - dk[None, :]) + mu[:, None]      # [i, j] = lam_i - d_j
    dd = dk[:, None] - dk[None, :]                                     # [i, j] = d_i - d_j
    np.fill_diagonal(dd, 1.0)
    ratio = lam_minus_d / dd
    zhat2 = np.prod(ratio, axis=0) / rho
    zhat = np.sign(zk) * np.sqrt(np.abs(zhat2))
    U = zhat[:, None] / (-lam_minus_d.T)                               # [j, i] = zhat_j / (d_j - lam_i)
    U /= np.linalg.norm(U, axis=0)[None, :]
    lam = dk[org] + mu
    Qnd = Q[:, col[nd]] @ U
    dnew = np.concatenate([lam, ds[defl]])
    Qnew = np.concatenate([Qnd, Q[:, col[defl]]], axis=1)
    o2 = np.argsort(dnew, kind="stable")
    return dnew[o2], Qnew[:, o2], dict(k=k, N=N, nits=nits, nrot=len(rots))


def stedc(d, e, leaf=32, stats=None):
    """Scale to unit max-norm first (as LAPACK dstedc does): the deflation tolerance 8 eps max(|d|, |z|) compares
    eigenvalue-scale quantities with the O(1) entries of z and is only meaningful for |T| ~ 1."""
    d = np.asarray(d, float); e = np.asarray(e, float)
    s = max(np.max(np.abs(d), initial=0.0), np.max(np.abs(e), initial=0.0))
    if s == 0.0:
        return d.copy(), np.eye(len(d))
    w, Q = _stedc_unit(d / s, e / s, leaf, stats)
    return w * s, Q


def _stedc_unit(d, e, leaf=32, stats=None):
    n = len(d)
    if n <= leaf:
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        w, Q = np.linalg.eigh(T)
        return w, Q
    n1 = n // 2
    r = e[n1 - 1]
    d1 = d[:n1].copy(); d2 = d[n1:].copy()
    d1[-1] -= abs(r); d2[0] -= abs(r)
    w1, Q1 = _stedc_unit(d1, e[:n1 - 1], leaf, stats)
    w2, Q2 = _stedc_unit(d2, e[n1:], leaf, stats)
    w, Q, st = merge(w1, Q1, w2, Q2, r)
    if stats is not None:
        stats.append(st)
    return w, Q


def eigh_fast(A, nb=64, leaf=32):
    d, e, Vst, tau = tridiagonalize(A, nb)
    w, Z = stedc(d, e, leaf)
    X = back_transform(Vst, tau, Z, nb)
    return w, X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for cplx in (False, True):
        for n in (1, 2, 3, 5, 33, 70, 200, 515):
            X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
            A = X + X.conj().T
            d, e, Vst, tau = tridiagonalize(A, nb=16)
            T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
            Q = back_transform(Vst, tau, np.eye(n), nb=16)
            err = np.linalg.norm(Q @ T @ Q.conj().T - A) / max(np.linalg.norm(A), 1e-300)
            w, U = eigh_fast(A, nb=16, leaf=8)
            wr = np.linalg.eigvalsh(A)
            res = np.linalg.norm(A @ U - U * w[None, :]) / max(np.linalg.norm(A), 1e-300)
            orth = np.linalg.norm(U.conj().T @ U - np.eye(n))
            print(f"cplx={cplx} n={n:4d} tridiag {err:.1e} eig {np.max(np.abs(w - wr)) / max(np.max(np.abs(wr)),1e-300):.1e} res {res:.1e} orth {orth:.1e}")
    # DMRG-like density matrix: rapidly decaying spectrum, big null space
    for n in (256, 600):
        X = rng.standard_normal((n, n))
        U0, _, V0 = np.linalg.svd(X)
        s = 10.0 ** (-np.arange(n) / 8.0)
        M = (U0 * s) @ V0
        A = M @ M.T
        stats = []
        d, e, Vst, tau = tridiagonalize(A, nb=32)
        w, Z = stedc(d, e, 32, stats)
        U = back_transform(Vst, tau, Z, 32)
        wr = np.linalg.eigvalsh(A)
        res = np.linalg.norm(A @ U - U * w[None, :]) / np.linalg.norm(A)
        orth = np.linalg.norm(U.T @ U - np.eye(n))
        print(f"psd n={n} eig {np.max(np.abs(w - wr)) / wr[-1]:.1e} res {res:.1e} orth {orth:.1e}", [(s['N'], s['k']) for s in stats[-3:]],
              "avg its", np.mean([s.get('nits', 0) / max(s['k'], 1) for s in stats]))
    # clustered / glued Wilkinson-like tridiagonals
    for n in (100, 401):
        d = np.abs(np.arange(n) - n // 2).astype(float)
        e = np.ones(n - 1)
        stats = []
        w, Z = stedc(d, e, 16, stats)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        print(f"wilkinson n={n} res {np.linalg.norm(T @ Z - Z * w[None, :]):.1e} orth {np.linalg.norm(Z.T @ Z - np.eye(n)):.1e}",
              "avg its", np.mean([s.get('nits', 0) / max(s['k'], 1) for s in stats]))
