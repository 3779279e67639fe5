"""NumPy prototype of the TWO-STAGE tridiagonalisation planned for round 2.
Design aid only -- not shipped, not imported by the package.

Why: the one-stage reduction of csrc/tridiag.cu is bound by an n^3/3 * 8-byte matrix-vector stream
(256 ms at n = 8192).  The two-stage form moves almost all flops onto the FP64 tensor pipe:

  stage 1  full -> band of half-width b   (successive band reduction): per block column a QR of the (n-kb-b) x b
           panel below the band and a two-sided compact-WY update of the trailing matrix -- 4/3 n^3 flop, all GEMM
           (about 25 ms at n = 8192 on the 34 TFLOP/s contraction kernel);
  stage 2  band -> tridiagonal by bulge chasing: column j of the band is annihilated below its first subdiagonal by
           one Householder reflector of length <= b, and the bulge this creates is chased down the band by further
           reflectors of length <= b  (6 n^2 b flop, all of it on b x b .. 3b x b windows that fit shared memory;
           sweeps j and j+1 are independent once they are 2 windows apart, so about n/(3b) sweeps are in flight);
  back     eigenvectors: Z <- Q1 (Q2 Z), Q2 = product of the n^2/(2b) short stage-2 reflectors (grouped into
           diamond-shaped blocks they become GEMMs), Q1 = the n/b block reflectors of stage 1.

This file fixes the index conventions and checks that the composition reproduces A (and that the eigenvectors of T
back-transform to eigenvectors of A) before any CUDA is written, in the same way tools/proto_eigh.py did for the
one-stage solver.  Real symmetric and complex Hermitian."""
import numpy as np


def house(x):
    """LAPACK larfg: H = I - tau v v^H with v[0] = 1, H^H x = beta e_1, beta real."""
    x = np.asarray(x)
    alpha = x[0]
    xn2 = np.sum(np.abs(x[1:]) ** 2)
    if xn2 == 0.0 and (not np.iscomplexobj(x) or alpha.imag == 0.0):
        v = np.zeros_like(x); v[0] = 1.0
        return v, 0.0 * alpha, alpha.real
    nrm = np.sqrt(abs(alpha) ** 2 + xn2)
    beta = -nrm if alpha.real >= 0 else nrm
    tau = (beta - alpha) / beta
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, tau, beta


def apply_two_sided(A, v, tau, r0):
    """A <- H^H A H with H = I - tau v v^H acting on rows/cols r0 .. r0+len(v)-1 (dense; the kernel touches a window)."""
    k = len(v)
    sl = slice(r0, r0 + k)
    A[sl, :] -= np.conj(tau) * np.outer(v, v.conj() @ A[sl, :])
    A[:, sl] -= tau * np.outer(A[:, sl] @ v, v.conj())


def to_band(A, b):
    """Stage 1.  Returns the band matrix (full storage, entries outside the band zeroed) and the block reflectors
    [(r0, V, T)] with Q1 = prod_k (I - V_k T_k V_k^H) acting on rows r0.."""
    A = A.copy()
    n = A.shape[0]
    refl = []
    for k0 in range(0, n - b - 1, b):
        r0 = k0 + b
        P = A[r0:, k0:k0 + b].copy()                   # panel below the band
        m, w = P.shape
        V = np.zeros((m, w), dtype=A.dtype)
        taus = np.zeros(w, dtype=A.dtype)
        for c in range(min(w, m - 1) if m > 1 else 0):
            v, tau, beta = house(P[c:, c])
            V[c:, c] = v
            taus[c] = tau
            P[c:, c:] -= np.conj(tau) * np.outer(v, v.conj() @ P[c:, c:])
        Tm = np.zeros((w, w), dtype=A.dtype)            # larft
        G = V.conj().T @ V
        for c in range(w):
            Tm[c, c] = taus[c]
            if c:
                Tm[:c, c] = -taus[c] * (Tm[:c, :c] @ G[:c, c])
        # two-sided update of everything from row/col r0 on (on the GPU: 3 GEMMs, zher2k-like)
        Qb = np.eye(m, dtype=A.dtype) - V @ Tm @ V.conj().T
        A[r0:, :] = Qb.conj().T @ A[r0:, :]
        A[:, r0:] = A[:, r0:] @ Qb
        refl.append((r0, V, Tm))
    # clean numerical fuzz outside the band
    i, j = np.indices(A.shape)
    A[np.abs(i - j) > b] = 0.0
    return A, refl


def band_to_tridiag(B, b):
    """Stage 2 (bulge chasing).  Returns d, e and the list of short reflectors (r0, v, tau) in application order."""
    A = B.copy()
    n = A.shape[0]
    refl = []
    for j in range(n - 2):
        # annihilate A[j+2 : j+b+1, j]
        r0 = j + 1
        r1 = min(n, j + b + 1)
        if r1 - r0 > 1:
            v, tau, beta = house(A[r0:r1, j])
            if tau != 0:
                apply_two_sided(A, v, tau, r0)
                refl.append((r0, v, tau))
            # chase the bulge: the update filled A[r1 : r1+b, r0 : r1] (below the band); restore column by column block
            c0 = r0
            while True:
                q0 = c0 + b                                  # first row of the bulge block
                q1 = min(n, q0 + b)
                if q1 - q0 <= 1 and not (q0 < n and np.any(np.abs(A[q0 + 1:q1, c0]) > 0)):
                    if q0 >= n - 1:
                        break
                if q0 >= n:
                    break
                # eliminate the first column of the bulge (column c0) below row q0
                if q1 - q0 > 1:
                    v2, tau2, _ = house(A[q0:q1, c0])
                    if tau2 != 0:
                        apply_two_sided(A, v2, tau2, q0)
                        refl.append((q0, v2, tau2))
                c0 = q0
    d = np.real(np.diag(A)).copy()
    e = np.real(np.diag(A, -1)).copy()
    off = A - np.diag(np.diag(A)) - np.diag(np.diag(A, -1), -1) - np.diag(np.diag(A, 1), 1)
    return d, e, refl, np.linalg.norm(off), A


def back_transform(refl1, refl2, Z):
    """X = Q1 Q2 Z."""
    X = Z.astype(np.result_type(Z, refl1[0][1] if refl1 else Z)).copy()
    for (r0, v, tau) in reversed(refl2):
        k = len(v)
        X[r0:r0 + k, :] -= tau * np.outer(v, v.conj() @ X[r0:r0 + k, :])
    for (r0, V, Tm) in reversed(refl1):
        X[r0:, :] -= V @ (Tm @ (V.conj().T @ X[r0:, :]))
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for cplx in (False, True):
        for n, b in ((40, 4), (97, 8), (200, 16), (257, 32)):
            X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
            A = X + X.conj().T
            Bm, r1 = to_band(A, b)
            w_band = np.linalg.eigvalsh(Bm)
            d, e, r2, off, Tfull = band_to_tridiag(Bm, b)
            # the subdiagonal of a complex Hermitian band stays complex after real-beta reflectors only in its last entry
            T = np.diag(d) + np.diag(np.diag(Tfull, -1), -1) + np.diag(np.diag(Tfull, 1), 1)
            w = np.linalg.eigvalsh(A)
            wt, Zt = np.linalg.eigh(T)
            U = back_transform(r1, r2, Zt)
            res = np.linalg.norm(A @ U - U * wt[None, :]) / np.linalg.norm(A)
            orth = np.linalg.norm(U.conj().T @ U - np.eye(n))
            print(f"cplx={cplx} n={n:4d} b={b:3d}  band eig {np.max(np.abs(w_band - w)) / np.max(np.abs(w)):.1e}  off-tridiagonal {off:.1e}"
                  f"  eig {np.max(np.abs(wt - w)) / np.max(np.abs(w)):.1e}  res {res:.1e}  orth {orth:.1e}  stage-2 reflectors {len(r2)} (n^2/2b = {n * n // (2 * b)})")
