"""Oracle: rank-2 factorizations with truncation (svd / eigen / qr / factorize).

TEST INFRASTRUCTURE.  LAPACK drivers mirror the CPU path the reference is compared to:
``gesdd`` (Julia ``svd`` default, [EXT]), ``syevr``/``heevr`` (Julia ``eigen(Hermitian)``,
[EXT]), ``geqrf``+``orgqr``.  What is returned follows the reference's GPU methods
(``src/tensor/culinearalgebra.jl``) except where those diverge from the CPU path
(SURVEY.md section 8 a12): here V obeys ``A = U @ diag(S) @ V.T`` (the CPU ``conj!(MV)``
convention; the GPU code leaves it commented out at ``culinearalgebra.jl:44``) and S is
returned as a vector (CPU ``Diag`` storage) rather than a dense dS x dS matrix
(``culinearalgebra.jl:66-68``).
"""
import numpy as np
import scipy.linalg as sla

from .truncate import truncate


class Spectrum:
    """[EXT] ITensors ``Spectrum``: kept eigenvalues (P = S**2 or D) and truncation error."""

    def __init__(self, eigs, truncerr):
        self.eigs = np.asarray(eigs)
        self.truncerr = float(truncerr)


def svd(M, maxdim=None, mindim=1, cutoff=None, use_absolute_cutoff=False,
        use_relative_cutoff=True):
    """Thin SVD + truncation of P = S**2 (``svd``: ``src/tensor/culinearalgebra.jl:33-72``).

    Returns (U[m,k], S[k], V[n,k], spec) with M ~= U @ diag(S) @ V.T.
    Truncation is applied only if maxdim or cutoff is given ([EXT] NDTensors svd).
    """
    M = np.asarray(M)
    U, S, Vh = sla.svd(M, full_matrices=False, lapack_driver="gesdd")
    V = Vh.T  # = conj(MV) of Julia's  M = MU*Diagonal(MS)*MV'
    P = S ** 2
    truncerr = 0.0
    k = len(S)
    if maxdim is not None or cutoff is not None:
        truncerr, _, k = truncate(P, maxdim=maxdim, mindim=mindim,
                                  cutoff=0.0 if cutoff is None else cutoff,
                                  use_absolute_cutoff=use_absolute_cutoff,
                                  use_relative_cutoff=use_relative_cutoff)
    return U[:, :k], S[:k], V[:, :k], Spectrum(P[:k], truncerr)


def eigen(M, maxdim=None, mindim=1, cutoff=None, use_absolute_cutoff=False,
          use_relative_cutoff=True):
    """Hermitian eigendecomposition, eigenvalues DEscending, truncated.

    ``eigen(::Hermitian)``: ``src/tensor/culinearalgebra.jl:74-108`` (syevd/heevd, 'U'
    triangle, ``reverse`` at :90,:96, truncate at :92).  Returns (D[k], U[n,k], spec).
    """
    M = np.asarray(M)
    w, V = sla.eigh(M, lower=False, driver="evr")
    w = w[::-1].copy()
    V = V[:, ::-1]
    k = len(w)
    truncerr = 0.0
    if maxdim is not None or cutoff is not None:
        truncerr, _, k = truncate(w, maxdim=maxdim, mindim=mindim,
                                  cutoff=0.0 if cutoff is None else cutoff,
                                  use_absolute_cutoff=use_absolute_cutoff,
                                  use_relative_cutoff=use_relative_cutoff)
    return w[:k], np.ascontiguousarray(V[:, :k]), Spectrum(w[:k], truncerr)


def qr(M):
    """Thin QR, explicit Q (``qr``: ``src/tensor/culinearalgebra.jl:110-121``).  LAPACK sign
    convention (diag(R) not forced positive), like Julia's ``qr``."""
    Q, R = sla.qr(np.asarray(M), mode="economic")
    return Q, R


def qr_positive(Q, R):
    """Gauge-fix a QR pair so diag(R) >= 0 -- used by tests to compare two QRs."""
    d = np.diagonal(R).copy()
    ph = np.where(np.abs(d) > 0, d / np.where(np.abs(d) > 0, np.abs(d), 1), 1.0)
    return Q * ph[None, :], np.conj(ph)[:, None] * R


def factorize(M, ortho="left", which_decomp=None, maxdim=None, mindim=1, cutoff=None,
              eigen_perturbation=None):
    """[EXT] ITensors ``factorize`` on a matricised tensor M[(left),(right)].

    Rule: eigen if a perturbation is given or cutoff > 1e-12; svd if truncating with
    cutoff <= 1e-12; qr if no truncation is requested.
    Returns (L[m,k], R[k,n], spec) with M ~= L @ R; L is an isometry for ortho="left",
    R has orthonormal rows for ortho="right".
    reached via ``replacebond!`` (``test/test_cumps.jl:153-179``).
    """
    M = np.asarray(M)
    trunc = maxdim is not None or cutoff is not None
    if which_decomp is None:
        if eigen_perturbation is not None:
            which_decomp = "eigen"
        elif not trunc:
            which_decomp = "qr"
        elif (cutoff or 0.0) <= 1e-12:
            which_decomp = "svd"
        else:
            which_decomp = "eigen"
    kw = dict(maxdim=maxdim, mindim=mindim, cutoff=cutoff)
    if which_decomp == "svd":
        U, S, V, spec = svd(M, **kw)
        if ortho == "left":
            return U, S[:, None] * V.T, spec
        return U * S[None, :], V.T, spec
    if which_decomp == "eigen":
        if ortho == "left":
            A2 = M @ M.conj().T
            if eigen_perturbation is not None:
                A2 = A2 + eigen_perturbation
            D, U, spec = eigen(A2, **kw)
            return U, U.conj().T @ M, spec
        # ortho == "right":  rho over the right indices, A2[j,j'] = sum_i M[i,j] conj(M[i,j'])
        A2 = M.T @ M.conj()
        if eigen_perturbation is not None:
            A2 = A2 + eigen_perturbation
        D, U, spec = eigen(A2, **kw)
        return M @ U.conj(), U.T, spec
    if which_decomp == "qr":
        if ortho == "left":
            Q, R = qr(M)
            return Q, R, Spectrum(np.zeros(0), 0.0)
        Q, R = qr(M.T)
        return R.T, Q.T, Spectrum(np.zeros(0), 0.0)
    raise ValueError(which_decomp)
