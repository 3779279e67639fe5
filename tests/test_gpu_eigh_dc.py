"""GPU parity of the fast Hermitian eigensolver (tridiagonalisation + divide & conquer + back-transform)
behind tnb_eigh_trunc for n >= 128, stage by stage and end to end, against LAPACK (numpy eigh = syevd/heevd,
the CPU path of `eigen(::Hermitian)`; reference GPU call: src/tensor/culinearalgebra.jl:74-108).
Tolerances: eigenvalues eps*||A||-class (1e-13 relative to the largest), residual and orthogonality 1e-12
in the Frobenius norm (same class as test/test_cuitensor.jl:105-130)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import dev, rand

pytestmark = pytest.mark.gpu


def _h():
    from itensorsgpu_b200 import tn
    h = tn.handle()
    lib = h.lib
    for name in ("tnb_dbg_tridiag", "tnb_dbg_stedc", "tnb_dbg_backtransform"):
        getattr(lib, name).restype = C.c_int
    return h, lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _herm(rng, n, cplx):
    X = rand(rng, (n, n), cplx)
    return X + X.conj().T


@pytest.mark.parametrize("sym", ["0", "1", "100"])
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [2, 3, 17, 64, 65, 130, 257, 600])
def test_tridiag_and_backtransform(cplx, n, sym, monkeypatch):
    # sym: trailing size from which the lower-triangle-only (half-traffic) kernels are used -- 0: never,
    # 1: always, 100: switch (and rebuild of the upper triangle) part-way through the reduction
    monkeypatch.setenv("TNB_TD_SYM_MIN", sym)
    h, lib = _h()
    rng = np.random.default_rng(100 + n)
    A = _herm(rng, n, cplx)
    dA = dev(A)
    tdt = torch.complex128 if cplx else torch.float64
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    e = torch.zeros(n, dtype=torch.float64, device="cuda")
    tau = torch.zeros(n, dtype=tdt, device="cuda")
    h.check(lib.tnb_dbg_tridiag(h.h, 1 if cplx else 0, C.c_int64(n), _p(dA.data), _p(d), _p(e), _p(tau), None))
    dn, en = d.cpu().numpy(), e.cpu().numpy()[: n - 1]
    T = np.diag(dn) + np.diag(en, 1) + np.diag(en, -1)
    w = np.linalg.eigvalsh(A)
    assert np.max(np.abs(np.linalg.eigvalsh(T) - w)) < 1e-13 * max(1.0, np.max(np.abs(w)))
    # Q = H_0...H_{n-2} applied to the identity: A = Q T Q^H
    X = dev(np.eye(n, dtype=np.complex128 if cplx else np.float64))
    h.check(lib.tnb_dbg_backtransform(h.h, 1 if cplx else 0, C.c_int64(n), _p(dA.data), _p(tau), _p(X.data), C.c_int64(n), None))
    Q = X.numpy()
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(n)) < 1e-12
    assert np.linalg.norm(Q @ T @ Q.conj().T - A) < 1e-12 * np.linalg.norm(A)


def _stedc(d, e):
    h, lib = _h()
    n = len(d)
    dd = torch.from_numpy(np.array(d, dtype=np.float64)).cuda()
    ee = torch.zeros(n, dtype=torch.float64, device="cuda")
    ee[: n - 1] = torch.from_numpy(np.array(e, dtype=np.float64)).cuda()
    lam = torch.zeros(n, dtype=torch.float64, device="cuda")
    Z = torch.zeros(n * n, dtype=torch.float64, device="cuda")
    h.check(lib.tnb_dbg_stedc(h.h, C.c_int64(n), _p(dd), _p(ee), _p(lam), _p(Z), None))
    return lam.cpu().numpy(), Z.cpu().numpy().reshape((n, n), order="F")


@pytest.mark.parametrize("kind", ["random", "wilkinson", "graded", "constant", "decoupled", "tiny1e-100", "small1e-8", "huge1e100", "zero"])
@pytest.mark.parametrize("n", [1, 2, 40, 64, 65, 129, 300, 1000])
def test_stedc(kind, n):
    rng = np.random.default_rng(7 * n + len(kind))
    if kind == "random":
        d, e = rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))
    elif kind == "wilkinson":
        d, e = np.abs(np.arange(n) - n // 2).astype(float), np.ones(max(n - 1, 0))
    elif kind == "graded":
        d = 10.0 ** (-np.arange(n) / 6.0)
        e = 0.3 * np.sqrt(d[:-1] * d[1:]) if n > 1 else np.zeros(0)
    elif kind == "constant":
        d, e = np.full(n, 2.0), np.full(max(n - 1, 0), -1.0)
    elif kind in ("tiny1e-100", "small1e-8", "huge1e100"):
        # the deflation tolerance is defined for a unit-norm matrix: the solver must scale (LAPACK dstedc does)
        sc = {"tiny1e-100": 1e-100, "small1e-8": 1e-8, "huge1e100": 1e100}[kind]
        d, e = sc * rng.standard_normal(n), sc * rng.standard_normal(max(n - 1, 0))
    elif kind == "zero":
        d, e = np.zeros(n), np.zeros(max(n - 1, 0))
    else:
        d, e = rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))
        e[::7] = 0.0
    lam, Z = _stedc(d, e)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w = np.linalg.eigvalsh(T)[::-1]
    nrm = max(np.max(np.abs(w)), 1e-300) if w.size else 1.0
    assert np.max(np.abs(lam - w)) < 2e-14 * nrm * max(1, n / 100)
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) < 2e-13 * max(1, n / 100)
    assert np.linalg.norm(T @ Z - Z * lam[None, :]) < 2e-13 * nrm * max(1, n / 100)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [128, 200, 515, 1024])
def test_eigh_dc_end_to_end(cplx, n):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(5 + n)
    A = _herm(rng, n, cplx)
    D, U, err = tn.ops.eigh(dev(A))
    D, U = D.cpu().numpy(), U.numpy()
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.max(np.abs(D - w)) < 1e-13 * np.max(np.abs(w)) * max(1, n / 200)
    assert np.linalg.norm(U.conj().T @ U - np.eye(n)) < 1e-12 * max(1, n / 200)
    assert np.linalg.norm(A @ U - U * D[None, :]) < 1e-12 * np.linalg.norm(A)


@pytest.mark.parametrize("cplx", [False, True])
def test_eigh_dc_density_matrix_truncated(cplx):
    """DMRG-like rho = M M^H with a spectrum decaying over 30 decades; keep the top 100 of 400."""
    from itensorsgpu_b200 import tn
    from oracle import linalg as ol
    rng = np.random.default_rng(77)
    n = 400
    U0, _ = np.linalg.qr(rand(rng, (n, n), cplx))
    V0, _ = np.linalg.qr(rand(rng, (n, n), cplx))
    s = 10.0 ** (-np.arange(n) / 13.0)
    M = (U0 * s) @ V0.conj().T
    rho = M @ M.conj().T
    D, U, err = tn.ops.eigh(dev(rho), maxdim=100, cutoff=1e-10)
    Dr, Ur, spec = ol.eigen(rho, maxdim=100, cutoff=1e-10)
    D, U = D.cpu().numpy(), U.numpy()
    assert len(D) == len(Dr)
    assert np.max(np.abs(D - Dr)) < 1e-13 * Dr[0]
    assert err == pytest.approx(spec.truncerr, rel=1e-6, abs=n * 2.3e-16 * Dr[0])   # absolute accuracy eps*||rho|| per discarded weight
    k = len(D)
    assert np.linalg.norm(U.conj().T @ U - np.eye(k)) < 1e-12
    # the kept subspace agrees with LAPACK's: projectors equal up to the gap-limited accuracy
    P1, P2 = U @ U.conj().T, Ur @ Ur.conj().T
    assert np.linalg.norm(rho @ P1 - rho @ P2) < 1e-12 * Dr[0]


@pytest.mark.parametrize("scale", [1e-240, 1e-90, 1e-7, 1e60, 1e240])
def test_eigh_and_svd_are_scale_invariant(scale):
    """eigen / svd of s*A must be s * (eigen / svd of A): nothing on the path may assume a unit-norm input
    (1e+-240: entries whose squares leave the double range -- the drivers rescale, like LAPACK's lascl)."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(9)
    n = 300
    A = _herm(rng, n, False)
    D, U, _ = tn.ops.eigh(dev(scale * A))
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.max(np.abs(D.cpu().numpy() / scale - w)) < 1e-12 * np.max(np.abs(w))
    Un = U.numpy()
    assert np.linalg.norm(Un.T @ Un - np.eye(n)) < 1e-11
    B = rand(rng, (320, 300), False)
    Us, S, Vs, _ = tn.ops.svd(dev(np.sqrt(scale) * B))
    assert np.max(np.abs(S.cpu().numpy() / np.sqrt(scale) - np.linalg.svd(B, compute_uv=False))) < 1e-12 * np.linalg.norm(B, 2)
