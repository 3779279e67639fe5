"""GPU parity: svd / eigh / qr / factorize / bond step / gate step vs the oracle (LAPACK gesdd, syevr,
geqrf).  Mirrors test/test_cuitensor.jl:105-112,125-130 (reconstruction + isometry, 1e-14-class) and
pins singular values / eigenvalues / truncation against the CPU path."""
import numpy as np
import pytest

from gpu_util import dev, rand
from oracle import dmrg as od
from oracle import linalg as ol
from oracle import models, mps, tebd
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (3, 5), (40, 40), (130, 57), (57, 130), (300, 260), (512, 512)])
def test_svd_full(cplx, shape):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(41)
    A = rand(rng, shape, cplx)
    U, S, V, err = tn.ops.svd(dev(A))
    U, S, V = U.numpy(), S.cpu().numpy(), V.numpy()
    k = min(shape)
    assert U.shape == (shape[0], k) and V.shape == (shape[1], k) and err == 0.0
    assert ot.rel_err(U @ np.diag(S) @ V.T, A) < 1e-13
    assert np.linalg.norm(U.conj().T @ U - np.eye(k)) < 1e-12
    assert np.linalg.norm(V.conj().T @ V - np.eye(k)) < 1e-12
    Sref = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(S - Sref)) < 1e-13 * Sref[0]
    assert np.all(np.diff(S) <= 1e-300)


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_truncation_matches_cpu_rule(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(42)
    X = rand(rng, (120, 90), cplx)
    U0, s0, V0 = np.linalg.svd(X, full_matrices=False)
    s = 0.7 ** np.arange(90)
    A = (U0 * s) @ V0
    for kw in (dict(maxdim=17), dict(cutoff=1e-8), dict(maxdim=40, cutoff=1e-6, mindim=3), dict(cutoff=0.0)):
        U, S, V, err = tn.ops.svd(dev(A), **kw)
        Ur, Sr, Vr, spec = ol.svd(A, **kw)
        assert len(S) == len(Sr), kw
        assert err == pytest.approx(spec.truncerr, rel=1e-9, abs=1e-18)
        assert np.max(np.abs(S.cpu().numpy() - Sr)) < 1e-13
        rec = U.numpy() @ np.diag(S.cpu().numpy()) @ V.numpy().T
        assert ot.rel_err(rec, Ur @ np.diag(Sr) @ Vr.T) < 1e-11


def test_svd_rank_deficient_and_graded():
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(43)
    B = rand(rng, (200, 20), False) @ rand(rng, (20, 150), False)        # rank 20
    U, S, V, err = tn.ops.svd(dev(B), cutoff=1e-20)
    assert len(S) <= 24
    assert ot.rel_err(U.numpy() @ np.diag(S.cpu().numpy()) @ V.numpy().T, B) < 1e-12
    # graded spectrum over 16 decades: absolute accuracy eps*sigma_max (a dense product cannot hold more)
    Q1, _ = np.linalg.qr(rand(rng, (64, 64), False)); Q2, _ = np.linalg.qr(rand(rng, (64, 64), False))
    s = 10.0 ** (-np.arange(64) / 4.0)
    A = (Q1 * s) @ Q2.T
    _, S, _, _ = tn.ops.svd(dev(A))
    assert np.max(np.abs(S.cpu().numpy() - np.linalg.svd(A, compute_uv=False))) < 1e-15


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [1, 7, 56, 113, 300])
def test_eigh_psd(cplx, n):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(44)
    X = rand(rng, (n, max(1, n // 2 + 3)), cplx)
    rho = X @ X.conj().T
    D, U, err = tn.ops.eigh(dev(rho))
    D, U = D.cpu().numpy(), U.numpy()
    w = np.linalg.eigvalsh(rho)[::-1]
    assert np.max(np.abs(D - w)) < 1e-12 * max(1.0, w[0])
    assert np.linalg.norm(U.conj().T @ U - np.eye(n)) < 1e-11
    assert np.linalg.norm(rho @ U - U * D[None, :]) < 1e-11 * max(1.0, w[0])
    D2, U2, err2 = tn.ops.eigh(dev(rho), maxdim=max(1, n // 3), cutoff=1e-10)
    Dr, Ur, spec = ol.eigen(rho, maxdim=max(1, n // 3), cutoff=1e-10)
    assert len(D2) == len(Dr)
    assert err2 == pytest.approx(spec.truncerr, rel=1e-8, abs=1e-16)


def test_eigh_indefinite_hermitian():
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(45)
    A = rand(rng, (60, 60), True)
    A = A + A.conj().T
    D, U, _ = tn.ops.eigh(dev(A))
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.max(np.abs(D.cpu().numpy() - w)) < 1e-11
    assert np.linalg.norm(A @ U.numpy() - U.numpy() * D.cpu().numpy()[None, :]) < 1e-10


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (6, 4), (4, 6), (100, 33), (33, 100), (257, 129), (512, 300)])
def test_qr(cplx, shape):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(46)
    A = rand(rng, shape, cplx)
    Q, R = tn.ops.qr(dev(A))
    Q, R = Q.numpy(), R.numpy()
    k = min(shape)
    assert ot.rel_err(Q @ R, A) < 1e-13                                   # test_cuitensor.jl:128
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(k)) < 1e-12             # test_cuitensor.jl:129
    assert np.linalg.norm(np.tril(R, -1)) == 0.0 and np.all(np.diagonal(R).real >= 0)
    Qr, Rr = ol.qr_positive(*ol.qr(A))
    assert ot.rel_err(R, Rr) < 1e-11 and ot.rel_err(Q, Qr) < 1e-11


@pytest.mark.parametrize("ortho", ["left", "right"])
@pytest.mark.parametrize("decomp", ["svd", "eigen", "qr"])
@pytest.mark.parametrize("cplx", [False, True])
def test_factorize_bond(ortho, decomp, cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(47)
    cl, d, cr = 12, 2, 20
    phi = rand(rng, (cl, d, d, cr), cplx)
    M = phi.reshape(cl * d, d * cr, order="F")
    kw = dict(maxdim=14, cutoff=1e-12) if decomp != "qr" else {}
    A, B, err = tn.ops.factorize_bond(dev(phi), ortho=ortho, which_decomp=decomp, **kw)
    Lr, Rr, spec = ol.factorize(M, ortho=ortho, which_decomp=decomp, **kw)
    A, B = A.numpy(), B.numpy()
    k = A.shape[2]
    assert k == Lr.shape[1]
    Am, Bm = A.reshape(cl * d, k, order="F"), B.reshape(k, d * cr, order="F")
    assert ot.rel_err(Am @ Bm, Lr @ Rr) < 1e-10
    if decomp != "qr":
        assert err == pytest.approx(spec.truncerr, rel=1e-8, abs=1e-16)
    if ortho == "left":
        assert np.linalg.norm(Am.conj().T @ Am - np.eye(k)) < 1e-11
    else:
        assert np.linalg.norm(Bm @ Bm.conj().T - np.eye(k)) < 1e-11


def _physical_bond(N, b, chi, S=0.5):
    Ws = models.heisenberg_mpo(N, S)
    d = Ws[0].shape[1]
    psi = mps.random_mps(N, d, chi, np.random.default_rng(2024))
    psi = mps.orthogonalize(psi, b)
    Rs = od.build_right_envs(psi, Ws, upto=b + 1)
    L = np.ones((1, 1, 1))
    for j in range(b):
        L = od.env_left_update(L, psi[j], Ws[j])
    return L, Ws[b], Ws[b + 1], Rs[b + 1], psi[b], psi[b + 1]


@pytest.mark.parametrize("ortho", ["left", "right"])
@pytest.mark.parametrize("noise,cutoff", [(0.0, 0.0), (1e-8, 1e-11)])
def test_dmrg_bond_step_matches_oracle(ortho, noise, cutoff):
    from itensorsgpu_b200 import tn
    L, W1, W2, R, A1, A2 = _physical_bond(12, 5, 16)
    e_ref, Ar, Br, spec, nmv = od.bond_step(L, W1, W2, R, A1, A2, ortho, maxdim=12, cutoff=cutoff, noise=noise)
    e, A, B, err = tn.ops.dmrg_bond_step(dev(L), dev(W1), dev(W2), dev(R), dev(A1), dev(A2), ortho, maxdim=12,
                                         cutoff=cutoff, noise=noise)
    assert abs(e - e_ref) < 1e-11 * max(1.0, abs(e_ref))
    A, B = A.numpy(), B.numpy()
    assert A.shape == Ar.shape and B.shape == Br.shape
    got = np.tensordot(A, B, axes=(2, 0))
    want = np.tensordot(Ar, Br, axes=(2, 0))
    ph = np.vdot(want.ravel(), got.ravel())          # the eigenvector's global sign is free
    assert ot.rel_err(got, want * ph / abs(ph)) < 1e-8
    assert err == pytest.approx(spec.truncerr, rel=1e-6, abs=1e-15)


@pytest.mark.parametrize("cplx", [False, True])
def test_tebd_gate(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(48)
    psi = mps.random_mps(6, 2, 8, rng, dtype=np.complex128 if cplx else np.float64)
    psi = mps.orthogonalize(psi, 2)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=not cplx)
    ref = [a.copy() for a in psi]
    tebd.apply_gate(ref, 2, G, 2, maxdim=6, cutoff=1e-14)
    A1, A2, err = tn.ops.tebd_apply_gate(dev(G), dev(psi[2]), dev(psi[3]), maxdim=6, cutoff=1e-14)
    got = np.tensordot(A1.numpy(), A2.numpy(), axes=(2, 0))
    want = np.tensordot(ref[2], ref[3], axes=(2, 0))
    assert ot.rel_err(got, want) < 1e-10
    assert mps.left_orthogonality_error(A1.numpy()) < 1e-11
