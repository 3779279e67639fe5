"""GPU parity: fused DMRG entry points (H_eff*phi, environment updates, Lanczos, noise term)
against the oracle's restatement of [EXT] ProjMPO / KrylovKit (oracle/dmrg.py)."""
import numpy as np
import pytest

from gpu_util import dev, rand
from oracle import dmrg as od
from oracle import models, mps
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


def _random_bond(rng, cl, cr, d, w, cplx):
    L = rand(rng, (cl, cl, w), cplx)
    R = rand(rng, (cr, cr, w), cplx)
    W1 = rand(rng, (w, d, d, w), cplx)
    W2 = rand(rng, (w, d, d, w), cplx)
    phi = rand(rng, (cl, d, d, cr), cplx)
    return L, W1, W2, R, phi


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 4, 2, 1), (16, 8, 2, 5), (64, 96, 2, 5), (33, 47, 3, 5), (128, 128, 2, 3)])
def test_heff_apply(cplx, shape):
    from itensorsgpu_b200 import tn
    cl, cr, d, w = shape
    rng = np.random.default_rng(31)
    L, W1, W2, R, phi = _random_bond(rng, cl, cr, d, w, cplx)
    got = tn.ops.heff_apply(dev(L), dev(W1), dev(W2), dev(R), dev(phi)).numpy()
    assert ot.rel_err(got, od.heff_apply(L, W1, W2, R, phi)) < 1e-12


@pytest.mark.parametrize("shape", [(48, 40, 2, 5, False), (40, 1030, 2, 5, False), (24, 1024, 3, 5, True)])
def test_heff_apply_host_buffers(shape):
    """chiR >= 1024 takes the pipelined form: phi uploaded in chunks over r that overlap step 1 (strided T1
    windows), H*phi downloaded in chunks over r' that overlap step 4 (strided windows of R); 1030 is ragged."""
    import torch
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(32)
    cl, cr, d, w, cplx = shape
    L, W1, W2, R, phi = _random_bond(rng, cl, cr, d, w, cplx)
    ph = torch.from_numpy(np.ascontiguousarray(phi.ravel(order="F"))).pin_memory()
    out = torch.empty_like(ph).pin_memory()
    for _ in range(2):                   # twice: the second call reuses the events and the copy stream
        out.zero_()
        tn.ops.heff_apply_host(dev(L), dev(W1), dev(W2), dev(R), ph, out, (cl, d, d, cr))
        got = out.numpy().reshape((cl, d, d, cr), order="F")
        assert ot.rel_err(got, od.heff_apply(L, W1, W2, R, phi)) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
def test_env_updates(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(33)
    cl, cr, d, wl, wr = 40, 56, 2, 5, 3
    A = rand(rng, (cl, d, cr), cplx)
    W = rand(rng, (wl, d, d, wr), cplx)
    L = rand(rng, (cl, cl, wl), cplx)
    R = rand(rng, (cr, cr, wr), cplx)
    got = tn.ops.env_update_left(dev(L), dev(A), dev(W)).numpy()
    assert ot.rel_err(got, od.env_left_update(L, A, W)) < 1e-12
    got = tn.ops.env_update_right(dev(R), dev(A), dev(W)).numpy()
    assert ot.rel_err(got, od.env_right_update(R, A, W)) < 1e-12


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


@pytest.mark.parametrize("S", [0.5, 1.0])
def test_lanczos_matches_oracle(S):
    from itensorsgpu_b200 import tn
    L, W1, W2, R, A1, A2 = _physical_bond(12, 5, 24, S)
    phi = np.tensordot(A1, A2, axes=(2, 0))
    e_ref, x_ref, nmv_ref = od.lanczos(lambda v: od.heff_apply(L, W1, W2, R, v), phi, krylovdim=3, maxiter=1)
    dphi = dev(phi)
    e, nmv = tn.ops.eigsolve_lanczos(dev(L), dev(W1), dev(W2), dev(R), dphi, krylovdim=3, maxiter=1)
    assert nmv == nmv_ref == 3
    assert abs(e - e_ref) < 1e-11 * max(1.0, abs(e_ref))
    x = dphi.numpy()
    assert abs(abs(np.vdot(x.ravel(), x_ref.ravel())) - 1.0) < 1e-10
    assert abs(np.linalg.norm(x.ravel()) - 1.0) < 1e-13


def test_lanczos_tiny_krylov_space():
    """Edge bond whose vector space (dim 2) is smaller than krylovdim: beta hits zero."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(34)
    L = np.ones((1, 1, 1)); R = np.ones((1, 1, 1))
    W1 = rand(rng, (1, 2, 2, 1), False); W1 = W1 + W1.transpose(0, 2, 1, 3)
    W2 = np.zeros((1, 1, 1, 1)); W2[0, 0, 0, 0] = 1.0
    phi = rand(rng, (1, 2, 1, 1), False)
    e_ref, x_ref, _ = od.lanczos(lambda v: od.heff_apply(L, W1, W2, R, v), phi, krylovdim=3)
    dphi = dev(phi)
    e, _ = tn.ops.eigsolve_lanczos(dev(L), dev(W1), dev(W2), dev(R), dphi, krylovdim=3)
    assert abs(e - e_ref) < 1e-12
    assert abs(abs(np.vdot(dphi.numpy().ravel(), x_ref.ravel())) - 1.0) < 1e-12


@pytest.mark.parametrize("ortho", ["left", "right"])
@pytest.mark.parametrize("cplx", [False, True])
def test_noise_term(ortho, cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(35)
    L, W1, W2, R, phi = _random_bond(rng, 24, 20, 2, 5, cplx)
    got = tn.ops.noise_term(dev(L), dev(W1), dev(W2), dev(R), dev(phi), ortho, 1e-3).numpy()
    want = 1e-3 * od.noise_term(L, W1, W2, R, phi, ortho)
    # the perturbation is Hermitian and feeds the 'U'-convention eigensolver: only its upper triangle is computed
    n = int(round(np.sqrt(want.size)))
    g, w = np.triu(got.reshape(n, n, order="F")), np.triu(want.reshape(n, n, order="F"))
    assert ot.rel_err(g, w) < 1e-12

    # a larger case spanning several 64 x 128 tiles, through the eigen branch that consumes it
    L, W1, W2, R, phi = _random_bond(rng, 96, 80, 2, 5, cplx)
    got = tn.ops.noise_term(dev(L), dev(W1), dev(W2), dev(R), dev(phi), ortho, 1e-3).numpy()
    want = 1e-3 * od.noise_term(L, W1, W2, R, phi, ortho)
    n = int(round(np.sqrt(want.size)))
    assert ot.rel_err(np.triu(got.reshape(n, n, order="F")), np.triu(want.reshape(n, n, order="F"))) < 1e-12
