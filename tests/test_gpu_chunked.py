"""Fixed-workspace form of the DMRG tier (tnb_set_workspace_limit): above the limit H_eff*phi is cut into slabs of the
output bond, the environment updates and the noise term into chunks of a summed bond (beta = 1 accumulation) -- what
lets C5 (chi = 8192, MPO bond 30: 2 x 64 GB of temporaries unchunked) run.  Here the limit is forced down to one byte
so that moderate sizes take the chunked code paths, and every entry point is checked against the oracle."""
import numpy as np
import pytest

from gpu_util import dev, rand
from oracle import dmrg as od
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


@pytest.fixture
def tiny_limit():
    from itensorsgpu_b200 import tn
    h = tn.handle()
    h.set_workspace_limit(1)
    yield h
    h.set_workspace_limit(0)
    assert h.workspace_limit == 40 << 30


def _bond(rng, cl, cr, d, w, cplx, herm=False):
    L = rand(rng, (cl, cl, w), cplx); R = rand(rng, (cr, cr, w), cplx)
    W1 = rand(rng, (w, d, d, w), cplx); W2 = rand(rng, (w, d, d, w), cplx)
    if herm:
        L = L + np.conj(np.transpose(L, (1, 0, 2))); R = R + np.conj(np.transpose(R, (1, 0, 2)))
        W1 = W1 + np.conj(np.transpose(W1, (0, 2, 1, 3))); W2 = W2 + np.conj(np.transpose(W2, (0, 2, 1, 3)))
    return L, W1, W2, R, rand(rng, (cl, d, d, cr), cplx)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(256, 192, 2, 5), (128, 256, 3, 5), (192, 64, 2, 30)])
def test_chunked_heff_env_noise(tiny_limit, shape, cplx):
    import torch
    from itensorsgpu_b200 import tn
    cl, cr, d, w = shape
    rng = np.random.default_rng(41)
    L, W1, W2, R, phi = _bond(rng, cl, cr, d, w, cplx)
    dL, dW1, dW2, dR, dphi = dev(L), dev(W1), dev(W2), dev(R), dev(phi)
    l0 = tiny_limit.launches
    got = tn.ops.heff_apply(dL, dW1, dW2, dR, dphi).numpy()
    nl = tiny_limit.launches - l0
    assert nl >= 2 * 3, "the chunked path was not taken (%d launches)" % nl
    want = od.heff_apply(L, W1, W2, R, phi)
    assert ot.rel_err(got, want) < 1e-12
    # host-buffer entry point falls back to the un-pipelined form when chunked
    ph = torch.from_numpy(np.ascontiguousarray(phi.ravel(order="F"))).pin_memory()
    oh = torch.empty_like(ph).pin_memory()
    tn.ops.heff_apply_host(dL, dW1, dW2, dR, ph, oh, (cl, d, d, cr))
    assert ot.rel_err(oh.numpy().reshape((cl, d, d, cr), order="F"), want) < 1e-12
    # environment updates (chunks of l' / r' accumulate with beta = 1)
    A = rand(rng, (cl, d, cr), cplx); Wm = rand(rng, (w, d, d, w), cplx)
    assert ot.rel_err(tn.ops.env_update_left(dL, dev(A), dev(Wm)).numpy(), od.env_left_update(L, A, Wm)) < 1e-12
    assert ot.rel_err(tn.ops.env_update_right(dR, dev(A), dev(Wm)).numpy(), od.env_right_update(R, A, Wm)) < 1e-12
    # noise term (upper triangle is what the library computes)
    for ortho in ("left", "right"):
        rho = tn.ops.noise_term(dL, dW1, dW2, dR, dphi, ortho, 0.37).numpy()
        ref = 0.37 * od.noise_term(L, W1, W2, R, phi, ortho)
        iu = np.triu_indices(ref.shape[0])
        assert np.linalg.norm(rho[iu] - ref[iu]) / np.linalg.norm(ref[iu]) < 1e-12


@pytest.mark.parametrize("ortho", ["left", "right"])
def test_chunked_bond_step_matches_unchunked(ortho):
    """Lanczos + noise + factorize through the chunked matvec / noise term equal the unchunked call."""
    from itensorsgpu_b200 import tn
    h = tn.handle()
    rng = np.random.default_rng(43)
    cl, cm, cr, d, w = 128, 96, 192, 2, 5
    L, W1, W2, R, _ = _bond(rng, cl, cr, d, w, False, herm=True)
    A1 = rand(rng, (cl, d, cm), False); A2 = rand(rng, (cm, d, cr), False)
    kw = dict(maxdim=160, cutoff=1e-11, noise=1e-3)
    e1, a1, a2, err1 = tn.ops.dmrg_bond_step(dev(L), dev(W1), dev(W2), dev(R), dev(A1), dev(A2), ortho, **kw)
    h.set_workspace_limit(1)
    try:
        e2, b1, b2, err2 = tn.ops.dmrg_bond_step(dev(L), dev(W1), dev(W2), dev(R), dev(A1), dev(A2), ortho, **kw)
    finally:
        h.set_workspace_limit(0)
    assert a1.dims == b1.dims
    t1 = np.tensordot(a1.numpy(), a2.numpy(), axes=(2, 0)); t2 = np.tensordot(b1.numpy(), b2.numpy(), axes=(2, 0))
    assert abs(e1 - e2) < 1e-12 * abs(e1) and ot.rel_err(t2, t1) < 1e-10 and abs(err1 - err2) < 1e-14
