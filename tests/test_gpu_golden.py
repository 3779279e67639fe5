"""CUDA path against the COMMITTED golden fixture (tests/golden/hotpath_small.npz: seeded inputs + the outputs the
oracle produced when the fixture was generated, tests/golden/make_golden.py).  Nothing under oracle/ is executed here:
inputs go through the C ABI on the B200 and the results are compared with the stored arrays.  One case per entry point
of SURVEY.md section 8(a); bars: contractions 1e-12 relative (north star), energies 1e-11, factors up to the gauge."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_small.npz"))


def _dev(a):
    from itensorsgpu_b200 import tn
    return tn.DTensor.from_numpy(np.asarray(a))


def _rel(got, want):
    got, want = np.asarray(got), np.asarray(want)
    return float(np.linalg.norm((got - want).ravel()) / np.linalg.norm(want.ravel()))


def test_contract_real_permuted_modes():
    from itensorsgpu_b200 import tn
    out, lc = tn.ops.contract(_dev(G["c1_A"]), ("j", "a", "i"), _dev(G["c1_B"]), ("k", "a", "l"))
    assert tuple(lc) == ("j", "i", "k", "l")
    assert _rel(out.numpy(), G["c1_C"]) < 1e-12


def test_contract_complex_conj_alpha_beta_output_order():
    from itensorsgpu_b200 import tn
    out = _dev(G["c2_C0"])
    tn.ops.contract(_dev(G["c2_A"]), ("l", "s1", "s2", "r"), _dev(G["c2_B"]), ("l", "lp", "a"), lc=("s1", "lp", "a", "s2", "r"),
                    out=out, alpha=0.5 - 0.25j, beta=2.0, conj_b=True)
    assert _rel(out.numpy(), G["c2_C"]) < 1e-12


def test_permute_axpby():
    from itensorsgpu_b200 import tn
    Y = _dev(G["p_Y"])
    tn.ops.permute_axpby(_dev(G["p_X"]), ("a", "b", "c"), Y, ("c", "a", "b"), alpha=-1.5, beta=0.5)
    assert _rel(Y.numpy(), G["p_out"]) < 1e-15          # one fused multiply-add per element


@pytest.mark.parametrize("tag", ["r", "c"])
def test_heff_apply_and_environment_updates(tag):
    from itensorsgpu_b200 import tn
    L, R, W1, W2, phi = (_dev(G["h%s_%s" % (tag, k)]) for k in ("L", "R", "W1", "W2", "phi"))
    assert _rel(tn.ops.heff_apply(L, W1, W2, R, phi).numpy(), G["h%s_out" % tag]) < 1e-12
    A = _dev(G["e%s_A" % tag])
    assert _rel(tn.ops.env_update_left(L, A, W1).numpy(), G["e%s_Lnew" % tag]) < 1e-12
    assert _rel(tn.ops.env_update_right(R, A, W1).numpy(), G["e%s_Rnew" % tag]) < 1e-12


def _bond():
    return tuple(_dev(G["b_" + k]) for k in ("L", "W1", "W2", "R", "A1", "A2"))


def test_lanczos_energy_and_vector():
    from itensorsgpu_b200 import tn
    L, W1, W2, R, A1, A2 = _bond()
    phi, _ = tn.ops.contract(A1, ("l", "s1", "k"), A2, ("k", "s2", "r"))
    e, nmv = tn.ops.eigsolve_lanczos(L, W1, W2, R, phi)
    assert nmv == int(G["b_lanczos_nmv"])
    assert abs(e - float(G["b_lanczos_energy"])) < 1e-11
    got, want = phi.numpy(), G["b_lanczos_vec"]
    assert abs(abs(np.vdot(got.ravel(), want.ravel())) - 1.0) < 1e-10
    assert abs(np.linalg.norm(got.ravel()) - 1.0) < 1e-13


@pytest.mark.parametrize("tag,ortho,noise,cutoff", [("svdL", "left", 0.0, 0.0), ("eigR", "right", 1e-8, 1e-11)])
def test_bond_step(tag, ortho, noise, cutoff):
    from itensorsgpu_b200 import tn
    L, W1, W2, R, A1, A2 = _bond()
    e, A, B, err = tn.ops.dmrg_bond_step(L, W1, W2, R, A1, A2, ortho, maxdim=12, cutoff=cutoff, noise=noise)
    assert abs(e - float(G["b_%s_energy" % tag])) < 1e-11
    assert A.dims[2] == B.dims[0] == int(G["b_%s_keep" % tag])
    assert err == pytest.approx(float(G["b_%s_truncerr" % tag]), rel=1e-6, abs=1e-15)
    got, want = np.tensordot(A.numpy(), B.numpy(), axes=(2, 0)), G["b_%s_theta" % tag]
    s = np.sign(np.vdot(want.ravel(), got.ravel()).real)
    assert _rel(s * got, want) < 1e-8


def test_svd_and_eigen_spectra():
    from itensorsgpu_b200 import tn
    U, S, V, _ = tn.ops.svd(_dev(G["s_M"]))
    assert np.max(np.abs(S.cpu().numpy() - G["s_S"])) < 1e-12 * G["s_S"][0]
    M = (U.numpy() * S.cpu().numpy()[None, :]) @ V.numpy().T          # CPU convention: M = U diag(S) V^T
    assert _rel(M, G["s_M"]) < 1e-12
    D, Q, _ = tn.ops.eigh(_dev(G["s_H"]))
    assert np.max(np.abs(D.cpu().numpy() - G["s_D"])) < 1e-12 * np.max(np.abs(G["s_D"]))


def test_tebd_gate():
    from itensorsgpu_b200 import tn
    A1, A2, _ = tn.ops.tebd_apply_gate(_dev(G["t_G"]), _dev(G["t_A1"]), _dev(G["t_A2"]), maxdim=6, cutoff=1e-14)
    got = np.tensordot(A1.numpy(), A2.numpy(), axes=(2, 0))
    assert _rel(got, G["t_theta"]) < 1e-10


def test_truncate_rule():
    import torch
    from itensorsgpu_b200 import tn
    P = lambda v: torch.tensor(v, dtype=torch.float64, device="cuda")
    assert tn.ops.truncate(P([1.0, 0.5, 0.1, 0.05]), cutoff=0.2, use_absolute_cutoff=True, use_relative_cutoff=False) == \
        pytest.approx(tuple(G["tr1"]), abs=1e-13)
    assert tn.ops.truncate(P([0.5, 0.4, 0.1]), cutoff=0.2) == pytest.approx(tuple(G["tr2"]), abs=1e-13)
    assert tn.ops.truncate(P([0.4, 0.3, 0.2, 0.1]), maxdim=2) == pytest.approx(tuple(G["tr3"]), abs=1e-13)
