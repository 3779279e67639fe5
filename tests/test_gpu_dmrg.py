"""GPU parity at the algorithm level: DMRG energies after identical sweeps (north-star bar 1e-10),
the reference's own DMRG assertions (test/dmrg.jl:28,79-80), MPS gauge / inner tests
(test/test_cumps.jl, test/test_cumpo.jl), gate application (examples/gate_evolution.jl) and the
ITensor-level API (test/test_cuitensor.jl, test/test_cuiterativesolvers.jl)."""
import numpy as np
import pytest

from oracle import dmrg as od
from oracle import models, mps as omps, tebd as otebd
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


def _host_mps(tn, arrays):
    return tn.MPS([a.copy() for a in arrays], llim=-1, rlim=1)


def _host_mpo(tn, arrays):
    return tn.MPO([a.copy() for a in arrays])


def test_dmrg_spin_one_heisenberg_energy_parity():
    """test/dmrg.jl:5-29 -- S=1 Heisenberg N=10, 3 sweeps maxdim 10/20/40, cutoff 1e-11, noise 1e-10."""
    from itensorsgpu_b200 import tn
    N = 10
    Ws = models.heisenberg_mpo(N, 1.0)
    psi0 = omps.random_mps(N, 3, 1, np.random.default_rng(2024))
    kw = dict(maxdim=[10, 20, 40], mindim=[1, 10], cutoff=1e-11, noise=1e-10)
    e_ref, _, hist_ref = od.dmrg(Ws, psi0, od.Sweeps(3, **kw))
    hist = []
    e, psi = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(3, **kw),
                     observer=lambda sw, b, o, en, err: hist.append(en) if (o == "right" and b == 0) else None)
    assert e < -12.0                                          # the reference's assertion
    # Why not 1e-10 on this workload (measured in tests/test_gpu_dmrg_lockstep.py): with the noise term rho carries a
    # cluster of noise-lifted eigenvalues ~1e-10 whose eigenVECTORS two backward-stable eigensolvers (LAPACK syevr, the
    # GPU's divide & conquer) resolve only to an angle eps*|rho|/gap_abs ~ 1e-5.  They carry ~1e-10 of the weight of the
    # current two-site tensor (which agrees to 1e-11 bond by bond, as do all energies when the oracle is forced onto the
    # GPU's trajectory) but they span the next bond's variational space, so free-running energies differ at the 1e-6
    # level while the sweeps are unconverged and re-converge afterwards (5 sweeps: 1e-10).
    assert abs(e - e_ref) < 5e-9
    # sweeps 1-2 (unconverged): 1e-6-class; the final sweep: 5e-9
    assert np.max(np.abs(np.array(hist) - np.array(hist_ref))) < 2e-6
    # the returned state reproduces the energy
    H = tn.cu(_host_mpo(tn, Ws))
    assert abs(tn.inner(psi, psi, H) / tn.inner(psi, psi) - e) < 1e-9


def test_dmrg_svd_branch_energy_parity():
    """noise = 0, cutoff = 0 -> factorize takes the svd branch (maxdim-only truncation, config C3's rule)."""
    from itensorsgpu_b200 import tn
    N = 12
    Ws = models.heisenberg_mpo(N, 0.5)
    psi0 = omps.random_mps(N, 2, 4, np.random.default_rng(7))
    kw = dict(maxdim=[8, 16, 16], cutoff=0.0)
    e_ref, _, _ = od.dmrg(Ws, psi0, od.Sweeps(3, **kw))
    e, psi = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(3, **kw))
    assert abs(e - e_ref) < 1e-10
    e_ed = models.ed_ground_energy(Ws)
    assert e > e_ed - 1e-9 and e - e_ed < 1e-4


def test_dmrg_tfim_closed_form():
    """test/dmrg.jl:58-81 -- TFIM N=32, 5 sweeps maxdim 10/20, cutoff 1e-12, noise 1e-10."""
    from itensorsgpu_b200 import tn
    N = 32
    psi0 = tn.randomCuMPS(N, 2, seed=432)
    H = tn.cu(tn.tfim_mpo(N))
    e, psi = tn.dmrg(H, psi0, tn.Sweeps(5, maxdim=[10, 20], cutoff=1e-12, noise=1e-10))
    ex = models.tfim_exact_energy(N)
    assert abs((e - ex) / ex) < 1e-2                          # the reference's assertion
    e_ref, _, _ = od.dmrg(models.tfim_mpo(N), [t for t in psi0.cpu().tensors],
                          od.Sweeps(5, maxdim=[10, 20], cutoff=1e-12, noise=1e-10))
    assert abs(e - e_ref) < 2e-8      # maxdim 20 binds at criticality: agreement to the truncation error


@pytest.mark.parametrize("noise", [0.0, 1e-9])
def test_dmrg_strict_parity_generic_spectrum(noise):
    """The north-star bar taken literally: energies within 1e-10 after identical sweeps, every sweep, on
    both factorize branches (noise=0 -> svd, noise>0 -> eigen + perturbation).  Random site fields remove
    the SU(2) multiplets and maxdim binds with mindim = maxdim's worth of well-separated Schmidt values,
    so the kept subspace is unique and both implementations must follow the same trajectory.  (When the
    cut falls inside a degenerate multiplet or in the numerically-null space, WHICH vectors survive is
    implementation-defined -- LAPACK vs Jacobi -- and energies agree only to the truncation error; see
    test_dmrg_spin_one_heisenberg_energy_parity.)"""
    from itensorsgpu_b200 import tn
    N = 12
    Ws = models.heisenberg_mpo(N, 0.5)
    rng = np.random.default_rng(3)
    Sz = np.diag([0.5, -0.5])
    for j in range(N):
        Wj = Ws[j].copy()
        Wj[Wj.shape[0] - 1, :, :, 0] += 0.3 * rng.standard_normal() * Sz.T
        Ws[j] = Wj
    psi0 = omps.random_mps(N, 2, 4, np.random.default_rng(5))
    kw = dict(maxdim=[8, 12, 16], cutoff=0.0, noise=[noise, noise, 0.0])
    e_ref, _, hist_ref = od.dmrg(Ws, psi0, od.Sweeps(3, **kw))
    hist = []
    e, _ = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(3, **kw),
                   observer=lambda sw, b, o, en, err: hist.append(en) if (o == "right" and b == 0) else None)
    d = np.abs(np.array(hist) - np.array(hist_ref))
    # sweeps run WITH noise lift the exactly-null part of rho to a noise-dominated cluster that the
    # maxdim cut crosses, so those sweeps agree to O(noise); the bond step itself matches the oracle to
    # 1e-15 with noise on (test_dmrg_bond_step_matches_oracle).  The final, noise-free sweep is strict.
    assert np.all(d[:2] < max(1e-10, 5 * noise))
    assert d[2] < 1e-10 and abs(e - e_ref) < 1e-10


def test_dmrg_complex_dtype():
    from itensorsgpu_b200 import tn
    N = 8
    Ws = models.heisenberg_mpo(N, 0.5)
    psi0 = omps.random_mps(N, 2, 4, np.random.default_rng(9), dtype=np.complex128)
    kw = dict(maxdim=[8, 16], cutoff=1e-12, noise=[1e-9, 0.0])
    e_ref, _, _ = od.dmrg(Ws, psi0, od.Sweeps(2, **kw))
    e, _ = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(2, **kw))
    assert abs(e - e_ref) < 1e-10


def test_orthogonalize_and_inner():
    """test/test_cumps.jl:71-101,138-149,200-229 and test/test_cumpo.jl:42-91."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(12)
    N = 30
    arrs = [rng.standard_normal((1 if j == 0 else 4, 2, 1 if j == N - 1 else 4)) for j in range(N)]
    phis = [rng.standard_normal(a.shape) for a in arrs]
    psi = tn.cu(tn.MPS(arrs, llim=-1, rlim=N))
    phi = tn.cu(tn.MPS(phis, llim=-1, rlim=N))
    assert abs(tn.inner(phi, psi) - omps.inner(phis, arrs)) < 1e-10 * abs(omps.inner(phis, arrs))
    c = 14
    out = tn.orthogonalize(psi, c)
    assert out.llim == c - 1 and out.rlim == c + 1            # test_cumps.jl:140-149
    host = out.cpu().tensors
    for j in range(c):
        assert omps.left_orthogonality_error(host[j]) < 1e-12
    for j in range(c + 1, N):
        assert omps.right_orthogonality_error(host[j]) < 1e-12
    assert abs(tn.inner(out, out) - omps.inner(arrs, arrs)) < 1e-9 * abs(omps.inner(arrs, arrs))
    Ws = models.heisenberg_mpo(N, 0.5)
    H = tn.cu(tn.MPO(Ws))
    assert abs(tn.inner(psi, psi, H) - omps.expect_mpo(arrs, Ws)) < 1e-9 * abs(omps.expect_mpo(arrs, Ws))
    with pytest.raises(tn.DimensionMismatch):
        tn.inner(psi, tn.cu(tn.MPS(arrs[:-1])))


@pytest.mark.parametrize("cplx", [False, True])
def test_apply_gate_layers(cplx):
    """examples/gate_evolution.jl (one-site X gates) and a TEBD even/odd layer of two-site gates."""
    from itensorsgpu_b200 import tn
    N = 10
    psi = tn.productCuMPS(2, [0] * N)
    X = np.array([[0.0, 1.0], [1.0, 0.0]])
    out = tn.apply([(X, n) for n in range(N)], psi)
    dense = omps.to_dense(out.cpu().tensors)
    assert abs(dense[(1,) * N]) == pytest.approx(1.0)
    rng = np.random.default_rng(13)
    arrs = omps.random_mps(N, 2, 8, rng, dtype=np.complex128 if cplx else np.float64)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=not cplx)
    gates = otebd.tebd_layer_gates(N, G, 0) + otebd.tebd_layer_gates(N, G, 1)
    ref, _ = otebd.apply(gates, arrs, center=0, maxdim=12, cutoff=1e-13)
    got = tn.apply(gates, tn.cu(tn.MPS(arrs, llim=-1, rlim=1)), maxdim=12, cutoff=1e-13)
    a, b = omps.to_dense(got.cpu().tensors), omps.to_dense(ref)
    assert ot.rel_err(a, b) < 1e-9
    assert [t.shape for t in got.cpu().tensors] == [t.shape for t in ref]


@pytest.mark.parametrize("cplx", [False, True])
def test_itensor_level_api(cplx):
    """test/test_cuitensor.jl:29-40,75-79,89-112,125-130 through the mirrored ITensor interface."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(14)
    dt = np.complex128 if cplx else np.float64
    i, j, k, l = tn.Index(2, "i"), tn.Index(3, "j"), tn.Index(4, "k"), tn.Index(5, "l")
    A = tn.randomCuITensor(i, j, k, l, dtype=dt, rng=rng)
    B = tn.randomCuITensor(k, i, l, j, dtype=dt, rng=rng)
    a, b = A.array(), B.array()
    assert np.array_equal(tn.permute(A, (l, k, j, i)).array(), a.transpose(3, 2, 1, 0))
    assert tn.norm(A) == pytest.approx(np.linalg.norm(a), rel=1e-14)
    assert np.array_equal((A + B).array(), a + b.transpose(1, 3, 0, 2))
    assert np.array_equal((A - B).array(), a - b.transpose(1, 3, 0, 2))
    C = A * tn.dag(B)
    assert abs(C.scalar() - np.vdot(b.transpose(1, 3, 0, 2), a)) < 1e-12 * np.linalg.norm(a) * np.linalg.norm(b)
    U, S, V, spec = tn.svd(A, (j, l))
    u, v = tn.commonind(U, S), tn.commonind(S, V)
    assert ot.rel_err(tn.permute(U * S * V, A.inds).array(), a) < 1e-13          # A ~ U*S*V (CPU convention)
    UU = (U * tn.dag(U.prime(1, u))).array()
    VV = (V * tn.dag(V.prime(1, v))).array()
    assert np.linalg.norm(UU - np.eye(u.dim)) < 1e-12 and np.linalg.norm(VV - np.eye(v.dim)) < 1e-12
    Q, R = tn.qr(A, (i, l))
    q = tn.commonind(Q, R)
    assert ot.rel_err(tn.permute(Q * R, A.inds).array(), a) < 1e-13
    assert np.linalg.norm((Q * tn.dag(Q.prime(1, q))).array() - np.eye(q.dim)) < 1e-12
    with pytest.raises(tn.DimensionMismatch):
        A + tn.randomCuITensor(i, j, rng=rng)
    with pytest.raises(tn.TnbError):
        tn.cpu(A) * tn.cpu(B)                                  # no CPU arithmetic path


@pytest.mark.parametrize("cplx_start", [False, True])
def test_davidson_itensor_map(cplx_start):
    """test/test_cuiterativesolvers.jl:13-28."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(15)
    d = 10
    i = tn.Index(d, "i")
    A = tn.randomCuITensor(i, i.prime(), dtype=np.complex128, rng=rng)
    # A = mapprime(A*mapprime(dag(A),0,2),2,1)   (test_cuiterativesolvers.jl:17)
    A2 = (A * tn.dag(A).replaceinds((i,), (i.prime(2),))).replaceinds((i.prime(2),), (i.prime(),))
    M = lambda v: tn.noprime(A2 * v)
    v0 = tn.randomCuITensor(i, dtype=np.complex128 if cplx_start else np.float64, rng=rng)
    lam, v = tn.davidson(M, v0, maxiter=10)
    r = M(v) - v * complex(lam)
    assert tn.norm(r) < 1e-6 * abs(lam)


def _generic_heisenberg(N, seed=3):
    Ws = models.heisenberg_mpo(N, 0.5)
    rng = np.random.default_rng(seed)
    Sz = np.diag([0.5, -0.5])
    for j in range(N):
        Wj = Ws[j].copy()
        Wj[Wj.shape[0] - 1, :, :, 0] += 0.3 * rng.standard_normal() * Sz.T
        Ws[j] = Wj
    return Ws


@pytest.mark.parametrize("route", ["eigen_dc", "svd_gram", "small_forced"])
def test_dmrg_strict_parity_fast_eigensolver(route, monkeypatch):
    """Same 1e-10 bar with the factorization going through the tridiagonalisation + divide & conquer
    eigensolver (csrc/tridiag.cu, stedc.cu): (i) eigen branch at chi = 64 (rho is 128 x 128, the default
    switch-over size), (ii) svd branch routed through the Gram matrix (what chi = 4096 bonds use),
    (iii) both forced on at tiny sizes so that every bond of the sweep, edges included, takes them."""
    from itensorsgpu_b200 import tn
    if route == "small_forced":
        monkeypatch.setenv("TNB_EIGH", "dc")
        monkeypatch.setenv("TNB_SVD_GRAM_MIN", "2")
        N, kw, which = 12, dict(maxdim=[8, 12, 16], cutoff=0.0), None
    elif route == "svd_gram":
        monkeypatch.setenv("TNB_SVD_GRAM_MIN", "64")
        N, kw, which = 16, dict(maxdim=[16, 32, 64], cutoff=0.0), None
    else:
        # eigen branch: a relative cutoff of 1e-13 drops the numerically-null eigenvectors of rank-deficient
        # rho (edge bonds), whose choice is implementation-defined; the kept weight then agrees to ~1e-13
        N, kw, which = 16, dict(maxdim=[16, 32, 64], cutoff=1e-13), "eigen"
    Ws = _generic_heisenberg(N)
    psi0 = omps.random_mps(N, 2, 4, np.random.default_rng(5))
    e_ref, _, hist_ref = od.dmrg(Ws, psi0, od.Sweeps(3, **kw), which_decomp=which)
    hist = []
    e, psi = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(3, **kw), which_decomp=which,
                     observer=lambda sw, b, o, en, err: hist.append(en) if (o == "right" and b == 0) else None)
    assert np.max(np.abs(np.array(hist) - np.array(hist_ref))) < 1e-10
    assert abs(e - e_ref) < 1e-10
    for t in psi.cpu().tensors[1:]:
        assert omps.right_orthogonality_error(t) < 1e-11


def test_dmrg_config_c1_full_size():
    """BASELINE.json configs[0] at full size: S=1 Heisenberg chain, N=100, 5 sweeps maxdim 10/20/100/100/200,
    cutoff 1e-11 (examples/dmrg.jl's model with ITensors.jl's stock schedule).  Energies after identical sweeps
    within the north star's 1e-10 of the CPU path on every sweep, and the known ground-state energy of the
    N=100 open S=1 chain (-138.940086) reached."""
    from itensorsgpu_b200 import tn
    N = 100
    Ws = models.heisenberg_mpo(N, 1.0)
    psi0 = omps.random_mps(N, 3, 10, np.random.default_rng(2024))
    kw = dict(maxdim=[10, 20, 100, 100, 200], cutoff=1e-11)
    e_ref, _, hist_ref = od.dmrg(Ws, psi0, od.Sweeps(5, **kw))
    hist = []
    e, psi = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(5, **kw),
                     observer=lambda sw, b, o, en, err: hist.append(en) if (o == "right" and b == 0) else None)
    assert np.max(np.abs(np.array(hist) - np.array(hist_ref))) < 1e-10
    assert abs(e - e_ref) < 1e-10
    assert abs(e - (-138.940086)) < 2e-6


def test_dmrg_environment_offload_is_transparent():
    """env_store="host" (environment cache spilled to pinned host memory, prefetched one bond ahead) must give
    bit-identical energies to the all-in-HBM cache: it only moves data."""
    from itensorsgpu_b200 import tn
    N = 14
    Ws = _generic_heisenberg(N)
    psi0 = omps.random_mps(N, 2, 4, np.random.default_rng(8))
    kw = dict(maxdim=[8, 16, 24], cutoff=1e-12, noise=[1e-9, 0.0, 0.0])
    runs = []
    for store in ("device", "host"):
        hist = []
        e, psi = tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(3, **kw), env_store=store,
                         observer=lambda sw, b, o, en, err: hist.append(en))
        runs.append((e, hist, [t.numpy() for t in psi.tensors]))
    assert runs[0][0] == runs[1][0]
    assert runs[0][1] == runs[1][1]
    assert all(np.array_equal(a, b) for a, b in zip(runs[0][2], runs[1][2]))
    with pytest.raises(tn.TnbError):
        tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(1, maxdim=4), env_store="disk")


@pytest.mark.parametrize("cplx", [False, True])
def test_mps_algebra_add_truncate_contract(cplx):
    """[EXT] `+`, `truncate!`, `contract(::MPO, ::MPS)` on the GPU kernels vs the oracle / dense algebra, and the
    reference's consistency relation inner(phi, contract(H, psi)) == inner(phi, H, psi)
    (test/test_cumpo.jl:89-99,131; test/test_cumps.jl:196-246)."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(62)
    N = 8
    dt = np.complex128 if cplx else np.float64
    psi = omps.random_mps(N, 2, 6, rng, dtype=dt)
    phi = omps.random_mps(N, 2, 5, rng, dtype=dt)
    Ws = models.heisenberg_mpo(N, 0.5)
    gpsi, gphi, gH = tn.cu(_host_mps(tn, psi)), tn.cu(_host_mps(tn, phi)), tn.cu(_host_mpo(tn, Ws))
    dpsi, dphi = omps.to_dense(psi), omps.to_dense(phi)
    s = tn.add(gpsi, gphi)
    assert np.linalg.norm(omps.to_dense(s.cpu().tensors) - (dpsi + dphi)) < 1e-12
    t = tn.truncate(tn.add(gpsi, gpsi), cutoff=1e-14)
    assert t.maxlinkdim() <= 6 and np.linalg.norm(omps.to_dense(t.cpu().tensors) - 2 * dpsi) < 1e-11
    for tens in t.cpu().tensors[1:]:
        assert omps.right_orthogonality_error(tens) < 1e-12
    tm = tn.truncate(s, maxdim=4)
    ref = omps.truncate(omps.add(psi, phi), maxdim=4)
    assert ot.rel_err(omps.to_dense(tm.cpu().tensors), omps.to_dense(ref)) < 1e-9
    Hpsi = tn.contract(gH, gpsi)
    assert np.linalg.norm(omps.to_dense(Hpsi.cpu().tensors) - omps.to_dense(omps.contract_mpo_mps(Ws, psi))) < 1e-11
    lhs, rhs = tn.inner(gphi, Hpsi), tn.inner(gphi, gpsi, gH)
    assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(rhs))
    Ht = tn.contract(gH, gpsi, maxdim=8, cutoff=1e-13)
    assert Ht.maxlinkdim() <= 8
    assert ot.rel_err(omps.to_dense(Ht.cpu().tensors), omps.to_dense(omps.contract_mpo_mps(Ws, psi, maxdim=8, cutoff=1e-13))) < 1e-9
    with pytest.raises(tn.DimensionMismatch):
        tn.add(gpsi, tn.cu(tn.MPS(psi[:-1])))
