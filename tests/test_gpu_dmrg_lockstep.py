"""Where do the GPU and the oracle DMRG trajectories separate on the reference's OWN workloads (test/dmrg.jl:5-29 S=1
Heisenberg N=10; test/dmrg.jl:58-81 TFIM N=32), and why?  (VERDICT r1, weak #2 / next #10.)

Lockstep experiment: the GPU runs its sweep; at EVERY bond the oracle performs the same bond step from the GPU's
current state (environments and site tensors downloaded), i.e. the oracle is forced onto the GPU's trajectory.  Result
(measured, gpurun_out/lockstep_*.json; asserted below):

  * the Lanczos energy of every bond step agrees to 1e-12 (measured 2e-15), the kept dimension is always the same, the
    truncation errors agree to 1e-12 and the truncated two-site tensors theta = A1'*A2' agree to 1e-9 (measured 7e-12
    on S=1, 3e-10 on TFIM) -- on every bond of every sweep.  There is NO degenerate cut on these workloads (smallest
    relative gap at a cut: 4e-2), so round 1's "SU(2) multiplet" explanation does not apply here;
  * what differs is the KEPT SUBSPACE beyond theta: with the noise term (1e-10) rho carries a cluster of noise-lifted
    eigenvalues ~1e-10..1e-9 on top of |rho| = 1.  Both eigensolvers (LAPACK syevr, the GPU's tridiagonalisation + divide
    & conquer) are backward stable to eps*|rho| ~ 1e-16, which resolves the eigenVECTORS inside that cluster only to an
    angle ~ eps*|rho| / gap_abs ~ 1e-16 / 1e-11 ~ 1e-5.  Those vectors carry ~1e-10 of the weight of theta (hence theta
    agrees), but they are basis vectors of the NEXT bond's variational space, so the next energies of two free-running
    trajectories differ at the 1e-6 level while the sweeps are unconverged -- from the third bond of the first sweep on --
    and re-converge as the state converges (5 sweeps, noise off at the end: 1e-10).  The projector difference measured
    in lockstep obeys the Davis-Kahan bound c*eps*|rho|/gap_abs on every bond.

So the 2e-6-per-sweep difference of the free-running trajectories (tests/test_gpu_dmrg.py) is neither a truncation bug
nor an arbitrary choice inside a degenerate multiplet: it is the conditioning of the noise-lifted eigenvectors, common
to any two backward-stable eigensolvers; with the kept subspace forced equal, parity holds far below the 1e-10 bar."""
import numpy as np
import pytest

from gpu_util import dev
from oracle import dmrg as od
from oracle import models, mps as omps
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


def _full_spectrum(phi, ortho, drho):
    l, d1, d2, r = phi.shape
    M = phi.reshape(l * d1, d2 * r, order="F")
    rho = M @ M.conj().T if ortho == "left" else M.T @ M.conj()
    if drho is not None:
        rho = rho + drho
    return np.sort(np.linalg.eigvalsh(rho))[::-1]


def _lockstep(tn, Ws, psi0, sweeps):
    N = len(psi0)
    dW = [dev(W) for W in Ws]
    ts = [dev(A) for A in psi0]
    one = dev(np.ones((1, 1, 1)))
    Rs = [None] * N
    Rs[N - 1] = one
    for j in range(N - 1, 1, -1):
        Rs[j - 1] = tn.ops.env_update_right(Rs[j], ts[j], dW[j])
    Ls = [None] * N
    Ls[0] = one
    rec = []
    for sw in range(sweeps.nsweep):
        kw = dict(maxdim=sweeps.maxdim[sw], mindim=sweeps.mindim[sw], cutoff=sweeps.cutoff[sw], noise=sweeps.noise[sw])
        order = [(b, "left") for b in range(N - 1)] + [(b, "right") for b in range(N - 2, -1, -1)]
        for b, ortho in order:
            L, R, A1, A2 = Ls[b].numpy(), Rs[b + 1].numpy(), ts[b].numpy(), ts[b + 1].numpy()
            # --- oracle, from the GPU's state (pieces of od.bond_step, so that phi and rho's spectrum are visible)
            mv = lambda v: od.heff_apply(L, Ws[b], Ws[b + 1], R, v)
            e_o, phi_o, _ = od.lanczos(mv, np.tensordot(A1, A2, axes=(2, 0)), krylovdim=3, maxiter=1)
            drho = kw["noise"] * od.noise_term(L, Ws[b], Ws[b + 1], R, phi_o, ortho) if kw["noise"] > 0 else None
            Ao, Bo, spec = od.replacebond(phi_o, ortho, kw["maxdim"], kw["mindim"], kw["cutoff"], drho=drho)
            w = _full_spectrum(phi_o, ortho, drho)
            # --- GPU
            e_g, G1, G2, err_g = tn.ops.dmrg_bond_step(Ls[b], dW[b], dW[b + 1], Rs[b + 1], ts[b], ts[b + 1], ortho, **kw)
            ts[b], ts[b + 1] = G1, G2
            if ortho == "left":
                Ls[b + 1] = tn.ops.env_update_left(Ls[b], ts[b], dW[b])
            else:
                Rs[b] = tn.ops.env_update_right(Rs[b + 1], ts[b + 1], dW[b + 1])
            ko, kg = Ao.shape[2], G1.dims[2]
            g1, g2 = G1.numpy(), G2.numpy()
            th_o = np.tensordot(Ao, Bo, axes=(2, 0)); th_g = np.tensordot(g1, g2, axes=(2, 0))
            # projector onto the kept basis (the isometry side): what the NEXT bond's variational space is built from
            if ortho == "left":
                Uo, Ug = Ao.reshape(-1, ko, order="F"), g1.reshape(-1, kg, order="F")
            else:
                Uo, Ug = Bo.reshape(ko, -1, order="F").T, g2.reshape(kg, -1, order="F").T
            dP = float(np.linalg.norm(Uo @ Uo.conj().T - Ug @ Ug.conj().T, 2)) if ko == kg else 1.0
            dth = min(np.linalg.norm(th_g - th_o), np.linalg.norm(th_g + th_o))           # overall sign of phi is free
            k0, k1 = min(ko, kg), max(ko, kg)
            tot = w.sum()
            if k0 < len(w) and k0 > 0:
                gap = (w[k0 - 1] - w[k0]) / w[k0 - 1] if w[k0 - 1] > 0 else 0.0
            else:
                gap = 1.0                                                                   # nothing discarded
            ambiguous = (ko != kg) or gap < 1e-6
            # weight of the cluster straddling the cut: everything within 1e-6 (relative) of its two edge values,
            # plus whatever lies between the two cuts when the kept dimensions differ
            WE = 0.0
            if ambiguous and k0 > 0 and k0 < len(w):
                hi, lo = w[k0 - 1], w[min(k1, len(w) - 1)]
                sel = (w <= hi * (1 + 1e-6)) & (w >= lo * (1 - 1e-6))
                WE = float(w[sel].sum() / tot)
            gap_abs = float(w[k0 - 1] - w[k0]) if (0 < k0 < len(w)) else float(w[0])
            rec.append(dict(sweep=sw, bond=b, ortho=ortho, dP=dP, gap_abs=gap_abs, rho_norm=float(w[0]), noise=kw["noise"], de=abs(e_g - e_o) / max(1.0, abs(e_o)), dth=float(dth), ko=ko, kg=kg,
                            gap=float(gap), ambiguous=bool(ambiguous), WE=WE, err_o=float(spec.truncerr), err_g=float(err_g)))
    return rec


def _dump(name, rec):
    import json
    import os
    d = os.environ.get("TNB_TEST_TRACE")
    if d:
        json.dump(rec, open(os.path.join(d, "lockstep_%s.json" % name), "w"))


def _check(rec):
    assert max(r["de"] for r in rec) < 1e-12                                     # every Lanczos step: same energy
    clean = [r for r in rec if not r["ambiguous"]]
    amb = [r for r in rec if r["ambiguous"]]
    assert len(clean) > len(rec) // 3                                            # the experiment is not vacuous
    assert max(r["dth"] for r in clean) < 1e-9, max(clean, key=lambda r: r["dth"])
    assert all(abs(r["err_o"] - r["err_g"]) < 1e-12 for r in clean)
    for r in amb:                                                                # bounded by the straddling cluster
        assert r["dth"] <= 3.0 * np.sqrt(r["WE"]) + 1e-9, r
        assert abs(r["err_o"] - r["err_g"]) <= 2.0 * r["WE"] + 1e-12, r      # same truncation error up to the cluster
    # kept subspace: Davis-Kahan, sin(angle) <= |backward error| / gap with backward error ~ c * eps * |rho|
    eps = np.finfo(float).eps
    for r in clean:
        assert r["dP"] <= 2e3 * eps * r["rho_norm"] / max(r["gap_abs"], 1e-300) + 1e-12, r
    return clean, amb


def test_lockstep_spin_one_heisenberg_reference_workload():
    """test/dmrg.jl:5-29: S=1 Heisenberg N=10, maxdim 10/20/40, mindim 1/10, cutoff 1e-11, noise 1e-10."""
    from itensorsgpu_b200 import tn
    N = 10
    Ws = models.heisenberg_mpo(N, 1.0)
    psi0 = omps.random_mps(N, 3, 1, np.random.default_rng(2024))
    rec = _lockstep(tn, Ws, psi0, od.Sweeps(3, maxdim=[10, 20, 40], mindim=[1, 10], cutoff=1e-11, noise=1e-10))
    _dump("spin1", rec)
    clean, amb = _check(rec)
    assert not amb                                               # no degenerate cut on this workload
    # ... yet the kept subspaces differ far above rounding where the cut runs through the noise-lifted cluster
    assert max(r["dP"] for r in rec if r["sweep"] == 0) > 1e-9


def test_lockstep_tfim_reference_workload():
    """test/dmrg.jl:58-81: TFIM N=32, maxdim 10/20, cutoff 1e-12, noise 1e-10 (3 of the 5 sweeps)."""
    from itensorsgpu_b200 import tn
    N = 32
    Ws = models.tfim_mpo(N)
    psi0 = omps.random_mps(N, 2, 1, np.random.default_rng(432))
    rec = _lockstep(tn, Ws, psi0, od.Sweeps(3, maxdim=[10, 20], cutoff=1e-12, noise=1e-10))
    _dump("tfim", rec)
    _check(rec)


def _free_running(tn, Ws, psi0, sweeps_kw, nsweep):
    from test_gpu_dmrg import _host_mpo, _host_mps
    eo, eg = [], []
    od.dmrg(Ws, psi0, od.Sweeps(nsweep, **sweeps_kw), observer=lambda sw, b, o, en, spec: eo.append(en))
    tn.dmrg(tn.cu(_host_mpo(tn, Ws)), tn.cu(_host_mps(tn, psi0)), tn.Sweeps(nsweep, **sweeps_kw),
            observer=lambda sw, b, o, en, err: eg.append(en))
    return np.abs(np.array(eo) - np.array(eg))


def test_free_running_separation_and_reconvergence():
    """The two free-running trajectories of the S=1 workload (5 sweeps, noise switched off for the last two): identical
    until the first noise-lifted basis vectors enter the next bond's variational space (third bond of sweep 1), 1e-6-class
    while unconverged, and back under the north-star bar once converged."""
    from itensorsgpu_b200 import tn
    N = 10
    Ws = models.heisenberg_mpo(N, 1.0)
    psi0 = omps.random_mps(N, 3, 1, np.random.default_rng(2024))
    d = _free_running(tn, Ws, psi0, dict(maxdim=[10, 20, 40, 40, 40], mindim=[1, 10], cutoff=1e-11,
                                         noise=[1e-10, 1e-10, 1e-10, 0.0, 0.0]), 5)
    _dump("spin1_free", [float(x) for x in d])
    per = 2 * (N - 1)
    assert d[:2].max() < 1e-12                                   # identical start (nothing noise-lifted is in the basis yet)
    sweep_max = [d[i * per:(i + 1) * per].max() for i in range(5)]
    assert max(sweep_max) < 1e-5, sweep_max
    assert d[-1] < 5e-10, d[-6:]                                 # converged: final energies agree at the north-star bar
