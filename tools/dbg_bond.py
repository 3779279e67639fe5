import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from itensorsgpu_b200 import tn
from oracle import dmrg as od, models, mps, tensor as ot, linalg as ol
from test_gpu_factor import _physical_bond
dev = tn.DTensor.from_numpy
L, W1, W2, R, A1, A2 = _physical_bond(12, 5, 16)
for ortho in ("left", "right"):
    for noise, wd in ((0.0, "eigen"), (1e-8, None), (1e-6, None)):
        e_ref, Ar, Br, spec, nmv = od.bond_step(L, W1, W2, R, A1, A2, ortho, maxdim=12, cutoff=0.0, noise=noise, which_decomp=wd)
        e, A, B, err = tn.ops.dmrg_bond_step(dev(L), dev(W1), dev(W2), dev(R), dev(A1), dev(A2), ortho, maxdim=12, cutoff=0.0, noise=noise, which_decomp=wd)
        got = np.tensordot(A.numpy(), B.numpy(), axes=(2, 0)); want = np.tensordot(Ar, Br, axes=(2, 0))
        ph = np.sign(np.vdot(want.ravel(), got.ravel()))
        print(ortho, noise, wd, "dE", e - e_ref, "state", ot.rel_err(got, ph * want), "err", err, spec.truncerr, flush=True)
# isolate: noise term + factorize on the same phi
phi = np.tensordot(A1, A2, axes=(2, 0))
e_ref, x, _ = od.lanczos(lambda v: od.heff_apply(L, W1, W2, R, v), phi)
for ortho in ("left", "right"):
    drho = 1e-6 * od.noise_term(L, W1, W2, R, x, ortho)
    rho_g = tn.ops.noise_term(dev(L), dev(W1), dev(W2), dev(R), dev(x), ortho, 1e-6)
    print("noise term", ortho, ot.rel_err(rho_g.numpy(), drho))
    M = x.reshape(x.shape[0] * 2, -1, order="F")
    Lr, Rr, spec = ol.factorize(M, ortho=ortho, maxdim=12, cutoff=0.0, eigen_perturbation=drho)
    Ag, Bg, err = tn.ops.factorize_bond(dev(x), ortho=ortho, maxdim=12, cutoff=0.0, rho_pert=dev(drho))
    k = Lr.shape[1]
    got = Ag.numpy().reshape(-1, k, order="F") @ Bg.numpy().reshape(k, -1, order="F")
    print("factorize with pert", ortho, ot.rel_err(got, Lr @ Rr), err, spec.truncerr)
