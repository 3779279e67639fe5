#!/usr/bin/env python
"""Generates tests/golden/hotpath_small.npz: seeded INPUTS and the ORACLE's outputs for one small case of every
entry point on the hot path, plus the reference's own known-answer vectors.

The reference (Julia) cannot run in this image and holds no stored golden arrays, so the committed
fixture plays two roles:
  * tests/test_golden_cpu.py regenerates every case from oracle/ and compares it with the committed file -- a change in
    the oracle that moves a result is caught on the CPU before it can silently move the bar of the GPU parity tests;
  * tests/test_gpu_golden.py feeds the stored inputs through the C ABI on the B200 and compares with the stored outputs
    -- without executing oracle/ at all.
Reference-held numbers (test/test_cutruncate.jl:9-17, test/dmrg.jl:28,79-80) are stored verbatim under the `ref_*` keys.

usage: python tests/golden/make_golden.py [--check]      (writes the file; --check only compares)
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import dmrg as od          # noqa: E402
from oracle import models, mps, tebd   # noqa: E402
from oracle import tensor as ot        # noqa: E402
from oracle import truncate as otr     # noqa: E402

PATH = os.path.join(HERE, "hotpath_small.npz")


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = (a + 1j * rng.standard_normal(shape)) / np.sqrt(2)
    return a


def physical_bond(N, b, chi, S=0.5):
    Ws = models.heisenberg_mpo(N, S)
    d = Ws[0].shape[1]
    psi = mps.random_mps(N, d, chi, np.random.default_rng(2024))
    psi = mps.orthogonalize(psi, b)
    Rs = od.build_right_envs(psi, Ws, upto=b + 1)
    L = np.ones((1, 1, 1))
    for j in range(b):
        L = od.env_left_update(L, psi[j], Ws[j])
    return L, Ws[b], Ws[b + 1], Rs[b + 1], psi[b], psi[b + 1]


def build():
    g = {}
    # ---- a5 _contract!: rank-3 x rank-3 with permuted modes (real), H_eff step-1 layout (complex, conj + alpha/beta)
    rng = np.random.default_rng(1234)
    dims = {"i": 5, "j": 6, "k": 7, "l": 4, "a": 3}
    A = _rand(rng, [dims[x] for x in ("j", "a", "i")], False); B = _rand(rng, [dims[x] for x in ("k", "a", "l")], False)
    C, lc = ot.contract(A, ("j", "a", "i"), B, ("k", "a", "l"))
    g.update(c1_A=A, c1_B=B, c1_C=C); assert lc == ("j", "i", "k", "l")
    phi = _rand(rng, (6, 2, 2, 5), True); L = _rand(rng, (6, 4, 3), True); C0 = _rand(rng, (2, 4, 3, 2, 5), True)
    C = ot.contract_into(C0, ("s1", "lp", "a", "s2", "r"), phi, ("l", "s1", "s2", "r"), L, ("l", "lp", "a"),
                         alpha=0.5 - 0.25j, beta=2.0, conj_b=True)
    g.update(c2_A=phi, c2_B=L, c2_C0=C0, c2_C=C)
    # ---- a9/a10 permute + axpby
    X = _rand(rng, (3, 4, 5), False); Y = _rand(rng, (5, 3, 4), False)
    g.update(p_X=X, p_Y=Y, p_out=ot.axpby(-1.5, X, ("a", "b", "c"), 0.5, Y, ("c", "a", "b")))
    # ---- a20 product(PH, phi), environment updates (real and complex)
    for tag, cplx in (("r", False), ("c", True)):
        rng = np.random.default_rng(31)
        cl, cr, d, w = 16, 8, 2, 5
        L = _rand(rng, (cl, cl, w), cplx); R = _rand(rng, (cr, cr, w), cplx)
        W1 = _rand(rng, (w, d, d, w), cplx); W2 = _rand(rng, (w, d, d, w), cplx); phi = _rand(rng, (cl, d, d, cr), cplx)
        g.update({"h%s_L" % tag: L, "h%s_R" % tag: R, "h%s_W1" % tag: W1, "h%s_W2" % tag: W2, "h%s_phi" % tag: phi,
                  "h%s_out" % tag: od.heff_apply(L, W1, W2, R, phi)})
        Asite = _rand(rng, (cl, d, cr), cplx)
        g.update({"e%s_A" % tag: Asite, "e%s_Lnew" % tag: od.env_left_update(L, Asite, W1),
                  "e%s_Rnew" % tag: od.env_right_update(R, Asite, W1)})
    # ---- a21/a22 Lanczos + replacebond on a physical bond (S=1/2 chain, N=12, bond 5, chi 16)
    L, W1, W2, R, A1, A2 = physical_bond(12, 5, 16)
    g.update(b_L=L, b_W1=W1, b_W2=W2, b_R=R, b_A1=A1, b_A2=A2)
    phi0 = np.tensordot(A1, A2, axes=(2, 0))
    e, v, nmv = od.lanczos(lambda x: od.heff_apply(L, W1, W2, R, x), phi0)
    g.update(b_lanczos_energy=np.float64(e), b_lanczos_nmv=np.int64(nmv), b_lanczos_vec=v)
    for tag, ortho, noise, cutoff in (("svdL", "left", 0.0, 0.0), ("eigR", "right", 1e-8, 1e-11)):
        e, Ar, Br, spec, _ = od.bond_step(L, W1, W2, R, A1, A2, ortho, maxdim=12, cutoff=cutoff, noise=noise)
        g.update({"b_%s_energy" % tag: np.float64(e), "b_%s_theta" % tag: np.tensordot(Ar, Br, axes=(2, 0)),
                  "b_%s_truncerr" % tag: np.float64(spec.truncerr), "b_%s_keep" % tag: np.int64(Ar.shape[2])})
    # ---- a12/a13 svd / eigen spectra of a seeded matrix
    rng = np.random.default_rng(7)
    M = _rand(rng, (24, 18), True)
    g.update(s_M=M, s_S=np.linalg.svd(M, compute_uv=False))
    Hm = _rand(rng, (20, 20), True); Hm = Hm + Hm.conj().T
    g.update(s_H=Hm, s_D=np.linalg.eigvalsh(Hm)[::-1].copy())
    # ---- a24 apply(gate): two-site gate + split (complex time)
    rng = np.random.default_rng(48)
    psi = mps.random_mps(6, 2, 8, rng, dtype=np.complex128)
    psi = mps.orthogonalize(psi, 2)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=False)
    ref = [a.copy() for a in psi]
    tebd.apply_gate(ref, 2, G, 2, maxdim=6, cutoff=1e-14)
    g.update(t_G=G, t_A1=psi[2], t_A2=psi[3], t_theta=np.tensordot(ref[2], ref[3], axes=(2, 0)))
    # ---- a15 truncate!: the oracle's CPU rule on the reference's three vectors + a maxdim-bound one
    for i, (P, kw) in enumerate([([0.0], {}), ([1.0, 0.5, 0.1, 0.05], dict(cutoff=0.2, use_absolute_cutoff=True, use_relative_cutoff=False)),
                                 ([0.5, 0.4, 0.1], dict(cutoff=0.2)), ([0.4, 0.3, 0.2, 0.1], dict(maxdim=2))]):
        err, docut, nk = otr.truncate(np.array(P), **kw)
        g["tr%d" % i] = np.array([err, docut, nk], dtype=np.float64)
    # ---- reference-held numbers, verbatim
    g["ref_truncate_kat2"] = np.array([0.15, 0.3, 2.0])         # test/test_cutruncate.jl:10-13 (agrees with the CPU rule)
    g["ref_truncate_kat3_gpu_rule"] = np.array([0.1, 0.45, 1.0])   # :14-17 (the GPU rule's own answer; deliberate mismatch)
    g["ref_dmrg_spin1_n10_upper_bound"] = np.float64(-12.0)     # test/dmrg.jl:28
    g["ref_c1_energy_itensors_readme"] = np.float64(-138.940086)
    return g


def compare(old, new):
    bad = []
    for k in sorted(set(old) | set(new)):
        if k not in old or k not in new:
            bad.append((k, "missing"))
            continue
        a, b = np.asarray(old[k]), np.asarray(new[k])
        if a.shape != b.shape:
            bad.append((k, "shape %s vs %s" % (a.shape, b.shape)))
            continue
        scale = max(1.0, float(np.max(np.abs(a))) if a.size else 1.0)
        err = float(np.max(np.abs(a - b))) / scale if a.size else 0.0
        if err > 1e-12:
            bad.append((k, err))
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    g = build()
    if a.check:
        bad = compare(dict(np.load(PATH)), g)
        print("golden check:", "ok" if not bad else bad)
        sys.exit(1 if bad else 0)
    np.savez_compressed(PATH, **g)
    print("wrote %s: %d arrays, %d bytes" % (PATH, len(g), os.path.getsize(PATH)))
