"""Helpers shared by the -m gpu parity tests: all compute goes through the C ABI
(itensorsgpu.jl_b200 -> libtnb200.so); the oracle is only the checker."""
import numpy as np

from oracle import tensor as ot


def rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = (a + 1j * rng.standard_normal(shape)) / np.sqrt(2)
    return a


def dev(a):
    from itensorsgpu_b200 import tn
    return tn.DTensor.from_numpy(a)


def check_contract(rng, dims, la, lb, cplx, tol=1e-12, lc=None, alpha=None, beta=None, conj_a=False,
                   conj_b=False):
    from itensorsgpu_b200 import tn
    A = rand(rng, [dims[x] for x in la], cplx)
    B = rand(rng, [dims[x] for x in lb], cplx)
    out = None
    C0 = None
    if lc is None:
        lc = ot.output_labels(la, lb)
    if beta is not None:
        C0 = rand(rng, [dims[x] for x in lc], cplx)
        out = dev(C0)
    got, glc = tn.ops.contract(dev(A), la, dev(B), lb, lc=lc, out=out, alpha=alpha, beta=beta, conj_a=conj_a,
                               conj_b=conj_b)
    assert tuple(glc) == tuple(lc)
    want = ot.contract_into(C0, lc, A, la, B, lb, alpha=1.0 if alpha is None else alpha,
                            beta=0.0 if beta is None else beta, conj_a=conj_a, conj_b=conj_b)
    err = ot.rel_err(got.numpy(), want)
    assert err < tol, (la, lb, lc, err)
    return err
