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


# ---------------------------------------------------------------------------------------------
# Parity at sizes the oracle cannot evaluate in full: operands are generated ON the device, the oracle evaluates a
# grid of sampled output elements from the matching operand slices (downloaded), and those elements are compared.
# ---------------------------------------------------------------------------------------------
def dev_rand(shape, cplx, seed):
    """flat column-major device tensor with N(0,1) entries (complex: (N + iN)/sqrt 2), as a tn.DTensor"""
    import torch
    from itensorsgpu_b200 import tn
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = int(np.prod(shape))
    if cplx:
        re = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
        im = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
        t = torch.complex(re, im) / np.sqrt(2.0)
        del re, im
    else:
        t = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    return tn.DTensor(t, shape)


def pick_indices(rng, extent, nsamp):
    if extent <= nsamp:
        return np.arange(extent)
    idx = np.sort(rng.choice(extent, nsamp, replace=False))
    idx[0], idx[-1] = 0, extent - 1
    return np.unique(idx)


def restrict(T, labels, picks):
    """Logical ndarray of the device tensor T restricted to picks[label] (index arrays) along the given labels."""
    import torch
    n = len(T.dims)
    v = T.data.view(*reversed(T.dims))                       # row-major view of the column-major buffer
    for pos, lab in enumerate(labels):
        if lab in picks:
            v = v.index_select(n - 1 - pos, torch.as_tensor(picks[lab], device=v.device))
    return v.cpu().numpy().transpose(tuple(range(n))[::-1])


def sampled_contract_error(dims, la, lb, cplx, seed, nsamp=5):
    """GPU contraction at full size vs the oracle on sampled output elements.  Returns (rel err, n elements)."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(seed)
    A = dev_rand([dims[x] for x in la], cplx, seed)
    B = dev_rand([dims[x] for x in lb], cplx, seed + 1)
    out, lc = tn.ops.contract(A, la, B, lb)
    free = [x for x in la if x not in lb] + [x for x in lb if x not in la]
    picks = {x: pick_indices(rng, dims[x], nsamp) for x in free}
    want, lw = ot.contract(restrict(A, la, picks), la, restrict(B, lb, picks), lb)
    assert tuple(lw) == tuple(lc)
    got = restrict(out, lc, picks)
    return ot.rel_err(got, want), int(want.size)


def sampled_heff_error(cl, cr, d, w, cplx, seed, nsamp=12):
    """tnb_heff_apply at full size vs oracle/dmrg.heff_apply on the environments sliced to sampled (l', r')."""
    from itensorsgpu_b200 import tn
    from oracle import dmrg as od
    rng = np.random.default_rng(seed)
    L = dev_rand((cl, cl, w), cplx, seed); R = dev_rand((cr, cr, w), cplx, seed + 1)
    W1 = dev_rand((w, d, d, w), cplx, seed + 2); W2 = dev_rand((w, d, d, w), cplx, seed + 3)
    phi = dev_rand((cl, d, d, cr), cplx, seed + 4)
    out = tn.ops.heff_apply(L, W1, W2, R, phi)
    picks = {"lp": pick_indices(rng, cl, nsamp), "rp": pick_indices(rng, cr, nsamp)}
    want = od.heff_apply(restrict(L, ("l", "lp", "a"), picks), W1.numpy(), W2.numpy(), restrict(R, ("r", "rp", "c"), picks),
                         phi.numpy())
    got = restrict(out, ("lp", "s1", "s2", "rp"), picks)
    return ot.rel_err(got, want), int(want.size)
