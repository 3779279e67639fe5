"""GPU parity: permute / + / - / scale / dot / norm / truncate vs the oracle.  Mirrors
test/test_cudense.jl:16-42, test/test_cuitensor.jl:29-54,75-79,89-99 (exact equality for
permute and add) and the three known-answer vectors of test/test_cutruncate.jl:9-17."""
import itertools

import numpy as np
import pytest

from gpu_util import dev, rand
from oracle import tensor as ot
from oracle import truncate as otr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cplx", [False, True])
def test_permute_exact_all_orders(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(21)
    dims = dict(i=2, j=3, k=40, l=5)
    la = ("i", "j", "k", "l")
    A = rand(rng, [dims[x] for x in la], cplx)
    dA = dev(A)
    for lb in itertools.permutations(la):
        got = tn.ops.permute(dA, la, lb).numpy()
        assert np.array_equal(got, ot.permute(A, la, lb)), lb       # elementwise exact (test_cuitensor.jl:29-40)


@pytest.mark.parametrize("cplx", [False, True])
def test_permute_big_transpose_and_rank14(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(22)
    A = rand(rng, (257, 131), cplx)
    assert np.array_equal(tn.ops.permute(dev(A), ("a", "b"), ("b", "a")).numpy(), A.T)   # test_cudense.jl:34-42
    A = rand(rng, (70, 3, 65), cplx)
    assert np.array_equal(tn.ops.permute(dev(A), "abc", "cba").numpy(), A.transpose(2, 1, 0))
    labs = tuple(range(14))
    A = rand(rng, (2,) * 14, cplx)
    dA = dev(A)
    for _ in range(5):                                              # test_cuitensor.jl:41-54
        lb = tuple(rng.permutation(14))
        assert np.array_equal(tn.ops.permute(dA, labs, lb).numpy(), ot.permute(A, labs, lb))


@pytest.mark.parametrize("cplx", [False, True])
def test_add_sub_with_permuted_indices(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(23)
    dims = dict(i=20, j=3, k=17)
    la, lb = ("i", "j", "k"), ("k", "i", "j")
    A = rand(rng, [dims[x] for x in la], cplx)
    B = rand(rng, [dims[x] for x in lb], cplx)
    for sgn in (1.0, -1.0):                                          # B <- B +- A  (cudense.jl:333-445)
        dB = dev(B)
        tn.ops.permute_axpby(dev(A), la, dB, lb, alpha=sgn, beta=1.0)
        assert np.array_equal(dB.numpy(), ot.axpby(sgn, A, la, 1.0, B, lb))
    dB = dev(B)
    al = (0.3 - 1.1j) if cplx else 0.3
    be = (-0.6 + 0.2j) if cplx else -0.6
    tn.ops.permute_axpby(dev(A), la, dB, lb, alpha=al, beta=be)
    assert ot.rel_err(dB.numpy(), ot.axpby(al, A, la, be, B, lb)) < 1e-15
    # same layout, odd length (vector tail path)
    x = rand(rng, (1001,), cplx)
    y = rand(rng, (1001,), cplx)
    dy = dev(y)
    tn.ops.permute_axpby(dev(x), ("n",), dy, ("n",), alpha=al, beta=be)
    assert ot.rel_err(dy.numpy(), al * x + be * y) < 1e-15


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 1000, 1001, 1 << 20, (1 << 22) + 3])
def test_dot_norm_scale(cplx, n):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(24)
    x = rand(rng, (n,), cplx)
    y = rand(rng, (n,), cplx)
    dx, dy = dev(x), dev(y)
    want = np.vdot(x, y)
    got = tn.ops.dot(dx, dy)
    assert abs(got - want) <= 1e-12 * max(1.0, np.linalg.norm(x) * np.linalg.norm(y))
    assert abs(tn.ops.norm(dx) - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
    al = (1.5 - 0.5j) if cplx else 1.5
    tn.ops.scale(dx, al)
    assert ot.rel_err(dx.numpy(), al * x) < 1e-15
    # reduction is deterministic run to run
    assert tn.ops.dot(dev(x), dy) == tn.ops.dot(dev(x), dy)


def _trunc(P, **kw):
    import torch
    from itensorsgpu_b200 import tn
    return tn.ops.truncate(torch.tensor(np.asarray(P, dtype=np.float64), device="cuda"), **kw)


def test_truncate_known_answers():
    # KAT1 (test_cutruncate.jl:9)
    assert _trunc(np.zeros(10)) == (0.0, 0.0, 1)
    # KAT2 (test_cutruncate.jl:10-13)
    err, docut, n = _trunc([1.0, 0.5, 0.1, 0.05], use_absolute_cutoff=True, cutoff=0.2)
    assert err == pytest.approx(0.15, abs=1e-15) and docut == pytest.approx(0.3) and n == 2
    # KAT3 (test_cutruncate.jl:14-17): the CPU rule is the parity target -> keeps 2
    err, docut, n = _trunc([0.5, 0.4, 0.1], cutoff=0.2)
    assert err == pytest.approx(0.1) and docut == pytest.approx(0.25) and n == 2


def test_truncate_matches_cpu_rule_randomised():
    rng = np.random.default_rng(25)
    for trial in range(200):
        n = int(rng.integers(1, 300))
        P = np.sort(rng.random(n) ** rng.integers(1, 12))[::-1].copy()
        if trial % 7 == 0:
            P[-int(rng.integers(0, n)):] = 0.0
        if trial % 11 == 0 and n > 2:
            P[-1] = -1e-18
        kw = dict(maxdim=int(rng.integers(1, n + 5)) if trial % 2 else None, mindim=int(rng.integers(1, 4)),
                  cutoff=float(10.0 ** rng.integers(-16, -1)) if trial % 3 else 0.0,
                  use_absolute_cutoff=bool(trial % 5 == 0), use_relative_cutoff=bool(trial % 13 != 0))
        want = otr.truncate(P, **kw)
        got = _trunc(P, **kw)
        assert got[2] == want[2], (trial, kw, got, want)
        assert got[0] == pytest.approx(want[0], rel=1e-12, abs=1e-300)
        assert got[1] == pytest.approx(want[1], rel=1e-12, abs=1e-300)


@pytest.mark.parametrize("cplx", [False, True])
def test_diag_contract_and_diag_storage(cplx):
    """Diag x Dense as a scale along one mode (tnb_diag_contract) instead of the reference's densify + full
    contraction (src/tensor/cudiag.jl:105-161), and the Diag-storage S / D that svd / eigen return on the CPU
    path (a12, a13, a16); checked against plain dense algebra."""
    import torch
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(71)
    A = rand(rng, (5, 7, 6), cplx)
    dvec = rng.standard_normal(7)
    # raw entry point: same order, and with an output permutation
    out = tn.ops.diag_contract(dev(A), ("i", "j", "k"), "j", torch.from_numpy(dvec).cuda(), ("i", "j", "k")).numpy()
    assert np.array_equal(out, A * dvec[None, :, None])
    out = tn.ops.diag_contract(dev(A), ("i", "j", "k"), "j", torch.from_numpy(dvec).cuda(), ("j", "k", "i")).numpy()
    assert np.array_equal(out, (A * dvec[None, :, None]).transpose(1, 2, 0))
    if cplx:
        cvec = dvec + 1j * rng.standard_normal(7)
        out = tn.ops.diag_contract(dev(A), ("i", "j", "k"), "j", torch.from_numpy(cvec).cuda(), ("k", "i", "j")).numpy()
        assert np.allclose(out, (A * cvec[None, :, None]).transpose(2, 0, 1), rtol=1e-15, atol=0)
    with pytest.raises(tn.TnbError):
        tn.ops.diag_contract(dev(A), ("i", "j", "k"), "x", torch.from_numpy(dvec).cuda(), ("i", "j", "k"))
    # ITensor level: svd returns Diag S; U*S, S*V and U*S*V go through the fast path
    i, j, k = tn.Index(6, "i"), tn.Index(5, "j"), tn.Index(4, "k")
    T = tn.randomCuITensor(i, j, k, dtype=np.complex128 if cplx else np.float64, rng=rng)
    U, S, V, spec = tn.svd(T, (i, k))
    assert S.is_diag and not U.is_diag
    u, v = S.inds
    s = np.diag(S.array())
    US = U * S
    assert US.inds == (i, k, v) and np.allclose(US.array(), U.array() * s[None, None, :], rtol=1e-14, atol=1e-15)
    SV = S * V
    assert SV.inds == (u, j) and np.allclose(SV.array(), (V.array() * s[None, :]).T, rtol=1e-14, atol=1e-15)
    a = T.array()
    rec = tn.permute(U * S * V, T.inds).array()
    assert np.linalg.norm(rec - a) < 1e-13 * np.linalg.norm(a)
    # both Diag indices shared / none shared fall back to the general contraction and stay correct
    SS = S * tn.dag(S)
    assert abs(SS.scalar() - np.sum(s * s)) < 1e-12 * np.sum(s * s)
    # eigen returns Diag D
    M = tn.randomCuITensor(i, i.prime(), dtype=np.complex128 if cplx else np.float64, rng=rng)
    Hm = M * 1.0 + tn.dag(M).replaceinds((i, i.prime()), (i.prime(), i))
    D, Ue, _ = tn.eigen(Hm, (i,), (i.prime(),))
    assert D.is_diag
