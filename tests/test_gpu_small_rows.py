"""SURVEY.md section 8 rows a16 (Diag / UniformDiag contractions, test/test_cudiag.jl:28-95), a19 (plan cache +
autotune, src/tensor/cudense.jl:285-326), f3 (contract(::MPO, ::MPO), add(::MPO, ::MPO): test/test_cumpo.jl:133-173)
and f4 (on-disk format, checkpoint / resume of dmrg)."""
import os

import numpy as np
import pytest

from gpu_util import dev, rand
from oracle import mps as omps
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- a16: test/test_cudiag.jl:28-95
@pytest.mark.parametrize("c1", [False, True])
@pytest.mark.parametrize("c2", [False, True])
def test_cudiag_contraction_matrix(c1, c2):
    """Every case of the reference's Diag testset, for all (T1, T2) in {Float64, ComplexF64}^2, plus the two orders
    the reference marks @test_broken (test_cudiag.jl:49,95) -- none of them densifies the diagonal."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(61)
    mi, mj = 5, 3
    i, j = tn.Index(mi, "i"), tn.Index(mj, "j")
    ip, ipp = i.prime(), i.prime(2)
    A = rand(rng, (mi, mj), c1)
    Aij = tn.cuITensor(A, (i, j))
    Dv, Ev = rand(rng, (mi,), c2), rand(rng, (mi,), c2)
    D = tn.diagITensor(Dv, i, ip)
    E = tn.diagITensor(Ev, i, ipp)
    h = tn.handle()
    # Matrix*Diag -> Matrix   (C = Aij*D has indices (j, i'))
    l0 = h.launches
    C = Aij * D
    assert C.inds == (j, ip) and not C.is_diag
    assert h.launches - l0 <= 2                                   # scale (+ permute), no k x k densified GEMM
    assert ot.rel_err(C.array(), A.T @ np.diag(Dv)) < 1e-13
    C = D * Aij
    assert C.inds == (ip, j) and ot.rel_err(C.array(), np.diag(Dv) @ A) < 1e-13
    # Diag*Diag -> Diag
    C = E * D
    assert C.is_diag and C.inds == (ipp, ip)
    assert ot.rel_err(C.array(), np.diag(Ev) @ np.diag(Dv)) < 1e-13
    # UniformDiag*Diag -> Diag, both orders
    scal = tn.diagITensor(2.0, i, ipp)
    for C in (scal * D, D * scal):
        assert C.is_diag and set(C.inds) == {ipp, ip}
        assert ot.rel_err(C.array(), 2.0 * np.diag(Dv)) < 1e-13
    # Matrix*UniformDiag -> Matrix, both orders (the second one is @test_broken in the reference)
    scal = tn.diagITensor(2.0 + (0.5j if c2 else 0.0), i, ip)
    C = scal * Aij
    assert C.inds == (ip, j) and ot.rel_err(C.array(), scal.store.value * A) < 1e-13
    C = Aij * scal
    assert C.inds == (j, ip) and ot.rel_err(C.array(), scal.store.value * A.T) < 1e-13
    # uniform * uniform, dag, and the two-shared-index fallback (full contraction to a scalar)
    C = tn.diagITensor(3.0, ip, i) * tn.diagITensor(2.0, i, ipp)
    assert C.is_uniform_diag and C.store.value == 6.0
    tr = (tn.diagITensor(Dv, i, ip) * tn.diagITensor(Ev, i, ip)).scalar()
    assert abs(tr - np.sum(Dv * Ev)) < 1e-12 * max(1.0, abs(np.sum(Dv * Ev)))


# ---------------------------------------------------------------- a19: plan cache + autotune
def test_plan_cache_and_autotune():
    from itensorsgpu_b200 import tn
    h = tn.handle()
    h.plan_cache_clear()
    rng = np.random.default_rng(62)
    A = dev(rand(rng, (96, 2, 80), False)); B = dev(rand(rng, (96, 5, 72), False))
    want, _ = ot.contract(A.numpy(), ("x", "s", "r"), B.numpy(), ("x", "a", "q"))
    s0 = h.plan_cache_stats()
    outs = [tn.ops.contract(A, ("x", "s", "r"), B, ("x", "a", "q"))[0].numpy() for _ in range(3)]
    s1 = h.plan_cache_stats()
    assert s1["entries"] == s0["entries"] + 1 and s1["misses"] == s0["misses"] + 1 and s1["hits"] == s0["hits"] + 2
    assert all(np.array_equal(o, outs[0]) for o in outs) and ot.rel_err(outs[0], want) < 1e-12
    # different strides / alignment / flags are different plans
    tn.ops.contract(A, ("x", "s", "r"), B, ("x", "a", "q"), conj_b=True)
    tn.ops.contract(A, ("x", "s", "r"), B, ("x", "a", "q"), lc=("a", "q", "s", "r"))
    assert h.plan_cache_stats()["entries"] == s1["entries"] + 2
    # a shape above the autotune threshold is timed once; the tuned plan reproduces the heuristic plan's bits
    n = 1536
    X = dev(rand(rng, (n, n), False)); Y = dev(rand(rng, (n, n), False))
    h.set_autotune(False)
    h.plan_cache_clear()
    ref = tn.ops.contract(X, ("i", "k"), Y, ("k", "j"))[0].numpy()
    assert h.plan_cache_stats()["autotuned"] == 0
    h.set_autotune(True)
    h.plan_cache_clear()
    got = tn.ops.contract(X, ("i", "k"), Y, ("k", "j"))[0].numpy()
    again = tn.ops.contract(X, ("i", "k"), Y, ("k", "j"))[0].numpy()
    st = h.plan_cache_stats()
    assert st["autotuned"] == 1 and st["hits"] == 1
    assert np.array_equal(got, ref) and np.array_equal(again, ref)          # the tile choice never changes a bit
    # beta != 0 is not idempotent: never autotuned, still correct and cached
    h.plan_cache_clear()
    C0 = rand(rng, (n, n), False)
    out = dev(C0)
    tn.ops.contract(X, ("i", "k"), Y, ("k", "j"), out=out, alpha=0.5, beta=2.0)
    assert h.plan_cache_stats()["autotuned"] == 0
    assert ot.rel_err(out.numpy(), 0.5 * ref + 2.0 * C0) < 1e-12


# ---------------------------------------------------------------- f3: MPO algebra (test/test_cumpo.jl:133-173)
@pytest.mark.parametrize("cplx", [False, True])
def test_contract_mpo_mpo_and_add(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(63)
    N, d = 6, 2
    mk = lambda w: [rand(rng, (1 if j == 0 else w, d, d, 1 if j == N - 1 else w), cplx) for j in range(N)]
    Ks, Ls = mk(3), mk(2)
    K, L = tn.cu(tn.MPO(Ks)), tn.cu(tn.MPO(Ls))
    # untruncated: the dense operator of contract(K, L) is the matrix product (L first, then K) of the dense operators
    KL = tn.contract(K, L)
    got = omps.mpo_to_dense([t.numpy() for t in KL.tensors])
    dK, dL = omps.mpo_to_dense(Ks), omps.mpo_to_dense(Ls)
    assert ot.rel_err(got, dL @ dK) < 1e-12                       # O[(s),(s')] with s = ket: apply L (s->s'), then K (s'->s'')
    assert ot.rel_err(got, omps.mpo_to_dense(omps.contract_mpo_mpo(Ks, Ls))) < 1e-12
    # the reference's consistency relation: <psi| KL |psi> == <psi| K (L psi)>   (test_cumpo.jl:153-155)
    psi = tn.randomCuMPS(N, d, chi=4, seed=7, dtype=np.complex128 if cplx else np.float64)
    lhs = tn.inner(psi, psi, KL)
    rhs = tn.inner(psi, tn.contract(K, tn.contract(L, psi)))
    assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(lhs))
    # truncated to the exact rank (6 = 3*2): nothing is lost, bond dimensions shrink to <= 6
    KLt = tn.contract(K, L, maxdim=6, cutoff=0.0)
    assert max(t.dims[3] for t in KLt.tensors) <= 6
    assert ot.rel_err(omps.mpo_to_dense([t.numpy() for t in KLt.tensors]), dL @ dK) < 1e-11
    # reference case: link dimension 1 on both sides, maxdim = 1 (test_cumpo.jl:146-155)
    K1, L1 = tn.randomCuMPO(N, d, seed=1), tn.randomCuMPO(N, d, seed=2)
    assert max(t.dims[3] for t in K1.tensors) == 1
    KL1 = tn.contract(K1, L1, maxdim=1)
    p1 = tn.randomCuMPS(N, d, seed=3)
    a = tn.inner(p1, p1, KL1)
    b = tn.inner(p1, tn.contract(K1, tn.contract(L1, p1, maxdim=1), maxdim=1))
    assert abs(a - b) < 1e-10 * max(1.0, abs(a))
    # add(K, L) (test_cumpo.jl:133-143)
    M = tn.add(K, L)
    assert len(M) == N
    assert ot.rel_err(omps.mpo_to_dense([t.numpy() for t in M.tensors]), dK + dL) < 1e-12
    assert ot.rel_err(omps.mpo_to_dense(omps.add_mpo(Ks, Ls)), dK + dL) < 1e-12
    # DimensionMismatch on chains of different length (test_cumpo.jl:170-172)
    with pytest.raises(tn.DimensionMismatch):
        tn.contract(K, tn.randomCuMPO(N + 1, d, seed=4))


def test_inner_with_complex_mpo_and_real_mps():
    """inner(phi, psi, H) used to cast a complex MPO to the real dtype of the MPS (imaginary part lost)."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(64)
    N, d = 5, 2
    Ws = [rand(rng, (1 if j == 0 else 3, d, d, 1 if j == N - 1 else 3), True) for j in range(N)]
    psi = omps.random_mps(N, d, 4, rng)
    got = tn.inner(tn.cu(tn.MPS(psi)), tn.cu(tn.MPS(psi)), tn.cu(tn.MPO(Ws)))
    want = omps.expect_mpo(psi, Ws)
    assert isinstance(got, complex) and abs(got - want) < 1e-12 * abs(want)
    with pytest.raises(TypeError):
        dev(Ws[0]).astype(__import__("torch").float64)


# ---------------------------------------------------------------- f4: on-disk format, checkpoint / resume
def test_on_disk_round_trip_and_dmrg_resume(tmp_path):
    from itensorsgpu_b200 import tn
    N = 10
    H = tn.cu(tn.heisenberg_mpo(N, 0.5))
    psi0 = tn.randomCuMPS(N, 2, chi=8, seed=11)
    # ITensor with index metadata, Dense and Diag storage
    i, j = tn.Index(4, "i"), tn.Index(3, "j", plev=1)
    T = tn.randomCuITensor(i, j, dtype=np.complex128, rng=np.random.default_rng(1))
    p = str(tmp_path / "t.npz")
    tn.save_itensor(p, T, extra={"note": "x"})
    T2, ex = tn.load_itensor(p)
    assert ex == {"note": "x"} and T2.inds == T.inds and T2.inds[1].plev == 1 and np.array_equal(T2.array(), T.array())
    Dg = tn.diagITensor(np.arange(4.0), i, i.prime())
    tn.save_itensor(p, Dg)
    D2, _ = tn.load_itensor(p)
    assert D2.is_diag and np.array_equal(D2.array(), Dg.array())
    # MPO round trip
    pm = str(tmp_path / "H.npz")
    tn.save_chain(pm, H)
    H2, _ = tn.load_chain(pm)
    assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(H.tensors, H2.tensors))
    # 4 sweeps in one go == 2 sweeps, checkpoint, reload, 2 more sweeps (bit-identical energies)
    kw = dict(maxdim=[8, 16, 16, 16], cutoff=1e-12, noise=[1e-8, 1e-9, 0.0, 0.0])
    e_full, _ = tn.dmrg(H, psi0, tn.Sweeps(4, **kw))
    ck = str(tmp_path / "psi.npz")
    e_half, _ = tn.dmrg(H, psi0, tn.Sweeps(2, maxdim=kw["maxdim"][:2], cutoff=1e-12, noise=kw["noise"][:2]), checkpoint=ck)
    assert os.path.exists(ck) and not os.path.exists(ck + ".tmp")
    psi_ck, extra = tn.load_chain(ck)
    assert extra["sweeps_done"] == 2 and extra["energy"] == e_half and psi_ck.llim == -1 and psi_ck.rlim == 1
    e_res, _ = tn.dmrg(H, psi_ck, tn.Sweeps(2, maxdim=kw["maxdim"][2:], cutoff=1e-12, noise=kw["noise"][2:]))
    assert e_res == e_full


# ---------------------------------------------------------------- native sweep driver (tnb_dmrg_sweep)
@pytest.mark.parametrize("cplx", [False, True])
def test_native_sweep_driver_matches_python_loop(cplx):
    """The C++ sweep loop issues the same bond steps and environment updates as the Python loop: every bond energy, the
    truncation errors, the bond dimensions and the final state are bit-identical (svd rule, eigen + noise rule, growing
    maxdim, mindim; real and complex)."""
    from itensorsgpu_b200 import tn
    N = 10
    H = tn.cu(tn.heisenberg_mpo(N, 0.5))
    psi0 = tn.randomCuMPS(N, 2, chi=4, seed=21, dtype=np.complex128 if cplx else np.float64)
    for kw in (dict(maxdim=[4, 8, 16], cutoff=0.0), dict(maxdim=[6, 12, 20], mindim=[1, 4], cutoff=1e-10, noise=[1e-6, 1e-8, 0.0])):
        a, b = [], []
        e1, p1 = tn.dmrg(H, psi0, tn.Sweeps(3, **kw), observer=lambda s, bb, o, e, err: a.append((s, bb, o, e, err)))
        e2, p2 = tn.dmrg(H, psi0, tn.Sweeps(3, **kw), driver="native", observer=lambda s, bb, o, e, err: b.append((s, bb, o, e, err)))
        assert a == b and e1 == e2
        assert [t.dims for t in p1.tensors] == [t.dims for t in p2.tensors]
        assert all(np.array_equal(x.numpy(), y.numpy()) for x, y in zip(p1.tensors, p2.tensors))
    with pytest.raises(tn.TnbError):
        tn.dmrg(H, psi0, tn.Sweeps(1, maxdim=4), driver="native", env_store="host")
