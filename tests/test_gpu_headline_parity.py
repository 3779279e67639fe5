"""Oracle parity AT THE HEADLINE SIZES (BASELINE.json configs C2, C3, C4, C5): the kernel variants the benchmark
times (64x128x16 tiles, 16-byte copies, chi up to 8192) are checked against the oracle on sampled output elements --
operands are generated on the device, the oracle contracts the matching operand slices (tests/gpu_util.py), so one
check costs seconds on the CPU whatever chi is.  Semantics: the reference's GPU ~ CPU relational check
(/root/reference/test/test_cucontract.jl:170-196), bar 1e-12 (north star)."""
import numpy as np
import pytest
import torch

from gpu_util import dev_rand, pick_indices, restrict, sampled_contract_error, sampled_heff_error
from oracle import tensor as ot

pytestmark = pytest.mark.gpu
D_, W_ = 2, 5

# the four C2 shapes of SURVEY.md section 8(d)
C2 = {
    "i_rank3xrank3": (("x", "s", "r"), ("x", "a", "rp")),
    "ii_heff_step1": (("l", "s1", "s2", "r"), ("l", "lp", "a")),
    "iii_heff_step2_smallK": (("s1", "s2", "r", "lp", "a"), ("a", "s1", "s1p", "b")),
    "iv_heff_step4": (("r", "lp", "s1p", "s2p", "c"), ("r", "rp", "c")),
}


def _dims(chi):
    return {"x": chi, "s": D_, "r": chi, "a": W_, "rp": chi, "l": chi, "s1": D_, "s2": D_, "lp": chi, "s1p": D_, "b": W_,
            "s2p": D_, "c": W_}


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", sorted(C2))
@pytest.mark.parametrize("chi", [1024, 2048, 4096, 8192])
def test_c2_contraction_sweep_sampled(chi, shape, cplx):
    if cplx and chi == 8192 and shape != "i_rank3xrank3":
        chi = 6144        # ComplexF64 rank-5 operands at 8192 are 21 GB each; 6144 keeps the test within a few seconds
    la, lb = C2[shape]
    err, n = sampled_contract_error(_dims(chi), la, lb, cplx, seed=1234 + chi)
    torch.cuda.empty_cache()
    assert n >= 64 and err < 1e-12, (chi, shape, cplx, err)


@pytest.mark.parametrize("cplx,chi", [(False, 4096), (True, 2048)])
def test_c3_heff_at_benchmark_size(cplx, chi):
    err, n = sampled_heff_error(chi, chi, 2, 5, cplx, seed=2024)
    torch.cuda.empty_cache()
    assert n >= 256 and err < 1e-12, err


@pytest.mark.parametrize("chi", [1024, 2048])
def test_c5_heff_shape_w30(chi):
    """C5: dense MPO of bond dimension 30 (2D cylinder); generic small-K steps 2 and 3 (no fused instantiation)."""
    err, n = sampled_heff_error(chi, chi, 2, 30, False, seed=88)
    torch.cuda.empty_cache()
    assert n >= 256 and err < 1e-12, err


def _gate(cplx):
    from oracle import models
    return models.heisenberg_bond_gate(0.05, imaginary_time=not cplx)


@pytest.mark.parametrize("cplx", [True, False])
def test_c4_gate_and_split_chi2048(cplx):
    """C4 at maxdim 2048: theta = G (A1 A2) then the split, 4096 x 4096.  Untruncated (maxdim = 4096) the product
    A1' A2' must reproduce the oracle's theta on sampled elements, A1' must be an isometry and the singular values
    (row norms of A2') must match LAPACK's (NumPy gesdd on the host -- the reference's CPU path)."""
    from itensorsgpu_b200 import tn
    chi, d = 2048, 2
    A1 = dev_rand((chi, d, chi), cplx, 4242); A2 = dev_rand((chi, d, chi), cplx, 4243)
    tn.ops.scale(A1, 1.0 / np.sqrt(chi * d)); tn.ops.scale(A2, 1.0 / np.sqrt(chi * d))
    G = _gate(cplx)
    B1, B2, err = tn.ops.tebd_apply_gate(tn.DTensor.from_numpy(G), A1, A2, maxdim=2 * chi, cutoff=0.0)
    k = B1.dims[2]
    assert k == 2 * chi and err < 1e-20
    rng = np.random.default_rng(5)
    picks = {"l": pick_indices(rng, chi, 6), "r": pick_indices(rng, chi, 6)}
    a1 = restrict(A1, ("l", "s1", "k"), picks); a2 = restrict(A2, ("k", "s2", "r"), picks)
    theta = np.einsum("abcd,lck,kdr->labr", G, a1, a2)                     # oracle: gate on the sampled rows/columns
    b1 = restrict(B1, ("l", "s1", "k"), picks); b2 = restrict(B2, ("k", "s2", "r"), picks)
    got = np.einsum("lak,kbr->labr", b1, b2)
    assert ot.rel_err(got, theta) < 1e-12
    # left isometry: B1^H B1 = 1 on a sampled set of columns (torch matmul as an independent checker)
    M1 = B1.data.view(k, chi * d)                                            # row j = column j of the (chi d) x k matrix
    cols = torch.as_tensor(pick_indices(rng, k, 64), device="cuda")
    Gm = M1.index_select(0, cols).conj() @ M1.index_select(0, cols).T
    assert float((Gm - torch.eye(len(cols), device="cuda", dtype=Gm.dtype)).abs().max()) < 1e-12
    # singular values: row norms of B2 (k x (d chi)) vs svdvals of the full theta computed by the checker
    th = torch.einsum("abcd,kcl,rdk->rbal", torch.as_tensor(G, device="cuda").to(A1.dtype), A1.data.view(chi, d, chi),
                      A2.data.view(chi, d, chi)).reshape(d * chi, d * chi)   # row-major (r,s2',s1',l) == column-major theta
    sv = torch.from_numpy(np.linalg.svd(th.cpu().numpy(), compute_uv=False)).cuda()      # LAPACK gesdd: the CPU path
    mine = torch.linalg.vector_norm(B2.data.view(chi * d, k), dim=0)
    assert float((mine - sv).abs().max() / sv[0]) < 1e-12


def test_c4_bform_gate_chi2048_truncated():
    """The call bench.py's tebd_c4 times: B-form gate at chi = 2048, ComplexF64, truncated to maxdim 2048.  Schmidt
    values and the truncation error against LAPACK's singular values of lam_L * theta, B2' right-isometry."""
    from itensorsgpu_b200 import tn
    chi, d = 2048, 2
    g = torch.Generator(device="cuda").manual_seed(11)
    Bs = []
    for _ in range(2):
        G0 = torch.complex(torch.randn(d * chi, chi, dtype=torch.float64, device="cuda", generator=g),
                           torch.randn(d * chi, chi, dtype=torch.float64, device="cuda", generator=g))
        Q = torch.linalg.qr(G0).Q
        Bs.append(tn.DTensor(Q.contiguous().reshape(-1).clone(), (chi, d, chi)))      # right-isometry B[l,s,r]
        del G0, Q
    lam = torch.exp(-6.0 * torch.arange(chi, device="cuda", dtype=torch.float64) / chi)
    lam = lam / lam.norm()
    G = _gate(True)
    B1, B2, lam2, err = tn.ops.tebd_gate_bform(tn.DTensor.from_numpy(G), lam, Bs[0], Bs[1], maxdim=chi, cutoff=1e-12)
    k = B2.dims[0]
    assert k == chi
    tt = torch.einsum("abcd,kcl,rdk->rbal", torch.as_tensor(G, device="cuda"), Bs[0].data.view(chi, d, chi),
                      Bs[1].data.view(chi, d, chi))                          # (r, s2', s1', l)
    th = (tt * lam.view(1, 1, 1, chi)).reshape(d * chi, d * chi)
    sv = torch.from_numpy(np.linalg.svd(th.cpu().numpy(), compute_uv=False)).cuda()      # LAPACK gesdd: the CPU path
    kept = sv[:chi]
    want = kept / kept.norm()
    assert float((lam2 - want).abs().max()) < 1e-12
    assert err == pytest.approx(float((sv[chi:] ** 2).sum() / (sv ** 2).sum()), rel=1e-6, abs=1e-14)
    M2 = B2.data.view(chi * d, k)                                             # row-major view: rows = (r, s2), cols = k
    rows = torch.as_tensor(pick_indices(np.random.default_rng(3), k, 64), device="cuda")
    Gm = M2.index_select(1, rows).T @ M2.index_select(1, rows).conj()
    assert float((Gm - torch.eye(len(rows), device="cuda", dtype=Gm.dtype)).abs().max()) < 1e-12


def test_eigh_beyond_14000_and_global_merge_path():
    """the eigensolver used to reject n > 14000 (C5: chi = 8192, d = 2 gives 16384).  n = 14336 takes the
    large-merge path (deflation vectors in global memory); checked against cuSOLVER's eigvalsh (checker only)."""
    from itensorsgpu_b200 import tn
    n = 14336
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    A = (A + A.T) / np.sqrt(2.0 * n) + 2.5 * torch.eye(n, dtype=torch.float64, device="cuda")     # spectrum in (0.5, 4.5)
    Dv, U, _ = tn.ops.eigh(tn.DTensor(A.reshape(-1).clone(), (n, n)), maxdim=64)
    ref = torch.linalg.eigvalsh(A).flip(0)[:64]
    assert float((Dv - ref).abs().max()) < 1e-12 * float(ref.abs().max()) * 10
    Um = U.data.view(64, n)                                                   # rows = eigenvectors
    assert float((Um @ Um.T - torch.eye(64, device="cuda", dtype=torch.float64)).abs().max()) < 1e-12
    assert float((Um @ A - Dv.view(-1, 1) * Um).abs().max()) < 1e-12
