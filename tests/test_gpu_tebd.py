"""GPU parity of the B-form TEBD path (tnb_tebd_gate_bform, itensorsgpu.jl_b200/tebd.py) against the oracle's
B-form restatement (oracle/tebd.py, itself pinned to the sequential [EXT] `apply` in tests/test_oracle_pins.py)
and against `apply` directly (examples/gate_evolution.jl:46)."""
import os
import sys

import numpy as np
import pytest
import torch

from gpu_util import dev, rand
from oracle import models, mps as omps, tebd as otebd
from oracle import tensor as ot

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_host(state):
    return [t.numpy() for t in state.Bs], [l.cpu().numpy() for l in state.lams]


@pytest.mark.parametrize("cplx", [False, True])
def test_canonical_form(cplx):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(31)
    N = 8
    arrs = [rand(rng, (1 if j == 0 else 6, 2, 1 if j == N - 1 else 6), cplx) for j in range(N)]   # arbitrary gauge
    st = tn.tebd.canonical_form(tn.cu(tn.MPS(arrs)))
    Bs, lams = _to_host(st)
    dense = omps.to_dense(arrs)
    dense = dense / np.linalg.norm(dense)
    got = omps.to_dense(Bs)
    assert abs(np.vdot(got, dense)) == pytest.approx(1.0, abs=1e-12)       # same ray
    for j in range(1, N):
        assert omps.right_orthogonality_error(Bs[j]) < 1e-12
        s = np.linalg.svd(dense.reshape(2 ** j, -1), compute_uv=False)
        k = len(lams[j])
        assert np.max(np.abs(s[:k] - lams[j])) < 1e-12 and np.all(s[k:] < 1e-12)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("trunc", [dict(), dict(maxdim=6), dict(cutoff=1e-8)])
def test_gate_bform_matches_oracle(cplx, trunc):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(32)
    N = 8
    psi = omps.random_mps(N, 2, 8, rng, dtype=np.complex128 if cplx else np.float64)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=not cplx)
    Bs, lams = otebd.canonical_bform(psi)
    st = tn.tebd.BState([dev(b) for b in Bs], [torch.from_numpy(l).cuda() for l in lams])
    for n in (3, 0, 6):
        e_ref = otebd.apply_gate_bform(Bs, lams, G, n, **trunc)
        B1, B2, lam, err = tn.ops.tebd_gate_bform(dev(G), st.lams[n], st.Bs[n], st.Bs[n + 1], **trunc)
        st.Bs[n], st.Bs[n + 1], st.lams[n + 1] = B1, B2, lam
        assert B1.dims == Bs[n].shape and B2.dims == Bs[n + 1].shape
        assert np.max(np.abs(lam.cpu().numpy() - lams[n + 1])) < 1e-12
        assert err == pytest.approx(e_ref, rel=1e-6, abs=1e-15)
        # the two-site block agrees as a tensor (individual factors differ by a unitary gauge on the new bond)
        two = np.tensordot(B1.numpy(), B2.numpy(), axes=(2, 0))
        assert ot.rel_err(two, np.tensordot(Bs[n], Bs[n + 1], axes=(2, 0))) < 1e-11
        assert omps.right_orthogonality_error(B2.numpy()) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
def test_tebd_layers_match_apply(cplx):
    """even + odd layer in B form == sequential apply(gates, psi) (no truncation: same ray to 1e-11)."""
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(33)
    N = 10
    psi = omps.random_mps(N, 2, 8, rng, dtype=np.complex128 if cplx else np.float64)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=not cplx)
    gates = otebd.tebd_layer_gates(N, G, 0) + otebd.tebd_layer_gates(N, G, 1)
    ref, _ = otebd.apply(gates, psi, center=0)
    st = tn.tebd.canonical_form(tn.cu(tn.MPS(psi, llim=-1, rlim=1)))
    tn.tebd.tebd_layer(st, dev(G), 0)
    tn.tebd.tebd_layer(st, dev(G), 1)
    a, b = omps.to_dense([t.numpy() for t in st.Bs]), omps.to_dense(ref)
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    assert np.linalg.norm(a - b) < 1e-11
    got = tn.apply(gates, tn.cu(tn.MPS(psi, llim=-1, rlim=1)))
    c = omps.to_dense(got.cpu().tensors)
    assert np.linalg.norm(c / np.linalg.norm(c) - a) < 1e-11


def _worker(rank, world):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(34)
    N = max(12, 2 * world)          # dense comparison below: keep 2^N small
    psi = omps.random_mps(N, 2, 8, rng, dtype=np.complex128)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=False)
    Gd = tn.DTensor.from_numpy(G)
    full = tn.tebd.canonical_form(tn.cu(tn.MPS(psi, llim=-1, rlim=1)))
    sh = tn.tebd.ShardedTEBD.scatter_from(full, N)
    sh.warm_links()
    for step in range(2):
        sh.layer(Gd, 0, maxdim=12)
        sh.layer(Gd, 1, maxdim=12)
    out = sh.gather()
    single = tn.tebd.canonical_form(tn.cu(tn.MPS(psi, llim=-1, rlim=1)))
    for step in range(2):
        tn.tebd.tebd_layer(single, Gd, 0, maxdim=12)
        tn.tebd.tebd_layer(single, Gd, 1, maxdim=12)
    a = omps.to_dense([t.numpy() for t in out.Bs])
    b = omps.to_dense([t.numpy() for t in single.Bs])
    return float(np.linalg.norm(a - b)), [t.dims for t in out.Bs] == [t.dims for t in single.Bs]


# _worker runs at 2, 4 and 8 ranks inside tests/test_gpu_shard_dmrg.py::test_multi_rank_suite
