"""GPU parity of the output-bond-sharded H_eff*phi slab kernel path (single GPU computes every slab in
turn; the all-gather plumbing is covered by tests/test_shard_gloo.py and by bench.py --gpus N)."""
import numpy as np
import pytest

from gpu_util import dev, rand
from oracle import dmrg as od
from oracle import tensor as ot

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("cplx", [False, True])
def test_heff_apply_shard_slabs(world, cplx):
    import torch
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(51)
    chi, cr, d, w = 64, 48, 2, 5
    L = rand(rng, (chi, chi, w), cplx); R = rand(rng, (cr, cr, w), cplx)
    W1 = rand(rng, (w, d, d, w), cplx); W2 = rand(rng, (w, d, d, w), cplx)
    phi = rand(rng, (chi, d, d, cr), cplx)
    want = od.heff_apply(L, W1, W2, R, phi)
    dL = dev(L)
    slabs = []
    for rank in range(world):
        Ls = tn.shard.left_env_slab(dL.data, chi, w, rank, world)
        lo, hi = tn.shard.slab_range(chi, rank, world)
        out = tn.ops.heff_apply_shard(tn.DTensor(Ls, (chi, hi - lo, w)), dev(W1), dev(W2), dev(R), dev(phi))
        assert ot.rel_err(out.numpy(), want[lo:hi]) < 1e-12
        slabs.append(out.data)
    full = tn.shard.assemble_gathered(torch.cat(slabs), chi, d, d, cr, world)
    assert ot.rel_err(full.cpu().numpy().reshape((chi, d, d, cr), order="F"), want) < 1e-12


def _fused_worker(rank, world, cplx):
    import torch
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(52)
    chi, cr, d, w = 64, 48, 2, 5
    L = rand(rng, (chi, chi, w), cplx); R = rand(rng, (cr, cr, w), cplx)
    W1 = rand(rng, (w, d, d, w), cplx); W2 = rand(rng, (w, d, d, w), cplx)
    phis = [rand(rng, (chi, d, d, cr), cplx) for _ in range(3)]
    D = tn.DTensor.from_numpy
    dL = D(L)
    lo, hi = tn.shard.slab_range(chi, rank, world)
    Ls = tn.DTensor(tn.shard.left_env_slab(dL.data, chi, w, rank, world), (chi, hi - lo, w))
    fh = tn.shard.FusedShardedHeff((chi, d, d, cr), torch.complex128 if cplx else torch.float64)
    errs = []
    for phi in phis:                                   # three epochs, alternating buffers
        out = fh.apply(Ls, D(W1), D(W2), D(R), D(phi))
        torch.cuda.synchronize()
        errs.append(ot.rel_err(out.numpy(), od.heff_apply(L, W1, W2, R, phi)))
    fh.status()
    fh.close()
    return max(errs)


# _fused_worker / _mpo_worker run at 2, 4 and 8 ranks inside tests/test_gpu_shard_dmrg.py::test_multi_rank_suite
# (GEMM + all-gather fused over peer memory; every rank must hold the full H*phi with no library collective).


def _mpo_worker(rank, world):
    from itensorsgpu_b200 import tn
    rng = np.random.default_rng(53)
    chi, cr, d, w = 48, 40, 2, 5
    errs = []
    for cplx in (False, True):
        L = rand(rng, (chi, chi, w), cplx); R = rand(rng, (cr, cr, w), cplx)
        W1 = rand(rng, (w, d, d, w), cplx); W2 = rand(rng, (w, d, d, w), cplx)
        phi = rand(rng, (chi, d, d, cr), cplx)
        D = tn.DTensor.from_numpy
        ms = tn.shard.MpoSplitHeff(D(L), D(W1), D(W2), D(R))
        out = ms.apply(D(phi))
        errs.append(ot.rel_err(out.numpy(), od.heff_apply(L, W1, W2, R, phi)))
    return max(errs)


def test_mpo_split_ranges():
    from itensorsgpu_b200 import tn
    assert tn.shard.mpo_split_ranges(5, 2) == [(0, 3), (3, 5)]
    assert tn.shard.mpo_split_ranges(5, 8) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    assert tn.shard.mpo_split_ranges(30, 8)[-1] == (27, 30)
