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
