"""N>1 path on CPU: world_size-2 gloo processes exercise the sharding plumbing (slab extraction,
all-gather, reassembly) of the output-bond-sharded H_eff*phi.  The slab compute itself is the CUDA
kernel on a GPU box (tests/test_gpu_shard.py); here the oracle stands in for it so that only the
host-side layout logic is under test."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from itensorsgpu_b200 import tn
    from oracle import dmrg as od
    rng = np.random.default_rng(99)          # same operands on every rank
    chi, d, w = 12, 2, 5
    L = rng.standard_normal((chi, chi, w)); R = rng.standard_normal((chi, chi, w))
    W1 = rng.standard_normal((w, d, d, w)); W2 = rng.standard_normal((w, d, d, w))
    phi = rng.standard_normal((chi, d, d, chi))
    Lf = torch.from_numpy(np.ascontiguousarray(L.ravel(order="F")))
    slab = tn.shard.left_env_slab(Lf, chi, w, rank, world)
    lo, hi = tn.shard.slab_range(chi, rank, world)
    Ls = slab.numpy().reshape((chi, hi - lo, w), order="F")
    assert np.array_equal(Ls, L[:, lo:hi, :])
    out_slab = od.heff_apply(Ls, W1, W2, R, phi)                      # stand-in for the CUDA slab kernel
    mine = torch.from_numpy(np.ascontiguousarray(out_slab.ravel(order="F")))
    gathered = torch.empty(world * mine.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, mine)
    full = tn.shard.assemble_gathered(gathered, chi, d, d, chi, world).numpy().reshape((chi, d, d, chi), order="F")
    want = od.heff_apply(L, W1, W2, R, phi)
    q.put((rank, float(np.linalg.norm(full - want) / np.linalg.norm(want))))
    dist.destroy_process_group()


def test_sharded_heff_plumbing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(e < 1e-14 for _, e in res)


def test_slab_range_errors():
    sys.path.insert(0, ROOT)
    from itensorsgpu_b200 import tn
    with pytest.raises(ValueError):
        tn.shard.slab_range(10, 0, 4)
