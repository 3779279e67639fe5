"""N>1 TEBD path on CPU: world_size-2 gloo processes exercise the block partition, the halo exchange of the
boundary site tensor and the hand-back of the updated tensor + Schmidt values (itensorsgpu.jl_b200/tebd.py).
The gate compute is the CUDA entry point on a GPU box (tests/test_gpu_tebd.py); here the oracle's B-form gate is
injected in its place so that only the host-side distribution logic is under test."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_gate(tn):
    from oracle import tebd as otebd

    def gate(G, lamL, B1, B2, maxdim=None, mindim=1, cutoff=0.0):
        Bs = [B1.data.numpy().reshape(B1.dims, order="F"), B2.data.numpy().reshape(B2.dims, order="F")]
        lams = [lamL.numpy(), None, None]
        err = otebd.apply_gate_bform(Bs, lams, G.data.numpy().reshape(G.dims, order="F"), 0, maxdim=maxdim, cutoff=cutoff)
        f = lambda a: tn.DTensor(torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))), a.shape)
        return f(Bs[0]), f(Bs[1]), torch.from_numpy(lams[1].copy()), err
    return gate


def _worker(rank, world, port, q, N):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from itensorsgpu_b200 import tn
    from oracle import models, mps as omps, tebd as otebd
    rng = np.random.default_rng(55)
    psi = omps.random_mps(N, 2, 8, rng, dtype=np.complex128)
    G = models.heisenberg_bond_gate(0.05, imaginary_time=False)
    Bs, lams = otebd.canonical_bform(psi)
    f = lambda a: tn.DTensor(torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))), a.shape)
    full = tn.tebd.BState([f(b) for b in Bs], [torch.from_numpy(l.copy()) for l in lams])
    sh = tn.tebd.ShardedTEBD.scatter_from(full, N, gate_fn=_oracle_gate(tn))
    sh.warm_links()
    Gd = f(G)
    for _ in range(2):
        sh.layer(Gd, 0, maxdim=10)
        sh.layer(Gd, 1, maxdim=10)
    out = sh.gather()
    for _ in range(2):
        otebd.tebd_layer_bform(Bs, lams, G, 0, maxdim=10)
        otebd.tebd_layer_bform(Bs, lams, G, 1, maxdim=10)
    a = omps.to_dense([t.data.numpy().reshape(t.dims, order="F") for t in out.Bs])
    b = omps.to_dense(Bs)
    lam_err = max(float(np.max(np.abs(x.numpy() - y))) for x, y in zip(out.lams, lams))
    q.put((rank, float(np.linalg.norm(a - b)), lam_err))
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [8, 14])
def test_sharded_tebd_plumbing_world2(N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + N
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, N)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in res) == [0, 1]
    assert all(e < 1e-13 and le < 1e-13 for _, e, le in res)


def test_block_range():
    sys.path.insert(0, ROOT)
    from itensorsgpu_b200 import tn
    assert tn.tebd.block_range(128, 0, 8) == (0, 16) and tn.tebd.block_range(128, 7, 8) == (112, 128)
    assert tn.tebd.block_range(10, 1, 2) == (4, 10)
    with pytest.raises(ValueError):
        tn.tebd.block_range(6, 0, 4)
