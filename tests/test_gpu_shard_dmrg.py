"""Multi-GPU path of the DMRG tier (peer group, all-gather over peer memory, sharded environment updates, sharded
Lanczos, sharded bond step, dmrg(..., comm=...), end-to-end host-buffer matvec) against the oracle and against the
single-GPU path, at 2, 4 and 8 ranks.  Runs on whatever the box has: one rank per GPU when there are enough,
ranks sharing GPUs otherwise (tests/mp_util.py) -- no test here is skipped on a 1-GPU box."""
import numpy as np
import pytest

from mp_util import run_ranks

pytestmark = pytest.mark.gpu
WORLDS = [2, 4, 8]


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = (a + 1j * rng.standard_normal(shape)) / np.sqrt(2)
    return a


def _w_allgather(rank, world, comm):
    import torch
    from itensorsgpu_b200 import tn
    n = 1000
    bufs = comm.buffer("t", 8 * n * world + 4096)
    errs = 0
    for rep in range(3):
        loc = bufs.local(torch.float64)
        loc[rank * n:(rank + 1) * n] = torch.arange(n, dtype=torch.float64, device="cuda") + 1e6 * rank + 1e3 * rep
        comm.allgather(bufs, 8 * n)
        torch.cuda.synchronize()
        for g in range(world):
            want = torch.arange(n, dtype=torch.float64, device="cuda") + 1e6 * g + 1e3 * rep
            errs += int(not torch.equal(loc[g * n:(g + 1) * n], want))
    comm.status()
    return errs


def _w_env(rank, world, comm, cplx):
    import torch
    from itensorsgpu_b200 import tn
    from oracle import dmrg as od
    from oracle import tensor as ot
    rng = np.random.default_rng(71)
    cl, cr, d, wl, wr = 8 * world, 16 * world, 2, 5, 3
    A = _rand(rng, (cl, d, cr), cplx); W = _rand(rng, (wl, d, d, wr), cplx)
    L = _rand(rng, (cl, cl, wl), cplx); R = _rand(rng, (cr, cr, wr), cplx)
    D = tn.DTensor.from_numpy
    dt = torch.complex128 if cplx else torch.float64
    sh = tn.shard.ShardedSweep(comm, dt, max(cl, cr), d, max(wl, wr), min_chi=world)
    errs = []
    for rep in range(2):            # twice: staging buffers are reused
        Ls = sh.to_slab(D(L))
        got = sh.env_left(Ls, D(A), D(W))
        want = od.env_left_update(L, A, W)
        lo, hi = tn.shard.slab_range(cr, rank, world)
        assert got.dims == (cr, hi - lo, wr)
        errs.append(ot.rel_err(got.numpy(), want[:, lo:hi, :]))
        errs.append(ot.rel_err(sh.to_full(got).numpy(), want))            # slab -> replicated round trip
        gotR = sh.env_right(D(R), D(A), D(W))
        errs.append(ot.rel_err(gotR.numpy(), od.env_right_update(R, A, W)))
    comm.status()
    return max(errs)


def _w_lanczos_and_host(rank, world, comm, cplx):
    import ctypes as C
    import torch
    from itensorsgpu_b200 import tn
    from oracle import dmrg as od
    from oracle import tensor as ot
    rng = np.random.default_rng(72)
    cl, cr, d, w = 8 * world, 24, 2, 5
    L = _rand(rng, (cl, cl, w), cplx); L = L + np.conj(np.transpose(L, (1, 0, 2)))          # Hermitian H_eff
    R = _rand(rng, (cr, cr, w), cplx); R = R + np.conj(np.transpose(R, (1, 0, 2)))
    W1 = _rand(rng, (w, d, d, w), cplx); W1 = W1 + np.conj(np.transpose(W1, (0, 2, 1, 3)))
    W2 = _rand(rng, (w, d, d, w), cplx); W2 = W2 + np.conj(np.transpose(W2, (0, 2, 1, 3)))
    phi = _rand(rng, (cl, d, d, cr), cplx)
    D = tn.DTensor.from_numpy
    dt = torch.complex128 if cplx else torch.float64
    sh = tn.shard.ShardedSweep(comm, dt, max(cl, cr), d, w, min_chi=world)
    Ls = sh.to_slab(D(L))
    # sharded Lanczos vs the single-GPU call on the same operands
    p1 = D(phi); e1, n1 = tn.ops.eigsolve_lanczos(D(L), D(W1), D(W2), D(R), p1)
    p2 = D(phi)
    h = comm.h
    bd = tn.ops.BondDims(cl, cr, d, d, w, w, w)
    e2 = C.c_double(0.0); n2 = C.c_int(0)
    h.check(h.lib.tnb_eigsolve_lanczos_shard(h.h, tn.ops._dt(p2.data), C.byref(bd), tn.ops._ptr(Ls.data), tn.ops._ptr(D(W1).data),
                                             tn.ops._ptr(D(W2).data), tn.ops._ptr(D(R).data), tn.ops._ptr(p2.data),
                                             sh.out_a.c_array(), sh.out_b.c_array(), 3, 1, 1e-14, C.byref(e2), C.byref(n2),
                                             tn.ops._stream()))
    errs = [abs(e1 - e2.value) / abs(e1), ot.rel_err(p2.numpy(), p1.numpy()), float(n1 != n2.value)]
    # end-to-end host-buffer matvec: this rank's r-chunk up, this rank's l' slab down
    hh = tn.shard.ShardedHeffHost(comm, (cl, d, d, cr), dt)
    ph = torch.from_numpy(np.ascontiguousarray(phi.ravel(order="F"))).pin_memory()
    oh = torch.zeros_like(ph).pin_memory()
    want = od.heff_apply(L, W1, W2, R, phi)
    lo, hi = tn.shard.slab_range(cl, rank, world)
    for rep in range(2):
        oh.zero_()
        hh.apply_host(Ls, D(W1), D(W2), D(R), ph, oh)
        got = oh.numpy().reshape((cl, d, d, cr), order="F")
        errs.append(ot.rel_err(got[lo:hi], want[lo:hi]))
        mask = np.ones(cl, bool); mask[lo:hi] = False
        errs.append(float(np.abs(got[mask]).max()) if mask.any() else 0.0)       # nothing outside the slab is written
        errs.append(ot.rel_err(hh.device_result().numpy(), want))                # the device copy is the full vector
    comm.status()
    return max(errs)


def _w_dmrg(rank, world, comm, noise):
    import torch
    from itensorsgpu_b200 import tn
    N, chi = 12, 8 * world
    H = tn.cu(tn.heisenberg_mpo(N, 0.5))
    psi0 = tn.randomCuMPS(N, 2, chi=chi, seed=5)
    kw = dict(maxdim=chi, cutoff=1e-11 if noise else 0.0, noise=1e-8 if noise else 0.0)
    sw = tn.Sweeps(2, **kw)
    ref = []
    e1, _ = tn.dmrg(H, psi0, sw, observer=lambda s, b, o, e, err: ref.append(e))
    got = []
    e2, psi = tn.dmrg(H, psi0, sw, comm=comm, shard_min_chi=world, verify_ranks=True,
                      observer=lambda s, b, o, e, err: got.append(e))
    stats = psi.shard_stats
    comm.status()
    dev = max(abs(a - b) for a, b in zip(ref, got))
    return dev, stats["sharded_bond_steps"], len(ref) == len(got), e2


def _suite(rank, world):
    """every multi-rank check in ONE process group (spawning 8 CUDA processes costs more than the checks)"""
    import test_gpu_shard as tgs
    import test_gpu_tebd as tgt
    from itensorsgpu_b200 import tn
    out = {}
    for cplx in (False, True):
        out["fused_gather_%s" % ("c128" if cplx else "f64")] = tgs._fused_worker(rank, world, cplx)
    out["mpo_split"] = tgs._mpo_worker(rank, world)
    out["tebd"] = tgt._worker(rank, world)
    comm = tn.shard.ShardComm()
    out["allgather_mismatches"] = _w_allgather(rank, world, comm)
    for cplx in (False, True):
        tag = "c128" if cplx else "f64"
        out["env_updates_%s" % tag] = _w_env(rank, world, comm, cplx)
        out["lanczos_host_%s" % tag] = _w_lanczos_and_host(rank, world, comm, cplx)
    for noise in (False, True):
        out["dmrg_noise" if noise else "dmrg_svd"] = _w_dmrg(rank, world, comm, noise)
    comm.close()
    return out


@pytest.mark.parametrize("world", WORLDS)
def test_multi_rank_suite(world):
    """2, 4 and 8 ranks: fused GEMM + all-gather over peer memory (F64, C128), MPO-bond split, sharded TEBD layers,
    all-gather over peer memory, sharded environment updates, sharded Lanczos + host-buffer matvec, and a 2-sweep
    dmrg(comm=...) on both factorize branches: every bond energy equals the single-GPU sweep's to 1e-12, the ranks
    stay bit-identical (verify_ranks) and the sharded code path is actually taken."""
    res = run_ranks(_suite, world, timeout=1500)
    for r in res:
        for k in ("fused_gather_f64", "fused_gather_c128", "mpo_split", "env_updates_f64", "env_updates_c128",
                  "lanczos_host_f64", "lanczos_host_c128"):
            assert r[k] < 1e-12, (k, r[k])
        assert r["allgather_mismatches"] == 0
        assert r["tebd"][0] < 1e-12 and r["tebd"][1]
        for k in ("dmrg_svd", "dmrg_noise"):
            dev, nshard, same_len, _ = r[k]
            assert same_len and dev < 1e-12 and nshard > 0, (k, r[k])
    for k in ("dmrg_svd", "dmrg_noise"):
        assert len({r[k][3] for r in res}) == 1            # identical final energy on every rank
