"""Config C4: TEBD even/odd gate layers on an N-site chain at maxdim chi, distributed over the ranks
(one process per GPU; launch with torchrun for more than one).

  python tools/tebd_c4.py [--N 128] [--chi 2048] [--dtype c128|f64] [--pairs 1] [--out file.json]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/tebd_c4.py ...

Synthetic state: every site tensor a random right-isometry with bond dims
min(2^k, 2^(N-k), chi) (made ON the device with torch's QR: input generation, not the measured path) and
normalised decaying Schmidt vectors; gates exp(-i tau h) (c128) or exp(-tau h) (f64) of the Heisenberg bond
term, maxdim chi, cutoff 1e-12.  One step = one even layer + one odd layer (N-1 gates over all ranks).
Time = CUDA-synchronised wall clock, barrier on both sides, MAX over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from itensorsgpu_b200 import tn  # noqa: E402


def bond_gate(tau, real_time):
    Sz = np.diag([0.5, -0.5]); Sp = np.array([[0, 1.0], [0, 0]]); Sm = Sp.T
    h = np.kron(Sz, Sz) + 0.5 * (np.kron(Sp, Sm) + np.kron(Sm, Sp))
    w, v = np.linalg.eigh(h)
    U = (v * np.exp((-1j if real_time else -1.0) * tau * w)) @ v.conj().T
    U = U.reshape(2, 2, 2, 2)                     # [s1', s2', s1, s2] with row-major kron index (s1, s2)
    return U if real_time else U.real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--dtype", default="c128", choices=["c128", "f64"])
    ap.add_argument("--pairs", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cplx = a.dtype == "c128"
    dt = torch.complex128 if cplx else torch.float64
    N, chi, d = a.N, a.chi, 2
    D = [int(min(chi, d ** min(k, N - k, 40))) for k in range(N + 1)]
    lo, hi = tn.tebd.block_range(N, rank, world) if world > 1 else (0, N)
    g = torch.Generator(device="cuda").manual_seed(4242 + rank)
    Bs, lams = [], []
    for j in range(lo, hi):
        l, r = D[j], D[j + 1]
        G0 = torch.randn(d * r, l, dtype=dt, device="cuda", generator=g)
        Q = torch.linalg.qr(G0).Q if d * r >= l else G0 / G0.norm()
        Bs.append(tn.DTensor(Q.contiguous().reshape(-1).clone(), (l, d, r)))
        del G0, Q
    for j in range(lo, hi + 1):
        s = torch.exp(-6.0 * torch.arange(D[j], device="cuda", dtype=torch.float64) / max(D[j], 1))
        lams.append(s / s.norm())
    st = tn.tebd.BState(Bs, lams, first=lo)
    Gd = tn.DTensor.from_numpy(bond_gate(0.05, cplx))
    kw = dict(maxdim=chi, cutoff=1e-12)
    if world > 1:
        sh = tn.tebd.ShardedTEBD(st, N)
        layer = lambda p: sh.layer(Gd, p, **kw)
    else:
        layer = lambda p: tn.tebd.tebd_layer(st, Gd, p, **kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        layer(0); layer(1)
    h = tn.handle()
    l0 = h.launches
    barrier()
    t0 = time.perf_counter()
    te = 0.0
    for _ in range(a.pairs):
        t1 = time.perf_counter()
        layer(0)
        barrier()
        te += time.perf_counter() - t1
        layer(1)
        barrier()
    secs = time.perf_counter() - t0
    t = torch.tensor([secs, te], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs, te = t.tolist()
    if rank == 0:
        res = {"config": "C4 TEBD: N=%d chain, maxdim %d, %s, cutoff 1e-12, even+odd layer of Heisenberg bond gates" % (N, chi, a.dtype),
               "n_gpus": world, "gates_per_step": N - 1, "steps": a.pairs,
               "seconds_per_step": secs / a.pairs, "even_layer_seconds": te / a.pairs,
               "odd_layer_seconds": (secs - te) / a.pairs, "ms_per_gate_per_gpu": secs / a.pairs / ((N - 1) / world) * 1e3,
               "maxlinkdim_after": st.maxlinkdim(), "gpu_launches_rank0": h.launches - l0,
               "parallelism": "single GPU" if world == 1 else "contiguous site blocks x%d, halo send/recv of one site tensor per boundary (NCCL p2p), no collective" % world}
        print(json.dumps(res), flush=True)
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
