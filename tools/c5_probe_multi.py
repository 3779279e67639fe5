"""Config C5 central bond (chi = 8192, d = 2, dense MPO bond w = 30, Float64) on N GPUs, one process per GPU:
sharded Lanczos (3 matvecs with the gather fused into step 4), the sharded bond step (svd rule) and the sharded
environment update.  Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/c5_probe_multi.py [chi] [out.json]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itensorsgpu_b200 import tn  # noqa: E402

D, W = 2, 30
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/r02_c5_chi%d_multi.json" % chi
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = tn.handle()
g = torch.Generator(device="cuda").manual_seed(5)           # same operands on every rank


def r(*d):
    return tn.DTensor(torch.randn(int(np.prod(d)), device="cuda", dtype=torch.float64, generator=g), d)


def herm(E, n, w):
    v = E.data.view(w, n, n)
    for a in range(w):
        v[a].add_(v[a].T.clone())


L, R = r(chi, chi, W), r(chi, chi, W)
herm(L, chi, W); herm(R, chi, W)
W1, W2 = r(W, D, D, W), r(W, D, D, W)
for Wt in (W1, W2):
    v = Wt.data.view(W, D, D, W)
    v.add_(v.transpose(1, 2).clone())
A1 = r(chi, D, chi); A2 = r(chi, D, chi)
A1.data.mul_(1.0 / np.sqrt(chi * D)); A2.data.mul_(1.0 / np.sqrt(chi * D))
comm = tn.shard.ShardComm()
sh = tn.shard.ShardedSweep(comm, torch.float64, chi, D, W, min_chi=world)
Ls = sh.to_slab(L)
del L
torch.cuda.empty_cache()


def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


res = {"config": "C5 central bond: chi=%d, d=%d, dense MPO bond w=%d, Float64, %d GPUs (output bond sharded)" % (chi, D, W, world)}
# ---- sharded Lanczos: 3 matvecs
phi, _ = tn.ops.contract(A1, ("l", "s1", "k"), A2, ("k", "s2", "r"))
bd = tn.ops.BondDims(chi, chi, D, D, W, W, W)
e = C.c_double(0.0); nmv = C.c_int(0)
for rep in range(2):
    p2 = phi.clone()
    sync()
    t0 = time.perf_counter()
    h.check(h.lib.tnb_eigsolve_lanczos_shard(h.h, 0, C.byref(bd), tn.ops._ptr(Ls.data), tn.ops._ptr(W1.data), tn.ops._ptr(W2.data),
                                             tn.ops._ptr(R.data), tn.ops._ptr(p2.data), sh.out_a.c_array(), sh.out_b.c_array(), 3, 1, 1e-14,
                                             C.byref(e), C.byref(nmv), tn.ops._stream()))
    sync()
    lan = time.perf_counter() - t0
F = 2.0 * D * D * W * 2 * chi ** 3 + 4.0 * D ** 3 * W * W * chi * chi
res.update(lanczos_3_matvecs_s=lan, heff_tflops_aggregate_incl_vector_ops=3 * F / lan * 1e-12, energy=e.value)
del phi, p2
# ---- sharded bond step (svd rule) and environment update
sync()
t0 = time.perf_counter()
en, B1, B2, err = sh.bond_step(Ls, W1, W2, R, A1, A2, "left", maxdim=chi, mindim=1, cutoff=0.0, noise=0.0, krylovdim=3, maxiter=1,
                              which_decomp=None)
sync()
res["bond_step_svd_rule_s"] = time.perf_counter() - t0
res["bond_step_kept"] = B1.dims[2]
sync()
t0 = time.perf_counter()
Ln = sh.env_left(Ls, B1, W1)
sync()
res["env_update_left_s"] = time.perf_counter() - t0
res["workspace_GB"] = h.workspace_bytes / 1e9
free, total = torch.cuda.mem_get_info()
res["hbm_in_use_GB_rank0"] = (total - free) / 1e9
res["stage_buffer_GB"] = sh.stage.nbytes / 1e9
comm.status()
if rank == 0:
    print(json.dumps(res, indent=1), flush=True)
    json.dump(res, open(out_path, "w"), indent=1)
comm.close()
dist.destroy_process_group()
