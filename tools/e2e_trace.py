"""Phase breakdown of the sharded host-buffer matvec (TNB_E2E_TRACE=1).  torchrun --nproc-per-node N tools/e2e_trace.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
chi, D, W = 4096, 2, 5
g = torch.Generator(device="cuda").manual_seed(2024)
rnd = lambda n: torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
Lfull = rnd(chi * chi * W)
R = tn.DTensor(rnd(chi * chi * W), (chi, chi, W)); W1 = tn.DTensor(rnd(W * D * D * W), (W, D, D, W)); W2 = tn.DTensor(rnd(W * D * D * W), (W, D, D, W))
L = tn.DTensor(tn.shard.left_env_slab(Lfull, chi, W, rank, world), (chi, chi // world, W)); del Lfull
ph = (rnd(chi * D * D * chi) / (2.0 * chi)).cpu().pin_memory(); oh = torch.zeros_like(ph).pin_memory()
comm = tn.shard.ShardComm()
hh = tn.shard.ShardedHeffHost(comm, (chi, D, D, chi), torch.float64)
for i in range(6):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    hh.apply_host(L, W1, W2, R, ph, oh)
    if rank == 0: print("step %d host %.3f ms" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
comm.close(); dist.destroy_process_group()
