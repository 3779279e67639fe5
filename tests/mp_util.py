"""Spawn `world` ranks for the multi-GPU tests.  With at least `world` GPUs every rank gets its own device and the
process group is NCCL; with fewer (the driver's 1-GPU box) the ranks share the devices round-robin and the group is
gloo -- the library's data path never uses the process group (CUDA IPC peer buffers + device-side flag barriers),
so the same code runs either way; only the NVLink hop degenerates to a local copy."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _entry(rank, world, port, q, fn, args):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch
        import torch.distributed as dist
        ndev = torch.cuda.device_count()
        dev = rank % max(ndev, 1)
        torch.cuda.set_device(dev)
        if ndev >= world:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        res = fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", res))
    except Exception:   # noqa: BLE001 -- report to the parent instead of dying silently
        q.put((rank, "error", traceback.format_exc()))


def run_ranks(fn, world, *args, timeout=600):
    """Run fn(rank, world, *args) in `world` spawned processes; returns the list of results ordered by rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 20000 + (os.getpid() * 7 + world * 131 + hash(fn.__name__) % 977) % 20000
    procs = [ctx.Process(target=_entry, args=(r, world, port, q, fn, args)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in procs:
            rank, status, res = q.get(timeout=timeout)
            if status != "ok":
                raise AssertionError("rank %d failed:\n%s" % (rank, res))
            out[rank] = res
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    return [out[r] for r in range(world)]
