#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the hot path.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the central
bond of the S=1/2 Heisenberg chain DMRG at maxdim chi=4096, Float64 (d=2, MPO bond w=5).
One *step* = one H_eff*phi (the [EXT] `product(::ProjMPO, phi)` inside the eigensolver):
four pairwise contractions, F = 2 d^2 w (chiL^2 chiR + chiL chiR^2) + 4 d^3 w^2 chiL chiR
= 5.51e12 flop at chi=4096.

  value     : H_eff*phi FP64 TFLOP/s, operands resident in HBM, whole job (all ranks).
  e2e       : same through the C-ABI call with HOST buffers (tnb_heff_apply_host): pinned-host
              phi in, H phi out, H2D + D2H inside the timed region, environments resident
              (like `cu(psi)`/`cu(H)` once in the reference: src/mps/cumps.jl:1-9).
  roofline  : dominant kernel = the DMMA contraction kernel on H_eff steps 1 and 4; achieved
              = algorithmic flops / CUDA-event time of those launches; peak = FP64 DMMA peak
              measured in this run (tools/dmma_peak; MEASURED_PEAKS.json has no FP64 entry).
  N > 1     : the output bond l' is sharded (each rank 1/N of every contraction, no reduction)
              followed by an NCCL all-gather of H phi; strong scaling.
  --impl reference : the CPU restatement of the ITensors.jl path (oracle/dmrg.py, OpenBLAS
              dgemm through the same four pairwise contractions) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DMRG sweep seconds @ χ=4096; H_eff·ψ FP64 TFLOP/s vs tensor-core peak"
UNIT = "TFLOP/s (H_eff·ψ, FP64)"
D, W = 2, 5


def heff_flops(chi, d=D, w=W):
    return 2.0 * d * d * w * (2 * chi ** 3) + 4.0 * d ** 3 * w * w * chi * chi


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return len(os.sched_getaffinity(0))


# ------------------------------------------------------------------ CPU arm
def cpu_heff_tflops(chi, reps, warm=0, seed=2024):
    import numpy as np
    from oracle import dmrg as od
    rng = np.random.default_rng(seed)
    L = rng.standard_normal((chi, chi, W)); R = rng.standard_normal((chi, chi, W))
    W1 = rng.standard_normal((W, D, D, W)); W2 = rng.standard_normal((W, D, D, W))
    phi = rng.standard_normal((chi, D, D, chi))
    for _ in range(warm):
        od.heff_apply(L, W1, W2, R, phi)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        od.heff_apply(L, W1, W2, R, phi)
        ts.append(time.perf_counter() - t0)
    return heff_flops(chi) / (sum(ts) / len(ts)) * 1e-12, sum(ts)


def pick_cpu_chi(budget_s, steps):
    """Largest chi in {4096, 2048, 1024} whose `steps` applies fit the budget (calibrated at 1024)."""
    tf, _ = cpu_heff_tflops(1024, 1, warm=1)
    for chi in (4096, 2048, 1024):
        if heff_flops(chi) / (tf * 1e12) * steps <= budget_s:
            return chi, tf
    return 1024, tf


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    chi, _ = pick_cpu_chi(150.0, steps + warm)
    tf, total = cpu_heff_tflops(chi, steps, warm=warm)
    cores = host_threads()
    sample = "H_eff*phi at chi=%d, d=2, w=5 (%d timed applies, %.1f s); same 4 pairwise contractions" % (chi, steps, total)
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": total / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 central-bond H_eff*phi (S=1/2 Heisenberg, N=100, maxdim 4096, d=2, w=5)",
                   "chi": 4096, "reference_sample_chi": chi},
        "cpu_baseline": {"value": tf, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measure_fp64_peak():
    exe = os.path.join(ROOT, "tools", "dmma_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        best = 0.0
        for ln in out.splitlines():
            try:
                j = json.loads(ln)
            except ValueError:
                continue
            if j.get("kernel") == "dmma_8x8x4":
                best = max(best, j["tflops"])
        if best > 0:
            return best, "measured in this run: DMMA.8x8x4 issue-rate microbenchmark (tools/dmma_peak.cu)"
    except Exception:
        pass
    return 37.2, "fallback: 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (microbenchmark unavailable)"


def sweep_sample(tn, chi, branch):
    """Bounded sample of metric M1 (sweep seconds): one real two-site DMRG sweep through the public dmrg() on a
    SHORT S=1/2 Heisenberg chain whose bonds take every (chiL, chiR) shape of the N=100 chain (bond dims
    min(2^k, 2^(N-k), chi)); the N=100 sweep time is then re-assembled bond by bond from the measured
    per-shape times (198 bond steps).  tools/sweep_c3.py measures the full N=100 sweep directly
    (profiles/*_sweep_c3_*.json)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sweep_c3 import random_iso_mps
    lg = max(1, (chi - 1).bit_length())
    Ns = 2 * lg + 4                      # three full-size bonds per half sweep
    psi, Dm = random_iso_mps(Ns, 2, chi)
    H = tn.cu(tn.heisenberg_mpo(Ns, 0.5))
    kw = dict(maxdim=chi, cutoff=0.0, noise=0.0) if branch == "svd" else dict(maxdim=chi, cutoff=1e-11, noise=1e-10)
    marks = []
    h = tn.handle()
    l0 = h.launches
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e, _ = tn.dmrg(H, psi, tn.Sweeps(1, **kw), observer=lambda sw, b, o, en, err: marks.append((b, o, time.perf_counter())))
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    per = {}
    prev = t0
    first = True
    for (b, o, now) in marks:
        if not first:                    # the first callback also contains the right-environment build
            per.setdefault((o, Dm[b], Dm[b + 2]), []).append(now - prev)
        first = False
        prev = now
    tshape = {k: min(v) for k, v in per.items()}
    # re-assemble the N=100 chain
    N = 100
    D100 = [int(min(chi, 2 ** min(k, N - k, 40))) for k in range(N + 1)]
    est, missing = 0.0, 0
    for o in ("left", "right"):
        for b in range(N - 1):
            key = (o, D100[b], D100[b + 2])
            other = ("right" if o == "left" else "left", D100[b], D100[b + 2])
            if key in tshape:
                est += tshape[key]
            elif other in tshape:       # the very first bond's callback also holds the environment build: use the return pass
                est += tshape[other]
            else:
                missing += 1
    full = [tshape.get((o, chi, chi)) for o in ("left", "right")]
    return {"seconds_N100_reassembled": est, "bond_steps": 2 * (N - 1), "shapes_missing": missing,
            "central_bond_step_ms": [None if x is None else x * 1e3 for x in full],
            "branch": branch, "params": kw, "energy_after_sample_sweep": e,
            "sample": "one full DMRG sweep on a %d-site chain (all (chiL,chiR) shapes of the N=100 chain at maxdim %d), "
                      "%.1f s incl. environment build; per-shape bond-step times (bond step + environment update, "
                      "host clock around synchronous C calls) summed over the 198 bonds of N=100" % (Ns, chi, total),
            "gpu_launches": h.launches - l0}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from itensorsgpu_b200 import tn
    h = tn.handle()
    chi = args.chi
    if chi % world:
        raise SystemExit("chi must be divisible by the number of GPUs")
    steps, warm = args.steps, max(args.warmup, 3)
    g = torch.Generator(device="cuda").manual_seed(2024)   # same seed on every rank -> identical operands

    def rnd(n):
        return torch.randn(n, device="cuda", dtype=torch.float64, generator=g)

    Lfull = rnd(chi * chi * W)
    R = tn.DTensor(rnd(chi * chi * W), (chi, chi, W))
    W1 = tn.DTensor(rnd(W * D * D * W), (W, D, D, W))
    W2 = tn.DTensor(rnd(W * D * D * W), (W, D, D, W))
    phi = tn.DTensor(rnd(chi * D * D * chi) / (2.0 * chi), (chi, D, D, chi))
    F = heff_flops(chi)
    clp = chi // world
    if world == 1:
        L = tn.DTensor(Lfull, (chi, chi, W))
        out = tn.DTensor.empty(phi.dims)
        step = lambda: tn.ops.heff_apply(L, W1, W2, R, phi, out=out)
    else:
        # slab L[:, l'_shard, :] of this rank, made contiguous once (environments never move)
        L = tn.DTensor(tn.shard.left_env_slab(Lfull, chi, W, rank, world), (chi, clp, W))
        del Lfull
        slab = tn.DTensor.empty((clp, D, D, chi))
        gathered = torch.empty(world * clp * D * D * chi, device="cuda", dtype=torch.float64)

        def step_nccl():
            tn.ops.heff_apply_shard(L, W1, W2, R, phi, out=slab)
            dist.all_gather_into_tensor(gathered, slab.data)   # rank-major slabs == [r', s2', s1', G, l'_shard]

        # default: the all-gather fused into the last GEMM (NVLink peer stores + device-side flag barrier);
        # --gather nccl (or a peer-mapping failure, reported in config) uses library all-gather instead
        fused = None
        gather_mode = "nccl all-gather after the slab kernels"
        if args.gather == "fused":
            try:
                fused = tn.shard.FusedShardedHeff(phi.dims, torch.float64)
                gather_mode = "all-gather fused into the step-4 GEMM epilogue (NVLink peer stores, device flag barrier)"
            except Exception as ex:   # noqa: BLE001
                gather_mode = "nccl all-gather (peer mapping unavailable: %s)" % str(ex)[:80]
        if args.shard == "mpo":
            # the north star's MPO-bond split, for comparison (parallelism capped by w = 5)
            Lfull2 = tn.DTensor(rnd(chi * chi * W), (chi, chi, W))
            mpo = tn.shard.MpoSplitHeff(Lfull2, W1, W2, R)
            gather_mode = "MPO-bond split: c-plane reduce to owners + all-reduce of H*phi (NCCL)"
            fused = None

            holder = {}

            def step():
                holder["out"] = mpo.apply(phi).data
        elif fused is not None:
            def step():
                fused.apply(L, W1, W2, R, phi)
        else:
            step = step_nccl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = h.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    secs = e0.elapsed_time(e1) * 1e-3
    launches = h.launches - l0
    if world > 1:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = t.item()
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    clocks = sampler.stop() if rank == 0 else None
    value = F * steps / secs * 1e-12

    # ---- roofline of the dominant kernel (contract_kernel on steps 1 and 4), CUDA events on the launch stream
    T1 = tn.DTensor.empty((D, D, chi, clp, W))
    T3 = tn.DTensor(T1.data, (chi, clp, D, D, W))
    o4 = tn.DTensor.empty((clp, D, D, chi))
    k1 = lambda: tn.ops.contract(phi, ("l", "s1", "s2", "r"), L, ("l", "lp", "a"), out=T1)
    k4 = lambda: tn.ops.contract(T3, ("r", "lp", "s1p", "s2p", "c"), R, ("r", "rp", "c"), out=o4)
    for _ in range(2):
        k1(); k4()
    torch.cuda.synchronize()
    r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
    nrep = max(2, min(steps, 10))
    r0.record()
    for _ in range(nrep):
        k1(); k4()
    r1.record()
    torch.cuda.synchronize()
    kern_s = r0.elapsed_time(r1) * 1e-3 / (2 * nrep)
    flops_per_launch = 2.0 * (D * D * chi) * (clp * W) * chi      # identical for steps 1 and 4
    del T1, T3, o4

    # ---- e2e through the host-buffer C-ABI entry (N=1) / sharded + gather with host staging (N>1)
    ph = phi.data.cpu().pin_memory()
    oh = torch.empty_like(ph).pin_memory()
    nbytes = ph.numel() * 8
    if world == 1:
        e2e_step = lambda: tn.ops.heff_apply_host(L, W1, W2, R, ph, oh, phi.dims)
    else:
        def e2e_step():
            phi.data.copy_(ph, non_blocking=True)
            step()
            if rank == 0:
                if args.shard == "mpo":
                    src = holder["out"]
                else:
                    src = fused.outs[(fused.epoch - 1) % len(fused.outs)].local() if fused is not None else gathered
                oh.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ne2e = max(2, min(steps, 10))
    for _ in range(ne2e):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / ne2e
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()

    if rank == 0:
        peak, peak_src = measure_fp64_peak()
        ach = flops_per_launch / kern_s * 1e-12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cchi, _ = pick_cpu_chi(45.0, 2)
            ctf, ctot = cpu_heff_tflops(cchi, 2, warm=0)
            cpu = {"value": ctf, "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": "oracle (NumPy/OpenBLAS dgemm) H_eff*phi at chi=%d, 2 applies, %.1f s" % (cchi, ctot)}
        sweep = None
        if world == 1 and not args.no_sweep:
            del L, R, phi, out
            torch.cuda.empty_cache()
            sweep = sweep_sample(tn, chi, args.sweep_branch)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("contract_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 central-bond H_eff*phi (S=1/2 Heisenberg, N=100, maxdim 4096, d=2, w=5)",
                       "chi": chi, "d": D, "w": W, "flop_per_step": F,
                       "l2": "operands 0.5-2.7 GB per contraction, far larger than the 126 MB L2 (no flush needed)",
                       "parallelism": "single GPU" if world == 1 else ("output bond l' sharded x%d; %s" % (world, gather_mode)
                                                                       if args.shard == "lp" else "x%d; %s" % (world, gather_mode))},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": traffic, "kernel": "contract_kernel<f64, A K-major, B K-major, 16-byte copies, Cfg<64x128x16, warp 32x64, 3 stages, 2 CTA/SM>> (H_eff steps 1 and 4)",
                         "flop_per_launch": flops_per_launch, "ms_per_launch": kern_s * 1e3, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": F / e2e_s * 1e-12, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches,
            "sweep": sweep,
            "clocks": clocks,
            "heff_frac_of_fp64_peak": value / (peak * world),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        if fused is not None:
            try:
                fused.status()
                fused.close()
            except Exception as ex:   # noqa: BLE001 -- the JSON line is out already; do not turn teardown into a failure
                print("bench.py: peer teardown: %s" % ex, file=sys.stderr)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chi", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="lp", choices=["lp", "mpo"],
                    help="N>1: shard the output bond l' (default) or the MPO bond (north-star plan, for comparison)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: how the slabs of H*phi reach every rank")
    ap.add_argument("--no-sweep", action="store_true", help="skip the DMRG sweep-seconds sample (metric M1)")
    ap.add_argument("--sweep-branch", default="svd", choices=["svd", "eigen"],
                    help="factorize rule of the sweep sample: svd = cutoff 0 / noise 0 (C3), eigen = cutoff 1e-11 / noise 1e-10")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
