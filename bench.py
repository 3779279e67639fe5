#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the hot path.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the central
bond of the S=1/2 Heisenberg chain DMRG at maxdim chi=4096, Float64 (d=2, MPO bond w=5).
One *step* = one H_eff*phi (the [EXT] `product(::ProjMPO, phi)` inside the eigensolver):
four pairwise contractions, F = 2 d^2 w (chiL^2 chiR + chiL chiR^2) + 4 d^3 w^2 chiL chiR
= 5.51e12 flop at chi=4096.

  value     : H_eff*phi FP64 TFLOP/s, operands resident in HBM, whole job (all ranks).
  e2e       : same through the C-ABI call with HOST buffers: pinned-host phi in, H phi out, H2D + D2H inside
              the timed region, environments resident (like `cu(psi)`/`cu(H)` once in the reference:
              src/mps/cumps.jl:1-9).  N = 1: tnb_heff_apply_host.  N > 1: tnb_heff_apply_shard_host -- every rank
              uploads 1/N of phi and downloads 1/N of H phi; the rest crosses NVLink.
  parity    : after the timed loops the result AT THE BENCHMARK SIZE is checked against the oracle on a grid of
              sampled output elements (oracle/dmrg.py on sliced environments), against the single-GPU
              tnb_heff_apply of the same operands (N > 1), across ranks, and host result vs device result;
              the run FAILS above 1e-12.
  roofline  : dominant kernel = the DMMA contraction kernel on H_eff steps 1 and 4; achieved
              = algorithmic flops / CUDA-event time of those launches; peak = FP64 DMMA peak
              measured in this run (tools/dmma_peak; MEASURED_PEAKS.json has no FP64 entry).
  sweep     : metric M1 -- one full two-site DMRG sweep of the N=100 chain at maxdim 4096 through dmrg(), run
              DIRECTLY (no re-assembly), for the svd rule (cutoff 0) and the reference examples' rule (eigen branch,
              noise 1e-10, cutoff 1e-11); at N > 1 through dmrg(comm=...) with the matvec / environment flops
              sharded and the Amdahl split (sharded vs replicated seconds) reported.
  tebd_c4   : config C4 -- one even+odd TEBD layer pair, N=128, maxdim 2048, ComplexF64, layers spread over the ranks.
  N > 1     : the output bond l' is sharded (each rank 1/N of every contraction, no reduction); the all-gather of
              H phi is fused into the step-4 GEMM epilogue (NVLink peer stores); strong scaling.
  --impl reference : the CPU restatement of the ITensors.jl path (oracle/dmrg.py, OpenBLAS dgemm through the same
              four pairwise contractions) at the SAME chi = 4096 on ALL host cores.
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
WORKLOAD = "C3 central-bond H_eff*phi (S=1/2 Heisenberg, N=100, maxdim 4096, d=2, w=5)"
D, W = 2, 5
PARITY_TOL = 1e-12


def heff_flops(chi, d=D, w=W):
    return 2.0 * d * d * w * (2 * chi ** 3) + 4.0 * d ** 3 * w * w * chi * chi


def all_cores():
    return len(os.sched_getaffinity(0))


# ------------------------------------------------------------------ CPU arm
def _unleash_blas_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm must use every host core.  Called
    before NumPy is imported in this process (bench.py imports it lazily), then enforced with threadpoolctl."""
    n = str(all_cores())
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = n


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return all_cores()


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


def run_reference(args):
    """The reference's CPU implementation of the path (the oracle port: Julia is not in this image) on the SAME
    config as the GPU arm: chi = 4096, every step one full H_eff*phi, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _unleash_blas_threads()
    import numpy  # noqa: F401  (first import happens with the thread settings above)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=all_cores())
    except Exception:
        pass
    chi = args.chi
    steps, warm = args.steps, args.warmup
    tf, total = cpu_heff_tflops(chi, steps, warm=warm)
    cores = blas_threads()
    sample = "H_eff*phi at chi=%d, d=2, w=5 (%d timed applies, %.1f s); same 4 pairwise contractions" % (chi, steps, total)
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": total / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chi": chi, "d": D, "w": W, "flop_per_step": heff_flops(chi),
                   "reference_sample_chi": chi, "host_cores": all_cores()},
        "cpu_baseline": {"value": tf, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(chi, reps=2):
    """cpu_baseline leg of the GPU arm: the reference arm in a clean child process (own BLAS thread pool)."""
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--chi", str(chi),
                              "--steps", str(reps), "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
        j = json.loads(out.stdout.strip().splitlines()[-1])
        return j["cpu_baseline"]
    except Exception as ex:   # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": all_cores(), "kind": "port", "sample": "failed: %s" % str(ex)[:120]}


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


def sampled_oracle_error(Lfull_flat, R, W1, W2, phi_flat, got_flat, chi, nl=24, nr=24, seed=7):
    """Parity at the benchmark size: H*phi on the grid (l' in lps) x (all s1', s2') x (r' in rps) from the ORACLE
    (oracle/dmrg.heff_apply on the environments sliced to those l', r' -- the same four pairwise contractions, seconds
    on the CPU) against the same elements of the GPU result.  Returns (max relative error, number of elements)."""
    import numpy as np
    import torch
    from oracle import dmrg as od
    rng = np.random.default_rng(seed)
    lps = np.sort(rng.choice(chi, nl, replace=False)); rps = np.sort(rng.choice(chi, nr, replace=False))
    lps[0], lps[-1], rps[0], rps[-1] = 0, chi - 1, 0, chi - 1                       # include the corners
    tl = torch.as_tensor(lps, device=Lfull_flat.device); tr = torch.as_tensor(rps, device=Lfull_flat.device)
    # flat column-major [x, y, a] == row-major view (a, y, x)
    Ls = Lfull_flat.view(W, chi, chi).index_select(1, tl).cpu().numpy().transpose(2, 1, 0)      # [l, l'_s, a]
    Rs = R.data.view(W, chi, chi).index_select(1, tr).cpu().numpy().transpose(2, 1, 0)          # [r, r'_s, c]
    phi = phi_flat.cpu().numpy().reshape((chi, D, D, chi), order="F")
    w1 = W1.data.cpu().numpy().reshape((W, D, D, W), order="F"); w2 = W2.data.cpu().numpy().reshape((W, D, D, W), order="F")
    want = od.heff_apply(Ls, w1, w2, Rs, phi)                                                    # [l'_s, s1', s2', r'_s]
    got = got_flat.view(chi, D, D, chi).index_select(3, tl).index_select(0, tr).cpu().numpy().transpose(3, 2, 1, 0)
    scale = float(np.sqrt(np.mean(np.abs(want) ** 2)))
    return float(np.max(np.abs(got - want)) / scale), int(want.size)


def sweep_direct(tn, chi, branch, comm=None, nsites=100):
    """Metric M1 measured directly: ONE full two-site sweep (2(N-1) bond steps) of the N=100 S=1/2 Heisenberg chain at
    maxdim chi through the public dmrg(), from a synthetic random-isometry MPS with bond dims min(2^k, 2^(N-k), chi)
    (input generation is not timed).  Host clock around synchronous C calls."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sweep_c3 import random_iso_mps
    psi, Dm = random_iso_mps(nsites, 2, chi)
    H = tn.cu(tn.heisenberg_mpo(nsites, 0.5))
    kw = dict(maxdim=chi, cutoff=0.0, noise=0.0) if branch == "svd" else dict(maxdim=chi, cutoff=1e-11, noise=1e-10)
    marks = []
    h = tn.handle()
    l0 = h.launches
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e, out = tn.dmrg(H, psi, tn.Sweeps(1, **kw), comm=comm,
                     observer=lambda sw, b, o, en, err: marks.append((b, o, time.perf_counter())))
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    first = marks[0][2] - t0                    # right-environment build + first (tiny) bond
    second = marks[1][2] - marks[0][2]
    env_build = max(0.0, first - second)
    mid = [marks[i][2] - marks[i - 1][2] for i in range(1, len(marks)) if marks[i][0] in (nsites // 2 - 1, nsites // 2)]
    full = [marks[i][2] - marks[i - 1][2] for i in range(1, len(marks))
            if Dm[marks[i][0]] == chi and Dm[marks[i][0] + 2] == chi]
    res = {"seconds": total - env_build, "measured": "directly: one dmrg() call, %d bond steps" % (2 * (nsites - 1)),
           "bond_steps": 2 * (nsites - 1), "right_environment_build_seconds": env_build,
           "central_bond_step_ms": [x * 1e3 for x in mid], "full_size_bond_steps": len(full),
           "full_size_bond_step_ms_mean": (sum(full) / len(full) * 1e3) if full else None,
           "branch": branch, "params": kw, "energy_after_sweep": e, "gpu_launches": h.launches - l0,
           "max_memory_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}
    if comm is not None:
        res["shard_stats"] = getattr(out, "shard_stats", None)
    del psi, H, out
    torch.cuda.empty_cache()
    return res


def tebd_c4(tn, world, rank, chi=2048, nsites=128, warm_gates=True):
    """Config C4: one even + one odd TEBD layer (nsites-1 gates, exp(-i tau h) Heisenberg bond gates, ComplexF64,
    maxdim chi, cutoff 1e-12) on a synthetic B-form state, layers spread over the ranks (tebd.ShardedTEBD)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tebd_c4 import bond_gate
    dt = torch.complex128
    d = 2
    Dm = [int(min(chi, d ** min(k, nsites - k, 40))) for k in range(nsites + 1)]
    lo, hi = tn.tebd.block_range(nsites, rank, world) if world > 1 else (0, nsites)
    g = torch.Generator(device="cuda").manual_seed(4242 + rank)
    Bs, lams = [], []
    for j in range(lo, hi):
        l, r = Dm[j], Dm[j + 1]
        G0 = torch.randn(d * r, l, dtype=dt, device="cuda", generator=g)
        Q = torch.linalg.qr(G0).Q if d * r >= l else G0 / G0.norm()
        Bs.append(tn.DTensor(Q.contiguous().reshape(-1).clone(), (l, d, r)))
        del G0, Q
    for j in range(lo, hi + 1):
        s = torch.exp(-6.0 * torch.arange(Dm[j], device="cuda", dtype=torch.float64) / max(Dm[j], 1))
        lams.append(s / s.norm())
    st = tn.tebd.BState(Bs, lams, first=lo)
    Gd = tn.DTensor.from_numpy(bond_gate(0.05, True))
    kw = dict(maxdim=chi, cutoff=1e-12)
    if world > 1:
        sh = tn.tebd.ShardedTEBD(st, nsites)
        sh.warm_links()        # NCCL opens its point-to-point channels on first use: not part of a layer
        layer = lambda p: sh.layer(Gd, p, **kw)
    else:
        layer = lambda p: tn.tebd.tebd_layer(st, Gd, p, **kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if warm_gates:   # one full-size gate on scratch copies (sizes the workspace arena; not a layer of the state)
        j = min(len(Bs) // 2, len(Bs) - 2)
        tn.ops.tebd_gate_bform(Gd, st.lams[j], st.Bs[j].clone(), st.Bs[j + 1].clone(), **kw)
    h = tn.handle()
    l0 = h.launches
    barrier()
    t0 = time.perf_counter()
    layer(0)
    barrier()
    t1 = time.perf_counter()
    layer(1)
    barrier()
    t2 = time.perf_counter()
    t = torch.tensor([t2 - t0, t1 - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs, te = t.tolist()
    return {"config": "C4 TEBD: N=%d chain, maxdim %d, ComplexF64, cutoff 1e-12, one even + one odd layer of Heisenberg "
                      "bond gates exp(-i 0.05 h)" % (nsites, chi),
            "seconds_per_layer_pair": secs, "even_layer_seconds": te, "odd_layer_seconds": secs - te,
            "gates": nsites - 1, "ms_per_gate_per_gpu": secs / ((nsites - 1) / world) * 1e3,
            "maxlinkdim_after": st.maxlinkdim(), "gpu_launches_rank0": h.launches - l0,
            "parallelism": "single GPU" if world == 1 else
            "contiguous site blocks x%d, halo send/recv of one site tensor per boundary (NCCL p2p), no collective" % world}


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
    fused = None
    comm = None
    holder = {}
    if world == 1:
        L = tn.DTensor(Lfull, (chi, chi, W))
        out = tn.DTensor.empty(phi.dims)
        step = lambda: tn.ops.heff_apply(L, W1, W2, R, phi, out=out)
        result = lambda: out
        gather_mode = None
    else:
        # slab L[:, l'_shard, :] of this rank, made contiguous once (environments never move)
        L = tn.DTensor(tn.shard.left_env_slab(Lfull, chi, W, rank, world), (chi, clp, W))
        if rank != 0:
            del Lfull                                   # rank 0 keeps the full L for the parity check
        slab = tn.DTensor.empty((clp, D, D, chi))
        gathered = torch.empty(world * clp * D * D * chi, device="cuda", dtype=torch.float64)

        def step_nccl():
            tn.ops.heff_apply_shard(L, W1, W2, R, phi, out=slab)
            dist.all_gather_into_tensor(gathered, slab.data)   # rank-major slabs == [r', s2', s1', G, l'_shard]

        # default: the all-gather fused into the last GEMM (NVLink peer stores + device-side flag barrier);
        # --gather nccl (or a peer-mapping failure, reported in config) uses library all-gather instead
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

            def step():
                holder["out"] = mpo.apply(phi)
            result = lambda: holder["out"]
        elif fused is not None:
            def step():
                holder["out"] = fused.apply(L, W1, W2, R, phi)
            result = lambda: holder["out"]
        else:
            step = step_nccl
            result = lambda: tn.DTensor(tn.shard.assemble_gathered(gathered, chi, D, D, chi, world), phi.dims)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

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
    secs = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = h.launches - l0
    if world > 1:
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    clocks = sampler.stop() if rank == 0 else None
    value = F * steps / secs * 1e-12

    # ---- parity of the timed path at the benchmark size
    parity = {"tolerance": PARITY_TOL, "ranks": world}
    res = result()
    torch.cuda.synchronize()
    if args.shard == "lp":
        if world > 1:     # every rank must hold the same full vector (bitwise: same stores from the same GEMMs)
            sig = torch.stack([res.data.sum(), res.data.abs().sum(), res.data[:: 4099].sum()])
            sigs = [torch.empty_like(sig) for _ in range(world)]
            dist.all_gather(sigs, sig)
            parity["ranks_bit_identical"] = bool(all(torch.equal(s, sigs[0]) for s in sigs))
        if rank == 0:
            t0 = time.perf_counter()
            err, ns = sampled_oracle_error(Lfull, R, W1, W2, phi.data, res.data, chi)
            parity.update({"max_rel_err": err, "n_samples": ns, "oracle_seconds": time.perf_counter() - t0,
                           "against": "oracle/dmrg.heff_apply on %d sampled output elements at chi=%d" % (ns, chi)})
            if world > 1:   # and the full vector against the single-GPU entry point on the same operands
                ref = tn.ops.heff_apply(tn.DTensor(Lfull, (chi, chi, W)), W1, W2, R, phi)
                parity["vs_single_gpu_rel_err"] = float((res.data - ref.data).norm() / ref.data.norm())
                del ref
        if world > 1 and rank == 0:
            del Lfull
            torch.cuda.empty_cache()

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

    # ---- e2e through the host-buffer C-ABI entry points
    ph = phi.data.cpu().pin_memory()
    oh = torch.zeros_like(ph).pin_memory()
    nbytes = ph.numel() * 8
    if world == 1:
        e2e_step = lambda: tn.ops.heff_apply_host(L, W1, W2, R, ph, oh, phi.dims)
        e2e_api = "tnb_heff_apply_host"
        h2d = d2h = nbytes
    elif args.shard == "lp" and fused is not None:
        comm = tn.shard.ShardComm()
        hh = tn.shard.ShardedHeffHost(comm, phi.dims, torch.float64)
        e2e_step = lambda: hh.apply_host(L, W1, W2, R, ph, oh)
        e2e_api = "tnb_heff_apply_shard_host (per rank: 1/N of phi up, its own 1/N of H*phi down overlapped with step 4; NVLink for the rest)"
        h2d = d2h = nbytes          # whole job: exactly one vector each way
    else:
        def e2e_step():
            phi.data.copy_(ph, non_blocking=True)
            step()
            if rank == 0:
                oh.copy_(result().data, non_blocking=True)
            torch.cuda.synchronize()
        e2e_api = "H2D of phi on every rank + step + D2H on rank 0"
        h2d, d2h = nbytes * world, nbytes
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ne2e = max(2, min(steps, 10))
    for _ in range(ne2e):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / ne2e)
    # parity of the host path: what landed in the host buffer against the device result of the timed loop
    if args.shard == "lp":
        if world == 1:
            herr = float((oh.cuda() - res.data).norm() / res.data.norm())
        elif comm is not None:
            lo = rank * clp
            mine = oh.view(D * D * chi, chi)[:, lo:lo + clp].cuda()
            want = res.data.view(D * D * chi, chi)[:, lo:lo + clp]
            herr = max_over_ranks(float((mine - want).norm() / want.norm()))
        else:
            herr = None
        parity["host_path_rel_err"] = herr
    if comm is not None:
        comm.status()

    # ---- the remaining keys need the big operands gone
    del L, R, phi
    res = None
    if world == 1:
        del out
    else:
        del slab, gathered
    holder.clear()
    if fused is not None:
        try:
            fused.status()
            fused.close()
        except Exception as ex:   # noqa: BLE001
            print("bench.py: peer teardown: %s" % ex, file=sys.stderr)
        fused = None
    torch.cuda.empty_cache()

    sweeps = None
    if not args.no_sweep:
        branches = ["svd", "eigen"] if args.sweep_branch == "both" else [args.sweep_branch]
        if world > 1:
            if comm is None:
                comm = tn.shard.ShardComm()
            branches = branches[:1]
        sweeps = [sweep_direct(tn, chi, b, comm=comm if world > 1 else None) for b in branches]
    if comm is not None:
        comm.status()
        comm.close()
        comm = None
        torch.cuda.empty_cache()
    tebd = None
    if not args.no_tebd:
        tebd = tebd_c4(tn, world, rank)

    if rank == 0:
        peak, peak_src = measure_fp64_peak()
        ach = flops_per_launch / kern_s * 1e-12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess(chi)
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if world == 1 and os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("contract_kernel_dram_bytes_per_launch")
                traffic_src = tj.get("source")
            except Exception:
                traffic = None
        bad = [k for k in ("max_rel_err", "vs_single_gpu_rel_err", "host_path_rel_err")
               if parity.get(k) is not None and not (parity[k] < PARITY_TOL)]
        if parity.get("ranks_bit_identical") is False:
            bad.append("ranks_bit_identical")
        parity["ok"] = not bad
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chi": chi, "d": D, "w": W, "flop_per_step": F,
                       "l2": "operands 0.5-2.7 GB per contraction, far larger than the 126 MB L2 (no flush needed)",
                       "parallelism": "single GPU" if world == 1 else ("output bond l' sharded x%d; %s" % (world, gather_mode)
                                                                       if args.shard == "lp" else "x%d; %s" % (world, gather_mode))},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "DMMA contraction kernels on H_eff steps 1 and 4 (M=%d, N=%d, K=%d per launch); the plan cache "
                                   "picks the TMA-staged (contract_tma_kernel) or the LDGSTS form (contract_kernel) per shape" % (D * D * chi, clp * W, chi),
                         "kernel_family_calls": h.kernel_family_counts(), "plan_cache": h.plan_cache_stats(),
                         "flop_per_launch": flops_per_launch, "ms_per_launch": kern_s * 1e3, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": F / e2e_s * 1e-12, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "api": e2e_api},
            "parity": parity,
            "gpu_launches": launches,
            "sweep": sweeps[0] if sweeps else None,
            "sweep_eigen_noise": sweeps[1] if sweeps and len(sweeps) > 1 else None,
            "tebd_c4": tebd,
            "clocks": clocks,
            "heff_frac_of_fp64_peak": value / (peak * world),
        }
        print(json.dumps(line), flush=True)
        if bad:
            print("bench.py: PARITY FAILURE at the benchmark size: %s" % {k: parity.get(k) for k in bad}, file=sys.stderr)
    else:
        bad = []
    if world > 1:
        dist.destroy_process_group()
    if bad:
        raise SystemExit(3)


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
    ap.add_argument("--no-sweep", action="store_true", help="skip the direct DMRG sweep (metric M1)")
    ap.add_argument("--no-tebd", action="store_true", help="skip the C4 TEBD layer pair")
    ap.add_argument("--sweep-branch", default="both", choices=["both", "svd", "eigen"],
                    help="factorize rule of the direct sweep: svd = cutoff 0 / noise 0 (C3), eigen = cutoff 1e-11 / noise 1e-10 "
                         "(examples/dmrg.jl:20-24); N>1 runs the first one only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
