"""Measure one full two-site DMRG sweep on config C3 (S=1/2 Heisenberg chain, N sites, maxdim chi, Float64)
through the library's public dmrg() (one fused C call per bond + one environment update).

  python tools/sweep_c3.py [--N 100] [--chi 4096] [--branch eigen|svd] [--out profiles/rXX_sweep.json]

Synthetic input as in SURVEY.md section 8(d): bond dims min(2^k, 2^(N-k), chi), every site tensor a random
isometry (generated ON THE DEVICE with torch's QR -- input generation, not the measured path), orthogonality
centre at site 1.  branch svd: cutoff 0, noise 0 (maxdim-only truncation); branch eigen: cutoff 1e-11,
noise 1e-10.  Reports wall seconds (host clock around synchronous calls) for: the initial right-environment
build, the right-moving and left-moving half sweeps, and per-bond detail at the central bond."""
import argparse
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from itensorsgpu_b200 import tn  # noqa: E402


def random_iso_mps(N, d, chi, seed=2024):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = [int(min(chi, d ** min(k, N - k, 40))) for k in range(N + 1)]
    ts = []
    for j in range(N):
        l, r = D[j], D[j + 1]
        G = torch.randn(d * r, l, dtype=torch.float64, device="cuda", generator=g)
        if d * r >= l:
            Q = torch.linalg.qr(G).Q
        else:
            Q = G / G.norm()
        ts.append(tn.DTensor(Q.contiguous().reshape(-1).clone(), (l, d, r)))
        del G, Q
    t0 = ts[0]
    ts[0] = tn.DTensor(t0.data / t0.data.norm(), t0.dims)
    return tn.MPS(ts, llim=-1, rlim=1), D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--chi", type=int, default=4096)
    ap.add_argument("--branch", default="eigen", choices=["eigen", "svd"])
    ap.add_argument("--sweeps", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    h = tn.handle()
    t0 = time.perf_counter()
    psi, D = random_iso_mps(a.N, 2, a.chi)
    H = tn.cu(tn.heisenberg_mpo(a.N, 0.5))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    kw = dict(maxdim=a.chi, cutoff=0.0, noise=0.0) if a.branch == "svd" else dict(maxdim=a.chi, cutoff=1e-11, noise=1e-10)
    marks = []
    state = {"t": None, "env_done": None}

    def obs(sw, b, o, en, err):
        now = time.perf_counter()
        if state["env_done"] is None:
            state["env_done"] = now      # first callback: env build + first bond are behind us
        marks.append((sw, b, o, now, en, err))

    l0 = h.launches
    t_start = time.perf_counter()
    e, out = tn.dmrg(H, psi, tn.Sweeps(a.sweeps, **kw), observer=obs)
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t_start
    # per-bond durations from consecutive callbacks (the first one includes the environment build)
    per = []
    prev = t_start
    for (sw, b, o, now, en, err) in marks:
        per.append({"sweep": sw, "bond": b, "dir": o, "seconds": now - prev, "energy": en, "truncerr": err})
        prev = now
    first = per[0]["seconds"]
    typical_first = per[1]["seconds"] if len(per) > 1 else 0.0
    env_build = max(0.0, first - typical_first)
    sweeps = []
    for sw in range(a.sweeps):
        s = sum(p["seconds"] for p in per if p["sweep"] == sw)
        sweeps.append(s - (env_build if sw == 0 else 0.0))
    mid = [p for p in per if p["bond"] in (a.N // 2 - 1, a.N // 2) and p["sweep"] == a.sweeps - 1]
    res = {
        "config": "C3: S=1/2 Heisenberg chain N=%d, maxdim %d, Float64, branch %s (%s)" % (a.N, a.chi, a.branch, kw),
        "bond_dims_max": max(D), "n_bond_steps_per_sweep": 2 * (a.N - 1),
        "input_generation_seconds": t_gen,
        "right_environment_build_seconds": env_build,
        "sweep_seconds": sweeps,
        "total_dmrg_call_seconds": t_total,
        "central_bond_step_seconds": [p["seconds"] for p in mid],
        "energy": e, "gpu_launches": h.launches - l0,
        "max_memory_allocated_GB": torch.cuda.max_memory_allocated() / 1e9,
        "workspace_GB": h.workspace_bytes / 1e9,
        "per_bond": per if a.N <= 40 else per[:: max(1, len(per) // 60)],
    }
    print(json.dumps({k: v for k, v in res.items() if k != "per_bond"}, indent=1), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
