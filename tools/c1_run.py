"""Config C1 (BASELINE.json configs[0], the reference's own CPU-runnable case): S=1 Heisenberg chain, N=100,
5 sweeps with maxdim 10/20/100/100/200, cutoff 1e-11 (ITensors.jl's stock examples/dmrg.jl schedule), run through
the library's dmrg() on the GPU and through the CPU oracle on the host cores of the same box, same start state.
Reports per-sweep energies of both, wall seconds of both, and the literature value for orientation.
usage: c1_run.py [--N 100] [--skip-oracle]"""
import argparse, json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
from oracle import dmrg as od, models, mps as omps

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=100)
ap.add_argument("--skip-oracle", action="store_true")
ap.add_argument("--out", default=None)
a = ap.parse_args()
N = a.N
Ws = models.heisenberg_mpo(N, 1.0)
psi0 = omps.random_mps(N, 3, 10, np.random.default_rng(2024))
kw = dict(maxdim=[10, 20, 100, 100, 200], cutoff=1e-11)
H = tn.cu(tn.MPO([w.copy() for w in Ws])); p0 = tn.cu(tn.MPS([t.copy() for t in psi0], llim=-1, rlim=1))
hist = []
tn.dmrg(H, p0, tn.Sweeps(1, maxdim=[10], cutoff=1e-11))          # warm-up (kernel attributes, arena)
torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = tn.handle().launches
marks = []
e, psi = tn.dmrg(H, p0, tn.Sweeps(5, **kw), observer=lambda sw, b, o, en, err: (hist.append(en), marks.append(time.perf_counter())) if (o == "right" and b == 0) else None)
torch.cuda.synchronize(); tg = time.perf_counter() - t0
res = {"config": "C1: S=1 Heisenberg N=%d, 5 sweeps maxdim 10/20/100/100/200, cutoff 1e-11" % N, "gpu_energy_per_sweep": hist,
       "gpu_seconds_total": tg, "gpu_seconds_per_sweep": list(np.diff([t0] + marks)), "gpu_launches": tn.handle().launches - l0,
       "maxlinkdim": psi.maxlinkdim(), "literature_E0_N100_S1_OBC": -138.940086}
if not a.skip_oracle:
    t1 = time.perf_counter()
    e_ref, _, hist_ref = od.dmrg(Ws, psi0, od.Sweeps(5, **kw))
    res["cpu_oracle_seconds_total"] = time.perf_counter() - t1
    res["cpu_oracle_energy_per_sweep"] = [float(x) for x in hist_ref]
    res["abs_energy_difference_per_sweep"] = [abs(x - y) for x, y in zip(hist, hist_ref)]
print(json.dumps(res, indent=1))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
