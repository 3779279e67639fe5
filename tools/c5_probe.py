"""Config C5 (2D Heisenberg cylinder 6x24, dense MPO bond w ~ 30, maxdim 8192): the central bond at chi = 8192 on ONE
GPU in a fixed workspace (tnb_set_workspace_limit: H_eff cut into slabs of the output bond, environment update into
chunks of the summed bond).  Reports H_eff*phi TFLOP/s, the full bond step (phi = A1 A2, Lanczos with 3 matvecs,
truncated factorization: svd rule and the examples' eigen + noise rule), one environment update, the workspace arena
and the peak HBM in use.  The result is checked against the oracle on sampled output elements (the environments
sliced to a grid of (l', r'); tests/gpu_util.py).
usage: c5_probe.py [chi] [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from itensorsgpu_b200 import tn  # noqa: E402

D, W = 2, 30
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/r02_c5_chi%d.json" % chi
h = tn.handle()
g = torch.Generator(device="cuda").manual_seed(5)


def r(*d):
    return tn.DTensor(torch.randn(int(np.prod(d)), device="cuda", dtype=torch.float64, generator=g), d)


def herm(E, n, w):           # make an environment Hermitian in its first two modes (in place, slice by slice)
    v = E.data.view(w, n, n)
    for a in range(w):
        v[a].add_(v[a].T.clone())


L, R = r(chi, chi, W), r(chi, chi, W)
herm(L, chi, W); herm(R, chi, W)
W1, W2 = r(W, D, D, W), r(W, D, D, W)
for Wt in (W1, W2):
    v = Wt.data.view(W, D, D, W)
    v.add_(v.transpose(1, 2).clone())
phi = r(chi, D, D, chi)
phi.data.mul_(1.0 / phi.data.norm())
out = tn.DTensor.empty(phi.dims)
res = {"config": "C5 central bond: chi=%d, d=%d, dense MPO bond w=%d, Float64, one B200" % (chi, D, W),
       "workspace_limit_GB": h.workspace_limit / 1e9}

# ---- H_eff*phi
tn.ops.heff_apply(L, W1, W2, R, phi, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = h.launches
a.record(); tn.ops.heff_apply(L, W1, W2, R, phi, out=out); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
F = 2.0 * D * D * W * 2 * chi ** 3 + 4.0 * D ** 3 * W * W * chi * chi
res.update(heff_ms=ms, heff_tflops=F / ms * 1e-9, heff_flop=F, heff_launches=h.launches - l0)

# ---- parity on sampled output elements
from gpu_util import pick_indices, restrict  # noqa: E402
from oracle import dmrg as od  # noqa: E402
from oracle import tensor as ot  # noqa: E402
rng = np.random.default_rng(3)
picks = {"lp": pick_indices(rng, chi, 8), "rp": pick_indices(rng, chi, 8)}
t0 = time.perf_counter()
want = od.heff_apply(restrict(L, ("l", "lp", "a"), picks), W1.numpy(), W2.numpy(), restrict(R, ("r", "rp", "c"), picks), phi.numpy())
res["parity"] = {"max_rel_err": ot.rel_err(restrict(out, ("lp", "s1", "s2", "rp"), picks), want), "n_samples": int(want.size),
                 "oracle_seconds": time.perf_counter() - t0}
del out

# ---- environment update
A = r(chi, D, chi)
Wm = r(W, D, D, W)
tn.ops.env_update_left(L, A, Wm)
torch.cuda.synchronize()
t0 = time.perf_counter(); tn.ops.env_update_left(L, A, Wm); torch.cuda.synchronize()
res["env_update_left_s"] = time.perf_counter() - t0
res["env_update_flop"] = 2.0 * (2 * chi ** 3 * D * W) + 2.0 * chi * chi * D * D * W * W
del A, Wm

# ---- full bond step, both factorize rules (A1, A2 = an exact split of a random two-site tensor)
A1 = r(chi, D, chi); A2 = r(chi, D, chi)
A1.data.mul_(1.0 / np.sqrt(chi * D)); A2.data.mul_(1.0 / np.sqrt(chi * D))
for name, kw in (("svd_rule", dict(maxdim=chi, cutoff=0.0, noise=0.0)), ("eigen_noise_rule", dict(maxdim=chi, cutoff=1e-11, noise=1e-10))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e, B1, B2, err = tn.ops.dmrg_bond_step(L, W1, W2, R, A1, A2, "left", **kw)
    torch.cuda.synchronize()
    res["bond_step_%s_s" % name] = time.perf_counter() - t0
    res["bond_step_%s_energy" % name] = e
    res["bond_step_%s_kept" % name] = B1.dims[2]
    del B1, B2
t0 = time.perf_counter()
tn.ops.factorize_bond(phi, ortho="left", which_decomp="eigen", maxdim=chi, cutoff=1e-11)
torch.cuda.synchronize()
res["factorize_eigen_s"] = time.perf_counter() - t0
res["workspace_GB"] = h.workspace_bytes / 1e9
res["max_memory_allocated_GB_torch"] = torch.cuda.max_memory_allocated() / 1e9
free, total = torch.cuda.mem_get_info()
res["hbm_in_use_GB_at_end"] = (total - free) / 1e9
print(json.dumps(res, indent=1), flush=True)
json.dump(res, open(out_path, "w"), indent=1)
