"""Config C5 shapes (2D Heisenberg cylinder, dense MPO bond w ~ 30): single-bond H_eff*phi and factorize timings.
chi = 8192 needs 2 x 64 GB of temporaries + 2 x 16 GB of environments (> one GPU without chunking over the MPO
bond), so the sweep stops at chi = 6144 (92 GB).  usage: c5_probe.py [chi ...]"""
import json, sys
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
D, W = 2, 30
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
rows = []
for chi in [int(x) for x in sys.argv[1:]] or [2048, 4096, 6144]:
    g = torch.Generator(device="cuda").manual_seed(5)
    r = lambda *d: tn.DTensor(torch.randn(int(torch.tensor(d).prod()), device="cuda", dtype=torch.float64, generator=g), d)
    L, R = r(chi, chi, W), r(chi, chi, W)
    W1, W2 = r(W, D, D, W), r(W, D, D, W)
    phi = r(chi, D, D, chi); out = tn.DTensor.empty(phi.dims)
    ms = ev(lambda: tn.ops.heff_apply(L, W1, W2, R, phi, out=out))
    F = 2.0 * D * D * W * 2 * chi ** 3 + 4.0 * D ** 3 * W * W * chi * chi
    phi.data.mul_(1.0 / phi.data.norm())
    fms = ev(lambda: tn.ops.factorize_bond(phi, ortho="left", which_decomp="eigen", maxdim=chi, cutoff=1e-11), reps=1)
    rows.append(dict(chi=chi, w=W, d=D, heff_ms=ms, heff_tflops=F / ms * 1e-9, flop=F, factorize_eigen_ms=fms,
                     workspace_GB=tn.handle().workspace_bytes / 1e9))
    print(json.dumps(rows[-1]), flush=True)
    del L, R, phi, out
    torch.cuda.empty_cache()
json.dump(rows, open("gpurun_out/c5_probe.json", "w"), indent=1)
