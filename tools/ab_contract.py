"""A/B timing of the two big H_eff GEMMs (steps 1 and 4) and two plain GEMM forms.  usage: ab_contract.py [chi]"""
import json, os, sys
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D, W = 2, 5
g = torch.Generator(device="cuda").manual_seed(1)
r = lambda *d: tn.DTensor(torch.randn(int(torch.tensor(d).prod()), device="cuda", dtype=torch.float64, generator=g), d)
phi, L, R = r(chi, D, D, chi), r(chi, chi, W), r(chi, chi, W)
T1 = tn.DTensor.empty((D, D, chi, chi, W)); T3 = tn.DTensor(T1.data, (chi, chi, D, D, W)); o4 = tn.DTensor.empty((chi, D, D, chi))
def ev(fn, reps=8):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
F = 2.0 * (D * D * chi) * (chi * W) * chi
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("TNB_")}}
ms = ev(lambda: tn.ops.contract(phi, ("l", "s1", "s2", "r"), L, ("l", "lp", "a"), out=T1)); res["step1_ms"] = ms; res["step1_tflops"] = F / ms * 1e-9
ms = ev(lambda: tn.ops.contract(T3, ("r", "lp", "s1p", "s2p", "c"), R, ("r", "rp", "c"), out=o4)); res["step4_ms"] = ms; res["step4_tflops"] = F / ms * 1e-9
n = 8192
X, Y = r(n, n), r(n, n); Z = tn.DTensor.empty((n, n))
ms = ev(lambda: tn.ops.contract(X, ("k", "m"), Y, ("k", "n"), out=Z), 4); res["gemm_tn_8192_tflops"] = 2.0 * n ** 3 / ms * 1e-9
res["families"] = tn.handle().kernel_family_counts()
print(json.dumps(res))
