"""A/B of the tail-wave split-K on the step-4 GEMM of an 8-way shard (per-rank shape: M = 2048, N = 4096, K = 20480 ->
1024 tiles of 64x128 on 296 CTA slots = 3.46 waves).  Run once with TNB_SPLITK=off and once without."""
import json, os, sys
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
chi, clp, D, W = 4096, 512, 2, 5
g = torch.Generator(device="cuda").manual_seed(1)
r = lambda *d: tn.DTensor(torch.randn(int(torch.tensor(d).prod()), device="cuda", dtype=torch.float64, generator=g), d)
T3, R = r(chi, clp, D, D, W), r(chi, chi, W)
out = tn.DTensor.empty((clp, D, D, chi))
fn = lambda: tn.ops.contract(T3, ("r", "lp", "s1p", "s2p", "c"), R, ("r", "rp", "c"), out=out)
fn(); fn(); torch.cuda.synchronize()
h = tn.handle(); l0 = h.launches
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): fn()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
F = 2.0 * (clp * D * D) * chi * (chi * W)
print(json.dumps({"TNB_SPLITK": os.environ.get("TNB_SPLITK", "on"), "ms": ms, "tflops": F / ms * 1e-9, "launches_per_call": (h.launches - l0) / 20,
                  "families": h.kernel_family_counts()}))
