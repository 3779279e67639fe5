"""One H_eff*phi at chi (default 4096) for ncu captures.  usage: heff_probe.py [chi] [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
D, W = 2, 5
g = torch.Generator(device="cuda").manual_seed(1)
r = lambda n: torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
L = tn.DTensor(r(chi * chi * W), (chi, chi, W)); R = tn.DTensor(r(chi * chi * W), (chi, chi, W))
W1 = tn.DTensor(r(W * D * D * W), (W, D, D, W)); W2 = tn.DTensor(r(W * D * D * W), (W, D, D, W))
phi = tn.DTensor(r(chi * D * D * chi), (chi, D, D, chi)); out = tn.DTensor.empty(phi.dims)
for _ in range(reps):
    tn.ops.heff_apply(L, W1, W2, R, phi, out=out)
torch.cuda.synchronize()
