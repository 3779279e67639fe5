"""Phase timing of one chi x chi central DMRG bond step (C3 shapes) through the public ops.  usage: bond_probe.py [chi]"""
import sys, time
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D, W = 2, 5
g = torch.Generator(device="cuda").manual_seed(1)
r = lambda *dims: tn.DTensor(torch.randn(int(torch.tensor(dims).prod()), device="cuda", dtype=torch.float64, generator=g), dims)
L, R = r(chi, chi, W), r(chi, chi, W)
# symmetrise environments in (l,l') so that H_eff is at least well-behaved
for E in (L, R):
    v = E.data.view(W, chi, chi); v.copy_(0.5 * (v + v.transpose(1, 2)))
W1, W2 = r(W, D, D, W), r(W, D, D, W)
A1 = r(chi, D, chi); A2 = r(chi, D, chi)
A1.data.mul_(1.0 / A1.data.norm()); A2.data.mul_(1.0 / A2.data.norm() * 10)

def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out

ms, (phi, _) = t(lambda: tn.ops.contract(A1, ("l", "s1", "k"), A2, ("k", "s2", "r")))
print(f"phi = A1*A2            {ms:8.1f} ms")
ms, out = t(lambda: tn.ops.heff_apply(L, W1, W2, R, phi))
print(f"H_eff*phi              {ms:8.1f} ms")
ms, _ = t(lambda: tn.ops.eigsolve_lanczos(L, W1, W2, R, phi.clone()))
print(f"Lanczos (3 matvecs)    {ms:8.1f} ms")
for ortho in ("left", "right"):
    ms, (A, B, err) = t(lambda: tn.ops.factorize_bond(phi, ortho=ortho, maxdim=chi, cutoff=0.0, normalize=True), reps=1)
    print(f"factorize svd-route {ortho:5s} {ms:8.1f} ms")
ms, (A, B, err) = t(lambda: tn.ops.factorize_bond(phi, ortho="left", which_decomp="eigen", maxdim=chi, cutoff=1e-11, normalize=True), reps=1)
print(f"factorize eigen         {ms:8.1f} ms  (k={A.dims[2]})")
ms, _ = t(lambda: tn.ops.env_update_left(L, A, W1))
print(f"env_update_left        {ms:8.1f} ms")
ms, _ = t(lambda: tn.ops.env_update_right(R, B, W2))
print(f"env_update_right       {ms:8.1f} ms")
ms, _ = t(lambda: tn.ops.dmrg_bond_step(L, W1, W2, R, A1, A2, "left", maxdim=chi, cutoff=0.0), reps=1)
print(f"dmrg_bond_step left    {ms:8.1f} ms")
M = tn.DTensor(phi.data, (chi * D, D * chi))
ms, _ = t(lambda: tn.ops.contract(M, ("m", "k"), M, ("n", "k")), reps=2)
print(f"Gram M M^T (NT)        {ms:8.1f} ms")
ms, _ = t(lambda: tn.ops.eigh(tn.DTensor(torch.randn(4 * chi * chi, device='cuda', dtype=torch.float64), (2 * chi, 2 * chi)), maxdim=chi), reps=1)
print(f"eigh(2chi) top-chi     {ms:8.1f} ms")
