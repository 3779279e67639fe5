"""Time tnb_svd_trunc (tier-1 svd) vs cuSOLVER gesvd/gesvdj (torch) on the same box.  usage: svd_probe.py [n ...]"""
import sys, time
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
for n in [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096]:
    for cplx in (False, True):
        dt = torch.complex128 if cplx else torch.float64
        g = torch.Generator(device="cuda").manual_seed(3)
        A = torch.randn(n, n, dtype=dt, device="cuda", generator=g)
        Ad = tn.DTensor(A.T.contiguous().reshape(-1), (n, n))
        tn.ops.svd(tn.DTensor(Ad.data[: 64 * 64].clone(), (64, 64)))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        U, S, V, _ = tn.ops.svd(Ad)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        Um = U.data.view(n, n).T; Vm = V.data.view(n, n).T
        rec = ((Um * S.to(dt)[None, :]) @ Vm.T - A).norm().item() / A.norm().item()
        orth = (Um.conj().T @ Um - torch.eye(n, dtype=dt, device="cuda")).norm().item()
        Sref = torch.linalg.svdvals(A)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        torch.linalg.svd(A)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f"n={n} cplx={cplx}: tnb svd {1e3*(t1-t0):.0f} ms (rec {rec:.1e}, orth {orth:.1e}, S err {((S-Sref).abs().max()/Sref[0]).item():.1e}) | torch.linalg.svd (cuSOLVER) {1e3*(t3-t2):.0f} ms", flush=True)
