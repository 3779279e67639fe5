"""Time the fast Hermitian eigensolver stage by stage on the GPU (CUDA events), and compare to
torch.linalg.eigh (cuSOLVER syevd) on the same box.  usage: python tools/eigh_probe.py [n ...]"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from itensorsgpu_b200 import tn  # noqa: E402


def ev_time(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ns = [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096, 8192]
    h = tn.handle()
    lib = h.lib
    p = lambda t: C.c_void_p(t.data_ptr())
    for n in ns:
        for cplx in (False, True):
            if cplx and n > 4096:
                continue
            g = torch.Generator(device="cuda").manual_seed(1)
            dt = torch.complex128 if cplx else torch.float64
            # DMRG-like density matrix: M with decaying spectrum, rho = M M^H
            M = torch.randn(n, n, dtype=dt, device="cuda", generator=g)
            s = torch.pow(10.0, -torch.arange(n, device="cuda", dtype=torch.float64) / (n / 20.0))
            M = M * s[None, :].to(dt)
            rho = M @ M.conj().T
            rho = (rho + rho.conj().T) / 2
            A0 = rho.T.contiguous().reshape(-1)        # column-major flat
            d = torch.zeros(n, dtype=torch.float64, device="cuda")
            e = torch.zeros(n, dtype=torch.float64, device="cuda")
            tau = torch.zeros(n, dtype=dt, device="cuda")
            lam = torch.zeros(n, dtype=torch.float64, device="cuda")
            Z = torch.zeros(n * n, dtype=torch.float64, device="cuda")
            A = A0.clone()
            l0 = h.launches
            t_tri = ev_time(lambda: h.check(lib.tnb_dbg_tridiag(h.h, int(cplx), C.c_int64(n), p(A), p(d), p(e), p(tau), None)))
            l1 = h.launches
            d2, e2 = d.clone(), e.clone()
            t_dc = ev_time(lambda: h.check(lib.tnb_dbg_stedc(h.h, C.c_int64(n), p(d2), p(e2), p(lam), p(Z), None)))
            l2 = h.launches
            X = torch.zeros(n * (n // 2), dtype=dt, device="cuda")
            X[:: n + 1] = 1
            t_bt = ev_time(lambda: h.check(lib.tnb_dbg_backtransform(h.h, int(cplx), C.c_int64(n), p(A), p(tau), p(X), C.c_int64(n // 2), None)))
            l3 = h.launches
            # end to end, full and truncated to n/2
            Dt = tn.DTensor(A0.clone(), (n, n))
            t_full = ev_time(lambda: tn.ops.eigh(Dt))
            t_half = ev_time(lambda: tn.ops.eigh(Dt, maxdim=n // 2))
            Dv, U, _ = tn.ops.eigh(Dt, maxdim=n // 2)
            Un = U.data.reshape(n // 2, n).T      # logical (n, n/2)
            res = (rho @ Un - Un * Dv[None, :].to(dt)).norm().item() / rho.norm().item()
            orth = (Un.conj().T @ Un - torch.eye(n // 2, dtype=dt, device="cuda")).norm().item()
            t_ref = ev_time(lambda: torch.linalg.eigh(rho))
            print(f"n={n} cplx={cplx}: tridiag {t_tri:.1f} ms ({l1-l0} launches)  stedc {t_dc:.1f} ms ({l2-l1})  backtransf(n/2) {t_bt:.1f} ms ({l3-l2})"
                  f"  | eigh full {t_full:.1f} ms, top-n/2 {t_half:.1f} ms | cuSOLVER syevd {t_ref:.1f} ms | res {res:.1e} orth {orth:.1e}", flush=True)
            del M, rho, A0, A, Z, X, U


if __name__ == "__main__":
    main()
