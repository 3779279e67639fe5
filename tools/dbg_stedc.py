import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from itensorsgpu_b200 import tn
h = tn.handle(); lib = h.lib
p = lambda t: C.c_void_p(t.data_ptr())
def stedc(d, e):
    n = len(d)
    dd = torch.from_numpy(np.array(d, dtype=np.float64)).cuda()
    ee = torch.zeros(n, dtype=torch.float64, device="cuda"); ee[: n - 1] = torch.from_numpy(np.array(e, dtype=np.float64)).cuda()
    lam = torch.zeros(n, dtype=torch.float64, device="cuda"); Z = torch.zeros(n * n, dtype=torch.float64, device="cuda")
    h.check(lib.tnb_dbg_stedc(h.h, C.c_int64(n), p(dd), p(ee), p(lam), p(Z), None))
    return lam.cpu().numpy(), Z.cpu().numpy().reshape((n, n), order="F")
for kind in ["random", "wilkinson", "graded", "constant"]:
    for n in [65, 96, 128, 129, 256, 300, 1000]:
        rng = np.random.default_rng(7 * n + len(kind))
        if kind == "random": d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
        elif kind == "wilkinson": d, e = np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)
        elif kind == "graded":
            d = 10.0 ** (-np.arange(n) / 6.0); e = 0.3 * np.sqrt(d[:-1] * d[1:])
        else: d, e = np.full(n, 2.0), np.full(n - 1, -1.0)
        lam, Z = stedc(d, e)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        w = np.linalg.eigvalsh(T)[::-1]
        nrm = np.max(np.abs(w))
        print(f"{kind:10s} n={n:5d} eig {np.max(np.abs(lam - w))/nrm:.2e} orth {np.linalg.norm(Z.T @ Z - np.eye(n)):.2e} res {np.linalg.norm(T @ Z - Z * lam[None, :])/nrm:.2e} sorted {bool(np.all(np.diff(lam) <= 0))} nan {int(np.isnan(Z).sum())}", flush=True)
