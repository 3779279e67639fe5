"""Run the tridiagonalisation alone (for ncu launch lists / timing).  usage: tridiag_probe.py n [reps]"""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = tn.handle(); lib = h.lib
p = lambda t: C.c_void_p(t.data_ptr())
g = torch.Generator(device="cuda").manual_seed(1)
M = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
rho = (M + M.T).contiguous().reshape(-1)
d = torch.zeros(n, dtype=torch.float64, device="cuda"); e = torch.zeros_like(d); tau = torch.zeros_like(d)
for r in range(reps + 1):
    A = rho.clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    h.check(lib.tnb_dbg_tridiag(h.h, 0, C.c_int64(n), p(A), p(d), p(e), p(tau), None))
    b.record(); torch.cuda.synchronize()
    print(f"n={n} rep {r}: tridiag {a.elapsed_time(b):.1f} ms", flush=True)
