"""tnb_contract on the four staging forms of a plain GEMM (which operand is K-contiguous), CUDA events."""
import sys
import torch
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
m, n, k = 8192, 8192, 4096
def ev(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
r = lambda *d: tn.DTensor(torch.randn(int(torch.tensor(d).prod()), device="cuda", dtype=torch.float64), d)
C = tn.DTensor.empty((m, n))
for name, da, la, db, lb in [("A[m,k] B[k,n]  (A free-major, B K-major: phi = A1*A2)", (m, k), ("m", "k"), (k, n), ("k", "n")),
                             ("A[k,m] B[k,n]  (both K-major: H_eff steps 1,4)", (k, m), ("k", "m"), (k, n), ("k", "n")),
                             ("A[m,k] B[n,k]  (both free-major: Gram, rank-k updates)", (m, k), ("m", "k"), (n, k), ("n", "k")),
                             ("A[k,m] B[n,k]  (A K-major, B free-major)", (k, m), ("k", "m"), (n, k), ("n", "k"))]:
    A, B = r(*da), r(*db)
    ms = ev(lambda: tn.ops.contract(A, la, B, lb, out=C))
    print(f"{name:62s} {ms:7.2f} ms  {2.0*m*n*k/ms*1e-9:6.2f} TFLOP/s", flush=True)
