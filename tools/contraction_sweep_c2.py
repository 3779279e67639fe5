"""Config C2 (metric M4): random dense ITensor contraction sweep through tnb_contract, chi = 256..8192,
Float64 and ComplexF64, the four shapes of SURVEY.md section 8(d):
  (i)   A[chi,d,chi] * B[chi,w,chi] over the first mode          (rank-3 x rank-3)
  (ii)  phi[chi,d,d,chi] * L[chi,chi,w]            H_eff step 1  (rank-4 x rank-3, compute-bound)
  (iii) T[d,d,chi,chi,w] * W[w,d,d,w]              H_eff step 2  (rank-5 x rank-4, K = w d = 10, HBM-bound)
  (iv)  T[chi,chi,d,d,w] * R[chi,chi,w]            H_eff step 4  (rank-5 x rank-3, compute-bound)
CUDA events, 2 warm-ups, best and median of 5.  TFLOP/s counts 2MNK (x4 for complex); GB/s counts the
algorithmic bytes (MK + KN + MN) * sizeof."""
import json
import sys

import torch

sys.path.insert(0, ".")
from itensorsgpu_b200 import tn  # noqa: E402

D, W = 2, 5


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def rnd(dims, cplx):
    n = 1
    for x in dims:
        n *= x
    t = torch.randn(n, device="cuda", dtype=torch.float64)
    if cplx:
        t = torch.complex(t, torch.randn(n, device="cuda", dtype=torch.float64))
    return tn.DTensor(t, dims)


def main():
    chis = [int(x) for x in sys.argv[1:]] or [256, 512, 1024, 2048, 4096, 8192]
    rows = []
    for cplx in (False, True):
        es = 16 if cplx else 8
        for chi in chis:
            shapes = {
                "i_rank3xrank3": ((chi, D, chi), ("k", "s", "m"), (chi, W, chi), ("k", "a", "n"), D * chi, W * chi, chi),
                "ii_heff_step1": ((chi, D, D, chi), ("l", "s1", "s2", "r"), (chi, chi, W), ("l", "lp", "a"), D * D * chi, chi * W, chi),
                "iii_heff_step2_smallK": ((D, D, chi, chi, W), ("s1", "s2", "r", "lp", "a"), (W, D, D, W), ("a", "s1", "s1p", "b"),
                                          D * chi * chi, D * W, W * D),
                "iv_heff_step4": ((chi, chi, D, D, W), ("r", "lp", "s1p", "s2p", "c"), (chi, chi, W), ("r", "rp", "c"), chi * D * D, chi, chi * W),
            }
            for name, (da, la, db, lb, M, N, K) in shapes.items():
                if cplx and chi >= 8192 and name == "iii_heff_step2_smallK":
                    pass
                try:
                    A, B = rnd(da, cplx), rnd(db, cplx)
                    out, lc = tn.ops.contract(A, la, B, lb)
                    best, med = timeit(lambda: tn.ops.contract(A, la, B, lb, out=out))
                except Exception as ex:  # e.g. out of memory at the largest complex case
                    rows.append(dict(shape=name, chi=chi, dtype="c128" if cplx else "f64", error=str(ex)[:80]))
                    print(json.dumps(rows[-1]), flush=True)
                    torch.cuda.empty_cache()
                    continue
                fl = 2.0 * M * N * K * (4 if cplx else 1)
                by = (M * K + K * N + M * N) * es
                rows.append(dict(shape=name, chi=chi, dtype="c128" if cplx else "f64", M=M, N=N, K=K, ms_best=best * 1e3,
                                 ms_median=med * 1e3, tflops=fl / best * 1e-12, gbs_algorithmic=by / best * 1e-9))
                print(json.dumps(rows[-1]), flush=True)
                del A, B, out
                torch.cuda.empty_cache()
    json.dump(rows, open("gpurun_out/contraction_sweep_c2.json", "w"), indent=1)


if __name__ == "__main__":
    main()
