"""GPU probe: cuBLAS DGEMM (torch) vs tnb_contract on the H_eff GEMM shapes, H_eff*phi at
several chi, and vector-op bandwidth.  Prints JSON lines.  Not a bench (bench.py is)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from itensorsgpu_b200 import tn  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), sorted(ts)[len(ts) // 2]


def rnd(n, cplx=False):
    t = torch.randn(n, device="cuda", dtype=torch.float64)
    if cplx:
        t = torch.complex(t, torch.randn(n, device="cuda", dtype=torch.float64))
    return t


def main():
    out = []
    which = sys.argv[1:] or ["gemm", "heff", "vec", "cplx"]
    if "gemm" in which:
        for (m, n, k) in [(8192, 8192, 8192), (16384, 20480, 4096), (16384, 4096, 20480), (4096, 5120, 1024)]:
            A = rnd(m * k).view(m, k)
            B = rnd(k * n).view(k, n)
            best, med = timeit(lambda: torch.matmul(A, B))
            out.append(dict(what="cublas_dgemm_NN_rowmajor", m=m, n=n, k=k, tflops=2.0 * m * n * k / best * 1e-12,
                            tflops_median=2.0 * m * n * k / med * 1e-12))
            print(json.dumps(out[-1]), flush=True)
            # same flops through tnb_contract, TN layout (both operands K-major, C M-major)
            At = tn.DTensor(A.reshape(-1), (k, m))
            Bt = tn.DTensor(B.reshape(-1), (k, n))
            Ct = tn.DTensor.empty((m, n))
            best, med = timeit(lambda: tn.ops.contract(At, ("k", "m"), Bt, ("k", "n"), out=Ct))
            out.append(dict(what="tnb_contract_TN", m=m, n=n, k=k, tflops=2.0 * m * n * k / best * 1e-12,
                            tflops_median=2.0 * m * n * k / med * 1e-12))
            print(json.dumps(out[-1]), flush=True)
            At = tn.DTensor(A.reshape(-1), (m, k))
            Bt = tn.DTensor(B.reshape(-1), (n, k))
            best, med = timeit(lambda: tn.ops.contract(At, ("m", "k"), Bt, ("n", "k"), out=Ct))
            out.append(dict(what="tnb_contract_NT", m=m, n=n, k=k, tflops=2.0 * m * n * k / best * 1e-12,
                            tflops_median=2.0 * m * n * k / med * 1e-12))
            print(json.dumps(out[-1]), flush=True)
            del A, B, At, Bt, Ct
    if "heff" in which:
        for chi in (512, 1024, 2048, 4096):
            d, w = 2, 5
            L = tn.DTensor(rnd(chi * chi * w), (chi, chi, w))
            R = tn.DTensor(rnd(chi * chi * w), (chi, chi, w))
            W1 = tn.DTensor(rnd(w * d * d * w), (w, d, d, w))
            W2 = tn.DTensor(rnd(w * d * d * w), (w, d, d, w))
            phi = tn.DTensor(rnd(chi * d * d * chi), (chi, d, d, chi))
            o = tn.DTensor.empty(phi.dims)
            best, med = timeit(lambda: tn.ops.heff_apply(L, W1, W2, R, phi, out=o), reps=4)
            F = 2 * d * d * w * (2 * chi ** 3) + 4 * d ** 3 * w * w * chi * chi
            out.append(dict(what="heff_apply", chi=chi, ms=best * 1e3, tflops=F / best * 1e-12, tflops_median=F / med * 1e-12))
            print(json.dumps(out[-1]), flush=True)
            # per-step breakdown via the primitive tier
            T1 = tn.DTensor.empty((d, d, chi, chi, w))
            T2 = tn.DTensor.empty((d, chi, chi, d, w))
            b1, _ = timeit(lambda: tn.ops.contract(phi, ("l", "s1", "s2", "r"), L, ("l", "lp", "a"), out=T1), reps=3)
            b2, _ = timeit(lambda: tn.ops.contract(T1, ("s1", "s2", "r", "lp", "a"), W1, ("a", "s1", "s1p", "b"), out=T2), reps=3)
            out.append(dict(what="heff_steps", chi=chi, step1_ms=b1 * 1e3, step2_ms=b2 * 1e3,
                            step1_tflops=2.0 * (d * d * chi) * (chi * w) * chi / b1 * 1e-12,
                            step2_gbs=2 * 8.0 * d * d * chi * chi * w / b2 * 1e-9))
            print(json.dumps(out[-1]), flush=True)
            del L, R, phi, o, T1, T2
    if "cplx" in which:
        for (m, n, k) in [(4096, 4096, 4096), (8192, 10240, 2048)]:
            A = tn.DTensor(rnd(m * k, True), (k, m))
            B = tn.DTensor(rnd(k * n, True), (k, n))
            Cc = tn.DTensor.empty((m, n), torch.complex128)
            best, med = timeit(lambda: tn.ops.contract(A, ("k", "m"), B, ("k", "n"), out=Cc), reps=3)
            out.append(dict(what="tnb_contract_TN_c128", m=m, n=n, k=k, tflops=8.0 * m * n * k / best * 1e-12))
            print(json.dumps(out[-1]), flush=True)
            Am = A.data.view(m, k)
            Bm = B.data.view(n, k)
            best, med = timeit(lambda: torch.matmul(Am, Bm.t()), reps=3)
            out.append(dict(what="cublas_zgemm", m=m, n=n, k=k, tflops=8.0 * m * n * k / best * 1e-12))
            print(json.dumps(out[-1]), flush=True)
            del A, B, Cc
    if "factor" in which:
        sizes = [int(x) for x in (sys.argv[sys.argv.index("--sizes") + 1].split(",") if "--sizes" in sys.argv else ["1024", "2048", "4096"])]
        for n in sizes:
            A = rnd(n * n).view(n, n)
            M = tn.DTensor(A.reshape(-1).clone(), (n, n))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            U, S, V, err = tn.ops.svd(M, maxdim=n // 2)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            out.append(dict(what="tnb_svd_trunc(maxdim n/2)", n=n, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
            t0 = time.perf_counter(); torch.linalg.svd(A, full_matrices=False); torch.cuda.synchronize(); t1 = time.perf_counter()
            out.append(dict(what="cusolver_svd(torch.linalg.svd)", n=n, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
            X = rnd(n * (n // 2)).view(n, n // 2)
            rho = (X @ X.t()).contiguous()
            Mr = tn.DTensor(rho.reshape(-1).clone(), (n, n))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            D, Ue, err = tn.ops.eigh(Mr, maxdim=n // 2)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            out.append(dict(what="tnb_eigh_trunc(psd rank n/2, maxdim n/2)", n=n, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
            t0 = time.perf_counter(); torch.linalg.eigh(rho); torch.cuda.synchronize(); t1 = time.perf_counter()
            out.append(dict(what="cusolver_syevd(torch.linalg.eigh)", n=n, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
            Mq = tn.DTensor(A.reshape(-1)[: n * (n // 2)].clone(), (n, n // 2))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            tn.ops.qr(Mq)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            out.append(dict(what="tnb_qr(n x n/2)", n=n, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
            del A, M, X, rho, Mr, Mq
    if "bond" in which:
        sizes = [int(x) for x in (sys.argv[sys.argv.index("--sizes") + 1].split(",") if "--sizes" in sys.argv else ["512", "1024"])]
        for chi in sizes:
            d, w = 2, 5
            L = tn.DTensor(rnd(chi * chi * w), (chi, chi, w)); R = tn.DTensor(rnd(chi * chi * w), (chi, chi, w))
            W1 = tn.DTensor(rnd(w * d * d * w), (w, d, d, w)); W2 = tn.DTensor(rnd(w * d * d * w), (w, d, d, w))
            A1 = tn.DTensor(rnd(chi * d * chi) / chi, (chi, d, chi)); A2 = tn.DTensor(rnd(chi * d * chi) / chi, (chi, d, chi))
            for noise, cutoff, name in ((0.0, 0.0, "svd"), (1e-10, 1e-11, "eigen+noise")):
                tn.ops.dmrg_bond_step(L, W1, W2, R, A1, A2, "left", maxdim=chi, cutoff=cutoff, noise=noise)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                tn.ops.dmrg_bond_step(L, W1, W2, R, A1, A2, "left", maxdim=chi, cutoff=cutoff, noise=noise)
                torch.cuda.synchronize(); t1 = time.perf_counter()
                out.append(dict(what="dmrg_bond_step", branch=name, chi=chi, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
            t0 = time.perf_counter(); tn.ops.env_update_left(L, A1, W1); torch.cuda.synchronize(); t1 = time.perf_counter()
            out.append(dict(what="env_update_left", chi=chi, s=t1 - t0)); print(json.dumps(out[-1]), flush=True)
    if "vec" in which:
        n = 1 << 26
        x = tn.DTensor(rnd(n), (n,))
        y = tn.DTensor(rnd(n), (n,))
        b, _ = timeit(lambda: tn.ops.permute_axpby(x, ("n",), y, ("n",), 0.5, 1.0))
        out.append(dict(what="axpby", n=n, gbs=3 * 8.0 * n / b * 1e-9))
        b, _ = timeit(lambda: tn.ops.dot(x, y))
        out.append(dict(what="dot(+sync)", n=n, gbs=2 * 8.0 * n / b * 1e-9))
        b, _ = timeit(lambda: tn.ops.norm(x))
        out.append(dict(what="nrm2(+sync)", n=n, gbs=8.0 * n / b * 1e-9))
        X = tn.DTensor(x.data, (8192, 8192))
        Y = tn.DTensor(y.data, (8192, 8192))
        b, _ = timeit(lambda: tn.ops.permute_axpby(X, ("a", "b"), Y, ("b", "a"), 1.0, 0.0))
        out.append(dict(what="transpose8192", gbs=2 * 8.0 * n / b * 1e-9))
        b, _ = timeit(lambda: y.data.copy_(x.data))
        out.append(dict(what="torch_copy", gbs=2 * 8.0 * n / b * 1e-9))
        for o in out[-5:]:
            print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
