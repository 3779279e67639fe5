"""Functional host layer over the C ABI: device tensors + one Python function per entry point.

A ``DTensor`` is a flat, column-major device buffer plus its extents -- exactly what the
reference's ``CuDense`` storage is (``src/tensor/cudense.jl:1-2``: ``Dense`` whose data is a
``CuVector``).  torch is used only to own device memory and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BondDims, F64, C128


def _dt(t):
    if t.dtype == torch.float64:
        return F64
    if t.dtype == torch.complex128:
        return C128
    raise TypeError("only float64 / complex128 are supported (got %s)" % t.dtype)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _i64(xs):
    return (C.c_int64 * max(len(xs), 1))(*[int(x) for x in xs])


def _i32(xs):
    return (C.c_int32 * max(len(xs), 1))(*[int(x) for x in xs])


def _scalar(v, dtype):
    """host scalar -> ctypes buffer of one element of dtype"""
    if v is None:
        return None, None
    v = complex(v)
    buf = (C.c_double * 2)(v.real, v.imag)
    if dtype == F64 and v.imag != 0:
        raise TypeError("complex scalar with a real tensor")
    return buf, C.cast(buf, C.c_void_p)


class DTensor:
    """Flat column-major device tensor.  ``data``: 1-D torch CUDA tensor; ``dims``: extents."""

    __slots__ = ("data", "dims")

    def __init__(self, data, dims):
        dims = tuple(int(d) for d in dims)
        n = 1
        for d in dims:
            n *= d
        if data.numel() != n:
            raise _lib.DimensionMismatch(2, "buffer of %d elements for dims %s" % (data.numel(), dims))
        self.data = data
        self.dims = dims

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def size(self):
        return self.data.numel()

    @staticmethod
    def from_numpy(a, device="cuda"):
        """H2D of a logical ndarray (``cuITensor(A::Array, inds)``: ``src/cuitensor.jl:18-20``)."""
        a = np.asarray(a)
        if a.dtype not in (np.float64, np.complex128):
            a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
        flat = np.ascontiguousarray(a.ravel(order="F"))
        return DTensor(torch.from_numpy(flat).to(device), a.shape)

    def numpy(self):
        """D2H to a logical ndarray (``cpu``: ``src/cuitensor.jl:30-33``, ``cudense.jl:12``)."""
        return self.data.cpu().numpy().reshape(self.dims, order="F")

    @staticmethod
    def empty(dims, dtype=torch.float64, device="cuda"):
        n = 1
        for d in dims:
            n *= int(d)
        return DTensor(torch.empty(n, dtype=dtype, device=device), dims)

    @staticmethod
    def zeros(dims, dtype=torch.float64, device="cuda"):
        n = 1
        for d in dims:
            n *= int(d)
        return DTensor(torch.zeros(n, dtype=dtype, device=device), dims)

    def astype(self, dtype):
        if self.data.dtype == dtype:
            return self
        if self.data.is_complex() and not dtype.is_complex:
            raise TypeError("astype: refusing to drop the imaginary part of a complex tensor")
        return DTensor(self.data.to(dtype), self.dims)

    def clone(self):
        return DTensor(self.data.clone(), self.dims)

    def reshape(self, dims):
        return DTensor(self.data, dims)


def _promote(A, B):
    if A.dtype == B.dtype:
        return A, B
    return A.astype(torch.complex128), B.astype(torch.complex128)


def output_labels(la, lb):
    la, lb = list(la), list(lb)
    return tuple([l for l in la if l not in lb] + [l for l in lb if l not in la])


def contract(A, la, B, lb, lc=None, out=None, alpha=None, beta=None, conj_a=False, conj_b=False):
    """C[lc] <- alpha * A[la]*B[lb] (+ beta*C).  Default lc = A-free then B-free.
    Labels are arbitrary hashables; they are mapped to small ints for the ABI."""
    h = _lib.handle()
    A, B = _promote(A, B)
    la, lb = list(la), list(lb)
    if lc is None:
        lc = output_labels(la, lb)
    lc = list(lc)
    ids = {}
    for l in la + lb + lc:
        ids.setdefault(l, len(ids))
    ext = {}
    for l, d in list(zip(la, A.dims)) + list(zip(lb, B.dims)):
        ext.setdefault(l, d)
    for l in lc:
        if l not in ext:
            raise _lib.TnbError(1, "output label %r is in neither input" % (l,))
    cdims = tuple(ext[l] for l in lc)
    if out is None:
        out = DTensor.empty(cdims, A.dtype, A.data.device)
        if beta is not None and complex(beta) != 0:
            raise _lib.TnbError(1, "beta != 0 needs an `out` tensor")
    elif out.dtype != A.dtype:
        raise TypeError("out dtype mismatch")
    dt = _dt(A.data)
    _, pa = _scalar(alpha, dt)
    _, pb = _scalar(beta, dt)
    flags = (1 if conj_a else 0) | (2 if conj_b else 0)
    h.check(h.lib.tnb_contract(h.h, dt, len(la), _i64(A.dims), _i32([ids[l] for l in la]), _ptr(A.data),
                               len(lb), _i64(B.dims), _i32([ids[l] for l in lb]), _ptr(B.data),
                               len(lc), _i64(out.dims), _i32([ids[l] for l in lc]), _ptr(out.data),
                               pa, pb, flags, _stream()))
    return out, tuple(lc)


def permute_axpby(A, la, B, lb, alpha=1.0, beta=0.0):
    """B[lb] <- alpha*A[la] + beta*B[lb] in place; returns B."""
    h = _lib.handle()
    if A.dtype != B.dtype:
        raise TypeError("dtype mismatch")
    la, lb = list(la), list(lb)
    if len(la) != len(lb) or set(la) != set(lb):
        raise _lib.TnbError(1, "permute: label sets differ")
    ids = {l: i for i, l in enumerate(la)}
    want = tuple(A.dims[la.index(l)] for l in lb)
    if want != B.dims:
        raise _lib.DimensionMismatch(2, "permute: B has dims %s, expected %s" % (B.dims, want))
    dt = _dt(A.data)
    _, pa = _scalar(alpha, dt)
    _, pb = _scalar(beta, dt)
    h.check(h.lib.tnb_permute_axpby(h.h, dt, len(la), _i64(A.dims), _i32([ids[l] for l in la]), _ptr(A.data),
                                    _i32([ids[l] for l in lb]), _ptr(B.data), pa, pb, _stream()))
    return B


def permute(A, la, lb):
    out = DTensor.empty(tuple(A.dims[list(la).index(l)] for l in lb), A.dtype, A.data.device)
    return permute_axpby(A, la, out, lb)


def scale(A, alpha):
    h = _lib.handle()
    dt = _dt(A.data)
    _, pa = _scalar(alpha, dt)
    h.check(h.lib.tnb_scale(h.h, dt, A.size, _ptr(A.data), pa, _stream()))
    return A


def dot(A, B):
    """<A|B> = sum conj(A)*B, returned to the host (synchronises)."""
    h = _lib.handle()
    if A.dtype != B.dtype or A.size != B.size:
        raise _lib.DimensionMismatch(2, "dot: operands differ")
    dt = _dt(A.data)
    res = (C.c_double * 2)(0.0, 0.0)
    h.check(h.lib.tnb_dot(h.h, dt, A.size, _ptr(A.data), _ptr(B.data), None, C.cast(res, C.c_void_p), _stream()))
    return res[0] if dt == F64 else complex(res[0], res[1])


def norm(A):
    h = _lib.handle()
    res = C.c_double(0.0)
    h.check(h.lib.tnb_nrm2(h.h, _dt(A.data), A.size, _ptr(A.data), None, C.byref(res), _stream()))
    return res.value


def truncate(P, maxdim=None, mindim=1, cutoff=0.0, use_absolute_cutoff=False, use_relative_cutoff=True):
    """P: 1-D float64 CUDA tensor of descending weights -> (truncerr, docut, n_keep)."""
    h = _lib.handle()
    n = C.c_int64(0)
    err = C.c_double(0.0)
    docut = C.c_double(0.0)
    flags = (1 if use_absolute_cutoff else 0) | (0 if use_relative_cutoff else 2)
    h.check(h.lib.tnb_truncate(h.h, _ptr(P), P.numel(), int(maxdim) if maxdim is not None else 0, int(mindim),
                               float(cutoff), flags, C.byref(n), C.byref(err), C.byref(docut), _stream()))
    return err.value, docut.value, n.value


# ------------------------------------------------------------------ tier 2
def bond_dims(L, W1, W2, R, phi):
    cl, d1, d2, cr = phi.dims
    if L.dims[:2] != (cl, cl) or R.dims[:2] != (cr, cr):
        raise _lib.DimensionMismatch(2, "environment / phi bond dims differ")
    return BondDims(cl, cr, d1, d2, W1.dims[0], W1.dims[3], W2.dims[3])


def heff_apply(L, W1, W2, R, phi, out=None):
    h = _lib.handle()
    bd = bond_dims(L, W1, W2, R, phi)
    if out is None:
        out = DTensor.empty(phi.dims, phi.dtype, phi.data.device)
    h.check(h.lib.tnb_heff_apply(h.h, _dt(phi.data), C.byref(bd), _ptr(L.data), _ptr(W1.data), _ptr(W2.data),
                                 _ptr(R.data), _ptr(phi.data), _ptr(out.data), _stream()))
    return out


def heff_apply_shard(Lslab, W1, W2, R, phi, out=None):
    """Output-bond-sharded H_eff*phi: Lslab[l, l'_shard, a] -> out[l'_shard, s1', s2', r']."""
    h = _lib.handle()
    cl, d1, d2, cr = phi.dims
    clp = Lslab.dims[1]
    if Lslab.dims[0] != cl or R.dims[:2] != (cr, cr):
        raise _lib.DimensionMismatch(2, "environment / phi bond dims differ")
    bd = BondDims(cl, cr, d1, d2, W1.dims[0], W1.dims[3], W2.dims[3])
    if out is None:
        out = DTensor.empty((clp, d1, d2, cr), phi.dtype, phi.data.device)
    h.check(h.lib.tnb_heff_apply_shard(h.h, _dt(phi.data), C.byref(bd), clp, _ptr(Lslab.data), _ptr(W1.data),
                                       _ptr(W2.data), _ptr(R.data), _ptr(phi.data), _ptr(out.data), _stream()))
    return out


def heff_apply_host(L, W1, W2, R, phi_host, out_host, dims):
    """phi_host/out_host: pinned CPU torch tensors (flat); L.. on device.  Synchronous."""
    h = _lib.handle()
    cl, d1, d2, cr = dims
    bd = BondDims(cl, cr, d1, d2, W1.dims[0], W1.dims[3], W2.dims[3])
    h.check(h.lib.tnb_heff_apply_host(h.h, _dt(L.data), C.byref(bd), _ptr(L.data), _ptr(W1.data), _ptr(W2.data),
                                      _ptr(R.data), _ptr(phi_host), _ptr(out_host), _stream()))
    return out_host


def env_update_left(L, A, W):
    h = _lib.handle()
    cl, d, cr = A.dims
    out = DTensor.empty((cr, cr, W.dims[3]), A.dtype, A.data.device)
    h.check(h.lib.tnb_env_update_left(h.h, _dt(A.data), cl, cr, d, W.dims[0], W.dims[3], _ptr(L.data),
                                      _ptr(A.data), _ptr(W.data), _ptr(out.data), _stream()))
    return out


def env_update_right(R, A, W):
    h = _lib.handle()
    cl, d, cr = A.dims
    out = DTensor.empty((cl, cl, W.dims[0]), A.dtype, A.data.device)
    h.check(h.lib.tnb_env_update_right(h.h, _dt(A.data), cl, cr, d, W.dims[0], W.dims[3], _ptr(R.data),
                                       _ptr(A.data), _ptr(W.data), _ptr(out.data), _stream()))
    return out


def eigsolve_lanczos(L, W1, W2, R, phi, krylovdim=3, maxiter=1, tol=1e-14):
    """In place on phi.  Returns (energy, n_matvec)."""
    h = _lib.handle()
    bd = bond_dims(L, W1, W2, R, phi)
    e = C.c_double(0.0)
    nmv = C.c_int(0)
    h.check(h.lib.tnb_eigsolve_lanczos(h.h, _dt(phi.data), C.byref(bd), _ptr(L.data), _ptr(W1.data), _ptr(W2.data),
                                       _ptr(R.data), _ptr(phi.data), krylovdim, maxiter, tol, C.byref(e),
                                       C.byref(nmv), _stream()))
    return e.value, nmv.value


def noise_term(L, W1, W2, R, phi, ortho, noise, rho=None):
    h = _lib.handle()
    bd = bond_dims(L, W1, W2, R, phi)
    cl, d1, d2, cr = phi.dims
    m = cl * d1 if ortho == "left" else d2 * cr
    acc = 1
    if rho is None:
        rho = DTensor.empty((m, m), phi.dtype, phi.data.device)
        acc = 0
    h.check(h.lib.tnb_noise_term(h.h, _dt(phi.data), C.byref(bd), _ptr(L.data), _ptr(W1.data), _ptr(W2.data),
                                 _ptr(R.data), _ptr(phi.data), 0 if ortho == "left" else 1, float(noise), acc,
                                 _ptr(rho.data), _stream()))
    return rho


# ------------------------------------------------------------------ factorizations
def _matrix_dims(M):
    if len(M.dims) != 2:
        raise _lib.TnbError(1, "rank-2 tensor expected, got dims %s" % (M.dims,))
    return M.dims


def svd(M, maxdim=None, mindim=1, cutoff=None, use_absolute_cutoff=False, use_relative_cutoff=True):
    """Thin SVD + truncation.  Returns (U[m,k], S[k] (torch), V[n,k], truncerr) with M ~ U diag(S) V^T
    (``svd``: reference ``src/tensor/culinearalgebra.jl:33-72``; CPU ``conj!(MV)`` convention)."""
    h = _lib.handle()
    m, n = _matrix_dims(M)
    do_trunc = maxdim is not None or cutoff is not None
    kfull = min(m, n)
    kmax = min(kfull, int(maxdim)) if (do_trunc and maxdim is not None) else kfull
    A = M.clone()
    U = DTensor.empty((m, kmax), M.dtype, M.data.device)
    V = DTensor.empty((n, kmax), M.dtype, M.data.device)
    S = torch.empty(kmax, dtype=torch.float64, device=M.data.device)
    nk = C.c_int64(0)
    err = C.c_double(0.0)
    flags = (1 if use_absolute_cutoff else 0) | (0 if use_relative_cutoff else 2)
    h.check(h.lib.tnb_svd_trunc(h.h, _dt(M.data), m, n, _ptr(A.data), int(maxdim) if maxdim is not None else 0,
                                int(mindim), float(cutoff or 0.0), flags, 1 if do_trunc else 0, _ptr(U.data), _ptr(S),
                                _ptr(V.data), C.byref(nk), C.byref(err), _stream()))
    k = nk.value
    return (DTensor(U.data[: m * k], (m, k)), S[:k], DTensor(V.data[: n * k], (n, k)), err.value)


def eigh(M, maxdim=None, mindim=1, cutoff=None, use_absolute_cutoff=False, use_relative_cutoff=True):
    """Hermitian eigendecomposition, eigenvalues descending + truncation.  Returns (D[k] torch, U[n,k], truncerr)
    (``eigen``: reference ``src/tensor/culinearalgebra.jl:74-108``)."""
    h = _lib.handle()
    n, n2 = _matrix_dims(M)
    if n != n2:
        raise _lib.DimensionMismatch(2, "eigh: matrix is %dx%d" % (n, n2))
    do_trunc = maxdim is not None or cutoff is not None
    kmax = min(n, int(maxdim)) if (do_trunc and maxdim is not None) else n
    A = M.clone()
    U = DTensor.empty((n, kmax), M.dtype, M.data.device)
    D = torch.empty(kmax, dtype=torch.float64, device=M.data.device)
    nk = C.c_int64(0)
    err = C.c_double(0.0)
    flags = (1 if use_absolute_cutoff else 0) | (0 if use_relative_cutoff else 2)
    h.check(h.lib.tnb_eigh_trunc(h.h, _dt(M.data), n, _ptr(A.data), int(maxdim) if maxdim is not None else 0,
                                 int(mindim), float(cutoff or 0.0), flags, 1 if do_trunc else 0, _ptr(D), _ptr(U.data),
                                 C.byref(nk), C.byref(err), _stream()))
    k = nk.value
    return D[:k], DTensor(U.data[: n * k], (n, k)), err.value


def qr(M):
    """Thin QR, explicit Q, diag(R) >= 0 (``qr``: reference ``src/tensor/culinearalgebra.jl:110-121``)."""
    h = _lib.handle()
    m, n = _matrix_dims(M)
    k = min(m, n)
    Q = DTensor.empty((m, k), M.dtype, M.data.device)
    R = DTensor.empty((k, n), M.dtype, M.data.device)
    h.check(h.lib.tnb_qr(h.h, _dt(M.data), m, n, _ptr(M.data), _ptr(Q.data), _ptr(R.data), _stream()))
    return Q, R


_DECOMP = {None: 0, "automatic": 0, "svd": 1, "eigen": 2, "qr": 3}


def factorize_bond(phi, ortho="left", which_decomp=None, maxdim=None, mindim=1, cutoff=0.0, rho_pert=None,
                   normalize=False):
    """phi[l,s1,s2,r] -> A[l,s1,k], B[k,s2,r]  ([EXT] replacebond!/factorize).  Returns (A, B, truncerr)."""
    h = _lib.handle()
    cl, d1, d2, cr = phi.dims
    m, n = cl * d1, d2 * cr
    if which_decomp == "eigen" or (which_decomp in (None, "automatic") and (rho_pert is not None or (cutoff or 0.0) > 1e-12)):
        rfull = m if ortho == "left" else n
    else:
        rfull = min(m, n)
    kmax = min(rfull, int(maxdim)) if maxdim is not None else rfull
    kmax = max(kmax, 1)
    bd = BondDims(cl, cr, d1, d2, 1, 1, 1)
    work = phi.clone()
    A = DTensor.empty((m * kmax,), phi.dtype, phi.data.device)
    B = DTensor.empty((kmax * n,), phi.dtype, phi.data.device)
    nk = C.c_int64(0)
    err = C.c_double(0.0)
    h.check(h.lib.tnb_factorize_bond(h.h, _dt(phi.data), C.byref(bd), _ptr(work.data), 0 if ortho == "left" else 1,
                                     _DECOMP[which_decomp], int(maxdim) if maxdim is not None else 0, int(mindim),
                                     float(cutoff or 0.0), _ptr(rho_pert.data) if rho_pert is not None else None,
                                     1 if normalize else 0, _ptr(A.data), _ptr(B.data), C.byref(nk), C.byref(err),
                                     _stream()))
    k = nk.value
    return DTensor(A.data[: m * k], (cl, d1, k)), DTensor(B.data[: k * n], (k, d2, cr)), err.value


def dmrg_bond_step(L, W1, W2, R, A1, A2, ortho, maxdim, mindim=1, cutoff=0.0, noise=0.0, krylovdim=3, maxiter=1,
                   which_decomp=None):
    """One two-site DMRG bond update through the single fused C call.  Returns (energy, A1', A2', truncerr)."""
    h = _lib.handle()
    cl, d1, cm = A1.dims
    cm2, d2, cr = A2.dims
    if cm != cm2:
        raise _lib.DimensionMismatch(2, "A1/A2 middle bond differs")
    m, n = cl * d1, d2 * cr
    use_eigen = which_decomp == "eigen" or (which_decomp in (None, "automatic") and (noise > 0 or (cutoff or 0.0) > 1e-12))
    rfull = (m if ortho == "left" else n) if use_eigen else min(m, n)
    kmax = max(1, min(rfull, int(maxdim)))
    bd = BondDims(cl, cr, d1, d2, W1.dims[0], W1.dims[3], W2.dims[3])
    dev = A1.data.device
    b1 = torch.empty(max(A1.size, m * kmax), dtype=A1.dtype, device=dev)
    b2 = torch.empty(max(A2.size, kmax * n), dtype=A2.dtype, device=dev)
    b1[: A1.size].copy_(A1.data)
    b2[: A2.size].copy_(A2.data)
    e = C.c_double(0.0)
    nk = C.c_int64(0)
    err = C.c_double(0.0)
    h.check(h.lib.tnb_dmrg_bond_step(h.h, _dt(A1.data), C.byref(bd), cm, _ptr(L.data), _ptr(W1.data), _ptr(W2.data),
                                     _ptr(R.data), _ptr(b1), _ptr(b2), 0 if ortho == "left" else 1,
                                     _DECOMP[which_decomp], int(maxdim), int(mindim), float(cutoff or 0.0),
                                     float(noise), int(krylovdim), int(maxiter), C.byref(e), C.byref(nk),
                                     C.byref(err), _stream()))
    k = nk.value
    return e.value, DTensor(b1[: m * k], (cl, d1, k)), DTensor(b2[: k * n], (k, d2, cr)), err.value


def tebd_apply_gate(G, A1, A2, maxdim=None, mindim=1, cutoff=0.0):
    """theta = G * (A1*A2), left-orthogonal split ([EXT] apply / product(o, psi)).  Returns (A1', A2', truncerr)."""
    h = _lib.handle()
    cl, d1, cm = A1.dims
    _, d2, cr = A2.dims
    m, n = cl * d1, d2 * cr
    G, A1 = _promote(G, A1)
    G, A2 = _promote(G, A2)
    use_eigen = (cutoff or 0.0) > 1e-12
    rfull = m if use_eigen else min(m, n)
    kmax = max(1, min(rfull, int(maxdim)) if maxdim is not None else rfull)
    dev = A1.data.device
    b1 = torch.empty(max(A1.size, m * kmax), dtype=A1.dtype, device=dev)
    b2 = torch.empty(max(A2.size, kmax * n), dtype=A2.dtype, device=dev)
    b1[: A1.size].copy_(A1.data)
    b2[: A2.size].copy_(A2.data)
    nk = C.c_int64(0)
    err = C.c_double(0.0)
    h.check(h.lib.tnb_tebd_apply_gate(h.h, _dt(A1.data), cl, cm, cr, d1, d2, _ptr(G.data), _ptr(b1), _ptr(b2),
                                      int(maxdim) if maxdim is not None else 0, int(mindim), float(cutoff or 0.0),
                                      C.byref(nk), C.byref(err), _stream()))
    k = nk.value
    return DTensor(b1[: m * k], (cl, d1, k)), DTensor(b2[: k * n], (k, d2, cr)), err.value


def tebd_gate_bform(G, lamL, B1, B2, maxdim=None, mindim=1, cutoff=0.0):
    """Two-site gate in B form (``tnb_tebd_gate_bform``): B1, B2 right-canonical in the Schmidt bases, ``lamL`` the
    Schmidt values (1-D float64 device tensor) of the bond left of the first site.
    Returns (B1', B2', lam' (torch), truncerr)."""
    h = _lib.handle()
    cl, d1, cm = B1.dims
    _, d2, cr = B2.dims
    m, n = cl * d1, d2 * cr
    G, B1 = _promote(G, B1)
    G, B2 = _promote(G, B2)
    B1, B2 = _promote(B1, B2)
    if lamL.numel() != cl:
        raise _lib.DimensionMismatch(2, "tebd_gate_bform: %d Schmidt values for a bond of dimension %d" % (lamL.numel(), cl))
    kmax = max(1, min(m, n, int(maxdim)) if maxdim is not None else min(m, n))
    dev = B1.data.device
    b1 = torch.empty(max(B1.size, m * kmax), dtype=B1.dtype, device=dev)
    b2 = torch.empty(max(B2.size, kmax * n), dtype=B2.dtype, device=dev)
    b1[: B1.size].copy_(B1.data)
    b2[: B2.size].copy_(B2.data)
    lam = torch.empty(kmax, dtype=torch.float64, device=dev)
    lamL = lamL.to(torch.float64).contiguous()
    nk = C.c_int64(0)
    err = C.c_double(0.0)
    h.check(h.lib.tnb_tebd_gate_bform(h.h, _dt(B1.data), cl, cm, cr, d1, d2, _ptr(G.data), _ptr(lamL), _ptr(b1), _ptr(b2),
                                      int(maxdim) if maxdim is not None else 0, int(mindim), float(cutoff or 0.0),
                                      _ptr(lam), C.byref(nk), C.byref(err), _stream()))
    k = nk.value
    return DTensor(b1[: m * k], (cl, d1, k)), DTensor(b2[: k * n], (k, d2, cr)), lam[:k], err.value


def diag_contract(A, la, mode, diag, lc):
    """C[lc] <- A[la] * diag[index of ``mode``]  (``lc`` a permutation of ``la``; ``tnb_diag_contract``):
    the contraction of a dense tensor with a Diag tensor over one of its two indices.  ``diag``: 1-D torch CUDA
    tensor (float64, or complex128 with a complex A)."""
    h = _lib.handle()
    la, lc = list(la), list(lc)
    if len(la) != len(lc) or set(la) != set(lc):
        raise _lib.TnbError(1, "diag_contract: output labels must be a permutation of the input labels")
    if mode not in la:
        raise _lib.TnbError(1, "diag_contract: %r is not a mode of A" % (mode,))
    if diag.numel() != A.dims[la.index(mode)]:
        raise _lib.DimensionMismatch(2, "diag_contract: diagonal of length %d for a mode of extent %d" %
                                     (diag.numel(), A.dims[la.index(mode)]))
    if diag.dtype == torch.complex128 and A.dtype != torch.complex128:
        A = A.astype(torch.complex128)
    ids = {l: i for i, l in enumerate(la)}
    out = DTensor.empty(tuple(A.dims[la.index(l)] for l in lc), A.dtype, A.data.device)
    diag = diag.contiguous()
    h.check(h.lib.tnb_diag_contract(h.h, _dt(A.data), len(la), _i64(A.dims), _i32([ids[l] for l in la]), _ptr(A.data),
                                    ids[mode], _ptr(diag), _dt(diag), _i32([ids[l] for l in lc]), _ptr(out.data), _stream()))
    return out
