"""Host-side mirror of the reference's ITensor-level interface for the hot path.

Same names and argument meaning as the Julia API the reference plugs into (exports:
``src/ITensorsGPU.jl:57-65``; overridden operations: ``src/ITensorsGPU.jl:32-43``), so the parity
tests read like ``test/test_cuitensor.jl`` / ``test_cucontract.jl``.  Index bookkeeping (ids,
tags, prime levels) stays on the host exactly as in the reference; every
arithmetic operation is one call into libtnb200.so.  A CPU ITensor here is only a container for
H2D/D2H (``cu`` / ``cpu``); there is no CPU arithmetic path.
"""
import itertools

import numpy as np
import torch

from . import _lib, ops
from .ops import DTensor

_ids = itertools.count(1)


class Index:
    """[EXT] ITensors ``Index``: (id, dim, tags, prime level)."""

    __slots__ = ("id", "dim", "tags", "plev")

    def __init__(self, dim, tags="", plev=0, id=None):
        self.id = next(_ids) if id is None else id
        self.dim = int(dim)
        self.tags = tags
        self.plev = int(plev)

    def __eq__(self, o):
        return isinstance(o, Index) and (self.id, self.plev) == (o.id, o.plev)

    def __hash__(self):
        return hash((self.id, self.plev))

    def __repr__(self):
        return "(dim=%d|id=%d|%s)%s" % (self.dim, self.id, self.tags, "'" * self.plev)

    def prime(self, n=1):
        return Index(self.dim, self.tags, self.plev + n, self.id)

    def noprime(self):
        return Index(self.dim, self.tags, 0, self.id)

    def sim(self):
        return Index(self.dim, self.tags, self.plev)


def prime(x, *a, **k):
    return x.prime(*a, **k)


def dim(i):
    return i.dim


class Spectrum:
    def __init__(self, eigs, truncerr):
        self.eigs = eigs
        self.truncerr = truncerr


class DiagStore:
    """Diagonal storage ([EXT] NDTensors ``Diag``; GPU form ``src/tensor/cudiag.jl``): a rank-2 ITensor whose only
    non-zero entries are ``vec[j]`` at (j, j).  ``vec``: 1-D torch CUDA tensor.  This is what ``svd`` returns for S
    and ``eigen`` for D on the CPU path; contracting it with a dense tensor is a scale along one mode
    (``tnb_diag_contract``), not a GEMM with a densified k x k matrix."""

    def __init__(self, vec):
        self.vec = vec

    @property
    def dtype(self):
        return self.vec.dtype

    def dense(self):
        k = self.vec.numel()
        return DTensor(torch.diag(self.vec).reshape(-1).contiguous(), (k, k))


class UniformDiagStore:
    """Uniform diagonal storage ([EXT] NDTensors ``Diag(x::Number)``: every diagonal entry equals ``value``; the
    reference's ``UniformDiagTensor`` cases, ``src/tensor/cudiag.jl:105-133``).  Holds no device memory."""

    def __init__(self, value, k):
        self.value = complex(value) if isinstance(value, complex) and value.imag != 0 else float(np.real(value))
        self.k = int(k)

    @property
    def dtype(self):
        return torch.complex128 if isinstance(self.value, complex) else torch.float64

    def dense(self):
        return DTensor(torch.diag(torch.full((self.k,), self.value, dtype=self.dtype, device="cuda")).reshape(-1).contiguous(),
                       (self.k, self.k))


class ITensor:
    """ITensor.  ``store`` is a DTensor (GPU dense, ``CuDense``), a DiagStore (GPU diagonal), a UniformDiagStore
    or a NumPy array (CPU container)."""

    def __init__(self, store, inds):
        self.inds = tuple(inds)
        dims = tuple(i.dim for i in self.inds)
        if isinstance(store, UniformDiagStore):
            if len(dims) != 2 or dims[0] != dims[1] or dims[0] != store.k:
                raise _lib.DimensionMismatch(2, "uniform Diag storage of length %d for indices of dims %s" % (store.k, dims))
        elif isinstance(store, DiagStore):
            if len(dims) != 2 or dims[0] != dims[1] or dims[0] != store.vec.numel():
                raise _lib.DimensionMismatch(2, "Diag storage of length %d for indices of dims %s" % (store.vec.numel(), dims))
        elif isinstance(store, DTensor):
            if store.dims != dims:
                store = DTensor(store.data, dims)
        else:
            store = np.asarray(store)
            if store.shape != dims:
                raise _lib.DimensionMismatch(2, "array of shape %s for indices of dims %s" % (store.shape, dims))
        self.store = store

    # ---- placement
    @property
    def on_gpu(self):
        return isinstance(self.store, (DTensor, DiagStore, UniformDiagStore))

    @property
    def is_diag(self):
        return isinstance(self.store, (DiagStore, UniformDiagStore))

    @property
    def is_uniform_diag(self):
        return isinstance(self.store, UniformDiagStore)

    def _dev(self):
        if not self.on_gpu:
            raise _lib.TnbError(3, "arithmetic on a CPU ITensor: move it with cu(); there is no CPU path")
        return self.store.dense() if self.is_diag else self.store

    def array(self):
        """Logical ndarray on the host (``array(cpu(A))``)."""
        if self.is_uniform_diag:
            return np.eye(self.store.k) * self.store.value
        if self.is_diag:
            return np.diag(self.store.vec.cpu().numpy())
        return self.store.numpy() if self.on_gpu else np.array(self.store)

    def scalar(self):
        if len(self.inds) != 0:
            raise _lib.DimensionMismatch(2, "scalar() of a rank-%d ITensor" % len(self.inds))
        v = self.array().reshape(())
        return complex(v) if np.iscomplexobj(v) else float(v)

    # ---- index manipulation (host only)
    def prime(self, n=1, which=None):
        return ITensor(self.store, [i.prime(n) if (which is None or i == which) else i for i in self.inds])

    def noprime(self):
        return ITensor(self.store, [i.noprime() for i in self.inds])

    def replaceinds(self, old, new):
        m = dict(zip(old, new))
        return ITensor(self.store, [m.get(i, i) for i in self.inds])

    def dag(self):
        if self.is_uniform_diag:
            return ITensor(UniformDiagStore(np.conj(self.store.value), self.store.k), self.inds)
        if self.is_diag:
            return ITensor(DiagStore(torch.conj_physical(self.store.vec)), self.inds)
        s = self._dev()
        if s.dtype == torch.complex128:
            return ITensor(DTensor(torch.conj_physical(s.data), s.dims), self.inds)
        return self

    # ---- arithmetic: every op is one C-ABI call
    def _diag_times_dense(self, o, diag_first):
        """self: Diag (u, v); o: dense sharing exactly one of u, v.  Output order = NDTensors' (first operand's free
        indices, then the second's): one scale-along-a-mode pass (+ a permute when the mode moves)."""
        u, v = self.inds
        shared = u if u in o.inds else v
        other = v if shared == u else u
        free = [i for i in o.inds if i != shared]
        out_inds = ([other] + free) if diag_first else (free + [other])
        lc = [shared if i == other else i for i in out_inds]          # labels of the kernel: the relabel is host-side
        C = ops.diag_contract(o.store, o.inds, shared, self.store.vec, lc)
        return ITensor(C, out_inds)

    def _uniform_times_dense(self, o, diag_first):
        """self: UniformDiag (u, v); o: dense sharing exactly one of u, v: a scale fused into the (possibly identity)
        permutation to the NDTensors output order -- ONE pass (``cudiag.jl:105-118`` multiplies in place and ignores
        the order, which is why ``Aij*scal`` is @test_broken at ``test/test_cudiag.jl:49,95``; here both orders hold)."""
        u, v = self.inds
        shared = u if u in o.inds else v
        other = v if shared == u else u
        free = [i for i in o.inds if i != shared]
        out_inds = ([other] + free) if diag_first else (free + [other])
        lc = [shared if i == other else i for i in out_inds]
        src = o.store
        val = self.store.value
        if isinstance(val, complex) and src.dtype != torch.complex128:
            src = src.astype(torch.complex128)
        out = DTensor.empty(tuple(i.dim for i in out_inds), src.dtype, src.data.device)
        ops.permute_axpby(src, o.inds, out, lc, alpha=val, beta=0.0)
        return ITensor(out, out_inds)

    def _diag_times_diag(self, o):
        """Diag x Diag sharing exactly one index -> Diag over the two remaining indices: an elementwise product of the
        two diagonals (``cudiag.jl:120-145``), or a scale when one of them is uniform."""
        shared = [i for i in self.inds if i in o.inds][0]
        a = [i for i in self.inds if i != shared][0]
        b = [i for i in o.inds if i != shared][0]
        if self.is_uniform_diag and o.is_uniform_diag:
            return ITensor(UniformDiagStore(self.store.value * o.store.value, self.store.k), (a, b))
        if self.is_uniform_diag or o.is_uniform_diag:
            un, dg = (self, o) if self.is_uniform_diag else (o, self)
            vec = dg.store.vec.clone()
            if isinstance(un.store.value, complex) and vec.dtype != torch.complex128:
                vec = vec.to(torch.complex128)
            ops.scale(DTensor(vec, (vec.numel(),)), un.store.value)
            return ITensor(DiagStore(vec), (a, b))
        x, y = self.store.vec, o.store.vec
        if x.dtype != y.dtype:
            x, y = x.to(torch.complex128), y.to(torch.complex128)
        if y.dtype == torch.complex128 or x.dtype == torch.float64:
            out = ops.diag_contract(DTensor(x.contiguous(), (x.numel(),)), ("k",), "k", y, ("k",))
        else:
            out = ops.diag_contract(DTensor(y.contiguous(), (y.numel(),)), ("k",), "k", x, ("k",))
        return ITensor(DiagStore(out.data), (a, b))

    def __mul__(self, o):
        if isinstance(o, ITensor):
            if self.on_gpu and o.on_gpu and self.is_diag and o.is_diag and sum(i in o.inds for i in self.inds) == 1:
                return self._diag_times_diag(o)
            if self.on_gpu and o.on_gpu and (self.is_diag != o.is_diag):            # cudiag.jl:105-161 without densifying
                dg, dn = (self, o) if self.is_diag else (o, self)
                if sum(i in dn.inds for i in dg.inds) == 1:
                    if dg.is_uniform_diag:
                        return dg._uniform_times_dense(dn, diag_first=self.is_diag)
                    return dg._diag_times_dense(dn, diag_first=self.is_diag)
            C, lc = ops.contract(self._dev(), self.inds, o._dev(), o.inds)        # contract!! (cudense.jl:83-110)
            return ITensor(C, lc)
        out = self._dev().clone()
        if isinstance(o, complex) and out.dtype != torch.complex128:
            out = out.astype(torch.complex128)
        return ITensor(ops.scale(out, o), self.inds)                              # cudense.jl:22

    __rmul__ = __mul__

    def __truediv__(self, x):
        return self * (1.0 / x)                                                   # cudense.jl:502

    def _addsub(self, o, sgn):
        if set(self.inds) != set(o.inds) or len(self.inds) != len(o.inds):
            raise _lib.DimensionMismatch(2, "cannot add ITensors with different index sets")
        a, b = self._dev(), o._dev()
        if a.dtype != b.dtype:
            a, b = a.astype(torch.complex128), b.astype(torch.complex128)
        out = a.clone()
        ops.permute_axpby(b, o.inds, out, self.inds, alpha=sgn, beta=1.0)         # cudense.jl:333-445
        return ITensor(out, self.inds)

    def __add__(self, o):
        return self._addsub(o, 1.0)

    def __sub__(self, o):
        return self._addsub(o, -1.0)

    def __neg__(self):
        return self * -1.0


def norm(A):
    return ops.norm(A._dev())                                                     # cudense.jl:27


def dot(A, B):
    """<A|B> = scalar(dag(A)*B)."""
    if set(A.inds) != set(B.inds):
        raise _lib.DimensionMismatch(2, "dot of ITensors with different index sets")
    b = B if A.inds == B.inds else permute(B, A.inds)
    a, bb = A._dev(), b._dev()
    if a.dtype != bb.dtype:
        a, bb = a.astype(torch.complex128), bb.astype(torch.complex128)
    return ops.dot(a, bb)


def permute(A, inds):
    return ITensor(ops.permute(A._dev(), A.inds, tuple(inds)), inds)             # permute! (cudense.jl:447-478)


def dag(A):
    return A.dag()


def noprime(A):
    return A.noprime()


def commonind(A, B):
    c = [i for i in A.inds if i in B.inds]
    return c[0] if c else None


def delta(i, j, dtype=np.float64):
    return cuITensor(np.eye(i.dim, j.dim, dtype=dtype), (i, j))


def diagITensor(x, i, j):
    """``itensor(tensor(Diag(x), IndexSet(i, j)))`` (``test/test_cudiag.jl:30-44``): a vector gives (non-uniform) Diag
    storage on the GPU, a number gives uniform Diag storage."""
    if np.isscalar(x):
        if i.dim != j.dim:
            raise _lib.DimensionMismatch(2, "diagITensor: indices of dims %d, %d" % (i.dim, j.dim))
        return ITensor(UniformDiagStore(x, i.dim), (i, j))
    v = np.asarray(x)
    v = v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)
    return ITensor(DiagStore(torch.from_numpy(np.ascontiguousarray(v)).cuda()), (i, j))


# ---- constructors / transfer (src/cuitensor.jl)
def cuITensor(x=None, inds=()):
    """``cuITensor(inds...)`` zero-filled, ``cuITensor(x::Number, inds)`` constant, ``cuITensor(A::Array, inds)``
    (``src/cuitensor.jl:1-26``)."""
    inds = tuple(inds)
    dims = tuple(i.dim for i in inds)
    if x is None:
        return ITensor(DTensor.zeros(dims), inds)
    if np.isscalar(x):
        return ITensor(DTensor.from_numpy(np.full(dims, x)), inds)
    x = np.asarray(x)
    if x.size != int(np.prod(dims, dtype=np.int64)):
        raise _lib.DimensionMismatch(2, "array has %d elements, indices need %d" % (x.size, int(np.prod(dims))))
    return ITensor(DTensor.from_numpy(x.reshape(dims, order="F") if x.shape != dims else x), inds)


def randomCuITensor(*inds, dtype=np.float64, rng=None):
    """``randomCuITensor`` (``src/cuitensor.jl:35-49``); values are drawn on the host and uploaded."""
    rng = rng or np.random.default_rng()
    dims = tuple(i.dim for i in inds)
    a = rng.standard_normal(dims)
    if np.issubdtype(dtype, np.complexfloating):
        a = a + 1j * rng.standard_normal(dims)
    return ITensor(DTensor.from_numpy(a), inds)


def cu(x):
    """``cu`` (``src/cuitensor.jl:28``, ``src/mps/cumps.jl:9``, ``src/mps/cumpo.jl:13``)."""
    from . import mps as _mps
    if isinstance(x, ITensor):
        return x if x.on_gpu else ITensor(DTensor.from_numpy(x.store), x.inds)
    if isinstance(x, (_mps.MPS, _mps.MPO)):
        return x.cu()
    raise TypeError("cu: unsupported %r" % type(x))


def cpu(x):
    """``cpu`` (``src/cuitensor.jl:30-33``, ``src/mps/cumpo.jl:25-31``)."""
    from . import mps as _mps
    if isinstance(x, ITensor):
        return ITensor(x.array(), x.inds)
    if isinstance(x, (_mps.MPS, _mps.MPO)):
        return x.cpu()
    raise TypeError("cpu: unsupported %r" % type(x))


# ---- factorizations ([EXT] decomp.jl: combiner to rank 2, then the CuDense methods)
def _matricize(A, Linds):
    Linds = [i for i in Linds if i in A.inds]
    Rinds = [i for i in A.inds if i not in Linds]
    order = tuple(Linds) + tuple(Rinds)
    T = A if A.inds == order else permute(A, order)          # the combiner's permute (K3)
    m = int(np.prod([i.dim for i in Linds], dtype=np.int64))
    n = int(np.prod([i.dim for i in Rinds], dtype=np.int64))
    return DTensor(T._dev().data, (m, n)), tuple(Linds), tuple(Rinds)


def svd(A, Linds, **kw):
    """``U,S,V,spec = svd(A, Linds...; maxdim, mindim, cutoff)`` with ``A ~ U*S*V`` (CPU convention;
    GPU body replaced: ``src/tensor/culinearalgebra.jl:33-72``).  S is a Diag-storage ITensor."""
    M, L, R = _matricize(A, Linds)
    U, S, V, err = ops.svd(M, **kw)
    k = U.dims[1]
    u, v = Index(k, "Link,u"), Index(k, "Link,v")
    Ut = ITensor(DTensor(U.data, tuple(i.dim for i in L) + (k,)), L + (u,))
    Vt = ITensor(DTensor(V.data, tuple(i.dim for i in R) + (k,)), R + (v,))
    St = ITensor(DiagStore(S), (u, v))            # Diag storage, as on the CPU path (the reference's GPU path densifies)
    return Ut, St, Vt, Spectrum((S ** 2).cpu().numpy(), err)


def eigen(A, Linds, Rinds, **kw):
    """``D,U,spec = eigen(A, Linds, Rinds; ishermitian=true, ...)`` (``culinearalgebra.jl:74-108``)."""
    kw.pop("ishermitian", None)
    order = tuple(Linds) + tuple(Rinds)
    T = A if A.inds == order else permute(A, order)
    n = int(np.prod([i.dim for i in Linds], dtype=np.int64))
    D, U, err = ops.eigh(DTensor(T._dev().data, (n, n)), **kw)
    k = U.dims[1]
    l, r = Index(k, "Link,eigen"), Index(k, "Link,eigen")
    Ut = ITensor(DTensor(U.data, tuple(i.dim for i in Rinds) + (k,)), tuple(Rinds) + (r,))
    Dt = ITensor(DiagStore(D), (l, r))            # Diag(real(D)), as culinearalgebra.jl:106 / the CPU path
    return Dt, Ut, Spectrum(D.cpu().numpy(), err)


def qr(A, Linds):
    """``Q,R = qr(A, Linds...)`` (``src/tensor/culinearalgebra.jl:110-121``)."""
    M, L, R = _matricize(A, Linds)
    Q, Rm = ops.qr(M)
    k = Q.dims[1]
    q = Index(k, "Link,qr")
    return (ITensor(DTensor(Q.data, tuple(i.dim for i in L) + (k,)), L + (q,)),
            ITensor(DTensor(Rm.data, (k,) + tuple(i.dim for i in R)), (q,) + R))


def davidson(A, phi0, maxiter=2, miniter=1, errgoal=1e-14):
    """[EXT] ITensors ``davidson(A, phi0; maxiter)`` on a callable ``A(v::ITensor)`` -- every vector
    operation (``A(v)``, dot, axpy, norm) is a device call (``test/test_cuiterativesolvers.jl:13-28``)."""
    phi = phi0 / norm(phi0)
    V, AV = [phi], [A(phi)]
    lam = dot(V[0], AV[0]).real if isinstance(dot(V[0], AV[0]), complex) else dot(V[0], AV[0])
    q = AV[0] - V[0] * lam
    M = np.array([[lam]], dtype=complex)
    last = lam
    for ni in range(1, maxiter + 1):
        if norm(q) < 1e-12 and ni > miniter:
            break
        for _ in range(2):
            for u in V:
                q = q - u * dot(u, q)
        qn = norm(q)
        if qn < 1e-10:
            break
        q = q / qn
        V.append(q)
        AV.append(A(q))
        k = len(V)
        Mn = np.zeros((k, k), dtype=complex)
        Mn[: k - 1, : k - 1] = M
        for i in range(k):
            Mn[i, k - 1] = dot(V[i], AV[k - 1])
            Mn[k - 1, i] = np.conj(Mn[i, k - 1])
        M = Mn
        ev, U = np.linalg.eigh(M)
        lam = float(ev[0])
        y = U[:, 0]
        cplx = any(v._dev().dtype == torch.complex128 for v in V + AV)
        if not cplx:
            y = y.real
        phi = V[0] * complex(y[0]) if cplx else V[0] * float(y[0])
        Aphi = AV[0] * (complex(y[0]) if cplx else float(y[0]))
        for i in range(1, k):
            c = complex(y[i]) if cplx else float(y[i])
            phi = phi + V[i] * c
            Aphi = Aphi + AV[i] * c
        q = Aphi - phi * lam
        if abs(lam - last) < errgoal and ni >= miniter and norm(q) < np.sqrt(errgoal):
            break
        last = lam
    return lam, phi / norm(phi)
