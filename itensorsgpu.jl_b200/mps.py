"""Host-side mirror of the MPS / MPO level of the hot path: ``cu(psi)``, ``dmrg``, ``apply``,
``inner``, ``orthogonalize`` ([EXT] ITensors.jl 0.2 algorithms the reference reaches through its
overrides -- call sites ``examples/dmrg.jl:25``, ``examples/gate_evolution.jl:46``,
``test/dmrg.jl:27,75``, ``test/test_cumps.jl``, ``test/test_cumpo.jl``).

Fixed layouts (column-major flat buffers): MPS site ``A[l,s,r]``, MPO site ``W[a,s,s',b]`` (s = ket),
environments ``L[l,l',a]`` / ``R[r,r',c]``.  All arithmetic is libtnb200.so calls; the sweep loop is
host control flow only (which bond, which environment), like the Julia ``dmrg`` driver.
"""
import numpy as np
import torch

from . import _lib, ops
from .ops import DTensor


def _to_dev(a):
    return a if isinstance(a, DTensor) else DTensor.from_numpy(a)


def _to_host(a):
    return a.numpy() if isinstance(a, DTensor) else np.asarray(a)


class _Chain:
    def __init__(self, tensors):
        self.tensors = list(tensors)

    def __len__(self):
        return len(self.tensors)

    def __getitem__(self, j):
        return self.tensors[j]

    def __setitem__(self, j, v):
        self.tensors[j] = v

    @property
    def on_gpu(self):
        return all(isinstance(t, DTensor) for t in self.tensors)

    def _moved(self, f):
        out = type(self).__new__(type(self))
        out.__dict__.update(self.__dict__)
        out.tensors = [f(t) for t in self.tensors]
        return out

    def cu(self):
        """site-wise H2D (``cuMPS``/``cuMPO``: ``src/mps/cumps.jl:1-7``, ``src/mps/cumpo.jl:1-7``)."""
        return self._moved(_to_dev)

    def cpu(self):
        """site-wise D2H (``cpu(::MPS/MPO)``: ``src/mps/cumpo.jl:25-31``)."""
        return self._moved(_to_host)

    def dims(self, j):
        t = self.tensors[j]
        return t.dims if isinstance(t, DTensor) else t.shape


class MPS(_Chain):
    """``llim``/``rlim`` as in ITensors: sites <= llim are left-orthogonal, sites >= rlim right-orthogonal."""

    def __init__(self, tensors, llim=-1, rlim=None):
        super().__init__(tensors)
        self.llim = llim
        self.rlim = len(self.tensors) if rlim is None else rlim

    def maxlinkdim(self):
        return max(self.dims(j)[2] for j in range(len(self) - 1)) if len(self) > 1 else 1

    @property
    def center(self):
        return self.llim + 1 if self.rlim == self.llim + 2 else None


class MPO(_Chain):
    pass


# ---- constructors (src/mps/cumps.jl, src/mps/cumpo.jl); values are made on the host and uploaded
def cuMPS(psi):
    return psi.cu()


def cuMPO(H):
    return H.cu()


def productCuMPS(d, states):
    """``productCuMPS(sites, states)`` (``src/mps/cumps.jl:58-134``): bond dimension 1 product state."""
    ts = []
    for s in states:
        A = np.zeros((1, d, 1))
        A[0, int(s), 0] = 1.0
        ts.append(A)
    return MPS(ts, llim=-1, rlim=1).cu()


def randomCuMPS(N, d, chi=1, seed=None, dtype=np.float64):
    """``randomCuMPS(sites)`` (``src/mps/cumps.jl:44-56``: bond dimension 1; ``chi`` generalises it).
    Site tensors are random isometries, orthogonality centre at site 0."""
    rng = np.random.default_rng(seed)
    D = [int(min(chi, d ** min(k, N - k, 40))) for k in range(N + 1)]
    ts = []
    for j in range(N):
        l, r = D[j], D[j + 1]
        G = rng.standard_normal((d * r, l))
        if np.issubdtype(dtype, np.complexfloating):
            G = (G + 1j * rng.standard_normal((d * r, l))) / np.sqrt(2)
        Q, _ = np.linalg.qr(G)
        ts.append(np.ascontiguousarray(Q.T.reshape(l, d, r, order="F").astype(dtype)))
    ts[0] = ts[0] / np.linalg.norm(ts[0])
    return MPS(ts, llim=-1, rlim=1).cu()


def randomCuMPO(N, d, seed=None, linkdim=1):
    """``randomCuMPO(sites)`` (``src/mps/cumpo.jl:15-23``: link dimension 1 only, where anything else is an
    ArgumentError, ``:21``; ``linkdim`` generalises it)."""
    if linkdim < 1:
        raise _lib.TnbError(1, "randomCuMPO: link dimension must be >= 1")
    rng = np.random.default_rng(seed)
    return MPO([rng.standard_normal((1 if j == 0 else linkdim, d, d, 1 if j == N - 1 else linkdim)) for j in range(N)]).cu()


# ---- measurements
def inner(phi, psi, H=None):
    """<phi|psi> or <phi|H|psi> (``test/test_cumps.jl:71-101``, ``test/test_cumpo.jl:42-91``)."""
    if len(phi) != len(psi) or (H is not None and len(H) != len(psi)):
        raise _lib.DimensionMismatch(2, "inner: chains of different length")
    cplx = any(t.dtype == torch.complex128 for t in list(phi.tensors) + list(psi.tensors) +
               (list(H.tensors) if H is not None else []))
    dt = torch.complex128 if cplx else torch.float64
    if H is None:
        E = DTensor(torch.ones(1, dtype=dt, device="cuda"), (1, 1))
        for A, B in zip(phi.tensors, psi.tensors):
            if A.dims[1] != B.dims[1]:
                raise _lib.DimensionMismatch(2, "inner: site dimensions differ")
            T, _ = ops.contract(E, ("lp", "l"), B.astype(dt), ("l", "s", "r"))
            E, _ = ops.contract(A.astype(dt), ("lp", "s", "rp"), T, ("lp", "s", "r"), conj_a=True)
        v = E.numpy().reshape(())
    else:
        E = DTensor(torch.ones(1, dtype=dt, device="cuda"), (1, 1, 1))
        for A, W, B in zip(phi.tensors, H.tensors, psi.tensors):
            if A is B or (A.data.data_ptr() == B.data.data_ptr()):
                E = ops.env_update_left(E, B.astype(dt), W.astype(dt))
            else:
                T, _ = ops.contract(E, ("l", "lp", "a"), B.astype(dt), ("l", "s", "r"))
                T, _ = ops.contract(T, ("lp", "a", "s", "r"), W.astype(dt), ("a", "s", "sp", "b"))
                E, _ = ops.contract(T, ("lp", "r", "sp", "b"), A.astype(dt), ("lp", "sp", "rp"), lc=("r", "rp", "b"),
                                    conj_b=True)
        v = E.numpy().reshape(())
    return complex(v) if cplx else float(v)


def orthogonalize(psi, j):
    """``orthogonalize!(psi, j)`` -- QR sweeps towards site j (``test/test_cumps.jl:138-149,200-229``)."""
    ts = list(psi.tensors)
    N = len(ts)
    lo = min(max(psi.llim + 1, 0), j)          # sites <= llim are already left-orthogonal
    hi = max(min(psi.rlim - 1, N - 1), j)      # sites >= rlim are already right-orthogonal
    for b in range(lo, j):
        l, d, r = ts[b].dims
        Q, R = ops.qr(DTensor(ts[b].data, (l * d, r)))
        k = Q.dims[1]
        ts[b] = DTensor(Q.data, (l, d, k))
        nxt, _ = ops.contract(R, ("k", "r"), ts[b + 1], ("r", "s", "rr"))
        ts[b + 1] = nxt
    for b in range(hi, j, -1):
        l, d, r = ts[b].dims
        At = ops.permute(DTensor(ts[b].data, (l, d * r)), ("l", "x"), ("x", "l"))     # (d r) x l
        Q, R = ops.qr(At)
        k = Q.dims[1]
        ts[b] = DTensor(ops.permute(Q, ("x", "k"), ("k", "x")).data, (k, d, r))
        prv, _ = ops.contract(ts[b - 1], ("ll", "s", "l"), R, ("k", "l"))            # R[k,l]: A_prev * R^T
        ts[b - 1] = prv
    return MPS(ts, llim=j - 1, rlim=j + 1)


# ---- MPS / MPO algebra on the same kernels ([EXT] ITensors `+`, `truncate!`, `contract(::MPO, ::MPS)`;
#      the reference exercises them in test/test_cumpo.jl:42-173 and test/test_cumps.jl:196-246)
def add(psi, phi):
    """|psi> + |phi>: direct sum of the bond spaces (exact; follow with ``truncate``).  Block placement only.
    ``add(K::MPO, L::MPO)`` dispatches to ``add_mpo``."""
    if isinstance(psi, MPO) and isinstance(phi, MPO):
        return add_mpo(psi, phi)
    if not (psi.on_gpu and phi.on_gpu):
        raise _lib.TnbError(3, "add: move both MPS to the GPU with cu(); there is no CPU path")
    N = len(psi)
    if len(phi) != N:
        raise _lib.DimensionMismatch(2, "add: MPS lengths differ")
    out = []
    for j in range(N):
        A, B = psi.tensors[j], phi.tensors[j]
        l1, d, r1 = A.dims
        l2, d2, r2 = B.dims
        if d != d2:
            raise _lib.DimensionMismatch(2, "add: site dimension differs at site %d" % j)
        dt = torch.complex128 if torch.complex128 in (A.dtype, B.dtype) else torch.float64
        L = l1 if j == 0 else l1 + l2
        R = r1 if j == N - 1 else r1 + r2
        if (j == 0 and l1 != l2) or (j == N - 1 and r1 != r2):
            raise _lib.DimensionMismatch(2, "add: boundary bond dimensions differ")
        C = torch.zeros(R, d, L, dtype=dt, device=A.data.device)           # row-major (R,d,L) == column-major [L,d,R]
        lo_l, lo_r = (0 if j == 0 else l1), (0 if j == N - 1 else r1)
        C[:r1, :, :l1] = A.data.view(r1, d, l1)
        C[lo_r:lo_r + r2, :, lo_l:lo_l + l2] = B.data.view(r2, d, l2)
        out.append(DTensor(C.reshape(-1), (L, d, R)))
    return MPS(out)


def truncate(psi, maxdim=None, cutoff=None):
    """[EXT] ``truncate!(psi; maxdim, cutoff)``: orthogonalise to the last site, then split every bond by a
    truncated SVD on the way back.  Returns a right-canonical MPS (centre at site 0)."""
    if not psi.on_gpu:
        raise _lib.TnbError(3, "truncate: move psi to the GPU with cu(); there is no CPU path")
    N = len(psi)
    ts = list(orthogonalize(psi, N - 1).tensors)
    Cc = ts[N - 1]
    for j in range(N - 1, 0, -1):
        l, d, r = Cc.dims
        A, B, _ = ops.factorize_bond(DTensor(Cc.data, (l, 1, d, r)), ortho="right", which_decomp="svd", maxdim=maxdim,
                                     cutoff=cutoff or 0.0)
        k = B.dims[0]
        ts[j] = B
        Cc, _ = ops.contract(ts[j - 1], ("a", "s", "l"), DTensor(A.data, (l, k)), ("l", "k"), lc=("a", "s", "k"))
    ts[0] = Cc
    return MPS(ts, llim=-1, rlim=1)


def _mpo_as_mps(H):
    """MPO site W[a,s,s',b] viewed as an MPS site [a,(s,s'),b] (same buffer) -- what [EXT] truncate!(::MPO) works on"""
    return MPS([DTensor(W.data, (W.dims[0], W.dims[1] * W.dims[2], W.dims[3])) for W in H.tensors]), \
        [(W.dims[1], W.dims[2]) for W in H.tensors]


def truncate_mpo(H, maxdim=None, cutoff=None):
    """[EXT] ``truncate!(::MPO; maxdim, cutoff)``: the MPS algorithm on the fused site index (s, s')."""
    if not H.on_gpu:
        raise _lib.TnbError(3, "truncate_mpo: move H to the GPU with cu(); there is no CPU path")
    psi, sd = _mpo_as_mps(H)
    out = truncate(psi, maxdim=maxdim, cutoff=cutoff)
    return MPO([DTensor(t.data, (t.dims[0], d1, d2, t.dims[2])) for t, (d1, d2) in zip(out.tensors, sd)])


def contract_mpo(K, L, maxdim=None, cutoff=None):
    """[EXT] ``contract(K::MPO, L::MPO; maxdim, cutoff)`` = the operator product K*L (L acts first) as an MPO
    (``test/test_cumpo.jl:145-173``): site-wise product
    M[(aL aK), s, s'', (bL bK)] = sum_{s'} L[aL,s,s',bL] K[aK,s',s'',bK] (one contraction per site), then truncation."""
    if not (K.on_gpu and L.on_gpu):
        raise _lib.TnbError(3, "contract: move both MPOs to the GPU with cu(); there is no CPU path")
    if len(K) != len(L):
        raise _lib.DimensionMismatch(2, "contract: MPOs of different length (%d, %d)" % (len(K), len(L)))
    out = []
    for j, (Wk, Wl) in enumerate(zip(K.tensors, L.tensors)):
        if Wl.dims[2] != Wk.dims[1]:
            raise _lib.DimensionMismatch(2, "contract: site dimensions differ at site %d" % j)
        T, _ = ops.contract(Wl, ("al", "s", "t", "bl"), Wk, ("ak", "t", "u", "bk"), lc=("al", "ak", "s", "u", "bl", "bk"))
        al, ak, sdim, u, bl, bk = T.dims
        out.append(DTensor(T.data, (al * ak, sdim, u, bl * bk)))
    res = MPO(out)
    if maxdim is None and cutoff is None:
        return res
    return truncate_mpo(res, maxdim=maxdim, cutoff=cutoff)


def add_mpo(K, L):
    """[EXT] ``add(K::MPO, L::MPO)`` (``test/test_cumpo.jl:133-143``): direct sum of the bond spaces."""
    if not (K.on_gpu and L.on_gpu):
        raise _lib.TnbError(3, "add: move both MPOs to the GPU with cu(); there is no CPU path")
    a, sd = _mpo_as_mps(K)
    b, sd2 = _mpo_as_mps(L)
    if sd != sd2:
        raise _lib.DimensionMismatch(2, "add: site dimensions of the two MPOs differ")
    out = add(a, b)
    return MPO([DTensor(t.data, (t.dims[0], d1, d2, t.dims[2])) for t, (d1, d2) in zip(out.tensors, sd)])


def contract(H, psi, maxdim=None, cutoff=None):
    """[EXT] ``contract(H::MPO, psi::MPS; maxdim, cutoff)`` = H|psi> as an MPS: site-wise product
    B[(l a), s', (r b)] = sum_s W[a,s,s',b] A[l,s,r] (one contraction per site), then ``truncate``.
    ``contract(K::MPO, L::MPO)`` dispatches to ``contract_mpo``."""
    if isinstance(psi, MPO):
        return contract_mpo(H, psi, maxdim=maxdim, cutoff=cutoff)
    if not (H.on_gpu and psi.on_gpu):
        raise _lib.TnbError(3, "contract: move H and psi to the GPU with cu(); there is no CPU path")
    if len(H) != len(psi):
        raise _lib.DimensionMismatch(2, "contract: MPO and MPS lengths differ")
    out = []
    for W, A in zip(H.tensors, psi.tensors):
        T, _ = ops.contract(A, ("l", "s", "r"), W, ("a", "s", "u", "b"), lc=("l", "a", "u", "r", "b"))
        l, a, u, r, b = T.dims
        out.append(DTensor(T.data, (l * a, u, r * b)))
    res = MPS(out)
    if maxdim is None and cutoff is None:
        return res
    return truncate(res, maxdim=maxdim, cutoff=cutoff)


class Sweeps:
    """[EXT] ``Sweeps(n)`` + ``maxdim!/mindim!/cutoff!/noise!`` (``examples/dmrg.jl:20-24``)."""

    def __init__(self, nsweep, maxdim=(1,), mindim=(1,), cutoff=(0.0,), noise=(0.0,)):
        self.nsweep = int(nsweep)

        def ext(v):
            v = list(np.atleast_1d(v))
            return [v[min(i, len(v) - 1)] for i in range(self.nsweep)]
        self.maxdim = [int(x) for x in ext(maxdim)]
        self.mindim = [int(x) for x in ext(mindim)]
        self.cutoff = [float(x) for x in ext(cutoff)]
        self.noise = [float(x) for x in ext(noise)]

    def __len__(self):
        return self.nsweep


class _EnvCache:
    """Left/right environments of a sweep ([EXT] ProjMPO's LR cache filled by makeL!/makeR!).  ``store="device"``
    keeps every environment in HBM (C3: 99 x 640 MiB at chi = 4096).  ``store="host"`` keeps only the ones about
    to be used on the device: each new environment is copied to pinned host memory on a side stream and dropped,
    and the next one a sweep direction will need is prefetched one bond ahead, so the PCIe traffic (one
    environment each way per bond) hides behind the bond step.  Needed when N * w * chi^2 * 16 B exceeds HBM
    (C5: 16 GB per environment at chi = 8192, w = 30)."""

    def __init__(self, N, store="device"):
        if store not in ("device", "host"):
            raise _lib.TnbError(1, "env_store must be 'device' or 'host'")
        self.store = store
        self.dev = [None] * N          # DTensor on the device (or None)
        self.host = [None] * N         # (pinned tensor, dims) when offloaded
        self.ready = [None] * N        # event: prefetch finished
        self.side = torch.cuda.Stream() if store == "host" else None

    def put(self, j, t):
        self.dev[j] = t
        if self.store == "host" and t.size > 1:
            buf = torch.empty(t.data.shape, dtype=t.data.dtype, pin_memory=True)
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                buf.copy_(t.data, non_blocking=True)
            t.data.record_stream(self.side)
            self.host[j] = (buf, t.dims)

    def drop(self, j):
        """forget environment j entirely (it is stale)"""
        self.dev[j] = None
        self.host[j] = None
        self.ready[j] = None

    def evict(self, j):
        """keep only the host copy of environment j"""
        if self.store == "host" and self.host[j] is not None:
            self.dev[j] = None

    def prefetch(self, j):
        if self.store != "host" or j < 0 or j >= len(self.dev) or self.dev[j] is not None or self.host[j] is None:
            return
        buf, dims = self.host[j]
        with torch.cuda.stream(self.side):
            d = torch.empty(buf.shape, dtype=buf.dtype, device="cuda")
            d.copy_(buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.dev[j] = DTensor(d, dims)
        self.ready[j] = ev

    def get(self, j):
        if self.dev[j] is None:
            self.prefetch(j)
        if self.ready[j] is not None:
            torch.cuda.current_stream().wait_event(self.ready[j])
            self.dev[j].data.record_stream(torch.cuda.current_stream())
            self.ready[j] = None
        if self.dev[j] is None:
            raise _lib.TnbError(1, "environment %d is not available" % j)
        return self.dev[j]


def _dmrg_native(ts, Ws, dt, sweeps, krylovdim, maxiter, which_decomp, outputlevel, observer, checkpoint):
    """The sweep loop inside the library (``tnb_dmrg_sweep``): Python only allocates the buffers -- site tensors and one
    environment per boundary at their maxdim capacity -- and makes ONE C call per sweep."""
    import ctypes as C
    h = _lib.handle()
    N = len(ts)
    dev = ts[0].data.device
    d = [t.dims[1] for t in ts]
    w = [W.dims[0] for W in Ws] + [Ws[-1].dims[3]]
    chi = [ts[0].dims[0]] + [t.dims[2] for t in ts]
    mx = max(max(sweeps.maxdim), max(chi))
    cap = [1] + [mx] * (N - 1) + [1]                       # the eigen branch may keep up to maxdim on any inner bond
    A, E = [], []
    for j in range(N):
        buf = torch.empty(cap[j] * d[j] * cap[j + 1], dtype=dt, device=dev)
        buf[: ts[j].size].copy_(ts[j].data)
        A.append(buf)
    for t in range(N + 1):
        E.append(torch.empty(max(1, cap[t] * cap[t] * w[t]), dtype=dt, device=dev))
    vp = lambda xs: (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    i64 = lambda xs: (C.c_int64 * len(xs))(*[int(x) for x in xs])
    i32 = lambda xs: (C.c_int32 * len(xs))(*[int(x) for x in xs])
    cchi = i64(chi)
    pA, pW, pE = vp(A), vp([W.data for W in Ws]), vp(E)
    capA, capE = i64([a.numel() for a in A]), i64([e.numel() for e in E])
    nb = 2 * (N - 1)
    energy = None
    for sw in range(sweeps.nsweep):
        e, merr = C.c_double(0.0), C.c_double(0.0)
        be, bt = (C.c_double * nb)(), (C.c_double * nb)()
        h.check(h.lib.tnb_dmrg_sweep(h.h, ops._dt(A[0]), N, cchi, i32(d), i32(w), pA, capA, pW, pE, capE, 1 if sw == 0 else 0,
                                     sweeps.maxdim[sw], sweeps.mindim[sw], sweeps.cutoff[sw], sweeps.noise[sw],
                                     ops._DECOMP[which_decomp], krylovdim, maxiter, C.byref(e), C.byref(merr), be, bt,
                                     ops._stream()))
        energy = e.value
        if observer:                                        # per-bond data of the finished sweep, in sweep order
            for i in range(nb):
                b, o = (i, "left") if i < N - 1 else (nb - 1 - i, "right")
                observer(sw, b, o, be[i], bt[i])
        if outputlevel > 0:
            print("After sweep %d energy=%.12f maxlinkdim=%d maxerr=%.2E" % (sw + 1, energy, max(cchi[1:N]), merr.value))
        if checkpoint:
            from .io import save_chain
            cur = [DTensor(A[j][: cchi[j] * d[j] * cchi[j + 1]], (cchi[j], d[j], cchi[j + 1])) for j in range(N)]
            save_chain(checkpoint, MPS(cur, llim=-1, rlim=1), extra={"sweeps_done": sw + 1, "energy": energy})
    out = [DTensor(A[j][: cchi[j] * d[j] * cchi[j + 1]].clone(), (cchi[j], d[j], cchi[j + 1])) for j in range(N)]
    return energy, MPS(out, llim=-1, rlim=1)


def dmrg(H, psi0, sweeps, krylovdim=3, maxiter=1, which_decomp=None, outputlevel=0, observer=None, env_store="device",
         comm=None, shard_min_chi=256, verify_ranks=False, checkpoint=None, driver="python"):
    """``energy, psi = dmrg(H, psi0, sweeps)`` ([EXT] ITensors 0.2 two-site DMRG; reference call sites
    ``examples/dmrg.jl:25``, ``test/dmrg.jl:27,75``).  Per bond: ONE fused C call (phi = A1*A2, Lanczos with
    krylovdim matvecs, optional noise term, truncated factorization) plus one environment update.
    ``env_store="host"`` spills the environment cache to pinned host memory (see ``_EnvCache``).

    ``comm`` (a ``shard.ShardComm``; one process per GPU, every rank calls dmrg with the same arguments): the
    matvecs, the noise term's big contractions and the environment updates of every bond whose left bond dimension
    is divisible by the number of ranks and >= ``shard_min_chi`` are sharded over the output bond (``shard.ShardedSweep``);
    the factorization is replicated.  Energies and the MPS are bit-identical on every rank and equal to the 1-GPU
    sweep's up to the summation order of the sharded GEMMs (tests: 1e-12).  ``verify_ranks`` cross-checks
    (energy, n_keep) over the ranks after every bond.

    ``driver="native"``: the whole sweep loop runs inside the library (``tnb_dmrg_sweep``, one C call per sweep; site
    tensors and one environment per boundary are allocated once at their maxdim capacity).  Same kernels in the same
    order, so energies are bit-identical to the Python loop; the observer is then called after each sweep with that
    sweep's per-bond data.  Not combined with ``comm`` / ``env_store="host"``.

    ``checkpoint`` (a path): after every sweep the MPS (orthogonality centre back at site 0) is written with
    ``io.save_chain`` together with {"sweeps_done", "energy"}; ``load_chain`` + ``dmrg`` with the remaining sweeps
    resumes bit-identically (rank 0 writes when sharded)."""
    if not (H.on_gpu and psi0.on_gpu):
        raise _lib.TnbError(3, "dmrg: move H and psi0 to the GPU with cu(); there is no CPU path")
    N = len(psi0)
    if len(H) != N:
        raise _lib.DimensionMismatch(2, "dmrg: MPO and MPS lengths differ")
    psi = orthogonalize(psi0, 0) if psi0.center != 0 else MPS(list(psi0.tensors), -1, 1)
    ts = list(psi.tensors)
    Ws = H.tensors
    dt = torch.complex128 if any(t.dtype == torch.complex128 for t in ts + list(Ws)) else torch.float64
    ts = [t.astype(dt) for t in ts]
    Ws = [w.astype(dt) for w in Ws]
    if driver not in ("python", "native"):
        raise _lib.TnbError(1, "dmrg: driver must be 'python' or 'native'")
    if driver == "native":
        if (comm is not None and comm.world > 1) or env_store != "device":
            raise _lib.TnbError(3, "dmrg: driver='native' runs on one GPU with the environments on the device")
        return _dmrg_native(ts, Ws, dt, sweeps, krylovdim, maxiter, which_decomp, outputlevel, observer, checkpoint)
    one = DTensor(torch.ones(1, dtype=dt, device="cuda"), (1, 1, 1))
    sh = None
    if comm is not None and comm.world > 1:
        from .shard import ShardedSweep
        chi_max = max([max(sweeps.maxdim)] + [t.dims[2] for t in ts])
        sh = ShardedSweep(comm, dt, chi_max, max(t.dims[1] for t in ts), max(max(w.dims[0], w.dims[3]) for w in Ws),
                          min_chi=shard_min_chi)
    env_l = sh.env_left if sh else ops.env_update_left
    env_r = sh.env_right if sh else ops.env_update_right

    def bond_step(L, W1, W2, R, A1, A2, ortho, **kw):
        if sh is None:
            return ops.dmrg_bond_step(L, W1, W2, R, A1, A2, ortho, **kw)
        out = sh.bond_step(L, W1, W2, R, A1, A2, ortho, **kw)
        if verify_ranks:
            import torch.distributed as dist
            seen = [None] * comm.world
            dist.all_gather_object(seen, (out[0], out[1].dims, out[3]), group=comm.group)
            if any(x != seen[0] for x in seen):
                raise _lib.TnbError(5, "dmrg: ranks diverged at a bond step: %r" % (seen,))
        return out

    Rs = _EnvCache(N, env_store)
    Ls = _EnvCache(N, env_store)
    Rs.put(N - 1, one)
    for j in range(N - 1, 1, -1):
        Rs.put(j - 1, env_r(Rs.get(j), ts[j], Ws[j]))
        if j < N - 1:
            Rs.evict(j)
    Ls.put(0, one)
    energy = None
    for sw in range(sweeps.nsweep):
        kw = dict(maxdim=sweeps.maxdim[sw], mindim=sweeps.mindim[sw], cutoff=sweeps.cutoff[sw],
                  noise=sweeps.noise[sw], krylovdim=krylovdim, maxiter=maxiter, which_decomp=which_decomp)
        maxerr = 0.0
        for b in range(0, N - 1):
            Rs.prefetch(b + 2)                      # the next bond's right environment, one bond ahead
            energy, ts[b], ts[b + 1], err = bond_step(Ls.get(b), Ws[b], Ws[b + 1], Rs.get(b + 1), ts[b], ts[b + 1], "left", **kw)
            Ls.put(b + 1, env_l(Ls.get(b), ts[b], Ws[b]))
            if b + 1 < N - 1:
                Rs.drop(b + 1)        # stale now (site b+1 changed): one environment per bond stays alive
            if b > 0:
                Ls.evict(b)
            maxerr = max(maxerr, err)
            if observer:
                observer(sw, b, "left", energy, err)
        for b in range(N - 2, -1, -1):
            Ls.prefetch(b - 1)
            energy, ts[b], ts[b + 1], err = bond_step(Ls.get(b), Ws[b], Ws[b + 1], Rs.get(b + 1), ts[b], ts[b + 1], "right", **kw)
            Rs.put(b, env_r(Rs.get(b + 1), ts[b + 1], Ws[b + 1]))
            if b > 0:
                Ls.drop(b)
            if b + 1 < N - 1:
                Rs.evict(b + 1)
            maxerr = max(maxerr, err)
            if observer:
                observer(sw, b, "right", energy, err)
        if outputlevel > 0:
            print("After sweep %d energy=%.12f maxlinkdim=%d maxerr=%.2E" %
                  (sw + 1, energy, max(t.dims[2] for t in ts[:-1]), maxerr))
        if checkpoint and (comm is None or comm.rank == 0):
            from .io import save_chain
            save_chain(checkpoint, MPS(ts, llim=-1, rlim=1), extra={"sweeps_done": sw + 1, "energy": energy})
    out = MPS(ts, llim=-1, rlim=1)
    if sh is not None:
        out.shard_stats = {"sharded_bond_steps": sh.sharded_steps, "replicated_bond_steps": sh.replicated_steps}
    return energy, out


def apply(gates, psi, cutoff=None, maxdim=None):
    """``apply(gates, psi; cutoff, maxdim)`` ([EXT] ``product``; ``examples/gate_evolution.jl:46``).
    gates: list of (G, n): one-site ``G[s',s]`` at site n, or two-site ``G[s1',s2',s1,s2]`` on (n, n+1)."""
    if not psi.on_gpu:
        raise _lib.TnbError(3, "apply: move psi to the GPU with cu(); there is no CPU path")
    cur = psi
    for G, n in gates:
        G = _to_dev(G)
        c = cur.center
        if c != n:
            cur = orthogonalize(cur, n)
        ts = list(cur.tensors)
        if len(G.dims) == 2:
            A, _ = ops.contract(ts[n], ("l", "s", "r"), G, ("sp", "s"), lc=("l", "sp", "r"))
            ts[n] = A
            cur = MPS(ts, llim=n - 1, rlim=n + 1)
        elif len(G.dims) == 4:
            if n + 1 >= len(ts):
                raise _lib.TnbError(1, "apply: two-site gate at the last site")
            A1, A2, _ = ops.tebd_apply_gate(G, ts[n], ts[n + 1], maxdim=maxdim, cutoff=cutoff or 0.0)
            ts[n], ts[n + 1] = A1, A2
            cur = MPS(ts, llim=n, rlim=n + 2)
        else:
            raise _lib.TnbError(3, "apply: only one- and two-site gates")
    return cur


# ---- host-side model builders (inputs are made on the host and uploaded, like MPO(ampo, sites) then cu())
def spin_ops(S):
    d = int(round(2 * S + 1))
    m = S - np.arange(d)
    Sz = np.diag(m)
    Sp = np.zeros((d, d))
    for i in range(1, d):
        Sp[i - 1, i] = np.sqrt(S * (S + 1) - m[i] * (m[i] + 1))
    return dict(Sz=Sz, Sp=Sp, Sm=Sp.T.copy(), Sx=0.5 * (Sp + Sp.T), Id=np.eye(d), d=d)


def _finish(Wb, N):
    out = []
    for j in range(N):
        W = Wb
        if j == 0:
            W = W[-1:, :, :, :]
        if j == N - 1:
            W = W[:, :, :, :1]
        out.append(np.ascontiguousarray(W))
    return MPO(out)


def heisenberg_mpo(N, S=0.5):
    """sum_j Sz Sz + 1/2 (S+ S- + S- S+) (``examples/dmrg.jl:9-15``), MPO bond 5, W[a,s,s',b] = <s'|O|s>."""
    o = spin_ops(S)
    d = o["d"]
    W = np.zeros((5, d, d, 5))
    for (a, b, O) in [(0, 0, o["Id"]), (1, 0, o["Sm"]), (2, 0, o["Sp"]), (3, 0, o["Sz"]), (4, 1, 0.5 * o["Sp"]),
                      (4, 2, 0.5 * o["Sm"]), (4, 3, o["Sz"]), (4, 4, o["Id"])]:
        W[a, :, :, b] += O.T
    return _finish(W, N)


def tfim_mpo(N, J=-1.0, h=-0.5):
    """J sum Sz Sz + h sum Sx (``test/dmrg.jl:63-67``), MPO bond 3."""
    o = spin_ops(0.5)
    W = np.zeros((3, 2, 2, 3))
    for (a, b, O) in [(0, 0, o["Id"]), (1, 0, o["Sz"]), (2, 0, h * o["Sx"]), (2, 1, J * o["Sz"]), (2, 2, o["Id"])]:
        W[a, :, :, b] += O.T
    return _finish(W, N)
