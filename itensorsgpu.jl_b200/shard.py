"""Multi-GPU plumbing for the sharded H_eff*phi (one process per GPU, torch.distributed).

Partitioning ("load-balanced alternative"): the OUTPUT bond l' of H_eff is split
into ``world`` contiguous slabs.  Rank g keeps the slab ``L[:, l'_g, :]`` of the left environment (it
never moves), the full right environment and MPO tensors (replicated), and the full Krylov vector.
Every contraction of the matvec then does exactly 1/world of the work with NO reduction; the slabs of
H*phi are all-gathered (NCCL over NVLink) when the next matvec needs the full vector.  The reference
has no live multi-GPU path (dead cuBLASMg code: ``src/tensor/dense.jl:195-265``).

Everything here is layout arithmetic on flat column-major buffers, so it is testable on CPU with the
gloo backend; the compute call is ``ops.heff_apply_shard``.
"""
import torch


def slab_range(chi, rank, world):
    if chi % world:
        raise ValueError("bond dimension %d is not divisible by %d ranks" % (chi, world))
    n = chi // world
    return rank * n, (rank + 1) * n


def left_env_slab(L_flat, chi, w, rank, world):
    """flat column-major L[l, l', a]  ->  flat column-major L[l, l'_slab, a] (contiguous copy)."""
    lo, hi = slab_range(chi, rank, world)
    return L_flat.view(w, chi, chi)[:, lo:hi, :].contiguous().reshape(-1)


def assemble_gathered(gathered, chi, d1, d2, chiR, world):
    """all_gather_into_tensor output (rank-major slabs out_g[l'_slab, s1, s2, r]) -> flat column-major
    full vector Hphi[l', s1, s2, r]."""
    n = chi // world
    g = gathered.view(world, chiR * d2 * d1, n)          # [rank][(r,s2,s1)][l'_slab]  (row-major view of F buffers)
    return g.permute(1, 0, 2).contiguous().reshape(-1)    # [(r,s2,s1)][rank][l'_slab] == F-order (l', s1, s2, r)


def heff_apply_sharded(ops, L_slab, W1, W2, R, phi, group=None):
    """One sharded matvec: local slab compute + all-gather.  Returns the full flat H*phi."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    slab = ops.heff_apply_shard(L_slab, W1, W2, R, phi)
    gathered = torch.empty(world * slab.data.numel(), dtype=slab.data.dtype, device=slab.data.device)
    dist.all_gather_into_tensor(gathered, slab.data, group=group)
    cl, d1, d2, cr = phi.dims
    return assemble_gathered(gathered, cl, d1, d2, cr, world)


# ---------------------------------------------------------------------------------------------
# Fused GEMM + all-gather over NVLink peer memory (tnb_heff_apply_shard_fused)
# ---------------------------------------------------------------------------------------------
class _RawCuda:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerBuffers:
    """One buffer of ``nbytes`` per rank, each mapped into every process of the group (CUDA IPC): ``ptrs[g]`` is
    rank g's buffer as seen from this process.  The 64-byte IPC handles travel through the host-side process
    group; after that no library collective touches the data path."""

    def __init__(self, nbytes, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        self.h = _lib.handle()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes)
        own = C.c_void_p()
        hbuf = C.create_string_buffer(64)
        self.h.check(self.h.lib.tnb_peer_alloc(self.h.h, self.nbytes, C.byref(own), hbuf))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(hbuf.raw), group=group)
        self.ptrs = []
        for g, hb in enumerate(handles):
            if g == self.rank:
                self.ptrs.append(own.value)
            else:
                p = C.c_void_p()
                self.h.check(self.h.lib.tnb_peer_open(self.h.h, hb, C.byref(p)))
                self.ptrs.append(p.value)
        self._arr = (C.c_void_p * self.world)(*self.ptrs)
        dist.barrier(group=group)

    def c_array(self):
        return self._arr

    def local(self, dtype=torch.float64):
        """torch view of this rank's own buffer."""
        return torch.as_tensor(_RawCuda(self.ptrs[self.rank], self.nbytes), device="cuda").view(dtype)

    def close(self):
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for g, p in enumerate(self.ptrs):
            if g != self.rank:
                self.h.lib.tnb_peer_close(self.h.h, p)
        dist.barrier(group=self.group)
        self.h.lib.tnb_peer_free(self.h.h, self.ptrs[self.rank])
        self.ptrs = []


class FusedShardedHeff:
    """H_eff*phi with the output bond sharded over the ranks and the all-gather fused into the last GEMM:
    every rank ends up with the FULL H*phi in its own ``out`` buffer (same layout as phi), written tile by tile by
    all ranks through NVLink peer stores; a device-side flag barrier replaces the collective."""

    def __init__(self, phi_dims, dtype=torch.float64, group=None, nbuf=2):
        n = 1
        for d in phi_dims:
            n *= int(d)
        es = 16 if dtype == torch.complex128 else 8
        self.dims, self.dtype = tuple(phi_dims), dtype
        self.outs = [PeerBuffers(n * es, group) for _ in range(nbuf)]
        self.flags = PeerBuffers(8 * 8, group)
        self.epoch = 0
        self.rank, self.world = self.flags.rank, self.flags.world

    def apply(self, Lslab, W1, W2, R, phi):
        """Returns a DTensor view of this rank's full-vector buffer holding H*phi (valid for work queued behind the
        call on the current stream; buffers alternate between calls)."""
        import ctypes as C
        from . import ops
        from .ops import DTensor, BondDims
        h = self.flags.h
        cl, d1, d2, cr = phi.dims
        clp = Lslab.dims[1]
        bd = BondDims(cl, cr, d1, d2, W1.dims[0], W1.dims[3], W2.dims[3])
        buf = self.outs[self.epoch % len(self.outs)]
        self.epoch += 1
        h.check(h.lib.tnb_heff_apply_shard_fused(h.h, ops._dt(phi.data), C.byref(bd), self.rank, self.world, clp,
                                                 ops._ptr(Lslab.data), ops._ptr(W1.data), ops._ptr(W2.data), ops._ptr(R.data),
                                                 ops._ptr(phi.data), buf.c_array(), self.flags.c_array(),
                                                 C.c_uint64(self.epoch), ops._stream()))
        return DTensor(buf.local(self.dtype), self.dims)

    def status(self):
        h = self.flags.h
        from . import ops
        h.check(h.lib.tnb_peer_status(h.h, ops._stream()))

    def close(self):
        for b in self.outs:
            b.close()
        self.flags.close()


# ---------------------------------------------------------------------------------------------
# MPO-bond split (the north star's plan) -- kept for comparison with the l' split
# ---------------------------------------------------------------------------------------------
def mpo_split_ranges(w, world):
    """Contiguous, as-even-as-possible split of an MPO bond of dimension w over the ranks (some may be empty)."""
    base, rem = divmod(w, world)
    out, lo = [], 0
    for g in range(world):
        n = base + (1 if g < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


class MpoSplitHeff:
    """H*phi = sum_{a,c} L[a] (What[a,c] o phi) R[c] with rank g owning L[:,:,a in A_g] and R[:,:,c in C_g]
    (environments never move).  (1) Y_g = phi*L[A_g] (compute-bound, local); (2,3) partial Z_g[c] for ALL c from
    a in A_g (HBM-bound, local); reduce the c-planes to their owners (payload w d^2 chi^2 elements); (4) partial
    H*phi_g from c in C_g; all-reduce H*phi.  Parallelism is capped by w (w = 5: ideal 1.67x / 2.5x / 5x on
    2 / 4 / 8 GPUs), which is why the output-bond split above is the default."""

    def __init__(self, L, W1, W2, R, group=None):
        import torch.distributed as dist
        from .ops import DTensor
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        cl, _, wl = L.dims
        cr, _, wr = R.dims
        self.ar = mpo_split_ranges(wl, self.world)
        self.cr_ = mpo_split_ranges(wr, self.world)
        a0, a1 = self.ar[self.rank]
        c0, c1 = self.cr_[self.rank]
        # a and c are the slowest modes of L[l,l',a] / R[r,r',c]: the owned slabs are contiguous views
        self.Ls = DTensor(L.data[a0 * cl * cl: a1 * cl * cl], (cl, cl, a1 - a0)) if a1 > a0 else None
        self.Rs = DTensor(R.data[c0 * cr * cr: c1 * cr * cr], (cr, cr, c1 - c0)) if c1 > c0 else None
        # W1[a,s,s',b] restricted to a in A_g (a is the FASTEST mode: small strided gather, done once)
        w1 = W1.data.view(W1.dims[3], W1.dims[2], W1.dims[1], W1.dims[0])[..., a0:a1].contiguous()
        self.W1s = DTensor(w1.reshape(-1), (a1 - a0, W1.dims[1], W1.dims[2], W1.dims[3])) if a1 > a0 else None
        self.W2 = W2
        self.wr = wr

    def apply(self, phi):
        """Returns the full H*phi (flat DTensor, phi's layout) on every rank."""
        from . import ops
        from .ops import DTensor
        dist = self.dist
        cl, d1, d2, cr = phi.dims
        npl = cr * cl * d1 * d2                                   # elements of one c-plane of Z[r,l',s1',s2',c]
        Z = torch.zeros(npl * self.wr, dtype=phi.dtype, device=phi.data.device)
        if self.Ls is not None:
            Y, _ = ops.contract(phi, ("l", "s1", "s2", "r"), self.Ls, ("l", "lp", "a"), lc=("s1", "s2", "r", "lp", "a"))
            T2, _ = ops.contract(Y, ("s1", "s2", "r", "lp", "a"), self.W1s, ("a", "s1", "s1p", "b"),
                                 lc=("s2", "r", "lp", "s1p", "b"))
            ops.contract(T2, ("s2", "r", "lp", "s1p", "b"), self.W2, ("b", "s2", "s2p", "c"),
                         lc=("r", "lp", "s1p", "s2p", "c"), out=DTensor(Z, (cr, cl, d1, d2, self.wr)))
        # reduce every c-chunk to its owner (unequal chunk sizes: one reduce per destination)
        staged = dist.get_backend(self.group) == "gloo"     # gloo (ranks sharing a GPU in the tests) reduces on the host
        work = []
        for g, (c0, c1) in enumerate(self.cr_):
            if c1 > c0:
                if staged:
                    hz = Z[c0 * npl: c1 * npl].cpu()
                    hz = torch.view_as_real(hz) if hz.is_complex() else hz
                    dist.reduce(hz, dst=g, group=self.group)
                    if g == self.rank:
                        Z[c0 * npl: c1 * npl].copy_(torch.view_as_complex(hz) if Z.is_complex() else hz)
                else:
                    work.append(dist.reduce(Z[c0 * npl: c1 * npl], dst=g, group=self.group, async_op=True))
        for wk in work:
            wk.wait()
        out = torch.zeros(cl * d1 * d2 * cr, dtype=phi.dtype, device=phi.data.device)
        if self.Rs is not None:
            c0, c1 = self.cr_[self.rank]
            Zs = DTensor(Z[c0 * npl: c1 * npl], (cr, cl, d1, d2, c1 - c0))
            ops.contract(Zs, ("r", "lp", "s1p", "s2p", "c"), self.Rs, ("r", "rp", "c"), lc=("lp", "s1p", "s2p", "rp"),
                         out=DTensor(out, (cl, d1, d2, cr)))
        if staged:
            ho = out.cpu()
            hr = torch.view_as_real(ho) if ho.is_complex() else ho
            dist.all_reduce(hr, group=self.group)
            out.copy_(ho)
        else:
            dist.all_reduce(out, group=self.group)
        return DTensor(out, (cl, d1, d2, cr))


# ---------------------------------------------------------------------------------------------
# Peer group on the handle (tnb_comm_init) and the sharded DMRG sweep built on it
# ---------------------------------------------------------------------------------------------
class ShardComm:
    """The peer group of this process (one process per GPU; ``torch.distributed`` is used only to exchange the
    64-byte IPC handles and for host-side barriers -- gloo or nccl).  Registers the peer-mapped flag arrays on the
    library handle; after that every collective step of the sharded path is a device-side barrier, a peer store
    from a GEMM epilogue or a copy-engine write into a peer-mapped buffer."""

    def __init__(self, group=None):
        import torch.distributed as dist
        from . import _lib
        self.h = _lib.handle()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.flags = PeerBuffers(256, group)
        self.h.check(self.h.lib.tnb_comm_init(self.h.h, self.rank, self.world, self.flags.c_array()))
        self._bufs = {}

    def buffer(self, name, nbytes):
        """Peer-mapped buffer set ``name`` of at least ``nbytes`` per rank (collective: every rank must ask for the
        same names and sizes in the same order)."""
        b = self._bufs.get(name)
        if b is not None and b.nbytes >= nbytes:
            return b
        if b is not None:
            b.close()
        b = PeerBuffers(int(nbytes), self.group)
        self._bufs[name] = b
        return b

    def barrier(self):
        from . import ops
        self.h.check(self.h.lib.tnb_comm_barrier(self.h.h, ops._stream()))

    def allgather(self, bufs, nbytes_per_rank, offset=0):
        from . import ops
        self.h.check(self.h.lib.tnb_comm_allgather(self.h.h, bufs.c_array(), int(offset), int(nbytes_per_rank), ops._stream()))

    def status(self):
        from . import ops
        self.h.check(self.h.lib.tnb_peer_status(self.h.h, ops._stream()))

    def close(self):
        torch.cuda.synchronize()
        for b in self._bufs.values():
            b.close()
        self._bufs = {}
        self.h.lib.tnb_comm_finalize(self.h.h)
        self.flags.close()


class ShardedHeffHost:
    """End-to-end sharded H_eff*phi with HOST buffers (``tnb_heff_apply_shard_host``): each rank moves 1/world of phi
    up and 1/world of H*phi down over PCIe; the rest crosses NVLink."""

    def __init__(self, comm, phi_dims, dtype=torch.float64):
        n = 1
        for d in phi_dims:
            n *= int(d)
        es = 16 if dtype == torch.complex128 else 8
        self.comm, self.dims, self.dtype = comm, tuple(phi_dims), dtype
        self.phis = comm.buffer("e2e_phi", n * es)
        self.outs = comm.buffer("e2e_out", n * es)

    def apply_host(self, Lslab, W1, W2, R, phi_host, out_host):
        """phi_host / out_host: pinned CPU tensors in phi's layout; only this rank's r-chunk of phi_host is read and
        only its own l' slab of out_host (``slab_range(chiL, rank, world)`` of the fastest mode) is written.  Synchronous."""
        import ctypes as C
        from . import ops
        from .ops import BondDims
        h = self.comm.h
        cl, d1, d2, cr = self.dims
        bd = BondDims(cl, cr, d1, d2, W1.dims[0], W1.dims[3], W2.dims[3])
        h.check(h.lib.tnb_heff_apply_shard_host(h.h, ops._dt(Lslab.data), C.byref(bd), ops._ptr(Lslab.data), ops._ptr(W1.data),
                                                ops._ptr(W2.data), ops._ptr(R.data), ops._ptr(phi_host), self.phis.c_array(),
                                                self.outs.c_array(), ops._ptr(out_host), ops._stream()))
        return out_host

    def chunk_range(self):
        """[r0, r1): the slice of the slowest mode r this rank moves over PCIe"""
        cr, w, g = self.dims[3], self.comm.world, self.comm.rank
        rc = (cr + w - 1) // w
        r0 = min(cr, g * rc)
        return r0, min(cr, r0 + rc)

    def device_result(self):
        """DTensor view of the full H*phi in this rank's own result buffer (after apply_host)."""
        from .ops import DTensor
        n = 1
        for d in self.dims:
            n *= d
        return DTensor(self.outs.local(self.dtype)[:n], self.dims)


def _env_kind(E):
    return "slab" if E.dims[0] != E.dims[1] else "full"


class ShardedSweep:
    """Per-bond operations of a multi-GPU two-site DMRG sweep (used by ``mps.dmrg(..., comm=...)``).

    Left environments of bonds whose dimension is divisible by the number of ranks (and at least ``min_chi``) are
    kept as l' slabs -- 1/world of the memory, never moved; smaller ones and all right environments are replicated.
    On a sharded bond the three Lanczos matvecs, the noise term's big contractions and both environment updates run
    at 1/world of the flops per rank; the truncated factorization is replicated (deterministic kernels, identical
    inputs), so the MPS stays bit-identical on every rank without a broadcast."""

    def __init__(self, comm, dtype, chi_max, d, w, min_chi=256):
        self.comm, self.dtype = comm, dtype
        self.world, self.rank = comm.world, comm.rank
        self.min_chi = max(int(min_chi), self.world)
        es = 16 if dtype == torch.complex128 else 8
        h = comm.h
        n = int(chi_max) * int(chi_max) * d * d
        self.out_a = comm.buffer("lanczos_a", n * es)
        self.out_b = comm.buffer("lanczos_b", n * es)
        self.stage = comm.buffer("stage", int(h.lib.tnb_shard_stage_bytes(1 if es == 16 else 0, int(chi_max), d, w, self.world)))
        self.sharded_steps = 0
        self.replicated_steps = 0

    def shardable(self, chi):
        return chi % self.world == 0 and chi >= self.min_chi

    # ---- layout conversions (rare: only where the chain's bond dimension crosses min_chi)
    def to_slab(self, L):
        from .ops import DTensor
        if _env_kind(L) == "slab":
            return L
        cl, _, w = L.dims
        lo, hi = slab_range(cl, self.rank, self.world)
        return DTensor(left_env_slab(L.data, cl, w, self.rank, self.world), (cl, hi - lo, w))

    def to_full(self, Ls):
        """all-gather the l' slabs of a left environment into the replicated full tensor"""
        from .ops import DTensor
        if _env_kind(Ls) == "full":
            return Ls
        cl, clp, w = Ls.dims
        es = Ls.data.element_size()
        nb = Ls.data.numel() * es
        mine = self.stage.local(Ls.dtype)
        mine[self.rank * Ls.data.numel(): (self.rank + 1) * Ls.data.numel()].copy_(Ls.data)
        self.comm.allgather(self.stage, nb)
        g = mine[: self.world * Ls.data.numel()].view(self.world, w, clp, cl)           # [g][a][l'_s][l]
        return DTensor(g.permute(1, 0, 2, 3).contiguous().reshape(-1), (cl, cl, w))      # [a][(g,l'_s)][l]

    # ---- one bond
    def bond_step(self, L, W1, W2, R, A1, A2, ortho, maxdim, mindim, cutoff, noise, krylovdim, maxiter, which_decomp):
        import ctypes as C
        from . import ops
        from .ops import DTensor, BondDims
        cl, d1, cm = A1.dims
        _, d2, cr = A2.dims
        sharded = _env_kind(L) == "slab" and (noise <= 0 or ortho == "left" or cr % self.world == 0)
        if not sharded:
            self.replicated_steps += 1
            return ops.dmrg_bond_step(self.to_full(L), W1, W2, R, A1, A2, ortho, maxdim=maxdim, mindim=mindim, cutoff=cutoff,
                                      noise=noise, krylovdim=krylovdim, maxiter=maxiter, which_decomp=which_decomp)
        self.sharded_steps += 1
        h = self.comm.h
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
        h.check(h.lib.tnb_dmrg_bond_step_shard(h.h, ops._dt(A1.data), C.byref(bd), cm, ops._ptr(L.data), ops._ptr(W1.data),
                                               ops._ptr(W2.data), ops._ptr(R.data), ops._ptr(b1), ops._ptr(b2),
                                               0 if ortho == "left" else 1, ops._DECOMP[which_decomp], int(maxdim), int(mindim),
                                               float(cutoff or 0.0), float(noise), int(krylovdim), int(maxiter),
                                               self.out_a.c_array(), self.out_b.c_array(), self.stage.c_array(),
                                               C.byref(e), C.byref(nk), C.byref(err), ops._stream()))
        k = nk.value
        return e.value, DTensor(b1[: m * k], (cl, d1, k)), DTensor(b2[: k * n], (k, d2, cr)), err.value

    def env_left(self, L, A, W):
        """new left environment for the bond to the right of site tensor A (slab if that bond is shardable)"""
        from . import ops
        from .ops import DTensor
        cl, d, cr = A.dims
        want_slab = self.shardable(cr)
        if want_slab and _env_kind(L) == "slab":
            h = self.comm.h
            out = DTensor.empty((cr, cr // self.world, W.dims[3]), A.dtype, A.data.device)
            h.check(h.lib.tnb_env_update_left_shard(h.h, ops._dt(A.data), cl, cr, d, W.dims[0], W.dims[3], ops._ptr(L.data),
                                                    ops._ptr(A.data), ops._ptr(W.data), self.stage.c_array(),
                                                    ops._ptr(out.data), ops._stream()))
            return out
        full = ops.env_update_left(self.to_full(L), A, W)
        return self.to_slab(full) if want_slab else full

    def env_right(self, R, A, W):
        """new (replicated) right environment for the bond to the left of site tensor A"""
        from . import ops
        from .ops import DTensor
        cl, d, cr = A.dims
        if self.shardable(cl) and self.shardable(cr):
            h = self.comm.h
            out = DTensor.empty((cl, cl, W.dims[0]), A.dtype, A.data.device)
            h.check(h.lib.tnb_env_update_right_shard(h.h, ops._dt(A.data), cl, cr, d, W.dims[0], W.dims[3], ops._ptr(R.data),
                                                     ops._ptr(A.data), ops._ptr(W.data), self.stage.c_array(),
                                                     ops._ptr(out.data), ops._stream()))
            return out
        return ops.env_update_right(R, A, W)
