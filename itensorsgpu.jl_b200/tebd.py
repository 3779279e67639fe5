"""TEBD even/odd gate layers in B form, on one GPU or spread over the GPUs of a node.

[EXT] ITensors ``apply(gates, psi; cutoff, maxdim)`` (reference call site ``examples/gate_evolution.jl:46``)
applies gates one after the other while moving the orthogonality centre.  In B form -- right-canonical site
tensors in the Schmidt bases plus the Schmidt values ``lam[j]`` of every bond (bond j = left of site j) -- a gate
on (n, n+1) touches only ``B[n], B[n+1], lam[n]`` and produces ``lam[n+1]``, so all gates of one layer are
independent.  ``B[0..N-1]`` is at all times a right-canonical MPS of the state, so the
result is directly an ``MPS`` with ``llim=-1, rlim=1``.

Multi-GPU (``ShardedTEBD``): contiguous blocks of an even number of sites per rank, one process per GPU.
Even layer: every gate is local.  Odd layer: the one bond straddling two blocks needs the right neighbour's
first site tensor -- a point-to-point halo exchange (``dist.isend`` / ``dist.recv``; NCCL over NVLink on GPUs,
gloo in the CPU plumbing test) of one site tensor each way; there is no global collective.
The compute call is injectable so that the plumbing can be exercised without a GPU (tests/test_tebd_gloo.py).
"""
import torch

from . import _lib, ops
from .mps import MPS, orthogonalize
from .ops import DTensor


class BState:
    """``Bs[j]``: DTensor [chi_j, d, chi_{j+1}];  ``lams[j]``: 1-D float64 tensor, Schmidt values of the bond left of
    site ``first + j`` (``len(lams) == len(Bs) + 1``); ``first``: global index of ``Bs[0]`` (0 unless sharded)."""

    def __init__(self, Bs, lams, first=0):
        if len(lams) != len(Bs) + 1:
            raise _lib.DimensionMismatch(2, "BState: %d tensors need %d Schmidt vectors" % (len(Bs), len(Bs) + 1))
        self.Bs, self.lams, self.first = list(Bs), list(lams), first

    def __len__(self):
        return len(self.Bs)

    def mps(self):
        return MPS(list(self.Bs), llim=-1, rlim=1)

    def maxlinkdim(self):
        return max(int(l.numel()) for l in self.lams)


def canonical_form(psi):
    """MPS (any gauge, on the GPU) -> BState.  QR sweep to the last site (``orthogonalize``), then one SVD sweep
    back: the SVD at the orthogonality centre IS the Schmidt decomposition, so no division by Schmidt values."""
    if not psi.on_gpu:
        raise _lib.TnbError(3, "canonical_form: move psi to the GPU with cu(); there is no CPU path")
    N = len(psi)
    left = orthogonalize(psi, N - 1)
    ts = list(left.tensors)
    Bs = [None] * N
    dev = ts[0].data.device
    lams = [None] * (N + 1)
    lams[0] = torch.ones(1, dtype=torch.float64, device=dev)
    lams[N] = torch.ones(1, dtype=torch.float64, device=dev)
    Cc = ts[N - 1]
    for j in range(N - 1, 0, -1):
        l, d, r = Cc.dims
        A, B, _ = ops.factorize_bond(DTensor(Cc.data, (l, 1, d, r)), ortho="right", which_decomp="svd", cutoff=0.0,
                                     normalize=False)
        k = B.dims[0]
        Bs[j] = B
        us = A.data.view(k, l)                                   # column-major (l, k): row i of the view = column i
        s = torch.linalg.vector_norm(us, dim=1)
        lams[j] = (s / torch.linalg.vector_norm(s)).to(torch.float64)
        Cc, _ = ops.contract(ts[j - 1], ("a", "s", "l"), DTensor(A.data, (l, k)), ("l", "k"), lc=("a", "s", "k"))
    nrm = ops.norm(Cc)
    Bs[0] = ops.scale(Cc, 1.0 / nrm)
    return BState(Bs, lams)


def tebd_layer(state, G, parity, maxdim=None, cutoff=0.0, gate_fn=None):
    """Apply the uniform two-site gate G[s1',s2',s1,s2] on every bond (n, n+1) with GLOBAL n of the given parity
    inside this state's block.  Returns the largest truncation error."""
    gate_fn = gate_fn or ops.tebd_gate_bform
    worst = 0.0
    n0 = 0 if (state.first % 2) == parity else 1
    for j in range(n0, len(state) - 1, 2):
        B1, B2, lam, err = gate_fn(G, state.lams[j], state.Bs[j], state.Bs[j + 1], maxdim=maxdim, cutoff=cutoff)
        state.Bs[j], state.Bs[j + 1], state.lams[j + 1] = B1, B2, lam
        worst = max(worst, err)
    return worst


def block_range(N, rank, world):
    """Contiguous block of an even number of sites per rank (the last rank takes the remainder)."""
    per = (N // world) & ~1
    if per < 2:
        raise ValueError("%d sites cannot give every one of %d ranks an even block of >= 2 sites" % (N, world))
    lo = rank * per
    hi = N if rank == world - 1 else lo + per
    return lo, hi


class ShardedTEBD:
    """One rank's block of a distributed B-form chain."""

    def __init__(self, state, N, group=None, gate_fn=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.N = N
        self.state = state
        self.gate_fn = gate_fn or ops.tebd_gate_bform
        # gloo moves host memory only (CPU plumbing test; ranks sharing one GPU in the -m gpu tests): stage through the host
        self.staged = dist.get_backend(group) == "gloo"
        lo, hi = block_range(N, self.rank, self.world)
        if state.first != lo or len(state) != hi - lo:
            raise _lib.DimensionMismatch(2, "rank %d holds sites [%d,%d), expected [%d,%d)" %
                                         (self.rank, state.first, state.first + len(state), lo, hi))

    @staticmethod
    def scatter_from(full, N, group=None, gate_fn=None):
        """Every rank passes the same full BState (e.g. canonicalised on each rank from the same input); each keeps
        its own block."""
        import torch.distributed as dist
        lo, hi = block_range(N, dist.get_rank(group), dist.get_world_size(group))
        return ShardedTEBD(BState(full.Bs[lo:hi], full.lams[lo:hi + 1], first=lo), N, group, gate_fn)

    # ---- halo exchange of one site tensor (dims header + payload)
    def _send_tensor(self, t, dst):
        d = self.dist
        cplx = t.data.is_complex()
        hdr = torch.tensor(list(t.dims) + [1 if cplx else 0], dtype=torch.int64, device=t.data.device)
        payload = torch.view_as_real(t.data).reshape(-1) if cplx else t.data
        if self.staged:
            hdr, payload = hdr.cpu(), payload.cpu()
        return [d.isend(hdr, dst, group=self.group), d.isend(payload.contiguous(), dst, group=self.group)]

    def _recv_tensor(self, src, device):
        d = self.dist
        rdev = "cpu" if self.staged else device
        hdr = torch.empty(4, dtype=torch.int64, device=rdev)
        d.recv(hdr, src, group=self.group)
        a, b, c, cplx = [int(x) for x in hdr.tolist()]
        n = a * b * c
        buf = torch.empty(2 * n if cplx else n, dtype=torch.float64, device=rdev)
        d.recv(buf, src, group=self.group)
        buf = buf.to(device)
        data = torch.view_as_complex(buf.view(n, 2)) if cplx else buf
        return DTensor(data, (a, b, c))

    def _send_vec(self, v, dst):
        d = self.dist
        hdr = torch.tensor([v.numel()], dtype=torch.int64, device=v.device)
        if self.staged:
            hdr, v = hdr.cpu(), v.cpu()
        return [d.isend(hdr, dst, group=self.group), d.isend(v.contiguous(), dst, group=self.group)]

    def _recv_vec(self, src, device):
        d = self.dist
        rdev = "cpu" if self.staged else device
        hdr = torch.empty(1, dtype=torch.int64, device=rdev)
        d.recv(hdr, src, group=self.group)
        v = torch.empty(int(hdr.item()), dtype=torch.float64, device=rdev)
        d.recv(v, src, group=self.group)
        return v.to(device)

    def warm_links(self):
        """Exchange one element with each neighbour: NCCL opens its point-to-point channels lazily (seconds on the
        first send/recv), which would otherwise be charged to the first odd layer."""
        dev = self.state.Bs[0].data.device
        one = torch.zeros(1, dtype=torch.float64, device=dev)
        pending = []
        if self.rank > 0:
            pending += self._send_vec(one, self.rank - 1)
        if self.rank < self.world - 1:
            self._recv_vec(self.rank + 1, dev)
            pending += self._send_vec(one, self.rank + 1)
        if self.rank > 0:
            self._recv_vec(self.rank - 1, dev)
        for p in pending:
            p.wait()

    def layer(self, G, parity, maxdim=None, cutoff=0.0):
        """One even (0) or odd (1) layer over the whole chain; returns this rank's largest truncation error."""
        st = self.state
        lo = st.first
        hi = lo + len(st)
        dev = st.Bs[0].data.device
        # does the bond (hi-1, hi) belong to this layer and straddle two blocks?
        right_boundary = self.rank < self.world - 1 and ((hi - 1) % 2) == parity
        left_boundary = self.rank > 0 and ((lo - 1) % 2) == parity
        pending = []
        if left_boundary:            # my first site goes to the left neighbour, which owns the gate
            pending += self._send_tensor(st.Bs[0], self.rank - 1)
        halo = self._recv_tensor(self.rank + 1, dev) if right_boundary else None
        for p in pending:
            p.wait()
        worst = tebd_layer(st, G, parity, maxdim=maxdim, cutoff=cutoff, gate_fn=self.gate_fn)
        pending = []
        if right_boundary:
            B1, B2, lam, err = self.gate_fn(G, st.lams[-2], st.Bs[-1], halo, maxdim=maxdim, cutoff=cutoff)
            st.Bs[-1], st.lams[-1] = B1, lam
            worst = max(worst, err)
            pending += self._send_tensor(B2, self.rank + 1) + self._send_vec(lam, self.rank + 1)
        if left_boundary:
            st.Bs[0] = self._recv_tensor(self.rank - 1, dev)
            st.lams[0] = self._recv_vec(self.rank - 1, dev)
        for p in pending:
            p.wait()
        return worst

    def gather(self):
        """All blocks to every rank (verification / output; not part of a time step).  Returns a full BState."""
        d = self.dist
        objs = [None] * self.world
        mine = ([(t.data.cpu(), t.dims) for t in self.state.Bs], [l.cpu() for l in self.state.lams])
        d.all_gather_object(objs, mine, group=self.group)
        dev = self.state.Bs[0].data.device
        Bs, lams = [], []
        for g, (bs, ls) in enumerate(objs):
            Bs += [DTensor(x.to(dev), dims) for x, dims in bs]
            lams += [l.to(dev) for l in (ls if g == self.world - 1 else ls[:-1])]
        return BState(Bs, lams)
