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
