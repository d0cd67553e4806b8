"""Multi-GPU use of the engine: the batch of independent NMPC instances is the only parallel axis.

One process per GPU (torch.distributed, backend nccl; gloo in the CPU tests).  Rank r owns the contiguous slice
`shard_range(B, r, world)` of every batched input, runs the single-GPU engine on it, and ONE all-gather of the packed
results (x, u, status, sqp_iter, qp_iter, 4 residuals per instance) gives every rank the whole batch.  There is no
exchange inside the solve, so nothing else is communicated.
"""
import numpy as np


def shard_range(B, rank, world):
    """contiguous, balanced slice [lo, hi) of a batch of B instances owned by `rank`"""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def packed_width(N, nx, nu):
    return (N + 1) * nx + N * nu + 7


def pack_results(torch, x, u, stats):
    """[b, (N+1)*nx + N*nu + 7]: trajectories then status, sqp_iter, qp_iter, res_stat, res_eq, res_ineq, res_comp"""
    b = x.shape[0]
    return torch.cat([x.reshape(b, -1), u.reshape(b, -1), stats[:, :7]], dim=1).contiguous()


def unpack_results(packed, N, nx, nu):
    a = (N + 1) * nx
    b = a + N * nu
    return packed[:, :a].reshape(-1, N + 1, nx), packed[:, a:b].reshape(-1, N, nu), packed[:, b:]


def gather_buffer(torch, B_local, world, rank, N, nx, nu, device):
    """(gathered [world*B_local, width], own slice) for equal shards: the engine's epilogue writes this rank's packed
    result rows straight into its slice of the gathered tensor (solver.set_result_buffer(own slice)), and the
    all-gather runs in place -- no pack kernel, no concatenation."""
    width = packed_width(N, nx, nu)
    out = torch.zeros((world * B_local, width), dtype=torch.float64, device=device)
    return out, out[rank * B_local:(rank + 1) * B_local]


def all_gather_in_place(dist, gathered, own):
    """the one collective of the path"""
    dist.all_gather_into_tensor(gathered, own)
    return gathered


def all_gather_results(torch, dist, packed, B, world):
    """All-gather the per-rank packed results into [B, width] in global instance order (ragged shards padded)."""
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    m = max(sizes)
    width = packed.shape[1]
    pad = packed
    if packed.shape[0] < m:
        pad = torch.cat([packed, packed.new_zeros((m - packed.shape[0], width))], dim=0)
    out = packed.new_empty((world * m, width))
    dist.all_gather_into_tensor(out, pad.contiguous())
    if all(s == m for s in sizes):
        return out
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)], dim=0)
