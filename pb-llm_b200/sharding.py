"""Row-sharded execution of one packed linear across the GPUs of a node.

Output rows are independent, so rank g owns rows [g*n_loc, (g+1)*n_loc) of the packed weight
(planes, affine and salient values split for free because the packed layout is row-tile-major),
x is replicated, every rank computes y[:, its rows] with the same libpbllm kernel, and ONE
all-gather of the [M, n_loc] slices per linear rebuilds y on every rank -- K is never split so no
all-reduce exists on this path. Collective: torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


def shard_rows(N: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Rows [r0, r1) owned by `rank` and the uniform padded slice width n_loc = ceil(N/world)
    (the last ranks may own fewer -- even zero -- real rows; their slice is zero-padded)."""
    n_loc = (N + world - 1) // world
    r0 = min(N, rank * n_loc)
    r1 = min(N, r0 + n_loc)
    return r0, r1, n_loc


def gather_rows(y_loc: torch.Tensor, N: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather the per-rank [M, n_loc] slices into the full [M, N] output."""
    world = dist.get_world_size(group)
    M, n_loc = y_loc.shape
    buf = torch.empty((world * M, n_loc), dtype=y_loc.dtype, device=y_loc.device)
    dist.all_gather_into_tensor(buf, y_loc.contiguous(), group=group)
    y = buf.view(world, M, n_loc).permute(1, 0, 2).reshape(M, world * n_loc)
    if world * n_loc != N:
        y = y[:, :N]
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


class RowShardedLinear(nn.Module):
    """Wraps the rank-local shard module (any callable x[M,K] -> y[M, rows_owned]) and performs
    the per-linear all-gather. Build with `from_dense` on the GPU box."""

    def __init__(self, local: Callable, N: int, K: int, rank: int, world: int, group=None):
        super().__init__()
        self.local, self.N, self.K, self.rank, self.world, self.group = local, N, K, rank, world, group
        self.r0, self.r1, self.n_loc = shard_rows(N, world, rank)

    @classmethod
    def from_dense(cls, w_sim: torch.Tensor, bias=None, low_mask=None, groupsize: int = -1, group=None):
        """Pack only this rank's rows of the dense fake-quant weight."""
        from .packing import PackedLinear
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        N, K = w_sim.shape
        r0, r1, _ = shard_rows(N, world, rank)
        local = None
        if r1 > r0:
            p = PackedLinear.from_dense(w_sim[r0:r1], None if bias is None else bias[r0:r1],
                                        None if low_mask is None else low_mask[r0:r1], groupsize)
            local = p.forward
        return cls(local, N, K, rank, world, group)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.K)
        M = x2.shape[0]
        y_loc = torch.zeros((M, self.n_loc), dtype=x.dtype, device=x.device)
        rows = self.r1 - self.r0
        if rows > 0:
            y_loc[:, :rows] = self.local(x2)
        y = gather_rows(y_loc, self.N, self.group)
        return y.view(*lead, self.N)
