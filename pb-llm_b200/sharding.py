"""Row-sharded execution of one packed linear across the GPUs of a node.

Output rows are independent, so rank g owns rows [g*n_loc, (g+1)*n_loc) of the packed weight
(planes, affine and salient values split for free because the packed layout is row-tile-major),
x is replicated, every rank computes y[:, its rows] with the same libpbllm kernel, and ONE
all-gather of the [M, n_loc] slices per linear rebuilds y on every rank -- K is never split so no
all-reduce exists on this path. Two gather implementations:
  * `RowShardedLinear`: torch.distributed all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) -- any M;
    the prefill regime, where the slices are megabytes and the collective is bandwidth-bound.
  * `PushLinear` (+ `PeerContext`): the per-token regime (M <= 16), where a slice is a few KB and a collective call costs
    more than the kernel.  The all-gather is FUSED into the decode kernel (pbl_linear_forward_push): its epilogue stores
    the slice straight into every rank's output buffer through peer-mapped (symmetric) memory over NVLink / NVSwitch and
    the completion is an in-kernel flag exchange -- no NCCL call on the path."""
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


class PeerContext:
    """Per-process plumbing of the fused push path: a symmetric (peer-mapped) arena for the gathered outputs, the flag
    arrays of the in-kernel completion exchange, and the local sync counters. torch's symmetric memory does the handle
    exchange; every arithmetic / communication step then happens inside libpbllm.so kernels."""

    def __init__(self, device, group=None, arena_bytes: int = 64 << 20):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"the push path addresses at most {_lib.MAX_PEERS} ranks (one NVSwitch node)")
        self.device = device
        self.arena = symm.empty(arena_bytes, dtype=torch.uint8, device=device)
        self.arena.zero_()
        self._ah = symm.rendezvous(self.arena, self.group)
        self.flags = symm.empty(_lib.MAX_PEERS, dtype=torch.int32, device=device)
        self.flags.zero_()
        self._fh = symm.rendezvous(self.flags, self.group)
        self.arena_ptrs = [int(p) for p in self._ah.buffer_ptrs]
        self.flag_ptrs = [int(p) for p in self._fh.buffer_ptrs]
        self.sync_ctr = torch.zeros(2, dtype=torch.int32, device=device)
        self._off = 0
        torch.cuda.synchronize(device)
        dist.barrier(self.group)            # every rank's arena and flags are zeroed before anybody pushes

    def alloc(self, nbytes: int) -> int:
        """Byte offset of a fresh 256-byte-aligned region at the SAME offset in every rank's arena (all ranks must
        allocate in the same order)."""
        off = self._off
        self._off = (off + nbytes + 255) // 256 * 256
        if self._off > self.arena.numel():
            raise RuntimeError("PeerContext arena exhausted: pass a larger arena_bytes")
        return off

    def push_desc(self, out_off: int, col0_bytes: int, wait_prev: bool):
        from . import _lib
        d = _lib.PblPeerPush()
        for r in range(self.world):
            d.y[r] = self.arena_ptrs[r] + out_off + col0_bytes
            d.flags[r] = self.flag_ptrs[r]
        d.sync_ctr = self.sync_ctr.data_ptr()
        d.n_ranks, d.rank, d.wait_prev = self.world, self.rank, 1 if wait_prev else 0
        return d

    def wait(self):
        """Stream-ordered: returns (on the stream) once every rank's latest push has landed here."""
        import ctypes as C
        from . import _lib
        d = self.push_desc(0, 0, True)
        _lib.check(_lib.load().pbl_peer_wait(C.byref(d), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                   "pbl_peer_wait")


class PushLinear:
    """One row-sharded linear of the per-token regime: this rank's rows packed locally, outputs gathered on every rank by
    the kernel itself. `forward(x)` returns this rank's [M, N] output view, complete once the next pushed kernel has
    started or `ctx.wait()` has been enqueued."""

    def __init__(self, ctx: PeerContext, w_sim: torch.Tensor, bias=None, low_mask=None, groupsize: int = -1, max_tokens: int = 16):
        from .packing import PackedLinear
        self.ctx = ctx
        self.N, self.K = w_sim.shape
        self.r0, self.r1, self.n_loc = shard_rows(self.N, ctx.world, ctx.rank)
        if self.r1 <= self.r0:
            raise RuntimeError("PushLinear: every rank must own at least one row (the completion exchange needs all ranks)")
        self.p = PackedLinear.from_dense(w_sim[self.r0:self.r1], None if bias is None else bias[self.r0:self.r1],
                                         None if low_mask is None else low_mask[self.r0:self.r1], groupsize)
        if not self.p.stream_layout:
            raise RuntimeError("PushLinear needs an fp16 / bf16 layer")
        self.max_tokens = max_tokens
        self.es = w_sim.element_size()
        self.out_off = ctx.alloc(max_tokens * self.N * self.es)
        nb = max_tokens * self.N * self.es
        self.out = ctx.arena[self.out_off:self.out_off + nb].view(w_sim.dtype).view(max_tokens, self.N)
        self._desc = {True: ctx.push_desc(self.out_off, self.r0 * self.es, True),
                      False: ctx.push_desc(self.out_off, self.r0 * self.es, False)}

    def forward(self, x: torch.Tensor, wait_prev: bool = True) -> torch.Tensor:
        import ctypes as C
        from . import _lib, packing
        M = x.shape[0]
        if M > self.max_tokens or x.dtype != self.p.dtype or x.shape[1] != self.K or x.stride(1) != 1:
            raise RuntimeError("PushLinear.forward: x must be [M <= max_tokens, K] of the layer's dtype")
        dev = x.device
        st = torch.cuda.current_stream(dev).cuda_stream
        ws = packing._decode_workspace(dev, st, self.p._dws_bytes)
        rc = _lib.load().pbl_linear_forward_push(self.p.handle, x.data_ptr(), x.stride(0) if M > 1 else self.K,
                                                 C.byref(self._desc[bool(wait_prev)]), self.N, M, ws.data_ptr(), ws.numel(), st)
        _lib.check(rc, "pbl_linear_forward_push")
        return self.out[:M]
