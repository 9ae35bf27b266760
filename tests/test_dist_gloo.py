"""N>1 host logic on CPU: world_size-2 gloo. The row-shard wrapper's partitioning, padding and
all-gather reassembly are exercised with the CPU oracle standing in for the rank-local kernel
(the CUDA kernel itself is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pbllm_b200 as pb
from pbllm_b200.sharding import RowShardedLinear, shard_rows
from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, K, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(7)
        W = (rs.standard_normal((N, K)) * 0.02).astype(np.float32)
        b = (rs.standard_normal(N) * 0.1).astype(np.float32)
        x = rs.standard_normal((M, K)).astype(np.float32)
        w_sim, _, _ = orc.xnor_wsim(W)
        r0, r1, n_loc = shard_rows(N, world, rank)

        def local(xt):  # stand-in for PackedLinear.forward on this rank's rows
            return torch.from_numpy(orc.linear(xt.numpy(), w_sim[r0:r1], b[r0:r1]))

        layer = RowShardedLinear(local if r1 > r0 else None, N, K, rank, world)
        y = layer(torch.from_numpy(x).view(1, M, K))
        ref = orc.linear(x, w_sim, b)
        ok = y.shape == (1, M, N) and np.array_equal(y.view(M, N).numpy(), ref)
        q.put((rank, bool(ok), (r0, r1, n_loc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,K,M", [(96, 64, 5), (67, 32, 3), (1, 32, 2)])
def test_row_shard_all_gather_world2(N, K, M):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, K, M, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(30) for p in procs]
    assert all(r[1] for r in res), res
    (r0a, r1a, nl), (r0b, r1b, _) = res[0][2], res[1][2]
    assert r0a == 0 and r1a == r0b and r1b == N and nl == (N + 1) // 2


def test_shard_rows_partition():
    for N in (1, 7, 128, 4096, 11008, 13824):
        for world in (1, 2, 4, 8):
            spans = [shard_rows(N, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert len({s[2] for s in spans}) == 1 and spans[0][2] * world >= N
