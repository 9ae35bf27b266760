"""Multi-GPU parity of the row-sharded paths (needs >= 2 B200s on one node; skipped otherwise): ranks are spawned here,
one process per GPU, NCCL for the bootstrap. Checks (i) RowShardedLinear (NCCL all-gather) and (ii) PushLinear -- the decode
kernel with the fused peer-store epilogue and in-kernel completion flags -- against the oracle's dense forward."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import pbllm_b200 as pb
        from pbllm_b200.sharding import PeerContext, PushLinear, RowShardedLinear
        from oracle import oracle as orc
        from test_gpu_parity import synth_wsim, rounded
        from oracle.gen_golden import make_x
        errs = {}
        ctx = PeerContext(dev, arena_bytes=8 << 20)
        layers = []
        for (N, K, M) in [(768, 768, 8), (1024, 2048, 3), (2048, 512, 16), (264, 520, 5)]:
            w, low = synth_wsim(N, K, -1, torch.float16, seed=N + K)
            b = rounded(np.random.RandomState(N).standard_normal(N).astype(np.float32) * 0.1, torch.float16)
            x = rounded(make_x(N + M, (M, K)), torch.float16)
            ref = orc.linear(x, w, b)
            wd, lowd, bd = (torch.from_numpy(a).to(dev) for a in (w, low, b))
            xd = torch.from_numpy(x).to(dev).half()
            pl = PushLinear(ctx, wd.half(), bd.half(), lowd)
            layers.append((pl, xd, ref, wd, lowd, bd))
        # three rounds through all layers back to back: buffers are reused, epochs advance, every round must be exact again
        for rnd in range(3):
            outs = [pl.forward(xd, wait_prev=True) for (pl, xd, *_rest) in layers]
            ctx.wait()
            torch.cuda.synchronize()
            for i, ((pl, xd, ref, *_r), y) in enumerate(zip(layers, outs)):
                e = float(np.abs(y.float().cpu().numpy() - ref).max() / np.abs(ref).max())
                errs[f"push{i}r{rnd}"] = e
            dist.barrier()
        # CUDA-graph replay of the pushed chain (the per-token step is launch-bound from Python)
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for (pl, xd, *_r) in layers:
                pl.forward(xd)
            ctx.wait()
            torch.cuda.synchronize()
            dist.barrier()
            with torch.cuda.graph(g, stream=side):
                for (pl, xd, *_r) in layers:
                    pl.forward(xd)
                ctx.wait()
        for _ in range(3):
            for (pl, *_r) in layers:
                pl.out.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            g.replay()
            torch.cuda.synchronize()
            dist.barrier()
        for i, (pl, xd, ref, *_r) in enumerate(layers):
            errs[f"graph{i}"] = float(np.abs(pl.out[: xd.shape[0]].float().cpu().numpy() - ref).max() / np.abs(ref).max())
        # NCCL all-gather variant at a prefill size
        pl, xd, ref, wd, lowd, bd = layers[1]
        xs = torch.from_numpy(rounded(make_x(5, (300, wd.shape[1])), torch.float16)).to(dev).half()
        rs = RowShardedLinear.from_dense(wd.half(), bd.half(), lowd)
        y = rs(xs)
        ref2 = orc.linear(xs.float().cpu().numpy(), wd.float().cpu().numpy(), bd.float().cpu().numpy())
        errs["allgather"] = float(np.abs(y.float().cpu().numpy() - ref2).max() / np.abs(ref2).max())
        q.put((rank, errs))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, {"exception": traceback.format_exc() + str(e)}))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_paths_match_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on the node")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, errs in res:
        assert "exception" not in errs, errs["exception"]
        assert len(errs) >= 17 and max(errs.values()) <= 1e-3, (rank, errs)
