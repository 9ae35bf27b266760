"""Host logic of the decode kernel (no GPU): the launch plan libpbllm computes (pbl_decode_plan) against a Python
mirror of the kernel's work partition and reductions (csrc/pbllm_decode.cu).  For every shape: each (row group,
k-block) is accumulated exactly once, every row group is finished exactly once, the cross-CTA slot numbering stays
inside the planned slot count, and only the LAST contributor of a shared row group ever waits (on CTAs dispatched
before it)."""
import ctypes as C

import numpy as np
import pytest

from pbllm_b200 import _lib

KW = 8


def plan(N, K, M, sms, ctas):
    out = (C.c_uint32 * 8)()
    assert _lib.load().pbl_decode_plan(N, K, M, sms, ctas, out) == 0
    return dict(zip(["blocks", "rgs", "grid", "passes", "q", "rem", "slots", "ws_kib"], list(out)))


def mirror(pl, TC, order_seed=0):
    B, G, q, rem, slots, rgs = pl["blocks"], pl["grid"], pl["q"], pl["rem"], pl["slots"], pl["rgs"]
    assert B == rgs * TC and q == B // (G * KW) and rem == B % (G * KW)
    wstart = lambda gw: gw * q + min(gw, rem)
    assert wstart(G * KW) == B

    def owner(b):
        cut = rem * (q + 1)
        gw = b // (q + 1) if b < cut else rem + (b - cut) // q
        return gw // KW

    y, part = {}, {}
    waits = []
    order = list(range(G))
    np.random.RandomState(order_seed).shuffle(order)            # CTAs may finish in any order
    for c in order:
        gw0 = c * KW
        c_lo, c_hi = wstart(gw0), wstart(gw0 + KW)
        assert c_lo < c_hi, "every CTA of the grid has work"
        head, tail = {}, {}
        for w in range(KW):
            w_lo, w_hi = wstart(gw0 + w), wstart(gw0 + w + 1)
            if w_lo >= w_hi:
                continue
            rg, kb = divmod(w_lo, TC)
            rg_first, acc = rg, 0
            for blk in range(w_lo, w_hi):
                assert owner(blk) == c
                more = blk + 1 < w_hi
                acc += blk + 1
                kb += 1
                rg_end = kb == TC
                if rg_end or not more:
                    if rg_end and w_lo <= rg * TC:
                        assert rg not in y
                        y[rg] = acc                                # the warp saw the whole row group
                    elif rg == rg_first:
                        head[w] = (rg, acc)
                    else:
                        assert not more
                        tail[w] = (rg, acc)
                    acc = 0
                    if rg_end:
                        kb, rg = 0, rg + 1
        rg_a, rg_b = c_lo // TC, (c_hi - 1) // TC
        hs = c_lo > rg_a * TC or c_hi < rg_a * TC + TC
        ts = rg_b != rg_a and c_hi < rg_b * TC + TC
        for r in range(rg_a, rg_b + 1):
            vals = [v for d in (head, tail) for (rr, v) in d.values() if rr == r]
            if not vals:
                continue
            v = sum(vals)
            split = (r == rg_a and hs) or (r == rg_b and ts)
            if not split:
                assert r not in y
                y[r] = v
                continue
            first = owner(r * TC)
            slot, expected = c - first, owner(r * TC + TC - 1) - first + 1
            assert 0 <= slot < expected <= slots
            if slot + 1 < expected:
                part[(r, slot)] = v                                # writer: never waits
            else:
                waits.append((c, r, expected))
                part[(r, slot)] = v
    for c, r, expected in waits:                                   # the finalizer only waits for lower-numbered CTAs
        assert r not in y
        y[r] = sum(part[(r, k)] for k in range(expected))
    for r in range(rgs):
        assert y.get(r) == sum(b + 1 for b in range(r * TC, (r + 1) * TC)), r
    return True


@pytest.mark.parametrize("N,K", [(4096, 4096), (11008, 4096), (4096, 11008), (768, 768), (3072, 768), (768, 3072), (2048, 8192),
                                 (50272, 2048), (100, 70), (5000, 64), (64, 8192), (32, 64), (13824, 5120)])
@pytest.mark.parametrize("ctas", [1, 2, 3, 4])
def test_plan_partition_is_exact_and_slots_suffice(N, K, ctas):
    pl = plan(N, K, 8, 148, ctas)
    TC = (K + 63) // 64
    assert pl["rgs"] == ((N + 127) // 128) * 4 and pl["grid"] <= 148 * ctas and pl["grid"] * KW <= max(pl["blocks"], KW)
    assert mirror(pl, TC, order_seed=N + K + ctas)
    assert pl["ws_kib"] * 1024 >= pl["passes"] * pl["rgs"] * pl["slots"] * 256 * 8 - 1023


def test_plan_token_passes_and_errors():
    # up to 8 tokens: one group per pass; 9..16: two groups against one expansion of each tile, still one pass
    assert plan(4096, 4096, 8, 148, 3)["passes"] == 1 and plan(4096, 4096, 9, 148, 3)["passes"] == 1
    assert plan(4096, 4096, 16, 148, 3)["passes"] == 1 and plan(4096, 4096, 17, 148, 3)["passes"] == 2
    assert plan(4096, 4096, 16, 148, 2)["ws_kib"] == 2 * plan(4096, 4096, 8, 148, 2)["ws_kib"]   # same grid, twice the outputs
    out = (C.c_uint32 * 8)()
    assert _lib.load().pbl_decode_plan(0, 64, 1, 148, 3, out) == -3
    assert _lib.load().pbl_decode_plan(64, 64, 1, 148, 3, None) == -1
