"""CPU checks of the block-stream layout (csrc/pbllm_stream.cuh) against the documented register layouts of
mma.sync.m16n8k16 and ldmatrix: the packer's bit / slot positions (pbl_stream_position, host-only) must be exactly what
the decode kernel's index arithmetic reads back. The kernel side is restated here in numpy, line by line, from
csrc/pbllm_decode.cu; no GPU is needed."""
import ctypes as C

import numpy as np

from pbllm_b200 import _lib


def positions():
    lib = _lib.load()
    out = (C.c_uint32 * 4)()
    pos = np.zeros((32, 64, 4), np.int64)
    for r in range(32):
        for c in range(64):
            assert lib.pbl_stream_position(r, c, out) == 0
            pos[r, c] = list(out)
    return pos


def mma_m16n8k16(a_regs, b_regs):
    """PTX ISA 'Matrix fragments for mma.m16n8k16 with .f16': a_regs[lane][i] = (lo, hi) pair of register a_i,
    b_regs[lane][j] = pair of b_j. Returns D [16 rows][8 cols]."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(4):
            row = g + 8 * (i & 1)
            k0 = 2 * t + 8 * (i >> 1)
            A[row, k0], A[row, k0 + 1] = a_regs[lane][i]
        for j in range(2):
            k0 = 2 * t + 8 * j
            B[k0, g], B[k0 + 1, g] = b_regs[lane][j]
    return A @ B


def test_positions_are_a_bijection():
    pos = positions()
    bits = {(int(p[0]), int(p[1]), int(p[2])) for p in pos.reshape(-1, 4)}
    assert len(bits) == 2048 and all(0 <= l < 32 and w in (0, 1) and 0 <= b < 32 for l, w, b in bits)
    assert sorted(int(s) for s in pos[..., 3].ravel()) == list(range(2048))


def test_dense_fragments_from_sign_words_reproduce_the_block_product():
    """Kernel: a_i of k16 step q, row half h = ((word_h << (4q+i)) & 0x80008000) ^ {1.0,1.0}; B fragments = the lane's own
    16 consecutive activations (token g, columns 16t..16t+15), words 2q and 2q+1."""
    pos = positions()
    rs = np.random.RandomState(0)
    low = rs.rand(32, 64) < 0.5                                # bit 1 = LOW level = -1
    x = rs.standard_normal((8, 64))                            # [token][column]
    words = np.zeros((32, 2), np.uint64)
    for r in range(32):
        for c in range(64):
            if low[r, c]:
                lane, w, b = pos[r, c, :3]
                words[lane, w] |= np.uint64(1) << np.uint64(b)
    ref = np.where(low, -1.0, 1.0) @ x.T                       # [row][token]
    got = np.zeros((32, 8))
    for h in range(2):
        for q in range(4):
            a_regs, b_regs = [], []
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                regs = []
                for i in range(4):
                    sh = (int(words[lane, h]) << (4 * q + i)) & 0x80008000
                    regs.append((-1.0 if sh & 0x8000 else 1.0, -1.0 if sh & 0x80000000 else 1.0))
                a_regs.append(regs)
                xw = [(x[g, 16 * t + 2 * w], x[g, 16 * t + 2 * w + 1]) for w in range(8)]
                b_regs.append([xw[2 * q], xw[2 * q + 1]])
            got[16 * h:16 * h + 16] += mma_m16n8k16(a_regs, b_regs)
    assert np.allclose(got, ref, atol=1e-12)


def test_correction_tile_slots_match_ldmatrix_addresses():
    """Kernel: lm_row = (lane & 7) + ((lane >> 3) & 1) * 8; address = tile + lm_row*128 + (((lane >> 4) ^ (lm_row & 7)) << 4),
    XOR (q << 5), + h*2048; ldmatrix.x4 matrix i takes its eight 16-byte rows from lanes 8i..8i+7 and hands thread (g, t)
    elements 2t, 2t+1 of row g."""
    pos = positions()
    rs = np.random.RandomState(1)
    corr = np.where(rs.rand(32, 64) < 0.1, rs.standard_normal((32, 64)), 0.0)
    x = rs.standard_normal((8, 64))
    tile = np.zeros(2048)
    for r in range(32):
        for c in range(64):
            tile[pos[r, c, 3]] = corr[r, c]
    ref = corr @ x.T
    got = np.zeros((32, 8))
    for h in range(2):
        for q in range(4):
            addr = []
            for lane in range(32):
                lm_row = (lane & 7) + ((lane >> 3) & 1) * 8
                base = lm_row * 128 + (((lane >> 4) ^ (lm_row & 7)) << 4)
                addr.append((base ^ (q << 5)) + h * 2048)
            a_regs, b_regs = [], []
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                regs = []
                for i in range(4):
                    row_addr = addr[8 * i + g] // 2             # 16-bit slot of the matrix row's first element
                    regs.append((tile[row_addr + 2 * t], tile[row_addr + 2 * t + 1]))
                a_regs.append(regs)
                xw = [(x[g, 16 * t + 2 * w], x[g, 16 * t + 2 * w + 1]) for w in range(8)]
                b_regs.append([xw[2 * q], xw[2 * q + 1]])
            got[16 * h:16 * h + 16] += mma_m16n8k16(a_regs, b_regs)
    assert np.allclose(got, ref, atol=1e-12)


def test_entry_patch_address_is_twice_the_slot():
    """entry = slot << 21 | k << 16 | c16 with bit 20 clear: the kernel's patch address is tile + (entry >> 20)."""
    for slot in (0, 1, 777, 2047):
        for k in (-8, -1, 0, 7):
            e = (slot << 21) | ((k & 15) << 16) | 0xBEEF
            assert (e >> 20) == 2 * slot and (e & 0xFFFF) == 0xBEEF
            k4 = (e >> 16) & 15
            assert (k4 - 16 if k4 & 8 else k4) == k
