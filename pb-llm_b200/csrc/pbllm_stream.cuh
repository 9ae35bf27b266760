// Block-stream layout of 16-bit (fp16 / bf16) layers: the ONE resident form of a packed layer that the decode kernel
// streams and that unpack / the prefill expansion read back.
//
// A layer N x K is cut into blocks of 32 output rows x 64 input columns, stored in ROW-GROUP-MAJOR order
// (block id = rg * tiles_c + kb, rg = row / 32, kb = col / 64) so that a warp's run of consecutive blocks is one
// contiguous stream:
//   fsign uint2 [blocks][32]   sign bits in MMA-FRAGMENT order: lane (g = lane>>2, t = lane&3) of the warp that
//                              processes the block owns rows {g, g+8, g+16, g+24} x columns 16t..16t+15 -- exactly the
//                              weights of its mma.sync.m16n8k16 A fragments -- and finds the bit of the fragment
//                              register rho = 4q + i (k16 step q, A register i) at bit 15-rho (low half) and 31-rho
//                              (high half) of word h (rows 0-15 / 16-31): one shift + one LOP3 per register.
//                              bit 1 = the LOW level of the (row, group), bit 0 = the HIGH level.
//   eptr  u32   [blocks + 1]   offset of each block's salient entries in 16-byte units (4 entries)
//   ent   u32   [..]           one entry per salient weight (value that is neither level, or outside the low mask):
//                                bits 31..21  slot: 16-bit position in the kernel's swizzled 32x64 tile
//                                bit  20      0
//                                bits 19..16  k: ulp correction of the reconstruction (two's complement, -7..7; -8 = the
//                                             exact value lives in the exception list)
//                                bits 15..0   tau = fl16((v - mid) / half): the weight in the +-1 units of its (row, group),
//                                             mid = (lo+hi)/2, half = (hi-lo)/2 (fp32); the sign bit of a salient position is 0
//                              v == step(fl16(mid + half * tau), k) reproduces the salient value bit-exactly (checked at
//                              pack time).  The decode kernel's tile holds +1.0 everywhere except tau at the salient slots
//                              and is XORed with the sign bits: ONE LOP3 per fragment register yields the block in +-1
//                              units, salient weights included, and the levels are applied to the fp32 accumulators.
//                              Blocks are padded to 4 entries with copies of one of their entries (idempotent stores).
//   exc   u32   [n_exc][2]     {block id, slot << 16 | exact 16-bit value} for the rare entries whose |k| > 7
//   affine float2 [n_pad][groups] {lo, hi} as in the plane layout; a (row, group) with ONE level (lo == hi) that also has
//                              salient weights is rewritten to {mid-1, mid+1} at pack time, which turns all its positions
//                              into entries (tau = v - mid): half = 0 could not carry them.
#pragma once
#include "pbllm_common.cuh"

namespace pbl {
namespace st {

// position (r, c) of a block -> owner lane, word and bit of the fragment-ordered sign words
__host__ __device__ inline void sign_pos(uint32_t r, uint32_t c, uint32_t& lane, uint32_t& word, uint32_t& bit) {
    const uint32_t g = r & 7u, j = r >> 3, t = c >> 4, o = c & 15u;
    const uint32_t q = o >> 2, hi = (o >> 1) & 1u, e = o & 1u;
    lane = 4u * g + t;
    word = j >> 1;
    bit = (15u - (4u * q + (j & 1u) + 2u * hi)) + 16u * e;
}

// position (r, c) -> 16-bit slot in the swizzled correction tile (128-byte rows, 16-byte chunks XOR-swizzled by r & 7;
// chunk pc = 2q + hi holds, at word t, the columns 16t + 4q + 2hi + {0,1}: see decode kernel "Tile layout")
__host__ __device__ inline uint32_t tile_slot(uint32_t r, uint32_t c) {
    const uint32_t pc = (c >> 1) & 7u;
    return r * 64u + ((pc ^ (r & 7u)) << 3) + (c >> 4) * 2u + (c & 1u);
}
__host__ __device__ inline void slot_pos(uint32_t slot, uint32_t& r, uint32_t& c) {
    r = slot >> 6;
    const uint32_t pc = ((slot >> 3) & 7u) ^ (r & 7u), t = (slot >> 1) & 3u, e = slot & 1u;
    c = 16u * t + 2u * pc + e;
}

// monotone integer order of 16-bit floats (sign-magnitude), for the ulp correction k
__host__ __device__ inline int ord16(uint32_t b) { return (b & 0x8000u) ? -(int)(b & 0x7FFFu) : (int)(b & 0x7FFFu); }
__host__ __device__ inline uint32_t unord16(int n) { return n < 0 ? (0x8000u | (uint32_t)(-n)) : (uint32_t)n; }

constexpr int kMaxK = 7;          // |k| <= 7 is stored in the entry; otherwise k = -8 and the value is in the exception list
constexpr uint32_t kExcK = 8u;    // 4-bit pattern of -8

__host__ __device__ inline uint32_t make_entry(uint32_t slot, int k, uint32_t tau16) {
    return (slot << 21) | (((uint32_t)k & 15u) << 16) | (tau16 & 0xFFFFu);
}
__host__ __device__ inline uint32_t entry_slot(uint32_t e) { return e >> 21; }
__host__ __device__ inline int entry_k(uint32_t e) { const int k4 = (int)((e >> 16) & 15u); return (k4 & 8) ? k4 - 16 : k4; }

}  // namespace st
}  // namespace pbl
