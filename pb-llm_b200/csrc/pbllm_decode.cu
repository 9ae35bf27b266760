// Decode kernel (M <= 16 tokens per call, fp16 / bf16): the HBM-bound regime of the bit-plane forward.
//
// Same math as every other kernel of the library (y = x . w_sim^T + b, the EXACT w_sim tile rebuilt in shared
// memory and multiplied on the tensor cores with fp32 accumulation), reorganised around what bounds a
// 2-microsecond kernel: instructions per weight, dependent DRAM round trips, and SM load balance.
//
//  * decode index (built once at pack time, pbl_decode_index_*): the layer in ROW-GROUP-MAJOR block order --
//    block (rg, kb) = 32 output rows x 64 input columns, linear id rg*tiles_c + kb:
//      dsign uint2 [blocks][32]        the sign words of the block's rows (1 bit / weight)
//      eptr  u32   [blocks + 1]        offset of each block's salient entries, in 16-byte units
//      ent   u32   [..]                one entry per salient weight: (byte offset in the 4 KB swizzled tile) << 16
//                                      | the value's 16 bits; blocks padded to 4 entries with copies of their last
//                                      entry (an idempotent store).
//    With explicit positions the salient patch is lane-balanced: entry e of a block is handled by lane e/4 % 32,
//    3 instructions per entry, no per-row bit walking, no warp scan, no divergence.
//  * warp-granular stream-K: the blocks of the layer are dealt out in contiguous, equal (+-1) runs to the
//    warps of a fixed grid (3 CTAs per SM), so every SM gets the same number of blocks whatever N and K are.
//    A warp's run covers at most two partial row groups (head / tail) plus whole ones; partials are reduced
//    across the warps of the CTA in shared memory, and across CTAs through a small global workspace: each
//    contributor parks {partial, valid tag} with one 64-bit store per output (no fence, no counter), the last
//    contributor polls the slots and sums them in CTA order, so the result is deterministic.
//  * every weight-side load of a warp's first blocks is issued before griddepcontrol.wait: under programmatic
//    dependent launch the packed stream of layer i+1 is in flight while layer i still computes.
#include <cstdlib>
#include <type_traits>

#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace dk {
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTok = 8;                       // tokens per pass (mma N)
constexpr int kTileBytes = kRgRows * kTileCols * 2;   // 4096: the warp's 32x64 16-bit weight tile (128B rows, swizzled)
constexpr int kHeadBytes = kRgRows * kTok * 4;        // 1024 per token group: the warp's head-segment partial (fp32 [tokens][32 rows])
constexpr int kWarpBytes = kTileBytes + kHeadBytes;   // 5120 (one token group per pass); two groups: + kHeadBytes
constexpr int kOut = kRgRows * kTok;          // 256 outputs per (row group, token pass) == kThreads
static_assert(kOut == kThreads, "one thread per output in the cross-warp reduction");

struct Params {
    const uint2* dsign;
    const uint32_t* eptr;
    const uint4* ent;
    const float2* affine;
    const float* bias;
    const void* x;
    void* y;
    int64_t ldx, ldy;
    void* ws_part;        // [token pass][row group][slots][256] x {fp32 partial, valid tag}: all zero between kernels
    int M, N, K;
    uint32_t tiles_c, groups, tiles_per_group;
    uint32_t nblocks;     // row groups * tiles_c
    uint32_t rgs;
    uint32_t slots;       // partial slots per row group
    uint32_t q, rem;      // blocks per warp: nblocks / (grid * 8) and the remainder (the first `rem` warps take one more)
    unsigned long long* trace;   // kTrace builds only: [cta][warp][8] globaltimer stamps of this launch
};
}  // namespace dk

template <typename T>
__device__ __forceinline__ void dk_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
    if constexpr (std::is_same<T, __half>::value) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
}

__device__ __forceinline__ void dk_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr)
                 : "memory");
}

// Tile layout.  The warp's 32x64 tile has 128-byte rows of eight 16-byte chunks, chunk pc stored at
// row*128 + ((pc ^ (row & 7)) << 4) (conflict-free for the row-per-lane stores and for ldmatrix).  Columns are PERMUTED
// inside a row so that the activations never go through shared memory: chunk pc, 32-bit word t holds the logical
// columns 16t + 2pc + {0,1}.  ldmatrix then hands lane (g, t) of k16-step q exactly the columns 16t + 4q + {0,1} (a0/a1)
// and 16t + 4q + {2,3} (a2/a3) -- the columns of words 2q and 2q+1 of the 16 consecutive activations that lane loaded
// from global memory (token g, columns 16t..16t+15), which therefore ARE its B fragments.
//
// 64 sign bits of one weight row -> 64 exact {lo,hi} 16-bit values, 8 STS.128: PRMT byte-sign replicate + LOP3 select
// as in expand_row; `brow` already carries (row & 7) << 4, so chunk pc is at brow ^ (pc << 4).
__device__ __forceinline__ void dk_expand_dense(const uint2 sg, const uint32_t LL, const uint32_t DD, const uint32_t brow) {
#pragma unroll
    for (int jq = 0; jq < 4; ++jq) {                    // bit 2pc (+16 for odd t) of a sign word = byte pc>>2 (+2), bit 2*(pc&3)
        const int j = 2 * jq;
        const uint32_t xa0 = sg.x << (7 - j), xb0 = sg.x << (6 - j), xa1 = sg.y << (7 - j), xb1 = sg.y << (6 - j);
#pragma unroll
        for (int by0 = 0; by0 < 2; ++by0) {
            const int pc = jq + 4 * by0;
            uint32_t h[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t by = (uint32_t)(by0 + 2 * (t & 1));
                const uint32_t sel = 0x8888u | by | (by << 4) | ((4u + by) << 8) | ((4u + by) << 12);
                h[t] = sel_xor_and(LL, DD, (t >> 1) ? prmt(xa1, xb1, sel) : prmt(xa0, xb0, sel));
            }
            sts_v4(brow ^ ((uint32_t)pc << 4), h[0], h[1], h[2], h[3]);
        }
    }
}

// four salient entries: store each value's 16 bits at its (pre-swizzled) byte offset in the tile
__device__ __forceinline__ void dk_patch4(const uint32_t tile_s, const uint4 e) {
    sts_u16(tile_s + (e.x >> 16), (uint16_t)e.x);
    sts_u16(tile_s + (e.y >> 16), (uint16_t)e.y);
    sts_u16(tile_s + (e.z >> 16), (uint16_t)e.z);
    sts_u16(tile_s + (e.w >> 16), (uint16_t)e.w);
}

// kOcc = CTAs per SM the register budget is sized for: 3 -> 85 registers, 4 -> 64
__device__ __forceinline__ unsigned long long dk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// kNT = token groups of 8 per pass: 1 (M <= 8), or 2 (9..16 tokens against ONE expansion of each tile; the second group
// lives in its own variables, so the one-group instance compiles to exactly the code it had before)
template <typename T, int kOcc, bool kTrace = false, int kNT = 1>
__global__ void __launch_bounds__(dk::kThreads, kOcc) decode_mma_kernel(const dk::Params p) {
    using namespace dk;
    constexpr int kWarpBytes = dk::kWarpBytes + (kNT - 1) * kHeadBytes;
    constexpr int kTokP = kTok * kNT, kOutP = kOut * kNT;
    unsigned long long tr[8];
    if (kTrace) { tr[0] = dk_now(); }
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_hrg[kWarps], s_trg[kWarps];   // row group of each warp's head / tail partial (or kNone)
    __shared__ uint32_t s_meta[8];                // {rg_a, rg_b, head split?, slot, expected, tail split?, slot, expected}
    constexpr uint32_t kNone = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint8_t* wsm = smem + wid * kWarpBytes;
    const uint32_t tile_s = smem_u32(wsm);
    float* head_red = reinterpret_cast<float*>(wsm + kTileBytes);
    float* tail_red = reinterpret_cast<float*>(wsm);      // aliases the tile: written only after the warp's last block

    // ---- work partition (division-free): warp gw owns blocks [gw*q + min(gw,rem), ...), q = B / warps, rem = B % warps ----
    const uint32_t TC = p.tiles_c, q = p.q, rem = p.rem;
    auto wstart = [&](uint32_t gw) { return gw * q + min(gw, rem); };
    const uint32_t gw0 = blockIdx.x * kWarps;
    const uint32_t c_lo = wstart(gw0), c_hi = wstart(gw0 + kWarps);
    const uint32_t w_lo = wstart(gw0 + wid), w_hi = wstart(gw0 + wid + 1u);
    const int m0 = blockIdx.y * kTokP;

    // Programmatic dependent launch: the next kernel in the stream may start its own weight prefetch now; everything
    // below up to griddepcontrol.wait touches only immutable packed weights.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // first loads of this warp's stream: sign words and entry offsets of its run (eptr lives in registers, one per lane)
    const uint2* sgp = p.dsign + (size_t)w_lo * kRgRows + lane;
    uint2 sg = make_uint2(0, 0);
    uint32_t epr = 0;
    if (w_lo < w_hi) {
        sg = __ldg(sgp);
        if (w_lo + lane <= w_hi) epr = __ldg(p.eptr + w_lo + lane);
    }
    uint32_t rg = 0, kb = 0;
    if (w_lo < w_hi) { rg = w_lo / TC; kb = w_lo - rg * TC; }
    const uint32_t rg_first = rg;
    const bool grouped = p.groups > 1;
    uint32_t cur_g = grouped ? kb / p.tiles_per_group : 0u;
    float2 af = make_float2(0.f, 0.f);
    if (w_lo < w_hi) af = __ldg(p.affine + (size_t)(rg * kRgRows + lane) * p.groups + cur_g);

    if (lane == 0) { s_hrg[wid] = kNone; s_trg[wid] = kNone; }
    if (tid == 0) {                               // which of this CTA's row groups are shared with other CTAs, and how
        const uint32_t rg_a = c_lo / TC, rg_b = (c_hi - 1u) / TC;
        auto owner = [&](uint32_t b) {            // CTA whose run contains block b
            const uint32_t cut = rem * (q + 1u);
            const uint32_t gw = b < cut ? b / (q + 1u) : rem + (b - cut) / q;
            return gw / kWarps;
        };
        s_meta[0] = rg_a; s_meta[1] = rg_b;
        const bool hs = c_lo > rg_a * TC || c_hi < rg_a * TC + TC;
        const bool ts = rg_b != rg_a && c_hi < rg_b * TC + TC;
        s_meta[2] = hs; s_meta[5] = ts;
        if (hs) { const uint32_t f = owner(rg_a * TC); s_meta[3] = blockIdx.x - f; s_meta[4] = owner(rg_a * TC + TC - 1u) - f + 1u; }
        if (ts) { const uint32_t f = owner(rg_b * TC); s_meta[6] = blockIdx.x - f; s_meta[7] = owner(rg_b * TC + TC - 1u) - f + 1u; }
    }

    uint32_t ci = 0;                               // index of the current block in the eptr register chunk
    // the first min(n4, 64) units of a block sit in two register sets of h1 = ceil/2 and the rest: unit `lane` and unit
    // `h1 + lane` (the index builder deals entries to units so that each of the 8 patch stores is bank-conflict free)
    struct Ent { uint4 a, c; uint32_t eb, n4; };        // one block's entries: units `lane` and `h1 + lane`, offset, unit count
    auto ent_load = [&](Ent& E, uint32_t i) {           // i = index of the block in this warp's eptr register chunk
        E.eb = __shfl_sync(0xffffffffu, epr, i);
        E.n4 = __shfl_sync(0xffffffffu, epr, i + 1u) - E.eb;
        const uint32_t n1 = min(E.n4, 64u), h1 = (n1 + 1u) >> 1;
        const uint4* e = p.ent + (E.eb + lane);
        asm volatile("" : "+l"(e));                     // one address computation for both predicated loads
        if (lane < h1) E.a = __ldg(e);
        if (lane + h1 < n1) E.c = __ldg(e + h1);
    };
    Ent E0, E1;                                         // two blocks of entries in flight: each set is reloaded for the block
    E0.a = E0.c = E1.a = E1.c = make_uint4(0, 0, 0, 0); // after next right after its patch -- two blocks of cover
    E0.eb = E0.n4 = E1.eb = E1.n4 = 0;
    if (w_lo < w_hi) ent_load(E0, 0);
    if (w_lo + 1u < w_hi) ent_load(E1, 1);
    uint32_t LL, DD;
    {
        const uint32_t lo = bits16<T>(af.x), hi = bits16<T>(af.y);
        LL = lo | (lo << 16);
        DD = (lo ^ hi) * 0x10001u;
    }

    // ---- activation loads: lane -> (token = lane>>2, 16-column segment = lane&3) of the 8 x 64 block ----------
    const uint32_t xtok = lane >> 2, xseg = lane & 3u;
    const bool x_fast = ((p.ldx & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15u) == 0) && ((p.K & 63) == 0);
    const bool x_fast256 = x_fast && ((p.ldx & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 31u) == 0);   // one 32 B load
    const bool x_tok_ok = (m0 + (int)xtok) < p.M;
    // byte offset of (my token row, my 16-column segment, k-block 0); 32-bit (checked by the launcher) and opaque to the
    // compiler so it stays in a register instead of being recomputed every block
    uint32_t xoff_row = (uint32_t)(((int64_t)(m0 + (x_tok_ok ? (int)xtok : 0)) * p.ldx + 16 * xseg) * 2);
    asm volatile("" : "+r"(xoff_row));
    uint32_t xoff = xoff_row + kb * (kTileCols * 2u);          // loop-carried: advances one k-block per iteration
    const bool x_tok_ok2 = kNT == 2 && (m0 + kTok + (int)xtok) < p.M;                      // second token group
    uint32_t xoff_row2 = (uint32_t)(((int64_t)(x_tok_ok2 ? m0 + kTok + (int)xtok : m0) * p.ldx + 16 * xseg) * 2);
    if constexpr (kNT == 2) asm volatile("" : "+r"(xoff_row2));
    uint32_t xoff2 = xoff_row2 + kb * (kTileCols * 2u);
    const uint8_t* xbytes = reinterpret_cast<const uint8_t*>(p.x);
    // fast path (x_fast: 16 B aligned rows, K a multiple of 64): two unconditional 16 B loads -- rows past M read token
    // m0's row, whose products land in output columns that are never stored.  Anything else: bounds-checked elements.
    auto load_x_at = [&](uint32_t kblk, uint32_t off, bool tok_ok, uint4& xa, uint4& xb) {
        if (x_fast256) {                               // sm_100 256-bit load: half the L1 wavefronts of two 128-bit loads
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(xa.x), "=r"(xa.y), "=r"(xa.z), "=r"(xa.w), "=r"(xb.x), "=r"(xb.y), "=r"(xb.z), "=r"(xb.w)
                         : "l"(xbytes + off));
        } else if (x_fast) {
            const uint4* p4 = reinterpret_cast<const uint4*>(xbytes + off);
            xa = __ldg(p4);
            xb = __ldg(p4 + 1);
        } else {
            const uint16_t* qx = reinterpret_cast<const uint16_t*>(xbytes + off);
            const int col = (int)(kblk * kTileCols + 16 * xseg);
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cc = col + 2 * i;
                uint32_t v = 0;
                if (tok_ok && cc < p.K) v = (uint32_t)qx[2 * i];
                if (tok_ok && cc + 1 < p.K) v |= (uint32_t)qx[2 * i + 1] << 16;
                w[i] = v;
            }
            xa = make_uint4(w[0], w[1], w[2], w[3]);
            xb = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };
    auto load_x = [&](uint32_t kblk, uint4& xa, uint4& xb) { load_x_at(kblk, xoff, x_tok_ok, xa, xb); };

    // shared-memory addresses of this lane
    const uint32_t r7 = lane & 7u;
    const uint32_t brow = (tile_s + lane * 128u) | (r7 << 4);                     // my weight row; chunk c lives at brow ^ (c << 4)
    const uint32_t lm_row = (lane & 7u) + ((lane >> 3) & 1u) * 8u;                // A fragments (weights)
    const uint32_t lm_base0 = tile_s + lm_row * 128u + (((lane >> 4) ^ (lm_row & 7u)) << 4);
    const uint32_t g4 = lane >> 2, t4 = lane & 3u;
    uint32_t lm_q[4];                                                             // ldmatrix row address per k16 step
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
        lm_q[qi] = lm_base0 ^ ((uint32_t)qi << 5);
        asm volatile("" : "+r"(lm_q[qi]));                                        // keep in a register (no rematerialisation)
    }
    uint32_t brow_r = brow;
    asm volatile("" : "+r"(brow_r));

    float acc[2][4], acc2[2][4];                  // acc2: second token group (kNT == 2)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = acc2[0][i] = acc2[1][i] = 0.f;

    // fp32 [token][row] layout of one row group's outputs: index m*32 + r
    auto store_frag = [&](float* dst) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t r = 16u * h + g4;
            dst[(2u * t4) * kRgRows + r] = acc[h][0];
            dst[(2u * t4 + 1u) * kRgRows + r] = acc[h][1];
            dst[(2u * t4) * kRgRows + r + 8u] = acc[h][2];
            dst[(2u * t4 + 1u) * kRgRows + r + 8u] = acc[h][3];
            if constexpr (kNT == 2) {
                dst[kOut + (2u * t4) * kRgRows + r] = acc2[h][0];
                dst[kOut + (2u * t4 + 1u) * kRgRows + r] = acc2[h][1];
                dst[kOut + (2u * t4) * kRgRows + r + 8u] = acc2[h][2];
                dst[kOut + (2u * t4 + 1u) * kRgRows + r + 8u] = acc2[h][3];
            }
        }
    };

    // activations (and y, and the workspace) belong to the stream's earlier kernels: wait before the first touch
    if (kTrace) tr[1] = dk_now();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (kTrace) tr[2] = dk_now();
    uint4 xa = make_uint4(0, 0, 0, 0), xb = make_uint4(0, 0, 0, 0);
    uint4 xa2 = make_uint4(0, 0, 0, 0), xb2 = make_uint4(0, 0, 0, 0);
    if (w_lo < w_hi) load_x(kb, xa, xb);
    if constexpr (kNT == 2) { if (w_lo < w_hi) load_x_at(kb, xoff2, x_tok_ok2, xa2, xb2); }

    // Every stream is prefetched IN PLACE: a register set is reloaded right after its last use (sign words, activations:
    // one block of cover; salient entries, the stream that comes from DRAM with a dependent address: two sets, two blocks).
    auto do_block = [&](const uint32_t blk, Ent& E) {
        const bool more = blk + 1 < w_hi;
        if (grouped) {
            const uint32_t g = kb / p.tiles_per_group;
            if (g != cur_g) {
                cur_g = g;
                af = __ldg(p.affine + (size_t)(rg * kRgRows + lane) * p.groups + g);
                const uint32_t lo = bits16<T>(af.x), hi = bits16<T>(af.y);
                LL = lo | (lo << 16);
                DD = (lo ^ hi) * 0x10001u;
            }
        }

        __syncwarp();                                   // the previous block's ldmatrix reads are done
        dk_expand_dense(sg, LL, DD, brow_r);
        sgp += more ? kRgRows : 0;                      // unconditional reload (the last block re-reads itself): the load
        sg = __ldg(sgp);                                // must land in `sg` directly, not in a temporary that is moved at once
        __syncwarp();                                   // dense rows land before other lanes patch them
        {
            const uint32_t n1 = min(E.n4, 64u), h1 = (n1 + 1u) >> 1;
            if (lane < h1) dk_patch4(tile_s, E.a);
            if (lane + h1 < n1) dk_patch4(tile_s, E.c);
            for (uint32_t i = 64u + lane; i < E.n4; i += 32u) dk_patch4(tile_s, __ldg(p.ent + (E.eb + i)));   // rare: > 256 salient in a block
        }
        ++ci;
        if (blk + 2u < w_hi) {                          // this set's next block is the one after next
            if (ci + 2u > 31u) {                        // rare: refill the eptr registers (runs longer than 30 blocks)
                if (blk + 1u + lane <= w_hi) epr = __ldg(p.eptr + blk + 1u + lane);
                ci = 0;
            }
            ent_load(E, ci + 1u);
        }
        __syncwarp();                                   // the tile is complete and visible to the whole warp

        // tensor cores: A fragments by ldmatrix from the tile, B fragments straight from the activation registers
        {
            const uint32_t xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const uint32_t xw2[8] = {xa2.x, xa2.y, xa2.z, xa2.w, xb2.x, xb2.y, xb2.z, xb2.w};
#pragma unroll
            for (int qi = 0; qi < 4; ++qi) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t a0, a1, a2, a3;
                    dk_ldsm4(lm_q[qi] + (uint32_t)h * 2048u, a0, a1, a2, a3);
                    dk_mma<T>(acc[h], a0, a1, a2, a3, xw[2 * qi], xw[2 * qi + 1]);
                    if constexpr (kNT == 2) dk_mma<T>(acc2[h], a0, a1, a2, a3, xw2[2 * qi], xw2[2 * qi + 1]);
                }
            }
        }
        ++kb;
        const bool rg_end = kb == TC;
        xoff = rg_end ? xoff_row : xoff + kTileCols * 2u;
        load_x(rg_end ? 0u : kb, xa, xb);               // unconditional: the address after the last block is still inside x
        if constexpr (kNT == 2) {
            xoff2 = rg_end ? xoff_row2 : xoff2 + kTileCols * 2u;
            load_x_at(rg_end ? 0u : kb, xoff2, x_tok_ok2, xa2, xb2);
        }

        // ---- end of this warp's part of the row group? -----------------------------------------------------
        if (rg_end || !more) {
            const bool whole = rg_end && (w_lo <= rg * TC);      // this warp saw every k-block of the row group
            if (whole) {                                         // finish it here: + bias, round, store
                __syncwarp();
                store_frag(tail_red);
                __syncwarp();
                const int orow = (int)(rg * kRgRows + lane);
                if (orow < p.N) {
                    const float bv = p.bias ? p.bias[orow] : 0.f;
#pragma unroll
                    for (int m = 0; m < kTokP; ++m)
                        if (m0 + m < p.M)
                            reinterpret_cast<T*>(p.y)[(int64_t)(m0 + m) * p.ldy + orow] = from_f32<T>(bv + tail_red[m * kRgRows + lane]);
                }
            } else if (rg == rg_first) {
                store_frag(head_red);                            // head partial: its own buffer, the warp may go on
                if (lane == 0) s_hrg[wid] = rg;
            } else {
                __syncwarp();
                store_frag(tail_red);                            // tail partial: last thing this warp does
                if (lane == 0) s_trg[wid] = rg;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = acc2[0][i] = acc2[1][i] = 0.f;
            if (rg_end) {
                kb = 0;
                ++rg;
                if (more) {
                    cur_g = 0;
                    af = __ldg(p.affine + (size_t)(rg * kRgRows + lane) * p.groups);
                    const uint32_t lo = bits16<T>(af.x), hi = bits16<T>(af.y);
                    LL = lo | (lo << 16);
                    DD = (lo ^ hi) * 0x10001u;
                }
            }
        }
    };
    for (uint32_t blk = w_lo; blk < w_hi; blk += 2u) {   // unrolled by two: the entry sets alternate without register moves
        do_block(blk, E0);
        if (blk + 1u < w_hi) do_block(blk + 1u, E1);
    }

    // ---- cross-warp reduction in shared memory; row groups shared with other CTAs go through the workspace ----------
    if (kTrace) tr[3] = dk_now();
    __syncthreads();
    if (kTrace) tr[4] = dk_now();
    if (c_lo >= c_hi) return;
    auto trace_out = [&]() {
        if (kTrace && lane == 0) {
            tr[6] = dk_now();
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            tr[7] = ((unsigned long long)smid << 32) | (w_hi - w_lo);
            unsigned long long* o = p.trace + ((size_t)blockIdx.x * kWarps + wid) * 8u;
            for (int i = 0; i < 8; ++i) o[i] = tr[i];
        }
    };
    // Two warps finish the CTA: thread t < 64 owns four consecutive outputs per token group -- token 8u + (t>>3), rows
    // 4*(t&7)..+3 of the row group, i.e. float4 number 64u + t of every [token][row] partial buffer -- so each partial costs
    // one LDS.128 per thread and group.
    if (tid >= 64u) { if (kTrace) tr[5] = tr[4]; trace_out(); return; }
    const uint32_t rg_a = s_meta[0], rg_b = s_meta[1];
    const uint32_t om = tid >> 3, or4 = (tid & 7u) * 4u;
    uint32_t hrg[kWarps], trg[kWarps];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { hrg[w] = s_hrg[w]; trg[w] = s_trg[w]; }
    const bool hs = s_meta[2] != 0u, ts = s_meta[5] != 0u;
#pragma unroll
    for (int u = 0; u < kNT; ++u) {                          // one round per token group: float4 number 64u + t of the buffers
        T* yout = reinterpret_cast<T*>(p.y) + (int64_t)(m0 + kTok * u + om) * p.ldy;
        const bool tok_ok = (m0 + kTok * u + (int)om) < p.M;
        auto emit = [&](uint32_t r, const float4 v) {            // + bias, round, store the four outputs of row group r
            const int orow = (int)(r * kRgRows + or4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
            if (tok_ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (orow + j < p.N) yout[orow + j] = from_f32<T>((p.bias ? p.bias[orow + j] : 0.f) + vv[j]);
            }
        };
        float4 v_split[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};   // head / tail row group (when shared)
        for (uint32_t r = rg_a; r <= rg_b; ++r) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            bool any = false;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {                   // fixed order: warp 0's partial first
                if (hrg[w] == r) {
                    const float4 a4 = reinterpret_cast<const float4*>(smem + w * kWarpBytes + kTileBytes)[64 * u + tid];
                    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w; any = true;
                }
                if (trg[w] == r) {
                    const float4 a4 = reinterpret_cast<const float4*>(smem + w * kWarpBytes)[64 * u + tid];
                    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w; any = true;
                }
            }
            if (!any) continue;                                  // stored by the single warp that owned it
            const bool split = (r == rg_a && s_meta[2]) || (r == rg_b && s_meta[5]);
            if (!split) emit(r, v);                              // the whole row group lives in this CTA
            else if (r == rg_a) v_split[0] = v;
            else v_split[1] = v;
        }
        if (kTrace) tr[5] = dk_now();
        if (!hs && !ts) continue;
        // Row groups shared with other CTAs.  Every contributor but the last parks its partial in its own slot as 64-bit
        // stores {value, valid tag}: data and flag travel together, so there is no fence, no counter and no barrier.  The
        // last contributor (highest CTA index, so everything it waits for was scheduled before it) polls the slots, sums them
        // in CTA order (deterministic), clears them for the next kernel, and writes y.
        unsigned long long* ws = reinterpret_cast<unsigned long long*>(p.ws_part);
#pragma unroll
        for (int f = 1; f >= 0; --f) {                           // the tail group first: this CTA is never its last contributor
            if (!(f == 0 ? hs : ts)) continue;
            const uint32_t r = f == 0 ? rg_a : rg_b, slot = s_meta[f == 0 ? 3 : 6], expected = s_meta[f == 0 ? 4 : 7];
            unsigned long long* part = ws + (((size_t)blockIdx.y * p.rgs + r) * p.slots) * kOutP + kOut * u + tid * 4u;
            const float4 v = v_split[f];
            const float vv[4] = {v.x, v.y, v.z, v.w};
            if (slot + 1u < expected) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long w64 = (1ull << 32) | (unsigned long long)__float_as_uint(vv[j]);
                    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(part + (size_t)slot * kOutP + j), "l"(w64) : "memory");
                }
            } else {
                float sum[4] = {0.f, 0.f, 0.f, 0.f};
                for (uint32_t k = 0; k + 1u < expected; ++k) {
                    unsigned long long w64[4];
                    do {                                         // four independent loads in flight per poll
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w64[j]) : "l"(part + (size_t)k * kOutP + j) : "memory");
                    } while (((w64[0] & w64[1] & w64[2] & w64[3]) >> 32) == 0ull);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sum[j] += __uint_as_float((uint32_t)w64[j]);
                        asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(part + (size_t)k * kOutP + j), "l"(0ull) : "memory");
                    }
                }
                emit(r, make_float4(sum[0] + vv[0], sum[1] + vv[1], sum[2] + vv[2], sum[3] + vv[3]));
            }
        }
    }
    trace_out();
}

// ---- decode index construction (one-time, from the packed form) ---------------------------------------------------
// pass 1: 16-byte units of salient entries per block, row-group-major order; scanned in place afterwards
__global__ void decode_index_count_kernel(const uint32_t* __restrict__ vptr, uint32_t* __restrict__ eptr, uint32_t tiles_c,
                                          uint32_t nblocks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint32_t rg = i / tiles_c, kb = i - rg * tiles_c;
    const uint32_t tr = rg / kRgPerTile, rgi = rg % kRgPerTile;
    const size_t old = ((size_t)tr * tiles_c + kb) * kRgPerTile + rgi;
    const uint32_t cnt = vptr[old + 1] - vptr[old];
    eptr[i] = (cnt + 3u) / 4u;
}

// pass 2: sign words and entries. CTA = one 128x64 plane tile, warp = one 32-row group, lane = row.
// Entry order inside a block is chosen for the kernel's patch stores: store j of register set a writes the entries at
// slots 4*(a*h1 + lane) + j, lane = 0..31 -- a "group" of up to 32 entries that should fall in 32 different shared-memory
// banks.  Entries are ranked by (bank, row, column) and rank k goes to group k % 8, position k / 8: the <= 8 entries of
// one bank land in 8 different groups.  (A row touches each bank at most twice, so per-bank counts come from two
// ballots.)  Blocks with more than 256 entries keep ranks >= 256 in rank order behind the first 64 units.
__global__ void __launch_bounds__(128) decode_index_fill_kernel(const uint4* __restrict__ planes, const uint32_t* __restrict__ vptr,
                                                                const uint16_t* __restrict__ vals, const uint32_t* __restrict__ eptr,
                                                                uint32_t tiles_c, uint2* __restrict__ dsign,
                                                                uint32_t* __restrict__ ent) {
    const uint32_t lane = threadIdx.x & 31u, rgi = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const uint32_t tr = (uint32_t)(tile / tiles_c), kb = (uint32_t)(tile % tiles_c);
    const uint4 pw = planes[tile * kTileRows + rgi * kRgRows + lane];
    const size_t blk = (size_t)(tr * kRgPerTile + rgi) * tiles_c + kb;
    dsign[blk * kRgRows + lane] = make_uint2(pw.x, pw.y);
    const size_t old = tile * kRgPerTile + rgi;
    const uint32_t vbase = vptr[old], cnt = vptr[old + 1] - vbase;
    if (cnt == 0) return;                                   // warp-uniform
    const uint32_t mine = (uint32_t)(__popc(pw.z) + __popc(pw.w));
    const uint32_t voff = warp_excl_scan(mine, lane);       // my row's first value in the packed (row, column) order
    uint32_t* dst = ent + (size_t)eptr[blk] * 4u;
    const uint32_t n4 = (cnt + 3u) / 4u, n1 = min(n4, 64u), h1 = (n1 + 1u) >> 1;

    // my row's entries and the banks they hit (each bank at most twice per row: two columns per 32-bit word)
    uint32_t my_e[64];
    uint32_t m1 = 0, m2 = 0, ne = 0;
#pragma unroll 1
    for (int wd = 0; wd < 2; ++wd) {
        uint32_t m = wd ? pw.w : pw.z;
        while (m) {
            const uint32_t col = (uint32_t)(__ffs(m) - 1) + 32u * wd;
            m &= m - 1u;
            const uint32_t pc = (col >> 1) & 7u;                                             // chunk, word, half: see "Tile layout"
            const uint32_t pos = lane * 128u + ((pc ^ (lane & 7u)) << 4) + (col >> 4) * 4u + (col & 1u) * 2u;
            my_e[ne] = (pos << 16) | (uint32_t)vals[vbase + voff + ne];
            ++ne;
            const uint32_t bit = 1u << ((pos >> 2) & 31u);
            m2 |= m1 & bit;
            m1 |= bit;
        }
    }
    // rank of my first entry in every bank: entries of lower banks + entries of this bank in lower rows
    uint16_t start[32];
    uint32_t base = 0;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll 1
    for (int bnk = 0; bnk < 32; ++bnk) {
        const uint32_t b1 = __ballot_sync(0xffffffffu, (m1 >> bnk) & 1u), b2 = __ballot_sync(0xffffffffu, (m2 >> bnk) & 1u);
        start[bnk] = (uint16_t)(base + __popc(b1 & lt) + __popc(b2 & lt));
        base += __popc(b1) + __popc(b2);
    }
    // every slot first gets a copy of one real entry (padding must be an idempotent store), then the real entries land
    const uint32_t first_lane = (uint32_t)__ffs(__ballot_sync(0xffffffffu, ne > 0)) - 1u;
    const uint32_t pad = __shfl_sync(0xffffffffu, ne ? my_e[0] : 0u, first_lane);
    for (uint32_t sl = lane; sl < n4 * 4u; sl += 32u) dst[sl] = pad;
    __syncwarp();
#pragma unroll 1
    for (uint32_t i = 0; i < ne; ++i) {
        const uint32_t e = my_e[i];
        const uint32_t k = start[(e >> 18) & 31u]++;
        uint32_t slot = k;
        if (k < 256u) {
            const uint32_t g = k & 7u, idx = k >> 3;
            slot = 4u * ((g >> 2) * h1 + idx) + (g & 3u);
        }
        dst[slot] = e;
    }
}

void launch_scan_counts(uint32_t* v, int64_t n, cudaStream_t s);   // pbllm_pack.cu

int launch_decode_index_count(const Layer& L, uint32_t* eptr, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)(L.tiles_r * kRgPerTile * L.tiles_c);
    decode_index_count_kernel<<<(nblocks + 255u) / 256u, 256, 0, s>>>(L.vptr, eptr, (uint32_t)L.tiles_c, nblocks);
    int rc = check_cuda(cudaGetLastError(), "decode_index_count launch");
    if (rc) return rc;
    launch_scan_counts(eptr, nblocks, s);
    count_launch(2);
    return check_cuda(cudaGetLastError(), "decode_index scan launch");
}

int launch_decode_index_fill(const Layer& L, const uint32_t* eptr, uint2* dsign, uint32_t* ent, cudaStream_t s) {
    decode_index_fill_kernel<<<(unsigned)(L.tiles_r * L.tiles_c), 128, 0, s>>>(L.planes, L.vptr, (const uint16_t*)L.vals, eptr,
                                                                               (uint32_t)L.tiles_c, dsign, ent);
    count_launch();
    return check_cuda(cudaGetLastError(), "decode_index_fill launch");
}

// ---- host side -------------------------------------------------------------------------------------------------
static int dk_ctas_per_sm() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PBL_DK_CTAS");
        v = (e && *e) ? atoi(e) : 3;                  // measured on B200: 3 CTAs (24 warps) per SM is the fastest grid
        if (v < 1) v = 1;
        if (v > 16) v = 16;
    }
    return v;
}

static int dk_occupancy() {      // which register budget / launch-bounds variant to launch
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PBL_DK_OCC");
        v = (e && *e) ? atoi(e) : 3;
        if (v != 4) v = 3;
    }
    return v;
}

static int dk_num_sms() {
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev] > 0 ? sms[dev] : 148;
}

// profiling aid (tools/decode_trace.py): when a device buffer is registered, launches go through the kTrace build and
// launch i writes its per-warp globaltimer stamps at trace + i * kTraceStride
static unsigned long long* g_trace = nullptr;
static size_t g_trace_launches = 0, g_trace_next = 0;
constexpr size_t kTraceStride = 16u * 148u * dk::kWarps * 8u;     // u64 per launch (grid <= 16 CTAs per SM)
void decode_set_trace(void* buf, size_t bytes) {
    g_trace = reinterpret_cast<unsigned long long*>(buf);
    g_trace_launches = buf ? bytes / (kTraceStride * 8u) : 0;
    g_trace_next = 0;
}

struct DecodeGeom { uint32_t nblocks, rgs, grid, slots, passes, q, rem, nt; size_t ws_bytes; };

static DecodeGeom decode_geom_raw(int64_t tiles_r, int64_t tiles_c, int64_t M, uint32_t want);

static DecodeGeom decode_geom(const Layer& L, int64_t M) {
    const int ctas = M > dk::kTok ? (dk_ctas_per_sm() < 2 ? dk_ctas_per_sm() : 2) : dk_ctas_per_sm();   // 16-token passes: 2 CTAs per SM
    return decode_geom_raw(L.tiles_r, L.tiles_c, M, (uint32_t)(dk_num_sms() * ctas));
}

// host-only: the launch plan of the decode kernel for an N x K layer on a device with `sms` SMs (pbl_decode_plan)
void decode_plan(int64_t N, int64_t K, int64_t M, int sms, int ctas_per_sm, uint32_t out[8]) {
    const int64_t tr = (N + kTileRows - 1) / kTileRows, tc = (K + kTileCols - 1) / kTileCols;
    const DecodeGeom g = decode_geom_raw(tr, tc, M, (uint32_t)(sms * ctas_per_sm));
    out[0] = g.nblocks; out[1] = g.rgs; out[2] = g.grid; out[3] = g.passes; out[4] = g.q; out[5] = g.rem; out[6] = g.slots;
    out[7] = (uint32_t)(g.ws_bytes >> 10);
}

static DecodeGeom decode_geom_raw(int64_t tiles_r, int64_t tiles_c, int64_t M, uint32_t want) {
    struct { int64_t tiles_r, tiles_c; } L = {tiles_r, tiles_c};
    DecodeGeom g;
    g.rgs = (uint32_t)(L.tiles_r * kRgPerTile);
    g.nblocks = g.rgs * (uint32_t)L.tiles_c;
    const uint32_t cap = g.nblocks / dk::kWarps > 0 ? g.nblocks / dk::kWarps : 1u;   // at least one block per warp
    g.grid = cap < want ? cap : want;
    g.q = g.nblocks / (g.grid * dk::kWarps);
    g.rem = g.nblocks % (g.grid * dk::kWarps);
    // a CTA's run is at least 8*q blocks long, so a row group (tiles_c blocks) meets at most this many CTAs
    g.slots = g.q ? ((uint32_t)L.tiles_c + 8u * g.q - 1u) / (8u * g.q) + 1u : 2u;
    g.nt = M > dk::kTok ? 2u : 1u;                                                     // token groups of 8 per pass
    g.passes = (uint32_t)((M + dk::kTok * g.nt - 1) / (dk::kTok * g.nt));
    g.ws_bytes = (size_t)g.passes * g.rgs * g.slots * dk::kOut * g.nt * 8u;
    return g;
}

bool decode_supported(const Layer& L, int64_t ldx, int64_t M) {
    if (!L.dsign || !L.eptr || !L.ent) return false;
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) return false;
    if (M <= 0 || M > 64) return false;
    if (ldx <= 0 || (uint64_t)M * (uint64_t)ldx * 2u >= (1ull << 31)) return false;   // 32-bit activation offsets
    return true;
}

size_t decode_workspace_bytes(const Layer& L, int64_t M) {
    if (!decode_supported(L, L.K, M)) return 0;
    return decode_geom(L, M).ws_bytes;
}

template <typename T, int kOcc, int kNT>
static int launch_decode_t(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws, cudaStream_t s) {
    const DecodeGeom g = decode_geom(L, M);
    const int smem = dk::kWarps * (dk::kWarpBytes + (kNT - 1) * dk::kHeadBytes);
    static int attr_smem_dev[64] = {};   // function attributes are per device
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    if (attr_smem_dev[cur_dev] < smem) {
        int rc = check_cuda(cudaFuncSetAttribute(decode_mma_kernel<T, kOcc, false, kNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                            "cudaFuncSetAttribute(decode smem)");
        if (rc) return rc;
        attr_smem_dev[cur_dev] = smem;
    }
    dk::Params p;
    p.dsign = L.dsign; p.eptr = L.eptr; p.ent = reinterpret_cast<const uint4*>(L.ent);
    p.affine = L.affine; p.bias = L.bias; p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy;
    p.ws_part = ws;
    p.M = (int)M; p.N = (int)L.N; p.K = (int)L.K;
    p.tiles_c = (uint32_t)L.tiles_c; p.groups = (uint32_t)L.groups; p.tiles_per_group = (uint32_t)L.tiles_per_group;
    p.nblocks = g.nblocks; p.rgs = g.rgs; p.slots = g.slots; p.q = g.q; p.rem = g.rem;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.grid, g.passes);
    cfg.blockDim = dim3(dk::kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static int pdl = -1;
    if (pdl < 0) { const char* e = getenv("PBL_PDL"); pdl = (e && *e) ? atoi(e) : 1; }
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t le;
    if (g_trace && g_trace_next < g_trace_launches && kOcc == 3 && kNT == 1) {
        p.trace = g_trace + (g_trace_next++) * kTraceStride;
        static bool tattr = false;
        if (!tattr) { cudaFuncSetAttribute(decode_mma_kernel<T, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); tattr = true; }
        le = cudaLaunchKernelEx(&cfg, decode_mma_kernel<T, 3, true>, p);
    } else {
        p.trace = nullptr;
        le = cudaLaunchKernelEx(&cfg, decode_mma_kernel<T, kOcc, false, kNT>, p);
    }
    count_launch();
    return check_cuda(le, "decode launch");
}

// ws == nullptr: take a transient workspace from the stream-ordered pool and zero it (slower: the memset
// sits between consecutive decode kernels); callers on the hot path pass a persistent zero-initialised workspace.
int launch_decode(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
    const DecodeGeom g = decode_geom(L, M);
    void* own = nullptr;
    if (ws) {
        if (ws_bytes < g.ws_bytes) { set_error("decode workspace too small: %zu < %zu bytes", ws_bytes, g.ws_bytes); return PBL_ERR_SHAPE; }
        if (reinterpret_cast<uintptr_t>(ws) & 15u) { set_error("decode workspace must be 16 B aligned"); return PBL_ERR_ALIGN; }
    } else {
        int rc = check_cuda(cudaMallocAsync(&own, g.ws_bytes, s), "cudaMallocAsync(decode workspace)");
        if (rc) return rc;
        rc = check_cuda(cudaMemsetAsync(own, 0, g.ws_bytes, s), "cudaMemsetAsync(decode workspace)");
        if (rc) { cudaFreeAsync(own, s); return rc; }
        ws = own;
    }
    int rc;
    if (g.nt == 2)                                       // 9..16 tokens in one pass: 128 registers, 2 CTAs per SM
        rc = (L.dtype == PBL_F16) ? launch_decode_t<__half, 2, 2>(L, x, ldx, y, ldy, M, ws, s)
                                  : launch_decode_t<__nv_bfloat16, 2, 2>(L, x, ldx, y, ldy, M, ws, s);
    else if (dk_occupancy() == 4)
        rc = (L.dtype == PBL_F16) ? launch_decode_t<__half, 4, 1>(L, x, ldx, y, ldy, M, ws, s)
                                  : launch_decode_t<__nv_bfloat16, 4, 1>(L, x, ldx, y, ldy, M, ws, s);
    else
        rc = (L.dtype == PBL_F16) ? launch_decode_t<__half, 3, 1>(L, x, ldx, y, ldy, M, ws, s)
                                  : launch_decode_t<__nv_bfloat16, 3, 1>(L, x, ldx, y, ldy, M, ws, s);
    if (own) {
        const int rf = check_cuda(cudaFreeAsync(own, s), "cudaFreeAsync(decode workspace)");
        if (!rc) rc = rf;
    }
    return rc;
}

}  // namespace pbl
