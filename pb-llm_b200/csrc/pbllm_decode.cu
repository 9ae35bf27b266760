// Decode kernel (M <= 16 tokens per pass, fp16 / bf16): the HBM-bound regime of the bit-plane forward.
//
//   y[m][i] = sum_g ( mid_ig * sum_{j in g} x[m][j] + half_ig * sum_{j in g} t_ij x[m][j] ) + b_i
//
// with t_ij = +-1 at binarized positions (the sign plane; mid = (lo+hi)/2, half = (hi-lo)/2 of the row's two levels) and
// t_ij = tau_ij = (w_ij - mid)/half at salient positions (block-stream layout, pbllm_stream.cuh) -- the same w_sim as every
// other kernel of the library, summed in fp32 on the tensor cores.  What bounds a 2-microsecond kernel is instructions and
// shared-memory wavefronts per weight:
//
//  * ONE instruction pair per two weights.  Each warp keeps a 4 KB tile that holds +1.0 everywhere; a block's salient
//    entries are patched into it (one LEA + one STS.U16 each, lane-balanced, bank-spread at pack time).  The sign words are
//    stored in MMA-fragment order, so after ldmatrix a lane turns its 64 sign bits and the tile words into the 32 A-fragment
//    registers of the block's eight mma.sync.m16n8k16 with one shift + one LOP3 each: ((w << rho) & 0x80008000) ^ tile.
//    Nothing of the 90 % binarized weights is ever written to shared memory; the entries are then reset to +1.0 (one STS
//    each).  Shared-memory traffic per weight drops from write + patch + read of the full tile to read + 2 x patch.
//  * the levels are applied to the fp32 accumulators when a (row group, group) segment ends; sum_j x[m][j] comes from the
//    tensor cores too (an all-ones A fragment), and only for layers that have an asymmetric level pair.
//  * activations never touch shared memory: lane (token g, segment t) loads its 16 consecutive activations with one
//    256-bit load; the column permutation of the layout makes those registers the B fragments.  They are loaded one block
//    ahead into a second register set.
//  * warp-granular stream-K (unchanged): the blocks of the layer are dealt out in contiguous, equal (+-1) runs to the warps
//    of a fixed grid, partial row groups are reduced across the CTA's warps in shared memory and across CTAs through a
//    small zeroed workspace with tagged 64-bit slots, summed in CTA order (deterministic).
//  * the packed stream (sign words + entries of a warp's next blocks) is moved by the bulk-copy engine (cp.async.bulk ->
//    UBLKCP) into a per-warp shared-memory ring, three blocks deep, completed on mbarriers: no registers are spent on
//    prefetch, so 24 warps per SM fit.  The first copies are issued before griddepcontrol.wait: under programmatic
//    dependent launch the packed stream of layer i+1 is in flight while layer i still computes.
#include <cstdlib>
#include <type_traits>

#include "pbllm_stream.cuh"
#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace dk {
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTok = 8;                       // tokens per group (mma N)
constexpr int kTileBytes = kRgRows * kTileCols * 2;   // 4096: the warp's 32x64 16-bit tile (128B rows, swizzled)
constexpr int kHeadBytes = kRgRows * kTok * 4;        // 1024 per token group: the warp's head-segment partial (fp32 [tokens][32 rows])
// per-warp ring of the packed stream: kStages blocks in flight, each = 256 B of sign words + up to kEntCap 16-byte entry units
// (blocks with more entries read the rest straight from global memory), filled by cp.async.bulk, completed on mbarriers
constexpr int kStages = 3;
constexpr int kEntCap = 64;
constexpr int kStageBytes = 256 + kEntCap * 16;       // 1280
constexpr int kRingBytes = kStages * kStageBytes + 128;  // + the stages' mbarriers, padded: the tile's XOR-swizzled addressing
                                                         // needs every warp's region to start 128-byte aligned
constexpr int kWarpBytes = kTileBytes + kHeadBytes + kRingBytes;   // 9088 (one token group per pass); two groups: + kHeadBytes
static_assert(kWarpBytes % 128 == 0 && kHeadBytes % 128 == 0, "per-warp regions must keep the tile 128-byte aligned");
constexpr int kOut = kRgRows * kTok;          // 256 outputs per (row group, token group) == kThreads
static_assert(kOut == kThreads, "one thread per output in the cross-warp reduction");

struct Params {
    const uint2* fsign;
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
    uint32_t has_mid;     // some (row, group) has lo != -hi: the mid * sum(x) term is needed
    unsigned long long* trace;   // kTrace builds only: [cta][warp][8] globaltimer stamps of this launch
    // row-sharded execution (kPush): this rank's [M, N] slice is stored straight into the y of every rank of the node
    // (peer-mapped pointers, already offset to the slice's first column), and the all-gather's completion is an in-kernel
    // flag exchange: the last CTA of the grid publishes a new epoch in every rank's flag array, the next pushed kernel
    // waits for all ranks' flags of the previous one before it touches its activations.
    void* y_dst[PBL_MAX_PEERS];
    uint32_t* flag_dst[PBL_MAX_PEERS];   // flag array (u32 [PBL_MAX_PEERS]) of every destination rank; we write entry [rank]
    uint32_t* flag_local;                // this rank's own flag array
    uint32_t* sync_ctr;                  // local: [0] = warps done (the last one resets it), [1] = epoch of the last push
    uint32_t n_dst, rank, wait_prev;
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

// Tile layout (st::tile_slot).  The warp's 32x64 tile has 128-byte rows of eight 16-byte chunks, chunk pc
// stored at row*128 + ((pc ^ (row & 7)) << 4) (conflict-free for ldmatrix).  Columns are PERMUTED inside a row so that the
// activations never go through shared memory: chunk pc, 32-bit word t holds the logical columns 16t + 2pc + {0,1}.
// ldmatrix then hands lane (g, t) of k16-step q exactly the columns 16t + 4q + {0,1} (a0/a1) and 16t + 4q + {2,3} (a2/a3)
// -- the columns of words 2q and 2q+1 of the 16 consecutive activations that lane loaded from global memory (token g,
// columns 16t..16t+15), which therefore ARE its B fragments.  The dense A fragments built from the sign words use the same
// column assignment (st::sign_pos).

// {+1,+1} in the layer's 16-bit type: the rest state of every tile word
template <typename T> __device__ __forceinline__ constexpr uint32_t dk_one2() { return std::is_same<T, __half>::value ? 0x3C003C00u : 0x3F803F80u; }

// four salient entries: store each tau at its slot of the tile (entry >> 20 = 2 * slot) / reset the slots to +1.0
__device__ __forceinline__ void dk_patch4(const uint32_t tile_s, const uint4 e) {
    sts_u16(tile_s + (e.x >> 20), (uint16_t)e.x);
    sts_u16(tile_s + (e.y >> 20), (uint16_t)e.y);
    sts_u16(tile_s + (e.z >> 20), (uint16_t)e.z);
    sts_u16(tile_s + (e.w >> 20), (uint16_t)e.w);
}
__device__ __forceinline__ void dk_unpatch4(const uint32_t tile_s, const uint4 e, const uint16_t one) {
    sts_u16(tile_s + (e.x >> 20), one);
    sts_u16(tile_s + (e.y >> 20), one);
    sts_u16(tile_s + (e.z >> 20), one);
    sts_u16(tile_s + (e.w >> 20), one);
}

__device__ __forceinline__ void dk_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint2 dk_lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 dk_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// one lane of the (converged) warp
__device__ __forceinline__ bool dk_elect() {
    uint32_t r;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(r));
    return r != 0u;
}

__device__ __forceinline__ unsigned long long dk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// kNT = token groups of 8 per pass: 1 (M <= 8), or 2 (9..16 tokens against ONE expansion of each block; the second group
// lives in its own variables).  kLean = the common case compiled without its branches: one group per row, activations
// 32-byte aligned with K a multiple of 64 (the launcher checks).
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* a) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* a, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}

// Which of a CTA's row groups are shared with other CTAs, and how: {rg_a, rg_b, head split?, slot, expected, tail split?,
// slot, expected}.  Units are what the kernel deals out to its warps (blocks, or pairs of blocks): the CTA owns units
// [c_lo, c_hi), a row group has TC of them, warp gw of the grid owns q (+1 for the first rem warps).
__device__ __forceinline__ void dk_meta(uint32_t* s_meta, uint32_t c_lo, uint32_t c_hi, uint32_t TC, uint32_t q, uint32_t rem) {
    constexpr uint32_t kWarps = dk::kWarps;
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

// The end of a decode kernel, shared by its variants: cross-warp reduction in shared memory; row groups shared with other
// CTAs go through the workspace.  Warp w's region starts at smem + w * kWarpStride: its tail partial at offset 0 (aliasing
// the tile), its head partial at kHeadOff, both fp32 [token][row].
template <typename T, int kNT, bool kPush, bool kTrace, int kWarpStride, int kHeadOff>
__device__ __forceinline__ void dk_finish(const dk::Params& p, uint8_t* smem, const uint32_t* s_hrg, const uint32_t* s_trg,
                                          const uint32_t* s_meta, const uint32_t c_lo, const uint32_t c_hi, const int m0,
                                          unsigned long long (&tr)[8], const uint32_t my_units) {
    using namespace dk;
    constexpr int kOutP = kOut * kNT;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    // ---- cross-warp reduction in shared memory; row groups shared with other CTAs go through the workspace ----------
    if (kTrace) tr[3] = dk_now();
    __syncthreads();
    if (kTrace) tr[4] = dk_now();
    auto trace_out = [&]() {
        if (kTrace && lane == 0) {
            tr[6] = dk_now();
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            tr[7] = ((unsigned long long)smid << 32) | my_units;
            unsigned long long* o = p.trace + ((size_t)blockIdx.x * kWarps + wid) * 8u;
            for (int i = 0; i < 8; ++i) o[i] = tr[i];
        }
    };
    // kPush: when every warp that stores outputs (warps 0 and 1 of every CTA, below; the loop's stores are ordered before
    // them by the barrier above) has passed its system-scope fence, the last one publishes the new epoch to every rank.
    auto publish = [&]() {
        if constexpr (kPush) {
            __syncwarp();
            if (lane == 0) {
                __threadfence_system();
                const uint32_t total = 2u * gridDim.x * gridDim.y;
                if (atomicAdd(p.sync_ctr, 1u) == total - 1u) {
                    p.sync_ctr[0] = 0u;
                    const uint32_t epoch = p.sync_ctr[1] + 1u;
                    p.sync_ctr[1] = epoch;
                    __threadfence_system();
                    for (uint32_t d = 0; d < p.n_dst; ++d) st_release_sys(p.flag_dst[d] + p.rank, epoch);
                }
            }
        }
    };
    // Two warps finish the CTA: thread t < 64 owns four consecutive outputs per token group -- token 8u + (t>>3), rows
    // 4*(t&7)..+3 of the row group, i.e. float4 number 64u + t of every [token][row] partial buffer -- so each partial costs
    // one LDS.128 per thread and group.
    if (tid >= 64u) { if (kTrace) tr[5] = tr[4]; trace_out(); return; }
    if (c_lo >= c_hi) { publish(); return; }
    const uint32_t rg_a = s_meta[0], rg_b = s_meta[1];
    const uint32_t om = tid >> 3, or4 = (tid & 7u) * 4u;
    uint32_t hrg[kWarps], trg[kWarps];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { hrg[w] = s_hrg[w]; trg[w] = s_trg[w]; }
    const bool hs = s_meta[2] != 0u, ts = s_meta[5] != 0u;
#pragma unroll
    for (int u = 0; u < kNT; ++u) {                          // one round per token group: float4 number 64u + t of the buffers
        const int64_t yoff = (int64_t)(m0 + kTok * u + om) * p.ldy;
        const bool tok_ok = (m0 + kTok * u + (int)om) < p.M;
        auto emit = [&](uint32_t r, const float4 v) {            // + bias, round, store the four outputs of row group r
            const int orow = (int)(r * kRgRows + or4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
            if (tok_ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (orow + j < p.N) {
                        const T o = from_f32<T>((p.bias ? p.bias[orow + j] : 0.f) + vv[j]);
                        if constexpr (kPush) {
                            for (uint32_t d = 0; d < p.n_dst; ++d) (reinterpret_cast<T*>(p.y_dst[d]) + yoff)[orow + j] = o;
                        } else {
                            (reinterpret_cast<T*>(p.y) + yoff)[orow + j] = o;
                        }
                    }
            }
        };
        float4 v_split[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};   // head / tail row group (when shared)
        for (uint32_t r = rg_a; r <= rg_b; ++r) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            bool any = false;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {                   // fixed order: warp 0's partial first
                if (hrg[w] == r) {
                    const float4 a4 = reinterpret_cast<const float4*>(smem + w * kWarpStride + kHeadOff)[64 * u + tid];
                    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w; any = true;
                }
                if (trg[w] == r) {
                    const float4 a4 = reinterpret_cast<const float4*>(smem + w * kWarpStride)[64 * u + tid];
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
    publish();
    trace_out();
}

template <typename T, int kOcc, bool kTrace = false, int kNT = 1, bool kLean = false, bool kPush = false>
__global__ void __launch_bounds__(dk::kThreads, kOcc) decode_mma_kernel(const dk::Params p) {
    using namespace dk;
    constexpr int kWarpBytes = dk::kWarpBytes + (kNT - 1) * kHeadBytes;
    constexpr int kTokP = kTok * kNT;
    constexpr uint32_t kOne2 = dk_one2<T>();
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

    // the tile starts (and, after every block, is again) +1.0 everywhere
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_v4(tile_s + lane * 16u + (uint32_t)i * 512u, kOne2, kOne2, kOne2, kOne2);

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

    // this warp's ring and its mbarriers
    const uint32_t ring_s = tile_s + kTileBytes + kNT * kHeadBytes, bar_s = ring_s + kStages * kStageBytes;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kStages; ++st) mbar_init(bar_s + 8u * st, 1);
        fence_barrier_init();
    }
    // entry offsets of the run: eptr lives in registers, one per lane (refilled every 28 blocks)
    uint32_t epr = 0;
    if (w_lo < w_hi && w_lo + lane <= w_hi) epr = __ldg(p.eptr + w_lo + lane);
    __syncwarp();
    // producer side (lane 0): one block = its 256 B of sign words + its first min(n4, kEntCap) entry units
    auto issue = [&](uint32_t blk, uint32_t st, uint32_t eb, uint32_t n4) {
        const uint32_t n = min(n4, (uint32_t)kEntCap), dst = ring_s + st * kStageBytes, bar = bar_s + 8u * st;
        mbar_arrive_expect_tx(bar, 256u + n * 16u);
        dk_bulk_g2s(dst, p.fsign + (size_t)blk * kRgRows, 256u, bar);
        if (n) dk_bulk_g2s(dst + 256u, p.ent + eb, n * 16u, bar);
    };
#pragma unroll
    for (int st = 0; st < kStages; ++st) {              // fill the ring: kStages blocks of DRAM latency in flight at once
        const uint32_t eb = __shfl_sync(0xffffffffu, epr, st), n4 = __shfl_sync(0xffffffffu, epr, st + 1) - eb;
        if (lane == 0 && w_lo + st < w_hi) issue(w_lo + st, st, eb, n4);
    }
    uint32_t rg = 0, kb = 0;
    if (w_lo < w_hi) { rg = w_lo / TC; kb = w_lo - rg * TC; }
    const uint32_t rg_first = rg;
    const bool grouped = !kLean && p.groups > 1;
    const bool has_mid = p.has_mid != 0u;
    uint32_t cur_g = grouped ? kb / p.tiles_per_group : 0u;
    const uint32_t g4 = lane >> 2, t4 = lane & 3u;
    // the {lo, hi} of this lane's four accumulator rows g4 + 8j of the current (row group, group): loaded when the segment
    // starts, first used when it ends (fold)
    float2 lv[4];
    auto load_levels = [&](uint32_t rgi, uint32_t grp) {
#pragma unroll
        for (int j = 0; j < 4; ++j) lv[j] = __ldg(p.affine + (size_t)(rgi * kRgRows + g4 + 8u * j) * p.groups + grp);
    };
#pragma unroll
    for (int j = 0; j < 4; ++j) lv[j] = make_float2(0.f, 0.f);
    if (w_lo < w_hi) load_levels(rg, cur_g);

    if (lane == 0) { s_hrg[wid] = kNone; s_trg[wid] = kNone; }
    if (tid == 0) dk_meta(s_meta, c_lo, c_hi, TC, q, rem);

    uint32_t ci = 0;                               // index of the current block in the eptr register chunk
    uint32_t cs = 0, cph = 0;                      // ring stage of the current block and the parity its mbarrier completes with

    // ---- activation loads: lane -> (token = lane>>2, 16-column segment = lane&3) of the 8 x 64 block ----------
    const uint32_t xtok = lane >> 2, xseg = lane & 3u;
    const bool x_fast = kLean || (((p.ldx & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15u) == 0) && ((p.K & 63) == 0));
    const bool x_fast256 = kLean || (x_fast && ((p.ldx & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 31u) == 0));   // one 32 B load
    const bool x_tok_ok = (m0 + (int)xtok) < p.M;
    // byte offset of (my token row, my 16-column segment, k-block 0); 32-bit (checked by the launcher) and opaque to the
    // compiler so it stays in a register instead of being recomputed every block
    uint32_t xoff_row = (uint32_t)(((int64_t)(m0 + (x_tok_ok ? (int)xtok : 0)) * p.ldx + 16 * xseg) * 2);
    asm volatile("" : "+r"(xoff_row));
    uint32_t xoff = xoff_row + kb * (kTileCols * 2u);          // loop-carried: offset of the NEXT block to load
    uint32_t xkb = kb;                                         // and its k-block
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
    struct XF { uint4 a, b, a2, b2; };                   // the B fragments of one block (second token group: a2, b2)
    auto load_x_next = [&](XF& X) {                      // load the next block in line, then advance the line
        load_x_at(xkb, xoff, x_tok_ok, X.a, X.b);
        if constexpr (kNT == 2) load_x_at(xkb, xoff2, x_tok_ok2, X.a2, X.b2);
        ++xkb;
        const bool wrap = xkb == TC;
        xkb = wrap ? 0u : xkb;
        xoff = wrap ? xoff_row : xoff + kTileCols * 2u;
        if constexpr (kNT == 2) xoff2 = wrap ? xoff_row2 : xoff2 + kTileCols * 2u;
    };

    // ldmatrix row addresses of this lane (A fragments of the correction tile), one per k16 step
    const uint32_t lm_row = (lane & 7u) + ((lane >> 3) & 1u) * 8u;
    const uint32_t lm_base0 = tile_s + lm_row * 128u + (((lane >> 4) ^ (lm_row & 7u)) << 4);
    uint32_t lm_q[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
        lm_q[qi] = lm_base0 ^ ((uint32_t)qi << 5);
        asm volatile("" : "+r"(lm_q[qi]));                                        // keep in a register (no rematerialisation)
    }

    // accumulators, all in the mma C layout (h = rows 0-15 / 16-31; i: row g4 + 8*(i>>1), token 2*t4 + (i&1)):
    //   acc_d  sum of t_ij x (+-1 plane and tau) of the current (row group, group) segment
    //   acc_x  sum of x over the segment's columns (identical in every row)
    //   acc    the folded result: mid * acc_x + half * acc_d of the finished segments
    float acc[2][4], acc_d[2][4], acc_x[4];
    float acc2[2][4], acc_d2[2][4], acc_x2[4];     // second token group (kNT == 2)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[0][i] = acc[1][i] = acc_d[0][i] = acc_d[1][i] = acc_x[i] = 0.f;
        acc2[0][i] = acc2[1][i] = acc_d2[0][i] = acc_d2[1][i] = acc_x2[i] = 0.f;
    }
    auto fold = [&]() {                             // apply the levels of the finished (row group, group) segment
        float mid[4], half[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { mid[j] = 0.5f * (lv[j].x + lv[j].y); half[j] = 0.5f * (lv[j].y - lv[j].x); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 2 * h + (i >> 1);
                acc[h][i] = fmaf(half[j], acc_d[h][i], acc[h][i]);
                if (has_mid) acc[h][i] = fmaf(mid[j], acc_x[i & 1], acc[h][i]);
                acc_d[h][i] = 0.f;
                if constexpr (kNT == 2) {
                    acc2[h][i] = fmaf(half[j], acc_d2[h][i], acc2[h][i]);
                    if (has_mid) acc2[h][i] = fmaf(mid[j], acc_x2[i & 1], acc2[h][i]);
                    acc_d2[h][i] = 0.f;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_x[i] = acc_x2[i] = 0.f;
    };

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
    if constexpr (kPush) {                              // the previous linear's slices from every rank have landed?
        if (p.wait_prev) {
            if (tid < p.n_dst) {
                const uint32_t want = p.sync_ctr[1];    // epoch our own previous pushed kernel published
                while ((int32_t)(ld_acquire_sys(p.flag_local + tid) - want) < 0) { }
            }
            __syncthreads();
        }
    }
    XF X0, X1;
    X0.a = X0.b = X0.a2 = X0.b2 = X1.a = X1.b = X1.a2 = X1.b2 = make_uint4(0, 0, 0, 0);
    if (w_lo < w_hi) load_x_next(X0);
    __syncwarp();                                       // the initialised tile is visible to the whole warp

    // The packed stream arrives through the ring (three blocks ahead, no registers); the activations are loaded one block
    // ahead into a second register set.
    const uint16_t one16 = (uint16_t)kOne2;
    auto do_block = [&](const uint32_t blk, const XF& X, XF& Xn) {
        const bool more = blk + 1 < w_hi;
        if (grouped) {
            const uint32_t g = kb / p.tiles_per_group;
            if (g != cur_g) {                           // a new group of this row group: fold the finished one, fetch the new levels
                fold();
                cur_g = g;
                load_levels(rg, g);
            }
        }
        if (more) load_x_next(Xn);                      // the next block's activations: a whole block of cover
        if (ci + (uint32_t)kStages + 1u > 31u) {        // rare: refill the eptr registers (runs longer than 28 blocks)
            if (blk + lane <= w_hi) epr = __ldg(p.eptr + blk + lane);
            ci = 0;
        }
        const uint32_t eb = __shfl_sync(0xffffffffu, epr, ci), n4 = __shfl_sync(0xffffffffu, epr, ci + 1u) - eb;
        const uint32_t n1 = min(n4, (uint32_t)kEntCap), h1 = (n1 + 1u) >> 1;
        const uint32_t stage = ring_s + cs * kStageBytes;

        // this block's record has landed: sign words and salient entries out of the ring, entries into the tile
        mbar_wait(bar_s + 8u * cs, cph);
        const uint2 sg = dk_lds64(stage + lane * 8u);
        uint4 ea = make_uint4(0, 0, 0, 0), ec = make_uint4(0, 0, 0, 0);
        if (lane < h1) { ea = dk_lds128(stage + 256u + lane * 16u); dk_patch4(tile_s, ea); }
        if (lane + h1 < n1) { ec = dk_lds128(stage + 256u + (lane + h1) * 16u); dk_patch4(tile_s, ec); }
        for (uint32_t i = (uint32_t)kEntCap + lane; i < n4; i += 32u) dk_patch4(tile_s, __ldg(p.ent + (eb + i)));   // rare: > 256 salient in a block
        __syncwarp();                                   // the patched tile is complete; every lane is done reading the stage
        {                                               // refill the stage with the block kStages ahead
            const uint32_t eb2 = __shfl_sync(0xffffffffu, epr, ci + (uint32_t)kStages);
            const uint32_t n42 = __shfl_sync(0xffffffffu, epr, ci + (uint32_t)kStages + 1u) - eb2;
            if (lane == 0 && blk + (uint32_t)kStages < w_hi) issue(blk + (uint32_t)kStages, cs, eb2, n42);
        }

        // the block on the tensor cores: A fragments = tile words with the sign bits XORed in, B fragments = the activation registers
        const uint32_t xw[8] = {X.a.x, X.a.y, X.a.z, X.a.w, X.b.x, X.b.y, X.b.z, X.b.w};
        const uint32_t xw2[8] = {X.a2.x, X.a2.y, X.a2.z, X.a2.w, X.b2.x, X.b2.y, X.b2.z, X.b2.w};
#pragma unroll
        for (int qi = 0; qi < 4; ++qi) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t wd = h ? sg.y : sg.x;
                uint32_t t0, t1, t2, t3;
                dk_ldsm4(lm_q[qi] + (uint32_t)h * 2048u, t0, t1, t2, t3);
                const uint32_t a0 = ((wd << (4 * qi + 0)) & 0x80008000u) ^ t0;
                const uint32_t a1 = ((wd << (4 * qi + 1)) & 0x80008000u) ^ t1;
                const uint32_t a2 = ((wd << (4 * qi + 2)) & 0x80008000u) ^ t2;
                const uint32_t a3 = ((wd << (4 * qi + 3)) & 0x80008000u) ^ t3;
                dk_mma<T>(acc_d[h], a0, a1, a2, a3, xw[2 * qi], xw[2 * qi + 1]);
                if constexpr (kNT == 2) dk_mma<T>(acc_d2[h], a0, a1, a2, a3, xw2[2 * qi], xw2[2 * qi + 1]);
            }
            if (has_mid) {                              // sum of x over the block's columns: all-ones A fragment
                dk_mma<T>(acc_x, kOne2, kOne2, kOne2, kOne2, xw[2 * qi], xw[2 * qi + 1]);
                if constexpr (kNT == 2) dk_mma<T>(acc_x2, kOne2, kOne2, kOne2, kOne2, xw2[2 * qi], xw2[2 * qi + 1]);
            }
        }

        __syncwarp();                                   // every lane's ldmatrix reads are done: reset the entries to +1.0
        if (lane < h1) dk_unpatch4(tile_s, ea, one16);
        if (lane + h1 < n1) dk_unpatch4(tile_s, ec, one16);
        for (uint32_t i = (uint32_t)kEntCap + lane; i < n4; i += 32u) dk_unpatch4(tile_s, __ldg(p.ent + (eb + i)), one16);
        ++ci;
        ++cs;
        if (cs == (uint32_t)kStages) { cs = 0; cph ^= 1u; }
        __syncwarp();                                   // the resets land before the next block's patch stores (other lanes, same slots)

        ++kb;
        const bool rg_end = kb == TC;

        // ---- end of this warp's part of the row group? -----------------------------------------------------
        if (rg_end || !more) {
            fold();
            const bool whole = rg_end && (w_lo <= rg * TC);      // this warp saw every k-block of the row group
            if (whole) {                                         // finish it here: + bias, round, store
                store_frag(tail_red);                            // (the tile is at rest and not read again before the next patch)
                __syncwarp();
                const int orow = (int)(rg * kRgRows + lane);
                if (orow < p.N) {
                    const float bv = p.bias ? p.bias[orow] : 0.f;
#pragma unroll
                    for (int m = 0; m < kTokP; ++m)
                        if (m0 + m < p.M) {
                            const T v = from_f32<T>(bv + tail_red[m * kRgRows + lane]);
                            if constexpr (kPush) {
                                for (uint32_t d = 0; d < p.n_dst; ++d) reinterpret_cast<T*>(p.y_dst[d])[(int64_t)(m0 + m) * p.ldy + orow] = v;
                            } else {
                                reinterpret_cast<T*>(p.y)[(int64_t)(m0 + m) * p.ldy + orow] = v;
                            }
                        }
                }
                __syncwarp();
                if (more) {                                      // the staging area is the tile: back to +1.0
#pragma unroll
                    for (int i = 0; i < kNT * 2; ++i) sts_v4(tile_s + lane * 16u + (uint32_t)i * 512u, kOne2, kOne2, kOne2, kOne2);
                    __syncwarp();
                }
            } else if (rg == rg_first) {
                store_frag(head_red);                            // head partial: its own buffer, the warp may go on
                if (lane == 0) s_hrg[wid] = rg;
            } else {
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
                    load_levels(rg, 0u);
                }
            }
        }
    };
    for (uint32_t blk = w_lo; blk < w_hi; blk += 2u) {   // unrolled by two: the activation register sets alternate without moves
        do_block(blk, X0, X1);
        if (blk + 1u < w_hi) do_block(blk + 1u, X1, X0);
    }

    dk_finish<T, kNT, kPush, kTrace, kWarpBytes, kTileBytes>(p, smem, s_hrg, s_trg, s_meta, c_lo, c_hi, m0, tr, w_hi - w_lo);
}

// ---- the pair kernel: the common case (one group per row, K a multiple of 128, M <= 8, aligned activations) -----------
//
// Same layout, same arithmetic, same stream-K partition and end of kernel as decode_mma_kernel, but the unit of work is a
// PAIR of k-adjacent blocks (32 rows x 128 columns).  A third of the block kernel's instructions are per-block bookkeeping
// (ring, barriers, offsets, activation prefetch, branches for the rare cases) and its launches end on stragglers; the loop
// itself is bound by the shared-memory pipe (ldmatrix of the tile + the 16-bit patch stores: profiles/
// r02_decode_pair_ablation.md), which this kernel does not change.  So here
//   * one ring stage, one mbarrier wait, one pair of bulk copies, one activation prefetch and three warp barriers serve two
//     blocks; the warp's tile is 8 KB (block A | block B), its entries patched in one go;
//   * the sixteen MMAs of a pair run on four independent accumulator chains (block x row half) instead of two;
//   * everything rare (groups, ragged K, unaligned activations, 9..16 tokens) stays in the general kernel, so the loop
//     has no branches for it.
namespace dk2 {
constexpr int kWarps = dk::kWarps, kThreads = dk::kThreads, kTok = dk::kTok;
constexpr int kTileBytes = 2 * dk::kTileBytes;            // 8192: block A of the pair at 0, block B at 4096
constexpr int kHeadBytes = dk::kHeadBytes;
constexpr int kStages = 2;                                // pairs in flight per warp (= 4 blocks)
// 16-byte entry units of a pair held by a ring stage (the rest comes straight from global memory, a microsecond-class stall):
// at 10 % salient weights a pair has 103 +- 5 units, so 124 covers all but one pair in 10^5 where 112 lost 3.6 % of them; it is
// also all that two CTAs per SM leave: 2 x (8 x 14336 + 1 KB reserved + static) = 231616 of the SM's 233472 bytes
constexpr int kEntCap = 124;
constexpr int kStageBytes = 512 + kEntCap * 16;           // 2496
constexpr int kWarpBytes = kTileBytes + kHeadBytes + kStages * kStageBytes + 128;   // 14336: + mbarriers, padded to 128
constexpr int kUnitsPerLane = (kEntCap + 31) / 32;        // 4 entry units per lane and pair
static_assert(kWarpBytes % 128 == 0, "per-warp regions must keep the tile 128-byte aligned");
}  // namespace dk2

template <typename T, bool kTrace = false, bool kPush = false>
__global__ void __launch_bounds__(dk2::kThreads, 2) decode_pair_kernel(const dk::Params p) {
    using namespace dk2;
    using dk::kOut;
    constexpr uint32_t kOne2 = dk_one2<T>();
    constexpr uint32_t kNone = 0xffffffffu;
    unsigned long long tr[8];
    if (kTrace) { tr[0] = dk_now(); }
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_hrg[kWarps], s_trg[kWarps];
    __shared__ uint32_t s_meta[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint8_t* wsm = smem + wid * kWarpBytes;
    const uint32_t tile_s = smem_u32(wsm);
    float* head_red = reinterpret_cast<float*>(wsm + kTileBytes);
    float* tail_red = reinterpret_cast<float*>(wsm);      // aliases the tile: written only when the tile is at rest

    // ---- work partition, in pairs: warp gw owns pairs [gw*q + min(gw,rem), ...) --------------------------------------
    const uint32_t TCp = p.tiles_c >> 1, q = p.q, rem = p.rem;
    auto wstart = [&](uint32_t gw) { return gw * q + min(gw, rem); };
    const uint32_t gw0 = blockIdx.x * kWarps;
    const uint32_t c_lo = wstart(gw0), c_hi = wstart(gw0 + kWarps);
    const uint32_t w_lo = wstart(gw0 + wid), w_hi = wstart(gw0 + wid + 1u);
    const int m0 = blockIdx.y * kTok;

    // entry offsets first (the bulk copies below depend on them: the one DRAM round trip before any weight byte moves):
    // lane l holds eptr[2 * chunk + l], i.e. the three offsets of pair chunk + j sit in lanes 2j, 2j+1, 2j+2; a chunk serves
    // 13 pairs (the refill two pairs ahead reads lanes up to 2j + 6)
    uint32_t epr = 0;
    if (w_lo < w_hi && 2u * w_lo + lane <= 2u * w_hi) epr = __ldg(p.eptr + 2u * w_lo + lane);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

#pragma unroll
    for (int i = 0; i < 16; ++i) sts_v4(tile_s + lane * 16u + (uint32_t)i * 512u, kOne2, kOne2, kOne2, kOne2);

    const uint32_t ring_s = tile_s + kTileBytes + kHeadBytes, bar_s = ring_s + kStages * kStageBytes;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kStages; ++st) mbar_init(bar_s + 8u * st, 1);
        fence_barrier_init();
    }
    __syncwarp();
    auto issue = [&](uint32_t pr, uint32_t st, uint32_t eb, uint32_t n) {       // lane 0: one pair = 512 B of sign words + its entries
        const uint32_t n1 = min(n, (uint32_t)kEntCap), dst = ring_s + st * kStageBytes, bar = bar_s + 8u * st;
        mbar_arrive_expect_tx(bar, 512u + n1 * 16u);
        dk_bulk_g2s(dst, p.fsign + (size_t)pr * (2u * kRgRows), 512u, bar);
        if (n1) dk_bulk_g2s(dst + 512u, p.ent + eb, n1 * 16u, bar);
    };
#pragma unroll
    for (int st = 0; st < kStages; ++st) {
        const uint32_t eb = __shfl_sync(0xffffffffu, epr, 2 * st), n = __shfl_sync(0xffffffffu, epr, 2 * st + 2) - eb;
        if (w_lo + st < w_hi) { if (dk_elect()) issue(w_lo + st, st, eb, n); }
    }
    uint32_t rg = 0, kp = 0;
    if (w_lo < w_hi) { rg = w_lo / TCp; kp = w_lo - rg * TCp; }
    const uint32_t rg_first = rg;
    const bool has_mid = p.has_mid != 0u;
    const uint32_t g4 = lane >> 2, t4 = lane & 3u;
    float2 lv[4];                                   // {lo, hi} of this lane's four accumulator rows of the current row group
    auto load_levels = [&](uint32_t rgi) {
#pragma unroll
        for (int j = 0; j < 4; ++j) lv[j] = __ldg(p.affine + (size_t)(rgi * kRgRows + g4 + 8u * j));
    };
#pragma unroll
    for (int j = 0; j < 4; ++j) lv[j] = make_float2(0.f, 0.f);
    if (w_lo < w_hi) load_levels(rg);

    if (lane == 0) { s_hrg[wid] = kNone; s_trg[wid] = kNone; }
    if (tid == 0) dk_meta(s_meta, c_lo, c_hi, TCp, q, rem);

    // activations: lane -> (token lane>>2, 16-column segment lane&3) of each of the pair's two 8 x 64 blocks
    const uint32_t xtok = lane >> 2, xseg = lane & 3u;
    const bool x_tok_ok = (m0 + (int)xtok) < p.M;
    uint32_t xoff_row = (uint32_t)(((int64_t)(m0 + (x_tok_ok ? (int)xtok : 0)) * p.ldx + 16 * xseg) * 2);
    asm volatile("" : "+r"(xoff_row));
    uint32_t xoff = xoff_row + kp * (2u * kTileCols * 2u);      // loop-carried: byte offset of the next pair to load
    uint32_t xkp = kp;
    const uint8_t* xbytes = reinterpret_cast<const uint8_t*>(p.x);
    struct XF { uint32_t a[8], b[8]; };                   // B fragments of block A / block B
    auto load_x_next = [&](XF& X) {
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(X.a[0]), "=r"(X.a[1]), "=r"(X.a[2]), "=r"(X.a[3]), "=r"(X.a[4]), "=r"(X.a[5]), "=r"(X.a[6]), "=r"(X.a[7])
                     : "l"(xbytes + xoff));
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(X.b[0]), "=r"(X.b[1]), "=r"(X.b[2]), "=r"(X.b[3]), "=r"(X.b[4]), "=r"(X.b[5]), "=r"(X.b[6]), "=r"(X.b[7])
                     : "l"(xbytes + xoff + kTileCols * 2u));
        ++xkp;
        const bool wrap = xkp == TCp;
        xkp = wrap ? 0u : xkp;
        xoff = wrap ? xoff_row : xoff + 2u * kTileCols * 2u;
    };

    const uint32_t lm_row = (lane & 7u) + ((lane >> 3) & 1u) * 8u;
    const uint32_t lm_base0 = tile_s + lm_row * 128u + (((lane >> 4) ^ (lm_row & 7u)) << 4);
    uint32_t lm_q[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
        lm_q[qi] = lm_base0 ^ ((uint32_t)qi << 5);
        asm volatile("" : "+r"(lm_q[qi]));
    }

    // accumulators in the mma C layout: acc_d[block of the pair][row half][4], acc_x the sum of x, acc the folded result
    float acc[2][4], acc_d[2][2][4], acc_x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[0][i] = acc[1][i] = acc_x[i] = 0.f;
        acc_d[0][0][i] = acc_d[0][1][i] = acc_d[1][0][i] = acc_d[1][1][i] = 0.f;
    }
    auto fold = [&]() {
        float mid[4], half[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { mid[j] = 0.5f * (lv[j].x + lv[j].y); half[j] = 0.5f * (lv[j].y - lv[j].x); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 2 * h + (i >> 1);
                acc[h][i] = fmaf(half[j], acc_d[0][h][i] + acc_d[1][h][i], acc[h][i]);
                if (has_mid) acc[h][i] = fmaf(mid[j], acc_x[i & 1], acc[h][i]);
                acc_d[0][h][i] = acc_d[1][h][i] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_x[i] = 0.f;
    };
    auto store_frag = [&](float* dst) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t r = 16u * h + g4;
            dst[(2u * t4) * kRgRows + r] = acc[h][0];
            dst[(2u * t4 + 1u) * kRgRows + r] = acc[h][1];
            dst[(2u * t4) * kRgRows + r + 8u] = acc[h][2];
            dst[(2u * t4 + 1u) * kRgRows + r + 8u] = acc[h][3];
        }
    };

    if (kTrace) tr[1] = dk_now();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (kTrace) tr[2] = dk_now();
    if constexpr (kPush) {
        if (p.wait_prev) {
            if (tid < p.n_dst) {
                const uint32_t want = p.sync_ctr[1];
                while ((int32_t)(ld_acquire_sys(p.flag_local + tid) - want) < 0) { }
            }
            __syncthreads();
        }
    }
    XF X0, X1;
#pragma unroll
    for (int i = 0; i < 8; ++i) X0.a[i] = X0.b[i] = X1.a[i] = X1.b[i] = 0u;
    if (w_lo < w_hi) load_x_next(X0);
    __syncwarp();

    uint32_t ci = 0;                               // pair index inside the eptr chunk
    uint32_t cs = 0, cph = 0;                      // ring stage of the current pair and its mbarrier parity
    const uint16_t one16 = (uint16_t)kOne2;
    auto do_pair = [&](const uint32_t pr, const XF& X, XF& Xn) {
        const bool more = pr + 1u < w_hi;
        if (more) load_x_next(Xn);
        if (ci > 12u) {                                 // rare: a run longer than 13 pairs refills the offsets
            if (2u * pr + lane <= 2u * w_hi) epr = __ldg(p.eptr + 2u * pr + lane);
            ci = 0;
        }
        const uint32_t eb = __shfl_sync(0xffffffffu, epr, 2u * ci);
        const uint32_t nA = __shfl_sync(0xffffffffu, epr, 2u * ci + 1u) - eb, n = __shfl_sync(0xffffffffu, epr, 2u * ci + 2u) - eb;
        const uint32_t stage = ring_s + cs * kStageBytes;

        mbar_wait(bar_s + 8u * cs, cph);
        const uint2 sgA = dk_lds64(stage + lane * 8u), sgB = dk_lds64(stage + 256u + lane * 8u);
        const uint32_t n1 = min(n, (uint32_t)kEntCap);
        uint32_t ad[kUnitsPerLane][4];                  // where this lane's entries went: the reset needs no arithmetic
#pragma unroll
        for (int j = 0; j < kUnitsPerLane; ++j) {       // entry unit lane + 32 j of the pair: block A's units first, then block B's
            const uint32_t i = lane + 32u * j;
            const uint32_t tb = tile_s + (i >= nA ? (uint32_t)dk::kTileBytes : 0u);
            ad[j][0] = ad[j][1] = ad[j][2] = ad[j][3] = tb;
            if (i < n1) {
                const uint4 e = dk_lds128(stage + 512u + i * 16u);
                ad[j][0] = tb + (e.x >> 20); ad[j][1] = tb + (e.y >> 20); ad[j][2] = tb + (e.z >> 20); ad[j][3] = tb + (e.w >> 20);
                sts_u16(ad[j][0], (uint16_t)e.x); sts_u16(ad[j][1], (uint16_t)e.y);
                sts_u16(ad[j][2], (uint16_t)e.z); sts_u16(ad[j][3], (uint16_t)e.w);
            }
        }
#pragma unroll 1
        for (uint32_t i = (uint32_t)kEntCap + lane; i < n; i += 32u)            // rare: more than 496 salient weights in the pair
            dk_patch4(tile_s + (i >= nA ? (uint32_t)dk::kTileBytes : 0u), __ldg(p.ent + (eb + i)));
        __syncwarp();                                   // the patched tile is complete; every lane is done reading the stage
        {
            const uint32_t eb2 = __shfl_sync(0xffffffffu, epr, 2u * ci + 2u * kStages);
            const uint32_t n2 = __shfl_sync(0xffffffffu, epr, 2u * ci + 2u * kStages + 2u) - eb2;
            if (pr + (uint32_t)kStages < w_hi) { if (dk_elect()) issue(pr + (uint32_t)kStages, cs, eb2, n2); }
        }

#pragma unroll
        for (int qi = 0; qi < 4; ++qi) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const uint32_t* xw = b ? X.b : X.a;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t wd = b ? (h ? sgB.y : sgB.x) : (h ? sgA.y : sgA.x);
                    uint32_t t0, t1, t2, t3;
                    dk_ldsm4(lm_q[qi] + (uint32_t)b * 4096u + (uint32_t)h * 2048u, t0, t1, t2, t3);
                    const uint32_t a0 = ((wd << (4 * qi + 0)) & 0x80008000u) ^ t0;
                    const uint32_t a1 = ((wd << (4 * qi + 1)) & 0x80008000u) ^ t1;
                    const uint32_t a2 = ((wd << (4 * qi + 2)) & 0x80008000u) ^ t2;
                    const uint32_t a3 = ((wd << (4 * qi + 3)) & 0x80008000u) ^ t3;
                    dk_mma<T>(acc_d[b][h], a0, a1, a2, a3, xw[2 * qi], xw[2 * qi + 1]);
                }
                if (has_mid) dk_mma<T>(acc_x, kOne2, kOne2, kOne2, kOne2, xw[2 * qi], xw[2 * qi + 1]);
            }
        }

        __syncwarp();                                   // every lane's ldmatrix reads are done: reset the entries to +1.0
#pragma unroll
        for (int j = 0; j < kUnitsPerLane; ++j)
            if (lane + 32u * j < n1) {
                sts_u16(ad[j][0], one16); sts_u16(ad[j][1], one16); sts_u16(ad[j][2], one16); sts_u16(ad[j][3], one16);
            }
#pragma unroll 1
        for (uint32_t i = (uint32_t)kEntCap + lane; i < n; i += 32u)
            dk_unpatch4(tile_s + (i >= nA ? (uint32_t)dk::kTileBytes : 0u), __ldg(p.ent + (eb + i)), one16);
        ++ci;
        cs ^= 1u;
        cph ^= (cs == 0u) ? 1u : 0u;
        __syncwarp();                                   // the resets land before the next pair's patch stores

        ++kp;
        const bool rg_end = kp == TCp;
        if (rg_end || !more) {                          // end of this warp's part of the row group
            fold();
            const bool whole = rg_end && (w_lo <= rg * TCp);
            if (whole) {                                // this warp saw the whole row group: + bias, round, store
                store_frag(tail_red);
                __syncwarp();
                const int orow = (int)(rg * kRgRows + lane);
                if (orow < p.N) {
                    const float bv = p.bias ? p.bias[orow] : 0.f;
#pragma unroll
                    for (int m = 0; m < kTok; ++m)
                        if (m0 + m < p.M) {
                            const T v = from_f32<T>(bv + tail_red[m * kRgRows + lane]);
                            if constexpr (kPush) {
                                for (uint32_t d = 0; d < p.n_dst; ++d) reinterpret_cast<T*>(p.y_dst[d])[(int64_t)(m0 + m) * p.ldy + orow] = v;
                            } else {
                                reinterpret_cast<T*>(p.y)[(int64_t)(m0 + m) * p.ldy + orow] = v;
                            }
                        }
                }
                __syncwarp();
                if (more) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) sts_v4(tile_s + lane * 16u + (uint32_t)i * 512u, kOne2, kOne2, kOne2, kOne2);
                    __syncwarp();
                }
            } else if (rg == rg_first) {
                store_frag(head_red);
                if (lane == 0) s_hrg[wid] = rg;
            } else {
                store_frag(tail_red);
                if (lane == 0) s_trg[wid] = rg;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = 0.f;
            if (rg_end) {
                kp = 0;
                ++rg;
                if (more) load_levels(rg);
            }
        }
    };
    for (uint32_t pr = w_lo; pr < w_hi; pr += 2u) {
        do_pair(pr, X0, X1);
        if (pr + 1u < w_hi) do_pair(pr + 1u, X1, X0);
    }
    dk_finish<T, 1, kPush, kTrace, kWarpBytes, kTileBytes>(p, smem, s_hrg, s_trg, s_meta, c_lo, c_hi, m0, tr, w_hi - w_lo);
}

// ---- host side -------------------------------------------------------------------------------------------------
static int dk_occupancy();
static int dk_ctas_per_sm() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PBL_DK_CTAS");
        v = (e && *e) ? atoi(e) : dk_occupancy();     // grid = SMs x resident CTAs per SM
        if (v < 1) v = 1;
        if (v > 16) v = 16;
    }
    return v;
}

static int dk_occupancy() {      // which register budget / launch-bounds variant to launch
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PBL_DK_OCC");
        v = (e && *e) ? atoi(e) : 2;                  // 2 -> up to 128 registers, 16 warps per SM (measured best); 3 -> 80 registers, 24 warps
        if (v != 3) v = 2;
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
    if (!L.fsign || !L.eptr || !L.ent) return false;
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) return false;
    if (M <= 0 || M > 65535LL * 16) return false;                                      // grid.y = token passes
    if (ldx <= 0 || (uint64_t)M * (uint64_t)ldx * 2u >= (1ull << 31)) return false;   // 32-bit activation offsets
    return true;
}

// the pair kernel's layer-side conditions (the launch adds the activation alignment): one group per row, K a multiple of
// 128 (pairs of 64-column blocks never straddle a row group), one pass of at most 8 tokens
static bool dk_pair_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PBL_DK_PAIR"); v = (e && *e) ? atoi(e) : 1; }
    return v != 0;
}
static bool pair_layer_ok(const Layer& L, int64_t M) {
    return dk_pair_enabled() && L.groups == 1 && (L.K & 127) == 0 && (L.tiles_c & 1) == 0 && M <= dk::kTok;
}
static DecodeGeom pair_geom(const Layer& L) {        // the partition in pairs of blocks: q, rem, slots in pair units
    DecodeGeom g;
    g.rgs = (uint32_t)(L.tiles_r * kRgPerTile);
    g.nblocks = g.rgs * (uint32_t)L.tiles_c;
    const uint32_t npairs = g.nblocks / 2u, tcp = (uint32_t)L.tiles_c / 2u;
    const uint32_t want = (uint32_t)(dk_num_sms() * (dk_ctas_per_sm() < 2 ? dk_ctas_per_sm() : 2));   // shared memory: 2 CTAs per SM
    const uint32_t cap = npairs / dk::kWarps > 0 ? npairs / dk::kWarps : 1u;
    g.grid = cap < want ? cap : want;
    g.q = npairs / (g.grid * dk::kWarps);
    g.rem = npairs % (g.grid * dk::kWarps);
    g.slots = g.q ? (tcp + 8u * g.q - 1u) / (8u * g.q) + 1u : 2u;
    g.nt = 1u;
    g.passes = 1u;
    g.ws_bytes = (size_t)g.rgs * g.slots * dk::kOut * 8u;
    return g;
}

size_t decode_workspace_bytes(const Layer& L, int64_t M) {
    if (!decode_supported(L, L.K, M)) return 0;
    size_t b = decode_geom(L, M).ws_bytes;              // whichever kernel the launch picks (it depends on the activations' alignment)
    if (pair_layer_ok(L, M)) { const size_t pb = pair_geom(L).ws_bytes; b = pb > b ? pb : b; }
    return b;
}

static void dk_fill_params(dk::Params& p, const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws,
                           const DecodeGeom& g, const pbl_peer_push* push) {
    p.fsign = L.fsign; p.eptr = L.eptr; p.ent = reinterpret_cast<const uint4*>(L.ent);
    p.has_mid = (L.flags & PBL_LAYER_HAS_MID) ? 1u : 0u;
    p.affine = L.affine; p.bias = L.bias; p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy;
    p.ws_part = ws;
    p.M = (int)M; p.N = (int)L.N; p.K = (int)L.K;
    p.tiles_c = (uint32_t)L.tiles_c; p.groups = (uint32_t)L.groups; p.tiles_per_group = (uint32_t)L.tiles_per_group;
    p.nblocks = g.nblocks; p.rgs = g.rgs; p.slots = g.slots; p.q = g.q; p.rem = g.rem;
    p.n_dst = 0; p.rank = 0; p.wait_prev = 0; p.flag_local = nullptr; p.sync_ctr = nullptr; p.trace = nullptr;
    for (int d = 0; d < PBL_MAX_PEERS; ++d) { p.y_dst[d] = nullptr; p.flag_dst[d] = nullptr; }
    if (push) {
        p.n_dst = (uint32_t)push->n_ranks; p.rank = (uint32_t)push->rank; p.wait_prev = push->wait_prev ? 1u : 0u;
        p.flag_local = push->flags[push->rank]; p.sync_ctr = push->sync_ctr;
        for (int d = 0; d < push->n_ranks; ++d) { p.y_dst[d] = push->y[d]; p.flag_dst[d] = push->flags[d]; }
    }
}

template <typename K>
static cudaError_t dk_launch(K kernel, dim3 grid, int smem, cudaStream_t s, const dk::Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
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
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <typename K>
static int dk_set_smem(K kernel, int smem, int (&done)[64]) {       // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (done[dev] >= smem) return PBL_OK;
    const int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute(decode smem)");
    if (!rc) done[dev] = smem;
    return rc;
}

template <typename T, bool kPush>
static int launch_decode_pair_t(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws, cudaStream_t s,
                                const pbl_peer_push* push = nullptr) {
    const DecodeGeom g = pair_geom(L);
    const int smem = dk2::kWarps * dk2::kWarpBytes;
    dk::Params p;
    dk_fill_params(p, L, x, ldx, y, ldy, M, ws, g, kPush ? push : nullptr);
    cudaError_t le;
    if (g_trace && g_trace_next < g_trace_launches && !kPush) {
        static int done_t[64] = {};
        if (int rc = dk_set_smem(decode_pair_kernel<T, true, false>, smem, done_t)) return rc;
        p.trace = g_trace + (g_trace_next++) * kTraceStride;
        le = dk_launch(decode_pair_kernel<T, true, false>, dim3(g.grid, 1), smem, s, p);
    } else {
        static int done[64] = {};
        if (int rc = dk_set_smem(decode_pair_kernel<T, false, kPush>, smem, done)) return rc;
        le = dk_launch(decode_pair_kernel<T, false, kPush>, dim3(g.grid, 1), smem, s, p);
    }
    count_launch();
    return check_cuda(le, "decode (pair) launch");
}

template <typename T, int kOcc, int kNT, bool kLean, bool kPush = false>
static int launch_decode_t(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws, cudaStream_t s,
                           const pbl_peer_push* push = nullptr) {
    const DecodeGeom g = decode_geom(L, M);
    const int smem = dk::kWarps * (dk::kWarpBytes + (kNT - 1) * dk::kHeadBytes);
    dk::Params p;
    dk_fill_params(p, L, x, ldx, y, ldy, M, ws, g, kPush ? push : nullptr);
    cudaError_t le;
    if (g_trace && g_trace_next < g_trace_launches && kOcc == 2 && kNT == 1 && !kLean && !kPush) {
        static int done_t[64] = {};
        if (int rc = dk_set_smem(decode_mma_kernel<T, 2, true>, smem, done_t)) return rc;
        p.trace = g_trace + (g_trace_next++) * kTraceStride;
        le = dk_launch(decode_mma_kernel<T, 2, true>, dim3(g.grid, g.passes), smem, s, p);
    } else {
        static int done[64] = {};
        if (int rc = dk_set_smem(decode_mma_kernel<T, kOcc, false, kNT, kLean, kPush>, smem, done)) return rc;
        le = dk_launch(decode_mma_kernel<T, kOcc, false, kNT, kLean, kPush>, dim3(g.grid, g.passes), smem, s, p);
    }
    count_launch();
    return check_cuda(le, "decode launch");
}

// ws == nullptr: take a transient workspace from the stream-ordered pool and zero it (slower: the memset
// sits between consecutive decode kernels); callers on the hot path pass a persistent zero-initialised workspace.
__global__ void peer_wait_kernel(const uint32_t* __restrict__ flag_local, const uint32_t* __restrict__ sync_ctr, uint32_t n) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < n) {
        const uint32_t want = sync_ctr[1];
        while ((int32_t)(ld_acquire_sys(flag_local + threadIdx.x) - want) < 0) { }
    }
}

int launch_peer_wait(const pbl_peer_push& push, cudaStream_t s) {
    peer_wait_kernel<<<1, 32, 0, s>>>(push.flags[push.rank], push.sync_ctr, (uint32_t)push.n_ranks);
    count_launch();
    return check_cuda(cudaGetLastError(), "peer_wait launch");
}

// the lean instances: one group per row, 32-byte aligned activation rows, K a multiple of 64 ...
static bool dk_aligned(const Layer& L, const void* x, int64_t ldx) {
    return L.groups == 1 && (L.K & 63) == 0 && (ldx & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 31u) == 0;
}

// host-only (pbl_decode_variant): 2 = pair kernel, 1 = block kernel, 0 = the decode kernel does not take this call
int decode_variant(const Layer& L, const void* x, int64_t ldx, int64_t M) {
    if (!decode_supported(L, ldx, M)) return 0;
    return dk_aligned(L, x, ldx) && pair_layer_ok(L, M) ? 2 : 1;
}

int launch_decode(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws, size_t ws_bytes,
                  cudaStream_t s, const pbl_peer_push* push) {
    // ... and with K a multiple of 128 and at most 8 tokens, the pair kernel
    const bool aligned = dk_aligned(L, x, ldx);
    const bool pair = aligned && pair_layer_ok(L, M);
    const DecodeGeom g = pair ? pair_geom(L) : decode_geom(L, M);
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
    const bool lean = !g_trace && aligned;               // (tracing the block kernel uses its generic instance)
    const bool f16 = L.dtype == PBL_F16;
    int rc;
#define PBL_DK_LAUNCH(OCC, NT)                                                                                                 \
    rc = lean ? (f16 ? launch_decode_t<__half, OCC, NT, true>(L, x, ldx, y, ldy, M, ws, s)                                     \
                     : launch_decode_t<__nv_bfloat16, OCC, NT, true>(L, x, ldx, y, ldy, M, ws, s))                             \
              : (f16 ? launch_decode_t<__half, OCC, NT, false>(L, x, ldx, y, ldy, M, ws, s)                                    \
                     : launch_decode_t<__nv_bfloat16, OCC, NT, false>(L, x, ldx, y, ldy, M, ws, s))
#define PBL_DK_PUSH(NT)                                                                                                        \
    rc = lean ? (f16 ? launch_decode_t<__half, 2, NT, true, true>(L, x, ldx, y, ldy, M, ws, s, push)                           \
                     : launch_decode_t<__nv_bfloat16, 2, NT, true, true>(L, x, ldx, y, ldy, M, ws, s, push))                   \
              : (f16 ? launch_decode_t<__half, 2, NT, false, true>(L, x, ldx, y, ldy, M, ws, s, push)                          \
                     : launch_decode_t<__nv_bfloat16, 2, NT, false, true>(L, x, ldx, y, ldy, M, ws, s, push))
    if (pair) {
        rc = push ? (f16 ? launch_decode_pair_t<__half, true>(L, x, ldx, y, ldy, M, ws, s, push)
                         : launch_decode_pair_t<__nv_bfloat16, true>(L, x, ldx, y, ldy, M, ws, s, push))
                  : (f16 ? launch_decode_pair_t<__half, false>(L, x, ldx, y, ldy, M, ws, s)
                         : launch_decode_pair_t<__nv_bfloat16, false>(L, x, ldx, y, ldy, M, ws, s));
    }
    else if (push) {
        if (g.passes != 1) { set_error("pbl_linear_forward_push: one decode pass only (M <= 16)"); rc = PBL_ERR_UNSUPPORTED; }
        else if (g.nt == 2) { PBL_DK_PUSH(2); }
        else { PBL_DK_PUSH(1); }
    }
    else if (g.nt == 2) { PBL_DK_LAUNCH(2, 2); }         // 9..16 tokens in one pass
    else if (dk_occupancy() == 3) { PBL_DK_LAUNCH(3, 1); }
    else { PBL_DK_LAUNCH(2, 1); }
#undef PBL_DK_LAUNCH
#undef PBL_DK_PUSH
    if (own) {
        const int rf = check_cuda(cudaFreeAsync(own, s), "cudaFreeAsync(decode workspace)");
        if (!rc) rc = rf;
    }
    return rc;
}

}  // namespace pbl
