// XNOR-popcount forward for BiRealLinear (reference quant/quantizer.py:131-169), the one layer of
// the reference whose activations are binarized too:
//     y[m,i] = sum_j sign(x[m,j]) * w_sim[i,j],   w_sim = mean_row|W| * sign(W),   NO bias (:168)
// With the packed sign plane b (1 = hi) and x's sign planes xp = (x>0), xn = (x<0):
//     d1 = popc(b & xp) - popc(b & xn)            (sum of sign(x) over the row's `hi` positions)
//     d0 = popc(nb & xp) - popc(nb & xn)          (same over its `lo` positions, nb = ~b & ~salient)
//     y  = hi * d1 + lo * d0                      == alpha * (2*popc(xnor) - K) when lo = -hi, no zeros
// Integer counting is exact; the only rounding is the final multiply.  Requires a packed layer whose
// salient values are all exactly zero (sign(0) = 0 weights), which pack(alpha*sign(W)) guarantees.
// HBM-bound on the 0.25 B/weight plane stream: lane = weight row, warps split K, x bits are
// broadcast 16 B loads.
#include <cstdlib>

#include "pbllm_common.cuh"

namespace pbl {

constexpr int kBrWarps = 16;

// x [M][ldx] (any dtype) -> xb [M][tiles_c] uint4 {xp[0:32], xp[32:64], xn[0:32], xn[32:64]}, dx [M][tiles_c] int
template <typename T>
__global__ void __launch_bounds__(64) bireal_binarize_kernel(const T* __restrict__ x, int64_t ldx, int64_t K, int tiles_c,
                                                             uint4* __restrict__ xb, int* __restrict__ dx) {
    const int kb = blockIdx.x, m = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t c = (int64_t)kb * kTileCols + w * 32 + lane;
    float v = 0.f;
    if (c < K) v = to_f32(x[(int64_t)m * ldx + c]);
    const uint32_t p = __ballot_sync(0xffffffffu, v > 0.f), n = __ballot_sync(0xffffffffu, v < 0.f);
    __shared__ uint32_t sh[4];
    if (lane == 0) { sh[w] = p; sh[2 + w] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        xb[(int64_t)m * tiles_c + kb] = make_uint4(sh[0], sh[1], sh[2], sh[3]);
        dx[(int64_t)m * tiles_c + kb] = __popc(sh[0]) + __popc(sh[1]) - __popc(sh[2]) - __popc(sh[3]);
    }
}

template <int MT, bool kCompact>
__global__ void __launch_bounds__(kBrWarps * 32)
bireal_xnor_kernel(const uint4* __restrict__ planes, const uint2* __restrict__ sign_planes, const float2* __restrict__ affine, const uint4* __restrict__ xb,
                   const int* __restrict__ dx, float* __restrict__ y, int64_t ldy, int64_t M, int64_t N, int tiles_c, int groups,
                   int tiles_per_group) {
    __shared__ float red[kBrWarps][MT][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t rg = blockIdx.x, tr = rg / kRgPerTile;
    const int rgi = (int)(rg % kRgPerTile);
    const int64_t row = rg * kRgRows + lane;
    const int64_t m0 = (int64_t)blockIdx.y * MT;

    float acc[MT];
    int d1[MT], d0[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) { acc[m] = 0.f; d1[m] = d0[m] = 0; }
    int cur_g = -1;
    float2 a = make_float2(0.f, 0.f);

    auto load_planes = [&](int kb) {
        uint4 p = make_uint4(0u, 0u, 0u, 0u);
        if (kb < tiles_c) {
            if constexpr (kCompact) {   // sign words only (pure binary layer): half the HBM bytes
                const uint2 sp = __ldg(sign_planes + (tr * tiles_c + kb) * kTileRows + rgi * kRgRows + lane);
                p = make_uint4(sp.x, sp.y, 0u, 0u);
            } else {
                p = __ldg(planes + (tr * tiles_c + kb) * kTileRows + rgi * kRgRows + lane);
            }
        }
        return p;
    };
    uint4 pn0 = load_planes(wid), pn1 = load_planes(wid + kBrWarps);   // two k-blocks in flight per warp
    for (int kb = wid; kb < tiles_c; kb += kBrWarps) {
        const uint4 p = pn0;
        pn0 = pn1;
        pn1 = load_planes(kb + 2 * kBrWarps);
        const int g = kb / tiles_per_group;
        if (g != cur_g) {   // fold the finished group, fetch the next {lo,hi}
#pragma unroll
            for (int m = 0; m < MT; ++m) { acc[m] += a.y * (float)d1[m] + a.x * (float)d0[m]; d1[m] = d0[m] = 0; }
            a = __ldg(affine + row * groups + g);
            cur_g = g;
        }
        bool any_sal = false;
        if constexpr (!kCompact) any_sal = __any_sync(0xffffffffu, (p.z | p.w) != 0u);
        const uint32_t nb0 = ~(p.x | p.z), nb1 = ~(p.y | p.w);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m0 + m < M) {   // uniform
                const uint4 xv = __ldg(xb + (m0 + m) * tiles_c + kb);
                const int t1 = __popc(p.x & xv.x) + __popc(p.y & xv.y) - __popc(p.x & xv.z) - __popc(p.y & xv.w);
                d1[m] += t1;
                if (any_sal) d0[m] += __popc(nb0 & xv.x) + __popc(nb1 & xv.y) - __popc(nb0 & xv.z) - __popc(nb1 & xv.w);
                else d0[m] += __ldg(dx + (m0 + m) * tiles_c + kb) - t1;   // no zero weights in this slab: lo set = complement
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) red[wid][m][lane] = acc[m] + a.y * (float)d1[m] + a.x * (float)d0[m];
    __syncthreads();
    if (wid == 0 && row < N) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m0 + m < M) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kBrWarps; ++w) s += red[w][m][lane];
                y[(m0 + m) * ldy + row] = s;
            }
        }
    }
}


// ---- stream-K XNOR-popcount kernel ----------------------------------------------------------------------------------------
// Same work partition and cross-CTA reduction as decode_mma_kernel (csrc/pbllm_decode.cu): the layer's 32x64 blocks are
// dealt in contiguous, equal runs to the warps of a fixed grid; a warp keeps integer counts d1 (sum of sign(x) over the
// `hi` positions of its rows) and S (sum of sign(x) over the whole block; the `lo` positions are the complement), folds
// them with the row's {lo,hi} when its part of a row group ends, and partial row groups are reduced across the CTA's
// warps in shared memory and across CTAs through tagged 64-bit workspace slots summed in CTA order (deterministic).
// The activation bit planes of a warp's run are staged eight blocks at a time in its private shared memory.
namespace bsk {
constexpr int kWarps = 8, kThreads = 256, kTok = 8, kOut = 256;
constexpr int kChunk = 8;                                   // blocks of activation bits staged at once
constexpr int kXbBytes = kChunk * kTok * 16;                // 1024: uint4 {xp0, xp1, xn0, xn1} per (block, token)
constexpr int kDxBytes = kChunk * kTok * 4;                 // 256
constexpr int kRedBytes = 2 * kOut * 4;                     // 2048: head and tail partial of the warp
constexpr int kWarpBytes = kXbBytes + kDxBytes + kRedBytes; // 3328
struct Params {
    const uint4* planes;            // {sign0, sign1, salient0, salient1} per row (layers with zero weights) ...
    const uint2* sign_planes;       // ... or the compact sign words only (kCompact)
    const float2* affine;
    const uint4* xb;
    const int* dx;
    float* y;
    int64_t ldy;
    unsigned long long* ws;
    int M, N;
    uint32_t tiles_c, groups, tiles_per_group, rgs, slots, q, rem;
};
}  // namespace bsk

template <bool kCompact>
__global__ void __launch_bounds__(bsk::kThreads, 4) bireal_sk_kernel(const bsk::Params p) {
    using namespace bsk;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_hrg[kWarps], s_trg[kWarps];
    __shared__ uint32_t s_meta[8];
    constexpr uint32_t kNone = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint8_t* wsm = smem + wid * kWarpBytes;
    uint4* xs = reinterpret_cast<uint4*>(wsm);                               // [kChunk][kTok]
    int* dxs = reinterpret_cast<int*>(wsm + kXbBytes);                       // [kChunk][kTok]
    float* head_red = reinterpret_cast<float*>(wsm + kXbBytes + kDxBytes);   // [kTok][32]
    float* tail_red = head_red + kOut;

    const uint32_t TC = p.tiles_c, q = p.q, rem = p.rem;
    auto wstart = [&](uint32_t gw) { return gw * q + min(gw, rem); };
    const uint32_t gw0 = blockIdx.x * kWarps;
    const uint32_t c_lo = wstart(gw0), c_hi = wstart(gw0 + kWarps);
    const uint32_t w_lo = wstart(gw0 + wid), w_hi = wstart(gw0 + wid + 1u);
    const int m0 = blockIdx.y * kTok;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    uint32_t rg = 0, kb = 0;
    if (w_lo < w_hi) { rg = w_lo / TC; kb = w_lo - rg * TC; }
    const uint32_t rg_first = rg;
    auto load_sg = [&](uint32_t r, uint32_t k) {             // tile-major planes: [tile row][k-block][128 rows]
        const size_t i = ((size_t)(r / kRgPerTile) * TC + k) * kTileRows + (r % kRgPerTile) * kRgRows + lane;
        if constexpr (kCompact) {
            const uint2 v = __ldg(p.sign_planes + i);
            return make_uint4(v.x, v.y, 0u, 0u);
        } else {
            return __ldg(p.planes + i);
        }
    };
    uint4 sg = make_uint4(0, 0, 0, 0);
    if (w_lo < w_hi) sg = load_sg(rg, kb);
    const bool grouped = p.groups > 1;
    uint32_t cur_g = grouped ? kb / p.tiles_per_group : 0u;
    float2 af = make_float2(0.f, 0.f);
    if (w_lo < w_hi) af = __ldg(p.affine + (size_t)(rg * kRgRows + lane) * p.groups + cur_g);
    if (lane == 0) { s_hrg[wid] = kNone; s_trg[wid] = kNone; }
    if (tid == 0) {
        const uint32_t rg_a = c_lo / TC, rg_b = (c_hi - 1u) / TC;
        auto owner = [&](uint32_t b) {
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

    int d1[kTok], d0[kTok];                                   // sum of sign(x) over the row's hi / lo positions
    float acc[kTok];
#pragma unroll
    for (int m = 0; m < kTok; ++m) { d1[m] = d0[m] = 0; acc[m] = 0.f; }
    auto fold = [&]() {                                       // counts -> values with the current {lo,hi}
#pragma unroll
        for (int m = 0; m < kTok; ++m) { acc[m] += af.y * (float)d1[m] + af.x * (float)d0[m]; d1[m] = d0[m] = 0; }
    };

    // the activation bits come from the binarize kernel just before this one in the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const uint32_t sm_tok = lane >> 2, sm_j = lane & 3u;      // staging: lane -> token, blocks j and j+4 of the chunk
    const bool sm_ok = (m0 + (int)sm_tok) < p.M;
    uint32_t ci = kChunk;                                     // index of the current block in the staged chunk
    uint32_t kb_c = kb;                                       // k-block of the chunk's first block
    for (uint32_t blk = w_lo; blk < w_hi; ++blk) {
        const bool more = blk + 1 < w_hi;
        if (ci == kChunk) {                                   // stage the bit planes of the next (up to) 8 blocks of the run
            ci = 0;
            kb_c = kb;
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t i = sm_j + 4u * h;
                uint32_t k = kb_c + i;
                while (k >= TC) k -= TC;                      // the run may continue in the next row group
                uint4 v = make_uint4(0, 0, 0, 0);
                int dv = 0;
                if (sm_ok && blk + i < w_hi) {
                    v = __ldg(p.xb + (size_t)(m0 + sm_tok) * TC + k);
                    dv = __ldg(p.dx + (size_t)(m0 + sm_tok) * TC + k);
                }
                xs[i * kTok + sm_tok] = v;
                dxs[i * kTok + sm_tok] = dv;
            }
            __syncwarp();
        }
        if (grouped) {
            const uint32_t g = kb / p.tiles_per_group;
            if (g != cur_g) {
                fold();
                cur_g = g;
                af = __ldg(p.affine + (size_t)(rg * kRgRows + lane) * p.groups + g);
            }
        }
        const uint4 b = sg;
        ++kb;
        const bool rg_end = kb == TC;
        if (more) sg = load_sg(rg_end ? rg + 1u : rg, rg_end ? 0u : kb);             // in-place prefetch of the next block
        bool any_sal = false;                                 // zero weights (sign(0) = 0) in this block of rows?
        if constexpr (!kCompact) any_sal = __any_sync(0xffffffffu, (b.z | b.w) != 0u);
        const uint32_t nb0 = ~(b.x | b.z), nb1 = ~(b.y | b.w);
#pragma unroll
        for (int m = 0; m < kTok; ++m) {
            const uint4 xv = xs[ci * kTok + m];               // same address in every lane: a broadcast
            const int t1 = __popc(b.x & xv.x) + __popc(b.y & xv.y) - __popc(b.x & xv.z) - __popc(b.y & xv.w);
            d1[m] += t1;
            if (any_sal) d0[m] += __popc(nb0 & xv.x) + __popc(nb1 & xv.y) - __popc(nb0 & xv.z) - __popc(nb1 & xv.w);
            else d0[m] += dxs[ci * kTok + m] - t1;            // no zero weights here: the lo set is the complement
        }
        ++ci;
        if (rg_end || !more) {
            fold();
            const bool whole = rg_end && (w_lo <= rg * TC);
            if (whole) {
                const int orow = (int)(rg * kRgRows + lane);
                if (orow < p.N) {
#pragma unroll
                    for (int m = 0; m < kTok; ++m)
                        if (m0 + m < p.M) p.y[(int64_t)(m0 + m) * p.ldy + orow] = acc[m];
                }
            } else {
                float* dst = (rg == rg_first) ? head_red : tail_red;
#pragma unroll
                for (int m = 0; m < kTok; ++m) dst[m * kRgRows + lane] = acc[m];
                if (lane == 0) { if (rg == rg_first) s_hrg[wid] = rg; else s_trg[wid] = rg; }
            }
#pragma unroll
            for (int m = 0; m < kTok; ++m) acc[m] = 0.f;
            if (rg_end) {
                kb = 0;
                ++rg;
                if (more) {
                    cur_g = 0;
                    af = __ldg(p.affine + (size_t)(rg * kRgRows + lane) * p.groups);
                }
            }
        }
    }

    __syncthreads();
    if (c_lo >= c_hi) return;
    const uint32_t rg_a = s_meta[0], rg_b = s_meta[1];
    const uint32_t om = tid >> 5, orr = tid & 31u;
    float* yout = p.y + (int64_t)(m0 + om) * p.ldy;
    const bool tok_ok = (m0 + (int)om) < p.M;
    uint32_t hrg[kWarps], trg[kWarps];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { hrg[w] = s_hrg[w]; trg[w] = s_trg[w]; }
    float v_split[2] = {0.f, 0.f};
    for (uint32_t r = rg_a; r <= rg_b; ++r) {
        float v = 0.f;
        bool any = false;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const float* red = reinterpret_cast<const float*>(smem + w * kWarpBytes + kXbBytes + kDxBytes);
            if (hrg[w] == r) { v += red[tid]; any = true; }
            if (trg[w] == r) { v += red[kOut + tid]; any = true; }
        }
        if (!any) continue;
        const bool split = (r == rg_a && s_meta[2]) || (r == rg_b && s_meta[5]);
        if (!split) {
            const int orow = (int)(r * kRgRows + orr);
            if (tok_ok && orow < p.N) yout[orow] = v;
        } else if (r == rg_a) {
            v_split[0] = v;
        } else {
            v_split[1] = v;
        }
    }
    const bool hs = s_meta[2] != 0u, ts = s_meta[5] != 0u;
    if (!hs && !ts) return;
#pragma unroll
    for (int f = 1; f >= 0; --f) {
        if (!(f == 0 ? hs : ts)) continue;
        const uint32_t r = f == 0 ? rg_a : rg_b, slot = s_meta[f == 0 ? 3 : 6], expected = s_meta[f == 0 ? 4 : 7];
        unsigned long long* part = p.ws + (((size_t)blockIdx.y * p.rgs + r) * p.slots) * kOut + tid;
        const float v = v_split[f];
        if (slot + 1u < expected) {
            const unsigned long long u = (1ull << 32) | (unsigned long long)__float_as_uint(v);
            asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(part + (size_t)slot * kOut), "l"(u) : "memory");
        } else {
            const int orow = (int)(r * kRgRows + orr);
            float sum = 0.f;
            for (uint32_t k = 0; k + 1u < expected; ++k) {
                unsigned long long u;
                do {
                    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(u) : "l"(part + (size_t)k * kOut) : "memory");
                } while ((u >> 32) == 0ull);
                sum += __uint_as_float((uint32_t)u);
                asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(part + (size_t)k * kOut), "l"(0ull) : "memory");
            }
            sum += v;
            if (tok_ok && orow < p.N) yout[orow] = sum;
        }
    }
}

static int bireal_sk_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PBL_BIREAL_SK"); v = (e && *e) ? atoi(e) : 1; }
    return v;
}

// bireal_sk_kernel handles bsk::kTok = 8 tokens per blockIdx.y whatever M is (the decode kernel's plan switches to
// 16-token passes above 8 tokens, so its `passes` must not be used here)
static uint32_t bireal_passes(int64_t M) { return (uint32_t)((M + bsk::kTok - 1) / bsk::kTok); }

static size_t bireal_fixup_bytes(const Layer& L, int64_t M) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { (void)cudaGetLastError(); sms = 148; }
    uint32_t pl[8];
    decode_plan(L.N, L.K, 1, sms, 4, pl);       // block partition only: this kernel's token passes are its own (8 tokens each)
    return (size_t)bireal_passes(M) * pl[1] * pl[6] * bsk::kOut * 8u;
}

size_t bireal_workspace_bytes(const Layer& L, int64_t M) {
    return (size_t)M * (size_t)L.tiles_c * (sizeof(uint4) + sizeof(int)) + 256;
}

// zero-initialised workspace the stream-K kernel needs for its cross-CTA reduction (same contract as the decode
// kernel's: left zero by every call, one stream at a time); 0 when the layer does not qualify for that kernel
size_t bireal_fixup_workspace_bytes(const Layer& L, int64_t M) {
    if (!bireal_sk_enabled() || M <= 0 || M > 64) return 0;   // larger M: the passes would re-read the planes
    return bireal_fixup_bytes(L, M);
}
int launch_bireal(const Layer& L, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy, int64_t M, void* workspace,
                  void* fixup_ws, size_t fixup_bytes, cudaStream_t s) {
    uint4* xb = reinterpret_cast<uint4*>(workspace);
    int* dx = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + (((size_t)M * L.tiles_c * sizeof(uint4) + 255) / 256) * 256);
    // (xb and dx fit in bireal_scratch_bytes: M*tiles_c*20 + 256 >= the 256-rounded xb region + M*tiles_c*4)
    const dim3 gb((unsigned)L.tiles_c, (unsigned)M);
    switch (x_dtype) {
        case PBL_F16: bireal_binarize_kernel<__half><<<gb, 64, 0, s>>>((const __half*)x, ldx, L.K, (int)L.tiles_c, xb, dx); break;
        case PBL_BF16: bireal_binarize_kernel<__nv_bfloat16><<<gb, 64, 0, s>>>((const __nv_bfloat16*)x, ldx, L.K, (int)L.tiles_c, xb, dx); break;
        case PBL_F32: bireal_binarize_kernel<float><<<gb, 64, 0, s>>>((const float*)x, ldx, L.K, (int)L.tiles_c, xb, dx); break;
        default: set_error("pbl_bireal_forward: bad x dtype %d", x_dtype); return PBL_ERR_DTYPE;
    }
    int rc = check_cuda(cudaGetLastError(), "bireal binarize launch");
    if (rc) return rc;
    if (fixup_ws && fixup_bytes >= bireal_fixup_workspace_bytes(L, M) && bireal_fixup_workspace_bytes(L, M) > 0) {
        // a zeroed reduction workspace was given: stream-K kernel, balanced over all SMs
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        uint32_t pl[8];
        decode_plan(L.N, L.K, 1, sms, 4, pl);
        bsk::Params p;
        p.planes = L.planes; p.sign_planes = L.sign_planes; p.affine = L.affine; p.xb = xb; p.dx = dx; p.y = y; p.ldy = ldy;
        p.ws = reinterpret_cast<unsigned long long*>(fixup_ws);
        p.M = (int)M; p.N = (int)L.N;
        p.tiles_c = (uint32_t)L.tiles_c; p.groups = (uint32_t)L.groups; p.tiles_per_group = (uint32_t)L.tiles_per_group;
        p.rgs = pl[1]; p.slots = pl[6]; p.q = pl[4]; p.rem = pl[5];
        const int smem = bsk::kWarps * bsk::kWarpBytes;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pl[2], bireal_passes(M));
        cfg.blockDim = dim3(bsk::kThreads);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = L.sign_planes ? cudaLaunchKernelEx(&cfg, bireal_sk_kernel<true>, p) : cudaLaunchKernelEx(&cfg, bireal_sk_kernel<false>, p);
        count_launch(2);
        return check_cuda(le, "bireal stream-K launch");
    }
    const unsigned gx = (unsigned)(L.n_pad / kRgRows);
#define PBL_BR_LAUNCH(MT)                                                                                                      \
    do {                                                                                                                       \
        const dim3 g2(gx, (unsigned)((M + MT - 1) / MT));                                                                      \
        if (L.sign_planes)                                                                                                     \
            bireal_xnor_kernel<MT, true><<<g2, kBrWarps * 32, 0, s>>>(L.planes, L.sign_planes, L.affine, xb, dx, y, ldy, M, L.N, \
                                                                      (int)L.tiles_c, (int)L.groups, L.tiles_per_group);       \
        else                                                                                                                   \
            bireal_xnor_kernel<MT, false><<<g2, kBrWarps * 32, 0, s>>>(L.planes, L.sign_planes, L.affine, xb, dx, y, ldy, M, L.N, \
                                                                       (int)L.tiles_c, (int)L.groups, L.tiles_per_group);      \
    } while (0)
    if (M <= 1) PBL_BR_LAUNCH(1);
    else if (M <= 2) PBL_BR_LAUNCH(2);
    else if (M <= 4) PBL_BR_LAUNCH(4);
    else PBL_BR_LAUNCH(8);
#undef PBL_BR_LAUNCH
    count_launch(2);
    return check_cuda(cudaGetLastError(), "bireal xnor launch");
}

}  // namespace pbl
