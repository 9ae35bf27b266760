// Bit-plane skinny kernel on tensor cores (decode regime: M <= 8 tokens per pass), fp16 / bf16.
//
// Same math as gemv_kernel (y = x . w_sim^T + b from the packed form, fp32 accumulation), organised
// as a per-warp mini-GEMM: each warp rebuilds its 32x64 block of the EXACT fp16/bf16 w_sim tile in a
// private 4 KB shared-memory buffer (bit -> PRMT byte-sign replicate -> LOP3 {lo,hi} select, 8
// conflict-free STS.128 per row, then the row's salient values patched over their positions -- the
// same expansion as the tcgen05 kernels, so the tile is bit-identical to theirs), reads it back as
// mma.sync m16n8k16 A fragments with ldmatrix (XOR-swizzled rows, conflict-free) and multiplies with
// the activations (B fragment, 8 tokens).  The salient entries therefore ride the tensor cores too:
// no per-element FMA chain, cost independent of M <= 8.
// HBM-bound target: the packed stream (planes 0.25 B/weight + values) is read exactly once.
// CTA = one 32-row group x (up to) 8 tokens; kWarps warps split the k-blocks, deterministic
// shared-memory reduction at the end.
#include <cstdlib>
#include <type_traits>

#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace sk {
constexpr int kTok = 8;                         // tokens per pass (mma N)
constexpr int kXrStride = kTileCols + 8;        // halves per token row (+8: conflict-free B-fragment LDS)
constexpr int kXrBytes = kTok * kXrStride * 2;  // 1152
constexpr int kTileBytes = kRgRows * kTileCols * 2;  // 4096: the warp's 32x64 16-bit weight tile (swizzled)
constexpr int kScrBytes = 1024;                 // staged salient values (512 x 16 bit)
constexpr int kWarpBytes = kTileBytes + kXrBytes + kScrBytes;  // 6272 per warp
static_assert(kWarpBytes % 128 == 0, "alignment");
}  // namespace sk

__device__ __forceinline__ uint32_t sk_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t sk_sel(uint32_t a, uint32_t b, uint32_t c) {  // a ^ (b & c)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
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

template <typename T> __device__ __forceinline__ uint32_t sk_bits16(float v);
template <> __device__ __forceinline__ uint32_t sk_bits16<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ uint32_t sk_bits16<__nv_bfloat16>(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

template <typename T> __device__ __forceinline__ float2 sk_unpack2(uint32_t v);
template <> __device__ __forceinline__ float2 sk_unpack2<__half>(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
template <> __device__ __forceinline__ float2 sk_unpack2<__nv_bfloat16>(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ float sk_val(uint16_t v);
template <> __device__ __forceinline__ float sk_val<__half>(uint16_t v) { return __half2float(__ushort_as_half(v)); }
template <> __device__ __forceinline__ float sk_val<__nv_bfloat16>(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }

template <typename T, int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
skinny_mma_kernel(const uint4* __restrict__ planes, const uint32_t* __restrict__ vptr, const uint16_t* __restrict__ vals,
                  const float2* __restrict__ affine, const float* __restrict__ bias, const T* __restrict__ x, int64_t ldx,
                  T* __restrict__ y, int64_t ldy, int M, int N, int K, int tiles_c, int groups, int tiles_per_group) {
    using namespace sk;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* wsm = smem + wid * kWarpBytes;
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(wsm);                    // [32 rows][64] 16-bit, swizzled
    uint16_t* xr = reinterpret_cast<uint16_t*>(wsm + kTileBytes);                       // [8][72] 16-bit
    const uint32_t scr = (uint32_t)__cvta_generic_to_shared(wsm + kTileBytes + kXrBytes);

    const int rg = blockIdx.x;                   // global 32-row group
    const int tr = rg / kRgPerTile, rgi = rg % kRgPerTile;
    const int row = rg * kRgRows + lane;         // lane's own row (expansion: thread = weight row)
    const int m0 = blockIdx.y * kTok;
    const int g4 = lane >> 2, t4 = lane & 3;     // mma fragment coordinates
    const uint16_t* x16 = reinterpret_cast<const uint16_t*>(x);
    const uint32_t r7 = (uint32_t)(lane & 7);
    const uint32_t brow = tile_s + (uint32_t)lane * 128u;      // my row of the tile (128 B), chunks XOR-swizzled by r7

    float cacc[2][4];                            // accumulators (two m16 tiles x 8 tokens)
#pragma unroll
    for (int i = 0; i < 4; ++i) cacc[0][i] = cacc[1][i] = 0.f;

    // ---- prefetch helpers: pointers advance by kWarps k-blocks per iteration (no per-iteration index math) ----
    struct Meta { uint4 pw; uint32_t cs, ce; };
    const uint4* pl_ptr = planes + ((int64_t)tr * tiles_c + wid) * kTileRows + rgi * kRgRows + lane;
    const uint32_t* vp_ptr = vptr + ((int64_t)tr * tiles_c + wid) * kRgPerTile + rgi;
    int kb_meta = wid;                                   // k-block the meta pointers refer to
    auto load_meta = [&]() {
        Meta mt;
        mt.pw = make_uint4(0, 0, 0, 0);
        mt.cs = mt.ce = 0;
        if (kb_meta < tiles_c) {
            mt.pw = __ldg(pl_ptr);
            mt.cs = __ldg(vp_ptr);
            mt.ce = __ldg(vp_ptr + 1);
        }
        pl_ptr += kWarps * kTileRows;
        vp_ptr += kWarps * kRgPerTile;
        kb_meta += kWarps;
        return mt;
    };
    auto load_vals = [&](const Meta& mt, uint4& q0, uint4& q1) {
        const uint32_t b0 = (mt.cs * 2u) & ~15u, b1 = mt.ce * 2u;
        const uint8_t* base = reinterpret_cast<const uint8_t*>(vals);
        const uint32_t o0 = b0 + 16u * lane, o1 = o0 + 512u;
        if (o0 < b1) q0 = __ldg(reinterpret_cast<const uint4*>(base + o0));
        if (o1 < b1) q1 = __ldg(reinterpret_cast<const uint4*>(base + o1));
    };
    // activations: lane -> (token = lane>>2, 16-column segment = lane&3) of the 8 x 64 tile: two 16 B loads.
    // Fast path needs 16 B-aligned rows and whole 16-column segments; anything else takes the generic path.
    const int xtok = lane >> 2, xseg = lane & 3;
    const bool x_fast = ((ldx & 7) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((K & 15) == 0);
    const bool x_tok_ok = (m0 + xtok) < M;
    const uint16_t* x_lane = x16 + (int64_t)(m0 + (x_tok_ok ? xtok : 0)) * ldx + 16 * xseg;
    auto load_x = [&](int kb, uint4& xa, uint4& xb2) {
        xa = make_uint4(0, 0, 0, 0);
        xb2 = make_uint4(0, 0, 0, 0);
        if (kb >= tiles_c) return;
        const int col = kb * kTileCols + 16 * xseg;
        if (x_fast) {
            if (x_tok_ok && col < K) {
                const uint4* p4 = reinterpret_cast<const uint4*>(x_lane + (int64_t)kb * kTileCols);
                xa = __ldg(p4);
                xb2 = __ldg(p4 + 1);
            }
        } else if (x_tok_ok) {                            // generic: element loads with bounds checks
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = col + 2 * i;
                const uint16_t* p = x_lane + (int64_t)kb * kTileCols + 2 * i;
                uint32_t v = 0;
                if (c < K) v = (uint32_t)p[0];
                if (c + 1 < K) v |= (uint32_t)p[1] << 16;
                w[i] = v;
            }
            xa = make_uint4(w[0], w[1], w[2], w[3]);
            xb2 = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };

    // Programmatic dependent launch: let the next kernel in the stream start its own weight prefetch now;
    // everything below up to griddepcontrol.wait touches only immutable packed weights.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const bool grouped = groups > 1;
    int cur_g = (grouped && wid < tiles_c) ? wid / tiles_per_group : 0;
    float2 a_first = __ldg(affine + (int64_t)row * groups + cur_g);      // overlaps the meta / value loads
    Meta mt0 = load_meta(), mt1 = load_meta();
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
    load_vals(mt0, q0, q1);
    uint32_t LL, DD;
    {
        const uint32_t lo = sk_bits16<T>(a_first.x), hi = sk_bits16<T>(a_first.y);
        LL = lo | (lo << 16);
        DD = (lo ^ hi) * 0x10001u;
    }
    // activations (and y) belong to the producer kernel: wait for it before the first read of x
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint4 xn0, xn1;
    load_x(wid, xn0, xn1);

    // ldmatrix source address for this lane: matrix i = lane>>3 -> (row block i&1, k chunk i>>1); the k16 step q
    // only flips address bits 5-6 (chunk = (kc ^ r7) ^ 2q), so addr(q) = base ^ (q << 5).
    const uint32_t lm_row = (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8);
    const uint32_t lm_kc = (uint32_t)(lane >> 4);
    const uint32_t lm_base0 = tile_s + lm_row * 128u + (((lm_kc ^ (lm_row & 7u))) << 4);
    const uint32_t lm_base1 = lm_base0 + 16u * 128u;
    const uint32_t xr_s = (uint32_t)__cvta_generic_to_shared(xr);
    const uint32_t xr_st = xr_s + (uint32_t)(xtok * kXrStride + 16 * xseg) * 2u;      // where my two 16 B x segments go
    const uint32_t bfr_s = xr_s + (uint32_t)(g4 * kXrStride + 2 * t4) * 2u;           // my B-fragment words

    for (int kb = wid; kb < tiles_c; kb += kWarps) {
        const uint4 pw = mt0.pw;
        const uint32_t cs = mt0.cs, ce = mt0.ce;
        const uint4 v0 = q0, v1 = q1;
        const uint4 xc0 = xn0, xc1 = xn1;
        // next items' global loads first
        const Meta mt2 = load_meta();
        load_vals(mt1, q0, q1);
        load_x(kb + kWarps, xn0, xn1);
        mt0 = mt1;
        mt1 = mt2;

        if (grouped) {
            const int g = kb / tiles_per_group;
            if (g != cur_g) {
                cur_g = g;
                const float2 a = __ldg(affine + (int64_t)row * groups + g);
                const uint32_t lo = sk_bits16<T>(a.x), hi = sk_bits16<T>(a.y);
                LL = lo | (lo << 16);
                DD = (lo ^ hi) * 0x10001u;
            }
        }

        // ---- stage activations (row layout for the B fragments), then rebuild my row of the exact tile ----
        __syncwarp();                                   // previous iteration's ldmatrix / B-fragment reads are done
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(xr_st), "r"(xc0.x), "r"(xc0.y), "r"(xc0.z), "r"(xc0.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(xr_st + 16u), "r"(xc1.x), "r"(xc1.y), "r"(xc1.z), "r"(xc1.w) : "memory");
        expand_row(pw, LL, DD, brow, r7, cs, ce, v0, v1, scr, vals, (uint32_t)lane);
        __syncwarp();                                   // tile and activations visible to the whole warp

        // ---- tensor cores: A fragments by ldmatrix from the swizzled tile, B fragments from xr ------------
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t bq0, bq1;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(bq0) : "r"(bfr_s + (uint32_t)(32 * q)) : "memory");
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(bq1) : "r"(bfr_s + (uint32_t)(32 * q + 16)) : "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t addr = (h ? lm_base1 : lm_base0) ^ ((uint32_t)q << 5);
                uint32_t a0, a1, a2, a3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                             : "r"(addr)
                             : "memory");
                mma_16816<T>(cacc[h], a0, a1, a2, a3, bq0, bq1);
            }
        }
    }

    // ---- per-warp partials -> shared memory, then reduce across warps (split-K), add bias, store ----
    __syncwarp();
    float* red = reinterpret_cast<float*>(wsm);   // [32 rows][8 tokens] fp32 = 1 KB, reuses the tile area
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        red[(16 * h + g4) * kTok + 2 * t4] = cacc[h][0];
        red[(16 * h + g4) * kTok + 2 * t4 + 1] = cacc[h][1];
        red[(16 * h + g4 + 8) * kTok + 2 * t4] = cacc[h][2];
        red[(16 * h + g4 + 8) * kTok + 2 * t4 + 1] = cacc[h][3];
    }
    __syncthreads();
    if (threadIdx.x < 32 * kTok) {
        const int m = threadIdx.x >> 5, r = threadIdx.x & 31;
        const int orow = rg * kRgRows + r;
        if (orow < N && m0 + m < M) {
            float s = bias ? bias[orow] : 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += reinterpret_cast<const float*>(smem + w * kWarpBytes)[r * kTok + m];
            y[(int64_t)(m0 + m) * ldy + orow] = from_f32<T>(s);
        }
    }
}

bool skinny_supported(const Layer& L, int64_t M) {
    return (L.dtype == PBL_F16 || L.dtype == PBL_BF16) && M > 0 && M <= 65535LL * sk::kTok;
}

template <typename T, int kWarps>
static int launch_skinny_t(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s, int which) {
    const dim3 grid((unsigned)(L.n_pad / kRgRows), (unsigned)((M + sk::kTok - 1) / sk::kTok));
    const int smem = kWarps * sk::kWarpBytes;
    static bool attr_set_dev[64] = {};   // function attributes are per device
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool attr_local = false;
    bool& attr_done = (cur_dev >= 0 && cur_dev < 64) ? attr_set_dev[cur_dev] : attr_local;
    (void)which;
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(skinny_mma_kernel<T, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                            "cudaFuncSetAttribute(skinny smem)");
        if (rc) return rc;
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static int pdl = -1;
    if (pdl < 0) { const char* e = getenv("PBL_PDL"); pdl = (e && *e) ? atoi(e) : 1; }
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const uint16_t* vals16 = (const uint16_t*)L.vals;
    const int Mi = (int)M, Ni = (int)L.N, Ki = (int)L.K, tci = (int)L.tiles_c, gi = (int)L.groups, tpg = L.tiles_per_group;
    const T* xx = (const T*)x;
    T* yy = (T*)y;
    cudaError_t le = cudaLaunchKernelEx(&cfg, skinny_mma_kernel<T, kWarps>, L.planes, L.vptr, vals16, L.affine, L.bias, xx, ldx, yy,
                                        ldy, Mi, Ni, Ki, tci, gi, tpg);
    count_launch();
    return check_cuda(le, "skinny launch");
}

int launch_skinny(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    // 16 warps per CTA (one CTA per SM). 8-warp CTAs (two per SM) measured equal for N >= 11008 and slower for
    // N = 4096 on B200 (profiles/r01 notes); PBL_SK_WARPS=8 keeps the variant reachable for experiments.
    int warps = 16;
    const char* e = getenv("PBL_SK_WARPS");
    if (e && *e) { const int w = atoi(e); warps = (w == 8 || w == 24) ? w : 16; }
    if (L.dtype == PBL_F16) {
        if (warps == 8) return launch_skinny_t<__half, 8>(L, x, ldx, y, ldy, M, s, 0);
        if (warps == 24) return launch_skinny_t<__half, 24>(L, x, ldx, y, ldy, M, s, 0);
        return launch_skinny_t<__half, 16>(L, x, ldx, y, ldy, M, s, 0);
    }
    if (warps == 8) return launch_skinny_t<__nv_bfloat16, 8>(L, x, ldx, y, ldy, M, s, 1);
    if (warps == 24) return launch_skinny_t<__nv_bfloat16, 24>(L, x, ldx, y, ldy, M, s, 1);
    return launch_skinny_t<__nv_bfloat16, 16>(L, x, ldx, y, ldy, M, s, 1);
}

}  // namespace pbl
