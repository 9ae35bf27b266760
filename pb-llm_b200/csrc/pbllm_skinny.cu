// Bit-plane skinny kernel on tensor cores (decode regime: M <= 8 tokens per pass), fp16 / bf16.
//
// Same math as gemv_kernel (y = x . w_sim^T + b from the packed form, fp32 accumulation) but the
// dense {lo,hi} part no longer costs one CUDA-core op per weight per token:
//   * dense part: the warp's 32x64 weight block is rebuilt IN REGISTERS as mma.sync m16n8k16 A
//     fragments (bit -> PRMT byte-sign replicate -> LOP3 {lo,hi} select, ~1.25 ALU ops/weight for
//     all 8 tokens at once) and multiplied with the activations as the 16x8 B operand; nothing
//     of the weight tile ever touches shared memory.
//   * salient part: lane = row walks its salient bits and accumulates (v - lo) * x[m] from an fp32
//     transposed activation tile (one pair of LDS.128 gives all 8 tokens).
// HBM-bound target: the packed stream (planes 0.25 B/weight + values) is read exactly once.
// CTA = one 32-row group x (up to) 8 tokens; kWarps warps split the k-blocks, deterministic
// shared-memory reduction at the end.
#include <type_traits>

#include "pbllm_common.cuh"

namespace pbl {

namespace sk {
constexpr int kWarps = 16;
constexpr int kTok = 8;                         // tokens per pass (mma N)
constexpr int kXrStride = kTileCols + 8;        // halves per token row (+8: conflict-free B-fragment LDS)
constexpr int kXrBytes = kTok * kXrStride * 2;  // 1152
constexpr int kXtBytes = kTileCols * kTok * 4;  // 2048: fp32 [64 cols][8 tokens]
constexpr int kScrBytes = 1024;                 // staged salient values (512 x 16 bit)
constexpr int kWarpBytes = kXrBytes + kXtBytes + kScrBytes;  // 4224
static_assert(kWarpBytes % 16 == 0, "alignment");
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

template <typename T>
__global__ void __launch_bounds__(sk::kWarps * 32)
skinny_mma_kernel(const uint4* __restrict__ planes, const uint32_t* __restrict__ vptr, const uint16_t* __restrict__ vals,
                  const float2* __restrict__ affine, const float* __restrict__ bias, const T* __restrict__ x, int64_t ldx,
                  T* __restrict__ y, int64_t ldy, int M, int N, int K, int tiles_c, int groups, int tiles_per_group) {
    using namespace sk;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* wsm = smem + wid * kWarpBytes;
    uint16_t* xr = reinterpret_cast<uint16_t*>(wsm);                       // [8][72] 16-bit
    float* xt = reinterpret_cast<float*>(wsm + kXrBytes);                  // [64][8] fp32
    const uint32_t scr = (uint32_t)__cvta_generic_to_shared(wsm + kXrBytes + kXtBytes);

    const int rg = blockIdx.x;                   // global 32-row group
    const int tr = rg / kRgPerTile, rgi = rg % kRgPerTile;
    const int row = rg * kRgRows + lane;         // lane's own row (salient part)
    const int m0 = blockIdx.y * kTok;
    const int g4 = lane >> 2, t4 = lane & 3;     // mma fragment coordinates
    const uint16_t* x16 = reinterpret_cast<const uint16_t*>(x);

    float cacc[2][4];                            // dense accumulators (two m16 tiles)
    float sacc[kTok];                            // salient accumulators (lane = row)
#pragma unroll
    for (int i = 0; i < 4; ++i) cacc[0][i] = cacc[1][i] = 0.f;
#pragma unroll
    for (int m = 0; m < kTok; ++m) sacc[m] = 0.f;

    // ---- prefetch helpers -----------------------------------------------------------------------
    struct Meta { uint4 pw; uint32_t cs, ce; };
    auto load_meta = [&](int kb) {
        Meta mt;
        mt.pw = make_uint4(0, 0, 0, 0);
        mt.cs = mt.ce = 0;
        if (kb < tiles_c) {
            const int64_t tile = (int64_t)tr * tiles_c + kb;
            mt.pw = __ldg(planes + tile * kTileRows + rgi * kRgRows + lane);
            mt.cs = __ldg(vptr + tile * kRgPerTile + rgi);
            mt.ce = __ldg(vptr + tile * kRgPerTile + rgi + 1);
        }
        return mt;
    };
    auto load_vals = [&](const Meta& mt, uint4& q0, uint4& q1) {
        const uint32_t b0 = (mt.cs * 2u) & ~15u, b1 = mt.ce * 2u;
        const uint8_t* base = reinterpret_cast<const uint8_t*>(vals);
        const uint32_t o0 = b0 + 16u * lane, o1 = o0 + 512u;
        if (o0 < b1) q0 = __ldg(reinterpret_cast<const uint4*>(base + o0));
        if (o1 < b1) q1 = __ldg(reinterpret_cast<const uint4*>(base + o1));
    };
    const bool x_al32 = ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 3u) == 0);
    auto load_x = [&](int kb, uint32_t (&xv)[kTok]) {   // lane holds columns 2*lane, 2*lane+1 of every token
        const int col = kb * kTileCols + 2 * lane;
#pragma unroll
        for (int m = 0; m < kTok; ++m) {
            uint32_t v = 0;
            if (kb < tiles_c && m0 + m < M) {
                const uint16_t* p = x16 + (int64_t)(m0 + m) * ldx + col;
                if (col + 1 < K) {
                    if (x_al32) v = __ldg(reinterpret_cast<const uint32_t*>(p));
                    else v = (uint32_t)p[0] | ((uint32_t)p[1] << 16);
                } else if (col < K) v = (uint32_t)p[0];
            }
            xv[m] = v;
        }
    };

    int cur_g = -1;
    uint32_t LLa[2] = {0, 0}, DDa[2] = {0, 0}, LLb[2] = {0, 0}, DDb[2] = {0, 0};   // fragment rows g4 / g4+8 of each tile
    float my_lo = 0.f;

    Meta mt0 = load_meta(wid), mt1 = load_meta(wid + kWarps);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
    uint32_t xv[kTok];
    load_vals(mt0, q0, q1);
    load_x(wid, xv);

    for (int kb = wid; kb < tiles_c; kb += kWarps) {
        const uint4 pw = mt0.pw;
        const uint32_t cs = mt0.cs, ce = mt0.ce;
        const uint4 v0 = q0, v1 = q1;
        uint32_t xc[kTok];
#pragma unroll
        for (int m = 0; m < kTok; ++m) xc[m] = xv[m];
        // next items' global loads first
        const Meta mt2 = load_meta(kb + 2 * kWarps);
        load_vals(mt1, q0, q1);
        load_x(kb + kWarps, xv);
        mt0 = mt1;
        mt1 = mt2;

        const int g = kb / tiles_per_group;
        if (g != cur_g) {
            cur_g = g;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 aa = __ldg(affine + (int64_t)(rg * kRgRows + 16 * h + g4) * groups + g);
                const float2 ab = __ldg(affine + (int64_t)(rg * kRgRows + 16 * h + g4 + 8) * groups + g);
                uint32_t lo = sk_bits16<T>(aa.x), hi = sk_bits16<T>(aa.y);
                LLa[h] = lo | (lo << 16); DDa[h] = (lo ^ hi) * 0x10001u;
                lo = sk_bits16<T>(ab.x); hi = sk_bits16<T>(ab.y);
                LLb[h] = lo | (lo << 16); DDb[h] = (lo ^ hi) * 0x10001u;
            }
            my_lo = __ldg(affine + (int64_t)row * groups + g).x;
        }

        // ---- stage activations (row layout for B fragments, fp32 transposed for the salient part)
        //      and this row group's salient values
        __syncwarp();
        {
            float4 lo4a, lo4b, hi4a, hi4b;
            float2 f;
#pragma unroll
            for (int m = 0; m < kTok; ++m) *reinterpret_cast<uint32_t*>(xr + m * kXrStride + 2 * lane) = xc[m];
            f = sk_unpack2<T>(xc[0]); lo4a.x = f.x; hi4a.x = f.y;
            f = sk_unpack2<T>(xc[1]); lo4a.y = f.x; hi4a.y = f.y;
            f = sk_unpack2<T>(xc[2]); lo4a.z = f.x; hi4a.z = f.y;
            f = sk_unpack2<T>(xc[3]); lo4a.w = f.x; hi4a.w = f.y;
            f = sk_unpack2<T>(xc[4]); lo4b.x = f.x; hi4b.x = f.y;
            f = sk_unpack2<T>(xc[5]); lo4b.y = f.x; hi4b.y = f.y;
            f = sk_unpack2<T>(xc[6]); lo4b.z = f.x; hi4b.z = f.y;
            f = sk_unpack2<T>(xc[7]); lo4b.w = f.x; hi4b.w = f.y;
            float4* dst = reinterpret_cast<float4*>(xt + (2 * lane) * kTok);
            dst[0] = lo4a; dst[1] = lo4b; dst[2] = hi4a; dst[3] = hi4b;
            const uint32_t b0 = (cs * 2u) & ~15u;
            const uint32_t o0 = 16u * lane, o1 = o0 + 512u;
            if (b0 + o0 < ce * 2u) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(scr + o0), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w) : "memory");
            if (b0 + o1 < ce * 2u) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(scr + o1), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
        }
        __syncwarp();

        // ---- dense part on tensor cores: A fragments from bits, B fragments from xr -------------------
        uint32_t bfr[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bfr[q][0] = *reinterpret_cast<const uint32_t*>(xr + g4 * kXrStride + 16 * q + 2 * t4);
            bfr[q][1] = *reinterpret_cast<const uint32_t*>(xr + g4 * kXrStride + 16 * q + 2 * t4 + 8);
        }
        const uint32_t sh7 = 7u - 2u * t4, sh6 = 6u - 2u * t4;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ra = 16 * h + g4, rb = ra + 8;
#pragma unroll
            for (int wd = 0; wd < 2; ++wd) {
                const uint32_t own = wd ? pw.y : pw.x;
                const uint32_t sa = __shfl_sync(0xffffffffu, own, ra), sb = __shfl_sync(0xffffffffu, own, rb);
                const uint32_t a7 = sa << sh7, a6 = sa << sh6, b7 = sb << sh7, b6 = sb << sh6;
                uint32_t fa[4], fb[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t sel = 0x8888u | (uint32_t)b | ((uint32_t)b << 4) | ((uint32_t)(4 + b) << 8) | ((uint32_t)(4 + b) << 12);
                    fa[b] = sk_sel(LLa[h], DDa[h], sk_prmt(a7, a6, sel));
                    fb[b] = sk_sel(LLb[h], DDb[h], sk_prmt(b7, b6, sel));
                }
                mma_16816<T>(cacc[h], fa[0], fb[0], fa[1], fb[1], bfr[2 * wd][0], bfr[2 * wd][1]);
                mma_16816<T>(cacc[h], fa[2], fb[2], fa[3], fb[3], bfr[2 * wd + 1][0], bfr[2 * wd + 1][1]);
            }
        }

        // ---- salient part: lane = row, (v - lo) * x over the row's salient columns -----------------------
        {
            const uint32_t b0 = (cs * 2u) & ~15u;
            const uint32_t idx0 = (cs - (b0 >> 1)) + warp_excl_scan(__popc(pw.z) + __popc(pw.w), lane);
            uint32_t rm0 = __brev(pw.z), rm1 = __brev(pw.w);
            auto fma8 = [&](uint32_t j, uint16_t v16) {
                const float c = sk_val<T>(v16) - my_lo;
                const float4 xa = *reinterpret_cast<const float4*>(xt + j * kTok);
                const float4 xb = *reinterpret_cast<const float4*>(xt + j * kTok + 4);
                sacc[0] = fmaf(c, xa.x, sacc[0]); sacc[1] = fmaf(c, xa.y, sacc[1]);
                sacc[2] = fmaf(c, xa.z, sacc[2]); sacc[3] = fmaf(c, xa.w, sacc[3]);
                sacc[4] = fmaf(c, xb.x, sacc[4]); sacc[5] = fmaf(c, xb.y, sacc[5]);
                sacc[6] = fmaf(c, xb.z, sacc[6]); sacc[7] = fmaf(c, xb.w, sacc[7]);
            };
            if (ce - (b0 >> 1) <= 512u) {            // warp-uniform: the whole chunk is staged in shared memory
                uint32_t sa = scr + idx0 * 2u;
                while (rm0 | rm1) {
                    uint32_t j;
                    if (rm0) { j = (uint32_t)__clz(rm0); rm0 &= ~(0x80000000u >> j); }
                    else { j = (uint32_t)__clz(rm1); rm1 &= ~(0x80000000u >> j); j += 32u; }
                    uint16_t v16;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v16) : "r"(sa) : "memory");
                    sa += 2u;
                    fma8(j, v16);
                }
            } else {                                   // rare: very dense chunk, tail read from global
                uint32_t idx = idx0;
                while (rm0 | rm1) {
                    uint32_t j;
                    if (rm0) { j = (uint32_t)__clz(rm0); rm0 &= ~(0x80000000u >> j); }
                    else { j = (uint32_t)__clz(rm1); rm1 &= ~(0x80000000u >> j); j += 32u; }
                    uint16_t v16;
                    if (idx < 512u) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v16) : "r"(scr + idx * 2u) : "memory");
                    else v16 = __ldg(vals + (b0 >> 1) + idx);
                    ++idx;
                    fma8(j, v16);
                }
            }
        }
    }

    // ---- combine dense fragments + salient partials per warp, then reduce across warps (split-K) ----
    __syncwarp();
    float* red = reinterpret_cast<float*>(wsm);   // [32 rows][8 tokens] fp32 = 1 KB, reuses the xr/xt area
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        red[(16 * h + g4) * kTok + 2 * t4] = cacc[h][0];
        red[(16 * h + g4) * kTok + 2 * t4 + 1] = cacc[h][1];
        red[(16 * h + g4 + 8) * kTok + 2 * t4] = cacc[h][2];
        red[(16 * h + g4 + 8) * kTok + 2 * t4 + 1] = cacc[h][3];
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < kTok; ++m) red[lane * kTok + m] += sacc[m];
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

int launch_skinny(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    const dim3 grid((unsigned)(L.n_pad / kRgRows), (unsigned)((M + sk::kTok - 1) / sk::kTok));
    const int smem = sk::kWarps * sk::kWarpBytes;
    static bool attr_set[2] = {false, false};
    const int which = L.dtype == PBL_F16 ? 0 : 1;
    if (!attr_set[which]) {
        cudaError_t e = which == 0
            ? cudaFuncSetAttribute(skinny_mma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
            : cudaFuncSetAttribute(skinny_mma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int rc = check_cuda(e, "cudaFuncSetAttribute(skinny smem)");
        if (rc) return rc;
        attr_set[which] = true;
    }
    if (which == 0)
        skinny_mma_kernel<__half><<<grid, sk::kWarps * 32, smem, s>>>(L.planes, L.vptr, (const uint16_t*)L.vals, L.affine, L.bias,
                                                                     (const __half*)x, ldx, (__half*)y, ldy, (int)M, (int)L.N,
                                                                     (int)L.K, (int)L.tiles_c, (int)L.groups, L.tiles_per_group);
    else
        skinny_mma_kernel<__nv_bfloat16><<<grid, sk::kWarps * 32, smem, s>>>(L.planes, L.vptr, (const uint16_t*)L.vals, L.affine,
                                                                             L.bias, (const __nv_bfloat16*)x, ldx,
                                                                             (__nv_bfloat16*)y, ldy, (int)M, (int)L.N, (int)L.K,
                                                                             (int)L.tiles_c, (int)L.groups, L.tiles_per_group);
    count_launch();
    return check_cuda(cudaGetLastError(), "skinny launch");
}

}  // namespace pbl
