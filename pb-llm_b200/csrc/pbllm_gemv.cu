// Bit-plane skinny kernel (decode regime, M small; also the fp32 path for any M).
//
//   y[m][i] = sum_g ( lo_ig * sum_{j in g} x[m][j] + (hi_ig - lo_ig) * sum_{j in g, bit_ij} x[m][j] )
//           + sum_{(i,j) salient} (v_ij - lo_ig) * x[m][j] + bias[i]
// which is x . w_sim^T + b of quant/quantizer.py:86,193 / quant/outlier_quantizer.py:105 with
// w_sim never materialised.  HBM-bound on the packed weight stream (planes 0.25 B/weight +
// salient values); activations are real-valued (W1A16/W1A32), fp32 accumulation.
//
// Mapping: CTA = one 32-row group x one chunk of MT tokens; its kWarps warps split the 64-column
// tiles of the row group (split-K inside the CTA, reduced through shared memory).  Lane = output
// row, so each warp reads one coalesced 512 B plane slab per tile and every x read in the dense
// loop is a shared-memory broadcast.
#include <type_traits>

#include "pbllm_common.cuh"

namespace pbl {

constexpr int kGemvWarps = 8;

template <typename T, int MT>
__global__ void __launch_bounds__(kGemvWarps * 32)
gemv_kernel(const uint4* __restrict__ planes, const uint32_t* __restrict__ vptr, const T* __restrict__ vals,
            const float2* __restrict__ affine, const float* __restrict__ bias, const T* __restrict__ x, int64_t ldx,
            T* __restrict__ y, int64_t ldy, int64_t M, int64_t N, int64_t K, int tiles_c, int groups,
            int tiles_per_group) {
    __shared__ __align__(16) float xs[kGemvWarps][MT][kTileCols];
    __shared__ float red[kGemvWarps][MT][32];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t rg = blockIdx.x;                 // global 32-row group
    const int64_t tr = rg / kRgPerTile;
    const int rgi = (int)(rg % kRgPerTile);
    const int64_t row = rg * kRgRows + lane;
    const int64_t m0 = (int64_t)blockIdx.y * MT;

    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.0f;

    for (int tc = wid; tc < tiles_c; tc += kGemvWarps) {
        const int64_t tile = tr * tiles_c + tc;
        const uint4 p = planes[tile * kTileRows + rgi * kRgRows + lane];
        const uint32_t vbase = vptr[tile * kRgPerTile + rgi];
        const float2 a = affine[row * groups + tc / tiles_per_group];
        const int64_t col0 = (int64_t)tc * kTileCols;

        // stage this tile's x slice (MT x 64) as fp32; sx = its row sum (same for every output row)
        float sx[MT];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float v0 = 0.0f, v1 = 0.0f;
            if (m0 + m < M) {
                const T* xr = x + (m0 + m) * ldx + col0;
                if (col0 + lane < K) v0 = to_f32(xr[lane]);
                if (col0 + 32 + lane < K) v1 = to_f32(xr[32 + lane]);
            }
            xs[wid][m][lane] = v0;
            xs[wid][m][32 + lane] = v1;
            float s = v0 + v1;
#pragma unroll
            for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            sx[m] = s;
        }
        __syncwarp();

        // dense part: sum of x over the set sign bits of this row
        float dacc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) dacc[m] = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t s = h ? p.y : p.x;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[wid][m][h * 32 + j4 * 4]);
                    if (s & (1u << (j4 * 4 + 0))) dacc[m] += xv.x;
                    if (s & (1u << (j4 * 4 + 1))) dacc[m] += xv.y;
                    if (s & (1u << (j4 * 4 + 2))) dacc[m] += xv.z;
                    if (s & (1u << (j4 * 4 + 3))) dacc[m] += xv.w;
                }
            }
        }

        // salient part: exact stored values, (v - lo) * x
        float sacc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) sacc[m] = 0.0f;
        uint32_t off = vbase + warp_excl_scan(__popc(p.z) + __popc(p.w), lane);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t mk = h ? p.w : p.z;
            while (mk) {
                const int j = __ffs(mk) - 1;
                mk &= mk - 1;
                const float c = to_f32(vals[off++]) - a.x;
#pragma unroll
                for (int m = 0; m < MT; ++m) sacc[m] = fmaf(c, xs[wid][m][h * 32 + j], sacc[m]);
            }
        }
        const float d = a.y - a.x;
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] += fmaf(a.x, sx[m], fmaf(d, dacc[m], sacc[m]));
    }

    // split-K reduction across the CTA's warps
#pragma unroll
    for (int m = 0; m < MT; ++m) red[wid][m][lane] = acc[m];
    __syncthreads();
    if (wid == 0 && row < N) {
        const float b = bias ? bias[row] : 0.0f;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m0 + m < M) {
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < kGemvWarps; ++w) s += red[w][m][lane];
                y[(m0 + m) * ldy + row] = from_f32<T>(s + b);
            }
        }
    }
}

template <typename T>
static int launch_gemv_t(const Layer& L, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t M, cudaStream_t s) {
    const unsigned gx = (unsigned)(L.n_pad / kRgRows);
    auto go = [&](auto mt_tag) {
        constexpr int MT = decltype(mt_tag)::value;
        const unsigned gy = (unsigned)((M + MT - 1) / MT);
        dim3 grid(gx, gy);
        gemv_kernel<T, MT><<<grid, kGemvWarps * 32, 0, s>>>(L.planes, L.vptr, (const T*)L.vals, L.affine, L.bias, x,
                                                            ldx, y, ldy, M, L.N, L.K, (int)L.tiles_c, (int)L.groups,
                                                            L.tiles_per_group);
    };
    if (M <= 1) go(std::integral_constant<int, 1>{});
    else if (M <= 2) go(std::integral_constant<int, 2>{});
    else if (M <= 4) go(std::integral_constant<int, 4>{});
    else go(std::integral_constant<int, 8>{});
    count_launch();
    return check_cuda(cudaGetLastError(), "gemv launch");
}

int launch_gemv(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    if (M > 65535 * 8) { set_error("gemv path: M=%lld too large", (long long)M); return PBL_ERR_SHAPE; }
    switch (L.dtype) {
        case PBL_F16: return launch_gemv_t<__half>(L, (const __half*)x, ldx, (__half*)y, ldy, M, s);
        case PBL_BF16: return launch_gemv_t<__nv_bfloat16>(L, (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, M, s);
        case PBL_F32: return launch_gemv_t<float>(L, (const float*)x, ldx, (float*)y, ldy, M, s);
    }
    set_error("unsupported dtype %d", L.dtype);
    return PBL_ERR_DTYPE;
}

}  // namespace pbl
