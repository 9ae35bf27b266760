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

size_t bireal_workspace_bytes(const Layer& L, int64_t M) {
    return (size_t)M * (size_t)L.tiles_c * (sizeof(uint4) + sizeof(int)) + 256;
}

int launch_bireal(const Layer& L, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy, int64_t M, void* workspace,
                  cudaStream_t s) {
    uint4* xb = reinterpret_cast<uint4*>(workspace);
    int* dx = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + (((size_t)M * L.tiles_c * sizeof(uint4) + 255) / 256) * 256);
    const dim3 gb((unsigned)L.tiles_c, (unsigned)M);
    switch (x_dtype) {
        case PBL_F16: bireal_binarize_kernel<__half><<<gb, 64, 0, s>>>((const __half*)x, ldx, L.K, (int)L.tiles_c, xb, dx); break;
        case PBL_BF16: bireal_binarize_kernel<__nv_bfloat16><<<gb, 64, 0, s>>>((const __nv_bfloat16*)x, ldx, L.K, (int)L.tiles_c, xb, dx); break;
        case PBL_F32: bireal_binarize_kernel<float><<<gb, 64, 0, s>>>((const float*)x, ldx, L.K, (int)L.tiles_c, xb, dx); break;
        default: set_error("pbl_bireal_forward: bad x dtype %d", x_dtype); return PBL_ERR_DTYPE;
    }
    int rc = check_cuda(cudaGetLastError(), "bireal binarize launch");
    if (rc) return rc;
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
