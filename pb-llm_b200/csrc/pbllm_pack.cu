// One-time packing of the dense fake-quant weight w_sim (the tensor the reference feeds to
// F.linear: quant/quantizer.py:86,193; quant/outlier_quantizer.py:98,105; gptq_pb/gptq.py:180-184)
// into the HBM layout the forward kernels stream:
//   planes  uint4 [tiles_r][tiles_c][128 rows] = {sign bits cols 0-31, 32-63, salient bits 0-31, 32-63}
//   vptr    u32   [tiles_r*tiles_c*4 + 1]       value offset of each (tile, 32-row group)
//   vals    T     salient values, (tile, row-group, row, column) order -- exact copies of w_sim
//   affine  float2 [n_pad][groups] = {lo, hi}   the two binarized levels of each (row, group)
// This replaces the reference's per-forward re-binarisation passes (mean/abs/sign/where over
// the whole [N,K] fp32 weight on every call) with a single pass at load time.
#include "pbllm_common.cuh"

namespace pbl {

// ---- affine: {min,max} of the binarized positions of each (row, group) ------------------------
template <typename T>
__global__ void pack_affine_kernel(const T* __restrict__ w, int64_t ldw, const uint8_t* __restrict__ low_mask,
                                   int64_t N, int64_t K, int64_t gs, int64_t groups, int64_t n_pad,
                                   float2* __restrict__ affine) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_pad * groups) return;
    const int64_t row = item / groups, g = item % groups;
    float lo = INFINITY, hi = -INFINITY;
    if (row < N) {
        const int64_t c0 = g * gs, c1 = min(K, c0 + gs);
        for (int64_t c = c0 + lane; c < c1; c += 32) {
            if (low_mask == nullptr || low_mask[row * K + c]) {
                float v = to_f32(w[row * ldw + c]);
                lo = fminf(lo, v);
                hi = fmaxf(hi, v);
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if (lane == 0) {
        if (!(lo <= hi)) lo = hi = 0.0f;  // no binarized position in this group (or padding row)
        affine[row * groups + g] = make_float2(lo, hi);
    }
}

// ---- planes + per-(tile,row-group) salient counts ---------------------------------------------
// Block = one 128x64 tile, warp = one 32-row group; lanes sweep columns so global reads are
// coalesced, __ballot_sync assembles the row words.
template <typename T>
__global__ void __launch_bounds__(128) pack_planes_kernel(const T* __restrict__ w, int64_t ldw,
                                                          const uint8_t* __restrict__ low_mask,
                                                          const float2* __restrict__ affine, int64_t N, int64_t K,
                                                          int tiles_c, int64_t groups, int tiles_per_group,
                                                          uint4* __restrict__ planes, uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31, rgi = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t tr = tile / tiles_c, tc = tile % tiles_c;
    const int64_t g = tc / tiles_per_group;
    const int64_t row0 = tr * kTileRows + rgi * kRgRows, col0 = tc * kTileCols;
    uint4 mine = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < kRgRows; ++r) {
        const int64_t row = row0 + r;
        uint32_t bit[2] = {0, 0}, sal[2] = {0, 0};
        if (row < N) {  // warp-uniform
            const float2 a = affine[row * groups + g];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t c = col0 + h * 32 + lane;
                bool is_bit = false, is_sal = false;
                if (c < K) {
                    const float v = to_f32(w[row * ldw + c]);
                    const bool low = (low_mask == nullptr) || low_mask[row * K + c];
                    const bool bin = low && (v == a.x || v == a.y);
                    is_sal = !bin;
                    is_bit = bin && (v == a.y) && (a.y != a.x);
                }
                bit[h] = __ballot_sync(0xffffffffu, is_bit);
                sal[h] = __ballot_sync(0xffffffffu, is_sal);
            }
        }
        if (lane == r) mine = make_uint4(bit[0], bit[1], sal[0], sal[1]);
    }
    planes[tile * kTileRows + rgi * kRgRows + lane] = mine;
    uint32_t cnt = __popc(mine.z) + __popc(mine.w);
#pragma unroll
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if (lane == 0) counts[tile * kRgPerTile + rgi] = cnt;
}

// In-place exclusive scan of counts[0..n) (n <= a few 100k), counts[n] = total. One block.
__global__ void __launch_bounds__(1024) scan_counts_kernel(uint32_t* __restrict__ v, int64_t n) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t per = (n + 1023) / 1024;
    const int64_t b = (int64_t)tid * per, e = min(n, b + per);
    uint32_t local = 0;
    for (int64_t i = b; i < e; ++i) local += v[i];
    uint32_t ex = warp_excl_scan(local, lane);
    if (lane == 31) warp_sums[wid] = ex + local;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = warp_sums[lane];
        uint32_t es = warp_excl_scan(s, lane);
        warp_sums[lane] = es;
        if (lane == 31) carry = es + s;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + ex;
    for (int64_t i = b; i < e; ++i) {
        uint32_t c = v[i];
        v[i] = run;
        run += c;
    }
    if (tid == 0) v[n] = carry;
}

void launch_scan_counts(uint32_t* v, int64_t n, cudaStream_t s) { scan_counts_kernel<<<1, 1024, 0, s>>>(v, n); }

// ---- salient values -------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) pack_vals_kernel(const T* __restrict__ w, int64_t ldw,
                                                        const uint4* __restrict__ planes,
                                                        const uint32_t* __restrict__ vptr, int tiles_c,
                                                        T* __restrict__ vals) {
    const int lane = threadIdx.x & 31, rgi = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t tr = tile / tiles_c, tc = tile % tiles_c;
    const uint4 p = planes[tile * kTileRows + rgi * kRgRows + lane];
    uint32_t off = vptr[tile * kRgPerTile + rgi] + warp_excl_scan(__popc(p.z) + __popc(p.w), lane);
    const int64_t row = tr * kTileRows + rgi * kRgRows + lane, col0 = tc * kTileCols;
    uint32_t m = p.z;
    while (m) {
        int j = __ffs(m) - 1;
        m &= m - 1;
        vals[off++] = w[row * ldw + col0 + j];
    }
    m = p.w;
    while (m) {
        int j = __ffs(m) - 1;
        m &= m - 1;
        vals[off++] = w[row * ldw + col0 + 32 + j];
    }
}

// ---- unpack: dense w_sim back from the packed form (pack invariant, `.weight`, to_regular_linear)
template <typename T>
__global__ void __launch_bounds__(128) unpack_kernel(const uint4* __restrict__ planes,
                                                     const uint32_t* __restrict__ vptr, const T* __restrict__ vals,
                                                     const float2* __restrict__ affine, int64_t N, int64_t K,
                                                     int tiles_c, int64_t groups, int tiles_per_group,
                                                     T* __restrict__ w_out, int64_t ldw) {
    const int lane = threadIdx.x & 31, rgi = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t tr = tile / tiles_c, tc = tile % tiles_c;
    const uint4 p = planes[tile * kTileRows + rgi * kRgRows + lane];
    uint32_t off = vptr[tile * kRgPerTile + rgi] + warp_excl_scan(__popc(p.z) + __popc(p.w), lane);
    const int64_t row = tr * kTileRows + rgi * kRgRows + lane, col0 = tc * kTileCols;
    if (row >= N) return;
    const float2 a = affine[row * groups + tc / tiles_per_group];
    const T lo = from_f32<T>(a.x), hi = from_f32<T>(a.y);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t s = h ? p.y : p.x, m = h ? p.w : p.z;
        for (int j = 0; j < 32; ++j) {
            const int64_t c = col0 + h * 32 + j;
            if (c >= K) break;
            T v;
            if ((m >> j) & 1u) v = vals[off++];
            else v = ((s >> j) & 1u) ? hi : lo;
            w_out[row * ldw + c] = v;
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------
#define PBL_DISPATCH_DTYPE(dtype, ...)                                        \
    switch (dtype) {                                                          \
        case PBL_F16: { using T = __half; __VA_ARGS__; break; }               \
        case PBL_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }       \
        case PBL_F32: { using T = float; __VA_ARGS__; break; }                \
        default: set_error("unsupported dtype %d", dtype); return PBL_ERR_DTYPE; \
    }

int launch_pack_affine(const void* w, int64_t ldw, const uint8_t* low_mask, int64_t N, int64_t K, int64_t gs,
                       int dtype, float2* affine, int64_t n_pad, int64_t groups, cudaStream_t s) {
    const int64_t items = n_pad * groups;
    const int warps = 8;
    const unsigned grid = (unsigned)((items + warps - 1) / warps);
    PBL_DISPATCH_DTYPE(dtype, (pack_affine_kernel<T><<<grid, warps * 32, 0, s>>>((const T*)w, ldw, low_mask, N, K, gs,
                                                                                 groups, n_pad, affine)));
    count_launch();
    return check_cuda(cudaGetLastError(), "pack_affine launch");
}

int launch_pack_planes(const void* w, int64_t ldw, const uint8_t* low_mask, const float2* affine, int64_t N,
                       int64_t K, int64_t gs, int dtype, uint4* planes, uint32_t* vptr, const pbl_sizes& sz,
                       cudaStream_t s) {
    const int64_t tiles = sz.tiles_r * sz.tiles_c;
    const int tpg = (sz.groups == 1) ? (int)sz.tiles_c : (int)(gs / kTileCols);
    PBL_DISPATCH_DTYPE(dtype, (pack_planes_kernel<T><<<(unsigned)tiles, 128, 0, s>>>(
                                  (const T*)w, ldw, low_mask, affine, N, K, (int)sz.tiles_c, sz.groups, tpg, planes, vptr)));
    int rc = check_cuda(cudaGetLastError(), "pack_planes launch");
    if (rc) return rc;
    scan_counts_kernel<<<1, 1024, 0, s>>>(vptr, tiles * kRgPerTile);
    count_launch(2);
    return check_cuda(cudaGetLastError(), "scan_counts launch");
}

int launch_pack_vals(const void* w, int64_t ldw, const uint4* planes, const uint32_t* vptr, int64_t N, int64_t K,
                     int dtype, void* vals, const pbl_sizes& sz, cudaStream_t s) {
    (void)N; (void)K;
    const int64_t tiles = sz.tiles_r * sz.tiles_c;
    PBL_DISPATCH_DTYPE(dtype, (pack_vals_kernel<T><<<(unsigned)tiles, 128, 0, s>>>((const T*)w, ldw, planes, vptr,
                                                                                  (int)sz.tiles_c, (T*)vals)));
    count_launch();
    return check_cuda(cudaGetLastError(), "pack_vals launch");
}

int launch_unpack(const Layer& L, void* w_out, int64_t ldw, cudaStream_t s) {
    const int64_t tiles = L.tiles_r * L.tiles_c;
    PBL_DISPATCH_DTYPE(L.dtype, (unpack_kernel<T><<<(unsigned)tiles, 128, 0, s>>>(
                                    L.planes, L.vptr, (const T*)L.vals, L.affine, L.N, L.K, (int)L.tiles_c, L.groups,
                                    L.tiles_per_group, (T*)w_out, ldw)));
    count_launch();
    return check_cuda(cudaGetLastError(), "unpack launch");
}

}  // namespace pbl
