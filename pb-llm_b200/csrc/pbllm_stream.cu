// Packing into / unpacking from the block-stream layout of 16-bit layers (pbllm_stream.cuh): one-time kernels that
// replace the reference's per-forward re-binarisation (quant/quantizer.py:183-188, quant/outlier_quantizer.py:94-98).
// Input is the dense fake-quant weight w_sim the reference feeds to F.linear plus the optional low mask and the
// {lo,hi} table (pack_affine_kernel); output is fsign / eptr / ent / exc.  unpack reproduces w_sim bit-exactly; the same
// kernel writes the dense scratch of the two-phase prefill path.
#include "pbllm_stream.cuh"

namespace pbl {

template <typename T> __device__ __forceinline__ uint32_t bits_of(T v);
template <> __device__ __forceinline__ uint32_t bits_of<__half>(__half v) { return __half_as_ushort(v); }
template <> __device__ __forceinline__ uint32_t bits_of<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat16_as_ushort(v); }
template <typename T> __device__ __forceinline__ T of_bits(uint32_t b);
template <> __device__ __forceinline__ __half of_bits<__half>(uint32_t b) { return __ushort_as_half((unsigned short)b); }
template <> __device__ __forceinline__ __nv_bfloat16 of_bits<__nv_bfloat16>(uint32_t b) { return __ushort_as_bfloat16((unsigned short)b); }

// The salient weight v of a (row, group) with levels {lo, hi} in the +-1 units of the decode kernel:
//   tau = fl16((v - mid) / half),  mid = (lo+hi)/2, half = (hi-lo)/2 in fp32      (pack)
//   v'  = fl16(mid + half * tau)                                                   (unpack; one fma, one rounding)
// and k = ord(v) - ord(v') closes the gap in ulps.  Pack and unpack share these functions, and pack checks the round trip.
template <typename T> __device__ __forceinline__ uint32_t tau_of(uint32_t v, float lo, float hi) {
    const float mid = 0.5f * (lo + hi), half = 0.5f * (hi - lo);
    if (half == 0.f) return 0u;
    return bits_of<T>(from_f32<T>(__fdiv_rn(__fsub_rn(to_f32(of_bits<T>(v)), mid), half)));
}
template <typename T> __device__ __forceinline__ uint32_t value_of(uint32_t tau, float lo, float hi) {
    const float mid = 0.5f * (lo + hi), half = 0.5f * (hi - lo);
    return bits_of<T>(from_f32<T>(__fmaf_rn(half, to_f32(of_bits<T>(tau)), mid)));
}

// One salient value: its entry payload {tau, k}; k = -8 marks an exception (|k| > 7: exact value in the exception list).
template <typename T>
__device__ __forceinline__ void encode_salient(uint32_t v, float lo, float hi, uint32_t& tau16, int& k) {
    tau16 = tau_of<T>(v, lo, hi);
    k = st::ord16(v) - st::ord16(value_of<T>(tau16, lo, hi));
    if (k > st::kMaxK || k < -st::kMaxK) k = -8;
}

// ---- pass 0: (row, group)s with a single level (lo == hi) that also hold salient weights get the pair {mid-1, mid+1}:
// half = 0 cannot carry tau, and with the new pair no element equals a level any more, so every position of the
// (row, group) becomes an entry with tau = v - mid (exact for 16-bit v near mid, k-corrected otherwise).
template <typename T>
__global__ void stream_fix_affine_kernel(const T* __restrict__ w, int64_t ldw, const uint8_t* __restrict__ low_mask, int64_t N,
                                         int64_t K, int64_t gs, int64_t groups, float2* __restrict__ affine) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= N * groups) return;
    const int64_t row = item / groups, g = item % groups;
    const float2 a = affine[row * groups + g];
    if (a.x != a.y) return;                                  // warp-uniform
    const int64_t c0 = g * gs, c1 = min(K, c0 + gs);
    bool any = false;
    for (int64_t c = c0 + lane; c < c1; c += 32) {
        const bool low = (low_mask == nullptr) || low_mask[row * K + c];
        any |= !(low && to_f32(w[row * ldw + c]) == a.x);
    }
    if (__any_sync(0xffffffffu, any) && lane == 0) affine[row * groups + g] = make_float2(a.x - 1.f, a.x + 1.f);
}

// Classification sweep of one 32x64 block, lanes = columns (coalesced): after it lane r holds row r's words.
//   lowb  bit c = 1 when (r, c) is binarized at the LOW level        sal  bit c = 1 when (r, c) is salient
// An element is binarized iff (low_mask == NULL || low_mask[r][c]) && (w == lo || w == hi)  (as in the plane layout).
template <typename T>
__device__ __forceinline__ void classify_block(const T* __restrict__ w, int64_t ldw, const uint8_t* __restrict__ low_mask,
                                               const float2* __restrict__ affine, int64_t N, int64_t K, int64_t groups, int64_t g,
                                               int64_t row0, int64_t col0, uint32_t lane, uint32_t (&lowb)[2], uint32_t (&sal)[2],
                                               float2& my_aff) {
    lowb[0] = lowb[1] = sal[0] = sal[1] = 0u;
    my_aff = make_float2(0.f, 0.f);
    for (int r = 0; r < kRgRows; ++r) {
        const int64_t row = row0 + r;
        uint32_t lb[2] = {0, 0}, sb[2] = {0, 0};
        float2 a = make_float2(0.f, 0.f);
        if (row < N) {                                           // warp-uniform
            a = affine[row * groups + g];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t c = col0 + h * 32 + lane;
                bool is_low = false, is_sal = false;
                if (c < K) {
                    const float v = to_f32(w[row * ldw + c]);
                    const bool low = (low_mask == nullptr) || low_mask[row * K + c];
                    const bool bin = low && (v == a.x || v == a.y);
                    is_sal = !bin;
                    is_low = bin && (v == a.x) && (a.y != a.x);
                }
                lb[h] = __ballot_sync(0xffffffffu, is_low);
                sb[h] = __ballot_sync(0xffffffffu, is_sal);
            }
        }
        if (lane == (uint32_t)r) { lowb[0] = lb[0]; lowb[1] = lb[1]; sal[0] = sb[0]; sal[1] = sb[1]; my_aff = a; }
    }
}

// ---- pass 1: entry units per block (scanned in place afterwards), exception count, "some level pair is asymmetric" flag
// stats: [0] = exceptions, [1] = 1 when any (row, group) has lo != -hi (the decode kernel then needs the mid*sum(x) term)
template <typename T>
__global__ void __launch_bounds__(128) stream_count_kernel(const T* __restrict__ w, int64_t ldw, const uint8_t* __restrict__ low_mask,
                                                           const float2* __restrict__ affine, int64_t N, int64_t K, int tiles_c,
                                                           int64_t groups, int tiles_per_group, uint32_t nblocks,
                                                           uint32_t* __restrict__ eptr, uint32_t* __restrict__ stats) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t blk = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (blk >= nblocks) return;
    const uint32_t rg = blk / (uint32_t)tiles_c, kb = blk - rg * (uint32_t)tiles_c;
    const int64_t row0 = (int64_t)rg * kRgRows, col0 = (int64_t)kb * kTileCols;
    uint32_t lowb[2], sal[2];
    float2 a;
    classify_block<T>(w, ldw, low_mask, affine, N, K, groups, kb / tiles_per_group, row0, col0, lane, lowb, sal, a);
    uint32_t cnt = (uint32_t)(__popc(sal[0]) + __popc(sal[1])), nexc = 0;
    const int64_t row = row0 + lane;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        uint32_t m = sal[h];
        while (m) {
            const uint32_t j = (uint32_t)__ffs(m) - 1u;
            m &= m - 1u;
            uint32_t tau16;
            int k;
            encode_salient<T>(bits_of<T>(w[row * ldw + col0 + h * 32 + j]), a.x, a.y, tau16, k);
            nexc += (k == -8);
        }
    }
    const bool asym = (row < N) && (a.x != -a.y);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        nexc += __shfl_xor_sync(0xffffffffu, nexc, d);
    }
    const bool any_asym = __any_sync(0xffffffffu, asym);
    if (lane == 0) {
        eptr[blk] = (cnt + 3u) / 4u;
        if (nexc) atomicAdd(stats + 0, nexc);
        if (any_asym) atomicOr(stats + 1, 1u);
    }
}

// ---- pass 2: fragment-ordered sign words, entries (ranked for conflict-free patch stores), exceptions ---------------------
// Entry order inside a block (chosen for the decode kernel's patch stores): store j of register set a writes the entries at
// slots 4*(a*h1 + lane) + j, lane = 0..31 -- a "group" of up to 32 entries that should fall in 32 different shared-memory
// banks.  Entries are ranked by (bank, row, column) and rank k goes to group k % 8, position k / 8: the <= 8 entries of
// one bank land in 8 different groups.  (A row touches each bank at most twice: two columns per 32-bit word.)  Blocks with
// more than 256 entries keep ranks >= 256 in rank order behind the first 64 units.
template <typename T>
__global__ void __launch_bounds__(128) stream_fill_kernel(const T* __restrict__ w, int64_t ldw, const uint8_t* __restrict__ low_mask,
                                                          const float2* __restrict__ affine, int64_t N, int64_t K, int tiles_c,
                                                          int64_t groups, int tiles_per_group, uint32_t nblocks,
                                                          const uint32_t* __restrict__ eptr, uint2* __restrict__ fsign,
                                                          uint32_t* __restrict__ ent, uint32_t* __restrict__ exc, uint32_t exc_cap,
                                                          uint32_t* __restrict__ exc_cursor) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t blk = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (blk >= nblocks) return;
    const uint32_t rg = blk / (uint32_t)tiles_c, kb = blk - rg * (uint32_t)tiles_c;
    const int64_t row0 = (int64_t)rg * kRgRows, col0 = (int64_t)kb * kTileCols;
    uint32_t lowb[2], sal[2];
    float2 a;
    classify_block<T>(w, ldw, low_mask, affine, N, K, groups, kb / tiles_per_group, row0, col0, lane, lowb, sal, a);
    const int64_t row = row0 + lane;

    // my row's entries (the sign bit of a salient position stays 0: the kernel XORs the sign bits into the tile)
    uint32_t my_e[64];
    uint32_t m1 = 0, m2 = 0, ne = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        uint32_t m = sal[h];
        while (m) {
            const uint32_t j = (uint32_t)__ffs(m) - 1u;
            m &= m - 1u;
            const uint32_t col = 32u * h + j;
            const uint32_t v = bits_of<T>(w[row * ldw + col0 + col]);
            uint32_t tau16;
            int k;
            encode_salient<T>(v, a.x, a.y, tau16, k);
            const uint32_t slot = st::tile_slot(lane, col);
            my_e[ne++] = st::make_entry(slot, k, tau16);
            if (k == -8) {
                const uint32_t at = atomicAdd(exc_cursor, 1u);
                if (at < exc_cap) { exc[2 * at] = blk; exc[2 * at + 1] = (slot << 16) | v; }
            }
            const uint32_t bit = 1u << ((slot >> 1) & 31u);          // shared-memory bank of the entry's 16-bit slot
            m2 |= m1 & bit;
            m1 |= bit;
        }
    }

    // fragment-ordered sign words: lane (g, t) collects rows g + 8j, columns 16t .. 16t+15
    {
        const uint32_t g = lane >> 2, t = lane & 3u;
        uint32_t out[2] = {0u, 0u};
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t w0 = __shfl_sync(0xffffffffu, lowb[0], g + 8u * j), w1 = __shfl_sync(0xffffffffu, lowb[1], g + 8u * j);
            const uint32_t bits = ((t & 2u) ? w1 : w0) >> (16u * (t & 1u));
#pragma unroll
            for (uint32_t o = 0; o < 16; ++o) {
                const uint32_t q = o >> 2, hi2 = (o >> 1) & 1u, e = o & 1u;
                const uint32_t pos = (15u - (4u * q + (j & 1u) + 2u * hi2)) + 16u * e;
                out[j >> 1] |= ((bits >> o) & 1u) << pos;
            }
        }
        fsign[(size_t)blk * kRgRows + lane] = make_uint2(out[0], out[1]);
    }

    uint32_t cnt = ne;
#pragma unroll
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if (cnt == 0) return;                                   // warp-uniform
    uint32_t* dst = ent + (size_t)eptr[blk] * 4u;
    const uint32_t n4 = (cnt + 3u) / 4u, n1 = min(n4, 64u), h1 = (n1 + 1u) >> 1;

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
        const uint32_t k = start[(e >> 22) & 31u]++;          // bank = (slot >> 1) & 31, slot = e >> 21
        uint32_t s = k;
        if (k < 256u) {
            const uint32_t gq = k & 7u, idx = k >> 3;
            s = 4u * ((gq >> 2) * h1 + idx) + (gq & 3u);
        }
        dst[s] = e;
    }
}

// ---- unpack / expand: dense w_sim [rows][ldw] back from the stream -----------------------------------------------------------
// One warp per block.  The 32 x 64 block is rebuilt in a row-major shared-memory tile (128-byte rows, 16-byte chunks
// XOR-swizzled by row & 7): lane (g, t) turns its fragment-ordered sign words into the levels of its four rows x sixteen
// columns with the decode kernel's shift trick -- ((w << rho) & 0x80008000) marks the two columns of a 32-bit word -- and
// writes them as two 16-byte stores per row; the salient entries are reconstructed over them (value_of + k), and the tile
// leaves as coalesced 16-byte stores (scalar stores for ragged / unaligned destinations).  Entries marked as exceptions
// are left at their level here and overwritten by stream_exc_kernel (same stream, launched right after).
template <typename T>
__global__ void __launch_bounds__(128) stream_unpack_kernel(const uint2* __restrict__ fsign, const uint32_t* __restrict__ eptr,
                                                            const uint32_t* __restrict__ ent, const float2* __restrict__ affine,
                                                            int64_t n_rows, int64_t n_cols, int tiles_c, int64_t groups,
                                                            int tiles_per_group, uint32_t nblocks, T* __restrict__ out, int64_t ldw) {
    __shared__ __align__(16) uint8_t tiles[4][kRgRows * kTileCols * 2];
    __shared__ float2 affs[4][kRgRows];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t blk = blockIdx.x * 4u + wid;
    if (blk >= nblocks) return;
    uint8_t* tile = tiles[wid];
    const uint32_t rg = blk / (uint32_t)tiles_c, kb = blk - rg * (uint32_t)tiles_c;
    const uint32_t g = lane >> 2, t = lane & 3u;
    const uint2 sg = fsign[(size_t)blk * kRgRows + lane];
    affs[wid][lane] = affine[((int64_t)rg * kRgRows + lane) * groups + kb / tiles_per_group];
    __syncwarp();
    auto phys = [](uint32_t r, uint32_t c) { return r * 128u + ((((c >> 3) ^ r) & 7u) << 4) + (c & 7u) * 2u; };   // byte offset of (r, c)
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t r = g + 8u * j;
        const float2 a = affs[wid][r];
        const uint32_t lo = bits_of<T>(from_f32<T>(a.x)), hi = bits_of<T>(from_f32<T>(a.y));
        const uint32_t HH = hi * 0x10001u, DD = (lo ^ hi) * 0x10001u;
        const uint32_t wd = (j >> 1) ? sg.y : sg.x;
        uint32_t v[8];
#pragma unroll
        for (uint32_t w = 0; w < 8; ++w) {                    // word w = columns 16t + 2w, 2w+1: fragment register rho of the decode kernel
            const uint32_t rho = 4u * (w >> 1) + (j & 1u) + 2u * (w & 1u);
            const uint32_t m = ((wd << rho) >> 15) & 0x00010001u;          // bit 1 = LOW level
            v[w] = HH ^ (DD & (m * 0xFFFFu));
        }
        *reinterpret_cast<uint4*>(tile + phys(r, 16u * t)) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(tile + phys(r, 16u * t + 8u)) = make_uint4(v[4], v[5], v[6], v[7]);
    }
    __syncwarp();
    // salient values over the levels (padding copies of an entry rewrite the same value: no hazard)
    const uint32_t e0 = eptr[blk] * 4u, e1 = eptr[blk + 1] * 4u;
    for (uint32_t i = e0 + lane; i < e1; i += 32u) {
        const uint32_t e = ent[i];
        const int k = st::entry_k(e);
        if (k == -8) continue;
        uint32_t r, c;
        st::slot_pos(st::entry_slot(e), r, c);
        const float2 a = affs[wid][r];
        *reinterpret_cast<uint16_t*>(tile + phys(r, c)) = (uint16_t)st::unord16(st::ord16(value_of<T>(e & 0xFFFFu, a.x, a.y)) + k);
    }
    __syncwarp();
    const int64_t row0 = (int64_t)rg * kRgRows, col0 = (int64_t)kb * kTileCols;
    const bool fast = row0 + kRgRows <= n_rows && col0 + kTileCols <= n_cols && (ldw & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (fast) {                                               // 4 rows x 8 chunks of 16 bytes per pass
        const uint32_t rl = lane >> 3, ch = lane & 7u;
#pragma unroll
        for (uint32_t r4 = 0; r4 < (uint32_t)kRgRows; r4 += 4) {
            const uint32_t r = r4 + rl;
            const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128u + (((ch ^ r) & 7u) << 4));
            *reinterpret_cast<uint4*>(out + (row0 + r) * ldw + col0 + ch * 8u) = v;
        }
    } else {
        for (uint32_t r = 0; r < (uint32_t)kRgRows; ++r) {
            const int64_t row = row0 + r;
            if (row >= n_rows) break;
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t c = h * 32u + lane;
                if (col0 + c < n_cols) out[row * ldw + col0 + c] = of_bits<T>(*reinterpret_cast<const uint16_t*>(tile + phys(r, c)));
            }
        }
    }
    // Launched early (programmatic dependent launch, see launch_stream_unpack): the grid touches nothing an earlier kernel
    // produces, but it must not COMPLETE before the kernels ahead of it have, or the kernel after it would be released
    // too soon.  One thread waiting is enough to hold the grid open; without the launch attribute this is a no-op.
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename T>
__global__ void stream_exc_kernel(const uint32_t* __restrict__ exc, uint32_t n_exc, int tiles_c, int64_t n_rows, int64_t n_cols,
                                  T* __restrict__ out, int64_t ldw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_exc) return;
    const uint32_t blk = exc[2 * i], sv = exc[2 * i + 1];
    uint32_t r, c;
    st::slot_pos(sv >> 16, r, c);
    const uint32_t rg = blk / (uint32_t)tiles_c, kb = blk - rg * (uint32_t)tiles_c;
    const int64_t row = (int64_t)rg * kRgRows + r, col = (int64_t)kb * kTileCols + c;
    if (row < n_rows && col < n_cols) out[row * ldw + col] = of_bits<T>(sv & 0xFFFFu);
}

void launch_scan_counts(uint32_t* v, int64_t n, cudaStream_t s);   // pbllm_pack.cu

#define PBL_DISPATCH_16(dtype, ...)                                                                     \
    switch (dtype) {                                                                                    \
        case PBL_F16: { using T = __half; __VA_ARGS__; break; }                                         \
        case PBL_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }                                 \
        default: set_error("the block-stream layout holds fp16 / bf16 layers (dtype %d)", dtype); return PBL_ERR_DTYPE; \
    }

// NOTE: rewrites `affine` in place for single-level (row, group)s that hold salient weights (stream_fix_affine_kernel)
int launch_stream_count(const void* w, int64_t ldw, const uint8_t* low_mask, float2* affine, int64_t N, int64_t K, int dtype,
                        const pbl_sizes& sz, int tiles_per_group, uint32_t* eptr, uint32_t* stats, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)(sz.tiles_r * kRgPerTile * sz.tiles_c);
    int rc = check_cuda(cudaMemsetAsync(stats, 0, 4 * sizeof(uint32_t), s), "memset(stream stats)");
    if (rc) return rc;
    {
        const int64_t items = N * sz.groups, gs = (int64_t)tiles_per_group * kTileCols;
        PBL_DISPATCH_16(dtype, (stream_fix_affine_kernel<T><<<(unsigned)((items + 7) / 8), 256, 0, s>>>((const T*)w, ldw, low_mask, N, K,
                                                                                                     sz.groups == 1 ? K : gs, sz.groups, affine)));
        rc = check_cuda(cudaGetLastError(), "stream_fix_affine launch");
        if (rc) return rc;
        count_launch();
    }
    PBL_DISPATCH_16(dtype, (stream_count_kernel<T><<<(nblocks + 3u) / 4u, 128, 0, s>>>((const T*)w, ldw, low_mask, affine, N, K, (int)sz.tiles_c,
                                                                                       sz.groups, tiles_per_group, nblocks, eptr, stats)));
    rc = check_cuda(cudaGetLastError(), "stream_count launch");
    if (rc) return rc;
    launch_scan_counts(eptr, nblocks, s);
    count_launch(2);
    return check_cuda(cudaGetLastError(), "stream scan launch");
}

int launch_stream_fill(const void* w, int64_t ldw, const uint8_t* low_mask, const float2* affine, int64_t N, int64_t K, int dtype,
                       const pbl_sizes& sz, int tiles_per_group, const uint32_t* eptr, uint2* fsign, uint32_t* ent, uint32_t* exc,
                       uint32_t exc_cap, uint32_t* stats, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)(sz.tiles_r * kRgPerTile * sz.tiles_c);
    int rc = check_cuda(cudaMemsetAsync(stats + 2, 0, sizeof(uint32_t), s), "memset(exception cursor)");
    if (rc) return rc;
    PBL_DISPATCH_16(dtype, (stream_fill_kernel<T><<<(nblocks + 3u) / 4u, 128, 0, s>>>((const T*)w, ldw, low_mask, affine, N, K, (int)sz.tiles_c,
                                                                                      sz.groups, tiles_per_group, nblocks, eptr, fsign, ent,
                                                                                      exc, exc_cap, stats + 2)));
    count_launch();
    return check_cuda(cudaGetLastError(), "stream_fill launch");
}

// dense [n_rows][ldw] <- stream; n_rows / n_cols bound the rows / columns written (N, K for unpack; n_pad, k_pad for the
// prefill scratch)
// early: launch with the programmatic-stream-serialization attribute, so that the expansion may run beside the kernel ahead
// of it in the stream once that kernel has executed griddepcontrol.launch_dependents (gemm_tt_kernel does, first thing).
// Only for destinations no kernel in flight can be reading (the caller's double-buffered scratch).
int launch_stream_unpack(const Layer& L, void* out, int64_t ldw, int64_t n_rows, int64_t n_cols, cudaStream_t s, bool early) {
    const uint32_t nblocks = (uint32_t)(L.tiles_r * kRgPerTile * L.tiles_c);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((nblocks + 3u) / 4u);
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = early ? 1 : 0;
    cudaError_t le = cudaSuccess;
    PBL_DISPATCH_16(L.dtype, (le = cudaLaunchKernelEx(&cfg, stream_unpack_kernel<T>, (const uint2*)L.fsign, (const uint32_t*)L.eptr,
                                                      (const uint32_t*)L.ent, (const float2*)L.affine, n_rows, n_cols, (int)L.tiles_c,
                                                      (int64_t)L.groups, (int)L.tiles_per_group, nblocks, (T*)out, ldw)));
    int rc = check_cuda(le, "stream_unpack launch");
    if (rc) return rc;
    count_launch();
    if (L.n_exc) {
        PBL_DISPATCH_16(L.dtype, (stream_exc_kernel<T><<<(unsigned)((L.n_exc + 255) / 256), 256, 0, s>>>(L.exc, (uint32_t)L.n_exc, (int)L.tiles_c,
                                                                                                       n_rows, n_cols, (T*)out, ldw)));
        count_launch();
        rc = check_cuda(cudaGetLastError(), "stream_exc launch");
    }
    return rc;
}

}  // namespace pbl
