// k-th smallest element of a flat tensor (exact, 1-based like torch.kthvalue) by radix select on order-preserving integer
// keys: the two magnitude thresholds of BinaryXnorExceptOutliersLinear.gen_outlier_mask (reference
// quant/outlier_quantizer.py:58-67 calls torch.kthvalue twice over up to 56 M weights per layer).  8 bits per pass (two
// passes for fp16 / bf16, four for fp32); every pass is one streaming histogram kernel over the elements that still match
// the selected prefix plus a one-warp pick kernel; the running state {prefix, mask, k} lives in device memory, so the whole
// selection is stream-ordered with no host round trip.  One-time work, not on the per-forward path.
#include "pbllm_common.cuh"

namespace pbl {

struct KthState { uint32_t prefix, mask, k_lo, k_hi; };     // k as 64 bit (k_lo | k_hi << 32), 1-based rank among the matching elements

template <typename T> __device__ __forceinline__ uint32_t sel_key(T v);
template <> __device__ __forceinline__ uint32_t sel_key<float>(float v) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;             // NaN sorts last, as in torch
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
template <> __device__ __forceinline__ uint32_t sel_key<__half>(__half v) {
    uint32_t u = __half_as_ushort(v);
    if ((u & 0x7FFFu) > 0x7C00u) return 0xFFFFu;
    return (u & 0x8000u) ? (~u & 0xFFFFu) : (u | 0x8000u);
}
template <> __device__ __forceinline__ uint32_t sel_key<__nv_bfloat16>(__nv_bfloat16 v) {
    uint32_t u = __bfloat16_as_ushort(v);
    if ((u & 0x7FFFu) > 0x7F80u) return 0xFFFFu;
    return (u & 0x8000u) ? (~u & 0xFFFFu) : (u | 0x8000u);
}

template <typename T>
__global__ void __launch_bounds__(512) kth_hist_kernel(const T* __restrict__ x, int64_t n, const KthState* __restrict__ st, int shift,
                                                       uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[256];
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = st->prefix, mask = st->mask;
    constexpr int kVec = 16 / (int)sizeof(T);
    const int64_t nv = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? n / kVec : 0;      // 16-byte loads when aligned
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x) + i);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const uint32_t key = sel_key<T>(e[j]);
            if ((key & mask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
        }
    }
    for (int64_t i = nv * kVec + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t key = sel_key<T>(x[i]);
        if ((key & mask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// one warp: find the bin that holds rank k, descend into it, clear the histogram for the next pass
__global__ void kth_pick_kernel(uint32_t* __restrict__ hist, KthState* __restrict__ st, int shift) {
    const uint32_t lane = threadIdx.x;
    uint64_t k = (uint64_t)st->k_lo | ((uint64_t)st->k_hi << 32);
    uint64_t cum = 0;
    uint32_t chosen = 255;
    bool found = false;
    for (int base = 0; base < 256; base += 32) {
        const uint32_t c = hist[base + lane];
        uint64_t inc = c;                                       // inclusive scan of 32 bins
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        const uint32_t hit = __ballot_sync(0xffffffffu, !found && cum + inc >= k);
        if (hit && !found) {
            const uint32_t l = (uint32_t)__ffs(hit) - 1u;
            const uint64_t before = cum + __shfl_sync(0xffffffffu, inc, l) - __shfl_sync(0xffffffffu, (uint64_t)c, l);
            chosen = (uint32_t)base + l;
            k -= before;
            found = true;
        }
        cum += __shfl_sync(0xffffffffu, inc, 31);
        hist[base + lane] = 0;
    }
    if (lane == 0) {
        st->prefix |= chosen << shift;
        st->mask |= 255u << shift;
        st->k_lo = (uint32_t)k;
        st->k_hi = (uint32_t)(k >> 32);
    }
}

template <typename T> __global__ void kth_emit_kernel(const KthState* __restrict__ st, T* __restrict__ out);
template <> __global__ void kth_emit_kernel<float>(const KthState* __restrict__ st, float* __restrict__ out) {
    const uint32_t key = st->prefix;
    *out = __uint_as_float((key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key);
}
template <> __global__ void kth_emit_kernel<__half>(const KthState* __restrict__ st, __half* __restrict__ out) {
    const uint32_t key = st->prefix;
    *out = __ushort_as_half((unsigned short)((key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu)));
}
template <> __global__ void kth_emit_kernel<__nv_bfloat16>(const KthState* __restrict__ st, __nv_bfloat16* __restrict__ out) {
    const uint32_t key = st->prefix;
    *out = __ushort_as_bfloat16((unsigned short)((key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu)));
}

__global__ void kth_init_kernel(KthState* st, uint32_t* hist, uint64_t k) {
    if (threadIdx.x == 0) { st->prefix = 0; st->mask = 0; st->k_lo = (uint32_t)k; st->k_hi = (uint32_t)(k >> 32); }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
}

template <typename T>
static int kth_impl(const T* x, int64_t n, int64_t k, T* out, void* workspace, cudaStream_t s) {
    KthState* st = reinterpret_cast<KthState*>(workspace);
    uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + 64);
    kth_init_kernel<<<1, 256, 0, s>>>(st, hist, (uint64_t)k);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + 512 * 16 - 1) / (512 * 16);
    const unsigned grid = (unsigned)(want < (int64_t)sms * 4 ? (want > 0 ? want : 1) : (int64_t)sms * 4);
    const int bits = (int)sizeof(T) * 8;
    for (int shift = bits - 8; shift >= 0; shift -= 8) {
        kth_hist_kernel<T><<<grid, 512, 0, s>>>(x, n, st, shift, hist);
        kth_pick_kernel<<<1, 32, 0, s>>>(hist, st, shift);
    }
    kth_emit_kernel<T><<<1, 1, 0, s>>>(st, out);
    count_launch(2 + 2 * (bits / 8));
    return check_cuda(cudaGetLastError(), "kth_value launch");
}

size_t kth_workspace_bytes() { return 64 + 256 * sizeof(uint32_t); }

int launch_kth_value(const void* x, int64_t n, int64_t k, int dtype, void* out, void* workspace, cudaStream_t s) {
    switch (dtype) {
        case PBL_F16: return kth_impl<__half>((const __half*)x, n, k, (__half*)out, workspace, s);
        case PBL_BF16: return kth_impl<__nv_bfloat16>((const __nv_bfloat16*)x, n, k, (__nv_bfloat16*)out, workspace, s);
        case PBL_F32: return kth_impl<float>((const float*)x, n, k, (float*)out, workspace, s);
        default: set_error("pbl_kth_value: bad dtype %d", dtype); return PBL_ERR_DTYPE;
    }
}

}  // namespace pbl
