// Shared declarations for libpbllm.so (sm_100a only). See include/pbllm.h for the ABI and
// DESIGN.md for the packed layout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbllm.h"

namespace pbl {

constexpr int kTileRows = PBL_TILE_ROWS;  // 128 output rows per plane tile
constexpr int kTileCols = PBL_TILE_COLS;  // 64 input columns per plane tile
constexpr int kRgRows = PBL_RG_ROWS;      // 32 rows per value row-group (one warp)
constexpr int kRgPerTile = kTileRows / kRgRows;

struct Layer {  // the opaque pbl_layer
    int64_t N, K, groupsize;
    int dtype;
    int64_t n_pad, k_pad, tiles_r, tiles_c, groups;
    int tiles_per_group;  // groupsize / 64 (tiles_c when one group)
    const uint4* planes;  // [tiles_r][tiles_c][128] {sign0, sign1, sal0, sal1}
    const uint32_t* vptr; // [tiles_r*tiles_c*4 + 1]
    const void* vals;
    const float2* affine; // [n_pad][groups] {lo, hi}
    const float* bias;
    const uint2* sign_planes;  // optional compact sign-only planes (nnz == 0 layers)
    // block-stream layout (fp16 / bf16 layers; pbllm_stream.cuh): fragment-ordered sign words, entry offsets, salient
    // entries {slot, k, correction}, exception list; flags = PBL_LAYER_*
    const uint2* fsign;
    const uint32_t* eptr;
    const uint32_t* ent;
    const uint32_t* exc;
    int64_t n_exc;
    uint32_t flags;
};

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane) {
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= (uint32_t)d) s += t;
    }
    return s - v;
}

// kernels' host launchers (each returns a pbl_status)
int launch_pack_affine(const void* w, int64_t ldw, const uint8_t* low_mask, int64_t N, int64_t K, int64_t gs,
                       int dtype, float2* affine, int64_t n_pad, int64_t groups, cudaStream_t s);
int launch_pack_planes(const void* w, int64_t ldw, const uint8_t* low_mask, const float2* affine, int64_t N,
                       int64_t K, int64_t gs, int dtype, uint4* planes, uint32_t* vptr, const pbl_sizes& sz,
                       cudaStream_t s);
int launch_pack_vals(const void* w, int64_t ldw, const uint4* planes, const uint32_t* vptr, int64_t N, int64_t K,
                     int dtype, void* vals, const pbl_sizes& sz, cudaStream_t s);
int launch_unpack(const Layer& L, void* w_out, int64_t ldw, cudaStream_t s);
int launch_gemv(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s);
int launch_stream_count(const void* w, int64_t ldw, const uint8_t* low_mask, float2* affine, int64_t N, int64_t K, int dtype,
                        const pbl_sizes& sz, int tiles_per_group, uint32_t* eptr, uint32_t* stats, cudaStream_t s);
int launch_stream_fill(const void* w, int64_t ldw, const uint8_t* low_mask, const float2* affine, int64_t N, int64_t K, int dtype,
                       const pbl_sizes& sz, int tiles_per_group, const uint32_t* eptr, uint2* fsign, uint32_t* ent, uint32_t* exc,
                       uint32_t exc_cap, uint32_t* stats, cudaStream_t s);
int launch_stream_unpack(const Layer& L, void* out, int64_t ldw, int64_t n_rows, int64_t n_cols, cudaStream_t s, bool early = false);
int launch_gemm_twophase(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s);
bool gemm_twophase_supported(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M);
bool decode_supported(const Layer& L, int64_t ldx, int64_t M);
int decode_variant(const Layer& L, const void* x, int64_t ldx, int64_t M);
size_t decode_workspace_bytes(const Layer& L, int64_t M);
void decode_set_trace(void* buf, size_t bytes);
void decode_plan(int64_t N, int64_t K, int64_t M, int sms, int ctas_per_sm, uint32_t out[8]);
int launch_decode(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, void* ws, size_t ws_bytes,
                  cudaStream_t s, const pbl_peer_push* push = nullptr);
int launch_peer_wait(const pbl_peer_push& push, cudaStream_t s);
int launch_gptq_block(float* W, int64_t ldw, float* Err, int64_t lde, const float* Hinv, int64_t ldh, const uint8_t* mask, int64_t ldm,
                      const float* lmean, const float* lscale, const float* hscale, const float* hzero, float maxq, int64_t N, int nc,
                      float* losses, cudaStream_t s);
size_t kth_workspace_bytes();
int launch_kth_value(const void* x, int64_t n, int64_t k, int dtype, void* out, void* workspace, cudaStream_t s);
size_t bireal_workspace_bytes(const Layer& L, int64_t M);
size_t bireal_fixup_workspace_bytes(const Layer& L, int64_t M);
int launch_bireal(const Layer& L, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy, int64_t M, void* workspace,
                  void* fixup_ws, size_t fixup_bytes, cudaStream_t s);

}  // namespace pbl
