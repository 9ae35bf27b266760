/*
 * pbllm.h -- C ABI of libpbllm.so: B200 (sm_100a) partially-binarized linear forward.
 *
 * Drop-in boundary for the ONE hot path of hahnyuan/PB-LLM: the forward of
 *   quant/quantizer.py:75-86      BinaryLinear          (and FdaBinaryLinear  :112-128, same forward)
 *   quant/quantizer.py:172-193    XnorBinaryLinear      (and IrBinaryLinear   :89-109,  same forward)
 *   quant/outlier_quantizer.py:33-123  BinaryXnorExceptOutliersLinear (+ ...Hessian :126-143)
 *   gptq_pb/gptq.py:180-184       the fp16 nn.Linear that GPTQ-PB writes (format a10)
 * all of which end in  F.linear(x, w_sim, bias)  over a re-materialised dense w_sim
 * (quant/quantizer.py:86,193; quant/outlier_quantizer.py:105).  This library evaluates the same
 * y = x . w_sim^T + bias from a packed form of w_sim; unpack(pack(w_sim)) == w_sim bit-exactly.  Two layouts:
 *   fp16 / bf16 layers  "block stream": 1-bit sign plane in MMA-fragment order + one positioned 32-bit entry per salient
 *                       weight + per-(row,group) {lo,hi}  (pbl_stream_*; the ONE resident copy every 16-bit kernel reads)
 *   fp32 layers         "planes": 1-bit sign plane + salient bitmap + packed salient values + {lo,hi}  (pbl_pack_*)
 *
 * Conventions: plain pointers and sizes, no torch types, no exceptions across the ABI.
 * Every entry point returns PBL_OK (0) or a negative pbl_status; pbl_last_error() gives the
 * thread-local message.  Device pointers are borrowed (owned by the caller's allocator);
 * kernels run on the caller's stream with no internal synchronisation (CUDA-graph capturable).
 * There is NO CPU fallback: on a machine without an sm_100 device every compute entry point
 * returns PBL_ERR_NO_DEVICE / PBL_ERR_ARCH.
 */
#ifndef PBLLM_H_
#define PBLLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBL_ABI_VERSION 4

#if defined(__GNUC__)
#define PBL_API __attribute__((visibility("default")))
#else
#define PBL_API
#endif

typedef enum {
    PBL_OK = 0,
    PBL_ERR_NULL = -1,      /* null pointer argument */
    PBL_ERR_DTYPE = -2,     /* dtype not one of pbl_dtype / not supported by this path */
    PBL_ERR_SHAPE = -3,     /* bad N/K/M/groupsize/leading dimension */
    PBL_ERR_NO_DEVICE = -4, /* no CUDA device / driver */
    PBL_ERR_ARCH = -5,      /* device is not sm_100 (B200) */
    PBL_ERR_CUDA = -6,      /* a CUDA runtime call failed; message carries cudaGetErrorString */
    PBL_ERR_ALIGN = -7,     /* pointer not aligned as required (16 B for x, y, packed buffers) */
    PBL_ERR_UNSUPPORTED = -8
} pbl_status;

/* dtype of activations x, outputs y and packed salient values (the reference requires
 * x.dtype == weight.dtype, SURVEY.md 8b "Call"). */
typedef enum { PBL_F16 = 0, PBL_BF16 = 1, PBL_F32 = 2 } pbl_dtype;

/* Packed layout geometry: plane tiles are PBL_TILE_ROWS output rows x PBL_TILE_COLS input
 * columns; rows/cols are padded to whole tiles inside the packed buffers only. */
#define PBL_TILE_ROWS 128
#define PBL_TILE_COLS 64
#define PBL_RG_ROWS 32 /* rows per value row-group (one warp) */

typedef struct {
    int64_t n_pad, k_pad;       /* N, K rounded up to whole tiles */
    int64_t tiles_r, tiles_c;   /* n_pad/128, k_pad/64 */
    int64_t groups;             /* ceil(K / groupsize) */
    size_t planes_bytes;        /* u32 [tiles_r][tiles_c][128][4] = {sign0,sign1,sal0,sal1} per row */
    size_t vptr_bytes;          /* u32 [tiles_r*tiles_c*4 + 1] value offsets per (tile, row-group) */
    size_t affine_bytes;        /* float2 {lo,hi} [n_pad][groups] */
    size_t vals_elem_bytes;     /* sizeof(dtype); vals buffer = (nnz + 8) * vals_elem_bytes */
} pbl_sizes;

/* Host-only. groupsize <= 0 or >= K means one group per row; otherwise it must be a multiple
 * of PBL_TILE_COLS (GPTQ-PB requires groupsize % 128 == 0, gptq_pb/gptq.py:102). */
PBL_API int pbl_pack_sizes(int64_t N, int64_t K, int64_t groupsize, int dtype, pbl_sizes* out);

/* ---- packing (one-time; replaces the per-forward re-binarisation of quant/quantizer.py:183-188
 *      and quant/outlier_quantizer.py:94-98) -------------------------------------------------- */

/* {lo,hi}[row][group] = {min,max} of w_sim over the binarized positions of the group
 * (all positions if low_mask == NULL).  w_sim: device dense [N][ldw] of `dtype`;
 * low_mask: device uint8/bool [N][K], nonzero = binarized (the GPTQ-PB mask-file convention,
 * gptq_pb/gptq.py:92,99; the complement of outlier_mask, quant/outlier_quantizer.py:138). */
PBL_API int pbl_pack_affine(const void* w_sim, int64_t ldw, const uint8_t* low_mask, int64_t N, int64_t K,
                    int64_t groupsize, int dtype, void* affine_out, void* stream);

/* Pass 1: writes the planes and the per-(tile,row-group) salient counts, then scans them in
 * place into offsets; vptr_out[last] = nnz (read it back to size `vals`).  An element is
 * binarized iff (low_mask == NULL || low_mask[i][j]) && (w == lo || w == hi); everything else
 * (salient weights, sign(0) zeros, any third value) is salient and stored exactly. */
PBL_API int pbl_pack_planes(const void* w_sim, int64_t ldw, const uint8_t* low_mask, const void* affine, int64_t N,
                    int64_t K, int64_t groupsize, int dtype, void* planes_out, void* vptr_out, void* stream);

/* Pass 2: gathers the salient values (tile-major, row-group, row, column order). */
PBL_API int pbl_pack_vals(const void* w_sim, int64_t ldw, const void* planes, const void* vptr, int64_t N, int64_t K,
                  int dtype, void* vals_out, void* stream);

/* ---- layer handle ------------------------------------------------------------------------ */

#define PBL_LAYER_HAS_MID 1u /* some (row, group) has lo != -hi: reported by pbl_stream_count */

typedef struct {
    int64_t N, K;        /* out_features, in_features of the replaced nn.Linear */
    int64_t groupsize;   /* as given to pbl_pack_sizes */
    int32_t dtype;       /* pbl_dtype */
    uint32_t flags;      /* PBL_LAYER_* (block-stream layers) */
    const void* affine;  /* device float2 [n_pad][groups] */
    const void* bias;    /* device float32 [N] or NULL (bias added inside, as F.linear does) */
    /* -- planes layout (dtype PBL_F32; NULL for 16-bit layers) -- */
    const void* planes;  /* device, 16 B aligned */
    const void* vptr;    /* device */
    const void* vals;    /* device, dtype, 16 B aligned, >= nnz + 8 elements */
    const void* sign_planes; /* optional, device uint2 [tiles_r][tiles_c][128] = the sign words of `planes` only;
                              * may be given iff nnz == 0 (pure binary layer): halves the bytes pbl_bireal_forward streams */
    /* -- block-stream layout (dtype PBL_F16 / PBL_BF16; NULL for fp32 layers) -- */
    const void* fsign;   /* device uint2 [blocks][32], 16 B aligned */
    const void* eptr;    /* device u32 [blocks + 1] */
    const void* ent;     /* device u32 [4 * eptr[blocks]], 16 B aligned (at least 16 bytes) */
    const void* exc;     /* device u32 [n_exc][2] or NULL */
    int64_t n_exc;
} pbl_layer_desc;

typedef struct pbl_layer pbl_layer; /* opaque; borrows the descriptor's device pointers */

PBL_API int pbl_layer_create(const pbl_layer_desc* desc, pbl_layer** out);
PBL_API void pbl_layer_destroy(pbl_layer* layer);

/* Reconstruct the dense w_sim [N][ldw] (dtype) from the packed form -- the pack invariant
 * check and the backing of the modules' `.weight` / to_regular_linear()
 * (quant/outlier_quantizer.py:108-114). */
PBL_API int pbl_unpack(const pbl_layer* layer, void* w_out, int64_t ldw, void* stream);

/* ---- the hot path ------------------------------------------------------------------------ */

/* y[m][i] = sum_j x[m][j] * w_sim[i][j] + bias[i]   (fp32 accumulation, y rounded to dtype)
 * Replaces F.linear(x, w_sim, bias) at quant/quantizer.py:86,193 and
 * quant/outlier_quantizer.py:105.  x: device [M][ldx], y: device [M][ldy] (row-major, dtype).
 * M = product of the leading dims of the reference's x[..., K].  M == 0 is a no-op.
 * fp16 / bf16 layers: calls of up to PBL_DECODE_MAX_M tokens (default 64) run the decode kernel in passes of 16 tokens;
 * larger calls first expand the weight into a dense scratch of 2*n_pad*k_pad bytes and run the tcgen05 GEMM over it.
 * The scratch is one of two buffers the library keeps per (device, stream) and uses alternately, so that the expansion
 * of a call can run beside the GEMM of the call before it on the same stream (grown with cudaMalloc when a larger layer
 * arrives, kept for the life of the process; PBL_PREFILL_OVERLAP=0 turns this off).  Under stream capture, and for more
 * than 8 streams per device, the scratch is transient: taken from and returned to the device's stream-ordered memory
 * pool on `stream` (cudaMallocAsync / cudaFreeAsync: no synchronisation, CUDA-graph capturable).  fp32 layers run the CUDA-core bit-plane kernel. */
PBL_API int pbl_linear_forward(const pbl_layer* layer, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M,
                       void* stream);

/* Same call with HOST buffers (the reference-facing end-to-end form, used for bench.py's
 * `e2e`): copies x host->device, runs the kernel, copies y device->host on `stream` and
 * synchronises it.  x_host/y_host should be pinned for full PCIe rate.  `workspace` is a
 * device buffer of at least pbl_forward_host_workspace(layer, M) bytes. */
PBL_API size_t pbl_forward_host_workspace(const pbl_layer* layer, int64_t M);
PBL_API int pbl_linear_forward_host(const pbl_layer* layer, const void* x_host, void* y_host, int64_t M, void* workspace,
                            void* stream);

/* ---- block-stream layout of fp16 / bf16 layers (csrc/pbllm_stream.cuh): blocks of 32 rows x 64 columns in
 *      row-group-major order.  It is what the decode kernel streams for calls of a few tokens (the per-token step of
 *      generation, where the reference's F.linear re-reads the whole dense weight: quant/quantizer.py:86,193,
 *      outlier_quantizer.py:105) and what pbl_unpack / the prefill expansion read back.  Packing (one-time):
 *        pbl_pack_affine -> pbl_stream_count (eptr; read back eptr[blocks] = units and stats) -> pbl_stream_fill.
 *      Sizes: fsign = blocks*32*8 bytes, eptr = (blocks+1)*4 bytes, ent = 16 bytes per unit, exc = 8 bytes per exception.
 *      pbl_stream_count rewrites `affine` in place for single-level (row, group)s that hold salient weights ({mid-1, mid+1}:
 *      all their positions become entries).  stats (device u32[4]): [0] exceptions, [1] PBL_LAYER_HAS_MID flag, [2] exceptions written by pbl_stream_fill. ---- */
typedef struct {
    int64_t blocks;      /* (n_pad/32) * tiles_c blocks of 32 rows x 64 columns */
    size_t fsign_bytes;  /* uint2 [blocks][32] fragment-ordered sign words */
    size_t eptr_bytes;   /* u32 [blocks + 1] entry offsets in 16-byte units */
} pbl_stream_sizes;
PBL_API int pbl_stream_layout(int64_t N, int64_t K, int64_t groupsize, int dtype, pbl_stream_sizes* out);
PBL_API int pbl_stream_count(const void* w_sim, int64_t ldw, const uint8_t* low_mask, void* affine, int64_t N, int64_t K,
                             int64_t groupsize, int dtype, void* eptr_out, void* stats_out, void* stream);
PBL_API int pbl_stream_fill(const void* w_sim, int64_t ldw, const uint8_t* low_mask, const void* affine, int64_t N, int64_t K,
                            int64_t groupsize, int dtype, const void* eptr, void* fsign_out, void* ent_out, void* exc_out,
                            int64_t exc_capacity, void* stats, void* stream);
/* Host-only: where position (r, c) of a block lives -- out4 = {owner lane, sign word, sign bit, correction-tile slot}
 * (tests/test_stream_layout.py checks the fragment mapping against mma.sync's documented layout on the CPU). */
PBL_API int pbl_stream_position(int r, int c, uint32_t* out4);

/* pbl_linear_forward with a caller-owned device workspace for the decode kernel's cross-CTA reduction:
 * >= pbl_decode_workspace_bytes(layer, M) bytes, 16 B aligned, ZERO-INITIALISED ONCE by the caller (the kernel
 * leaves it all zero again), used by one stream at a time; it may be shared by all layers of a device.  Calls that do not take the decode kernel ignore it.  Without a workspace (pbl_linear_forward) the
 * decode kernel takes a transient one from the stream-ordered pool and zeroes it on every call. */
PBL_API size_t pbl_decode_workspace_bytes(const pbl_layer* layer, int64_t M);
PBL_API int pbl_linear_forward_ws(const pbl_layer* layer, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- row-sharded execution across the GPUs of a node: rank g holds the packed rows [g*N/G, (g+1)*N/G) of a
 *      linear, x is replicated, and the all-gather of the [M, N/G] slices is FUSED into the decode kernel: its epilogue
 *      stores the slice straight into the y buffer of every rank through peer-mapped pointers (NVLink / NVSwitch), the last
 *      CTA of the grid publishes a new epoch in every rank's flag array, and the next pushed kernel (or pbl_peer_wait)
 *      waits until the flags of all ranks show the previous push.  No NCCL call on the per-token path.
 *      M <= 16 (one decode pass); larger calls gather with the caller's collective library. ---- */
#define PBL_MAX_PEERS 8
typedef struct {
    void* y[PBL_MAX_PEERS];         /* y [M][ldy] of every rank, peer-mapped, already offset to THIS rank's first output column */
    uint32_t* flags[PBL_MAX_PEERS]; /* flag array u32 [PBL_MAX_PEERS] of every rank, peer-mapped, zero-initialised once */
    uint32_t* sync_ctr;             /* local device u32 [2], zero-initialised once; shared by all pushed layers of this rank */
    int32_t n_ranks, rank;
    int32_t wait_prev;              /* nonzero: wait for every rank's previous push before reading x (x was produced by it) */
    int32_t reserved;
} pbl_peer_push;
PBL_API int pbl_linear_forward_push(const pbl_layer* layer, const void* x, int64_t ldx, const pbl_peer_push* push, int64_t ldy,
                                    int64_t M, void* workspace, size_t workspace_bytes, void* stream);
/* Stream-ordered wait until every rank's latest push has landed in this rank's buffers (for consumers that are not
 * pushed kernels, and for the end of a timed step). */
PBL_API int pbl_peer_wait(const pbl_peer_push* push, void* stream);

/* Host-only: the decode kernel's launch plan for an N x K layer, M tokens, on a device with `sms` SMs and
 * `ctas_per_sm` CTAs of 8 warps per SM (a pass handles 8 tokens, or 16 when M > 8).  out8 = {blocks, row groups, grid.x,
 * token passes, q, rem, slots, workspace KiB}:
 * warp g of the grid owns blocks [g*q + min(g, rem), (g+1)*q + min(g+1, rem)) of the row-group-major block order, and a
 * row group's partials need at most `slots` workspace slots (tests/test_decode_plan.py checks both on the CPU). */
PBL_API int pbl_decode_plan(int64_t N, int64_t K, int64_t M, int sms, int ctas_per_sm, uint32_t* out8);

/* Profiling aid for the decode kernel (tools/decode_trace.py): with a device buffer registered, the following decode
 * launches write per-warp %globaltimer stamps (start, before/after the dependency wait, loop end, after the CTA barrier,
 * after the cross-warp reduction, end; SM id) -- launch i at byte offset i * 16*148*8*8*8.  NULL switches it off. */
PBL_API void pbl_decode_set_trace(void* device_buf, size_t bytes);

/* XNOR-popcount forward of BiRealLinear (quant/quantizer.py:151-169): activations are binarized too,
 *   y[m][i] = sum_j sign(x[m][j]) * w_sim[i][j]        (fp32 out, NO bias -- the reference drops it, :168)
 * evaluated as hi*(popc(b&xp)-popc(b&xn)) + lo*(popc(nb&xp)-popc(nb&xn)) over the packed sign plane.
 * `layer` must be packed from alpha_i*sign(W) (all salient values exactly zero).  x: device [M][ldx] of
 * x_dtype (any pbl_dtype, independent of the layer's); y: device float [M][ldy]; workspace: device,
 * 16 B aligned, >= pbl_bireal_workspace(layer, M) bytes (packed activation sign planes; plain scratch).
 * pbl_bireal_forward_ws additionally takes a ZERO-INITIALISED reduction workspace of >= pbl_bireal_fixup_workspace(layer, M)
 * bytes (same contract as pbl_linear_forward_ws's: left zero by every call, one stream at a time, shareable between
 * layers and with the decode kernel); with it (and M <= 64) the stream-K XNOR kernel runs, which balances the layer's
 * blocks over all SMs.  Without it the row-group-per-CTA kernel runs. */
PBL_API size_t pbl_bireal_workspace(const pbl_layer* layer, int64_t M);
PBL_API int pbl_bireal_forward(const pbl_layer* layer, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy,
                               int64_t M, void* workspace, void* stream);
PBL_API size_t pbl_bireal_fixup_workspace(const pbl_layer* layer, int64_t M);
PBL_API int pbl_bireal_forward_ws(const pbl_layer* layer, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy,
                                  int64_t M, void* workspace, void* fixup_workspace, size_t fixup_workspace_bytes, void* stream);

/* ---- magnitude thresholds on the device: the k-th smallest element (k 1-based, exact, NaN last -- the
 *      semantics of torch.kthvalue, which BinaryXnorExceptOutliersLinear.gen_outlier_mask calls twice over the flattened
 *      weight, quant/outlier_quantizer.py:58-67) by radix select; x: device, n elements of `dtype`; out: device, one element;
 *      workspace: device, >= pbl_kth_workspace() bytes, 16 B aligned.  Stream-ordered, no host synchronisation. ---- */
PBL_API size_t pbl_kth_workspace(void);
PBL_API int pbl_kth_value(const void* x, int64_t n, int64_t k, int dtype, void* out, void* workspace, void* stream);

/* ---- GPTQ-PB calibration: the column loop of LowHighGPT.fasterquant (gptq_pb/gptq.py:116-168) for ONE block
 *      of nc <= 128 columns, all rows, as one kernel.  W1 [N][ldw] fp32 holds the block's current weights and receives the
 *      quantised values Q1 (gptq.py:166); err_out [N][lde] receives Err1 (gptq.py:163) for the caller's cross-block update
 *      W[:, col_ed:] -= Err1 @ Hinv[col_st:col_ed, col_ed:] (gptq.py:168); hinv_block = Hinv[col_st:col_ed, col_st:col_ed]
 *      (upper Cholesky factor of the inverse Hessian, row stride ldh); mask1 [N][ldm] nonzero = low (binarized) position
 *      (gptq.py:92,99); per-row parameters: the xnor low quantizer's mean / scale of the block's group (low_quant.py:25-32) and
 *      the 8-bit high quantizer's scale / zero (high_quant.py:29-67); losses [N] (optional) accumulates gptq.py:167. ---- */
PBL_API int pbl_gptq_block(float* W1, int64_t ldw, float* err_out, int64_t lde, const float* hinv_block, int64_t ldh,
                           const uint8_t* mask1, int64_t ldm, const float* low_mean, const float* low_scale,
                           const float* high_scale, const float* high_zero, float maxq, int64_t N, int nc, float* losses,
                           void* stream);

/* Which kernel pbl_linear_forward would launch for this (layer, M): 0 = CUDA-core bit-plane kernel (fp32 layers),
 * 1 = two-phase prefill (expansion + tcgen05 GEMM; fp16 / bf16 layers, M above PBL_DECODE_MAX_M, default 64),
 * 4 = decode kernel.  PBL_FORCE_KERNEL=0|1|4 overrides (tests). */
PBL_API int pbl_select_kernel(const pbl_layer* layer, int64_t M);

/* Which instance of the decode kernel a call with these activations would run (host-only, nothing is launched):
 * 2 = pair kernel (one group per row, K a multiple of 128, M <= 8, x 32-byte aligned with ldx a multiple of 16 elements),
 * 1 = general block kernel, 0 = the decode kernel does not apply (fp32 layer, M out of range).  PBL_DK_PAIR=0 disables 2. */
PBL_API int pbl_decode_variant(const pbl_layer* layer, const void* x, int64_t ldx, int64_t M);

/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
PBL_API int64_t pbl_launch_count(void);

PBL_API const char* pbl_last_error(void);
PBL_API int pbl_abi_version(void);
/* 0 when an sm_100 device is usable, else PBL_ERR_NO_DEVICE / PBL_ERR_ARCH. */
PBL_API int pbl_device_check(void);

#ifdef __cplusplus
}
#endif
#endif /* PBLLM_H_ */
