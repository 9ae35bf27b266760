// C ABI of libpbllm.so (see include/pbllm.h). Host-side glue only: argument checking, the
// opaque layer handle, kernel selection, thread-local error string. No torch types, no
// exceptions across the boundary, no CPU fallback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pbllm_stream.cuh"

namespace pbl {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PBL_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PBL_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int device_check_impl() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); libpbllm has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PBL_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed"); return PBL_ERR_NO_DEVICE; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        set_error("device %d is sm_%d%d; libpbllm is built for sm_100a (B200) only", dev, major, minor);
        return PBL_ERR_ARCH;
    }
    return PBL_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool valid_dtype(int d) { return d == PBL_F16 || d == PBL_BF16 || d == PBL_F32; }
static size_t dtype_size(int d) { return d == PBL_F32 ? 4 : 2; }

static int sizes_impl(int64_t N, int64_t K, int64_t gs, int dtype, pbl_sizes* out) {
    if (!out) { set_error("pbl_pack_sizes: out is NULL"); return PBL_ERR_NULL; }
    if (!valid_dtype(dtype)) { set_error("pbl_pack_sizes: bad dtype %d", dtype); return PBL_ERR_DTYPE; }
    if (N <= 0 || K <= 0) { set_error("pbl_pack_sizes: N=%lld K=%lld must be positive", (long long)N, (long long)K); return PBL_ERR_SHAPE; }
    if (gs <= 0 || gs >= K) gs = K;
    else if (gs % kTileCols != 0) {
        set_error("groupsize %lld must be a multiple of %d (or cover the whole row)", (long long)gs, kTileCols);
        return PBL_ERR_SHAPE;
    }
    pbl_sizes s;
    s.n_pad = (N + kTileRows - 1) / kTileRows * kTileRows;
    s.k_pad = (K + kTileCols - 1) / kTileCols * kTileCols;
    s.tiles_r = s.n_pad / kTileRows;
    s.tiles_c = s.k_pad / kTileCols;
    s.groups = (K + gs - 1) / gs;
    if (s.tiles_r * s.tiles_c * kRgPerTile >= (int64_t)1 << 31 || N * K >= (int64_t)1 << 32) {
        set_error("layer too large for 32-bit value offsets"); return PBL_ERR_SHAPE;
    }
    s.planes_bytes = (size_t)(s.tiles_r * s.tiles_c * kTileRows) * sizeof(uint4);
    s.vptr_bytes = (size_t)(s.tiles_r * s.tiles_c * kRgPerTile + 1) * sizeof(uint32_t);
    s.affine_bytes = (size_t)(s.n_pad * s.groups) * sizeof(float2);
    s.vals_elem_bytes = dtype_size(dtype);
    *out = s;
    return PBL_OK;
}

}  // namespace pbl

using namespace pbl;

extern "C" {

int pbl_abi_version(void) { return PBL_ABI_VERSION; }
const char* pbl_last_error(void) { return g_err; }
int64_t pbl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int pbl_device_check(void) { return device_check_impl(); }

int pbl_pack_sizes(int64_t N, int64_t K, int64_t groupsize, int dtype, pbl_sizes* out) {
    return sizes_impl(N, K, groupsize, dtype, out);
}

int pbl_pack_affine(const void* w_sim, int64_t ldw, const uint8_t* low_mask, int64_t N, int64_t K,
                    int64_t groupsize, int dtype, void* affine_out, void* stream) {
    pbl_sizes sz;
    int rc = sizes_impl(N, K, groupsize, dtype, &sz);
    if (rc) return rc;
    if (!w_sim || !affine_out) { set_error("pbl_pack_affine: null pointer"); return PBL_ERR_NULL; }
    if (ldw < K) { set_error("pbl_pack_affine: ldw %lld < K %lld", (long long)ldw, (long long)K); return PBL_ERR_SHAPE; }
    if ((rc = device_check_impl())) return rc;
    const int64_t gs = (groupsize <= 0 || groupsize >= K) ? K : groupsize;
    return launch_pack_affine(w_sim, ldw, low_mask, N, K, gs, dtype, (float2*)affine_out, sz.n_pad, sz.groups,
                              (cudaStream_t)stream);
}

int pbl_pack_planes(const void* w_sim, int64_t ldw, const uint8_t* low_mask, const void* affine, int64_t N,
                    int64_t K, int64_t groupsize, int dtype, void* planes_out, void* vptr_out, void* stream) {
    pbl_sizes sz;
    int rc = sizes_impl(N, K, groupsize, dtype, &sz);
    if (rc) return rc;
    if (!w_sim || !affine || !planes_out || !vptr_out) { set_error("pbl_pack_planes: null pointer"); return PBL_ERR_NULL; }
    if (ldw < K) { set_error("pbl_pack_planes: ldw %lld < K %lld", (long long)ldw, (long long)K); return PBL_ERR_SHAPE; }
    if (!aligned16(planes_out)) { set_error("pbl_pack_planes: planes_out must be 16 B aligned"); return PBL_ERR_ALIGN; }
    if ((rc = device_check_impl())) return rc;
    const int64_t gs = (groupsize <= 0 || groupsize >= K) ? K : groupsize;
    return launch_pack_planes(w_sim, ldw, low_mask, (const float2*)affine, N, K, gs, dtype, (uint4*)planes_out,
                              (uint32_t*)vptr_out, sz, (cudaStream_t)stream);
}

int pbl_pack_vals(const void* w_sim, int64_t ldw, const void* planes, const void* vptr, int64_t N, int64_t K,
                  int dtype, void* vals_out, void* stream) {
    pbl_sizes sz;
    int rc = sizes_impl(N, K, 0, dtype, &sz);
    if (rc) return rc;
    if (!w_sim || !planes || !vptr || !vals_out) { set_error("pbl_pack_vals: null pointer"); return PBL_ERR_NULL; }
    if (ldw < K) { set_error("pbl_pack_vals: ldw %lld < K %lld", (long long)ldw, (long long)K); return PBL_ERR_SHAPE; }
    if ((rc = device_check_impl())) return rc;
    return launch_pack_vals(w_sim, ldw, (const uint4*)planes, (const uint32_t*)vptr, N, K, dtype, vals_out, sz,
                            (cudaStream_t)stream);
}

int pbl_layer_create(const pbl_layer_desc* d, pbl_layer** out) {
    if (!d || !out) { set_error("pbl_layer_create: null pointer"); return PBL_ERR_NULL; }
    *out = nullptr;
    pbl_sizes sz;
    int rc = sizes_impl(d->N, d->K, d->groupsize, d->dtype, &sz);
    if (rc) return rc;
    if (!d->affine) { set_error("pbl_layer_create: null affine table"); return PBL_ERR_NULL; }
    const bool stream_layout = d->dtype != PBL_F32;
    if (stream_layout) {       // fp16 / bf16: block-stream layout
        if (!d->fsign || !d->eptr || !d->ent) { set_error("pbl_layer_create: fp16 / bf16 layers need fsign / eptr / ent (pbl_stream_*)"); return PBL_ERR_NULL; }
        if (!aligned16(d->fsign) || !aligned16(d->ent)) { set_error("pbl_layer_create: fsign/ent must be 16 B aligned"); return PBL_ERR_ALIGN; }
        if (d->n_exc < 0 || (d->n_exc > 0 && !d->exc)) { set_error("pbl_layer_create: n_exc without an exception list"); return PBL_ERR_NULL; }
    } else {                   // fp32: planes layout
        if (!d->planes || !d->vptr || !d->vals) { set_error("pbl_layer_create: fp32 layers need planes / vptr / vals (pbl_pack_*)"); return PBL_ERR_NULL; }
        if (!aligned16(d->planes) || !aligned16(d->vals)) { set_error("pbl_layer_create: planes/vals must be 16 B aligned"); return PBL_ERR_ALIGN; }
    }
    Layer* L = new (std::nothrow) Layer();
    if (!L) { set_error("out of host memory"); return PBL_ERR_CUDA; }
    L->N = d->N; L->K = d->K;
    L->groupsize = (d->groupsize <= 0 || d->groupsize >= d->K) ? d->K : d->groupsize;
    L->dtype = d->dtype;
    L->n_pad = sz.n_pad; L->k_pad = sz.k_pad; L->tiles_r = sz.tiles_r; L->tiles_c = sz.tiles_c; L->groups = sz.groups;
    L->tiles_per_group = (sz.groups == 1) ? (int)sz.tiles_c : (int)(L->groupsize / kTileCols);
    L->affine = (const float2*)d->affine; L->bias = (const float*)d->bias;
    L->planes = nullptr; L->vptr = nullptr; L->vals = nullptr; L->sign_planes = nullptr;
    L->fsign = nullptr; L->eptr = nullptr; L->ent = nullptr; L->exc = nullptr; L->n_exc = 0; L->flags = d->flags;
    if (stream_layout) {
        L->fsign = (const uint2*)d->fsign; L->eptr = (const uint32_t*)d->eptr; L->ent = (const uint32_t*)d->ent;
        L->exc = (const uint32_t*)d->exc; L->n_exc = d->n_exc;
    } else {
        L->planes = (const uint4*)d->planes; L->vptr = (const uint32_t*)d->vptr; L->vals = d->vals;
        L->sign_planes = (const uint2*)d->sign_planes;
    }
    *out = reinterpret_cast<pbl_layer*>(L);
    return PBL_OK;
}

void pbl_layer_destroy(pbl_layer* layer) { delete reinterpret_cast<Layer*>(layer); }

int pbl_unpack(const pbl_layer* layer, void* w_out, int64_t ldw, void* stream) {
    if (!layer || !w_out) { set_error("pbl_unpack: null pointer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (ldw < L.K) { set_error("pbl_unpack: ldw < K"); return PBL_ERR_SHAPE; }
    int rc = device_check_impl();
    if (rc) return rc;
    if (L.fsign) return launch_stream_unpack(L, w_out, ldw, L.N, L.K, (cudaStream_t)stream);
    return launch_unpack(L, w_out, ldw, (cudaStream_t)stream);
}

static int forced_kernel() {
    const char* e = getenv("PBL_FORCE_KERNEL");
    if (!e || !*e) return -1;
    return atoi(e);
}

static int decode_max_m() {   // calls of up to this many tokens run the decode kernel (in passes of 16)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PBL_DECODE_MAX_M");
        v = (e && *e) ? atoi(e) : 64;
        if (v < 1) v = 1;
    }
    return v;
}

// 0 = CUDA-core bit-plane kernel (fp32 layers), 1 = two-phase prefill (expansion + tcgen05 GEMM), 4 = decode kernel;
// -1 = forced kernel unsupported
static int select_impl(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M) {
    const int f = forced_kernel();
    if (!L.fsign) return (f < 0 || f == 0) ? 0 : -1;                     // fp32 layer: planes layout, CUDA cores
    const bool tc_ok = gemm_twophase_supported(L, x, ldx, y, ldy, M);
    const bool dk_ok = decode_supported(L, ldx, M);
    if (f == 1) return tc_ok ? 1 : -1;
    if (f == 4) return dk_ok ? 4 : -1;
    if (f >= 0) return -1;
    if ((M <= decode_max_m() || !tc_ok) && dk_ok) return 4;
    return tc_ok ? 1 : -1;
}

int pbl_select_kernel(const pbl_layer* layer, int64_t M) {
    if (!layer) { set_error("pbl_select_kernel: null layer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    return select_impl(L, nullptr, L.K, nullptr, L.N, M);
}

int pbl_decode_variant(const pbl_layer* layer, const void* x, int64_t ldx, int64_t M) {
    if (!layer) { set_error("pbl_decode_variant: null layer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (!L.fsign || M <= 0) return 0;
    return decode_variant(L, x, ldx, M);
}

int pbl_linear_forward(const pbl_layer* layer, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M,
                       void* stream) {
    return pbl_linear_forward_ws(layer, x, ldx, y, ldy, M, nullptr, 0, stream);
}

int pbl_linear_forward_ws(const pbl_layer* layer, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!layer) { set_error("pbl_linear_forward: null layer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (M < 0) { set_error("pbl_linear_forward: M=%lld < 0", (long long)M); return PBL_ERR_SHAPE; }
    if (M == 0) return PBL_OK;
    if (!x || !y) { set_error("pbl_linear_forward: null x or y"); return PBL_ERR_NULL; }
    if (ldx < L.K || ldy < L.N) {
        set_error("pbl_linear_forward: ldx=%lld (K=%lld) or ldy=%lld (N=%lld) too small", (long long)ldx, (long long)L.K,
                  (long long)ldy, (long long)L.N);
        return PBL_ERR_SHAPE;
    }
    int rc = device_check_impl();
    if (rc) return rc;
    const int k = select_impl(L, x, ldx, y, ldy, M);
    if (k < 0) { set_error("no kernel supports this call (PBL_FORCE_KERNEL, or activations too large for 32-bit offsets)"); return PBL_ERR_UNSUPPORTED; }
    if (k == 1) return launch_gemm_twophase(L, x, ldx, y, ldy, M, (cudaStream_t)stream);
    if (k == 4) return launch_decode(L, x, ldx, y, ldy, M, workspace, workspace_bytes, (cudaStream_t)stream);
    return launch_gemv(L, x, ldx, y, ldy, M, (cudaStream_t)stream);
}

int pbl_linear_forward_push(const pbl_layer* layer, const void* x, int64_t ldx, const pbl_peer_push* push, int64_t ldy, int64_t M,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!layer || !push) { set_error("pbl_linear_forward_push: null layer / push"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (M <= 0 || M > 16) { set_error("pbl_linear_forward_push: M=%lld, one decode pass handles 1..16 tokens", (long long)M); return PBL_ERR_SHAPE; }
    if (!x) { set_error("pbl_linear_forward_push: null x"); return PBL_ERR_NULL; }
    if (push->n_ranks < 1 || push->n_ranks > PBL_MAX_PEERS || push->rank < 0 || push->rank >= push->n_ranks) {
        set_error("pbl_linear_forward_push: bad rank %d of %d", push->rank, push->n_ranks); return PBL_ERR_SHAPE;
    }
    if (!push->sync_ctr) { set_error("pbl_linear_forward_push: null sync_ctr"); return PBL_ERR_NULL; }
    for (int d = 0; d < push->n_ranks; ++d)
        if (!push->y[d] || !push->flags[d]) { set_error("pbl_linear_forward_push: null y / flags pointer of rank %d", d); return PBL_ERR_NULL; }
    if (ldx < L.K || ldy < L.N) { set_error("pbl_linear_forward_push: leading dimension too small"); return PBL_ERR_SHAPE; }
    int rc = device_check_impl();
    if (rc) return rc;
    if (!decode_supported(L, ldx, M)) { set_error("pbl_linear_forward_push needs a block-stream (fp16 / bf16) layer"); return PBL_ERR_UNSUPPORTED; }
    return launch_decode(L, x, ldx, nullptr, ldy, M, workspace, workspace_bytes, (cudaStream_t)stream, push);
}

int pbl_peer_wait(const pbl_peer_push* push, void* stream) {
    if (!push || !push->sync_ctr) { set_error("pbl_peer_wait: null push"); return PBL_ERR_NULL; }
    if (push->n_ranks < 1 || push->n_ranks > PBL_MAX_PEERS || push->rank < 0 || push->rank >= push->n_ranks || !push->flags[push->rank]) {
        set_error("pbl_peer_wait: bad rank / flags"); return PBL_ERR_SHAPE;
    }
    int rc = device_check_impl();
    if (rc) return rc;
    return launch_peer_wait(*push, (cudaStream_t)stream);
}

size_t pbl_kth_workspace(void) { return kth_workspace_bytes(); }

int pbl_kth_value(const void* x, int64_t n, int64_t k, int dtype, void* out, void* workspace, void* stream) {
    if (!x || !out || !workspace) { set_error("pbl_kth_value: null pointer"); return PBL_ERR_NULL; }
    if (!valid_dtype(dtype)) { set_error("pbl_kth_value: bad dtype %d", dtype); return PBL_ERR_DTYPE; }
    if (n <= 0 || k < 1 || k > n) { set_error("pbl_kth_value: k=%lld out of range for n=%lld (selected index k out of range)", (long long)k, (long long)n); return PBL_ERR_SHAPE; }
    if (!aligned16(workspace)) { set_error("pbl_kth_value: workspace must be 16 B aligned"); return PBL_ERR_ALIGN; }
    int rc = device_check_impl();
    if (rc) return rc;
    return launch_kth_value(x, n, k, dtype, out, workspace, (cudaStream_t)stream);
}

int pbl_gptq_block(float* W1, int64_t ldw, float* err_out, int64_t lde, const float* hinv_block, int64_t ldh, const uint8_t* mask1,
                   int64_t ldm, const float* low_mean, const float* low_scale, const float* high_scale, const float* high_zero,
                   float maxq, int64_t N, int nc, float* losses, void* stream) {
    if (!W1 || !err_out || !hinv_block || !mask1 || !low_mean || !low_scale || !high_scale || !high_zero) {
        set_error("pbl_gptq_block: null pointer"); return PBL_ERR_NULL;
    }
    if (N <= 0 || nc <= 0 || nc > 128 || ldw < nc || lde < nc || ldh < nc || ldm < nc) {
        set_error("pbl_gptq_block: bad shape (N=%lld, nc=%d must be 1..128, leading dimensions >= nc)", (long long)N, nc); return PBL_ERR_SHAPE;
    }
    int rc = device_check_impl();
    if (rc) return rc;
    return launch_gptq_block(W1, ldw, err_out, lde, hinv_block, ldh, mask1, ldm, low_mean, low_scale, high_scale, high_zero, maxq, N, nc,
                             losses, (cudaStream_t)stream);
}

int pbl_stream_layout(int64_t N, int64_t K, int64_t groupsize, int dtype, pbl_stream_sizes* out) {
    if (!out) { set_error("pbl_stream_layout: out is NULL"); return PBL_ERR_NULL; }
    if (dtype != PBL_F16 && dtype != PBL_BF16) { set_error("the block-stream layout holds fp16 / bf16 layers"); return PBL_ERR_DTYPE; }
    pbl_sizes sz;
    int rc = sizes_impl(N, K, groupsize, dtype, &sz);
    if (rc) return rc;
    out->blocks = sz.tiles_r * kRgPerTile * sz.tiles_c;
    out->fsign_bytes = (size_t)out->blocks * kRgRows * sizeof(uint2);
    out->eptr_bytes = (size_t)(out->blocks + 1) * sizeof(uint32_t);
    return PBL_OK;
}

static int stream_args(const char* who, const void* w, int64_t ldw, const void* affine, int64_t N, int64_t K, int64_t groupsize,
                       int dtype, pbl_sizes* sz, int* tpg) {
    if (dtype != PBL_F16 && dtype != PBL_BF16) { set_error("%s: the block-stream layout holds fp16 / bf16 layers", who); return PBL_ERR_DTYPE; }
    int rc = sizes_impl(N, K, groupsize, dtype, sz);
    if (rc) return rc;
    if (!w || !affine) { set_error("%s: null pointer", who); return PBL_ERR_NULL; }
    if (ldw < K) { set_error("%s: ldw %lld < K %lld", who, (long long)ldw, (long long)K); return PBL_ERR_SHAPE; }
    const int64_t gs = (groupsize <= 0 || groupsize >= K) ? K : groupsize;
    *tpg = (sz->groups == 1) ? (int)sz->tiles_c : (int)(gs / kTileCols);
    return device_check_impl();
}

int pbl_stream_count(const void* w_sim, int64_t ldw, const uint8_t* low_mask, void* affine, int64_t N, int64_t K,
                     int64_t groupsize, int dtype, void* eptr_out, void* stats_out, void* stream) {
    pbl_sizes sz;
    int tpg = 0;
    int rc = stream_args("pbl_stream_count", w_sim, ldw, affine, N, K, groupsize, dtype, &sz, &tpg);
    if (rc) return rc;
    if (!eptr_out || !stats_out) { set_error("pbl_stream_count: null output"); return PBL_ERR_NULL; }
    return launch_stream_count(w_sim, ldw, low_mask, (float2*)affine, N, K, dtype, sz, tpg, (uint32_t*)eptr_out,
                               (uint32_t*)stats_out, (cudaStream_t)stream);
}

int pbl_stream_fill(const void* w_sim, int64_t ldw, const uint8_t* low_mask, const void* affine, int64_t N, int64_t K,
                    int64_t groupsize, int dtype, const void* eptr, void* fsign_out, void* ent_out, void* exc_out,
                    int64_t exc_capacity, void* stats, void* stream) {
    pbl_sizes sz;
    int tpg = 0;
    int rc = stream_args("pbl_stream_fill", w_sim, ldw, affine, N, K, groupsize, dtype, &sz, &tpg);
    if (rc) return rc;
    if (!eptr || !fsign_out || !ent_out || !stats) { set_error("pbl_stream_fill: null pointer"); return PBL_ERR_NULL; }
    if (exc_capacity < 0 || (exc_capacity > 0 && !exc_out)) { set_error("pbl_stream_fill: exception capacity without a buffer"); return PBL_ERR_NULL; }
    if (!aligned16(fsign_out) || !aligned16(ent_out)) { set_error("pbl_stream_fill: fsign/ent must be 16 B aligned"); return PBL_ERR_ALIGN; }
    return launch_stream_fill(w_sim, ldw, low_mask, (const float2*)affine, N, K, dtype, sz, tpg, (const uint32_t*)eptr, (uint2*)fsign_out,
                              (uint32_t*)ent_out, (uint32_t*)exc_out, (uint32_t)exc_capacity, (uint32_t*)stats, (cudaStream_t)stream);
}

int pbl_stream_position(int r, int c, uint32_t* out4) {
    if (!out4) { set_error("pbl_stream_position: out is NULL"); return PBL_ERR_NULL; }
    if (r < 0 || r >= kRgRows || c < 0 || c >= kTileCols) { set_error("pbl_stream_position: (r, c) outside the 32 x 64 block"); return PBL_ERR_SHAPE; }
    st::sign_pos((uint32_t)r, (uint32_t)c, out4[0], out4[1], out4[2]);
    out4[3] = st::tile_slot((uint32_t)r, (uint32_t)c);
    return PBL_OK;
}

int pbl_decode_plan(int64_t N, int64_t K, int64_t M, int sms, int ctas_per_sm, uint32_t* out8) {
    if (!out8) { set_error("pbl_decode_plan: out is NULL"); return PBL_ERR_NULL; }
    if (N <= 0 || K <= 0 || M <= 0 || sms <= 0 || ctas_per_sm <= 0) { set_error("pbl_decode_plan: arguments must be positive"); return PBL_ERR_SHAPE; }
    decode_plan(N, K, M, sms, ctas_per_sm, out8);
    return PBL_OK;
}

void pbl_decode_set_trace(void* device_buf, size_t bytes) { decode_set_trace(device_buf, bytes); }

size_t pbl_decode_workspace_bytes(const pbl_layer* layer, int64_t M) {
    if (!layer || M <= 0) return 0;
    return decode_workspace_bytes(*reinterpret_cast<const Layer*>(layer), M);
}

size_t pbl_bireal_workspace(const pbl_layer* layer, int64_t M) {
    if (!layer || M <= 0) return 0;
    return bireal_workspace_bytes(*reinterpret_cast<const Layer*>(layer), M);
}

size_t pbl_bireal_fixup_workspace(const pbl_layer* layer, int64_t M) {
    if (!layer || M <= 0) return 0;
    return bireal_fixup_workspace_bytes(*reinterpret_cast<const Layer*>(layer), M);
}

int pbl_bireal_forward(const pbl_layer* layer, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy, int64_t M,
                       void* workspace, void* stream) {
    return pbl_bireal_forward_ws(layer, x, ldx, x_dtype, y, ldy, M, workspace, nullptr, 0, stream);
}

int pbl_bireal_forward_ws(const pbl_layer* layer, const void* x, int64_t ldx, int x_dtype, float* y, int64_t ldy, int64_t M,
                          void* workspace, void* fixup_ws, size_t fixup_bytes, void* stream) {
    if (!layer) { set_error("pbl_bireal_forward: null layer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (!L.planes) { set_error("pbl_bireal_forward needs a planes-layout (fp32) layer, as BiRealLinear packs"); return PBL_ERR_DTYPE; }
    if (M < 0 || M > 65535LL * 8) { set_error("pbl_bireal_forward: bad M=%lld", (long long)M); return PBL_ERR_SHAPE; }
    if (M == 0) return PBL_OK;
    if (!x || !y || !workspace) { set_error("pbl_bireal_forward: null pointer"); return PBL_ERR_NULL; }
    if (!valid_dtype(x_dtype)) { set_error("pbl_bireal_forward: bad x dtype %d", x_dtype); return PBL_ERR_DTYPE; }
    if (ldx < L.K || ldy < L.N) { set_error("pbl_bireal_forward: leading dimension too small"); return PBL_ERR_SHAPE; }
    if (!aligned16(workspace)) { set_error("pbl_bireal_forward: workspace must be 16 B aligned"); return PBL_ERR_ALIGN; }
    int rc = device_check_impl();
    if (rc) return rc;
    if (fixup_ws && (reinterpret_cast<uintptr_t>(fixup_ws) & 15u)) { set_error("pbl_bireal_forward: fixup workspace must be 16 B aligned"); return PBL_ERR_ALIGN; }
    return launch_bireal(L, x, ldx, x_dtype, y, ldy, M, workspace, fixup_ws, fixup_bytes, (cudaStream_t)stream);
}

size_t pbl_forward_host_workspace(const pbl_layer* layer, int64_t M) {
    if (!layer || M <= 0) return 0;
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    const size_t es = dtype_size(L.dtype);
    const size_t xb = ((size_t)M * (size_t)L.K * es + 255) / 256 * 256;
    const size_t yb = ((size_t)M * (size_t)L.N * es + 255) / 256 * 256;
    return xb + yb;
}

int pbl_linear_forward_host(const pbl_layer* layer, const void* x_host, void* y_host, int64_t M, void* workspace,
                            void* stream) {
    if (!layer) { set_error("pbl_linear_forward_host: null layer"); return PBL_ERR_NULL; }
    if (M == 0) return PBL_OK;
    if (!x_host || !y_host || !workspace) { set_error("pbl_linear_forward_host: null pointer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    int rc = device_check_impl();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t es = dtype_size(L.dtype);
    const size_t xb = (size_t)M * (size_t)L.K * es, yb = (size_t)M * (size_t)L.N * es;
    char* xd = (char*)workspace;
    char* yd = xd + (xb + 255) / 256 * 256;
    if ((rc = check_cuda(cudaMemcpyAsync(xd, x_host, xb, cudaMemcpyHostToDevice, s), "H2D x"))) return rc;
    if ((rc = pbl_linear_forward(layer, xd, L.K, yd, L.N, M, stream))) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(y_host, yd, yb, cudaMemcpyDeviceToHost, s), "D2H y"))) return rc;
    return check_cuda(cudaStreamSynchronize(s), "stream sync");
}

}  // extern "C"
