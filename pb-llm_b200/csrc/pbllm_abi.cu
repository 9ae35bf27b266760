// C ABI of libpbllm.so (see include/pbllm.h). Host-side glue only: argument checking, the
// opaque layer handle, kernel selection, thread-local error string. No torch types, no
// exceptions across the boundary, no CPU fallback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pbllm_common.cuh"

namespace pbl {
int launch_gemm_splitk(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s);
bool gemm_splitk_supported(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M);
}  // namespace pbl

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
    if (!d->planes || !d->vptr || !d->vals || !d->affine) { set_error("pbl_layer_create: null packed buffer"); return PBL_ERR_NULL; }
    if (!aligned16(d->planes) || !aligned16(d->vals)) { set_error("pbl_layer_create: planes/vals must be 16 B aligned"); return PBL_ERR_ALIGN; }
    Layer* L = new (std::nothrow) Layer();
    if (!L) { set_error("out of host memory"); return PBL_ERR_CUDA; }
    L->N = d->N; L->K = d->K;
    L->groupsize = (d->groupsize <= 0 || d->groupsize >= d->K) ? d->K : d->groupsize;
    L->dtype = d->dtype;
    L->n_pad = sz.n_pad; L->k_pad = sz.k_pad; L->tiles_r = sz.tiles_r; L->tiles_c = sz.tiles_c; L->groups = sz.groups;
    L->tiles_per_group = (sz.groups == 1) ? (int)sz.tiles_c : (int)(L->groupsize / kTileCols);
    L->planes = (const uint4*)d->planes; L->vptr = (const uint32_t*)d->vptr; L->vals = d->vals;
    L->affine = (const float2*)d->affine; L->bias = (const float*)d->bias;
    L->sign_planes = (const uint2*)d->sign_planes;
    L->dsign = nullptr; L->eptr = nullptr; L->ent = nullptr;
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
    return launch_unpack(L, w_out, ldw, (cudaStream_t)stream);
}

static int forced_kernel() {
    const char* e = getenv("PBL_FORCE_KERNEL");
    if (!e || !*e) return -1;
    return atoi(e);
}

static int skinny_max_m() {
    const char* e = getenv("PBL_SKINNY_MAX_M");
    if (e && *e) return atoi(e);
    return 16;
}

static int splitk_max_m() {   // M up to which the split-K cluster kernel is preferred (0 disables it)
    const char* e = getenv("PBL_SPLITK_MAX_M");
    if (e && *e) return atoi(e);
    return 128;
}

// 0 = CUDA-core bit-plane kernel, 1 = tcgen05 GEMM (single-CTA / CTA-pair), 2 = mma.sync skinny kernel,
// 3 = tcgen05 split-K cluster kernel (M <= 128), 4 = decode kernel (decode index attached); -1 = forced kernel unsupported
static int select_impl(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M) {
    const bool tc_ok = gemm_tc_supported(L, x, ldx, y, ldy, M);
    const bool sk_ok = skinny_supported(L, M);
    const bool ck_ok = gemm_splitk_supported(L, x, ldx, y, ldy, M);
    const int f = forced_kernel();
    if (f == 0) return 0;
    if (f == 1) return tc_ok ? 1 : -1;
    if (f == 2) return sk_ok ? 2 : -1;
    if (f == 3) return ck_ok ? 3 : -1;
    if (f == 4) return decode_supported(L, ldx, M) ? 4 : -1;
    if (!sk_ok) return 0;                       // fp32 I/O: CUDA cores
    if (M <= skinny_max_m() && decode_supported(L, ldx, M)) return 4;   // decode: positioned-entry kernel, stream-K over warps
    if (M <= skinny_max_m()) return 2;          // decode: mma.sync skinny kernel (measured faster than split-K up to 16 tokens)
    if (ck_ok && M <= splitk_max_m()) return 3; // short prompts: split-K cluster kernel (2x the single-CTA GEMM at M = 64)
    if (!tc_ok) return 2;
    return 1;
}

int pbl_select_kernel(const pbl_layer* layer, int64_t M) {
    if (!layer) { set_error("pbl_select_kernel: null layer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    return select_impl(L, nullptr, L.K, nullptr, L.N, M);
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
    if (k < 0) { set_error("PBL_FORCE_KERNEL names a kernel that does not support this call"); return PBL_ERR_UNSUPPORTED; }
    if (k == 1) return launch_gemm_tc(L, x, ldx, y, ldy, M, (cudaStream_t)stream);
    if (k == 2) return launch_skinny(L, x, ldx, y, ldy, M, (cudaStream_t)stream);
    if (k == 3) return launch_gemm_splitk(L, x, ldx, y, ldy, M, (cudaStream_t)stream);
    if (k == 4) return launch_decode(L, x, ldx, y, ldy, M, workspace, workspace_bytes, (cudaStream_t)stream);
    return launch_gemv(L, x, ldx, y, ldy, M, (cudaStream_t)stream);
}

int pbl_decode_index_sizes(const pbl_layer* layer, pbl_decode_sizes* out) {
    if (!layer || !out) { set_error("pbl_decode_index_sizes: null pointer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) { set_error("the decode index needs an fp16 / bf16 layer"); return PBL_ERR_DTYPE; }
    out->blocks = L.tiles_r * kRgPerTile * L.tiles_c;
    out->dsign_bytes = (size_t)out->blocks * kRgRows * sizeof(uint2);
    out->eptr_bytes = (size_t)(out->blocks + 1) * sizeof(uint32_t);
    return PBL_OK;
}

int pbl_decode_index_count(const pbl_layer* layer, void* eptr_out, void* stream) {
    if (!layer || !eptr_out) { set_error("pbl_decode_index_count: null pointer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) { set_error("the decode index needs an fp16 / bf16 layer"); return PBL_ERR_DTYPE; }
    int rc = device_check_impl();
    if (rc) return rc;
    return launch_decode_index_count(L, (uint32_t*)eptr_out, (cudaStream_t)stream);
}

int pbl_decode_index_fill(const pbl_layer* layer, const void* eptr, void* dsign_out, void* ent_out, void* stream) {
    if (!layer || !eptr || !dsign_out || !ent_out) { set_error("pbl_decode_index_fill: null pointer"); return PBL_ERR_NULL; }
    const Layer& L = *reinterpret_cast<const Layer*>(layer);
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) { set_error("the decode index needs an fp16 / bf16 layer"); return PBL_ERR_DTYPE; }
    if (!aligned16(dsign_out) || !aligned16(ent_out)) { set_error("pbl_decode_index_fill: dsign/ent must be 16 B aligned"); return PBL_ERR_ALIGN; }
    int rc = device_check_impl();
    if (rc) return rc;
    return launch_decode_index_fill(L, (const uint32_t*)eptr, (uint2*)dsign_out, (uint32_t*)ent_out, (cudaStream_t)stream);
}

int pbl_layer_attach_decode_index(pbl_layer* layer, const void* dsign, const void* eptr, const void* ent) {
    if (!layer) { set_error("pbl_layer_attach_decode_index: null layer"); return PBL_ERR_NULL; }
    Layer& L = *reinterpret_cast<Layer*>(layer);
    if (!dsign && !eptr && !ent) { L.dsign = nullptr; L.eptr = nullptr; L.ent = nullptr; return PBL_OK; }
    if (!dsign || !eptr || !ent) { set_error("pbl_layer_attach_decode_index: all three buffers or none"); return PBL_ERR_NULL; }
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) { set_error("the decode index needs an fp16 / bf16 layer"); return PBL_ERR_DTYPE; }
    if (!aligned16(dsign) || !aligned16(ent)) { set_error("pbl_layer_attach_decode_index: dsign/ent must be 16 B aligned"); return PBL_ERR_ALIGN; }
    L.dsign = (const uint2*)dsign; L.eptr = (const uint32_t*)eptr; L.ent = (const uint32_t*)ent;
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
