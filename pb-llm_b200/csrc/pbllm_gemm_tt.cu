// Two-phase prefill path: expand the packed weight ONCE into a dense fp16/bf16 scratch (L2-resident for the
// usual layer sizes), then run a plain tcgen05 CTA-pair GEMM with BOTH operands fed by TMA.
//
// Kernel 1 is stream_unpack_kernel (pbllm_stream.cu): the expansion cost is paid once per weight per call
// (L2-write bound, ~2*N*K bytes).  Kernel 2 is an ordinary K-major x K-major GEMM: TMA (SWIZZLE_128B) for x and
// for the scratch, cta_group::2 UMMA M=256 N=256, fp32 accumulators in TMEM.
// The scratch holds exactly w_sim (bit-identical to unpack()): with x = I the kernel reproduces w_sim^T bit for bit.
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace tt {
constexpr int BMC = 256, BN = 256, BNC = 128, BK = 64;
constexpr int kStages = 4;
constexpr int kAStage = BMC * BK * 2;  // 32 KB
constexpr int kBStage = BNC * BK * 2;  // 16 KB
constexpr int kEpiWarps = 8;   // two warps per TMEM lane quarter, each draining half of the accumulator columns (16 measured slower: 1334 vs 1357 TFLOP/s)
constexpr int kThreads = (2 + kEpiWarps) * 32;  // 320
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kStages * kAStage;
constexpr int kOffBar = kOffB + kStages * kBStage;
constexpr int kNumBars = 2 * kStages + 2;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;
static_assert(kSmemBytes <= 232448, "exceeds 227 KB");
}  // namespace tt

// Tile rasterisation: n-tiles are taken in groups of kGroupN; inside a group the tile index runs n-fastest, so
// the ~74 pair tiles in flight share one 16 x 256-row weight slab (<= 34 MB at K = 4096) and a few token tiles:
// the slab stays in L2 across the group's waves instead of the whole scratch (90 MB for 11008 x 4096) being
// re-fetched from HBM every wave.
constexpr int kGroupN = 16;
__device__ __forceinline__ void tile_mn(int t, const GemmParams& p, int& m_tile, int& n_tile) {
    const int per_group = kGroupN * p.m_tiles;
    const int g = t / per_group, r = t - g * per_group;
    const int gn = min(kGroupN, p.n_tiles - g * kGroupN);
    m_tile = r / gn;
    n_tile = g * kGroupN + (r - m_tile * gn);
}

// ---- kernel 2: plain CTA-pair tcgen05 GEMM, A = x (TMA), B = dense scratch (TMA) --------------------------------
template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tt::kThreads, 1)
gemm_tt_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
    using namespace tt;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    // the next call's weight expansion (stream_unpack_kernel, launched with the programmatic attribute into the OTHER scratch
    // buffer) may run beside this GEMM: it needs 17 KB of shared memory and a few percent of the issue slots we leave idle
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const uint32_t bar0 = smem_base + kOffBar;
    auto full = [&](int s) { return bar0 + 8u * s; };                        // leader only: A+B bytes of both CTAs
    auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };           // one per CTA (multicast commit)
    const uint32_t tmem_full = bar0 + 8u * (2 * kStages);
    const uint32_t tmem_empty = bar0 + 8u * (2 * kStages + 1);
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemPtr);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kOffTmemPtr), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_tiles = p.m_tiles * p.n_tiles;
    const int my_tiles = (num_tiles - cluster_id + num_clusters - 1) / num_clusters;
    const int KB = p.kblocks;
    const uint32_t stage_bytes = (uint32_t)p.bm * 128u + 2u * kBStage;       // both CTAs: x boxes + weight boxes

    if (warp == 0) {
        // ===== TMA producer: this CTA's x tile and its half of the weight tile; bytes complete on the leader =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                int mt, nt;
                tile_mn(cluster_id + ti * num_clusters, p, mt, nt);
                const int m0 = mt * p.bm + (int)crank * (p.bm >> 1);
                const int n0 = nt * BN + (int)crank * BNC;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty(s), ph ^ 1u);
                    if (leader) mbar_arrive_expect_tx(full(s), stage_bytes);
                    const uint32_t fb = full(s) & 0xFEFFFFFFu;
                    tma_load_2d_2sm(smem_base + kOffA + s * kAStage, &tmap_x, kb * BK, m0, fb);
                    tma_load_2d_2sm(smem_base + kOffB + s * kBStage, &tmap_w, kb * BK, n0, fb);
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader only) =====
        if (leader) {
            const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((256u >> 4) << 24);
            int s = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                int mt, nt;
                tile_mn(cluster_id + ti * num_clusters, p, mt, nt);
                const int m0 = mt * p.bm;
                const int halves = (p.bm == 2 * BMC && m0 + 128 < p.M) ? 2 : 1;
                mbar_wait_cluster(tmem_empty, acc_ph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_cluster(full(s), ph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_base + kOffA + s * kAStage, b_addr = smem_base + kOffB + s * kBStage;
                        for (int h = 0; h < halves; ++h) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16_2sm(tmem_base + h * 256, make_sw128_desc(a_addr + h * (128 * 128) + k * 32),
                                             make_sw128_desc(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_2sm(empty(s));
                        if (kb == KB - 1) umma_commit_2sm(tmem_full);
                    }
                    __syncwarp();
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
                acc_ph ^= 1u;
            }
        }
    } else {
        // ===== epilogue: this CTA's 256 tokens x 256 weight rows (TMEM -> regs -> +bias -> 16 bit -> global) =====
        const int q = warp & 3;                      // TMEM lane quarter this warp may read (warp id % 4)
        const int cpart = (warp - 2) >> 2;           // which part of the 256 accumulator columns it drains
        constexpr int kChunksPerPart = (BN / 32) / (kEpiWarps / 4);
        uint32_t acc_ph = 0;
        T* y = reinterpret_cast<T*>(p.y);
        const uint32_t tmem_empty_leader = mapa_rank0(tmem_empty);
        for (int ti = 0; ti < my_tiles; ++ti) {
            int mt, nt;
            tile_mn(cluster_id + ti * num_clusters, p, mt, nt);
            const int mp = mt * p.bm;
            const int m0 = mp + (int)crank * (p.bm >> 1), n0 = nt * BN;
            const int halves = (p.bm == 2 * BMC && mp + 128 < p.M) ? 2 : 1;      // which accumulators the MMA warp wrote
            mbar_wait(tmem_full, acc_ph);
            tc_fence_after();
            for (int h = 0; h < halves; ++h) {
                const int m = m0 + h * 128 + q * 32 + lane;
                if (m0 + h * 128 >= p.M) break;                // warp-uniform: no valid token in this half
#pragma unroll 1
                for (int cb = cpart * kChunksPerPart; cb < (cpart + 1) * kChunksPerPart; ++cb) {
                    const int n = n0 + cb * 32;
                    if (n >= p.N) break;
                    uint32_t acc[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256 + cb * 32), acc);
                    tmem_ld_wait();
                    if (m < p.M) {
                        T* yrow = y + (int64_t)m * p.ldy + n;
#pragma unroll
                        for (int v8 = 0; v8 < 4; ++v8) {
                            if (n + v8 * 8 + 8 <= p.N) {
                                float f[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(acc[v8 * 8 + i]);
                                if (p.bias) {
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n + v8 * 8));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + v8 * 8 + 4));
                                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                                    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                                }
                                uint4 o;
                                o.x = pack2<T>(f[0], f[1]); o.y = pack2<T>(f[2], f[3]);
                                o.z = pack2<T>(f[4], f[5]); o.w = pack2<T>(f[6], f[7]);
                                *reinterpret_cast<uint4*>(yrow + v8 * 8) = o;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
            acc_ph ^= 1u;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// TMA boxes and the 16-byte epilogue stores need aligned, 8-element-multiple strides; anything else runs the decode
// kernel in token passes
bool gemm_twophase_supported(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M) {
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) return false;
    if (!L.fsign) return false;
    if (M <= 0 || M > (1 << 30)) return false;
    if (L.N % 8 != 0 || ldy % 8 != 0 || ldx % 8 != 0) return false;
    if (x && (reinterpret_cast<uintptr_t>(x) & 15u)) return false;
    if (y && (reinterpret_cast<uintptr_t>(y) & 15u)) return false;
    if (L.bias && (reinterpret_cast<uintptr_t>(L.bias) & 15u)) return false;
    return true;
}

// ---- the dense scratch -------------------------------------------------------------------------------------------
// Two persistent buffers per (device, stream), used alternately: call i+1 expands into the buffer call i's GEMM is NOT
// reading, so its expansion kernel can be launched early and run beside that GEMM (which it could not with one
// stream-ordered allocation per call: the pool hands call i+1 the block call i has just freed).  Grown on demand
// (cudaFree + cudaMalloc, synchronising, once per new maximum), kept for the life of the process.  Streams beyond the
// table, streams under capture and allocation failures fall back to cudaMallocAsync / cudaFreeAsync without overlap.
namespace {
struct ScratchSlot { cudaStream_t stream; void* buf[2]; size_t bytes[2]; int flip; bool used; };
constexpr int kScratchStreams = 8;
ScratchSlot g_scratch[64][kScratchStreams] = {};
std::mutex g_scratch_mu;
bool scratch_overlap_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PBL_PREFILL_OVERLAP"); v = (e && *e) ? atoi(e) : 1; }
    return v != 0;
}
void* scratch_acquire(int dev, cudaStream_t s, size_t bytes) {
    if (!scratch_overlap_enabled() || dev < 0 || dev >= 64) return nullptr;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { (void)cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    ScratchSlot* slot = nullptr;
    for (int i = 0; i < kScratchStreams && !slot; ++i)
        if (g_scratch[dev][i].used && g_scratch[dev][i].stream == s) slot = &g_scratch[dev][i];
    for (int i = 0; i < kScratchStreams && !slot; ++i)
        if (!g_scratch[dev][i].used) { slot = &g_scratch[dev][i]; slot->used = true; slot->stream = s; }
    if (!slot) return nullptr;
    const int b = slot->flip ^= 1;
    if (slot->bytes[b] < bytes) {
        if (slot->buf[b]) cudaFree(slot->buf[b]);          // synchronises: nothing in flight still reads it
        slot->buf[b] = nullptr;
        slot->bytes[b] = 0;
        if (cudaMalloc(&slot->buf[b], bytes) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        slot->bytes[b] = bytes;
    }
    return slot->buf[b];
}
}  // namespace

int launch_gemm_twophase(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable"); return PBL_ERR_CUDA; }
    static int sms_dev[64] = {};
    static bool pool_set_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms_dev[dev]) cudaDeviceGetAttribute(&sms_dev[dev], cudaDevAttrMultiProcessorCount, dev);
    const int num_sms = sms_dev[dev] > 0 ? sms_dev[dev] : 148;
    bool& pool_set = pool_set_dev[dev];
    if (!pool_set) {   // keep freed scratch cached in the stream-ordered pool instead of returning it to the OS
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        (void)cudaGetLastError();
        pool_set = true;
    }
    const size_t scratch_bytes = (size_t)L.n_pad * (size_t)L.k_pad * 2;
    void* scratch = scratch_acquire(dev, s, scratch_bytes);
    const bool pooled = scratch == nullptr;                   // transient allocation, no overlap with the kernel ahead
    int rc = PBL_OK;
    if (pooled) {
        rc = check_cuda(cudaMallocAsync(&scratch, scratch_bytes, s), "cudaMallocAsync(weight scratch)");
        if (rc) return rc;
    }
    auto release = [&]() { return pooled ? cudaFreeAsync(scratch, s) : cudaSuccess; };

    const int which = L.dtype == PBL_F16 ? 0 : 1;
    rc = launch_stream_unpack(L, scratch, L.k_pad, L.n_pad, L.k_pad, s, !pooled);   // kernel 1: the exact w_sim, padded with level values
    if (rc) { release(); return rc; }

    const int n_tiles = (int)((L.N + tt::BN - 1) / tt::BN);
    const int bm = (((M + 511) / 512) * (int64_t)n_tiles >= num_sms / 2) ? 512 : 256;
    const CUtensorMapDataType dt = L.dtype == PBL_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap tmx, tmw;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)L.K, (cuuint64_t)M};
        const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
        const cuuint32_t box[2] = {(cuuint32_t)tt::BK, (cuuint32_t)(bm / 2)};
        CUresult cr = enc(&tmx, dt, 2, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { release(); set_error("cuTensorMapEncodeTiled(x) failed (%d)", (int)cr); return PBL_ERR_CUDA; }
    }
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)L.k_pad, (cuuint64_t)L.n_pad};
        const cuuint64_t gstr[1] = {(cuuint64_t)L.k_pad * 2};
        const cuuint32_t box[2] = {(cuuint32_t)tt::BK, (cuuint32_t)tt::BNC};
        CUresult cr = enc(&tmw, dt, 2, scratch, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { release(); set_error("cuTensorMapEncodeTiled(w) failed (%d)", (int)cr); return PBL_ERR_CUDA; }
    }
    GemmParams p;
    p.bias = L.bias; p.y = y; p.ldy = ldy; p.M = (int)M; p.N = (int)L.N; p.K = (int)L.K;
    p.tiles_r = (int)L.tiles_r; p.tiles_c = (int)L.tiles_c; p.groups = (int)L.groups; p.tiles_per_group = L.tiles_per_group;
    p.bm = bm;
    p.m_tiles = (int)((M + bm - 1) / bm);
    p.n_tiles = n_tiles;
    p.kblocks = (int)L.tiles_c;

    static bool attr_set_dev[2][64] = {};
    auto kern = which == 0 ? gemm_tt_kernel<__half> : gemm_tt_kernel<__nv_bfloat16>;
    bool attr_local = false;
    bool& attr_done = (dev >= 0 && dev < 64) ? attr_set_dev[which][dev] : attr_local;
    if (!attr_done) {
        rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tt::kSmemBytes),
                        "cudaFuncSetAttribute(smem, tt)");
        if (rc) { release(); return rc; }
        attr_done = true;
    }
    const int ntiles = p.m_tiles * p.n_tiles;
    const int max_clusters = num_sms / 2;
    const int clusters = ntiles < max_clusters ? ntiles : max_clusters;
    kern<<<2 * clusters, tt::kThreads, tt::kSmemBytes, s>>>(tmx, tmw, p);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "gemm_tt launch");
    cudaError_t fe = release();
    if (!rc) rc = check_cuda(fe, "cudaFreeAsync(weight scratch)");
    return rc;
}

}  // namespace pbl
