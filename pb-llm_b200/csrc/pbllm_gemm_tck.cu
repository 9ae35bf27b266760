// tcgen05 bit-plane GEMM, split-K cluster version: the small-M (decode / short-prompt) kernel, M <= 128.
//
// With few tokens there are only N/256 output tiles, far fewer than SMs, and each tile's cost is the
// weight expansion over the whole K extent.  Here a thread-block CLUSTER of C CTAs shares one
// 256-weight-row tile and splits K: CTA r expands and multiplies only k-blocks [r*KB/C, (r+1)*KB/C)
// (same TMA / in-smem exact-tile expansion / tcgen05.mma pipeline as gemm_tc_kernel, fp32 partial
// accumulator in TMEM), then the partials are reduced through distributed shared memory in a fixed
// order (deterministic): every CTA parks its [tokens x 256] partial in its own shared memory, and after
// a cluster barrier CTA r sums column slice r of all C partials (ld.shared::cluster), adds the bias and
// stores y.  The token tile is the UMMA M operand (M=128); only the first box rows are loaded by TMA --
// an output row depends only on its own A row, so the unloaded rows produce TMEM lanes nobody reads.
// Programmatic dependent launch: the weight prefetch of the next linear overlaps this one's tail.
#include <cstdlib>
#include <type_traits>

#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace tck {
constexpr int BN = 256, BK = 64;
constexpr int kStages = 4;
constexpr int kAStage = 128 * BK * 2;  // 16 KB (only the first box rows are written)
constexpr int kBStage = BN * BK * 2;   // 32 KB
constexpr int kTeams = 2;
constexpr int kExpWarps = 8 * kTeams, kEpiWarps = 4;
constexpr int kExpThreads = 256;
constexpr int kThreads = (2 + kExpWarps + kEpiWarps) * 32;  // 704
constexpr int kScratchVals = 512;
constexpr int kScratchBytes = kScratchVals * 2;
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kStages * kAStage;
constexpr int kOffScratch = kOffB + kStages * kBStage;
constexpr int kOffBar = kOffScratch + kExpWarps * kScratchBytes;
constexpr int kNumBars = 3 * kStages + 1;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;
constexpr int kPartStride = 260;        // floats per token row of the parked partial (conflict-free STS.128)
static_assert(kTeams <= kStages, "teams must not outnumber stages");
static_assert(128 * kPartStride * 4 <= kOffScratch, "partial tile must fit in the (finished) stage memory");
static_assert(kSmemBytes <= 232448, "exceeds 227 KB");
}  // namespace tck

template <typename T>
__global__ void __launch_bounds__(tck::kThreads, 1)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p, const int a_rows) {
    using namespace tck;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t csize;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const uint32_t crank = cluster_ctarank();
    const int n_tile = (int)(blockIdx.x / csize);
    const int KB = p.kblocks;
    const int kb_lo = (int)((int64_t)crank * KB / csize), kb_hi = (int)((int64_t)(crank + 1) * KB / csize);
    const int n_kb = kb_hi - kb_lo;

    const uint32_t bar0 = smem_base + kOffBar;
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto full_b = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    const uint32_t tmem_full = bar0 + 8u * (3 * kStages);
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemPtr);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_a(s), 1);
            mbar_init(full_b(s), kExpThreads);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kOffTmemPtr), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===== TMA producer: the token tile's first a_rows rows x 64 per k-block =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
            asm volatile("griddepcontrol.wait;" ::: "memory");     // x belongs to the producer kernel
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_kb; ++i) {
                mbar_wait(empty(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_a(s), (uint32_t)a_rows * 128u);
                tma_load_2d(smem_base + kOffA + s * kAStage, &tmap_x, (kb_lo + i) * BK, 0, full_a(s));
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: UMMA M=128 (tokens), N=256 (weight rows), K=16 =====
        const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_kb; ++i) {
            mbar_wait(full_a(s), ph);
            mbar_wait(full_b(s), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_base + kOffA + s * kAStage, b_addr = smem_base + kOffB + s * kBStage;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_f16(tmem_base, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                             (i | k) != 0 ? 1u : 0u);
                umma_commit(empty(s));
                if (i == n_kb - 1) umma_commit(tmem_full);
            }
            __syncwarp();
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    } else if (warp < 2 + kExpWarps) {
        // ===== weight expansion: thread = weight row of the 256-row tile; kTeams teams of 8 warps
        //       take alternate k-blocks so two shared-memory stages are being filled concurrently =====
        const int et = threadIdx.x - 64;
        const int team = et >> 8;              // 0..kTeams-1
        const int e = et & 255;                // weight row within the CTA tile
        const int ew = et >> 5;                // expansion warp (own scratch)
        const int r = e & 127;                 // row within its 128-row plane tile
        const int rgi = r >> 5;                // row group within the plane tile (warp-uniform)
        const uint32_t r7 = (uint32_t)(e & 7);
        const uint32_t row_off = (uint32_t)(e >> 3) * 1024u + r7 * 128u;
        const uint32_t scratch = smem_base + kOffScratch + ew * kScratchBytes;
        const bool grouped = p.groups > 1;

        // items = this CTA's k-blocks kb_lo + team, kb_lo + team + kTeams, ... < kb_hi of its one tile
        struct Meta { uint4 pw; uint32_t cs, ce; int tr, g; };
        int c_kb = kb_lo + team;               // prefetch cursor
        const int my_tr = n_tile * 2 + (e >> 7);
        auto load_meta = [&]() {               // loads the cursor's item and advances the cursor
            Meta m;
            m.pw = make_uint4(0, 0, 0, 0);
            m.cs = m.ce = 0;
            m.tr = -1; m.g = 0;
            if (c_kb < kb_hi) {
                if (my_tr < p.tiles_r) {
                    const int64_t tile = (int64_t)my_tr * p.tiles_c + c_kb;
                    m.pw = __ldg(p.planes + tile * kTileRows + r);
                    m.cs = __ldg(p.vptr + tile * kRgPerTile + rgi);
                    m.ce = __ldg(p.vptr + tile * kRgPerTile + rgi + 1);
                    m.tr = my_tr;
                    m.g = grouped ? c_kb / p.tiles_per_group : 0;
                }
                c_kb += kTeams;
            }
            return m;
        };
        auto load_vals = [&](const Meta& m, uint4& q0, uint4& q1) {   // coalesced prefetch of the value chunk
            const uint32_t b0 = (m.cs * 2u) & ~15u, b1 = m.ce * 2u;
            const uint8_t* base = reinterpret_cast<const uint8_t*>(p.vals);
            const uint32_t o0 = b0 + 16u * lane, o1 = o0 + 512u;
            if (o0 < b1) q0 = __ldg(reinterpret_cast<const uint4*>(base + o0));
            if (o1 < b1) q1 = __ldg(reinterpret_cast<const uint4*>(base + o1));
        };

        const int64_t total = n_kb;
        const int64_t my_items = (total - team + kTeams - 1) / kTeams;
        Meta m0 = load_meta(), m1 = load_meta();
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
        if (my_items > 0) load_vals(m0, q0, q1);

        int s = team % kStages;
        uint32_t ph = 0;
        int cur_g = -1, cur_tr = -2;
        uint32_t LL = 0, DD = 0;
        for (int64_t it = 0; it < my_items; ++it) {
            const uint4 pw = m0.pw;
            const uint32_t cs = m0.cs, ce = m0.ce;
            const int tr = m0.tr, g = m0.g;
            const uint4 v0 = q0, v1 = q1;
            // issue the next items' global loads before touching shared memory
            const Meta m2 = load_meta();
            if (it + 1 < my_items) load_vals(m1, q0, q1);
            m0 = m1;
            m1 = m2;

            if (g != cur_g || tr != cur_tr) {
                cur_g = g; cur_tr = tr;
                float2 a = make_float2(0.f, 0.f);
                if (tr >= 0) a = __ldg(p.affine + ((int64_t)tr * kTileRows + r) * p.groups + g);
                const uint32_t lo = bits16<T>(a.x), hi = bits16<T>(a.y);
                LL = lo | (lo << 16);
                DD = (lo ^ hi) * 0x10001u;
            }

            mbar_wait(empty(s), ph ^ 1u);

            const uint32_t brow = smem_base + kOffB + s * kBStage + row_off;
            expand_row(pw, LL, DD, brow, r7, cs, ce, v0, v1, scratch, p.vals, (uint32_t)lane);
            fence_proxy_async();
            mbar_arrive(full_b(s));
            s += kTeams;
            if (s >= kStages) { s -= kStages; ph ^= 1u; }
        }
    } else if (n_kb > 0) {
        // ===== epilogue part 1: park this CTA's fp32 partial [tokens x 256] in shared memory =====
        const int q = warp & 3;                    // TMEM lane quadrant = tokens 32q .. 32q+31
        mbar_wait(tmem_full, 0u);
        tc_fence_after();
        if (q * 32 < p.M) {
            const uint32_t part = smem_base + (uint32_t)(q * 32 + lane) * (uint32_t)(kPartStride * 4);
#pragma unroll 1
            for (int cb = 0; cb < BN / 32; ++cb) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), acc);
                tmem_ld_wait();
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    sts_v4(part + (uint32_t)(cb * 32 + v * 4) * 4u, acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            }
        }
        tc_fence_before();
    }

    // ===== epilogue part 2: cluster-wide reduction of the partials, column slice per CTA =====
    tc_fence_before();
    cluster_sync_all();
    asm volatile("griddepcontrol.wait;" ::: "memory");             // y may still be read by the previous kernel
    {
        const int n0 = n_tile * BN;
        const int c_lo = (int)((int64_t)crank * BN / csize), c_hi = (int)((int64_t)(crank + 1) * BN / csize);
        const int W = c_hi - c_lo, count = p.M * W;
        T* y = reinterpret_cast<T*>(p.y);
        // ranks whose k-range is empty parked nothing: skip them (n_kb > 0 <=> KB*(r+1)/C > KB*r/C)
        for (int idx = threadIdx.x; idx < count; idx += kThreads) {
            const int tok = idx / W, col = c_lo + idx - tok * W;
            const int n = n0 + col;
            if (n < p.N) {
                const uint32_t off = smem_base + (uint32_t)(tok * kPartStride + col) * 4u;
                float s = p.bias ? __ldg(p.bias + n) : 0.f;
                for (uint32_t c = 0; c < csize; ++c) {
                    const int lo_c = (int)((int64_t)c * KB / csize), hi_c = (int)((int64_t)(c + 1) * KB / csize);
                    if (hi_c > lo_c) s += ld_cluster_f32(mapa_rank(off, c));
                }
                y[(int64_t)tok * p.ldy + n] = from_f32<T>(s);
            }
        }
    }
    cluster_sync_all();       // peers may still be reading this CTA's partial
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------
bool gemm_splitk_supported(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M) {
    if (M <= 0 || M > 128) return false;
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) return false;
    if (ldx % 8 != 0) return false;
    if (x && (reinterpret_cast<uintptr_t>(x) & 15u)) return false;
    if (L.groups > 1 && L.groupsize % tck::BK != 0) return false;
    (void)y; (void)ldy;
    return true;
}

int launch_gemm_splitk(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable"); return PBL_ERR_CUDA; }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int a_rows = M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128));
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)L.K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)tck::BK, (cuuint32_t)a_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, L.dtype == PBL_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return PBL_ERR_CUDA; }

    GemmParams p;
    p.planes = L.planes; p.vptr = L.vptr; p.vals = reinterpret_cast<const uint16_t*>(L.vals); p.affine = L.affine;
    p.bias = L.bias; p.y = y; p.ldy = ldy; p.M = (int)M; p.N = (int)L.N; p.K = (int)L.K;
    p.tiles_r = (int)L.tiles_r; p.tiles_c = (int)L.tiles_c; p.groups = (int)L.groups; p.tiles_per_group = L.tiles_per_group;
    p.bm = 128;
    p.m_tiles = 1;
    p.n_tiles = (int)((L.N + tck::BN - 1) / tck::BN);
    p.kblocks = (int)L.tiles_c;

    // cluster size: as many K-splits as keep (tiles x splits) within one wave of SMs, at most 8 (portable limit)
    int C = num_sms / p.n_tiles;
    if (C > 8) C = 8;
    if (C > p.kblocks) C = p.kblocks;
    if (C < 1) C = 1;
    const char* e = getenv("PBL_SPLITK_C");
    if (e && *e) { C = atoi(e); if (C < 1) C = 1; if (C > 8) C = 8; if (C > p.kblocks) C = p.kblocks; }

    static bool attr_set_dev[2][64] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const int which = L.dtype == PBL_F16 ? 0 : 1;
    auto kern = which == 0 ? gemm_splitk_kernel<__half> : gemm_splitk_kernel<__nv_bfloat16>;
    bool attr_local = false;
    bool& attr_done = (cur_dev >= 0 && cur_dev < 64) ? attr_set_dev[which][cur_dev] : attr_local;
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tck::kSmemBytes),
                            "cudaFuncSetAttribute(smem, splitk)");
        if (rc) return rc;
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.n_tiles * C));
    cfg.blockDim = dim3(tck::kThreads);
    cfg.dynamicSmemBytes = tck::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    static int pdl = -1;
    if (pdl < 0) { const char* e2 = getenv("PBL_PDL"); pdl = (e2 && *e2) ? atoi(e2) : 1; }
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmap, p, a_rows);
    count_launch();
    return check_cuda(le, "gemm_splitk launch");
}

}  // namespace pbl
