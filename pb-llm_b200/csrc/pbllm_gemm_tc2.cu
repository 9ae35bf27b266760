// tcgen05 bit-plane GEMM, CTA-pair version (cta_group::2): the large-M prefill kernel.
//
// Same algorithm as gemm_tc_kernel (pbllm_gemm_tc.cu) -- activations by TMA, the exact fp16/bf16
// w_sim tile rebuilt in shared memory from bit planes + salient values, fp32 accumulators in TMEM --
// but two CTAs on an SM pair (a 2-CTA cluster) execute each MMA together: UMMA M=256 (128 tokens
// from each CTA), N=256 weight rows of which each CTA expands and holds only HALF (128 rows).
// Per CTA and k-block that halves the expansion work and the shared-memory traffic of the weight
// operand, which is what bounds the single-CTA kernel (ncu r01: smem pipe saturated).
//
// Pair tile: 512 tokens x 256 weight rows x 64 (k-block). CTA c owns tokens [m0+256c, m0+256c+256)
// (two UMMA halves -> both 256-column TMEM accumulators) and weight rows [n0+128c, n0+128c+128).
// The leader CTA (cluster rank 0) issues every tcgen05.mma.cta_group::2; completion is multicast to
// both CTAs' barriers; the peer's producers arrive on the leader's barriers through DSMEM.
// Warp roles per CTA: 0 TMA (own x tile), 1 TMEM alloc (+ MMA issue on the leader), 2..17 weight
// expansion (4 teams x 4 warps, thread = weight row, teams take every 4th k-block), 18..21 epilogue.
#include <type_traits>

#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace tc2 {
constexpr int BMC = 256;            // tokens per CTA
constexpr int BN = 256, BNC = 128;  // weight rows per pair / per CTA
constexpr int BK = 64;
constexpr int kStages = 4;
constexpr int kAStage = BMC * BK * 2;  // 32 KB
constexpr int kBStage = BNC * BK * 2;  // 16 KB
constexpr int kTeams = 4, kTeamWarps = 4;
constexpr int kExpWarps = kTeams * kTeamWarps, kEpiWarps = 4;
constexpr int kThreads = (2 + kExpWarps + kEpiWarps) * 32;  // 704
constexpr int kScratchVals = 512;
constexpr int kScratchBytes = kScratchVals * 2;
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kStages * kAStage;
constexpr int kOffScratch = kOffB + kStages * kBStage;
constexpr int kOffBar = kOffScratch + kExpWarps * kScratchBytes;
constexpr int kNumBars = 3 * kStages + 2;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;
static_assert(kTeams <= kStages, "teams must not outnumber stages");
static_assert(kSmemBytes <= 232448, "exceeds 227 KB");
}  // namespace tc2

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc2::kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p) {
    using namespace tc2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();          // 0 = leader
    const bool leader = crank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    const uint32_t bar0 = smem_base + kOffBar;
    auto full_a = [&](int s) { return bar0 + 8u * s; };                      // used on the leader only
    auto full_b = [&](int s) { return bar0 + 8u * (kStages + s); };          // used on the leader only
    auto empty = [&](int s) { return bar0 + 8u * (2 * kStages + s); };       // one per CTA (multicast commit)
    const uint32_t tmem_full = bar0 + 8u * (3 * kStages);                    // one per CTA (multicast commit)
    const uint32_t tmem_empty = bar0 + 8u * (3 * kStages + 1);               // used on the leader only
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemPtr);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_a(s), 1);                        // leader's arrive.expect_tx; both CTAs' TMA bytes
            mbar_init(full_b(s), 2 * kTeamWarps);           // one arrive per expansion warp, both CTAs
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * kEpiWarps);               // one arrive per epilogue warp, both CTAs
        fence_barrier_init();
    }
    if (warp == 1) {  // both CTAs of the pair allocate (same warp id, same smem offset)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kOffTmemPtr), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();   // barrier inits + TMEM address visible pair-wide before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_tiles = p.m_tiles * p.n_tiles;       // pair tiles: 512 tokens x 256 rows
    const int my_tiles = (num_tiles - cluster_id + num_clusters - 1) / num_clusters;
    const int KB = p.kblocks;

    if (warp == 0) {
        // ===== TMA producer: this CTA's x tile [256 tokens x 64]; bytes complete on the LEADER's barrier =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int t = cluster_id + ti * num_clusters;
                const int m0 = (t / p.n_tiles) * p.bm + (int)crank * (p.bm >> 1);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty(s), ph ^ 1u);
                    if (leader) mbar_arrive_expect_tx(full_a(s), (uint32_t)p.bm * 128u);   // both CTAs' boxes
                    tma_load_2d_2sm(smem_base + kOffA + s * kAStage, &tmap_x, kb * BK, m0, full_a(s) & 0xFEFFFFFFu);
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): UMMA M=256 (128 tokens per CTA), N=256, K=16 =====
        if (leader) {
            const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((256u >> 4) << 24);
            int s = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int t = cluster_id + ti * num_clusters;
                const int m0 = (t / p.n_tiles) * p.bm;
                const int halves = (p.bm == 2 * BMC && m0 + 128 < p.M) ? 2 : 1;
                mbar_wait_cluster(tmem_empty, acc_ph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_cluster(full_a(s), ph);
                    mbar_wait_cluster(full_b(s), ph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_base + kOffA + s * kAStage, b_addr = smem_base + kOffB + s * kBStage;
                        for (int h = 0; h < halves; ++h) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t da = make_sw128_desc(a_addr + h * (128 * 128) + k * 32);
                                const uint64_t db = make_sw128_desc(b_addr + k * 32);
                                umma_f16_2sm(tmem_base + h * 256, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
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
    } else if (warp < 2 + kExpWarps) {
        // ===== weight expansion: thread = one of this CTA's 128 weight rows; 4 teams on every 4th k-block =====
        const int et = threadIdx.x - 64;
        const int team = et >> 7;              // 0..3
        const int e = et & 127;                // weight row within this CTA's 128-row half (= plane-tile row)
        const int ew = et >> 5;                // expansion warp (own scratch)
        const int rgi = e >> 5;                // row group within the plane tile (warp-uniform)
        const uint32_t r7 = (uint32_t)(e & 7);
        const uint32_t row_off = (uint32_t)(e >> 3) * 1024u + r7 * 128u;
        const uint32_t scratch = smem_base + kOffScratch + ew * kScratchBytes;
        const bool grouped = p.groups > 1;
        const uint32_t full_b_leader0 = mapa_rank0(full_b(0));

        struct Meta { uint4 pw; uint32_t cs, ce; int tr, g; };
        int c_ti = 0, c_kb = team;             // prefetch cursor over this team's items
        auto cursor_norm = [&]() {
            while (c_kb >= KB && c_ti < my_tiles) { c_kb -= KB; ++c_ti; }
        };
        auto load_meta = [&]() {
            Meta m;
            m.pw = make_uint4(0, 0, 0, 0);
            m.cs = m.ce = 0;
            m.tr = -1; m.g = 0;
            cursor_norm();
            if (c_ti < my_tiles) {
                const int t = cluster_id + c_ti * num_clusters;
                const int tr = (t % p.n_tiles) * 2 + (int)crank;     // this CTA's 128-row plane tile
                if (tr < p.tiles_r) {
                    const int64_t tile = (int64_t)tr * p.tiles_c + c_kb;
                    m.pw = __ldg(p.planes + tile * kTileRows + e);
                    m.cs = __ldg(p.vptr + tile * kRgPerTile + rgi);
                    m.ce = __ldg(p.vptr + tile * kRgPerTile + rgi + 1);
                    m.tr = tr;
                    m.g = grouped ? c_kb / p.tiles_per_group : 0;
                }
                c_kb += kTeams;
            }
            return m;
        };
        auto load_vals = [&](const Meta& m, uint4& q0, uint4& q1) {
            const uint32_t b0 = (m.cs * 2u) & ~15u, b1 = m.ce * 2u;
            const uint8_t* base = reinterpret_cast<const uint8_t*>(p.vals);
            const uint32_t o0 = b0 + 16u * lane, o1 = o0 + 512u;
            if (o0 < b1) q0 = __ldg(reinterpret_cast<const uint4*>(base + o0));
            if (o1 < b1) q1 = __ldg(reinterpret_cast<const uint4*>(base + o1));
        };

        const int64_t total = (int64_t)my_tiles * KB;
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
            const Meta m2 = load_meta();
            if (it + 1 < my_items) load_vals(m1, q0, q1);
            m0 = m1;
            m1 = m2;

            if (g != cur_g || tr != cur_tr) {
                cur_g = g; cur_tr = tr;
                float2 a = make_float2(0.f, 0.f);
                if (tr >= 0) a = __ldg(p.affine + ((int64_t)tr * kTileRows + e) * p.groups + g);
                const uint32_t lo = bits16<T>(a.x), hi = bits16<T>(a.y);
                LL = lo | (lo << 16);
                DD = (lo ^ hi) * 0x10001u;
            }

            mbar_wait(empty(s), ph ^ 1u);

            const uint32_t brow = smem_base + kOffB + s * kBStage + row_off;
            expand_row(pw, LL, DD, brow, r7, cs, ce, v0, v1, scratch, p.vals, (uint32_t)lane);
            fence_proxy_async();          // this thread's generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(full_b_leader0 + 8u * s);   // one arrive per warp on the LEADER's barrier
            s += kTeams;
            if (s >= kStages) { s -= kStages; ph ^= 1u; }
        }
    } else {
        // ===== epilogue: this CTA's 256 tokens x 256 weight rows (TMEM -> regs -> +bias -> 16 bit -> global) =====
        const int q = warp & 3;
        uint32_t acc_ph = 0;
        T* y = reinterpret_cast<T*>(p.y);
        const uint32_t tmem_empty_leader = mapa_rank0(tmem_empty);
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int t = cluster_id + ti * num_clusters;
            const int mp = (t / p.n_tiles) * p.bm;
            const int m0 = mp + (int)crank * (p.bm >> 1), n0 = (t % p.n_tiles) * BN;
            const int halves = (p.bm == 2 * BMC && mp + 128 < p.M) ? 2 : 1;      // which accumulators the MMA warp wrote
            mbar_wait(tmem_full, acc_ph);
            tc_fence_after();
            for (int h = 0; h < halves; ++h) {
                const int m = m0 + h * 128 + q * 32 + lane;
                if (m0 + h * 128 >= p.M) break;                // warp-uniform: no valid token in this half
#pragma unroll 1
                for (int cb = 0; cb < BN / 32; ++cb) {
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
    cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still touch this CTA's smem / barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------
bool gemm_tc2_enabled(const Layer& L, int64_t M) {
    // PBL_GEMM_2CTA: 0 = never, 1 = when profitable (default), 2 = whenever the tcgen05 path is taken
    const char* e = getenv("PBL_GEMM_2CTA");
    const int mode = (e && *e) ? atoi(e) : 1;
    (void)L;
    if (mode == 0) return false;
    if (mode == 2) return true;
    return M > 256;
}

int launch_gemm_tc2(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable"); return PBL_ERR_CUDA; }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // pair tile = 512 tokens (each CTA two UMMA halves) when that fills the GPU, else 256 tokens (one half per
    // CTA): twice the tiles, and the per-CTA weight expansion still amortised over a 256-token pair MMA.
    const int n_tiles = (int)((L.N + tc2::BN - 1) / tc2::BN);
    const int bm = (((M + 511) / 512) * (int64_t)n_tiles >= num_sms / 2) ? 512 : 256;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)L.K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)tc2::BK, (cuuint32_t)(bm / 2)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, L.dtype == PBL_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return PBL_ERR_CUDA; }

    GemmParams p;
    p.planes = L.planes; p.vptr = L.vptr; p.vals = reinterpret_cast<const uint16_t*>(L.vals); p.affine = L.affine;
    p.bias = L.bias; p.y = y; p.ldy = ldy; p.M = (int)M; p.N = (int)L.N; p.K = (int)L.K;
    p.tiles_r = (int)L.tiles_r; p.tiles_c = (int)L.tiles_c; p.groups = (int)L.groups; p.tiles_per_group = L.tiles_per_group;
    p.bm = bm;
    p.m_tiles = (int)((M + p.bm - 1) / p.bm);
    p.n_tiles = n_tiles;
    p.kblocks = (int)L.tiles_c;

    static bool attr_set_dev[2][64] = {};   // function attributes are per device
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool* attr_set = nullptr;
    bool attr_local[2] = {false, false};
    attr_set = (cur_dev >= 0 && cur_dev < 64) ? nullptr : attr_local;
    const int which = L.dtype == PBL_F16 ? 0 : 1;
    auto kern = which == 0 ? gemm_tc2_kernel<__half> : gemm_tc2_kernel<__nv_bfloat16>;
    bool& attr_done = attr_set ? attr_set[which] : attr_set_dev[which][cur_dev];
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemBytes),
                            "cudaFuncSetAttribute(smem, 2cta)");
        if (rc) return rc;
        attr_done = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int max_clusters = num_sms / 2;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    kern<<<2 * clusters, tc2::kThreads, tc2::kSmemBytes, s>>>(tmap, p);   // __cluster_dims__(2,1,1) pairs adjacent CTAs
    count_launch();
    return check_cuda(cudaGetLastError(), "gemm_tc2 launch");
}

}  // namespace pbl
