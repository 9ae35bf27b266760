// tcgen05 bit-plane GEMM (prefill regime): y[M,N] = x[M,K] . w_sim[N,K]^T + bias, fp16/bf16.
//
// The weight operand never exists densely in HBM: per 64-column k-block each CTA streams the
// packed planes (0.25 B/weight) + salient values, and 8 "expansion" warps rebuild the EXACT
// fp16/bf16 w_sim tile (bit -> {lo,hi} select per row, salient values patched in) directly in
// shared memory in the 128B-swizzled K-major layout tcgen05.mma consumes.  Because the tile is
// bit-identical to the reference's materialised w_sim, the result differs from the reference's
// cuBLAS F.linear (quant/quantizer.py:193, quant/outlier_quantizer.py:105) only by fp32
// summation order.  Activations come in by TMA (SWIZZLE_128B), accumulators live in TMEM.
//
// CTA tile: 256 tokens (two UMMA M=128 halves) x 256 weight rows (N=256) x 64 (K block);
// all 512 TMEM columns hold the two fp32 accumulator halves.  Persistent grid, one CTA per SM.
// Warp roles: 0 = TMA producer (x), 1 = MMA issuer + TMEM owner, 2..17 = weight expansion
// (two teams of 8 warps on alternate k-blocks; thread = weight row), 18..21 = epilogue
// (TMEM -> regs -> +bias -> fp16 -> global).
#include "pbllm_tc_ptx.cuh"

namespace pbl {

namespace tc {
constexpr int BM = 256, BN = 256, BK = 64;
constexpr int kStages = 3;
constexpr int kAStage = BM * BK * 2;  // 32 KB
constexpr int kBStage = BN * BK * 2;  // 32 KB
constexpr int kTeams = 2;                                     // expansion teams (alternate k-blocks)
constexpr int kExpWarps = 8 * kTeams, kEpiWarps = 4;
constexpr int kExpThreads = 256;                              // arrivals per B stage (one team)
static_assert(kTeams <= kStages, "teams must not outnumber stages");
constexpr int kThreads = (2 + kExpWarps + kEpiWarps) * 32;  // 704
constexpr int kScratchVals = 512;                            // prefetched salient values per (row-group, k-block)
constexpr int kScratchBytes = kScratchVals * 2;              // two 512 B rows of 16 B lane slots
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kStages * kAStage;
constexpr int kOffScratch = kOffB + kStages * kBStage;
constexpr int kOffBar = kOffScratch + kExpWarps * kScratchBytes;
constexpr int kNumBars = 3 * kStages + 2;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;  // + slack for 1024 B alignment of the base
static_assert(kOffBar % 8 == 0, "barrier alignment");
static_assert(kSmemBytes <= 232448, "exceeds 227 KB");
}  // namespace tc

template <typename T>
__global__ void __launch_bounds__(tc::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024 B alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_base + kOffBar;
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto full_b = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    const uint32_t tmem_full = bar0 + 8u * (3 * kStages), tmem_empty = bar0 + 8u * (3 * kStages + 1);
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemPtr);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_a(s), 1);
            mbar_init(full_b(s), kExpThreads);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, kEpiWarps * 32);
        fence_barrier_init();
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 128x256 fp32 accumulator halves)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kOffTmemPtr), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_tiles = p.m_tiles * p.n_tiles;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int KB = p.kblocks;

    if (warp == 0) {
        // ===== TMA producer: x tile [256 tokens x 64] per k-block =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int t = blockIdx.x + ti * gridDim.x;
                const int m0 = (t / p.n_tiles) * p.bm;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_a(s), (uint32_t)p.bm * 128u);
                    tma_load_2d(smem_base + kOffA + s * kAStage, &tmap_x, kb * BK, m0, full_a(s));
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D=f32, A/B = f16|bf16, both K-major, N=256, M=128
        const uint32_t fmt = (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
        int s = 0;
        uint32_t ph = 0, acc_ph = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int t = blockIdx.x + ti * gridDim.x;
            const int m0 = (t / p.n_tiles) * p.bm;
            const int halves = (p.bm == 256 && m0 + 128 < p.M) ? 2 : 1;
            mbar_wait(tmem_empty, acc_ph ^ 1u);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_base + kOffA + s * kAStage, b_addr = smem_base + kOffB + s * kBStage;
                    for (int h = 0; h < halves; ++h) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da = make_sw128_desc(a_addr + h * (128 * 128) + k * 32);
                            const uint64_t db = make_sw128_desc(b_addr + k * 32);
                            umma_f16(tmem_base + h * 256, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(empty(s));
                    if (kb == KB - 1) umma_commit(tmem_full);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
            acc_ph ^= 1u;
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

        // item = (tile index ti, k-block kb); this team handles items team, team+kTeams, ... of the
        // CTA's flattened (tile, k-block) sequence.  Cursor arithmetic is incremental (no divisions).
        struct Meta { uint4 pw; uint32_t cs, ce; int tr, g; };
        int c_ti = 0, c_kb = team;             // prefetch cursor
        auto cursor_norm = [&]() {
            while (c_kb >= KB && c_ti < my_tiles) { c_kb -= KB; ++c_ti; }
        };
        auto load_meta = [&]() {               // loads the cursor's item and advances the cursor
            Meta m;
            m.pw = make_uint4(0, 0, 0, 0);
            m.cs = m.ce = 0;
            m.tr = -1; m.g = 0;
            cursor_norm();
            if (c_ti < my_tiles) {
                const int t = blockIdx.x + c_ti * gridDim.x;
                const int tr = (t % p.n_tiles) * 2 + (e >> 7);
                if (tr < p.tiles_r) {
                    const int64_t tile = (int64_t)tr * p.tiles_c + c_kb;
                    m.pw = __ldg(p.planes + tile * kTileRows + r);
                    m.cs = __ldg(p.vptr + tile * kRgPerTile + rgi);
                    m.ce = __ldg(p.vptr + tile * kRgPerTile + rgi + 1);
                    m.tr = tr;
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
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias) -> 16-bit -> global =====
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        uint32_t acc_ph = 0;
        T* y = reinterpret_cast<T*>(p.y);
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int t = blockIdx.x + ti * gridDim.x;
            const int m0 = (t / p.n_tiles) * p.bm, n0 = (t % p.n_tiles) * BN;
            const int halves = (p.bm == 256 && m0 + 128 < p.M) ? 2 : 1;
            mbar_wait(tmem_full, acc_ph);
            tc_fence_after();
            for (int h = 0; h < halves; ++h) {
                const int m = m0 + h * 128 + q * 32 + lane;
#pragma unroll 1
                for (int cb = 0; cb < BN / 32; ++cb) {
                    const int n = n0 + cb * 32;
                    if (n >= p.N) break;  // warp-uniform
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
            mbar_arrive(tmem_empty);
            acc_ph ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------
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

bool gemm_tc_supported(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M) {
    if (L.dtype != PBL_F16 && L.dtype != PBL_BF16) return false;
    if (M <= 0 || M > (1 << 30)) return false;
    if (L.N % 8 != 0 || ldy % 8 != 0 || ldx % 8 != 0) return false;
    if (x && (reinterpret_cast<uintptr_t>(x) & 15u)) return false;
    if (y && (reinterpret_cast<uintptr_t>(y) & 15u)) return false;
    if (L.bias && (reinterpret_cast<uintptr_t>(L.bias) & 15u)) return false;
    if (L.groups > 1 && L.groupsize % tc::BK != 0) return false;
    return true;
}

int launch_gemm_tc(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s) {
    if (gemm_twophase_enabled(L, M)) return launch_gemm_twophase(L, x, ldx, y, ldy, M, s);
    if (gemm_tc2_enabled(L, M)) return launch_gemm_tc2(L, x, ldx, y, ldy, M, s);
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable"); return PBL_ERR_CUDA; }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int n_tiles = (int)((L.N + tc::BN - 1) / tc::BN);
    // 256-token tiles amortise the weight expansion over two MMA halves; when that leaves SMs idle
    // (small M), fall back to 128-token tiles to double the CTA count.
    const int bm = (M <= 128 || ((M + 255) / 256) * (int64_t)n_tiles < num_sms) ? 128 : 256;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)L.K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)bm};
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
    p.m_tiles = (int)((M + bm - 1) / bm);
    p.n_tiles = n_tiles;
    p.kblocks = (int)L.tiles_c;

    static bool attr_set_dev[2][64] = {};   // function attributes are per device
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool* attr_set = nullptr;
    bool attr_local[2] = {false, false};
    attr_set = (cur_dev >= 0 && cur_dev < 64) ? nullptr : attr_local;
    const int which = L.dtype == PBL_F16 ? 0 : 1;
    auto kern = which == 0 ? gemm_tc_kernel<__half> : gemm_tc_kernel<__nv_bfloat16>;
    bool& attr_done = attr_set ? attr_set[which] : attr_set_dev[which][cur_dev];
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes),
                            "cudaFuncSetAttribute(smem)");
        if (rc) return rc;
        attr_done = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    kern<<<grid, tc::kThreads, tc::kSmemBytes, s>>>(tmap, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "gemm_tc launch");
}

}  // namespace pbl
