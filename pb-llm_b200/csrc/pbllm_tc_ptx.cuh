// Inline-PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, tcgen05.mma/ld/commit, PRMT/LOP3).
#pragma once
#include <cuda.h>

#include "pbllm_common.cuh"

namespace pbl {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t sel_xor_and(uint32_t a, uint32_t b, uint32_t c) {  // a ^ (b & c)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 64 16-bit elements = 128 B,
// 8-row groups 1024 B apart): start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32 | version(1)<<46 | layout(2)<<61
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

template <typename T> __device__ __forceinline__ uint32_t bits16(float v);
template <> __device__ __forceinline__ uint32_t bits16<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ uint32_t bits16<__nv_bfloat16>(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---- cluster / cta_group::2 PTX ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {   // same offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t local_addr) { return mapa_rank(local_addr, 0u); }
__device__ __forceinline__ float ld_cluster_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // release at cluster scope
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "DONE_C:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar_leader) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar_leader), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrive on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- the weight expansion shared by every tensor-core kernel ------------------------------------------------
// Rebuild ONE weight row (64 exact 16-bit values = 128 B) of a 128B-swizzled K-major tile in shared memory:
// sign bits -> byte-sign-replicating PRMT -> LOP3 select of the row's {lo,hi} pair (8 conflict-free STS.128),
// then the row's salient values are patched over their positions.  Called by all 32 lanes of a warp, lane =
// row of one 32-row group.
//   pw      {sign[0:32], sign[32:64], salient[0:32], salient[32:64]} of this row
//   LL, DD  {lo,lo} and {lo^hi, lo^hi} as packed 16-bit pairs
//   brow    shared address of the row (128 B aligned);  r7 = row & 7 (swizzle key)
//   cs, ce  the row group's value chunk [cs, ce) in `vals` (elements);  v0, v1 = this lane's prefetched 16 B
//           slots (lane, lane+32) of the chunk counted from its 16 B-aligned start
//   scratch the warp's private 1 KB staging buffer (512 values); values beyond it are read from `vals`
__device__ __forceinline__ void expand_row(const uint4 pw, const uint32_t LL, const uint32_t DD, const uint32_t brow,
                                           const uint32_t r7, const uint32_t cs, const uint32_t ce, const uint4 v0,
                                           const uint4 v1, const uint32_t scratch, const uint16_t* __restrict__ vals,
                                           const uint32_t lane) {
    const uint32_t b0 = (cs * 2u) & ~15u;
    const uint32_t npop = (uint32_t)(__popc(pw.z) + __popc(pw.w));
    const bool any_sal = ce != cs;                       // warp-uniform (chunk bounds are per row group)
    uint32_t idx0 = 0;
    if (any_sal) {
        __syncwarp();                                    // the previous item's scratch reads are finished
        const uint32_t o0 = 16u * lane, o1 = o0 + 512u;
        if (b0 + o0 < ce * 2u) sts_v4(scratch + o0, v0.x, v0.y, v0.z, v0.w);
        if (b0 + o1 < ce * 2u) sts_v4(scratch + o1, v1.x, v1.y, v1.z, v1.w);
        __syncwarp();                                    // staged values visible to every lane
    }
    // dense part: 64 bits -> 64 exact {lo,hi} values, 8 swizzled 16 B chunks
#pragma unroll
    for (int wd = 0; wd < 2; ++wd) {
        const uint32_t sg = wd ? pw.y : pw.x;
        const uint32_t X0 = sg, X1 = sg << 1, X2 = sg << 2, X3 = sg << 3, X4 = sg << 4, X5 = sg << 5, X6 = sg << 6, X7 = sg << 7;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t sel = 0x8888u | (uint32_t)c | ((uint32_t)c << 4) | ((uint32_t)(4 + c) << 8) | ((uint32_t)(4 + c) << 12);
            const uint32_t h0 = sel_xor_and(LL, DD, prmt(X7, X6, sel));
            const uint32_t h1 = sel_xor_and(LL, DD, prmt(X5, X4, sel));
            const uint32_t h2 = sel_xor_and(LL, DD, prmt(X3, X2, sel));
            const uint32_t h3 = sel_xor_and(LL, DD, prmt(X1, X0, sel));
            sts_v4(brow + ((((uint32_t)(wd * 4 + c)) ^ r7) << 4), h0, h1, h2, h3);
        }
    }
    if (!any_sal) return;
    idx0 = (cs - (b0 >> 1)) + warp_excl_scan(npop, lane);
    // salient part: patch the exact stored values over their positions (a 2-way unrolled loop measured slower)
    if (ce - (b0 >> 1) <= 512u) {                        // warp-uniform: the whole chunk is staged in scratch
        uint32_t sa = scratch + idx0 * 2u;
#pragma unroll
        for (int wd = 0; wd < 2; ++wd) {
            uint32_t rm = __brev(wd ? pw.w : pw.z);      // msb-first: clz gives the lowest column
            const uint32_t k1 = (r7 << 4) ^ (uint32_t)(wd * 64);
            while (rm) {
                const uint32_t j = (uint32_t)__clz(rm);
                rm &= ~(0x80000000u >> j);
                const uint16_t v = lds_u16(sa);
                sa += 2u;
                sts_u16(brow | ((j + j) ^ k1), v);
            }
        }
    } else {                                             // rare: very dense chunk, tail read from global
        uint32_t idx = idx0;
#pragma unroll
        for (int wd = 0; wd < 2; ++wd) {
            uint32_t mk = wd ? pw.w : pw.z;
            while (mk) {
                const uint32_t j = (uint32_t)__ffs(mk) - 1u;
                mk &= mk - 1u;
                uint16_t v;
                if (idx < 512u) v = lds_u16(scratch + idx * 2u);
                else v = __ldg(vals + (b0 >> 1) + idx);
                ++idx;
                sts_u16(brow + (((uint32_t)wd * 64u + (j << 1)) ^ (r7 << 4)), v);
            }
        }
    }
}

struct GemmParams {
    const uint4* planes;
    const uint32_t* vptr;
    const uint16_t* vals;
    const float2* affine;
    const float* bias;
    void* y;
    int64_t ldy;
    int M, N, K;
    int tiles_r, tiles_c, groups, tiles_per_group;
    int m_tiles, n_tiles, kblocks;
    int bm;  // tokens per CTA tile: 256 (two UMMA halves) or 128 (one half; more CTAs when M is small)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
int launch_gemm_tc2(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s);
bool gemm_tc2_enabled(const Layer& L, int64_t M);
int launch_gemm_twophase(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s);
bool gemm_twophase_enabled(const Layer& L, int64_t M);
int launch_gemm_splitk(const Layer& L, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, cudaStream_t s);
bool gemm_splitk_supported(const Layer& L, const void* x, int64_t ldx, const void* y, int64_t ldy, int64_t M);

}  // namespace pbl
