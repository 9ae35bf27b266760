// GPTQ-PB calibration, the column loop of LowHighGPT.fasterquant (reference gptq_pb/gptq.py:116-168) as ONE kernel per
// 128-column block: the reference walks the block column by column from Python (about ten small launches per column:
// quantise, error, rank-1 update of the remaining columns); here a warp owns a weight row, keeps the block's 128 values
// of that row in registers (4 per lane), and runs the whole column recurrence without leaving the SM:
//     q    = mask ? low.quantize(w) : high.quantize(w)            (low_quant.py:75-82 "xnor", high_quant.py:6-8)
//     err  = (w - q) / d,   d = Hinv1[i][i]
//     W1[:, j] -= err * Hinv1[i][j]   for j >= i                    (gptq.py:162)
// with the block of the inverse-Hessian Cholesky factor staged once per CTA in shared memory.  Arithmetic follows the
// reference operation by operation in fp32 (separate multiply and subtract, round-half-even), so quantisation decisions
// match the reference's; the cross-block update W[:, col_ed:] -= Err1 @ Hinv[col_st:col_ed, col_ed:] (gptq.py:168) stays a
// library GEMM on the host side.  One-time work, not on the per-forward path.
#include "pbllm_common.cuh"

namespace pbl {

constexpr int kGqCols = 128;   // blocksize of the reference (gptq.py:54)
constexpr int kGqWarps = 8;

__global__ void __launch_bounds__(kGqWarps * 32)
gptq_block_kernel(float* __restrict__ W, int64_t ldw, float* __restrict__ Err, int64_t lde, const float* __restrict__ Hinv, int64_t ldh,
                  const uint8_t* __restrict__ mask, int64_t ldm, const float* __restrict__ lmean, const float* __restrict__ lscale,
                  const float* __restrict__ hscale, const float* __restrict__ hzero, float maxq, int64_t N, int nc,
                  float* __restrict__ losses) {
    extern __shared__ float sh[];                       // Hinv1 [nc][kGqCols] (columns past nc are zero)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int idx = threadIdx.x; idx < nc * kGqCols; idx += blockDim.x) {
        const int i = idx / kGqCols, j = idx - i * kGqCols;
        sh[idx] = j < nc ? Hinv[(int64_t)i * ldh + j] : 0.f;
    }
    __syncthreads();
    for (int64_t row = (int64_t)blockIdx.x * kGqWarps + wid; row < N; row += (int64_t)gridDim.x * kGqWarps) {
        float wv[4], qv[4], ev[4];
        uint32_t mv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = lane + 32 * c;
            wv[c] = j < nc ? W[row * ldw + j] : 0.f;
            mv[c] = j < nc ? (uint32_t)mask[row * ldm + j] : 0u;
            qv[c] = ev[c] = 0.f;
        }
        const float mean = lmean[row], lsc = lscale[row], hsc = hscale[row], hz = hzero[row];
        float loss = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll 4
            for (int l = 0; l < 32; ++l) {
                const int i = 32 * c + l;
                if (i >= nc) break;                                               // warp-uniform
                const float w = __shfl_sync(0xffffffffu, wv[c], l);
                const uint32_t m = __shfl_sync(0xffffffffu, mv[c], l);
                const float d = sh[i * kGqCols + i];
                // low: (w - mean).sign() * scale + mean          high: scale * (clamp(round(w / scale) + zero, 0, maxq) - zero)
                const float wm = __fsub_rn(w, mean);
                const float sg = wm > 0.f ? 1.f : (wm < 0.f ? -1.f : 0.f);
                const float q_low = __fadd_rn(__fmul_rn(sg, lsc), mean);
                const float qi = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(w, hsc)), hz), 0.f), maxq);
                const float q_high = __fmul_rn(hsc, __fsub_rn(qi, hz));
                const float q = m ? q_low : q_high;
                const float diff = __fsub_rn(w, q);
                const float err = __fdiv_rn(diff, d);
                loss += __fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(d, d));
                if (lane == l) { qv[c] = q; ev[c] = err; }
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) {                                  // W1[:, j] -= err * Hinv1[i][j], j >= i
                    const int j = lane + 32 * c2;
                    if (c2 >= c && j >= i) wv[c2] = __fsub_rn(wv[c2], __fmul_rn(err, sh[i * kGqCols + j]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = lane + 32 * c;
            if (j < nc) { W[row * ldw + j] = qv[c]; Err[row * lde + j] = ev[c]; }
        }
        if (lane == 0 && losses) losses[row] += 0.5f * loss;
    }
}

int launch_gptq_block(float* W, int64_t ldw, float* Err, int64_t lde, const float* Hinv, int64_t ldh, const uint8_t* mask, int64_t ldm,
                      const float* lmean, const float* lscale, const float* hscale, const float* hzero, float maxq, int64_t N, int nc,
                      float* losses, cudaStream_t s) {
    const int smem = nc * kGqCols * (int)sizeof(float);
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr[dev]) {
        int rc = check_cuda(cudaFuncSetAttribute(gptq_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGqCols * kGqCols * 4),
                            "cudaFuncSetAttribute(gptq smem)");
        if (rc) return rc;
        attr[dev] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (N + kGqWarps - 1) / kGqWarps;
    const unsigned grid = (unsigned)(want < (int64_t)sms * 3 ? want : (int64_t)sms * 3);
    gptq_block_kernel<<<grid, kGqWarps * 32, smem, s>>>(W, ldw, Err, lde, Hinv, ldh, mask, ldm, lmean, lscale, hscale, hzero, maxq, N, nc, losses);
    count_launch();
    return check_cuda(cudaGetLastError(), "gptq_block launch");
}

}  // namespace pbl
