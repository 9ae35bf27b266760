/*
 * pbllm_oracle.c -- CPU restatement of the PB-LLM partially-binarized linear forward.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is a product path: it is the
 * checker that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs compare the CUDA path against (or time beside it).  The
 * product (pb-llm_b200/) never imports, links or executes this file and fails
 * loudly when its CUDA library is missing.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks every function below
 * against fixtures in tests/golden/ produced by executing the unmodified
 * reference modules (oracle/gen_golden.py, run in the build container where
 * /root/reference exists).
 *
 * Each function cites the reference file:line (relative to /root/reference) whose
 * arithmetic it restates.  All tensors are row-major float32 buffers.  In
 * "half_mode" the buffers hold fp16-representable values and every arithmetic
 * step is rounded to fp16 exactly where the reference's fp16 tensors round
 * (torch CPU half ops compute in fp32 and round the result of each op).
 *
 * Plain C, scalar, single thread, double accumulation for reductions.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline float h_round(float v) { return (float)(_Float16)v; } /* fp32 -> fp16 -> fp32, RNE */

/* quant/quantizer.py:18-21  STEBinary.forward: x.sign() in {-1, 0, +1} */
static inline float sgn(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

ORC_API void orc_sign(const float* w, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = sgn(w[i]);
}

/* quant/quantizer.py:84-85  BinaryLinear.forward: w = sign(W) */
ORC_API void orc_binary_wsim(const float* W, int64_t N, int64_t K, float* wsim) {
    orc_sign(W, wsim, N * K);
}

/* quant/quantizer.py:181-189  XnorBinaryLinear.quant_weight (outlier_mask=None):
 *   w = W - mean_row(W); alpha = mean_row(|w|); w_sim = sign(w) * alpha   (no mean add-back)
 * mu/alpha may be NULL. */
ORC_API void orc_xnor_wsim(const float* W, int64_t N, int64_t K, float* wsim, float* mu, float* alpha) {
    for (int64_t i = 0; i < N; ++i) {
        const float* r = W + i * K;
        double s = 0.0;
        for (int64_t j = 0; j < K; ++j) s += r[j];
        float m = (float)(s / (double)K);
        double a = 0.0;
        for (int64_t j = 0; j < K; ++j) a += fabsf(r[j] - m);
        float al = (float)(a / (double)K);
        for (int64_t j = 0; j < K; ++j) wsim[i * K + j] = sgn(r[j] - m) * al;
        if (mu) mu[i] = m;
        if (alpha) alpha[i] = al;
    }
}

/* float -> uint8 as the reference's CPU path does it (x86 truncating convert, then the
 * low 8 bits): round([-100.4,-1,255.6,300]).type(uint8) == [156,255,0,44]
 * (SURVEY.md fact 5).  quant/outlier_quantizer.py:18-20 */
static inline uint8_t wrap_u8(float v) { return (uint8_t)((int64_t)v & 0xFF); }

/* quant/outlier_quantizer.py:10-29  weight_quant_8bit(w, simulated):
 *   range = max_row - min_row (as float32); zp = round(min_row);
 *   q = uint8(round((w - zp) / range * 255)); clamp(0,255) is a no-op on uint8;
 *   simulated: q * (range / 255) + zp, cast back to the input dtype.
 * out_sim and out_codes may be NULL. */
ORC_API void orc_weight_quant_8bit(const float* W, int64_t N, int64_t K, float* out_sim,
                                   uint8_t* out_codes, int half_mode) {
    for (int64_t i = 0; i < N; ++i) {
        const float* r = W + i * K;
        float mx = r[0], mn = r[0];
        for (int64_t j = 1; j < K; ++j) {
            if (r[j] > mx) mx = r[j];
            if (r[j] < mn) mn = r[j];
        }
        float range = mx - mn;              /* :12-14, computed in the tensor dtype */
        if (half_mode) range = h_round(range);
        float zp = rintf(mn);               /* :16 torch.round = half-to-even */
        float step = range / 255.0f;        /* :24 (w_range / 255) is float32 in both modes */
        for (int64_t j = 0; j < K; ++j) {
            float d = r[j] - zp;            /* :18 in the tensor dtype */
            if (half_mode) d = h_round(d);
            float t = d / range;            /* promoted to float32 (range is float32, :15) */
            t = t * 255.0f;
            uint8_t q = wrap_u8(rintf(t));
            if (out_codes) out_codes[i * K + j] = q;
            if (out_sim) {
                float v = (float)q * step;
                v = v + zp;
                out_sim[i * K + j] = half_mode ? h_round(v) : v;
            }
        }
    }
}

static int cmp_f32(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* quant/outlier_quantizer.py:54-81  BinaryXnorExceptOutliersLinear.gen_outlier_mask:
 *   lower = kthvalue(flat, int(n*f/2)); upper = kthvalue(flat, int(n*(1-f/2)))   (1-indexed k)
 *   mask = (w < lower) | (w > upper)
 *   binary_scale = mean(|w[~mask]|)   -- ONE scalar, taken on the ORIGINAL weights (:72-74)
 *   weight <- weight_quant_8bit(w)    (:75)
 * Outputs: mask[N*K] (1 = salient), *binary_scale, w8[N*K] (8-bit fake-quant weight),
 * thr[2] = {lower, upper}.  Returns the salient count, or -1 if a k is out of range. */
ORC_API int64_t orc_outlier_gen_mask(const float* W, int64_t N, int64_t K, double frac, uint8_t* mask,
                                     float* binary_scale, float* w8, float* thr, int half_mode) {
    int64_t n = N * K;
    int64_t k_lo = (int64_t)((double)n * frac / 2.0);
    int64_t k_hi = (int64_t)((double)n * (1.0 - frac / 2.0));
    if (k_lo < 1 || k_hi < 1 || k_lo > n || k_hi > n) return -1;
    float* tmp = (float*)malloc(sizeof(float) * (size_t)n);
    memcpy(tmp, W, sizeof(float) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(float), cmp_f32);
    float lo = tmp[k_lo - 1], hi = tmp[k_hi - 1];
    free(tmp);
    int64_t cnt = 0, nn = 0;
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        uint8_t s = (W[i] < lo) || (W[i] > hi);
        mask[i] = s;
        cnt += s;
        if (!s) { acc += fabsf(W[i]); ++nn; }
    }
    float bs = (float)(acc / (double)nn);
    if (half_mode) bs = h_round(bs);
    *binary_scale = bs;
    if (thr) { thr[0] = lo; thr[1] = hi; }
    orc_weight_quant_8bit(W, N, K, w8, NULL, half_mode);
    return cnt;
}

/* quant/outlier_quantizer.py:83-99  binarize_except_outliers (eval mode, train_outlier=False):
 *   w_sim = where(mask, W8 * outlier_scale, sign(W8) * binary_scale) */
ORC_API void orc_outlier_wsim(const float* w8, const uint8_t* mask, float binary_scale, float outlier_scale,
                              int64_t N, int64_t K, float* wsim, int half_mode) {
    for (int64_t i = 0; i < N * K; ++i) {
        float sc = w8[i] * outlier_scale;
        float bn = sgn(w8[i]) * binary_scale;
        if (half_mode) { sc = h_round(sc); bn = h_round(bn); }
        wsim[i] = mask[i] ? sc : bn;
    }
}

/* quant/outlier_quantizer.py:90-93  training-mode recompute of the scalar scale from the
 * (already 8-bit) weights. */
ORC_API float orc_outlier_train_scale(const float* w8, const uint8_t* mask, int64_t n, int half_mode) {
    double acc = 0.0;
    int64_t nn = 0;
    for (int64_t i = 0; i < n; ++i)
        if (!mask[i]) { acc += fabsf(w8[i]); ++nn; }
    float bs = (float)(acc / (double)nn);
    return half_mode ? h_round(bs) : bs;
}

/* quant/outlier_quantizer.py:116-122  calc_memory_consumption: CSR of the uint8 codes
 * restricted to the mask -> (8 bit col + 8 bit value per stored entry + 8 bit per row
 * pointer) / numel.  to_sparse_csr() drops entries whose CODE is zero. */
ORC_API double orc_outlier_nbits(const float* w8_after, const uint8_t* mask, int64_t N, int64_t K, int half_mode) {
    uint8_t* codes = (uint8_t*)malloc((size_t)(N * K));
    orc_weight_quant_8bit(w8_after, N, K, NULL, codes, half_mode);
    int64_t nnz = 0;
    for (int64_t i = 0; i < N * K; ++i) nnz += (mask[i] && codes[i] != 0);
    free(codes);
    return ((double)nnz * 8.0 + (double)nnz * 8.0 + (double)(N + 1) * 8.0) / (double)(N * K);
}

/* torch.nn.functional.linear as called at quant/quantizer.py:86,193 and
 * quant/outlier_quantizer.py:105:  y[m,i] = sum_j x[m,j] * w[i,j] + b[i]
 * (double accumulation: the oracle is the more accurate side of every comparison). */
ORC_API void orc_linear(const float* x, const float* w, const float* bias, int64_t M, int64_t N, int64_t K,
                        float* y) {
    for (int64_t m = 0; m < M; ++m)
        for (int64_t i = 0; i < N; ++i) {
            const float* xr = x + m * K;
            const float* wr = w + i * K;
            double acc = 0.0;
            for (int64_t j = 0; j < K; ++j) acc += (double)xr[j] * (double)wr[j];
            if (bias) acc += (double)bias[i];
            y[m * N + i] = (float)acc;
        }
}

/* quant/quantizer.py:151-169  BiRealLinear.forward (inference value):
 *   input -> sign(x);  w = mean_row(|W|) * sign(W);  y = linear(sign(x), w)   -- NO bias (:168)
 * which equals alpha_i * (2*popcount(xnor(sign x, sign w)) - K) when no operand is 0. */
ORC_API void orc_bireal_forward(const float* x, const float* W, int64_t M, int64_t N, int64_t K, float* y) {
    float* al = (float*)malloc(sizeof(float) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) {
        double a = 0.0;
        for (int64_t j = 0; j < K; ++j) a += fabsf(W[i * K + j]);
        al[i] = (float)(a / (double)K);
    }
    for (int64_t m = 0; m < M; ++m)
        for (int64_t i = 0; i < N; ++i) {
            double acc = 0.0;
            for (int64_t j = 0; j < K; ++j) acc += (double)sgn(x[m * K + j]) * (double)(al[i] * sgn(W[i * K + j]));
            y[m * N + i] = (float)acc;
        }
    free(al);
}

/* ---- GPTQ-PB output format (input format of BASELINE configs 3-5) ------------------ */

/* gptq_pb/low_quant.py:25-32  LowQuantizer.calibrate, method "xnor", one column group:
 *   called with w = W[:, st:ed] * mask  (gptq_pb/gptq.py:103-105), so the mean / scale run
 *   over the masked-with-zeros row:  mean = mean_row(w); scale = mean_row(|w - mean|) */
ORC_API void orc_low_xnor_calibrate(const float* W, const uint8_t* low_mask, int64_t N, int64_t K, int64_t st,
                                    int64_t ed, float* mean, float* scale) {
    int64_t g = ed - st;
    for (int64_t i = 0; i < N; ++i) {
        double s = 0.0;
        for (int64_t j = st; j < ed; ++j) s += low_mask[i * K + j] ? W[i * K + j] : 0.0f;
        float m = (float)(s / (double)g);
        double a = 0.0;
        for (int64_t j = st; j < ed; ++j) a += fabsf((low_mask[i * K + j] ? W[i * K + j] : 0.0f) - m);
        mean[i] = m;
        scale[i] = (float)(a / (double)g);
    }
}

/* gptq_pb/low_quant.py:75-82  LowQuantizer.quantize "xnor":  q = mean + scale * sign(w - mean) */
static inline float low_xnor_q(float w, float mean, float scale) { return sgn(w - mean) * scale + mean; }

/* gptq_pb/high_quant.py:29-67  HighQuantizer.calibrate(weight=True, perchannel, sym=False, mse=False):
 *   xmin = min(min_row, 0); xmax = max(max_row, 0); (0,0) -> (-1,+1);
 *   scale = (xmax - xmin) / maxq; zero = round(-xmin / scale) */
ORC_API void orc_high_calibrate(const float* W, int64_t N, int64_t K, int bits, float* scale, float* zero) {
    float maxq = (float)((1 << bits) - 1);
    for (int64_t i = 0; i < N; ++i) {
        float mn = 0.0f, mx = 0.0f;
        for (int64_t j = 0; j < K; ++j) {
            float v = W[i * K + j];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mn == 0.0f && mx == 0.0f) { mn = -1.0f; mx = 1.0f; }
        scale[i] = (mx - mn) / maxq;
        zero[i] = rintf(-mn / scale[i]);
    }
}

/* gptq_pb/high_quant.py:6-8  quantize: scale * (clamp(round(x / scale) + zero, 0, maxq) - zero) */
static inline float high_q(float w, float scale, float zero, float maxq) {
    float q = rintf(w / scale) + zero;
    if (q < 0.0f) q = 0.0f;
    if (q > maxq) q = maxq;
    return scale * (q - zero);
}

/* gptq_pb/gptq.py:116-128 with disable_gptq=True (RTN) for one group layout:
 *   q = q_high * ~mask + q_low * mask  (mask True = binarized), then the result is cast to
 *   the layer dtype (gptq.py:180-184) -- fp16 when to_half.  groupsize<=0 means one group. */
ORC_API void orc_gptqpb_rtn(const float* W, const uint8_t* low_mask, int64_t N, int64_t K, int64_t groupsize,
                            int bits, float* out, float* lo, float* hi, int to_half) {
    if (groupsize <= 0) groupsize = K;
    int64_t G = (K + groupsize - 1) / groupsize;
    float* hs = (float*)malloc(sizeof(float) * (size_t)N);
    float* hz = (float*)malloc(sizeof(float) * (size_t)N);
    float* mean = (float*)malloc(sizeof(float) * (size_t)N);
    float* scale = (float*)malloc(sizeof(float) * (size_t)N);
    float maxq = (float)((1 << bits) - 1);
    orc_high_calibrate(W, N, K, bits, hs, hz);
    for (int64_t g = 0; g < G; ++g) {
        int64_t st = g * groupsize, ed = st + groupsize;
        if (ed > K) ed = K;
        orc_low_xnor_calibrate(W, low_mask, N, K, st, ed, mean, scale);
        for (int64_t i = 0; i < N; ++i) {
            for (int64_t j = st; j < ed; ++j) {
                float w = W[i * K + j];
                float v = low_mask[i * K + j] ? low_xnor_q(w, mean[i], scale[i]) : high_q(w, hs[i], hz[i], maxq);
                out[i * K + j] = to_half ? h_round(v) : v;
            }
            if (lo) {
                float a = -scale[i] + mean[i], b = scale[i] + mean[i];
                lo[i * G + g] = to_half ? h_round(a) : a;
                hi[i * G + g] = to_half ? h_round(b) : b;
            }
        }
    }
    free(hs); free(hz); free(mean); free(scale);
}
