/*
 * Plain-C restatement of the ACCEL-v1 BSR-INT8 golden arithmetic.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, smoke() and bench.py's CPU-baseline
 * leg (through oracle/c_oracle.py).  The product never links or calls this.
 *
 * Parity status: PINNED (tests/test_oracle_golden.py checks this library against
 * tests/golden/ fixtures produced by the reference's own Python / C++ code and
 * against oracle/_ref when that is built).
 *
 * Reference lines followed (relative to the reference root):
 *   orc_bsr_gemm_i32        sw/golden/golden_fc1_test.py:78-106 (Convention B)
 *   orc_bsr_gemm_i32_conva  hw/sim/cpp/src/golden_models.cpp:187-255 (Convention A)
 *   orc_im2col_chw          hw/sim/cpp/src/golden_models.cpp:801-842
 *   orc_requant_channel     hw/sim/cpp/src/golden_models.cpp:378-411 (factor per channel)
 *   orc_add_residual        hw/sim/cpp/src/golden_models.cpp:465-490
 *   orc_maxpool / orc_avgpool  hw/sim/cpp/src/golden_models.cpp:534-571 / 601-628
 *
 * Build: see oracle/Makefile (gcc -O2 -fPIC -shared, no -ffast-math, no FMA contraction).
 */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* Y[m, br*b + h] = sum over stored blocks j of row br, sum_w X[m, col_idx[j]*b + w] * blk[j][h][w].
 * X is [M, K] with leading dimension lda; a K-tile that runs past K is truncated.
 * Y is int32 [M, nbr*b] with leading dimension ldy.  Accumulation wraps (uint32 arithmetic). */
ORC_API void orc_bsr_gemm_i32(const int8_t *X, int64_t M, int64_t K, int64_t lda,
                              const int32_t *row_ptr, const int32_t *col_idx, const int8_t *blocks,
                              int32_t nbr, int32_t b, int32_t *Y, int64_t ldy)
{
    for (int64_t m = 0; m < M; ++m) {
        const int8_t *x = X + m * lda;
        for (int32_t br = 0; br < nbr; ++br) {
            uint32_t acc[64];
            for (int h = 0; h < b; ++h) acc[h] = 0u;
            for (int32_t j = row_ptr[br]; j < row_ptr[br + 1]; ++j) {
                int64_t k0 = (int64_t)col_idx[j] * b;
                int w_lim = (int)((K - k0) < b ? (K - k0) : b);
                const int8_t *blk = blocks + (int64_t)j * b * b;
                for (int h = 0; h < b; ++h) {
                    int32_t s = 0;
                    for (int w = 0; w < w_lim; ++w) s += (int32_t)x[k0 + w] * (int32_t)blk[h * b + w];
                    acc[h] += (uint32_t)s;
                }
            }
            for (int h = 0; h < b; ++h) Y[m * ldy + (int64_t)br * b + h] = (int32_t)acc[h];
        }
    }
}

/* Convention A: B[K, N] in BSR with block-rows over K, col_idx = N-tile. C is [M, N], zeroed first. */
ORC_API void orc_bsr_gemm_i32_conva(const int8_t *A, int64_t M, int64_t K, int64_t N,
                                    const int32_t *row_ptr, const int32_t *col_idx, const int8_t *blocks,
                                    int32_t nbr, int32_t b, int32_t *C)
{
    memset(C, 0, (size_t)(M * N) * sizeof(int32_t));
    for (int32_t br = 0; br < nbr; ++br)
        for (int32_t j = row_ptr[br]; j < row_ptr[br + 1]; ++j) {
            const int8_t *blk = blocks + (int64_t)j * b * b;
            for (int64_t m = 0; m < M; ++m)
                for (int jj = 0; jj < b; ++jj) {
                    int64_t n = (int64_t)col_idx[j] * b + jj;
                    if (n >= N) continue;
                    int32_t s = 0;
                    for (int i = 0; i < b; ++i) {
                        int64_t k = (int64_t)br * b + i;
                        if (k >= K) continue;
                        s += (int32_t)A[m * K + k] * (int32_t)blk[i * b + jj];
                    }
                    C[m * N + n] = (int32_t)((uint32_t)C[m * N + n] + (uint32_t)s);
                }
        }
}

/* CHW image -> patch matrix, TRANSPOSED relative to the reference's col buffer so that it can
 * feed orc_bsr_gemm_i32 directly: P[(oh*Wo+ow), (c*k*k + kh*k + kw)], leading dimension ldp. */
ORC_API void orc_im2col_chw(const int8_t *x, int32_t C, int32_t H, int32_t W, int32_t k, int32_t stride,
                            int32_t pad, int8_t *P, int64_t ldp)
{
    int32_t Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    for (int32_t oh = 0; oh < Ho; ++oh)
        for (int32_t ow = 0; ow < Wo; ++ow) {
            int8_t *row = P + ((int64_t)oh * Wo + ow) * ldp;
            int64_t kk = 0;
            for (int32_t c = 0; c < C; ++c)
                for (int32_t kh = 0; kh < k; ++kh)
                    for (int32_t kw = 0; kw < k; ++kw, ++kk) {
                        int32_t ih = oh * stride + kh - pad, iw = ow * stride + kw - pad;
                        row[kk] = (ih >= 0 && ih < H && iw >= 0 && iw < W)
                                      ? x[((int64_t)c * H + ih) * W + iw] : (int8_t)0;
                    }
        }
}

static inline int8_t sat8(int32_t v, uint64_t *sat)
{
    if (v > 127) { if (sat) ++*sat; return 127; }
    if (v < -128) { if (sat) ++*sat; return -128; }
    return (int8_t)v;
}

/* acc is [n_outer, n_chan, n_inner] int32; out same shape int8.
 * y = nearbyint((float)(relu? max(acc+bias,0) : acc+bias) * sf[c]); saturate.  Returns #saturated. */
ORC_API uint64_t orc_requant_channel(const int32_t *acc, int8_t *out, int64_t n_outer, int64_t n_chan,
                                     int64_t n_inner, const float *sf, const int32_t *bias, int32_t relu)
{
    uint64_t sat = 0;
    fesetround(FE_TONEAREST);
    for (int64_t o = 0; o < n_outer; ++o)
        for (int64_t c = 0; c < n_chan; ++c) {
            volatile float f = sf[c];
            for (int64_t i = 0; i < n_inner; ++i) {
                int64_t idx = (o * n_chan + c) * n_inner + i;
                int32_t a = acc[idx];
                if (bias) a = (int32_t)((uint32_t)a + (uint32_t)bias[c]);
                if (relu && a < 0) a = 0;
                volatile float scaled = (float)a * f;
                out[idx] = sat8((int32_t)nearbyintf(scaled), &sat);
            }
        }
    return sat;
}

ORC_API void orc_add_residual(const int8_t *main_, const int8_t *res, int8_t *out, int64_t n,
                              float s_main, float s_res, float s_out)
{
    fesetround(FE_TONEAREST);
    for (int64_t i = 0; i < n; ++i) {
        volatile float a = (float)main_[i] * s_main;
        volatile float b = (float)res[i] * s_res;
        volatile float s = a + b;
        volatile float q = s / s_out;
        out[i] = sat8((int32_t)nearbyintf(q), NULL);
    }
}

/* planes: [n_planes, H, W] int8; window pool x pool, stride, symmetric pad filled with -128. */
ORC_API void orc_maxpool(const int8_t *x, int8_t *out, int64_t n_planes, int32_t H, int32_t W,
                         int32_t pool, int32_t stride, int32_t pad)
{
    int32_t Ho = (H + 2 * pad - pool) / stride + 1, Wo = (W + 2 * pad - pool) / stride + 1;
    for (int64_t p = 0; p < n_planes; ++p)
        for (int32_t oh = 0; oh < Ho; ++oh)
            for (int32_t ow = 0; ow < Wo; ++ow) {
                int8_t best = -128;
                for (int32_t ph = 0; ph < pool; ++ph)
                    for (int32_t pw = 0; pw < pool; ++pw) {
                        int32_t ih = oh * stride + ph - pad, iw = ow * stride + pw - pad;
                        if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                        int8_t v = x[(p * H + ih) * W + iw];
                        if (v > best) best = v;
                    }
                out[(p * Ho + oh) * Wo + ow] = best;
            }
}

ORC_API void orc_avgpool(const int8_t *x, int8_t *out, int64_t n_planes, int32_t H, int32_t W)
{
    int32_t hw = H * W;
    for (int64_t p = 0; p < n_planes; ++p) {
        int32_t sum = 0;
        for (int32_t i = 0; i < hw; ++i) sum += x[p * hw + i];
        int32_t avg = (sum + hw / 2) / hw; /* C division: truncates toward zero */
        out[p] = sat8(avg, NULL);
    }
}

/* One whole BSR conv layer over a batch of CHW images (the CPU-baseline "port" leg):
 * im2col -> Convention-B BSR GEMM -> +bias -> ReLU -> per-channel requant -> optional residual.
 * x [B, C, H, W] int8; out [B, cout, Ho, Wo] int8.  Returns #saturated at the requant. */
ORC_API uint64_t orc_conv_bsr_layer(const int8_t *x, int32_t B, int32_t C, int32_t H, int32_t W,
                                    int32_t k, int32_t stride, int32_t pad,
                                    const int32_t *row_ptr, const int32_t *col_idx, const int8_t *blocks,
                                    int32_t nbr, int32_t b, int32_t cout,
                                    const int32_t *bias, int32_t relu, const float *sf,
                                    const int8_t *residual, float s_main, float s_res, float s_out,
                                    int8_t *out)
{
    int32_t Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    int64_t P = (int64_t)Ho * Wo, K = (int64_t)C * k * k, Np = (int64_t)nbr * b;
    int8_t *patches = (int8_t *)malloc((size_t)(P * K));
    int32_t *y = (int32_t *)malloc((size_t)(P * Np) * sizeof(int32_t));
    int32_t *yt = (int32_t *)malloc((size_t)(cout * P) * sizeof(int32_t));
    uint64_t sat = 0;
    for (int32_t n = 0; n < B; ++n) {
        orc_im2col_chw(x + (int64_t)n * C * H * W, C, H, W, k, stride, pad, patches, K);
        orc_bsr_gemm_i32(patches, P, K, K, row_ptr, col_idx, blocks, nbr, b, y, Np);
        for (int64_t p = 0; p < P; ++p)
            for (int32_t c = 0; c < cout; ++c) yt[c * P + p] = y[p * Np + c];
        int8_t *o = out + (int64_t)n * cout * P;
        sat += orc_requant_channel(yt, o, 1, cout, P, sf, bias, relu);
        if (residual) orc_add_residual(o, residual + (int64_t)n * cout * P, o, cout * P, s_main, s_res, s_out);
    }
    free(patches); free(y); free(yt);
    return sat;
}
