/*
 * CPU oracle, plain C restatement of the FLEXS surrogate forward pass.
 * TEST INFRASTRUCTURE ONLY: built into oracle/_build/liboracle.so and loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
 * product library (flexs_b200/csrc) never links or calls this file.
 *
 * PARITY STATUS: parity unpinned for the floating-point results (the reference computes
 * them inside TensorFlow/Keras, which is absent here; see oracle/flexs_oracle.py header).
 *
 * Follows (paths relative to /root/reference):
 *   flexs/baselines/models/cnn.py:23-54   layer stack, Keras channels-last semantics
 *   flexs/baselines/models/mlp.py:21-31
 *   flexs/baselines/models/keras_model.py:77-79  squeeze + nan_to_num
 *   flexs/ensemble.py:54-59, :24          mean over members (fp32, divide by M)
 *   flexs/utils/sequence_utils.py:32-47   index = alphabet.index(ch)
 *
 * All arithmetic is fp32 with sequential accumulation in the order written, which is
 * one legitimate fp32 evaluation order (TF's own order is unspecified).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline float relu_f(float x) { return x > 0.f ? x : 0.f; }

/* np.nan_to_num on float32: NaN -> 0, +-inf -> +-FLT_MAX (keras_model.py:77). */
static inline float nan_to_num_f(float x)
{
    if (isnan(x)) return 0.f;
    if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
    return x;
}

/* sequence_utils.py:44-47.  Returns -1 on success, else the flat position of the first
 * character that is not in the alphabet (the reference raises ValueError there). */
int64_t oracle_encode(const char *chars, int64_t n, int len, const char *alphabet, int a,
                      uint8_t *out)
{
    int lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = -1;
    for (int i = a - 1; i >= 0; --i) lut[(unsigned char)alphabet[i]] = i; /* first index wins */
    for (int64_t p = 0; p < n * (int64_t)len; ++p) {
        int v = lut[(unsigned char)chars[p]];
        if (v < 0) return p;
        out[p] = (uint8_t)v;
    }
    return -1;
}

/* One sequence through the CNN.  Weight pointers are in Keras get_weights() order/layout:
 * w1 (k,A,F) b1 | w2 (k,F,F) b2 | w3 (k3,F,F) b3 | wd1 (F,H) bd1 | wd2 (H,H) bd2 | wd3 (H,1) bd3 */
static float cnn_one(const uint8_t *idx, int len, int a, int f, int h, int k, int k3,
                     const float *const *w, float *scratch)
{
    const int t = len - k + 1;
    float *h1 = scratch, *h2 = h1 + (size_t)t * f, *feat = h2 + (size_t)t * f;
    float *d1 = feat + f, *d2 = d1 + h;
    const float *w1 = w[0], *b1 = w[1], *w2 = w[2], *b2 = w[3], *w3 = w[4], *b3 = w[5];
    const float *wd1 = w[6], *bd1 = w[7], *wd2 = w[8], *bd2 = w[9], *wd3 = w[10], *bd3 = w[11];

    /* conv1, valid (cnn.py:25-32): the one-hot input picks one row of w1 per tap */
    for (int tt = 0; tt < t; ++tt)
        for (int o = 0; o < f; ++o) {
            float acc = 0.f;
            for (int j = 0; j < k; ++j) acc += w1[((size_t)j * a + idx[tt + j]) * f + o];
            h1[(size_t)tt * f + o] = relu_f(acc + b1[o]);
        }
    /* conv2, same (cnn.py:33-39): pad_left=(k-1)/2, remainder right; :40 MaxPooling1D(1)=id */
    const int pl2 = (k - 1) / 2;
    for (int tt = 0; tt < t; ++tt)
        for (int o = 0; o < f; ++o) {
            float acc = 0.f;
            for (int j = 0; j < k; ++j) {
                int s = tt + j - pl2;
                if (s < 0 || s >= t) continue;
                for (int g = 0; g < f; ++g)
                    acc += h1[(size_t)s * f + g] * w2[((size_t)j * f + g) * f + o];
            }
            h2[(size_t)tt * f + o] = relu_f(acc + b2[o]);
        }
    /* conv3, same, width k3 = A-1 (cnn.py:41-47), fused with GlobalMaxPooling1D (:48) */
    const int pl3 = (k3 - 1) / 2;
    for (int o = 0; o < f; ++o) feat[o] = -INFINITY;
    for (int tt = 0; tt < t; ++tt)
        for (int o = 0; o < f; ++o) {
            float acc = 0.f;
            for (int j = 0; j < k3; ++j) {
                int s = tt + j - pl3;
                if (s < 0 || s >= t) continue;
                for (int g = 0; g < f; ++g)
                    acc += h2[(size_t)s * f + g] * w3[((size_t)j * f + g) * f + o];
            }
            float v = relu_f(acc + b3[o]);
            if (v > feat[o]) feat[o] = v;
        }
    /* dense head (cnn.py:49-52); Dropout(:51) is the identity at predict time */
    for (int o = 0; o < h; ++o) {
        float acc = 0.f;
        for (int g = 0; g < f; ++g) acc += feat[g] * wd1[(size_t)g * h + o];
        d1[o] = relu_f(acc + bd1[o]);
    }
    for (int o = 0; o < h; ++o) {
        float acc = 0.f;
        for (int g = 0; g < h; ++g) acc += d1[g] * wd2[(size_t)g * h + o];
        d2[o] = relu_f(acc + bd2[o]);
    }
    float acc = 0.f;
    for (int g = 0; g < h; ++g) acc += d2[g] * wd3[g];
    return acc + bd3[0];
}

static float mlp_one(const uint8_t *idx, int len, int a, int h, const float *const *w,
                     float *scratch)
{
    float *x1 = scratch, *x2 = x1 + h;
    const float *w1 = w[0], *b1 = w[1], *w2 = w[2], *b2 = w[3], *w3 = w[4], *b3 = w[5];
    const float *w4 = w[6], *b4 = w[7];
    /* Flatten is row-major l*A + c (mlp.py:23), so layer 1 is a gather of L rows of w1 */
    for (int o = 0; o < h; ++o) {
        float acc = 0.f;
        for (int l = 0; l < len; ++l) acc += w1[((size_t)l * a + idx[l]) * h + o];
        x1[o] = relu_f(acc + b1[o]);
    }
    for (int o = 0; o < h; ++o) {
        float acc = 0.f;
        for (int g = 0; g < h; ++g) acc += x1[g] * w2[(size_t)g * h + o];
        x2[o] = relu_f(acc + b2[o]);
    }
    for (int o = 0; o < h; ++o) {
        float acc = 0.f;
        for (int g = 0; g < h; ++g) acc += x2[g] * w3[(size_t)g * h + o];
        x1[o] = relu_f(acc + b3[o]);
    }
    float acc = 0.f;
    for (int g = 0; g < h; ++g) acc += x1[g] * w4[g];
    return acc + b4[0];
}

/* Ensemble of m CNNs: weights[12*m].  Output = nan_to_num(member) averaged like
 * np.mean(axis=1) on float32: pairwise-free left-to-right sum, then divide by m. */
void oracle_cnn_forward(const uint8_t *idx, int64_t n, int len, int a, int f, int h, int k,
                        int m, const float *const *weights, float *out)
{
    const int k3 = a - 1;
    const int t = len - k + 1;
    const size_t scratch_n = (size_t)2 * t * f + f + 2 * (size_t)h;
#pragma omp parallel
    {
        float *scratch = (float *)malloc(scratch_n * sizeof(float));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            float sum = 0.f;
            for (int mm = 0; mm < m; ++mm)
                sum += nan_to_num_f(cnn_one(idx + i * len, len, a, f, h, k, k3,
                                            weights + 12 * mm, scratch));
            out[i] = (m == 1) ? sum : sum / (float)m;
        }
        free(scratch);
    }
}

void oracle_mlp_forward(const uint8_t *idx, int64_t n, int len, int a, int h, int m,
                        const float *const *weights, float *out)
{
#pragma omp parallel
    {
        float *scratch = (float *)malloc((size_t)2 * h * sizeof(float));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            float sum = 0.f;
            for (int mm = 0; mm < m; ++mm)
                sum += nan_to_num_f(mlp_one(idx + i * len, len, a, h, weights + 8 * mm, scratch));
            out[i] = (m == 1) ? sum : sum / (float)m;
        }
        free(scratch);
    }
}
