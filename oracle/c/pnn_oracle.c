/*
 * Plain-C restatement of the integer / gather / epilogue parts of the PNN path and of the
 * fully-connected forward pass.  TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* reference hevc/hm_common/c++/source_common/extraction_context.cpp:3-208 (same control flow) */
int oracle_extract_context_portions(const int* roi_origin, float* above, float* left, const uint8_t* flags,
                                    int n_avail, int unit_w, int unit_h, int above_units, int left_units,
                                    int w, int h, int stride, float mean) {
    if (!roi_origin || !above || !left || !flags) return -1;          /* :17-37 */
    if (n_avail <= 0) return -1;                                      /* :43-47 */
    const int total = above_units + left_units + 1, cw = 3 * w;
    const int* p;
    int i, j, k;
    if (n_avail == total) {                                           /* :56-90 */
        p = roi_origin - h * stride - w;
        for (i = 0; i < h; i++, p += stride)
            for (j = 0; j < cw; j++) above[i * cw + j] = (float)p[j] - mean;
        p = roi_origin - w;
        for (i = 0; i < 2 * h; i++, p += stride)
            for (j = 0; j < w; j++) left[i * w + j] = (float)p[j] - mean;
        return 0;
    }
    for (i = 0; i < h * cw; i++) above[i] = 0.f;                      /* :95-102 */
    for (i = 0; i < 2 * h * w; i++) left[i] = 0.f;
    p = roi_origin - h * stride - w;                                  /* :119-127 */
    for (i = 0; i < h; i++, p += stride)
        for (j = 0; j < w; j++) above[i * cw + j] = (float)p[j] - mean;
    if (!flags[left_units]) return -1;                                /* :133-138 */
    for (i = 0; i < above_units; i++) {                               /* :149-166 */
        if (!flags[left_units + 1 + i]) continue;
        float* d = above + w + i * unit_w;
        p = roi_origin - h * stride + i * unit_w;
        for (j = 0; j < h; j++, d += cw, p += stride)
            for (k = 0; k < unit_w; k++) d[k] = (float)p[k] - mean;
    }
    {                                                                 /* :189-205 */
        float* d = left;
        p = roi_origin - w;
        for (i = 0; i < left_units; i++) {
            if (!flags[left_units - 1 - i]) continue;
            for (j = 0; j < unit_h; j++, d += w, p += stride)
                for (k = 0; k < w; k++) d[k] = (float)p[k] - mean;
        }
    }
    return 0;
}

/* reference TComPrediction.cpp(substitution):621-635 */
void oracle_epilogue_hm(const float* pred, int n, float mean, int32_t* out) {
    for (int i = 0; i < n; i++) {
        float v = pred[i] + mean;
        v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
        out[i] = (int32_t)roundf(v);
    }
}

/* reference tools/tools.py:49 (numpy.round = half to even) */
void oracle_epilogue_numpy(const float* pred, int n, float mean, uint8_t* out) {
    for (int i = 0; i < n; i++) {
        float v = pred[i] + mean;
        v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
        out[i] = (uint8_t)rintf(v);
    }
}

/* reference pnn/components.py:169-176: one fully-connected layer, y = act(x W + b), W [in][out] */
void oracle_fc_layer(const float* x, const float* w, const float* b, int n, int in, int out, int leaky, float* y) {
    for (int s = 0; s < n; s++) {
        float* ys = y + (size_t)s * out;
        for (int o = 0; o < out; o++) ys[o] = 0.f;
        for (int i = 0; i < in; i++) {
            const float xv = x[(size_t)s * in + i];
            const float* wr = w + (size_t)i * out;
            for (int o = 0; o < out; o++) ys[o] += xv * wr[o];
        }
        for (int o = 0; o < out; o++) {
            float v = ys[o] + b[o];
            ys[o] = leaky ? fmaxf(0.1f * v, v) : v;                    /* pnn/tfutils.py:192 */
        }
    }
}
