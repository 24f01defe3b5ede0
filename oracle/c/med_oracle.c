/* CPU oracle (plain C) for the integer / byte-moving part of the b200med hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).  Loaded with ctypes by the
 * tests, by __graft_entry__.smoke() and by bench.py's cpu_baseline leg only.
 *
 * Each function restates one reference routine of GonzaloPlaaza/Multimodal-Error-Detection:
 *   med_oracle_window_starts  <- MED/dataset/dataset_utils.py:206-240  (window_data walk)
 *   med_oracle_powerset       <- MED/dataset/dataset_utils.py:793-843  (powerset_error_labels)
 *   med_oracle_gather_norm    <- MED/dataset/dataset_utils.py:230-231 + MED/dataset/CustomWindowDataset.py:56-60
 *   med_oracle_window_mean    <- MED/modeling/modeling_utils.py:2752-2758 (window_predictions)
 * Pinned against the reference's own outputs through tests/golden/ (tests/test_oracle_golden.py).
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

/* Subjects are contiguous row ranges [offsets[s], offsets[s+1]).  Writes global start rows in
 * subject order; returns the number of windows (counting only when starts_out == NULL or when
 * capacity is exceeded).  Returns -1 - s if subject s has no non-zero gesture (the reference
 * raises IndexError there, dataset_utils.py:211-212). */
int64_t med_oracle_window_starts(const float *g, const int64_t *offsets, int64_t n_subjects,
                                 int64_t W, int64_t S, int64_t *starts_out, int64_t capacity)
{
    int64_t n_out = 0;
    for (int64_t s = 0; s < n_subjects; ++s) {
        const float *gs = g + offsets[s];
        int64_t n = offsets[s + 1] - offsets[s];
        int64_t pos = 0;
        while (pos < n && !(gs[pos] != 0.0f)) ++pos;          /* first non-zero gesture */
        if (pos >= n) return -1 - s;
        while (pos < n - W) {                                   /* strict bound (:214)    */
            if (gs[pos] != gs[pos + W - 1]) { pos += 1; continue; }   /* end points only (:220-226) */
            if (starts_out && n_out < capacity) starts_out[n_out] = offsets[s] + pos;
            ++n_out;
            pos += S;
        }
    }
    return n_out;
}

void med_oracle_powerset(const float *e5, int64_t n, int delete_nd, int32_t *e7, uint8_t *nd_mask)
{
    for (int64_t i = 0; i < n; ++i) {
        const float oov = e5[5 * i + 0], nd = e5[5 * i + 1], ma = e5[5 * i + 2],
                    np_ = e5[5 * i + 3], err = e5[5 * i + 4];
        int32_t *o = e7 + 7 * i;
        for (int j = 0; j < 7; ++j) o[j] = 0;
        nd_mask[i] = 0;
        if (err == 1.0f) {
            o[6] = 1;
            const int single = (((oov + nd) + ma) + np_) == 1.0f;
            if ((oov == 1.0f && single) || (oov == 1.0f && nd == 1.0f)) o[1] = 1;
            else if ((ma == 1.0f && single) || (ma == 1.0f && nd == 1.0f)) o[2] = 1;
            else if ((np_ == 1.0f && single) || (np_ == 1.0f && oov == 1.0f)) o[3] = 1;
            else if (oov == 1.0f && ma == 1.0f) o[4] = 1;
            else if (ma == 1.0f && np_ == 1.0f) o[5] = 1;
            else if (nd == 1.0f) { if (delete_nd) { o[6] = 0; nd_mask[i] = 1; } }
        } else {
            o[0] = 1;
        }
    }
}

/* out[b, t, :] = (table[starts[b] + t, :] - mean) / std, fp32, subtract then divide. */
void med_oracle_gather_norm(const float *table, int64_t dim, const float *mean, const float *std_,
                            const int64_t *starts, int64_t B, int64_t W, float *out)
{
    for (int64_t b = 0; b < B; ++b)
        for (int64_t t = 0; t < W; ++t) {
            const float *src = table + (starts[b] + t) * dim;
            float *dst = out + (b * W + t) * dim;
            for (int64_t d = 0; d < dim; ++d) dst[d] = (src[d] - mean[d]) / std_[d];
        }
}

/* Window value of frame-level predictions: float64 mean over the window, then ">= 0.5"
 * (binary) or round-half-to-even (multi-class, numpy np.round). */
void med_oracle_window_mean(const double *preds, const int64_t *starts, int64_t n, int64_t W,
                            int binary, double *out)
{
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int64_t t = 0; t < W; ++t) acc += preds[starts[i] + t];
        double m = acc / (double)W;
        out[i] = binary ? (m >= 0.5 ? 1.0 : 0.0) : nearbyint(m);
    }
}
