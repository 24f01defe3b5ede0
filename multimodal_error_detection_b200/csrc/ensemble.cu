// Frame -> window post-processing and ensemble fusion (config 5 of BASELINE.json).
//   window_vote  <- window_predictions        MED/modeling/modeling_utils.py:2752-2758
//   soft_vote    <- ensemble.ipynb cell 6     (p_video + p_kin) / 2 >= 0.5
//   cascade      <- ensemble.ipynb cell 15    multiclass where the binary model fired, else 0
//   confusion    <- sklearn.confusion_matrix inputs of the above
// Elementwise / tiny reductions: latency-bound, one thread per window.
#include "common.cuh"

namespace b200med {

__global__ void window_vote_kernel(const float *__restrict__ fp, const int32_t *__restrict__ starts, long long n,
                                   int W, int binary, float *__restrict__ out) {
    pdl_wait();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // numpy mean: pairwise summation degenerates to a plain left-to-right fp64 sum below 8 elements per
    // lane and is exact anyway for the 0/1 or small-integer frame predictions this is applied to.
    double acc = 0.0;
    const long long s = starts[i];
    for (int t = 0; t < W; ++t) acc += (double)fp[s + t];
    const double m = acc / (double)W;
    out[i] = binary ? (m >= 0.5 ? 1.0f : 0.0f) : (float)rint(m);  // np.round = half-to-even
}

__global__ void soft_vote_kernel(const float *__restrict__ pa, const float *__restrict__ pb,
                                 const float *__restrict__ labels, long long n, float *__restrict__ preds,
                                 unsigned long long *__restrict__ cnt) {
    pdl_wait();
    __shared__ unsigned int c[4];
    if (threadIdx.x < 4) c[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double m = ((double)pa[i] + (double)pb[i]) / 2.0;
        const int p = m >= 0.5 ? 1 : 0;
        if (preds) preds[i] = (float)p;
        if (labels) atomicAdd(&c[(labels[i] > 0.5f ? 1 : 0) * 2 + p], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 4 && c[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)c[threadIdx.x]);
}

__global__ void cascade_kernel(const int32_t *__restrict__ bin, const int32_t *__restrict__ multi, long long n,
                               int32_t *__restrict__ out) {
    pdl_wait();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = bin[i] == 1 ? multi[i] : 0;
}

__global__ void confusion_kernel(const int32_t *__restrict__ t, const int32_t *__restrict__ p, long long n, int C,
                                 unsigned long long *__restrict__ cm) {
    pdl_wait();
    __shared__ unsigned int c[64];
    for (int k = threadIdx.x; k < 64; k += blockDim.x) c[k] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int a = t[i], b = p[i];
        if (a >= 0 && a < C && b >= 0 && b < C) atomicAdd(&c[a * C + b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C * C; k += blockDim.x)
        if (c[k]) atomicAdd(&cm[k], (unsigned long long)c[k]);  // integer adds: order-independent
}

__global__ void zero_u64_kernel(unsigned long long *p, int n) {
    pdl_wait();
    if (threadIdx.x < n) p[threadIdx.x] = 0;
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int b200med_window_vote(const float *frame_preds, const int32_t *starts, int64_t n, int32_t W,
                                   int32_t binary, float *out, void *stream) {
    B200MED_REQUIRE(n >= 0 && W >= 1, "bad shape");
    if (n == 0) return B200MED_OK;
    B200MED_REQUIRE(frame_preds && starts && out, "null pointer");
    launch_k(window_vote_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, frame_preds, starts, n, W, binary, out);
    return after_launch("window_vote_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_soft_vote(const float *pa, const float *pb, const float *labels, int64_t n, float *preds,
                                 int64_t *counts, int32_t accumulate, void *, void *stream) {
    B200MED_REQUIRE(n >= 1, "empty input");
    B200MED_REQUIRE(pa && pb && (!labels || counts), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (counts && !accumulate) {
        launch_k(zero_u64_kernel, 1, 32, 0, st, (unsigned long long *)counts, 4);
        if (int e = after_launch("zero_u64_kernel")) return e;
    }
    const long long want = (n + 255) / 256, cap = (long long)num_sms() * 4;
    launch_k(soft_vote_kernel, (unsigned)(want < cap ? want : cap), 256, 0, st, pa, pb, labels, n, preds, (unsigned long long *)counts);
    return after_launch("soft_vote_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_cascade(const int32_t *binary, const int32_t *multiclass, int64_t n, int32_t *out, void *stream) {
    B200MED_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return B200MED_OK;
    B200MED_REQUIRE(binary && multiclass && out, "null pointer");
    launch_k(cascade_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, binary, multiclass, n, out);
    return after_launch("cascade_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_confusion(const int32_t *target, const int32_t *pred, int64_t n, int32_t C, int64_t *cm,
                                 int32_t accumulate, void *stream) {
    B200MED_REQUIRE(C >= 1 && C <= 8 && n >= 0, "1..8 classes");
    B200MED_REQUIRE(cm && (n == 0 || (target && pred)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) {
        launch_k(zero_u64_kernel, 1, 64, 0, st, (unsigned long long *)cm, C * C);
        if (int e = after_launch("zero_u64_kernel")) return e;
    }
    if (n == 0) return B200MED_OK;
    const long long want = (n + 255) / 256, cap = (long long)num_sms() * 4;
    launch_k(confusion_kernel, (unsigned)(want < cap ? want : cap), 256, 0, st, target, pred, n, C, (unsigned long long *)cm);
    return after_launch("confusion_kernel");
}
