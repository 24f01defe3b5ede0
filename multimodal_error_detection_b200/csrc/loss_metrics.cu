// K3: fused loss + dlogits + predictions + confusion counts (latency-bound, deterministic).
//
// One pass over the logits replaces: BCEWithLogitsLoss / CrossEntropyLoss forward and backward,
// sigmoid / softmax, thresholding / argmax and the inputs of sklearn's confusion_matrix / f1 /
// accuracy / jaccard (MED/modeling/modeling_utils.py:234-254, 265-297, 374-381, 493-528).  8-28 bytes
// per window, so the kernel is sized for launch latency: <= 64 CTAs, per-thread double accumulation,
// warp-shuffle + shared-memory block reduction, block partials combined by the LAST CTA in ascending
// block order (fixed order => bit-reproducible loss).
//
// Workspace: [0] u32 ticket (must be zero before the first call; the kernel leaves it zero),
//            [64..] double partials[kMaxBlocks][kSlots].
#include "common.cuh"

namespace b200med {

constexpr int kLossThreads = 256;
constexpr int kMaxBlocks = 64;
constexpr int kSlots = 8;           // doubles per block partial (scalar losses)
constexpr int kMaxClasses = 8;      // multi-class heads of the reference have 5 or 6 classes
constexpr int kCmSlots = kMaxClasses * kMaxClasses;

__device__ __forceinline__ double block_sum(double v, double *sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (kLossThreads >> 5) ? sh[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;  // valid in warp 0
}

// Publishes this block's partial vector and returns true in the last block to arrive.
__device__ __forceinline__ bool publish_and_elect(unsigned int *ticket) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) __threadfence();
    return last;
}

__global__ void __launch_bounds__(kLossThreads)
bce_logits_kernel(const float *__restrict__ logits, const float *__restrict__ labels, long long B, float pos_weight,
                  float grad_scale, float *__restrict__ loss, float *__restrict__ dlogits, float *__restrict__ probs,
                  float *__restrict__ preds, long long *__restrict__ counts, int accumulate,
                  unsigned int *ticket, double *partials) {
    pdl_wait();
    double l_sum = 0.0, c[4] = {0.0, 0.0, 0.0, 0.0};
    const float inv_b = grad_scale / (float)B;
    for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < B; i += (long long)gridDim.x * kLossThreads) {
        const float x = logits[i], y = labels[i];
        const float lw = (pos_weight - 1.0f) * y + 1.0f;
        const float softplus = log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.0f);
        l_sum += (double)((1.0f - y) * x + lw * softplus);
        const float sg = 1.0f / (1.0f + expf(-x));
        const float pred = sg > 0.5f ? 1.0f : 0.0f;
        if (dlogits) dlogits[i] = (lw * sg - pos_weight * y) * inv_b;
        if (probs) probs[i] = sg;
        if (preds) preds[i] = pred;
        const int yi = y > 0.5f ? 1 : 0, pi = (int)pred;
        c[yi * 2 + pi] += 1.0;
    }
    // the five block sums in ONE pass (same shuffle trees as block_sum, so the same bits): one barrier instead of ten
    double vals[5] = {l_sum, c[0], c[1], c[2], c[3]};
    __shared__ double shv[kLossThreads / 32][5];
    __shared__ double part_sh[kMaxBlocks * 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        vals[k] = warp_sum(vals[k]);
        if (lane == 0) shv[warp][k] = vals[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double t = warp_sum(lane < (kLossThreads >> 5) ? shv[lane][k] : 0.0);
            if (lane == 0) partials[blockIdx.x * kSlots + k] = t;
        }
    }
    if (publish_and_elect(ticket)) {
        // the last block: all partials fetched in one round trip, then summed by one thread in ascending block order
        for (int e = threadIdx.x; e < (int)gridDim.x * 5; e += kLossThreads) part_sh[e] = partials[(e / 5) * kSlots + e % 5];
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot[5] = {0, 0, 0, 0, 0};
            for (unsigned int b = 0; b < gridDim.x; ++b)
                for (int k = 0; k < 5; ++k) tot[k] += part_sh[b * 5 + k];
            loss[0] = (float)(tot[0] / (double)B);
            if (counts)
                for (int k = 0; k < 4; ++k) counts[k] = (accumulate ? counts[k] : 0) + (long long)(tot[1 + k] + 0.5);
            *ticket = 0;
        }
    }
}

__global__ void __launch_bounds__(kLossThreads)
ce_logits_kernel(const float *__restrict__ logits, const int32_t *__restrict__ target,
                 const float *__restrict__ class_weight, const float *__restrict__ mask, long long B, int C,
                 int target_shift, int reduction, float grad_scale, float *__restrict__ loss,
                 float *__restrict__ dlogits, float *__restrict__ probs, int32_t *__restrict__ preds, int pred_shift,
                 int pred_mask_mode, long long *__restrict__ cm, int cm_classes, int accumulate,
                 unsigned int *ticket, double *partials) {
    pdl_wait();
    __shared__ double sh[32];
    __shared__ double denom_sh;
    __shared__ unsigned int cm_sh[kCmSlots];
    // Every block derives the SAME denominator in the same order (B is small and L2 resident).
    {
        double d = 0.0;
        for (long long i = threadIdx.x; i < B; i += kLossThreads) {
            if (reduction == 0) {
                const int t = max(target[i] + target_shift, 0);
                d += class_weight ? (double)class_weight[t] : 1.0;
            } else {
                d += mask ? (double)mask[i] : 1.0;
            }
        }
        d = block_sum(d, sh);
        if (threadIdx.x == 0) denom_sh = d;
        for (int k = threadIdx.x; k < kCmSlots; k += kLossThreads) cm_sh[k] = 0;
        __syncthreads();
    }
    const double denom = denom_sh;
    // gradient scale per reduction mode
    float gden;
    if (reduction == 0) gden = (float)denom;
    else if (reduction == 1) gden = denom > 0.0 ? (float)denom : 1.0f;
    else if (reduction == 2) gden = 1.0f;
    else gden = denom > 0.0 ? 1.0f : (float)B;
    double l_sum = 0.0;
    for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < B; i += (long long)gridDim.x * kLossThreads) {
        const float *row = logits + i * C;
        float v[kMaxClasses];
        float mx = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
            if (c < C) {
                v[c] = row[c];
                if (v[c] > mx) { mx = v[c]; arg = c; }
            }
        float se = 0.0f;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
            if (c < C) se += expf(v[c] - mx);
        const float lse = logf(se);
        const int raw_t = target[i];
        const int t = max(raw_t + target_shift, 0);
        const float w = (reduction == 0 && class_weight) ? class_weight[t] : 1.0f;
        const float m = (reduction != 0 && reduction != 2 && mask) ? mask[i] : 1.0f;
        const float li = -(v[t] - mx - lse) * w;
        // modes 0,1,2 reduce l*m (m == 1 without mask); mode 3 reduces the plain l (see header)
        l_sum += (double)(reduction == 3 ? li : li * m);
        const float gscale = (reduction == 3 ? 1.0f : m) * w * grad_scale / gden;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
            if (c < C) {
                const float p = expf(v[c] - mx - lse);
                if (probs) probs[i * C + c] = p;
                if (dlogits) dlogits[i * C + c] = (p - (c == t ? 1.0f : 0.0f)) * gscale;
            }
        int pred = arg + pred_shift;
        if (pred_mask_mode == 1 && raw_t == 0) pred = 0;
        if (pred_mask_mode == 2 && mask && !(mask[i] > 0.0f)) pred = 0;
        if (preds) preds[i] = pred;
        if (cm && raw_t >= 0 && raw_t < cm_classes && pred >= 0 && pred < cm_classes)
            atomicAdd(&cm_sh[raw_t * cm_classes + pred], 1u);
    }
    const double t = block_sum(l_sum, sh);
    if (threadIdx.x == 0) partials[blockIdx.x * (kSlots + kCmSlots)] = t;
    __syncthreads();
    for (int k = threadIdx.x; k < kCmSlots; k += kLossThreads)
        partials[blockIdx.x * (kSlots + kCmSlots) + kSlots + k] = (double)cm_sh[k];
    if (publish_and_elect(ticket)) {
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) tot += partials[b * (kSlots + kCmSlots)];
            double out;
            if (reduction == 0) out = tot / denom;
            else if (reduction == 1) out = denom > 0.0 ? tot / denom : tot;
            else if (reduction == 2) out = tot;
            else out = denom > 0.0 ? tot : tot / (double)B;
            loss[0] = (float)out;
        }
        if (cm) {
            for (int k = threadIdx.x; k < cm_classes * cm_classes; k += kLossThreads) {
                double s = 0.0;
                for (unsigned int b = 0; b < gridDim.x; ++b) s += partials[b * (kSlots + kCmSlots) + kSlots + k];
                cm[k] = (accumulate ? cm[k] : 0) + (long long)(s + 0.5);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// logits [S, 2, T]; soft targets [1-e, e]; mean over T then over S (modeling_utils.py:278-295).
__global__ void __launch_bounds__(kLossThreads)
ce_frame_kernel(const float *__restrict__ logits, const float *__restrict__ e, int stages, long long T,
                float grad_scale, float *__restrict__ loss, float *__restrict__ dlogits, float *__restrict__ preds,
                long long *__restrict__ counts, int accumulate, unsigned int *ticket, double *partials) {
    pdl_wait();
    __shared__ double sh[32];
    double l_sum = 0.0, c[4] = {0.0, 0.0, 0.0, 0.0};
    const float gs = grad_scale / ((float)T * (float)stages);
    for (long long t = blockIdx.x * (long long)kLossThreads + threadIdx.x; t < T; t += (long long)gridDim.x * kLossThreads) {
        const float y1 = e[t], y0 = 1.0f - y1;
        for (int s = 0; s < stages; ++s) {
            const float a = logits[((long long)s * 2 + 0) * T + t], b = logits[((long long)s * 2 + 1) * T + t];
            const float mx = fmaxf(a, b);
            const float lse = logf(expf(a - mx) + expf(b - mx));
            const float lp0 = a - mx - lse, lp1 = b - mx - lse;
            l_sum += (double)(-(y0 * lp0 + y1 * lp1));
            if (dlogits) {
                const float ysum = y0 + y1;
                dlogits[((long long)s * 2 + 0) * T + t] = (expf(lp0) * ysum - y0) * gs;
                dlogits[((long long)s * 2 + 1) * T + t] = (expf(lp1) * ysum - y1) * gs;
            }
            if (s == stages - 1) {
                const int pi = b > a ? 1 : 0;  // torch.max returns the first index on ties
                if (preds) preds[t] = (float)pi;
                c[(y1 > 0.5f ? 1 : 0) * 2 + pi] += 1.0;
            }
        }
    }
    double vals[5] = {l_sum, c[0], c[1], c[2], c[3]};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double tt = block_sum(vals[k], sh);
        if (threadIdx.x == 0) partials[blockIdx.x * kSlots + k] = tt;
    }
    if (publish_and_elect(ticket) && threadIdx.x == 0) {
        double tot[5] = {0, 0, 0, 0, 0};
        for (unsigned int b = 0; b < gridDim.x; ++b)
            for (int k = 0; k < 5; ++k) tot[k] += partials[b * kSlots + k];
        loss[0] = (float)(tot[0] / ((double)T * stages));
        if (counts)
            for (int k = 0; k < 4; ++k) counts[k] = (accumulate ? counts[k] : 0) + (long long)(tot[1 + k] + 0.5);
        *ticket = 0;
    }
}

static int loss_grid(long long B) {
    long long g = (B + kLossThreads - 1) / kLossThreads;
    if (g < 1) g = 1;
    return (int)(g < kMaxBlocks ? g : kMaxBlocks);
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int64_t b200med_loss_ws_bytes(int64_t) { return 64 + (int64_t)kMaxBlocks * (kSlots + kCmSlots) * 8; }

extern "C" __attribute__((visibility("default"))) int b200med_bce_logits(const float *logits, const float *labels, int64_t B, float pos_weight,
                                  float grad_scale, float *loss, float *dlogits, float *probs, float *preds,
                                  int64_t *counts, int32_t accumulate, void *workspace, void *stream) {
    B200MED_REQUIRE(B >= 1, "empty batch");
    B200MED_REQUIRE(logits && labels && loss && workspace, "null pointer");
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    double *partials = reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + 64);
    launch_k(bce_logits_kernel, loss_grid(B), kLossThreads, 0, (cudaStream_t)stream, 
        logits, labels, B, pos_weight, grad_scale, loss, dlogits, probs, preds, (long long *)counts, accumulate,
        ticket, partials);
    return after_launch("bce_logits_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_ce_logits(const float *logits, const int32_t *target, const float *class_weight,
                                 const float *mask, int64_t B, int32_t C, int32_t target_shift, int32_t reduction,
                                 float grad_scale, float *loss, float *dlogits, float *probs, int32_t *preds,
                                 int32_t pred_shift, int32_t pred_mask_mode, int64_t *cm, int32_t cm_classes,
                                 int32_t accumulate, void *workspace, void *stream) {
    B200MED_REQUIRE(B >= 1, "empty batch");
    B200MED_REQUIRE(C >= 2 && C <= kMaxClasses, "2..8 classes supported");
    B200MED_REQUIRE(!cm || (cm_classes >= 2 && cm_classes <= kMaxClasses), "2..8 confusion classes supported");
    B200MED_REQUIRE(reduction >= 0 && reduction <= 3, "bad reduction mode");
    B200MED_REQUIRE(logits && target && loss && workspace, "null pointer");
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    double *partials = reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + 64);
    launch_k(ce_logits_kernel, loss_grid(B), kLossThreads, 0, (cudaStream_t)stream, 
        logits, target, class_weight, mask, B, C, target_shift, reduction, grad_scale, loss, dlogits, probs, preds,
        pred_shift, pred_mask_mode, (long long *)cm, cm_classes, accumulate, ticket, partials);
    return after_launch("ce_logits_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_ce_frame(const float *logits, const float *e, int32_t stages, int64_t T, float grad_scale,
                                float *loss, float *dlogits, float *preds, int64_t *counts, int32_t accumulate,
                                void *workspace, void *stream) {
    B200MED_REQUIRE(T >= 1 && stages >= 1, "empty video");
    B200MED_REQUIRE(logits && e && loss && workspace, "null pointer");
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    double *partials = reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + 64);
    launch_k(ce_frame_kernel, loss_grid(T), kLossThreads, 0, (cudaStream_t)stream, 
        logits, e, stages, T, grad_scale, loss, dlogits, preds, (long long *)counts, accumulate, ticket, partials);
    return after_launch("ce_frame_kernel");
}
