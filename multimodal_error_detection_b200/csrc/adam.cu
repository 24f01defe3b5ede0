// Fused Adam over the flat parameter buffer (torch.optim.Adam semantics, coupled L2 decay),
// replacing the per-tensor optimiser launches of MED/modeling/modeling_utils.py:221-222, 363-365.
//
// HBM-bound: 16 B read (p, g, m, v) + 12 B written per parameter = 28 B/param; 1.27-1.60 M parameters
// -> 36-45 MB per step, L2 resident on B200.  lr and the bias corrections live in a small device
// `state` vector so that a captured CUDA graph can be replayed across steps and epochs.
#include "common.cuh"

namespace b200med {

// state: {step, lr, bias_corr1, sqrt(bias_corr2)}
__global__ void adam_advance_kernel(float *state, float beta1, float beta2) {
    pdl_wait();
    const double step = (double)state[0] + 1.0;
    state[0] = (float)step;
    state[2] = (float)(1.0 - pow((double)beta1, step));
    state[3] = (float)sqrt(1.0 - pow((double)beta2, step));
}

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float lr_over_bc1, float bc2_sqrt,
                                         float beta1, float beta2, float eps, float wd, float gscale) {
    g = g * gscale;
    g = fmaf(wd, p, g);                          // grad = grad + weight_decay * param
    m = m + (g - m) * (1.0f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + (1.0f - beta2) * g * g;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - lr_over_bc1 * (m / denom);           // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
adam_step_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 long long n, const float *__restrict__ state, float beta1, float beta2, float eps, float wd,
                 float gscale) {
    pdl_wait();
    const float lr = state[1], bc1 = state[2], bc2s = state[3];
    const float step_size = lr / bc1;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4 *>(p)[i];
        const float4 G = reinterpret_cast<const float4 *>(g)[i];
        float4 M = reinterpret_cast<float4 *>(m)[i];
        float4 V = reinterpret_cast<float4 *>(v)[i];
        adam_one(P.x, G.x, M.x, V.x, step_size, bc2s, beta1, beta2, eps, wd, gscale);
        adam_one(P.y, G.y, M.y, V.y, step_size, bc2s, beta1, beta2, eps, wd, gscale);
        adam_one(P.z, G.z, M.z, V.z, step_size, bc2s, beta1, beta2, eps, wd, gscale);
        adam_one(P.w, G.w, M.w, V.w, step_size, bc2s, beta1, beta2, eps, wd, gscale);
        reinterpret_cast<float4 *>(p)[i] = P;
        reinterpret_cast<float4 *>(m)[i] = M;
        reinterpret_cast<float4 *>(v)[i] = V;
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
        adam_one(p[i], g[i], m[i], v[i], step_size, bc2s, beta1, beta2, eps, wd, gscale);
}

// ---- many small copies in ONE launch: the gradients autograd hands over (one tensor per parameter) are gathered into the flat
// gradient buffer; the bf16 copies of the FeatureExtractor weights are made in one pass.  Pointers, sizes and the first block
// of every pair travel in the kernel's parameter space (no pointer table in device memory: nothing to upload per step, and a
// captured graph keeps them by value).
constexpr int kMultiMax = 48;           // pairs per launch
constexpr int kMultiChunk = 8192;       // elements per block

struct MultiCopyArgs {
    const float *src[kMultiMax];
    void *dst[kMultiMax];
    long long n[kMultiMax];
    int blk0[kMultiMax + 1];
    int count;
};

template <typename TO>
__global__ void __launch_bounds__(256)
multi_copy_kernel(const __grid_constant__ MultiCopyArgs a, float *adam_state, float beta1, float beta2) {
    pdl_wait();
    if (adam_state && blockIdx.x == 0 && threadIdx.x == 0) {      // adam_advance_kernel's arithmetic
        const double step = (double)adam_state[0] + 1.0;
        adam_state[0] = (float)step;
        adam_state[2] = (float)(1.0 - pow((double)beta1, step));
        adam_state[3] = (float)sqrt(1.0 - pow((double)beta2, step));
    }
    int t = 0;
    while (t + 1 < a.count && (int)blockIdx.x >= a.blk0[t + 1]) ++t;
    const long long n = a.n[t], e0 = (long long)((int)blockIdx.x - a.blk0[t]) * kMultiChunk;
    const long long e1 = e0 + kMultiChunk < n ? e0 + kMultiChunk : n;
    const float *src = a.src[t];
    TO *dst = reinterpret_cast<TO *>(a.dst[t]);
    const bool vec = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    if (vec) {
        const long long v1 = e0 + ((e1 - e0) & ~3LL);
#pragma unroll 4
        for (long long e = e0 + 4LL * threadIdx.x; e < v1; e += 4LL * blockDim.x) {
            const float4 v = *reinterpret_cast<const float4 *>(src + e);
            if constexpr (sizeof(TO) == 4) {
                *reinterpret_cast<float4 *>(dst + e) = v;
            } else {
                uint2 o;
                o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
                *reinterpret_cast<uint2 *>(dst + e) = o;
            }
        }
        for (long long e = v1 + threadIdx.x; e < e1; e += blockDim.x) dst[e] = (TO)src[e];
    } else {
        for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) dst[e] = (TO)src[e];
    }
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int b200med_multi_copy_f32(const void *const *src, void *const *dst, const int64_t *n,
                                 int32_t count, int32_t dst_dtype, float *adam_state, float beta1, float beta2, void *stream) {
    B200MED_REQUIRE(count >= 0 && (count == 0 || (src && dst && n)), "bad arguments");
    B200MED_REQUIRE(dst_dtype == B200MED_F32 || dst_dtype == B200MED_BF16, "dst_dtype must be f32 or bf16");
    cudaStream_t st = (cudaStream_t)stream;
    bool advanced = adam_state == nullptr;
    for (int base = 0; base < count || !advanced; base += kMultiMax) {
        MultiCopyArgs a;
        a.count = 0;
        int blocks = 0;
        for (int i = base; i < count && i < base + kMultiMax; ++i) {
            if (n[i] <= 0) continue;
            B200MED_REQUIRE(src[i] && dst[i], "null pointer");
            a.src[a.count] = (const float *)src[i]; a.dst[a.count] = dst[i]; a.n[a.count] = n[i]; a.blk0[a.count] = blocks;
            blocks += (int)((n[i] + kMultiChunk - 1) / kMultiChunk);
            ++a.count;
        }
        a.blk0[a.count] = blocks;
        if (a.count == 0) {                    // nothing to copy in this batch: only the Adam scalars may be left to advance
            if (advanced) continue;
            a.src[0] = nullptr; a.dst[0] = nullptr; a.n[0] = 0; a.blk0[0] = 0; a.blk0[1] = 1; a.count = 1; blocks = 1;
        }
        float *state = advanced ? nullptr : adam_state;
        advanced = true;
        if (dst_dtype == B200MED_F32) launch_k(multi_copy_kernel<float>, (unsigned)blocks, 256, 0, st, a, state, beta1, beta2);
        else launch_k(multi_copy_kernel<__nv_bfloat16>, (unsigned)blocks, 256, 0, st, a, state, beta1, beta2);
        if (int e = after_launch("multi_copy_kernel")) return e;
    }
    return B200MED_OK;
}


extern "C" __attribute__((visibility("default"))) int b200med_adam_advance(float *state, float beta1, float beta2, void *stream) {
    B200MED_REQUIRE(state, "null state");
    launch_k(adam_advance_kernel, 1, 1, 0, (cudaStream_t)stream, state, beta1, beta2);
    return after_launch("adam_advance_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_adam_step(float *p, const float *g, float *m, float *v, int64_t n, const float *state,
                                 float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                 void *stream) {
    B200MED_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return B200MED_OK;
    B200MED_REQUIRE(p && g && m && v && state, "null pointer");
    B200MED_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "flat buffers must be 16-byte aligned");
    const long long want = ((n >> 2) + 255) / 256 + 1, cap = (long long)num_sms() * 8;
    launch_k(adam_step_kernel, (unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream, 
        p, g, m, v, n, state, beta1, beta2, eps, weight_decay, grad_scale);
    return after_launch("adam_step_kernel");
}
