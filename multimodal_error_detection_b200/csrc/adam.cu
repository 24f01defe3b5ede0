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

}  // namespace b200med

using namespace b200med;

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
