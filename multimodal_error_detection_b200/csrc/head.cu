// Window-classifier heads (MED/modeling/models.py:49-131 CNN, :166-186 / :204-210 the LSTM head's MLP) -- the layers around
// the GEMMs: BatchNorm1d (batch statistics, running-stat update, backward), MaxPool1d(2) + Dropout, and the weight
// re-packing that turns Conv1d(k = 3) into a GEMM over OVERLAPPING rows of the time-major activations.
//
// Layout.  Activations are row-major [rows, C] with the channel contiguous: a Linear layer's [B, C]; a convolution's
// time-major [B * L, C] (row = (window b, step l)) -- the reference's Conv1d runs over [B, C, L], but the head input is a
// permuted VIEW of a [B, W, F] tensor (modeling_utils.py:47), so time-major is the layout the data already has.
//   conv(k=3):  out[b, l, co] = sum_{k, ci} x[b, l+k, ci] w[co, ci, k]  = row (b, l) of a GEMM whose A row is the 3*Cin
//               CONTIGUOUS floats starting at x[b, l, 0] (row stride Cin: rows overlap, no im2col copy) and whose B is
//               w' [Cout, 3*Cin] with w'[co, k*Cin + ci] = w[co, ci, k] (conv_pack_kernel).  The product runs over all
//               B*L - 2 rows; rows with l >= L - 2 straddle two windows and are never read afterwards.
//   backward:   dx[b, l, ci] = sum_{k, co} dz[b, l-k, co] w[co, ci, k] = the same trick on dz with two zero rows in front
//               and w'' [Cin, 3*Cout], w''[ci, k'*Cout + co] = w[co, ci, 2-k'] (dz is zero on the straddling rows).
// BatchNorm statistics: per-slab (n, mean, M2) partials by a local two-pass, combined with Chan's formula in double and in
// a fixed order -> deterministic, and as accurate as the CPU reference's double accumulation for the 1e-5 bar.
#include "common.cuh"

namespace b200med {

constexpr int kBnCols = 32;   // columns per CTA (one per lane)
constexpr int kBnWarps = 8;

__device__ __forceinline__ bool head_keep(uint32_t seed, unsigned long long index, float p) {
    uint64_t z = index + 0x9E3779B97F4A7C15ull * (uint64_t)(seed + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
    return u >= p;
}

// ---- BatchNorm: three launches per direction -- vectorised per-slab partials, a tiny per-column finalize, an elementwise
// apply.  V = columns per thread (4 when C % 4 == 0: 16-byte accesses, 128 columns per warp row; else 1).

template <int V> struct BnVec;
template <> struct BnVec<4> {
    __device__ static void load(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ static void store(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct BnVec<1> {
    __device__ static void load(const float *p, float (&v)[1]) { v[0] = *p; }
    __device__ static void store(float *p, const float (&v)[1]) { *p = v[0]; }
};

// part [S][3][C] = (n, mean, M2) of every row slab (local two-pass; the slab is re-read from L1 / L2)
template <int V>
__global__ void __launch_bounds__(kBnCols * kBnWarps)
bn_stats_partial_kernel(const float *__restrict__ x, long long M, int C, long long rows_per_slab, float *__restrict__ part) {
    pdl_wait();
    __shared__ double sh[kBnWarps][kBnCols * V];
    __shared__ float sh_mean[kBnCols * V];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = (blockIdx.x * kBnCols + lane) * V;
    const long long m0 = (long long)blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
    const bool ok = col < C;
    float s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = 0.0f;
    if (ok) {
#pragma unroll 4
        for (long long m = m0 + w; m < m1; m += kBnWarps) {
            float v[V];
            BnVec<V>::load(x + m * C + col, v);
#pragma unroll
            for (int j = 0; j < V; ++j) s[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) sh[w][lane * V + j] = (double)s[j];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < kBnWarps; ++r) t += sh[r][lane * V + j];
            sh_mean[lane * V + j] = (float)(t / (double)max(1LL, m1 - m0));
        }
    }
    __syncthreads();
    float mu[V], q[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { mu[j] = sh_mean[lane * V + j]; q[j] = 0.0f; }
    if (ok) {
#pragma unroll 4
        for (long long m = m0 + w; m < m1; m += kBnWarps) {
            float v[V];
            BnVec<V>::load(x + m * C + col, v);
#pragma unroll
            for (int j = 0; j < V; ++j) { const float d = v[j] - mu[j]; q[j] = fmaf(d, d, q[j]); }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < V; ++j) sh[w][lane * V + j] = (double)q[j];
    __syncthreads();
    if (w == 0 && ok) {
        float *p = part + (long long)blockIdx.y * 3 * C;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < kBnWarps; ++r) t += sh[r][lane * V + j];
            p[col + j] = (float)(m1 - m0);
            p[C + col + j] = mu[j];
            p[2 * C + col + j] = (float)t;
        }
    }
}

// fixed-order warp sum of doubles (butterfly: every lane ends with the same bits)
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per column (ONE WARP per column, lanes over the slabs): combine the slab partials in double -- mean = sum n_s mean_s / M,
// M2 = sum [M2_s + n_s (mean_s - mean)^2] -- -> scale / shift of the normalisation, save_mean / save_rstd for the backward,
// running statistics (momentum, unbiased variance) like nn.BatchNorm1d in training mode.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float *__restrict__ part, int S, long long M, int C, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float eps, float momentum, float *__restrict__ scale_shift,
                   float *__restrict__ save_mean, float *__restrict__ save_rstd, float *__restrict__ running_mean,
                   float *__restrict__ running_var, long long *__restrict__ num_batches) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col == 0 && lane == 0 && num_batches) *num_batches += 1;
    if (col >= C) return;
    double ns[2], ms[2], qs[2];
    double wsum = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int s = lane + 32 * k;
        ns[k] = 0.0; ms[k] = 0.0; qs[k] = 0.0;
        if (s < S) {
            const float *p = part + (long long)s * 3 * C;
            ns[k] = p[col]; ms[k] = p[C + col]; qs[k] = p[2 * C + col];
        }
        wsum += ns[k] * ms[k];
    }
    const double mu = warp_sum_d(wsum) / (double)M;
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) { const double d = ms[k] - mu; q += qs[k] + ns[k] * d * d; }
    q = warp_sum_d(q);
    if (lane != 0) return;
    const double var = q / (double)M;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[col] : 1.0f, b = beta ? beta[col] : 0.0f;
    scale_shift[col] = rstd * g;
    scale_shift[C + col] = b - (float)mu * rstd * g;
    save_mean[col] = (float)mu;
    save_rstd[col] = rstd;
    if (running_mean) {
        const double unbiased = M > 1 ? q / (double)(M - 1) : var;
        running_mean[col] = (1.0f - momentum) * running_mean[col] + momentum * (float)mu;
        running_var[col] = (1.0f - momentum) * running_var[col] + momentum * (float)unbiased;
    }
}

// y = x * scale[c] + shift[c], elementwise (grid-stride, 16-byte accesses when V = 4)
template <int V>
__global__ void bn_apply_kernel(const float *__restrict__ x, long long M, int C, const float *__restrict__ scale_shift,
                                float *__restrict__ y) {
    pdl_wait();
    const long long total = M * C / V;
    const int cv = C / V;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % cv) * V;
        float v[V], o[V];
        BnVec<V>::load(x + e * V, v);
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = fmaf(v[j], __ldg(scale_shift + c + j), __ldg(scale_shift + C + c + j));
        BnVec<V>::store(y + e * V, o);
    }
}

// inference: y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta
__global__ void bn_eval_kernel(const float *__restrict__ x, long long M, int C, const float *__restrict__ rm,
                               const float *__restrict__ rv, const float *__restrict__ gamma, const float *__restrict__ beta,
                               float eps, float *__restrict__ y) {
    pdl_wait();
    const long long total = M * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const float rstd = 1.0f / sqrtf(rv[c] + eps);
        const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
        y[e] = (x[e] - rm[c]) * rstd * g + b;
    }
}

// part [S][2][C] = (sum dy, sum dy * xhat) per slab
template <int V>
__global__ void __launch_bounds__(kBnCols * kBnWarps)
bn_bwd_partial_kernel(const float *__restrict__ dy, const float *__restrict__ x, long long M, int C, long long rows_per_slab,
                      const float *__restrict__ save_mean, const float *__restrict__ save_rstd, float *__restrict__ part) {
    pdl_wait();
    __shared__ float sh[2][kBnWarps][kBnCols * V];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = (blockIdx.x * kBnCols + lane) * V;
    const long long m0 = (long long)blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
    float a[V], b[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { a[j] = 0.0f; b[j] = 0.0f; }
    if (col < C) {
        float mu[V], rs[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { mu[j] = save_mean[col + j]; rs[j] = save_rstd[col + j]; }
#pragma unroll 4
        for (long long m = m0 + w; m < m1; m += kBnWarps) {
            float g[V], xv[V];
            BnVec<V>::load(dy + m * C + col, g);
            BnVec<V>::load(x + m * C + col, xv);
#pragma unroll
            for (int j = 0; j < V; ++j) { a[j] += g[j]; b[j] = fmaf(g[j], (xv[j] - mu[j]) * rs[j], b[j]); }
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { sh[0][w][lane * V + j] = a[j]; sh[1][w][lane * V + j] = b[j]; }
    __syncthreads();
    if (w == 0 && col < C) {
        float *p = part + (long long)blockIdx.y * 2 * C;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float ta = 0.0f, tb = 0.0f;
#pragma unroll
            for (int r = 0; r < kBnWarps; ++r) { ta += sh[0][r][lane * V + j]; tb += sh[1][r][lane * V + j]; }
            p[col + j] = ta; p[C + col + j] = tb;
        }
    }
}

// Per column (one warp, lanes over the slabs): dgamma = sum dy xhat, dbeta = sum dy, and the three coefficients of
// dx = k0 * dy - k1 - k2 * x  (k0 = gamma rstd, k1 = k0 (mean(dy) - mu rstd mean(dy xhat)), k2 = k0 rstd mean(dy xhat)).
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float *__restrict__ part, int S, long long M, int C, const float *__restrict__ gamma,
                       const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
                       float *__restrict__ coef, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= C) return;
    double sa = 0.0, sb = 0.0;
    for (int s = lane; s < S; s += 32) { sa += part[(long long)s * 2 * C + col]; sb += part[(long long)s * 2 * C + C + col]; }
    sa = warp_sum_d(sa); sb = warp_sum_d(sb);
    if (lane != 0) return;
    if (dbeta) dbeta[col] = (float)sa;
    if (dgamma) dgamma[col] = (float)sb;
    const float mu = save_mean[col], rs = save_rstd[col], k0 = (gamma ? gamma[col] : 1.0f) * rs;
    const float ma = (float)(sa / (double)M), mb = (float)(sb / (double)M);
    coef[col] = k0;
    coef[C + col] = k0 * (ma - mu * rs * mb);
    coef[2 * C + col] = k0 * rs * mb;
}

// dx = k0 dy - k1 - k2 x; relu_mask: zeroed where x <= 0 -- the ReLU that sits between the Linear layer and this BatchNorm in
// the reference heads (x is that ReLU's output).
template <int V>
__global__ void bn_bwd_apply_kernel(const float *__restrict__ dy, const float *__restrict__ x, long long M, int C,
                                    const float *__restrict__ coef, int relu_mask, float *__restrict__ dx) {
    pdl_wait();
    const long long total = M * C / V;
    const int cv = C / V;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % cv) * V;
        float g[V], xv[V], o[V];
        BnVec<V>::load(dy + e * V, g);
        BnVec<V>::load(x + e * V, xv);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float v = fmaf(__ldg(coef + c + j), g[j], -__ldg(coef + C + c + j)) - __ldg(coef + 2 * C + c + j) * xv[j];
            if (relu_mask && !(xv[j] > 0.0f)) v = 0.0f;
            o[j] = v;
        }
        BnVec<V>::store(dx + e * V, o);
    }
}

// MaxPool1d(2, 2) over the steps of every window + Dropout: z [B*L rows, C] (valid steps l < Lc of every window) ->
// p [B*Lp, C], Lp = Lc / 2.
__global__ void pool_drop_fwd_kernel(const float *__restrict__ z, float *__restrict__ p, long long B, int L, int Lp, int C,
                                     float drop_p, const uint32_t *__restrict__ seed_dev, unsigned long long drop_base) {
    pdl_wait();
    const uint32_t seed = seed_dev ? *seed_dev : 0u;
    const long long total = B * (long long)Lp * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const long long r = e / C;
        const int j = (int)(r % Lp);
        const long long b = r / Lp;
        const float *src = z + ((b * L + 2 * j) * (long long)C + c);
        float v = fmaxf(src[0], src[C]);
        if (drop_p > 0.0f) v = head_keep(seed, drop_base + (unsigned long long)e, drop_p) ? v / (1.0f - drop_p) : 0.0f;
        p[e] = v;
    }
}

// dz [2 + B*L rows, C] <- dp [B*Lp, C]: the gradient goes to the FIRST maximum of each pair (torch's tie rule), through the
// dropout mask; every other row (odd tail, the straddling rows l >= Lc) is zero, and so are the TWO EXTRA ROWS IN FRONT that
// the data-gradient product of the convolution reads for l < 2 (csrc/head.cu header).
__global__ void pool_drop_bwd_kernel(const float *__restrict__ dp, const float *__restrict__ z, float *__restrict__ dz,
                                     long long B, int L, int Lp, int C, float drop_p, const uint32_t *__restrict__ seed_dev,
                                     unsigned long long drop_base) {
    pdl_wait();
    const uint32_t seed = seed_dev ? *seed_dev : 0u;
    const long long total = (B * (long long)L + 2) * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const long long r = e / C - 2;
        float v = 0.0f;
        if (r >= 0) {
            const int l = (int)(r % L);
            const long long b = r / L;
            const int j = l >> 1;
            if (j < Lp) {
                const float *pair = z + ((b * L + 2 * j) * (long long)C + c);
                const bool first_wins = !(pair[C] > pair[0]);
                if (((l & 1) == 0) == first_wins) {
                    const long long pe = (b * Lp + j) * (long long)C + c;
                    v = dp[pe];
                    if (drop_p > 0.0f) v = head_keep(seed, drop_base + (unsigned long long)pe, drop_p) ? v / (1.0f - drop_p) : 0.0f;
                }
            }
        }
        dz[e] = v;
    }
}

// w [Cout, Cin, 3] -> fwd [Cout, 3*Cin] (fwd[co, k*Cin + ci] = w[co, ci, k]) and bwd [Cin, 3*Cout]
// (bwd[ci, k*Cout + co] = w[co, ci, 2-k]); either destination may be null.
__global__ void conv_pack_kernel(const float *__restrict__ w, float *__restrict__ fwd, float *__restrict__ bwd, int Cout, int Cin) {
    pdl_wait();
    const int total = Cout * Cin * 3;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e % 3, ci = (e / 3) % Cin, co = e / (3 * Cin);
        const float v = w[e];
        if (fwd) fwd[(long long)co * 3 * Cin + k * Cin + ci] = v;
        if (bwd) bwd[(long long)ci * 3 * Cout + (2 - k) * Cout + co] = v;
    }
}
// gradient of the packed forward weight [Cout, 3*Cin] -> dw [Cout, Cin, 3]
__global__ void conv_unpack_grad_kernel(const float *__restrict__ dfwd, float *__restrict__ dw, int Cout, int Cin) {
    pdl_wait();
    const int total = Cout * Cin * 3;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e % 3, ci = (e / 3) % Cin, co = e / (3 * Cin);
        dw[e] = dfwd[(long long)co * 3 * Cin + k * Cin + ci];
    }
}

// y [B, C, R] <- x [B, R, C] (the reference's nn.Flatten runs over [B, C, L]; the native conv stack is time-major)
__global__ void transpose_last2_kernel(const float *__restrict__ x, float *__restrict__ y, long long B, int R, int C) {
    pdl_wait();
    const long long total = B * (long long)R * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % R);
        const long long q = e / R;
        const int c = (int)(q % C);
        const long long b = q / C;
        y[e] = x[(b * R + r) * (long long)C + c];
    }
}

// out [M, Ca + Cb] = [a [M, Ca] | b [M, Cb]] -- torch.cat((features, kinematics), dim=2) of define_inputs (modeling_utils.py:41-47)
__global__ void concat2_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long long M,
                               int Ca, int Cb) {
    pdl_wait();
    const int C = Ca + Cb;
    const long long total = M * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long m = e / C;
        const int c = (int)(e - m * C);
        out[e] = c < Ca ? a[m * Ca + c] : b[m * Cb + (c - Ca)];
    }
}
// da [M, Ca] = dout [M, ld] columns [col0, col0 + Ca)
__global__ void slice_cols_kernel(const float *__restrict__ dout, float *__restrict__ da, long long M, int ld, int col0, int Ca) {
    pdl_wait();
    const long long total = M * Ca;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long m = e / Ca;
        da[e] = dout[m * ld + col0 + (int)(e - m * Ca)];
    }
}
// out [n, C] = src [idx[i], :] for 4-byte elements (window-index -> start row / label lookups of a batch)
__global__ void take_rows_kernel(const uint32_t *__restrict__ src, const long long *__restrict__ idx, uint32_t *__restrict__ out,
                                 long long n, int C) {
    pdl_wait();
    const long long total = n * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / C;
        out[e] = src[idx[i] * C + (e - i * C)];
    }
}

static long long bn_slabs(long long M) {
    long long s = (M + 255) / 256;
    return s < 1 ? 1 : (s > 64 ? 64 : s);
}
static unsigned ew_grid(long long total) {
    const long long want = (total + 255) / 256, cap = (long long)num_sms() * 8;
    return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int64_t b200med_bn_ws_bytes(int64_t M, int32_t C) { return (bn_slabs(M) * 3 + 3) * (int64_t)C * 4 + 256; }

extern "C" __attribute__((visibility("default"))) int b200med_bn_fwd(const float *x, int64_t M, int32_t C, const float *gamma, const float *beta,
                               float eps, float momentum, int32_t training, float *running_mean, float *running_var,
                               int64_t *num_batches_tracked, float *y, float *save_mean, float *save_rstd, void *workspace,
                               void *stream) {
    B200MED_REQUIRE(M >= 1 && C >= 1, "bad shape");
    B200MED_REQUIRE(x && y, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (!training) {
        B200MED_REQUIRE(running_mean && running_var, "inference needs the running statistics");
        launch_k(bn_eval_kernel, ew_grid(M * C), 256, 0, st, x, M, C, running_mean, running_var, gamma, beta, eps, y);
        return after_launch("bn_eval_kernel");
    }
    B200MED_REQUIRE(save_mean && save_rstd && workspace, "training needs save_mean, save_rstd and a workspace");
    const int S = (int)bn_slabs(M);
    const long long rows = (M + S - 1) / S;
    float *part = reinterpret_cast<float *>(workspace);
    float *scale_shift = part + (long long)S * 3 * C;
    const bool vec = C % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
    const int V = vec ? 4 : 1;
    dim3 grid((unsigned)((C + kBnCols * V - 1) / (kBnCols * V)), (unsigned)S);
    if (vec) launch_k(bn_stats_partial_kernel<4>, grid, kBnCols * kBnWarps, 0, st, x, M, C, rows, part);
    else launch_k(bn_stats_partial_kernel<1>, grid, kBnCols * kBnWarps, 0, st, x, M, C, rows, part);
    if (int e = after_launch("bn_stats_partial_kernel")) return e;
    launch_k(bn_finalize_kernel, (C + 7) / 8, 256, 0, st, part, S, M, C, gamma, beta, eps, momentum, scale_shift, save_mean, save_rstd,
                                                      running_mean, running_var, (long long *)num_batches_tracked);
    if (int e = after_launch("bn_finalize_kernel")) return e;
    if (vec) launch_k(bn_apply_kernel<4>, ew_grid(M * C / 4), 256, 0, st, x, M, C, scale_shift, y);
    else launch_k(bn_apply_kernel<1>, ew_grid(M * C), 256, 0, st, x, M, C, scale_shift, y);
    return after_launch("bn_apply_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_bn_bwd(const float *dy, const float *x, int64_t M, int32_t C, const float *gamma,
                               const float *save_mean, const float *save_rstd, int32_t relu_mask, float *dx, float *dgamma,
                               float *dbeta, void *workspace, void *stream) {
    B200MED_REQUIRE(M >= 1 && C >= 1, "bad shape");
    B200MED_REQUIRE(dy && x && save_mean && save_rstd && dx && workspace, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int S = (int)bn_slabs(M);
    const long long rows = (M + S - 1) / S;
    float *part = reinterpret_cast<float *>(workspace);
    float *coef = part + (long long)S * 3 * C;
    const bool vec = C % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0);
    const int V = vec ? 4 : 1;
    dim3 grid((unsigned)((C + kBnCols * V - 1) / (kBnCols * V)), (unsigned)S);
    if (vec) launch_k(bn_bwd_partial_kernel<4>, grid, kBnCols * kBnWarps, 0, st, dy, x, M, C, rows, save_mean, save_rstd, part);
    else launch_k(bn_bwd_partial_kernel<1>, grid, kBnCols * kBnWarps, 0, st, dy, x, M, C, rows, save_mean, save_rstd, part);
    if (int e = after_launch("bn_bwd_partial_kernel")) return e;
    launch_k(bn_bwd_finalize_kernel, (C + 7) / 8, 256, 0, st, part, S, M, C, gamma, save_mean, save_rstd, coef, dgamma, dbeta);
    if (int e = after_launch("bn_bwd_finalize_kernel")) return e;
    if (vec) launch_k(bn_bwd_apply_kernel<4>, ew_grid(M * C / 4), 256, 0, st, dy, x, M, C, coef, relu_mask, dx);
    else launch_k(bn_bwd_apply_kernel<1>, ew_grid(M * C), 256, 0, st, dy, x, M, C, coef, relu_mask, dx);
    return after_launch("bn_bwd_apply_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_pool_drop_fwd(const float *z, float *p, int64_t B, int32_t L, int32_t Lc, int32_t C,
                                      float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(B >= 1 && L >= 2 && Lc >= 2 && Lc <= L && C >= 1 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(z && p, "null pointer");
    const int Lp = Lc / 2;
    launch_k(pool_drop_fwd_kernel, ew_grid(B * (long long)Lp * C), 256, 0, (cudaStream_t)stream, z, p, B, L, Lp, C, drop_p, seed, drop_base);
    return after_launch("pool_drop_fwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_pool_drop_bwd(const float *dp, const float *z, float *dz, int64_t B, int32_t L,
                                      int32_t Lc, int32_t C, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(B >= 1 && L >= 2 && Lc >= 2 && Lc <= L && C >= 1 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(dp && z && dz, "null pointer");
    launch_k(pool_drop_bwd_kernel, ew_grid((B * (long long)L + 2) * C), 256, 0, (cudaStream_t)stream, dp, z, dz, B, L, Lc / 2, C, drop_p, seed,
                                                                                          drop_base);
    return after_launch("pool_drop_bwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_conv_pack(const float *w, float *fwd, float *bwd, int32_t Cout, int32_t Cin, void *stream) {
    B200MED_REQUIRE(Cout >= 1 && Cin >= 1 && w && (fwd || bwd), "bad argument");
    launch_k(conv_pack_kernel, ew_grid((long long)Cout * Cin * 3), 256, 0, (cudaStream_t)stream, w, fwd, bwd, Cout, Cin);
    return after_launch("conv_pack_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_conv_unpack_grad(const float *dfwd, float *dw, int32_t Cout, int32_t Cin, void *stream) {
    B200MED_REQUIRE(Cout >= 1 && Cin >= 1 && dfwd && dw, "bad argument");
    launch_k(conv_unpack_grad_kernel, ew_grid((long long)Cout * Cin * 3), 256, 0, (cudaStream_t)stream, dfwd, dw, Cout, Cin);
    return after_launch("conv_unpack_grad_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_transpose_last2(const float *x, float *y, int64_t B, int32_t R, int32_t C, void *stream) {
    B200MED_REQUIRE(B >= 0 && R >= 1 && C >= 1, "bad shape");
    if (B == 0) return B200MED_OK;
    B200MED_REQUIRE(x && y, "null pointer");
    launch_k(transpose_last2_kernel, ew_grid(B * (long long)R * C), 256, 0, (cudaStream_t)stream, x, y, B, R, C);
    return after_launch("transpose_last2_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_concat2(const float *a, const float *b, float *out, int64_t M, int32_t Ca, int32_t Cb,
                                                                     void *stream) {
    B200MED_REQUIRE(M >= 0 && Ca >= 1 && Cb >= 1, "bad shape");
    if (M == 0) return B200MED_OK;
    B200MED_REQUIRE(a && b && out, "null pointer");
    launch_k(concat2_kernel, ew_grid(M * (Ca + Cb)), 256, 0, (cudaStream_t)stream, a, b, out, M, Ca, Cb);
    return after_launch("concat2_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_slice_cols(const float *x, float *out, int64_t M, int32_t ld, int32_t col0, int32_t C,
                                                                        void *stream) {
    B200MED_REQUIRE(M >= 0 && C >= 1 && col0 >= 0 && col0 + C <= ld, "bad shape");
    if (M == 0) return B200MED_OK;
    B200MED_REQUIRE(x && out, "null pointer");
    launch_k(slice_cols_kernel, ew_grid(M * C), 256, 0, (cudaStream_t)stream, x, out, M, ld, col0, C);
    return after_launch("slice_cols_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_take_rows(const void *src, const int64_t *idx, void *out, int64_t n, int32_t C,
                                                                       void *stream) {
    B200MED_REQUIRE(n >= 0 && C >= 1, "bad shape");
    if (n == 0) return B200MED_OK;
    B200MED_REQUIRE(src && idx && out, "null pointer");
    launch_k(take_rows_kernel, ew_grid(n * C), 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint32_t *>(src),
                                                                        reinterpret_cast<const long long *>(idx),
                                                                        reinterpret_cast<uint32_t *>(out), n, C);
    return after_launch("take_rows_kernel");
}
