// The window heads' Linear / ReLU / BatchNorm1d tail (MED/modeling/models.py:166-186, 204-210: ReLU -> Linear(128, 256) ->
// ReLU -> BatchNorm1d -> Linear(256, 64) -> ReLU -> BatchNorm1d -> Linear(64, C)) as THREE kernels per direction.
//
// Why.  With the layer-at-a-time kernels (csrc/head.cu BatchNorm passes around the GEMMs) the tail of the headline step was 30
// launches / 0.17 ms during which most SMs idle (profiles/r2_step_timeline_final.md): 0.8 GFLOP on 8192 rows is a latency
// chain, not a throughput problem.  BatchNorm needs the statistics of the WHOLE batch, so a kernel boundary sits after every
// ReLU; everything between two boundaries is one kernel here:
//
//   forward   F1  a1 = relu(relu(x) W1^T + b1)                        + per-CTA (n, mean, M2) of a1's columns
//             F2  [finalize BN1] y1 = BN1(a1), a2 = relu(y1 W2^T + b2)  + per-CTA (n, mean, M2) of a2's columns
//             F3  [finalize BN2] y2 = BN2(a2), out = y2 Wl^T + bl
//   backward  B1  per-CTA (sum g2, sum g2 xhat2), g2 = g Wl                          (g = d loss / d out, C <= 8 columns)
//             B2  [finalize] dz2 = (a2 > 0) BN2'(g2), g1 = dz2 W2                   + per-CTA (sum g1, sum g1 xhat1)
//             B3  [finalize] dz1 = (a1 > 0) BN1'(g1), dx = (x > 0) dz1 W1
//
// One CTA owns 64 batch rows (8192 rows = 128 CTAs), holds the layer's whole weight matrix and its input tile in shared memory
// and computes in fp32 FMAs (an 8 x 8 register tile per thread: the products are 0.4 GFLOP, and fp32 keeps the 1e-5 bar in
// both precision modes -- the split-bf16 tensor-core products this replaces needed three conversion kernels per layer).
// Between two of them a small finalize kernel (one warp per column, lanes over the per-CTA partials, Chan's combination in
// double and in a fixed order like bn_finalize_kernel: deterministic) turns the partials into save_mean / save_rstd / the
// running statistics (forward) or dgamma / dbeta (backward); every CTA of the consumer derives its normalisation coefficients
// from those 2 x K floats.  (First version: every consumer CTA combined the 128 partials itself -- a serial chain of L2 round
// trips, 18 us per kernel, and 50 MB of redundant L2 reads at K = 256.)  The weight and bias gradients (dz^T y, column sums of
// dz) stay on the side stream (heads.py), off the critical path.
#include "common.cuh"

namespace b200med {

constexpr int kTailThreads = 256;
constexpr int kTailRows = 64;
constexpr int kTailMaxC = 8;       // columns of the output layer

struct TailBn {
    const float *gamma, *beta;
    float eps;
    const float *running_mean, *running_var;
    const float *save_mean, *save_rstd;   // written by tail_bn_finalize_kernel in front of this kernel (mode 1)
    int mode;                             // 0 no BatchNorm in front, 1 batch statistics (training), 2 running statistics
};

// part [S][3][K] = (n, mean, M2) per CTA of the producing kernel -> save_mean / save_rstd, running statistics (momentum,
// unbiased variance), *num_batches += 1: nn.BatchNorm1d in training mode.  One warp per column, lanes over the partials.
__global__ void __launch_bounds__(256)
tail_bn_finalize_kernel(const float *__restrict__ part, int S, long long M, int K, float eps, float momentum,
                        float *__restrict__ save_mean, float *__restrict__ save_rstd, float *__restrict__ running_mean,
                        float *__restrict__ running_var, long long *__restrict__ num_batches) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col == 0 && lane == 0 && num_batches) *num_batches += 1;
    if (col >= K) return;
    // one pass: every lane merges its partials (Chan), then a butterfly over the lanes -- a fixed tree, so deterministic
    double n = 0.0, mu = 0.0, q = 0.0;
#pragma unroll 4
    for (int s = lane; s < S; s += 32) {
        const float *p = part + (long long)s * 3 * K;
        const double n2 = (double)p[col], m2 = (double)p[K + col], q2 = (double)p[2 * K + col];
        const double nn = n + n2, d = m2 - mu;
        mu += d * (n2 / nn);
        q += q2 + d * d * (n * n2 / nn);
        n = nn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double n2 = __shfl_xor_sync(0xffffffffu, n, o), m2 = __shfl_xor_sync(0xffffffffu, mu, o), q2 = __shfl_xor_sync(0xffffffffu, q, o);
        const double nn = n + n2;
        if (nn > 0.0) {
            // the pair is merged in lane order (lower lane first) on both sides, so all lanes hold the same bits
            const bool low = (lane & o) == 0;
            const double na = low ? n : n2, ma = low ? mu : m2, nb = low ? n2 : n, mb = low ? m2 : mu;
            const double d = mb - ma;
            mu = ma + d * (nb / nn);
            q = q + q2 + d * d * (na * nb / nn);
        }
        n = nn;
    }
    if (lane != 0) return;
    const double var = q / (double)M;          // n == M
    save_mean[col] = (float)mu;
    save_rstd[col] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = M > 1 ? q / (double)(M - 1) : var;
        running_mean[col] = (1.0f - momentum) * running_mean[col] + momentum * (float)mu;
        running_var[col] = (1.0f - momentum) * running_var[col] + momentum * (float)unbiased;
    }
}

// part [S][2][K] = (sum dy, sum dy xhat) per CTA -> dbeta, dgamma
__global__ void __launch_bounds__(256)
tail_bn_bwd_finalize_kernel(const float *__restrict__ part, int S, int K, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= K) return;
    double sa = 0.0, sb = 0.0;
    for (int s = lane; s < S; s += 32) { sa += (double)part[(long long)s * 2 * K + col]; sb += (double)part[(long long)s * 2 * K + K + col]; }
    sa = warp_sum(sa); sb = warp_sum(sb);
    if (lane != 0) return;
    dbeta[col] = (float)sa;
    dgamma[col] = (float)sb;
}

// Normalisation coefficients of the K input columns into c_s [4][K] (bn_finalize_kernel's / bn_eval_kernel's arithmetic)
__device__ __forceinline__ void tail_bn_coefs(const TailBn &bn, int K, float *c_s) {
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        const float g = bn.gamma ? bn.gamma[c] : 1.0f, b = bn.beta ? bn.beta[c] : 0.0f;
        if (bn.mode == 1) {
            const float mu = bn.save_mean[c], rstd = bn.save_rstd[c];
            c_s[c] = rstd * g;
            c_s[K + c] = b - mu * rstd * g;
        } else {
            c_s[c] = 1.0f / sqrtf(bn.running_var[c] + bn.eps);
            c_s[K + c] = bn.running_mean[c];
            c_s[2 * K + c] = g;
            c_s[3 * K + c] = b;
        }
    }
}

// y = BatchNorm(x) for four consecutive columns k .. k + 3 (k % 4 == 0): the coefficients by 16-byte shared loads (scalar loads at a stride of four
// floats across the lanes were 4-way bank conflicts: 14 % of the wavefronts of the second layer's kernel)
__device__ __forceinline__ void tail_bn_apply4(int mode, const float *c_s, int K, int k, float4 &v) {
    const float4 a = *reinterpret_cast<const float4 *>(c_s + k), b = *reinterpret_cast<const float4 *>(c_s + K + k);
    if (mode == 1) {
        v.x = fmaf(v.x, a.x, b.x); v.y = fmaf(v.y, a.y, b.y); v.z = fmaf(v.z, a.z, b.z); v.w = fmaf(v.w, a.w, b.w);
    } else {
        const float4 g = *reinterpret_cast<const float4 *>(c_s + 2 * K + k), h = *reinterpret_cast<const float4 *>(c_s + 3 * K + k);
        v.x = (v.x - b.x) * a.x * g.x + h.x; v.y = (v.y - b.y) * a.y * g.y + h.y;
        v.z = (v.z - b.z) * a.z * g.z + h.z; v.w = (v.w - b.w) * a.w * g.w + h.w;
    }
}

// acc[i][4j + q] = sum_k in_s[row_i][k] * w_s[k][col_j + q] over one 64-row tile.  8 warps as WARPS_M x WARPS_N; a warp's
// lanes as 4 (rows) x 8 (column quads): row_i = wm * 4TR + lr + 4i, col_j = wn * 8TC + 4lc + 32j.  Both operand reads are
// 16-byte shared loads without bank conflicts (rows of in_s are K + 4 floats apart; a w_s row is read contiguously).
template <int N, int TR, int TC>
__device__ __forceinline__ void tail_tile_gemm(const float *__restrict__ in_s, int ldi, const float *__restrict__ w_s, int K,
                                               float (&acc)[TR][TC]) {
    constexpr int WC = 8 * TC, WR = 4 * TR, WARPS_N = N / WC;
    static_assert(TC % 4 == 0 && N % WC == 0 && (8 / WARPS_N) * WR == kTailRows && 8 % WARPS_N == 0, "tile shape");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane >> 3, lc = lane & 7;
    const int wn = warp % WARPS_N, wm = warp / WARPS_N;
    const float *xp = in_s + (wm * WR + lr) * ldi;
    const float *wp = w_s + wn * WC + 4 * lc;
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = 0.0f;
#pragma unroll 1
    for (int k = 0; k < K; k += 4) {
        float4 xv[TR];
#pragma unroll
        for (int i = 0; i < TR; ++i) xv[i] = *reinterpret_cast<const float4 *>(xp + 4 * i * ldi + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 wv[TC / 4];
#pragma unroll
            for (int j = 0; j < TC / 4; ++j) wv[j] = *reinterpret_cast<const float4 *>(wp + (k + kk) * N + 32 * j);
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                const float xs = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
#pragma unroll
                for (int j = 0; j < TC / 4; ++j) {
                    acc[i][4 * j + 0] = fmaf(xs, wv[j].x, acc[i][4 * j + 0]);
                    acc[i][4 * j + 1] = fmaf(xs, wv[j].y, acc[i][4 * j + 1]);
                    acc[i][4 * j + 2] = fmaf(xs, wv[j].z, acc[i][4 * j + 2]);
                    acc[i][4 * j + 3] = fmaf(xs, wv[j].w, acc[i][4 * j + 3]);
                }
            }
        }
    }
}

__device__ __forceinline__ void tail_cp16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void tail_cp_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// The same product with the weights in their stored layout, w_s [N][ldw] (a Linear layer's W [N, K], rows K + 4 floats apart):
// a thread's columns are col_j = wn * 8TC + lc + 8j, so the 8 column lanes of a warp read 8 consecutive rows of w_s -- 16 bytes
// each, 32 distinct banks.  No transposition of W on its way into shared memory: it arrives by cp.async.
template <int N, int TR, int TC>
__device__ __forceinline__ void tail_tile_gemm_nmajor(const float *__restrict__ in_s, int ldi, const float *__restrict__ w_s, int ldw,
                                                      int K, float (&acc)[TR][TC]) {
    constexpr int WC = 8 * TC, WR = 4 * TR, WARPS_N = N / WC;
    static_assert(N % WC == 0 && (8 / WARPS_N) * WR == kTailRows && 8 % WARPS_N == 0, "tile shape");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane >> 3, lc = lane & 7;
    const int wn = warp % WARPS_N, wm = warp / WARPS_N;
    const float *xp = in_s + (wm * WR + lr) * ldi;
    const float *wp = w_s + (wn * WC + lc) * ldw;
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 xv[TR], wv[TC];
#pragma unroll
        for (int i = 0; i < TR; ++i) xv[i] = *reinterpret_cast<const float4 *>(xp + 4 * i * ldi + k);
#pragma unroll
        for (int j = 0; j < TC; ++j) wv[j] = *reinterpret_cast<const float4 *>(wp + 8 * j * ldw + k);
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(xv[i].x, wv[j].x, acc[i][j]);
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(xv[i].y, wv[j].y, acc[i][j]);
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(xv[i].z, wv[j].z, acc[i][j]);
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(xv[i].w, wv[j].w, acc[i][j]);
    }
}

// ------------------------------------------------------------------------------------------------ forward, hidden layer
// a [M, N] = relu(in' W^T + b), in' = relu(in) (relu_in) or BatchNorm(in) (bn.mode; y [M, K] = in' is written for the
// backward); part [CTAs][3][N] = (n, mean, M2) of a's columns over this CTA's rows (NULL: not wanted).
template <int N, int TR, int TC>
__global__ void __launch_bounds__(kTailThreads, 1)
tail_fwd_hidden_kernel(const float *__restrict__ in, long long M, int K, int relu_in, TailBn bn, float *__restrict__ y,
                       const float *__restrict__ W, const float *__restrict__ bias, float *__restrict__ a,
                       float *__restrict__ part) {
    pdl_wait();
    extern __shared__ __align__(16) float tail_smem[];
    const int ldi = K + 4, k4n = K >> 2, tid = threadIdx.x;
    float *c_s = tail_smem;                  // [4][K]
    float *in_s = c_s + 4 * K;               // [64][K + 4]
    float *w_s = in_s + kTailRows * ldi;     // [N][K + 4]; after the product: the output tile [64][N + 8]
    const long long r0 = (long long)blockIdx.x * kTailRows;
    const int nvalid = (int)min((long long)kTailRows, M - r0);
    // both operands by cp.async (no registers, everything in flight at once): W [N, K] as stored, the CTA's 64 input rows
    for (int idx = tid; idx < N * k4n; idx += kTailThreads) {
        const int n = idx / k4n, k = (idx - n * k4n) * 4;
        tail_cp16(w_s + n * ldi + k, W + (long long)n * K + k);
    }
    for (int idx = tid; idx < kTailRows * k4n; idx += kTailThreads) {
        const int r = idx / k4n, k = (idx - r * k4n) * 4;
        if (r < nvalid) tail_cp16(in_s + r * ldi + k, in + (r0 + r) * K + k);
        else *reinterpret_cast<float4 *>(in_s + r * ldi + k) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (bn.mode) tail_bn_coefs(bn, K, c_s);
    tail_cp_wait();
    __syncthreads();
    if (bn.mode || relu_in) {                 // the input transform, in place (each thread the chunks it copied itself)
        for (int idx = tid; idx < nvalid * k4n; idx += kTailThreads) {
            const int r = idx / k4n, k = (idx - r * k4n) * 4;
            float4 v = *reinterpret_cast<const float4 *>(in_s + r * ldi + k);
            if (bn.mode) {
                tail_bn_apply4(bn.mode, c_s, K, k, v);
                if (y) *reinterpret_cast<float4 *>(y + (r0 + r) * K + k) = v;
            } else {
                v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
            }
            *reinterpret_cast<float4 *>(in_s + r * ldi + k) = v;
        }
        __syncthreads();
    }
    float acc[TR][TC];
    tail_tile_gemm_nmajor<N, TR, TC>(in_s, ldi, w_s, ldi, K, acc);
    __syncthreads();                          // every warp is done with w_s: it becomes the output tile
    constexpr int WC = 8 * TC, WR = 4 * TR, WARPS_N = N / WC, lds = N + 8;
    const int lane = tid & 31, warp = tid >> 5, lr = lane >> 3, lc = lane & 7;
    const int wn = warp % WARPS_N, wm = warp / WARPS_N;
    float *st = w_s;
#pragma unroll
    for (int j = 0; j < TC; ++j) {
        const int col = wn * WC + lc + 8 * j;
        const float bv = bias ? bias[col] : 0.0f;
#pragma unroll
        for (int i = 0; i < TR; ++i) st[(wm * WR + lr + 4 * i) * lds + col] = fmaxf(acc[i][j] + bv, 0.0f);
    }
    __syncthreads();
    constexpr int n4 = N / 4;
    for (int idx = tid; idx < nvalid * n4; idx += kTailThreads) {
        const int r = idx / n4, c = (idx - r * n4) * 4;
        *reinterpret_cast<float4 *>(a + (r0 + r) * N + c) = *reinterpret_cast<const float4 *>(st + r * lds + c);
    }
    if (!part) return;
    for (int c = tid; c < N; c += kTailThreads) {      // local two-pass over this CTA's rows
        float s = 0.0f;
#pragma unroll 8
        for (int r = 0; r < nvalid; ++r) s += st[r * lds + c];
        const float mu = s / (float)nvalid;
        float q = 0.0f;
#pragma unroll 8
        for (int r = 0; r < nvalid; ++r) { const float d = st[r * lds + c] - mu; q = fmaf(d, d, q); }
        float *p = part + (long long)blockIdx.x * 3 * N;
        p[c] = (float)nvalid; p[N + c] = mu; p[2 * N + c] = q;
    }
}

// ------------------------------------------------------------------------------------------------ forward, output layer
// y = BatchNorm(in) (written for the backward when y != NULL), out [M, C] = y Wl^T + bl, C <= 8.  Four threads per row.
__global__ void __launch_bounds__(kTailThreads, 1)
tail_fwd_out_kernel(const float *__restrict__ in, long long M, int K, int relu_in, TailBn bn, float *__restrict__ y,
                    const float *__restrict__ W, const float *__restrict__ bias, int C, float *__restrict__ out) {
    pdl_wait();
    extern __shared__ __align__(16) float tail_smem[];
    const int ldi = K + 4, k4n = K >> 2, tid = threadIdx.x;
    float *c_s = tail_smem;                  // [4][K]
    float *in_s = c_s + 4 * K;               // [64][K + 4]
    float *w_s = in_s + kTailRows * ldi;     // [C][K]
    const long long r0 = (long long)blockIdx.x * kTailRows;
    const int nvalid = (int)min((long long)kTailRows, M - r0);
    if (bn.mode) tail_bn_coefs(bn, K, c_s);
    for (int idx = tid; idx < C * K; idx += kTailThreads) w_s[idx] = W[idx];
    if (bn.mode) __syncthreads();
#pragma unroll 4
    for (int idx = tid; idx < kTailRows * k4n; idx += kTailThreads) {
        const int r = idx / k4n, k = (idx - r * k4n) * 4;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (r < nvalid) {
            v = *reinterpret_cast<const float4 *>(in + (r0 + r) * K + k);
            if (bn.mode) {
                tail_bn_apply4(bn.mode, c_s, K, k, v);
                if (y) *reinterpret_cast<float4 *>(y + (r0 + r) * K + k) = v;
            } else if (relu_in) {
                v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
            }
        }
        *reinterpret_cast<float4 *>(in_s + r * ldi + k) = v;
    }
    __syncthreads();
    const int row = tid >> 2, q = tid & 3;
    for (int c = 0; c < C; ++c) {
        float p = 0.0f;
        for (int k = q; k < K; k += 4) p = fmaf(in_s[row * ldi + k], w_s[c * K + k], p);
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        if (q == 0 && row < nvalid) out[(r0 + row) * C + c] = p + (bias ? bias[c] : 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------ backward, output layer
// part [CTAs][2][N] = (sum g2, sum g2 xhat) over this CTA's rows, g2 = g [M, C] Wl [C, N], xhat = (a - mean) rstd:
// the two sums the backward of the LAST BatchNorm needs.  N in {64, 128, 256}: 256 / N row groups x N columns.
template <int N>
__global__ void __launch_bounds__(kTailThreads)
tail_bwd_out_kernel(const float *__restrict__ g, int C, const float *__restrict__ Wl, const float *__restrict__ a, long long M,
                    const float *__restrict__ save_mean, const float *__restrict__ save_rstd, float *__restrict__ part) {
    pdl_wait();
    constexpr int G = kTailThreads / N, T = kTailRows / G;
    __shared__ float g_s[kTailRows * kTailMaxC];
    __shared__ float red[2][G][N];
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * kTailRows;
    const int nvalid = (int)min((long long)kTailRows, M - r0);
    for (int idx = tid; idx < kTailRows * C; idx += kTailThreads) g_s[idx] = idx < nvalid * C ? g[r0 * C + idx] : 0.0f;
    const int n = tid % N, rg = tid / N;
    float av[T];                      // this thread's column of the BatchNorm input: all loads in flight at once
#pragma unroll
    for (int t = 0; t < T; ++t) { const int r = rg + t * G; av[t] = r < nvalid ? a[(r0 + r) * N + n] : 0.0f; }
    float wl[kTailMaxC];
#pragma unroll
    for (int c = 0; c < kTailMaxC; ++c) wl[c] = c < C ? Wl[c * N + n] : 0.0f;
    const float mu = save_mean[n], rs = save_rstd[n];
    __syncthreads();
    float sa = 0.0f, sb = 0.0f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int r = rg + t * G;
        float g2 = 0.0f;
#pragma unroll
        for (int c = 0; c < kTailMaxC; ++c) if (c < C) g2 = fmaf(g_s[r * C + c], wl[c], g2);      // zero on rows past the batch
        sa += g2;
        sb = fmaf(g2, (av[t] - mu) * rs, sb);
    }
    red[0][rg][n] = sa; red[1][rg][n] = sb;
    __syncthreads();
    if (rg == 0) {
        float ta = 0.0f, tb = 0.0f;
#pragma unroll
        for (int k = 0; k < G; ++k) { ta += red[0][k][n]; tb += red[1][k][n]; }
        float *p = part + (long long)blockIdx.x * 2 * N;
        p[n] = ta; p[N + n] = tb;
    }
}

// ------------------------------------------------------------------------------------------------ backward, hidden layer
// The layer is  in -> Linear(W [K, N]... stored [K rows = its outputs][N = its inputs]) -> ReLU (= a [M, K]) -> BatchNorm.
//   dy [M, K]   gradient of the BatchNorm's output: read from `dy`, or (FIRST: the last hidden layer) g [M, C] Wl [C, K]
//   dz [M, K] = (a > 0) (k0 dy - k1 - k2 a)      (BatchNorm backward from the column sums dbeta / dgamma, then the ReLU mask) OUT
//   dx [M, N] = dz W                              OUT (zeroed where relu_mask <= 0 when given: the ReLU in front of the tail)
//   part_prev [CTAs][2][N] = (sum dx, sum dx xhat_prev) for the BatchNorm in front of this layer (a_prev != NULL)
template <int N, int TR, int TC, bool FIRST>
__global__ void __launch_bounds__(kTailThreads, 1)
tail_bwd_hidden_kernel(const float *__restrict__ dy, const float *__restrict__ g, int C, const float *__restrict__ Wl,
                       const float *__restrict__ a, long long M, int K,
                       const float *__restrict__ gamma, const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
                       float *__restrict__ dz, const float *__restrict__ dgamma, const float *__restrict__ dbeta,
                       const float *__restrict__ W, float *__restrict__ dx, const float *__restrict__ a_prev,
                       const float *__restrict__ mean_prev, const float *__restrict__ rstd_prev, float *__restrict__ part_prev,
                       const float *__restrict__ relu_mask) {
    pdl_wait();
    extern __shared__ __align__(16) float tail_smem[];
    const int ldi = K + 4, k4n = K >> 2, tid = threadIdx.x;
    float *c_s = tail_smem;                  // [3][K]: k0, k1, k2   (+ [K] unused)
    float *in_s = c_s + 4 * K;               // [64][K + 4]: dz tile
    float *w_s = in_s + kTailRows * ldi;     // [K][N]
    float *g_s = w_s + K * N;                // FIRST: [64][C] then Wl [C][K]
    const long long r0 = (long long)blockIdx.x * kTailRows;
    const int nvalid = (int)min((long long)kTailRows, M - r0);
    // W [K, N] row-major is already the [k][n] operand: cp.async, in flight under the construction of the dz tile
    for (int idx = tid; idx < K * (N >> 2); idx += kTailThreads) tail_cp16(w_s + 4 * idx, W + 4 * idx);
    // BatchNorm backward coefficients (bn_bwd_finalize_kernel's arithmetic) from the column sums tail_bn_bwd_finalize_kernel left
    for (int c = tid; c < K; c += kTailThreads) {
        const float mu = save_mean[c], rs = save_rstd[c], k0 = (gamma ? gamma[c] : 1.0f) * rs;
        const float ma = (float)((double)dbeta[c] / (double)M), mb = (float)((double)dgamma[c] / (double)M);
        c_s[c] = k0;
        c_s[K + c] = k0 * (ma - mu * rs * mb);
        c_s[2 * K + c] = k0 * rs * mb;
    }
    if (FIRST) {
        for (int idx = tid; idx < nvalid * C; idx += kTailThreads) g_s[idx] = g[r0 * C + idx];
        for (int idx = tid; idx < C * K; idx += kTailThreads) g_s[kTailRows * kTailMaxC + idx] = Wl[idx];
    }
    __syncthreads();
    const float *wl_s = g_s + kTailRows * kTailMaxC;
#pragma unroll 4
    for (int idx = tid; idx < kTailRows * k4n; idx += kTailThreads) {
        const int r = idx / k4n, k = (idx - r * k4n) * 4;
        float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (r < nvalid) {
            const float4 av4 = *reinterpret_cast<const float4 *>(a + (r0 + r) * K + k);
            const float av[4] = {av4.x, av4.y, av4.z, av4.w};
            float gv[4];
            if (FIRST) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t = 0.0f;
                    for (int c = 0; c < C; ++c) t = fmaf(g_s[r * C + c], wl_s[c * K + k + e], t);
                    gv[e] = t;
                }
            } else {
                const float4 d4 = *reinterpret_cast<const float4 *>(dy + (r0 + r) * K + k);
                gv[0] = d4.x; gv[1] = d4.y; gv[2] = d4.z; gv[3] = d4.w;
            }
            const float4 q0 = *reinterpret_cast<const float4 *>(c_s + k), q1 = *reinterpret_cast<const float4 *>(c_s + K + k),
                         q2 = *reinterpret_cast<const float4 *>(c_s + 2 * K + k);
            const float k0[4] = {q0.x, q0.y, q0.z, q0.w}, k1[4] = {q1.x, q1.y, q1.z, q1.w}, k2[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = fmaf(k0[e], gv[e], -k1[e]) - k2[e] * av[e];
                if (!(av[e] > 0.0f)) v = 0.0f;
                o[e] = v;
            }
            *reinterpret_cast<float4 *>(dz + (r0 + r) * K + k) = make_float4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<float4 *>(in_s + r * ldi + k) = make_float4(o[0], o[1], o[2], o[3]);
    }
    tail_cp_wait();
    __syncthreads();
    float acc[TR][TC];
    tail_tile_gemm<N, TR, TC>(in_s, ldi, w_s, K, acc);
    constexpr int WC = 8 * TC, WR = 4 * TR, WARPS_N = N / WC;
    const int lane = tid & 31, warp = tid >> 5, lr = lane >> 3, lc = lane & 7;
    const int wn = warp % WARPS_N, wm = warp / WARPS_N;
    __shared__ float red[8 / WARPS_N][2][N];
    float sa[TC], sb[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) { sa[t] = 0.0f; sb[t] = 0.0f; }
#pragma unroll
    for (int j = 0; j < TC / 4; ++j) {
        const int col = wn * WC + 4 * lc + 32 * j;
        float mu[4] = {0.0f, 0.0f, 0.0f, 0.0f}, rs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (part_prev) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { mu[e] = mean_prev[col + e]; rs[e] = rstd_prev[col + e]; }
        }
        float4 ap[TR];                 // the rows' BatchNorm inputs in front of this layer: all loads in flight at once
        if (part_prev) {
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                const int row = wm * WR + lr + 4 * i;
                ap[i] = row < nvalid ? *reinterpret_cast<const float4 *>(a_prev + (r0 + row) * N + col) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            const int row = wm * WR + lr + 4 * i;
            float4 v = make_float4(acc[i][4 * j + 0], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
            if (row >= nvalid) continue;         // rows past the batch: dz is zero there, so v is zero too; nothing to store
            if (relu_mask) {
                const float4 m = *reinterpret_cast<const float4 *>(relu_mask + (r0 + row) * N + col);
                if (!(m.x > 0.0f)) v.x = 0.0f;
                if (!(m.y > 0.0f)) v.y = 0.0f;
                if (!(m.z > 0.0f)) v.z = 0.0f;
                if (!(m.w > 0.0f)) v.w = 0.0f;
            }
            *reinterpret_cast<float4 *>(dx + (r0 + row) * N + col) = v;
            if (part_prev) {
                sa[4 * j + 0] += v.x; sb[4 * j + 0] = fmaf(v.x, (ap[i].x - mu[0]) * rs[0], sb[4 * j + 0]);
                sa[4 * j + 1] += v.y; sb[4 * j + 1] = fmaf(v.y, (ap[i].y - mu[1]) * rs[1], sb[4 * j + 1]);
                sa[4 * j + 2] += v.z; sb[4 * j + 2] = fmaf(v.z, (ap[i].z - mu[2]) * rs[2], sb[4 * j + 2]);
                sa[4 * j + 3] += v.w; sb[4 * j + 3] = fmaf(v.w, (ap[i].w - mu[3]) * rs[3], sb[4 * j + 3]);
            }
        }
    }
    if (!part_prev) return;
    // column sums: the thread's TR rows (above), the warp's 4 row lanes (shuffles), the CTA's warp rows (shared memory), always
    // in the same order
#pragma unroll
    for (int t = 0; t < TC; ++t) {
        sa[t] += __shfl_xor_sync(0xffffffffu, sa[t], 8);  sb[t] += __shfl_xor_sync(0xffffffffu, sb[t], 8);
        sa[t] += __shfl_xor_sync(0xffffffffu, sa[t], 16); sb[t] += __shfl_xor_sync(0xffffffffu, sb[t], 16);
    }
    if (lr == 0) {
#pragma unroll
        for (int t = 0; t < TC; ++t) {
            const int col = wn * WC + 4 * lc + 32 * (t / 4) + (t % 4);
            red[wm][0][col] = sa[t]; red[wm][1][col] = sb[t];
        }
    }
    __syncthreads();
    for (int c = tid; c < N; c += kTailThreads) {
        float ta = 0.0f, tb = 0.0f;
#pragma unroll
        for (int w = 0; w < 8 / WARPS_N; ++w) { ta += red[w][0][c]; tb += red[w][1][c]; }
        float *p = part_prev + (long long)blockIdx.x * 2 * N;
        p[c] = ta; p[N + c] = tb;
    }
}

static size_t tail_smem_bytes(int K, int N, bool first) {          // backward hidden kernel
    long long f = 4LL * K + (long long)kTailRows * (K + 4) + (long long)K * N;
    if (first) f += kTailRows * kTailMaxC + (long long)kTailMaxC * K;
    return (size_t)f * 4;
}
static size_t tail_fwd_smem_bytes(int K, int N) {
    const long long w = (long long)N * (K + 4) > (long long)kTailRows * (N + 8) ? (long long)N * (K + 4) : (long long)kTailRows * (N + 8);
    return (size_t)(4LL * K + (long long)kTailRows * (K + 4) + w) * 4;
}
constexpr size_t kTailSmemMax = 223 * 1024;      // dynamic part; the kernels hold up to 4 KB of static shared memory

static bool tail_width_ok(int N) { return N == 64 || N == 128 || N == 256; }
static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

template <typename Kern>
static int tail_set_smem(Kern kern, size_t bytes) {
    return check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "cudaFuncSetAttribute");
}

}  // namespace b200med

using namespace b200med;

#define B200MED_API extern "C" __attribute__((visibility("default")))

B200MED_API int32_t b200med_tail_supported(int32_t K, int32_t N, int32_t kind) {
    // kind 0: forward hidden layer K -> N; 1: backward hidden layer (K = the layer's outputs, N = its inputs); 2: output layer
    // K -> N (N <= 8); 3: kind 1 for the LAST hidden layer (its dy is g Wl)
    if (K < 4 || K % 4 != 0 || K > 1024) return 0;
    if (kind == 2) return N >= 1 && N <= kTailMaxC && (4LL * K + (long long)kTailRows * (K + 4) + (long long)N * K) * 4 <= (long long)kTailSmemMax;
    if (!tail_width_ok(N)) return 0;
    if (kind == 3 && !tail_width_ok(K)) return 0;          // tail_bwd_out_kernel's thread layout over the last hidden width
    return (kind == 0 ? tail_fwd_smem_bytes(K, N) : tail_smem_bytes(K, N, kind == 3)) <= kTailSmemMax ? 1 : 0;
}

B200MED_API int64_t b200med_tail_slabs(int64_t M) { return (M + kTailRows - 1) / kTailRows; }

static TailBn make_bn(int32_t mode, const float *gamma, const float *beta, float eps, const float *rm, const float *rv,
                      const float *sm, const float *sr) {
    TailBn bn;
    bn.gamma = gamma; bn.beta = beta; bn.eps = eps; bn.running_mean = rm; bn.running_var = rv; bn.save_mean = sm; bn.save_rstd = sr;
    bn.mode = mode;
    return bn;
}

// bn_mode 1: the finalize kernel in front of the consumer
static int tail_bn_front(int32_t mode, const float *part, int64_t M, int32_t K, float eps, float momentum, float *rm, float *rv,
                         int64_t *nbt, float *sm, float *sr, cudaStream_t st) {
    B200MED_REQUIRE(mode >= 0 && mode <= 2, "bn_mode must be 0, 1 or 2");
    if (mode == 2) B200MED_REQUIRE(rm && rv, "inference needs the running statistics");
    if (mode != 1) return B200MED_OK;
    B200MED_REQUIRE(part && sm && sr, "batch statistics need the partials and save_mean / save_rstd");
    launch_k(tail_bn_finalize_kernel, (unsigned)((K + 7) / 8), 256, 0, st, part, (int)((M + kTailRows - 1) / kTailRows), (long long)M,
             K, eps, momentum, sm, sr, rm, rv, (long long *)nbt);
    return after_launch("tail_bn_finalize_kernel");
}

B200MED_API int b200med_tail_fwd_hidden(const float *in, int64_t M, int32_t K, int32_t relu_in, int32_t bn_mode, const float *bn_part,
                                        const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                        float *running_var, int64_t *num_batches_tracked, float *save_mean, float *save_rstd,
                                        float *y, const float *W, const float *b, int32_t N, float *a, float *part, void *stream) {
    B200MED_REQUIRE(M >= 1, "bad shape");
    B200MED_REQUIRE(b200med_tail_supported(K, N, 0), "unsupported layer shape (b200med_tail_supported)");
    B200MED_REQUIRE(in && W && a, "null pointer");
    B200MED_REQUIRE(aligned16(in) && aligned16(W) && aligned16(a) && aligned16(y), "pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = tail_bn_front(bn_mode, bn_part, M, K, eps, momentum, running_mean, running_var, num_batches_tracked, save_mean,
                              save_rstd, st)) return e;
    const TailBn bn = make_bn(bn_mode, gamma, beta, eps, running_mean, running_var, save_mean, save_rstd);
    const size_t smem = tail_fwd_smem_bytes(K, N);
    const unsigned grid = (unsigned)((M + kTailRows - 1) / kTailRows);
#define TAIL_FWD(NN, TR, TC)                                                                                              \
    do {                                                                                                                  \
        if (int e = tail_set_smem(tail_fwd_hidden_kernel<NN, TR, TC>, smem)) return e;                                    \
        launch_k(tail_fwd_hidden_kernel<NN, TR, TC>, grid, kTailThreads, smem, st, in, (long long)M, K, relu_in, bn, y, W, b, a, part); \
    } while (0)
    if (N == 256) TAIL_FWD(256, 8, 8);
    else if (N == 128) TAIL_FWD(128, 8, 4);
    else TAIL_FWD(64, 4, 4);
#undef TAIL_FWD
    return after_launch("tail_fwd_hidden_kernel");
}

B200MED_API int b200med_tail_fwd_out(const float *in, int64_t M, int32_t K, int32_t relu_in, int32_t bn_mode, const float *bn_part,
                                     const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                     float *running_var, int64_t *num_batches_tracked, float *save_mean, float *save_rstd, float *y,
                                     const float *W, const float *b, int32_t C, float *out, void *stream) {
    B200MED_REQUIRE(M >= 1, "bad shape");
    B200MED_REQUIRE(b200med_tail_supported(K, C, 2), "unsupported layer shape (b200med_tail_supported)");
    B200MED_REQUIRE(in && W && out, "null pointer");
    B200MED_REQUIRE(aligned16(in) && aligned16(y), "pointers must be 16-byte aligned");
    if (int e = tail_bn_front(bn_mode, bn_part, M, K, eps, momentum, running_mean, running_var, num_batches_tracked, save_mean,
                              save_rstd, (cudaStream_t)stream)) return e;
    const TailBn bn = make_bn(bn_mode, gamma, beta, eps, running_mean, running_var, save_mean, save_rstd);
    const size_t smem = (4 * (size_t)K + (size_t)kTailRows * (K + 4) + (size_t)C * K) * 4;
    if (int e = tail_set_smem(tail_fwd_out_kernel, smem)) return e;
    launch_k(tail_fwd_out_kernel, (unsigned)((M + kTailRows - 1) / kTailRows), kTailThreads, smem, (cudaStream_t)stream, in,
             (long long)M, K, relu_in, bn, y, W, b, C, out);
    return after_launch("tail_fwd_out_kernel");
}

B200MED_API int b200med_tail_bwd_out(const float *g, int32_t C, const float *Wl, const float *a, int64_t M, int32_t N,
                                     const float *save_mean, const float *save_rstd, float *part, void *stream) {
    B200MED_REQUIRE(M >= 1 && C >= 1 && C <= kTailMaxC && tail_width_ok(N), "bad shape");
    B200MED_REQUIRE(g && Wl && a && save_mean && save_rstd && part, "null pointer");
    const unsigned grid = (unsigned)((M + kTailRows - 1) / kTailRows);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 256) launch_k(tail_bwd_out_kernel<256>, grid, kTailThreads, 0, st, g, C, Wl, a, (long long)M, save_mean, save_rstd, part);
    else if (N == 128) launch_k(tail_bwd_out_kernel<128>, grid, kTailThreads, 0, st, g, C, Wl, a, (long long)M, save_mean, save_rstd, part);
    else launch_k(tail_bwd_out_kernel<64>, grid, kTailThreads, 0, st, g, C, Wl, a, (long long)M, save_mean, save_rstd, part);
    return after_launch("tail_bwd_out_kernel");
}

B200MED_API int b200med_tail_bwd_hidden(const float *dy, const float *g, int32_t C, const float *Wl, const float *a, int64_t M,
                                        int32_t K, const float *part, const float *gamma, const float *save_mean,
                                        const float *save_rstd, float *dz, float *dgamma, float *dbeta, const float *W, int32_t N,
                                        float *dx, const float *a_prev, const float *mean_prev, const float *rstd_prev,
                                        float *part_prev, const float *relu_mask, void *stream) {
    const bool first = dy == nullptr;
    B200MED_REQUIRE(M >= 1, "bad shape");
    B200MED_REQUIRE(b200med_tail_supported(K, N, first ? 3 : 1), "unsupported layer shape (b200med_tail_supported)");
    B200MED_REQUIRE(a && part && save_mean && save_rstd && dz && dgamma && dbeta && W && dx, "null pointer");
    if (first) B200MED_REQUIRE(g && Wl && C >= 1 && C <= kTailMaxC, "the last hidden layer needs g [M, C] and Wl [C, K], C <= 8");
    if (part_prev) B200MED_REQUIRE(a_prev && mean_prev && rstd_prev, "part_prev needs a_prev, mean_prev, rstd_prev");
    B200MED_REQUIRE(aligned16(dy) && aligned16(a) && aligned16(dz) && aligned16(W) && aligned16(dx) && aligned16(relu_mask),
                    "pointers must be 16-byte aligned");
    const size_t smem = tail_smem_bytes(K, N, first);
    const unsigned grid = (unsigned)((M + kTailRows - 1) / kTailRows);
    cudaStream_t st = (cudaStream_t)stream;
    launch_k(tail_bn_bwd_finalize_kernel, (unsigned)((K + 7) / 8), 256, 0, st, part, (int)grid, K, dgamma, dbeta);
    if (int e = after_launch("tail_bn_bwd_finalize_kernel")) return e;
#define TAIL_BWD(NN, TR, TC, F)                                                                                           \
    do {                                                                                                                  \
        if (int e = tail_set_smem(tail_bwd_hidden_kernel<NN, TR, TC, F>, smem)) return e;                                 \
        launch_k(tail_bwd_hidden_kernel<NN, TR, TC, F>, grid, kTailThreads, smem, st, dy, g, C, Wl, a, (long long)M, K, gamma, \
                 save_mean, save_rstd, dz, dgamma, dbeta, W, dx, a_prev, mean_prev, rstd_prev, part_prev, relu_mask);     \
    } while (0)
    if (first) {
        if (N == 256) TAIL_BWD(256, 8, 8, true);
        else if (N == 128) TAIL_BWD(128, 8, 4, true);
        else TAIL_BWD(64, 4, 4, true);
    } else {
        if (N == 256) TAIL_BWD(256, 8, 8, false);
        else if (N == 128) TAIL_BWD(128, 8, 4, false);
        else TAIL_BWD(64, 4, 4, false);
    }
#undef TAIL_BWD
    return after_launch("tail_bwd_hidden_kernel");
}
