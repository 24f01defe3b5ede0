// LSTM head (MED/modeling/models.py:135-210, nn.LSTM(58, 128, num_layers=3, dropout=.2)) in throughput mode.
//
// cuDNN runs this recurrence as fp32 SIMT "blockPersist" kernels: 23 ms of a 28 ms train step at B=8192
// (profiles/r1_launch_shares_cudnn_lstm.md).  Here every time step of every layer is ONE tcgen05 GEMM
//     gates[B, 4H] = [x_t | h_{t-1}] [B, Kp] * [W_ih | W_hh]^T [4H, Kp]  (+ b_ih + b_hh)
// (K2 kernel, bf16 operands, fp32 accumulation in TMEM) followed by ONE fused cell kernel below; the
// backward is the mirror image plus one big weight-gradient GEMM per layer over all W*B rows.
//
// Buffers are TIME-MAJOR so that each step's operand is a contiguous [B, Kp] matrix:
//   A_l   [W, B, Kp_l] bf16   columns [0,in_l) = layer input x_t, [in_l, in_l+H) = h_{t-1}, rest = 0 padding
//   G_l   [W, B, 4H]   f32    gate pre-activations, overwritten in place by the ACTIVATED gates (i, f, g, o)
//   C_l   [W, B, H]    f32    cell states
//   dG_l  [W, B, 4H]   bf16   gate gradients (operand of the data- and weight-gradient GEMMs)
//   dA_l  [W, B, Kp_l] f32    [dx_t | dh_{t-1}] produced by the data-gradient GEMM
// These cell kernels are HBM-bound: forward 4H*4 B read+write + ~6 H B per (b, t); they move 20-40 MB per
// step at B=8192, i.e. a few microseconds each.
#include "common.cuh"

namespace b200med {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
    // tanh(x) = 2*sigmoid(2x) - 1; |error| ~1e-6, far below the bf16 operand rounding of this mode
    return 2.0f / (1.0f + __expf(-2.0f * x)) - 1.0f;
}

// Counter-based dropout mask: keep(index) is a pure function of (seed, index), so the backward pass
// regenerates it instead of storing it.
__device__ __forceinline__ bool dropout_keep(uint32_t seed, unsigned long long index, float p) {
    uint64_t z = index + 0x9E3779B97F4A7C15ull * (uint64_t)(seed + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);  // 24 random bits -> [0, 1)
    return u >= p;
}

// x [B, F, W] f32 (the reference's [batch, features, time] layout) -> A0 [W, Bpad, Kp] bf16 columns [0, F), rows b < B;
// also zeroes the padding columns [F, hoff) and [hoff+H, Kp) of every step and the h_{-1} columns [hoff, hoff+H) of step 0.
// One block per 4 windows: x[b] (F*W contiguous floats) is read coalesced into shared memory and written back one
// (t, b) row per warp pass (time-major rows, 2 columns per lane).
constexpr int kPackWin = 4;
__global__ void __launch_bounds__(256)
lstm_pack_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ A0, long long B, long long Bpad, int F, int W, int H,
                 int Kp, int hoff) {
    pdl_wait();
    extern __shared__ float sh_pack[];            // [kPackWin][F][W + 1] (padded rows: column reads spread over the banks)
    const int FW = F * W, W1 = W + 1, FW1 = F * W1;
    for (long long b0 = (long long)blockIdx.x * kPackWin; b0 < B; b0 += (long long)gridDim.x * kPackWin) {
        const int nb = (int)min((long long)kPackWin, B - b0);
        __syncthreads();
        for (int e = threadIdx.x; e < nb * FW; e += blockDim.x) {
            const int bl = e / FW, r = e - bl * FW, k = r / W;
            sh_pack[bl * FW1 + k * W1 + (r - k * W)] = x[b0 * FW + e];
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int r = warp; r < nb * W; r += nwarps) {
            const int bl = r / W, t = r - bl * W;
            __nv_bfloat16 *dst = A0 + ((long long)t * Bpad + b0 + bl) * Kp;
            const float *src = sh_pack + bl * FW1 + t;  // element k at src[k*W1]
            for (int k = lane * 2; k < Kp; k += 64) {    // Kp is a multiple of 8: pairs never straddle the end
                const bool x0 = k < F, x1 = k + 1 < F;
                const bool h0 = k >= hoff && k < hoff + H, h1 = k + 1 >= hoff && k + 1 < hoff + H;
                if ((h0 || h1) && t != 0) {
                    // the h_{t-1} columns of steps t > 0 belong to the recurrence: leave them alone
                    if (!h0) dst[k] = __float2bfloat16_rn(x0 ? src[k * W1] : 0.0f);
                    if (!h1) dst[k + 1] = __float2bfloat16_rn(x1 ? src[(k + 1) * W1] : 0.0f);
                } else {
                    *reinterpret_cast<__nv_bfloat162 *>(dst + k) =
                        __floats2bfloat162_rn(x0 ? src[k * W1] : 0.0f, x1 ? src[(k + 1) * W1] : 0.0f);
                }
            }
        }
    }
}

// dx [B, F, W] f32 <- dA0 [W, Bpad, Kp] f32 columns [0, F): the mirror transpose through shared memory.
__global__ void __launch_bounds__(256)
lstm_unpack_kernel(const float *__restrict__ dA0, float *__restrict__ dx, long long B, long long Bpad, int F, int W, int Kp) {
    pdl_wait();
    extern __shared__ float sh_pack[];            // [kPackWin][F][W + 1]
    const int FW = F * W, W1 = W + 1, FW1 = F * W1;
    for (long long b0 = (long long)blockIdx.x * kPackWin; b0 < B; b0 += (long long)gridDim.x * kPackWin) {
        const int nb = (int)min((long long)kPackWin, B - b0);
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int r = warp; r < nb * W; r += nwarps) {
            const int bl = r / W, t = r - bl * W;
            const float *src = dA0 + ((long long)t * Bpad + b0 + bl) * Kp;
            for (int k = lane; k < F; k += 32) sh_pack[bl * FW1 + k * W1 + t] = src[k];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nb * FW; e += blockDim.x) {
            const int bl = e / FW, r = e - bl * FW, k = r / W;
            dx[b0 * FW + e] = sh_pack[bl * FW1 + k * W1 + (r - k * W)];
        }
    }
}

// Same, for a head input that is stored [B, W, F] (the reference builds it as cat(...).permute(0, 2, 1): the permuted tensor
// is a VIEW of this layout): one warp per (b, t) row, no transpose.
__global__ void __launch_bounds__(256)
lstm_pack_bwf_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ A0, long long B, long long Bpad, int F, int W, int H,
                     int Kp, int hoff) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long nrows = B * W, warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const long long b = r / W;
        const int t = (int)(r - b * W);
        const float *src = x + r * F;
        __nv_bfloat16 *dst = A0 + ((long long)t * Bpad + b) * Kp;
        for (int k = lane * 2; k < Kp; k += 64) {
            const bool x0 = k < F, x1 = k + 1 < F;
            const bool h0 = k >= hoff && k < hoff + H, h1 = k + 1 >= hoff && k + 1 < hoff + H;
            if ((h0 || h1) && t != 0) {
                if (!h0) dst[k] = __float2bfloat16_rn(x0 ? src[k] : 0.0f);
                if (!h1) dst[k + 1] = __float2bfloat16_rn(x1 ? src[k + 1] : 0.0f);
            } else {
                *reinterpret_cast<__nv_bfloat162 *>(dst + k) = __floats2bfloat162_rn(x0 ? src[k] : 0.0f, x1 ? src[k + 1] : 0.0f);
            }
        }
    }
}

// The head input of the window path built straight into the LSTM's first operand (define_inputs, MED/modeling/modeling_utils.py:40-47:
// torch.cat((features, kinematics), dim = 2).permute(0, 2, 1), then models.py:204 transposes it back):
//   A0[t][b][0:Ca]      = bf16(feats[b, t, :])                                   (FeatureExtractor output, fp32 [B, W, Ca])
//   A0[t][b][Ca:Ca+Cb]  = bf16((kin_table[starts[b] + t, :] - mean) / std)       (CustomWindowDataset.py:56-60, IEEE subtract / divide)
// plus the zero padding of lstm_pack_bwf_kernel.  One warp per (window, step) row.  Replaces gather (26 columns) + concat + pack.
__global__ void __launch_bounds__(256)
lstm_pack_parts_kernel(const float *__restrict__ feats, int Ca, const float *__restrict__ kin, long long table_rows, int Cb,
                       const float *__restrict__ mean, const float *__restrict__ stdv, int stat_rows, const int32_t *__restrict__ starts,
                       __nv_bfloat16 *__restrict__ A0, long long B, long long Bpad, int W, int H, int Kp, int hoff) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int F = Ca + Cb;
    const long long nrows = B * W, warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const long long b = r / W;
        const int t = (int)(r - b * W);
        const long long row = (long long)starts[b] + t;
        if (row < 0 || row >= table_rows) __trap();          // a window outside the table: the reference raises IndexError
        const float *fs = feats + r * Ca;
        const float *ks = kin + row * Cb;
        const float *mu = mean ? mean + (stat_rows > 1 ? (long long)t * Cb : 0) : nullptr;
        const float *sd = mean ? stdv + (stat_rows > 1 ? (long long)t * Cb : 0) : nullptr;
        __nv_bfloat16 *dst = A0 + ((long long)t * Bpad + b) * Kp;
        for (int k = lane * 2; k < Kp; k += 64) {
            float v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = k + j;
                float x = 0.0f;
                if (c < Ca) x = fs[c];
                else if (c < F) {
                    x = ks[c - Ca];
                    if (mu) x = (x - mu[c - Ca]) / sd[c - Ca];
                }
                v[j] = x;
            }
            const bool h0 = k >= hoff && k < hoff + H, h1 = k + 1 >= hoff && k + 1 < hoff + H;
            if ((h0 || h1) && t != 0) {          // the h_{t-1} columns of the steps t > 0 belong to the recurrence kernel
                if (!h0) dst[k] = __float2bfloat16_rn(v[0]);
                if (!h1) dst[k + 1] = __float2bfloat16_rn(v[1]);
            } else {
                *reinterpret_cast<__nv_bfloat162 *>(dst + k) = __floats2bfloat162_rn(v[0], v[1]);
            }
        }
    }
}

// The common geometry (Ca, Cb even, Ca + Cb <= 64 = hoff, one statistics row): lane l owns columns 2l, 2l + 1 of the 64-column
// x block -- one float2 load (feats or table row), mean / std in registers for the whole kernel, one 128-byte store per row and
// warp; the h_{-1} block is zeroed by the rows of step 0 only.  (The generic kernel above walks Kp columns with a branch per
// element: 52 us at B = 8192, W = 16 against ~80 MB of traffic.)
__device__ __forceinline__ float2 load_pair(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float2 load_pair(const __nv_bfloat16 *p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p)); }

template <typename TF>      // feats f32, or bf16 (the FeatureExtractor's last GEMM writes the rounding this kernel would apply)
__global__ void __launch_bounds__(256)
lstm_pack_parts64_kernel(const TF *__restrict__ feats, int Ca, const float *__restrict__ kin, long long table_rows, int Cb,
                         const float *__restrict__ mean, const float *__restrict__ stdv, const int32_t *__restrict__ starts,
                         __nv_bfloat16 *__restrict__ A0, long long B, long long Bpad, int W, int H, int Kp) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int c = lane * 2;
    const bool is_f = c < Ca, is_k = !is_f && c < Ca + Cb;
    float m0 = 0.0f, m1 = 0.0f, s0 = 1.0f, s1 = 1.0f;
    if (is_k && mean) { m0 = mean[c - Ca]; m1 = mean[c - Ca + 1]; s0 = stdv[c - Ca]; s1 = stdv[c - Ca + 1]; }
    const long long nrows = B * W, warps = (long long)gridDim.x * (blockDim.x >> 5);
    // four rows per warp and pass: the dependent loads (starts -> table row) of all four are in flight before the first store
    // (one row per pass was a latency chain: 25 us for 70 MB at B = 8192, W = 16)
    constexpr int U = 4;
    for (long long r0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r0 < nrows; r0 += U * warps) {
        long long bs[U], rows[U];
        int ts[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * warps;
            bs[u] = r < nrows ? r / W : -1;
            ts[u] = r < nrows ? (int)(r - bs[u] * W) : 0;
            rows[u] = r < nrows ? (long long)starts[bs[u]] + ts[u] : 0;
        }
        float2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u] = make_float2(0.0f, 0.0f);
            if (bs[u] < 0) continue;
            if (rows[u] < 0 || rows[u] >= table_rows) __trap();  // a window outside the table: the reference raises IndexError
            if (is_f) v[u] = load_pair(feats + (r0 + u * warps) * Ca + c);
            else if (is_k) v[u] = *reinterpret_cast<const float2 *>(kin + rows[u] * Cb + (c - Ca));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (bs[u] < 0) continue;
            if (is_k && mean) { v[u].x = (v[u].x - m0) / s0; v[u].y = (v[u].y - m1) / s1; }
            __nv_bfloat16 *dst = A0 + ((long long)ts[u] * Bpad + bs[u]) * Kp;
            *reinterpret_cast<__nv_bfloat162 *>(dst + c) = __floats2bfloat162_rn(v[u].x, v[u].y);
            if (ts[u] == 0)                                      // h_{-1} = 0 (and any padding behind it)
                for (int k = 64 + c; k < Kp; k += 64) *reinterpret_cast<__nv_bfloat162 *>(dst + k) = __floats2bfloat162_rn(0.0f, 0.0f);
        }
    }
}

// dx [B, W, F] f32 <- dA0 [W, Bpad, Kp] f32 columns [0, F)
template <typename TO>      // dx f32, or bf16 (the rounding the FeatureExtractor's backward applies to its incoming gradient)
__global__ void __launch_bounds__(256)
lstm_unpack_bwf_kernel(const float *__restrict__ dA0, TO *__restrict__ dx, long long B, long long Bpad, int F, int W, int Kp) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long nrows = B * W, warps = (long long)gridDim.x * (blockDim.x >> 5);
    if (F <= 32) {      // the window path (F = FeatureExtractor width): four rows per warp and pass in flight
        constexpr int U = 4;
        for (long long r0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r0 < nrows; r0 += U * warps) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * warps;
                const long long b = r / W;
                const int t = (int)(r - b * W);
                v[u] = (r < nrows && lane < F) ? dA0[((long long)t * Bpad + b) * Kp + lane] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * warps;
                if (r < nrows && lane < F) dx[r * F + lane] = (TO)v[u];
            }
        }
        return;
    }
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const long long b = r / W;
        const int t = (int)(r - b * W);
        const float *src = dA0 + ((long long)t * Bpad + b) * Kp;
        for (int k = lane; k < F; k += 32) dx[r * F + k] = (TO)src[k];
    }
}

__global__ void zero_cols_bf16_kernel(__nv_bfloat16 *__restrict__ A, long long rows, int ld, int col0, int ncols) {
    pdl_wait();
    const long long total = rows * ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        A[(e / ncols) * ld + col0 + (int)(e % ncols)] = __float2bfloat16_rn(0.0f);
}

// One thread per (b, j): gates -> (c_t, h_t); the activated gates replace the pre-activations in G.
__global__ void __launch_bounds__(256)
lstm_cell_fwd_kernel(float *__restrict__ G, const float *__restrict__ c_prev, float *__restrict__ c_out,
                     __nv_bfloat16 *__restrict__ h_next, int ld_next, __nv_bfloat16 *__restrict__ x_up, int ld_up,
                     float *__restrict__ h_out, long long B, int H, float drop_p, const uint32_t *__restrict__ seed_dev,
                     unsigned long long drop_base) {
    pdl_wait();
    const uint32_t seed = seed_dev ? *seed_dev : 0u;
    const long long total = B * (long long)H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int j = (int)(e - b * H);
        float *g = G + b * 4 * H;
        const float i_ = sigmoidf_(g[j]), f_ = sigmoidf_(g[H + j]), g_ = tanhf_(g[2 * H + j]), o_ = sigmoidf_(g[3 * H + j]);
        const float c = f_ * (c_prev ? c_prev[e] : 0.0f) + i_ * g_;
        const float h = o_ * tanhf_(c);
        g[j] = i_; g[H + j] = f_; g[2 * H + j] = g_; g[3 * H + j] = o_;
        c_out[e] = c;
        if (h_next) h_next[b * ld_next + j] = __float2bfloat16_rn(h);
        if (x_up) {
            float hv = h;
            if (drop_p > 0.0f) hv = dropout_keep(seed, drop_base + (unsigned long long)e, drop_p) ? h / (1.0f - drop_p) : 0.0f;
            x_up[b * ld_up + j] = __float2bfloat16_rn(hv);
        }
        if (h_out) h_out[e] = h;
    }
}

// Backward of one cell step.  dh = dh_up (gradient arriving from the layer above through its dropout, or
// the head's gradient) + dh_rec (from step t+1 of this layer); dc accumulates in place.
__global__ void __launch_bounds__(256)
lstm_cell_bwd_kernel(const float *__restrict__ Gact, const float *__restrict__ c, const float *__restrict__ c_prev,
                     const float *__restrict__ dh_up, int ld_up, const float *__restrict__ dh_rec, int ld_rec,
                     float *__restrict__ dc, int dc_init, __nv_bfloat16 *__restrict__ dG, long long B, int H,
                     float drop_p, const uint32_t *__restrict__ seed_dev, unsigned long long drop_base) {
    pdl_wait();
    const uint32_t seed = seed_dev ? *seed_dev : 0u;
    const long long total = B * (long long)H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int j = (int)(e - b * H);
        const float *g = Gact + b * 4 * H;
        const float i_ = g[j], f_ = g[H + j], g_ = g[2 * H + j], o_ = g[3 * H + j];
        float dh = 0.0f;
        if (dh_up) {
            float v = dh_up[b * ld_up + j];
            if (drop_p > 0.0f) v = dropout_keep(seed, drop_base + (unsigned long long)e, drop_p) ? v / (1.0f - drop_p) : 0.0f;
            dh += v;
        }
        if (dh_rec) dh += dh_rec[b * ld_rec + j];
        const float tc = tanhf_(c[e]);
        const float dct = (dc_init ? 0.0f : dc[e]) + dh * o_ * (1.0f - tc * tc);
        const float cp = c_prev ? c_prev[e] : 0.0f;
        __nv_bfloat16 *d = dG + b * 4 * H;
        d[j] = __float2bfloat16_rn(dct * g_ * i_ * (1.0f - i_));
        d[H + j] = __float2bfloat16_rn(dct * cp * f_ * (1.0f - f_));
        d[2 * H + j] = __float2bfloat16_rn(dct * i_ * (1.0f - g_ * g_));
        d[3 * H + j] = __float2bfloat16_rn(dh * tc * o_ * (1.0f - o_));
        dc[e] = dct * f_;
    }
}

// ---------------------------------------------------------------------------------------------------- fp32 parity mode
// The 1e-5 mode of the head (exp_kwargs['precision'] = "fp32"): same recurrence with fp32 operands on the SIMT GEMM
// (gemm_f32.cu) and EXACT-math cells (expf / tanhf / IEEE division; cuDNN's fast-math cell is ~20x noisier than the CPU
// reference, measured in round 1).  Buffers are time-major fp32: X_l [W, B, in_l], Hs_l [W, B, H], C_l [W, B, H],
// G_l [W, B, 4H] (pre-activations, overwritten by the activated gates), dG_l [W, B, 4H].

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

// x ([B, F, W] when layout == 0, [B, W, F] when layout == 1) -> X0 [W, B, F]
__global__ void lstm_pack_f32_kernel(const float *__restrict__ x, float *__restrict__ X0, long long B, int F, int W, int layout) {
    pdl_wait();
    const long long total = B * (long long)F * W;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e % F);
        const long long r = e / F;
        const long long b = r % B;
        const int t = (int)(r / B);
        X0[e] = layout ? x[(b * W + t) * F + k] : x[(b * F + k) * W + t];
    }
}
// dX0 [W, B, ld] columns [0, F) -> dx in the layout of x
__global__ void lstm_unpack_f32_kernel(const float *__restrict__ dX0, float *__restrict__ dx, long long B, int F, int W, int ld,
                                       int layout) {
    pdl_wait();
    const long long total = B * (long long)F * W;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long b; int t, k;
        if (layout) { k = (int)(e % F); const long long r = e / F; t = (int)(r % W); b = r / W; }
        else { t = (int)(e % W); const long long r = e / W; k = (int)(r % F); b = r / F; }
        dx[e] = dX0[((long long)t * B + b) * ld + k];
    }
}

__global__ void __launch_bounds__(256)
lstm_cell_fwd_f32_kernel(float *__restrict__ G, const float *__restrict__ c_prev, float *__restrict__ c_out,
                         float *__restrict__ h_out, float *__restrict__ x_up, long long B, int H, float drop_p,
                         const uint32_t *__restrict__ seed_dev, unsigned long long drop_base) {
    pdl_wait();
    const uint32_t seed = seed_dev ? *seed_dev : 0u;
    const long long total = B * (long long)H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int j = (int)(e - b * H);
        float *g = G + b * 4 * H;
        const float i_ = sigmoid_exact(g[j]), f_ = sigmoid_exact(g[H + j]), g_ = tanhf(g[2 * H + j]), o_ = sigmoid_exact(g[3 * H + j]);
        const float c = f_ * (c_prev ? c_prev[e] : 0.0f) + i_ * g_;
        const float h = o_ * tanhf(c);
        g[j] = i_; g[H + j] = f_; g[2 * H + j] = g_; g[3 * H + j] = o_;
        c_out[e] = c;
        h_out[e] = h;
        if (x_up) x_up[e] = (drop_p > 0.0f) ? (dropout_keep(seed, drop_base + (unsigned long long)e, drop_p) ? h / (1.0f - drop_p) : 0.0f) : h;
    }
}

__global__ void __launch_bounds__(256)
lstm_cell_bwd_f32_kernel(const float *__restrict__ Gact, const float *__restrict__ c, const float *__restrict__ c_prev,
                         const float *__restrict__ dh_up, int ld_up, const float *__restrict__ dh_rec,
                         float *__restrict__ dc, int dc_init, float *__restrict__ dG, long long B, int H, float drop_p,
                         const uint32_t *__restrict__ seed_dev, unsigned long long drop_base) {
    pdl_wait();
    const uint32_t seed = seed_dev ? *seed_dev : 0u;
    const long long total = B * (long long)H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int j = (int)(e - b * H);
        const float *g = Gact + b * 4 * H;
        const float i_ = g[j], f_ = g[H + j], g_ = g[2 * H + j], o_ = g[3 * H + j];
        float dh = 0.0f;
        if (dh_up) {
            float v = dh_up[b * ld_up + j];
            if (drop_p > 0.0f) v = dropout_keep(seed, drop_base + (unsigned long long)e, drop_p) ? v / (1.0f - drop_p) : 0.0f;
            dh += v;
        }
        if (dh_rec) dh += dh_rec[e];
        const float tc = tanhf(c[e]);
        const float dct = (dc_init ? 0.0f : dc[e]) + dh * o_ * (1.0f - tc * tc);
        const float cp = c_prev ? c_prev[e] : 0.0f;
        float *d = dG + b * 4 * H;
        d[j] = dct * g_ * i_ * (1.0f - i_);
        d[H + j] = dct * cp * f_ * (1.0f - f_);
        d[2 * H + j] = dct * i_ * (1.0f - g_ * g_);
        d[3 * H + j] = dh * tc * o_ * (1.0f - o_);
        dc[e] = dct * f_;
    }
}

static unsigned grid_for(long long total) {
    const long long want = (total + 255) / 256, cap = (long long)num_sms() * 8;
    return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int b200med_lstm_pack_inputs(const float *x, void *A0, int64_t B, int64_t Bpad,
                                                                               int32_t F, int32_t W, int32_t H, int32_t Kp,
                                                                               int32_t hoff, int32_t x_layout, void *stream) {
    B200MED_REQUIRE(B >= 1 && Bpad >= B && F >= 1 && W >= 1 && H >= 1 && hoff >= F && Kp >= hoff + H, "bad shape");
    B200MED_REQUIRE(x_layout == 0 || x_layout == 1, "x_layout: 0 = [B,F,W], 1 = [B,W,F]");
    if (x_layout == 1) {
        B200MED_REQUIRE(x && A0 && Kp % 8 == 0, "bad argument");
        const long long blocks = (B * (long long)W + 7) / 8, cap1 = (long long)num_sms() * 8;
        launch_k(lstm_pack_bwf_kernel, (unsigned)(blocks < cap1 ? blocks : cap1), 256, 0, (cudaStream_t)stream, 
            x, reinterpret_cast<__nv_bfloat16 *>(A0), B, Bpad, F, W, H, Kp, hoff);
        return after_launch("lstm_pack_bwf_kernel");
    }
    B200MED_REQUIRE(x && A0, "null pointer");
    B200MED_REQUIRE(Kp % 8 == 0 && (size_t)kPackWin * F * (W + 1) * 4 <= 48 * 1024, "bad shape");
    const long long blocks = (B + kPackWin - 1) / kPackWin, cap = (long long)num_sms() * 8;
    launch_k(lstm_pack_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, (size_t)kPackWin * F * (W + 1) * 4, (cudaStream_t)stream, 
        x, reinterpret_cast<__nv_bfloat16 *>(A0), B, Bpad, F, W, H, Kp, hoff);
    return after_launch("lstm_pack_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_pack_parts(
    const float *feats, int32_t Ca, const float *kin_table, int64_t table_rows, int32_t Cb, const float *mean, const float *stdv,
    int32_t stat_rows, const int32_t *starts, void *A0, int64_t B, int64_t Bpad, int32_t W, int32_t H, int32_t Kp, int32_t hoff,
    void *stream) {
    B200MED_REQUIRE(B >= 1 && Bpad >= B && Ca >= 1 && Cb >= 0 && W >= 1 && H >= 1, "bad shape");
    B200MED_REQUIRE(Kp % 2 == 0 && hoff % 2 == 0 && hoff >= Ca + Cb && hoff + H <= Kp, "bad operand geometry");
    B200MED_REQUIRE(feats && starts && A0 && (Cb == 0 || kin_table), "null pointer");
    B200MED_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must both be given or both NULL");
    B200MED_REQUIRE(stat_rows == 1 || stat_rows == W, "statistics: one row, or one per window step");
    const long long blocks = (B * (long long)W + 7) / 8, cap = (long long)num_sms() * 8;
    if (Ca % 2 == 0 && Cb % 2 == 0 && Ca + Cb <= 64 && hoff == 64 && hoff + H == Kp && stat_rows == 1 &&
        ((uintptr_t)feats % 8 == 0) && ((uintptr_t)kin_table % 8 == 0)) {
        launch_k(lstm_pack_parts64_kernel<float>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, 
            feats, Ca, kin_table, table_rows, Cb, mean, stdv, starts, reinterpret_cast<__nv_bfloat16 *>(A0), B, Bpad, W, H, Kp);
        return after_launch("lstm_pack_parts64_kernel");
    }
    launch_k(lstm_pack_parts_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, 
        feats, Ca, kin_table, table_rows, Cb, mean, stdv, stat_rows, starts, reinterpret_cast<__nv_bfloat16 *>(A0), B, Bpad, W, H, Kp, hoff);
    return after_launch("lstm_pack_parts_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_pack_parts_bf16(
    const void *feats, int32_t Ca, const float *kin_table, int64_t table_rows, int32_t Cb, const float *mean, const float *stdv,
    const int32_t *starts, void *A0, int64_t B, int64_t Bpad, int32_t W, int32_t H, int32_t Kp, void *stream) {
    B200MED_REQUIRE(B >= 1 && Bpad >= B && Ca >= 2 && Cb >= 0 && W >= 1 && H >= 1, "bad shape");
    B200MED_REQUIRE(Ca % 2 == 0 && Cb % 2 == 0 && Ca + Cb <= 64 && Kp == 64 + H, "geometry: even widths, Ca + Cb <= 64, Kp = 64 + H");
    B200MED_REQUIRE(feats && starts && A0 && (Cb == 0 || kin_table), "null pointer");
    B200MED_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must both be given or both NULL");
    B200MED_REQUIRE(((uintptr_t)feats % 4 == 0) && ((uintptr_t)kin_table % 8 == 0), "misaligned feats / table");
    const long long blocks = (B * (long long)W + 7) / 8, cap = (long long)num_sms() * 8;
    launch_k(lstm_pack_parts64_kernel<__nv_bfloat16>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream,
             reinterpret_cast<const __nv_bfloat16 *>(feats), Ca, kin_table, table_rows, Cb, mean, stdv, starts,
             reinterpret_cast<__nv_bfloat16 *>(A0), B, Bpad, W, H, Kp);
    return after_launch("lstm_pack_parts64_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_unpack_dx_bf16(const float *dA0, void *dx, int64_t B, int64_t Bpad,
                                                                                  int32_t F, int32_t W, int32_t Kp, void *stream) {
    B200MED_REQUIRE(B >= 1 && Bpad >= B && F >= 1 && W >= 1 && Kp >= F, "bad shape");
    B200MED_REQUIRE(dA0 && dx, "null pointer");
    const long long blocks = (B * (long long)W + 7) / 8, cap1 = (long long)num_sms() * 8;
    launch_k(lstm_unpack_bwf_kernel<__nv_bfloat16>, (unsigned)(blocks < cap1 ? blocks : cap1), 256, 0, (cudaStream_t)stream, dA0,
             reinterpret_cast<__nv_bfloat16 *>(dx), B, Bpad, F, W, Kp);
    return after_launch("lstm_unpack_bwf_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_unpack_dx(const float *dA0, float *dx, int64_t B, int64_t Bpad,
                                                                             int32_t F, int32_t W, int32_t Kp, int32_t x_layout,
                                                                             void *stream) {
    B200MED_REQUIRE(B >= 1 && Bpad >= B && F >= 1 && W >= 1 && Kp >= F, "bad shape");
    B200MED_REQUIRE(x_layout == 0 || x_layout == 1, "x_layout: 0 = [B,F,W], 1 = [B,W,F]");
    if (x_layout == 1) {
        B200MED_REQUIRE(dA0 && dx, "null pointer");
        const long long blocks = (B * (long long)W + 7) / 8, cap1 = (long long)num_sms() * 8;
        launch_k(lstm_unpack_bwf_kernel<float>, (unsigned)(blocks < cap1 ? blocks : cap1), 256, 0, (cudaStream_t)stream, dA0, dx, B, Bpad, F, W, Kp);
        return after_launch("lstm_unpack_bwf_kernel");
    }
    B200MED_REQUIRE(dA0 && dx, "null pointer");
    B200MED_REQUIRE((size_t)kPackWin * F * (W + 1) * 4 <= 48 * 1024, "bad shape");
    const long long blocks = (B + kPackWin - 1) / kPackWin, cap = (long long)num_sms() * 8;
    launch_k(lstm_unpack_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, (size_t)kPackWin * F * (W + 1) * 4, (cudaStream_t)stream, 
        dA0, dx, B, Bpad, F, W, Kp);
    return after_launch("lstm_unpack_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_zero_cols_bf16(void *A, int64_t rows, int32_t ld, int32_t col0,
                                                                             int32_t ncols, void *stream) {
    B200MED_REQUIRE(rows >= 0 && ncols >= 0 && col0 >= 0 && col0 + ncols <= ld, "bad shape");
    if (rows == 0 || ncols == 0) return B200MED_OK;
    B200MED_REQUIRE(A, "null pointer");
    launch_k(zero_cols_bf16_kernel, grid_for(rows * ncols), 256, 0, (cudaStream_t)stream, reinterpret_cast<__nv_bfloat16 *>(A), rows, ld,
                                                                                    col0, ncols);
    return after_launch("zero_cols_bf16_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_cell_fwd(float *G, const float *c_prev, float *c_out,
                                                                            void *h_next, int32_t ld_next, void *x_up,
                                                                            int32_t ld_up, float *h_out, int64_t B, int32_t H,
                                                                            float drop_p, const uint32_t *seed, uint64_t drop_base,
                                                                            void *stream) {
    B200MED_REQUIRE(B >= 1 && H >= 1 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(G && c_out, "null pointer");
    launch_k(lstm_cell_fwd_kernel, grid_for(B * (long long)H), 256, 0, (cudaStream_t)stream, 
        G, c_prev, c_out, reinterpret_cast<__nv_bfloat16 *>(h_next), ld_next, reinterpret_cast<__nv_bfloat16 *>(x_up), ld_up,
        h_out, B, H, drop_p, seed, drop_base);
    return after_launch("lstm_cell_fwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_cell_bwd(const float *Gact, const float *c, const float *c_prev,
                                                                            const float *dh_up, int32_t ld_up, const float *dh_rec,
                                                                            int32_t ld_rec, float *dc, int32_t dc_init, void *dG,
                                                                            int64_t B, int32_t H, float drop_p, const uint32_t *seed,
                                                                            uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(B >= 1 && H >= 1 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(Gact && c && dc && dG, "null pointer");
    launch_k(lstm_cell_bwd_kernel, grid_for(B * (long long)H), 256, 0, (cudaStream_t)stream, 
        Gact, c, c_prev, dh_up, ld_up, dh_rec, ld_rec, dc, dc_init, reinterpret_cast<__nv_bfloat16 *>(dG), B, H, drop_p, seed,
        drop_base);
    return after_launch("lstm_cell_bwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_pack_f32(const float *x, float *X0, int64_t B, int32_t F, int32_t W,
                                                                           int32_t x_layout, void *stream) {
    B200MED_REQUIRE(B >= 1 && F >= 1 && W >= 1 && (x_layout == 0 || x_layout == 1), "bad shape");
    B200MED_REQUIRE(x && X0, "null pointer");
    launch_k(lstm_pack_f32_kernel, grid_for(B * (long long)F * W), 256, 0, (cudaStream_t)stream, x, X0, B, F, W, x_layout);
    return after_launch("lstm_pack_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_unpack_f32(const float *dX0, float *dx, int64_t B, int32_t F, int32_t W,
                                                                             int32_t ld, int32_t x_layout, void *stream) {
    B200MED_REQUIRE(B >= 1 && F >= 1 && W >= 1 && ld >= F && (x_layout == 0 || x_layout == 1), "bad shape");
    B200MED_REQUIRE(dX0 && dx, "null pointer");
    launch_k(lstm_unpack_f32_kernel, grid_for(B * (long long)F * W), 256, 0, (cudaStream_t)stream, dX0, dx, B, F, W, ld, x_layout);
    return after_launch("lstm_unpack_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_cell_fwd_f32(float *G, const float *c_prev, float *c_out, float *h_out,
                                                                               float *x_up, int64_t B, int32_t H, float drop_p,
                                                                               const uint32_t *seed, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(B >= 1 && H >= 1 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(G && c_out && h_out, "null pointer");
    launch_k(lstm_cell_fwd_f32_kernel, grid_for(B * (long long)H), 256, 0, (cudaStream_t)stream, G, c_prev, c_out, h_out, x_up, B, H, drop_p,
                                                                                           seed, drop_base);
    return after_launch("lstm_cell_fwd_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_cell_bwd_f32(const float *Gact, const float *c, const float *c_prev,
                                                                               const float *dh_up, int32_t ld_up, const float *dh_rec,
                                                                               float *dc, int32_t dc_init, float *dG, int64_t B, int32_t H,
                                                                               float drop_p, const uint32_t *seed, uint64_t drop_base,
                                                                               void *stream) {
    B200MED_REQUIRE(B >= 1 && H >= 1 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(Gact && c && dc && dG, "null pointer");
    launch_k(lstm_cell_bwd_f32_kernel, grid_for(B * (long long)H), 256, 0, (cudaStream_t)stream, Gact, c, c_prev, dh_up, ld_up, dh_rec, dc,
                                                                                           dc_init, dG, B, H, drop_p, seed, drop_base);
    return after_launch("lstm_cell_bwd_f32_kernel");
}
