// K2 (parity mode): fp32 SIMT GEMMs for the FeatureExtractor MLP (MED/modeling/models.py:6-47).
//
// TF32 / bf16 tensor-core products cannot meet the 1e-5 relative bar of the fp32 parity mode
// ("hard parts"), so this path uses true fp32 FMAs on the CUDA cores with a fixed
// reduction order (deterministic).  The throughput mode is the tcgen05 kernel in gemm_tcgen05.cu.
//
// One tiled kernel covers the three products of a Linear layer:
//   NT  y [M,N] = x [M,K]  W[N,K]^T            (forward,           A row = m, B row = n, both K-contiguous)
//   NN  dx[M,K] = dy[M,N]  W[N,K]              (data gradient,     reduction over n)
//   TN  dW[N,K] = dy[M,N]^T x[M,K]             (weight gradient,   reduction over m, split in slabs)
// Tile 64x64x16, 256 threads, 4x4 outputs per thread, operands staged in shared memory as [k][row].
#include "common.cuh"

namespace b200med {

constexpr int BM = 64, BN = 64, BK = 16;   // the large tile (TM = 4); TM = 2 gives 32 x 32 tiles
constexpr int kGemmRelu = 1, kGemmAccum = 2, kGemmReluA = 4, kGemmReluB = 8, kGemmSplit = 16;   // `flags` bits (b200med.h)

// C[i,j] = sum_r A(i,r) * B(j,r)   with A(i,r) = A[i*a_rs + r*a_cs], B(j,r) = B[j*b_rs + r*b_cs].
// Epilogue: + bias[j], (+ the old C[i,j]: kGemmAccum), ReLU, * (mask[i,j] > 0).  Operand transforms on load: kGemmReluA /
// kGemmReluB clamp the A / B elements at zero (a ReLU that precedes the product is never materialised).
// blockIdx.z = reduction slab (split-R), partial results go to C + z*slab_stride when gridDim.z > 1.
// TM = micro-tile edge per thread (16 x 16 threads): TM = 4 -> 64 x 64 CTA tiles, TM = 2 -> 32 x 32 tiles for problems
// whose 64 x 64 grid would leave most SMs idle (the frame path's M = T ~ 600 rows).  Every output is the same ascending-r
// FMA chain in both, so the tile choice does not change a single bit of the result.
template <int TM>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
                long long I, long long J, long long R, long long a_rs, long long a_cs, long long b_rs,
                long long b_cs, long long ldc, const float *__restrict__ bias, int relu,
                const float *__restrict__ mask, long long ld_mask, long long r_per_slab,
                long long slab_stride) {
    pdl_wait();
    constexpr int TB = 16 * TM;   // tile edge
    __shared__ __align__(16) float As[BK][TB + 4];
    __shared__ __align__(16) float Bs[BK][TB + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a TM x TM micro-tile
    const long long i0 = (long long)blockIdx.y * TB, j0 = (long long)blockIdx.x * TB;
    const long long r_begin = (long long)blockIdx.z * r_per_slab;
    const long long r_end = min(R, r_begin + r_per_slab);

    float acc[TM][TM] = {};
    // loader mapping: when the reduction index is the contiguous one (cs == 1) let consecutive
    // threads walk r; otherwise let them walk the row index.
    const bool a_r_fast = (a_cs == 1), b_r_fast = (b_cs == 1);
    // Register-staged software pipeline: the global loads of tile i+1 are in flight while tile i is multiplied (a GEMM with
    // few CTAs is a chain of load latencies otherwise).  Summation order unchanged.
    constexpr int LD = (TB * BK) / 256;
    float ra[LD], rb[LD];
    auto fetch = [&](long long r0) {
#pragma unroll
        for (int l = 0; l < LD; ++l) {
            const int e = l * 256 + tid;
            const int rr = a_r_fast ? (e % BK) : (e / TB);
            const int ii = a_r_fast ? (e / BK) : (e % TB);
            const long long gi = i0 + ii, gr = r0 + rr;
            ra[l] = (gi < I && gr < r_end) ? __ldg(A + gi * a_rs + gr * a_cs) : 0.0f;
            if (relu & kGemmReluA) ra[l] = fmaxf(ra[l], 0.0f);
        }
#pragma unroll
        for (int l = 0; l < LD; ++l) {
            const int e = l * 256 + tid;
            const int rr = b_r_fast ? (e % BK) : (e / TB);
            const int jj = b_r_fast ? (e / BK) : (e % TB);
            const long long gj = j0 + jj, gr = r0 + rr;
            rb[l] = (gj < J && gr < r_end) ? __ldg(B + gj * b_rs + gr * b_cs) : 0.0f;
            if (relu & kGemmReluB) rb[l] = fmaxf(rb[l], 0.0f);
        }
    };
    if (r_begin < r_end) fetch(r_begin);
    for (long long r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
        for (int l = 0; l < LD; ++l) {
            const int e = l * 256 + tid;
            As[a_r_fast ? (e % BK) : (e / TB)][a_r_fast ? (e / BK) : (e % TB)] = ra[l];
            Bs[b_r_fast ? (e % BK) : (e / TB)][b_r_fast ? (e / BK) : (e % TB)] = rb[l];
        }
        __syncthreads();
        if (r0 + BK < r_end) fetch(r0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TM];
            if constexpr (TM == 4) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
                av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
                bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w;
            } else {
                const float2 a = *reinterpret_cast<const float2 *>(&As[k][ty * 2]);
                const float2 b = *reinterpret_cast<const float2 *>(&Bs[k][tx * 2]);
                av[0] = a.x; av[1] = a.y;
                bv[0] = b.x; bv[1] = b.y;
            }
#pragma unroll
            for (int u = 0; u < TM; ++u)
#pragma unroll
                for (int v = 0; v < TM; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        __syncthreads();
    }
    float *Cz = C + (long long)blockIdx.z * slab_stride;
#pragma unroll
    for (int u = 0; u < TM; ++u) {
        const long long gi = i0 + ty * TM + u;
        if (gi >= I) continue;
#pragma unroll
        for (int v = 0; v < TM; ++v) {
            const long long gj = j0 + tx * TM + v;
            if (gj >= J) continue;
            float y = acc[u][v];
            if (bias) y += bias[gj];
            if (relu & kGemmAccum) y += Cz[gi * ldc + gj];
            if (relu & kGemmRelu) y = fmaxf(y, 0.0f);
            if (mask && !(mask[gi * ld_mask + gj] > 0.0f)) y = 0.0f;
            Cz[gi * ldc + gj] = y;
        }
    }
}

// out[e] (+)= sum_z part[z*stride + e], z ascending (fixed order).
__global__ void slab_reduce_kernel(const float *__restrict__ part, float *__restrict__ out, long long n, int slabs,
                                   long long stride, int accumulate) {
    pdl_wait();
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int z = 0; z < slabs; ++z) s += part[z * stride + e];
        out[e] = accumulate ? out[e] + s : s;
    }
}

// Column sums of a [M, N] matrix (bias gradient).  Stage 1: each block sums a slab of rows for a tile
// of 32 columns; stage 2: slabs are added in ascending order.  T = float or bf16.
template <typename T>
__global__ void colsum_partial_kernel(const T *__restrict__ x, float *__restrict__ part, long long M, int N,
                                      long long ld, long long rows_per_slab) {
    pdl_wait();
    __shared__ float sh[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int lane_row = threadIdx.x >> 5;  // 8 row lanes
    const long long m0 = (long long)blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
    float s = 0.0f;
    if (col < N)
        for (long long m = m0 + lane_row; m < m1; m += 8) s += (float)x[m * ld + col];
    sh[lane_row][threadIdx.x & 31] = s;
    __syncthreads();
    if (lane_row == 0 && col < N) {
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += sh[r][threadIdx.x & 31];
        part[(long long)blockIdx.y * N + col] = t;
    }
}

// Vectorised variant: 16-byte loads (V = 8 bf16 / 4 f32 columns per thread), a warp covers 32*V columns of one row,
// the 8 warps of a block walk 8 rows at a time.  Needs N % V == 0, ld % V == 0 and a 16-byte aligned base.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_vec_kernel(const T *__restrict__ x, float *__restrict__ part, long long M, int N, long long ld,
                          long long rows_per_slab) {
    pdl_wait();
    constexpr int V = 16 / sizeof(T);
    __shared__ float sh[8][32 * V + 1];
    const int lane = threadIdx.x & 31, lane_row = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + lane) * V;
    const long long m0 = (long long)blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
    float s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = 0.0f;
    if (col < N) {
#pragma unroll 4
        for (long long m = m0 + lane_row; m < m1; m += 8) {
            const uint4 u = ldg_stream_u4(reinterpret_cast<const uint4 *>(x + m * ld + col));
            if constexpr (sizeof(T) == 2) {
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[2 * j] += __uint_as_float(w[j] << 16);
                    s[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                }
            } else {
                s[0] += __uint_as_float(u.x); s[1] += __uint_as_float(u.y); s[2] += __uint_as_float(u.z); s[3] += __uint_as_float(u.w);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) sh[lane_row][lane * V + j] = s[j];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * V; c += 256) {
        const int gc = blockIdx.x * 32 * V + c;
        if (gc < N) {
            float t = 0.0f;
#pragma unroll
            for (int r = 0; r < 8; ++r) t += sh[r][c];
            part[(long long)blockIdx.y * N + gc] = t;
        }
    }
}

// Flat variant for N / V a power of two <= 256: the 256 threads of a block tile (256 / lpr) rows x lpr column groups per pass
// (lpr = N / V), so that narrow matrices (N = 32) keep every lane busy; 4 passes of 16-byte loads in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_flat_kernel(const T *__restrict__ x, float *__restrict__ part, long long M, int N, long long ld,
                           long long rows_per_slab) {
    pdl_wait();
    constexpr int V = 16 / sizeof(T);
    __shared__ float sh[256][V + 1];
    const int lpr = N / V;                      // column groups (threads) per row
    const int rpp = 256 / lpr;                  // rows per pass
    const int cg = threadIdx.x % lpr, r0 = threadIdx.x / lpr;
    const long long m0 = (long long)blockIdx.x * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
    float s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = 0.0f;
#pragma unroll 4
    for (long long m = m0 + r0; m < m1; m += rpp) {
        const uint4 u = ldg_stream_u4(reinterpret_cast<const uint4 *>(x + m * ld + cg * V));
        if constexpr (sizeof(T) == 2) {
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[2 * j] += __uint_as_float(w[j] << 16);
                s[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
            }
        } else {
            s[0] += __uint_as_float(u.x); s[1] += __uint_as_float(u.y); s[2] += __uint_as_float(u.z); s[3] += __uint_as_float(u.w);
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) sh[threadIdx.x][j] = s[j];
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += 256) {        // fixed order over the block's row lanes: deterministic
        const int g = c / V, j = c - g * V;
        float t = 0.0f;
        for (int r = 0; r < rpp; ++r) t += sh[r * lpr + g][j];
        part[(long long)blockIdx.x * N + c] = t;
    }
}

// out[e] = sum_z part[z*n + e] in a fixed order: 8 strided partial sums per column (z = r, r+8, ...) added r = 0..7.
__global__ void __launch_bounds__(256)
slab_reduce8_kernel(const float *__restrict__ part, float *__restrict__ out, int n, int slabs) {
    pdl_wait();
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    float s = 0.0f;
    if (c < n)
        for (int z = r; z < slabs; z += 8) s += part[(long long)z * n + c];
    sh[r][threadIdx.x & 31] = s;
    __syncthreads();
    if (r == 0 && c < n) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x & 31];
        out[c] = t;
    }
}

__global__ void cast_f32_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, long long n) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void relu_cast_f32_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, long long n) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(fmaxf(x[i], 0.0f));
}
// x = hi + lo (+ 2^-17 |x|) with hi = bf16(x), lo = bf16(x - hi): three bf16 products hi*hi' + lo*hi' + hi*lo' carry an fp32-like
// product on the bf16 tensor cores.  The three terms are folded into ONE GEMM by concatenating along its reduction dimension:
// a LEFT operand is laid out (hi, lo, hi), a RIGHT operand (hi, hi, lo).  row3 [R, 3C]: the blocks side by side in every row
// (reduction over columns); stack3 [3R, C]: the blocks one under the other (reduction over rows).  order 0 = left, 1 = right.
__global__ void split_bf16x3_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ row3, __nv_bfloat16 *__restrict__ stack3,
                                    long long R, int C, int row_order, int stack_order, int relu) {
    pdl_wait();
    const long long n = R * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = x[i];
        if (relu) v = fmaxf(v, 0.0f);
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const long long r = i / C;
        const int c = (int)(i - r * C);
        if (row3) {
            __nv_bfloat16 *d = row3 + r * 3 * C + c;
            d[0] = hi; d[C] = row_order ? hi : lo; d[2 * C] = row_order ? lo : hi;
        }
        if (stack3) {
            stack3[i] = hi; stack3[n + i] = stack_order ? hi : lo; stack3[2 * n + i] = stack_order ? lo : hi;
        }
    }
}
// Three-way split x = h + m + l (h = bf16(x), m = bf16(x - h), l = bf16(x - h - m): exact, 8 + 8 + 8 mantissa bits) laid out for
// ONE bf16 tensor-core GEMM over a 6x longer reduction that reproduces the fp32 product: blocks of a left operand
// (m, l, h, m, h, h), of a right operand (m, h, l, h, m, h) -> products mm, lh, hl, mh, hm, hh (dropped: ml, lm, ll <= 2^-24).
// Small products FIRST: the tensor core adds into its fp32 accumulator with truncation, so the order matters -- with the hh
// block first every later (small) addition truncates against a full-size accumulator and the result is biased towards zero by
// 2e-5 at K = 2048; with hh last the error is 1.4e-6 (fp32 FMA chain: 6e-7; scripts/split6_accuracy.py).
template <int V>
__global__ void split_bf16x6_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ row6, __nv_bfloat16 *__restrict__ stack6,
                                    long long R, int C, int role, int relu) {
    pdl_wait();
    const long long n = R * C;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < n; i += (long long)gridDim.x * blockDim.x * V) {
        float v[V];
        if (V == 4) {
            const float4 q = *reinterpret_cast<const float4 *>(x + i);
            v[0] = q.x; v[1 % V] = q.y; v[2 % V] = q.z; v[3 % V] = q.w;
        } else {
            v[0] = x[i];
        }
        __nv_bfloat16 h[V], m[V], l[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float a = relu ? fmaxf(v[k], 0.0f) : v[k];
            h[k] = __float2bfloat16_rn(a);
            const float r1 = a - __bfloat162float(h[k]);
            m[k] = __float2bfloat16_rn(r1);
            l[k] = __float2bfloat16_rn(r1 - __bfloat162float(m[k]));
        }
        const __nv_bfloat16 *blk[6];
        if (role == 0) { blk[0] = m; blk[1] = l; blk[2] = h; blk[3] = m; blk[4] = h; blk[5] = h; }
        else           { blk[0] = m; blk[1] = h; blk[2] = l; blk[3] = h; blk[4] = m; blk[5] = h; }
        const long long r = i / C;
        const int c = (int)(i - r * C);
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            if (V == 4) {
                const uint2 u = make_uint2((uint32_t)__bfloat16_as_ushort(blk[b][0]) | ((uint32_t)__bfloat16_as_ushort(blk[b][1 % V]) << 16),
                                           (uint32_t)__bfloat16_as_ushort(blk[b][2 % V]) | ((uint32_t)__bfloat16_as_ushort(blk[b][3 % V]) << 16));
                if (row6) *reinterpret_cast<uint2 *>(row6 + r * 6 * C + (long long)b * C + c) = u;
                if (stack6) *reinterpret_cast<uint2 *>(stack6 + b * n + i) = u;
            } else {
                if (row6) row6[r * 6 * C + (long long)b * C + c] = blk[b][0];
                if (stack6) stack6[b * n + i] = blk[b][0];
            }
        }
    }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y, long long n) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __bfloat162float(x[i]);
}

static long long weight_slabs(long long M, long long N, long long K) {
    // Reduction slabs of the weight gradient dW[N,K] = dy^T x: enough CTAs (output tiles x slabs) for about two waves of
    // the 148 SMs, never fewer than one slab per 2048 rows, never fewer than 32 rows per slab.  (One slab for M < 2048 --
    // the first rule alone -- left the frame path's [64 x 64] gradients to ONE CTA walking all T rows: 84 us per launch.)
    const long long tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
    long long s = (296 + tiles - 1) / tiles;
    const long long by_rows = (M + 2047) / 2048, max_by_rows = (M + 31) / 32;
    if (s < by_rows) s = by_rows;
    if (s > max_by_rows) s = max_by_rows;
    if (s < 1) s = 1;
    if (s > 128) s = 128;
    return s;
}
static long long colsum_slabs_flat(long long M) {
    long long s = (M + 63) / 64;         // >= 64 rows per slab
    if (s < 1) s = 1;
    if (s > 592) s = 592;                // 4 CTAs on each of the 148 SMs
    return s;
}
static long long colsum_slabs(long long M) {
    long long s = (M + 63) / 64;     // >= 64 rows per slab
    if (s < 1) s = 1;
    if (s > 128) s = 128;
    return s;
}

}  // namespace b200med

using namespace b200med;

static int launch_gemm(const float *A, const float *B, float *C, long long I, long long J, long long R,
                       long long a_rs, long long a_cs, long long b_rs, long long b_cs, long long ldc,
                       const float *bias, int relu, const float *mask, long long ld_mask, int slabs,
                       long long slab_stride, cudaStream_t st) {
    const long long per = ((R + slabs - 1) / slabs + BK - 1) / BK * BK;
    const long long ctas64 = ((J + BN - 1) / BN) * ((I + BM - 1) / BM) * slabs;
    if (ctas64 < num_sms()) {   // too few 64 x 64 tiles to fill the GPU: 32 x 32 tiles (4x the CTAs, identical results)
        dim3 grid((unsigned)((J + 31) / 32), (unsigned)((I + 31) / 32), (unsigned)slabs);
        launch_k(gemm_f32_kernel<2>, grid, 256, 0, st, A, B, C, I, J, R, a_rs, a_cs, b_rs, b_cs, ldc, bias, relu, mask, ld_mask,
                                                  per, slab_stride);
    } else {
        dim3 grid((unsigned)((J + BN - 1) / BN), (unsigned)((I + BM - 1) / BM), (unsigned)slabs);
        launch_k(gemm_f32_kernel<4>, grid, 256, 0, st, A, B, C, I, J, R, a_rs, a_cs, b_rs, b_cs, ldc, bias, relu, mask, ld_mask,
                                                  per, slab_stride);
    }
    return after_launch("gemm_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_linear_fwd_f32(const float *x, const float *w, const float *bias, float *y, int64_t M,
                                      int32_t N, int32_t K, int32_t relu, void *stream) {
    B200MED_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad shape");
    if (M == 0) return B200MED_OK;
    B200MED_REQUIRE(x && w && y, "null pointer");
    return launch_gemm(x, w, y, M, N, K, K, 1, K, 1, N, bias, relu, nullptr, 0, 1, 0, (cudaStream_t)stream);
}

static long long split_slabs(long long I, long long J, long long R) { return weight_slabs(R, I, J); }

extern "C" __attribute__((visibility("default"))) int64_t b200med_gemm_f32_ws_bytes(int64_t I, int64_t J, int64_t R) {
    return split_slabs(I, J, R) * I * J * 4 + 256;
}

// General strided fp32 product (see b200med.h): C[i,j] = epilogue(sum_r A[i*a_rs + r*a_cs] * B[j*b_rs + r*b_cs]).
extern "C" __attribute__((visibility("default"))) int b200med_gemm_f32(const float *A, const float *B, float *C, int64_t I, int64_t J, int64_t R,
                                int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t ldc, const float *bias,
                                const float *mask, int64_t ld_mask, int32_t flags, void *workspace, void *stream) {
    B200MED_REQUIRE(I >= 0 && J >= 1 && R >= 1 && ldc >= J, "bad shape");
    if (I == 0) return B200MED_OK;
    B200MED_REQUIRE(A && B && C, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & kGemmSplit))
        return launch_gemm(A, B, C, I, J, R, a_rs, a_cs, b_rs, b_cs, ldc, bias, flags & 15, mask, ld_mask, 1, 0, st);
    // long reduction, small output (weight gradients): deterministic slabs, summed in ascending order
    B200MED_REQUIRE(workspace && !bias && !mask && ldc == J && !(flags & kGemmRelu), "split-R products take no epilogue and a dense C");
    const int slabs = (int)split_slabs(I, J, R);
    float *part = reinterpret_cast<float *>(workspace);
    if (int e = launch_gemm(A, B, part, I, J, R, a_rs, a_cs, b_rs, b_cs, J, nullptr, flags & (kGemmReluA | kGemmReluB), nullptr, 0,
                            slabs, I * J, st)) return e;
    const long long n = I * J, blocks = (n + 255) / 256;
    launch_k(slab_reduce_kernel, (unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st, part, C, n, slabs, n, (flags & kGemmAccum) ? 1 : 0);
    return after_launch("slab_reduce_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_linear_bwd_data_f32(const float *dy, const float *w, const float *relu_out, float *dx,
                                           int64_t M, int32_t N, int32_t K, void *stream) {
    B200MED_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad shape");
    if (M == 0) return B200MED_OK;
    B200MED_REQUIRE(dy && w && dx, "null pointer");
    // dx[m,k] = sum_n dy[m,n] * w[n,k]:  A = dy (row m, r = n contiguous), B(j=k, r=n) = w[n*K + k]
    return launch_gemm(dy, w, dx, M, K, N, N, 1, 1, K, K, nullptr, 0, relu_out, K, 1, 0, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int64_t b200med_linear_bwd_weight_ws_bytes(int64_t M, int32_t N, int32_t K) {
    const long long a = weight_slabs(M, N, K) * (long long)N * K * 4;
    const long long b = colsum_slabs(M) * (long long)N * 4;
    return a + b + 256;
}

extern "C" __attribute__((visibility("default"))) int b200med_linear_bwd_weight_f32(const float *dy, const float *x, float *dw, float *db, int64_t M,
                                             int32_t N, int32_t K, int32_t accumulate, void *workspace,
                                             void *stream) {
    B200MED_REQUIRE(M >= 1 && N >= 1 && K >= 1, "bad shape");
    B200MED_REQUIRE(dy && x && dw && workspace, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = (int)weight_slabs(M, N, K);
    float *part = reinterpret_cast<float *>(workspace);
    // dW[n,k] = sum_m dy[m,n] * x[m,k]:  A(i=n, r=m) = dy[m*N + n], B(j=k, r=m) = x[m*K + k]
    if (int e = launch_gemm(dy, x, part, N, K, M, 1, N, 1, K, K, nullptr, 0, nullptr, 0, slabs, (long long)N * K, st))
        return e;
    const long long n = (long long)N * K;
    const long long blocks = (n + 255) / 256;
    launch_k(slab_reduce_kernel, (unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, st, part, dw, n, slabs, n, accumulate);
    if (int e = after_launch("slab_reduce_kernel")) return e;
    if (db) {
        float *cpart = part + (long long)slabs * n;
        const int cs = (int)colsum_slabs(M);
        const long long rows = (M + cs - 1) / cs;
        dim3 grid((unsigned)((N + 31) / 32), (unsigned)cs);
        launch_k(colsum_partial_kernel<float>, grid, 256, 0, st, dy, cpart, M, N, N, rows);
        if (int e = after_launch("colsum_partial_kernel")) return e;
        launch_k(slab_reduce_kernel, (unsigned)((N + 255) / 256), 256, 0, st, cpart, db, N, cs, N, accumulate);
        if (int e = after_launch("slab_reduce_kernel")) return e;
    }
    return B200MED_OK;
}

extern "C" __attribute__((visibility("default"))) int64_t b200med_colsum_ws_bytes(int64_t M, int32_t N) {
    const long long s = colsum_slabs(M) > colsum_slabs_flat(M) ? colsum_slabs(M) : colsum_slabs_flat(M);
    return s * (long long)N * 4 + 256;
}

extern "C" __attribute__((visibility("default"))) int b200med_colsum(const void *dy, int32_t dtype, float *db, int64_t M, int32_t N, int64_t ld,
                              void *workspace, void *stream) {
    B200MED_REQUIRE(M >= 1 && N >= 1 && ld >= N, "bad shape");
    B200MED_REQUIRE(dy && db && workspace, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    float *cpart = reinterpret_cast<float *>(workspace);
    const int cs = (int)colsum_slabs(M);
    const long long rows = (M + cs - 1) / cs;
    const int V = dtype == B200MED_F32 ? 4 : 8;
    const int lpr = N % V == 0 ? N / V : 0;
    if (lpr >= 1 && lpr <= 256 && (lpr & (lpr - 1)) == 0 && ld % V == 0 && (uintptr_t)dy % 16 == 0 && M >= 4096) {
        // large matrices: flat kernel, ~4 CTAs per SM worth of row slabs (the workspace holds up to colsum_slabs_flat() slabs)
        const int slabs = (int)colsum_slabs_flat(M);
        const long long rps = (M + slabs - 1) / slabs;
        if (dtype == B200MED_F32)
            launch_k(colsum_partial_flat_kernel<float>, slabs, 256, 0, st, reinterpret_cast<const float *>(dy), cpart, M, N, ld, rps);
        else
            launch_k(colsum_partial_flat_kernel<__nv_bfloat16>, slabs, 256, 0, st, reinterpret_cast<const __nv_bfloat16 *>(dy), cpart, M, N, ld, rps);
        if (int e = after_launch("colsum_partial_flat_kernel")) return e;
        launch_k(slab_reduce8_kernel, (unsigned)((N + 31) / 32), 256, 0, st, cpart, db, N, slabs);
        return after_launch("slab_reduce8_kernel");
    }
    if (N % V == 0 && ld % V == 0 && (uintptr_t)dy % 16 == 0) {
        dim3 grid((unsigned)((N + 32 * V - 1) / (32 * V)), (unsigned)cs);
        if (dtype == B200MED_F32)
            launch_k(colsum_partial_vec_kernel<float>, grid, 256, 0, st, reinterpret_cast<const float *>(dy), cpart, M, N, ld, rows);
        else
            launch_k(colsum_partial_vec_kernel<__nv_bfloat16>, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16 *>(dy), cpart, M, N, ld, rows);
        if (int e = after_launch("colsum_partial_vec_kernel")) return e;
    } else {
        dim3 grid((unsigned)((N + 31) / 32), (unsigned)cs);
        if (dtype == B200MED_F32)
            launch_k(colsum_partial_kernel<float>, grid, 256, 0, st, reinterpret_cast<const float *>(dy), cpart, M, N, ld, rows);
        else
            launch_k(colsum_partial_kernel<__nv_bfloat16>, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16 *>(dy), cpart, M, N, ld, rows);
        if (int e = after_launch("colsum_partial_kernel")) return e;
    }
    launch_k(slab_reduce8_kernel, (unsigned)((N + 31) / 32), 256, 0, st, cpart, db, N, cs);
    return after_launch("slab_reduce8_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_cast_f32_to_bf16(const float *x, void *y, int64_t n, void *stream) {
    if (n <= 0) return B200MED_OK;
    B200MED_REQUIRE(x && y, "null pointer");
    const long long blocks = (n + 255) / 256, cap = (long long)num_sms() * 16;
    launch_k(cast_f32_bf16_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, 
        x, reinterpret_cast<__nv_bfloat16 *>(y), n);
    return after_launch("cast_f32_bf16_kernel");
}
extern "C" __attribute__((visibility("default"))) int b200med_relu_cast_f32_to_bf16(const float *x, void *y, int64_t n, void *stream) {
    if (n <= 0) return B200MED_OK;
    B200MED_REQUIRE(x && y, "null pointer");
    const long long blocks = (n + 255) / 256, cap = (long long)num_sms() * 16;
    launch_k(relu_cast_f32_bf16_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, 
        x, reinterpret_cast<__nv_bfloat16 *>(y), n);
    return after_launch("relu_cast_f32_bf16_kernel");
}
extern "C" __attribute__((visibility("default"))) int b200med_split_bf16x3(const float *x, void *row3, void *stack3, int64_t R, int32_t C,
                                                                           int32_t row_order, int32_t stack_order, int32_t relu,
                                                                           void *stream) {
    if (R <= 0 || C <= 0) return B200MED_OK;
    B200MED_REQUIRE(x && (row3 || stack3), "null pointer");
    B200MED_REQUIRE((row_order == 0 || row_order == 1) && (stack_order == 0 || stack_order == 1), "order: 0 = left operand, 1 = right operand");
    const long long n = R * (long long)C, blocks = (n + 255) / 256, cap = (long long)num_sms() * 16;
    launch_k(split_bf16x3_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, 
        x, reinterpret_cast<__nv_bfloat16 *>(row3), reinterpret_cast<__nv_bfloat16 *>(stack3), R, C, row_order, stack_order, relu);
    return after_launch("split_bf16x3_kernel");
}
extern "C" __attribute__((visibility("default"))) int b200med_split_bf16x6(const float *x, void *row6, void *stack6, int64_t R, int32_t C,
                                                                           int32_t role, int32_t relu, void *stream) {
    if (R <= 0 || C <= 0) return B200MED_OK;
    B200MED_REQUIRE(x && (row6 || stack6), "null pointer");
    B200MED_REQUIRE(role == 0 || role == 1, "role: 0 = left operand, 1 = right operand");
    const long long n = R * (long long)C, cap = (long long)num_sms() * 16;
    const bool vec = C % 4 == 0 && (uintptr_t)x % 16 == 0 && (!row6 || (uintptr_t)row6 % 8 == 0) && (!stack6 || (uintptr_t)stack6 % 8 == 0);
    const long long blocks = ((vec ? n / 4 : n) + 255) / 256;
    if (vec) launch_k(split_bf16x6_kernel<4>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream,
                      x, reinterpret_cast<__nv_bfloat16 *>(row6), reinterpret_cast<__nv_bfloat16 *>(stack6), R, C, role, relu);
    else launch_k(split_bf16x6_kernel<1>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream,
                  x, reinterpret_cast<__nv_bfloat16 *>(row6), reinterpret_cast<__nv_bfloat16 *>(stack6), R, C, role, relu);
    return after_launch("split_bf16x6_kernel");
}
extern "C" __attribute__((visibility("default"))) int b200med_cast_bf16_to_f32(const void *x, float *y, int64_t n, void *stream) {
    if (n <= 0) return B200MED_OK;
    B200MED_REQUIRE(x && y, "null pointer");
    const long long blocks = (n + 255) / 256, cap = (long long)num_sms() * 16;
    launch_k(cast_bf16_f32_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16 *>(x), y, n);
    return after_launch("cast_bf16_f32_kernel");
}
