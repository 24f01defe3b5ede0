// TeCNo frame head: the dilated residual stack of MultiStageModel (MED/modeling/models_TCN.py:76-137) as fused fp32
// kernels (path row a10).
//
// Layout.  The reference runs Conv1d over [1, C, T] (channel-major).  Here every activation of a stage is TIME-major,
// A[T, 64]: one frame = one 256-byte row, so a kernel tap "t + off" is a row offset, a tile of 16 frames is 4 KB of
// contiguous memory, and the 1x1 input convolution is the plain GEMM x[T, F] W[64, F]^T on the table rows as they lie
// in HBM (no [F, T] transpose of the 2048-d frame features is ever made).  Only the stage logits are written [C, T],
// the layout the reference returns and the frame loss kernel (b200med_ce_frame) reads.
//
// One DilatedResidualLayer = ONE launch (torch: conv, ReLU, slice, conv, dropout, add = 6):
//   y   = relu(b_d + sum_k Wd[:, :, k] x[t + off_k])           off = {-2d, -d, 0} causal, {-d, 0, +d} otherwise
//   out = x + dropout(b_1 + W1 y)
// A CTA (4 warps) owns TT = 4*RPW consecutive frames and all 64 channels.  The layer's weights (48 KB + 16 KB, pre-transposed
// once per step by tcn_pack_kernel so that the output channel is the contiguous index) arrive in shared memory by ONE
// cp.async.bulk (TMA 1-D copy, mbarrier complete_tx) issued by thread 0 while all threads stage the three tap tiles with
// 128-bit loads.  A warp owns RPW frames; a lane owns 4 output channels (lane & 15) and HALF of the input channels
// (lane >> 4), the two halves meet in one shuffle: with one video (T ~ 600) the kernel is a latency chain per thread, so
// the chain is cut (RPW = 1: 150 CTAs, 512 FMA per thread) instead of the tile grown; long ragged batches use RPW = 4
// (weights staged once per 16 frames).  First version (16 frames per CTA, 2 frames x 4 channels x all inputs per
// thread): 38 CTAs, 25 % issue utilisation, 10 us per layer (profiles/r1_ncu_tcn.md).
// Backward per layer = two launches:
//   tcn_layer_bwd_hidden: dz = dout * mask, dpre = (dz W1) * (y > 0), and the layer's weight/bias gradient PARTIALS
//                         (outer products of the frame tiles, 144 register accumulators per thread, the record staged
//                         through shared memory and copied out coalesced, slots summed later in fixed order by
//                         tcn_reduce_grads_kernel: deterministic, no atomics);
//   tcn_layer_bwd_input:  dx = dout + sum_k Wd[:, :, k]^T dpre[t - off_k]   (the same 3-tap tile product, transposed pack).
// Ragged batches: `tloc` / `trem` (frame index inside its video / frames left after it) let several videos be
// concatenated along T without taps crossing a video boundary (ensemble inference, BASELINE config 5); NULL = one video.
// Bulk inference in the bf16 mode runs the layer on the tensor cores instead (tcn_layer_fwd_bf16_kernel, further down).
#include "tcgen05.cuh"

namespace b200med {

constexpr int kF = 64;            // feature maps (mstcn_f_maps)
constexpr int kTcnThreads = 128;  // 4 warps; a warp = RPW frames x (16 channel groups x 2 input-channel halves)
constexpr int kWd = 3 * kF * kF;  // 12288
constexpr int kW1 = kF * kF;      // 4096
// pack of one layer (floats): WdF [k][ci][co] | W1F [ci][co] | WdB [k][co][ci] | W1B [co][ci] | b_d [64] | b_1 [64]
constexpr int kOffWdF = 0, kOffW1F = kWd, kOffWdB = kWd + kW1, kOffW1B = 2 * kWd + kW1, kOffBias = 2 * kWd + 2 * kW1;
constexpr int kPackFloats = kOffBias + 2 * kF;  // 32896 == B200MED_TCN_PACK_FLOATS
// gradient record of one layer (floats): dWd [co][ci][k] | dW1 [co][ci] | db_d [64] | db_1 [64]  (torch parameter layouts)
constexpr int kGradFloats = kWd + kW1 + 2 * kF;  // 16512 == B200MED_TCN_GRAD_FLOATS
constexpr int kMaxSlots = 80;
constexpr int kMaxClasses = 8;
static_assert(kPackFloats == B200MED_TCN_PACK_FLOATS && kGradFloats == B200MED_TCN_GRAD_FLOATS, "header constants");

struct TcnGeom {
    long long T;
    const int32_t *tloc, *trem;
    int off0, off1, off2;
    int centre;  // tap whose offset is 0
    __device__ __forceinline__ int off(int k) const { return k == 0 ? off0 : (k == 1 ? off1 : off2); }
};

__device__ __forceinline__ bool tap_ok(const TcnGeom &g, long long t, int off) {
    const long long tl = g.tloc ? (long long)g.tloc[t] : t;
    const long long tr = g.trem ? (long long)g.trem[t] : g.T - 1 - t;
    return tl + off >= 0 && off <= tr;
}

__device__ __forceinline__ bool tcn_keep(unsigned long long seed, unsigned long long index, float p) {
    uint64_t z = index + 0x9E3779B97F4A7C15ull * (seed + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f) >= p;
}

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(s_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(s_addr(bar)) : "memory");
}

// Arm `bar` and start the weight copy (thread 0 only; the caller syncs the CTA before anyone waits).
__device__ __forceinline__ void start_weight_copy(float *dst, const float *src, uint32_t bytes, uint64_t *bar) {
    if (threadIdx.x == 0) {
        bar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        bar_expect_tx(bar, bytes);
        bulk_g2s(dst, src, bytes, bar);
    }
}

// xs[k][r][0..63] = src[t0 + r + sign*off_k][:] (zero where the tap leaves the video or the tile leaves the table).
template <int TT>
__device__ __forceinline__ void stage_taps(float *xs, const float *__restrict__ src, const TcnGeom &g, long long t0, int sign) {
    for (int e = threadIdx.x; e < 3 * TT * (kF / 4); e += kTcnThreads) {
        const int k = e / (TT * (kF / 4)), r = (e / (kF / 4)) % TT, c4 = e % (kF / 4);
        const long long t = t0 + r;
        const int off = sign * g.off(k);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < g.T && tap_ok(g, t, off)) v = __ldg(reinterpret_cast<const float4 *>(src + (t + off) * kF) + c4);
        *reinterpret_cast<float4 *>(xs + (k * TT + r) * kF + c4 * 4) = v;
    }
}

// acc[j][c] += sum over this lane's half of i (32 of the 64 inputs) of rows[j][i] * w[i*64 + c]; w already points at the
// lane's 4 output channels.  half_sum() adds the two halves (lanes l and l ^ 16): afterwards both hold the full sums.
template <int RPW>
__device__ __forceinline__ void mac_rows(float (&acc)[RPW][4], const float *rows, const float *w, int half) {
#pragma unroll 2
    for (int i = half * (kF / 2); i < (half + 1) * (kF / 2); i += 4) {
        float xv[RPW][4];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const float4 a = *reinterpret_cast<const float4 *>(rows + j * kF + i);
            xv[j][0] = a.x; xv[j][1] = a.y; xv[j][2] = a.z; xv[j][3] = a.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 wv = *reinterpret_cast<const float4 *>(w + (i + q) * kF);
#pragma unroll
            for (int j = 0; j < RPW; ++j) {
                acc[j][0] = fmaf(xv[j][q], wv.x, acc[j][0]); acc[j][1] = fmaf(xv[j][q], wv.y, acc[j][1]);
                acc[j][2] = fmaf(xv[j][q], wv.z, acc[j][2]); acc[j][3] = fmaf(xv[j][q], wv.w, acc[j][3]);
            }
        }
    }
}
template <int RPW>
__device__ __forceinline__ void half_sum(float (&acc)[RPW][4]) {
#pragma unroll
    for (int j = 0; j < RPW; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] += __shfl_xor_sync(0xffffffffu, acc[j][c], 16);
}

// ------------------------------------------------------------------------------------------------ weight pack
// ptrs [L][4] = {conv_dilated.weight [64,64,3], conv_dilated.bias [64], conv_1x1.weight [64,64,1], conv_1x1.bias [64]}
__global__ void __launch_bounds__(256)
tcn_pack_kernel(const float *const *__restrict__ ptrs, float *__restrict__ packed) {
    pdl_wait();
    const int layer = blockIdx.y;
    const float *wd = ptrs[layer * 4 + 0], *bd = ptrs[layer * 4 + 1], *w1 = ptrs[layer * 4 + 2], *b1 = ptrs[layer * 4 + 3];
    float *out = packed + (long long)layer * kPackFloats;
    const int stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = first; i < kWd; i += stride) {
        const int co = i / (3 * kF), r = i % (3 * kF), ci = r / 3, k = r % 3;
        const float v = wd[i];
        out[kOffWdF + (k * kF + ci) * kF + co] = v;
        out[kOffWdB + (k * kF + co) * kF + ci] = v;
    }
    for (int i = first; i < kW1; i += stride) {
        const int co = i / kF, ci = i % kF;
        const float v = w1[i];
        out[kOffW1F + ci * kF + co] = v;
        out[kOffW1B + i] = v;
    }
    for (int i = first; i < kF; i += stride) {
        out[kOffBias + i] = bd[i];
        out[kOffBias + kF + i] = b1[i];
    }
}

// ------------------------------------------------------------------------------------------------ layer forward
template <int RPW>
__global__ void __launch_bounds__(kTcnThreads)
tcn_layer_fwd_kernel(const float *__restrict__ x, const float *__restrict__ pack, float *__restrict__ out,
                     float *__restrict__ y_save, TcnGeom g, float drop_p, unsigned long long seed,
                     const unsigned long long *__restrict__ seed_dev, unsigned long long drop_base) {
    pdl_wait();
    constexpr int TT = 4 * RPW;
    if (seed_dev) seed += *seed_dev;   // per-step counter in device memory: a replayed CUDA graph draws a fresh mask
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    float *ws = smem;                   // WdF | W1F : 16384 floats
    float *xs = ws + kWd + kW1;         // [3][TT][64]
    float *ys = xs + 3 * TT * kF;       // [TT][64]
    const long long t0 = (long long)blockIdx.x * TT;
    start_weight_copy(ws, pack + kOffWdF, (kWd + kW1) * 4, &bar);
    stage_taps<TT>(xs, x, g, t0, +1);
    __syncthreads();
    bar_wait(&bar, 0);

    const int lane = threadIdx.x & 31, cg = lane & 15, half = lane >> 4, r0 = (threadIdx.x >> 5) * RPW;
    float acc[RPW][4] = {};
#pragma unroll
    for (int k = 0; k < 3; ++k) mac_rows<RPW>(acc, xs + (k * TT + r0) * kF, ws + k * kW1 + cg * 4, half);
    half_sum<RPW>(acc);
    const float4 bd = __ldg(reinterpret_cast<const float4 *>(pack + kOffBias) + cg);
    const float bdv[4] = {bd.x, bd.y, bd.z, bd.w};
    if (half == 0) {
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            float4 y;
            y.x = fmaxf(acc[j][0] + bdv[0], 0.f); y.y = fmaxf(acc[j][1] + bdv[1], 0.f);
            y.z = fmaxf(acc[j][2] + bdv[2], 0.f); y.w = fmaxf(acc[j][3] + bdv[3], 0.f);
            *reinterpret_cast<float4 *>(ys + (r0 + j) * kF + cg * 4) = y;
            const long long t = t0 + r0 + j;
            if (y_save && t < g.T) *reinterpret_cast<float4 *>(y_save + t * kF + cg * 4) = y;
        }
    }
    __syncwarp();   // a warp only reads the y rows it wrote itself
    float z[RPW][4] = {};
    mac_rows<RPW>(z, ys + r0 * kF, ws + kWd + cg * 4, half);
    half_sum<RPW>(z);
    if (half != 0) return;
    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(pack + kOffBias + kF) + cg);
    const float b1v[4] = {b1.x, b1.y, b1.z, b1.w};
    const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const long long t = t0 + r0 + j;
        if (t >= g.T) continue;
        const float4 xc = *reinterpret_cast<const float4 *>(xs + (g.centre * TT + r0 + j) * kF + cg * 4);
        const float xv[4] = {xc.x, xc.y, xc.z, xc.w};
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = z[j][c] + b1v[c];
            if (drop_p > 0.f) v = tcn_keep(seed, drop_base + (unsigned long long)(t * kF + cg * 4 + c), drop_p) ? v * scale : 0.f;
            o[c] = xv[c] + v;
        }
        *reinterpret_cast<float4 *>(out + t * kF + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------------------ layer backward (hidden)
// Per TT-frame tile: dz = dout * dropout mask, dpre = (dz W1) * (y > 0) -> global; partial weight gradients of the tile
// accumulate in registers over the tiles this CTA (= slot) owns:
//   dWd[co][ci][k] += dpre[t][co] x[t + off_k][ci],  dW1[co][ci] += dz[t][co] y[t][ci],  db_d += dpre[t],  db_1 += dz[t].
template <int RPW>
__global__ void __launch_bounds__(kTcnThreads)
tcn_layer_bwd_hidden_kernel(const float *__restrict__ dout, const float *__restrict__ x, const float *__restrict__ y,
                            const float *__restrict__ pack, float *__restrict__ dpre, float *__restrict__ partials,
                            TcnGeom g, float drop_p, unsigned long long seed,
                            const unsigned long long *__restrict__ seed_dev, unsigned long long drop_base) {
    pdl_wait();
    constexpr int TT = 4 * RPW;
    if (seed_dev) seed += *seed_dev;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    float *w1b = smem;                  // [co][ci] 4096
    float *xs = w1b + kW1;              // [3][TT][64]
    float *ys = xs + 3 * TT * kF;       // [TT][64]
    float *dzs = ys + TT * kF;          // [TT][64]
    float *dps = dzs + TT * kF;         // [TT][64]
    start_weight_copy(w1b, pack + kOffW1B, kW1 * 4, &bar);
    __syncthreads();   // the barrier is initialised before any thread can reach a wait

    const int lane = threadIdx.x & 31, cg = lane & 15, half = lane >> 4, r0 = (threadIdx.x >> 5) * RPW;
    const int cog = threadIdx.x & 7, cig = threadIdx.x >> 3;   // weight-gradient tile: 8 co x (3 taps x 4 ci + 4 ci)
    float gD[8][3][4] = {}, g1[8][4] = {}, gbd[8] = {}, gb1[8] = {};
    const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const long long ntiles = (g.T + TT - 1) / TT;
    bool first = true;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long t0 = tile * TT;
        stage_taps<TT>(xs, x, g, t0, +1);
        for (int e = threadIdx.x; e < TT * (kF / 4); e += kTcnThreads) {
            const int r = e / (kF / 4), c4 = e % (kF / 4);
            const long long t = t0 + r;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f), yv = d;
            if (t < g.T) {
                d = __ldg(reinterpret_cast<const float4 *>(dout + t * kF) + c4);
                yv = __ldg(reinterpret_cast<const float4 *>(y + t * kF) + c4);
                if (drop_p > 0.f) {
                    const unsigned long long base = drop_base + (unsigned long long)(t * kF + c4 * 4);
                    d.x = tcn_keep(seed, base + 0, drop_p) ? d.x * scale : 0.f;
                    d.y = tcn_keep(seed, base + 1, drop_p) ? d.y * scale : 0.f;
                    d.z = tcn_keep(seed, base + 2, drop_p) ? d.z * scale : 0.f;
                    d.w = tcn_keep(seed, base + 3, drop_p) ? d.w * scale : 0.f;
                }
            }
            *reinterpret_cast<float4 *>(dzs + r * kF + c4 * 4) = d;
            *reinterpret_cast<float4 *>(ys + r * kF + c4 * 4) = yv;
        }
        __syncthreads();
        if (first) { bar_wait(&bar, 0); first = false; }
        float acc[RPW][4] = {};
        mac_rows<RPW>(acc, dzs + r0 * kF, w1b + cg * 4, half);
        half_sum<RPW>(acc);
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < RPW; ++j) {
                const float4 yv = *reinterpret_cast<const float4 *>(ys + (r0 + j) * kF + cg * 4);
                float4 p;
                p.x = yv.x > 0.f ? acc[j][0] : 0.f; p.y = yv.y > 0.f ? acc[j][1] : 0.f;
                p.z = yv.z > 0.f ? acc[j][2] : 0.f; p.w = yv.w > 0.f ? acc[j][3] : 0.f;
                *reinterpret_cast<float4 *>(dps + (r0 + j) * kF + cg * 4) = p;
                const long long t = t0 + r0 + j;
                if (t < g.T) *reinterpret_cast<float4 *>(dpre + t * kF + cg * 4) = p;
            }
        }
        __syncthreads();
        // weight-gradient partials of this tile (rows beyond T hold zeros in dzs / dps)
#pragma unroll 2
        for (int r = 0; r < TT; ++r) {
            const float4 p0 = *reinterpret_cast<const float4 *>(dps + r * kF + cog * 8);
            const float4 p1 = *reinterpret_cast<const float4 *>(dps + r * kF + cog * 8 + 4);
            const float4 z0 = *reinterpret_cast<const float4 *>(dzs + r * kF + cog * 8);
            const float4 z1 = *reinterpret_cast<const float4 *>(dzs + r * kF + cog * 8 + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            const float4 yv4 = *reinterpret_cast<const float4 *>(ys + r * kF + cig * 4);
            const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
            float xv[3][4];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 t4 = *reinterpret_cast<const float4 *>(xs + (k * TT + r) * kF + cig * 4);
                xv[k][0] = t4.x; xv[k][1] = t4.y; xv[k][2] = t4.z; xv[k][3] = t4.w;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i) gD[a][k][i] = fmaf(pv[a], xv[k][i], gD[a][k][i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) g1[a][i] = fmaf(zv[a], yv[i], g1[a][i]);
                gbd[a] += pv[a];
                gb1[a] += zv[a];
            }
        }
        __syncthreads();
    }
    if (first) bar_wait(&bar, 0);   // a CTA without tiles must still drain its copy before exiting
    // The record goes out through shared memory: written straight from the accumulators every lane of a store hits its own
    // 32-byte sector (ncu, first version: 27 sectors per request, 58 % of the stalls lg_throttle, 18 us per launch).  Rows
    // are padded (193 / 65 floats) against bank conflicts; the copy out is one coalesced sweep.
    __syncthreads();
    float *st = smem;
    constexpr int kRowD = 3 * kF + 1, kRow1 = kF + 1, kSt1 = kF * kRowD, kStB = kSt1 + kF * kRow1;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int co = cog * 8 + a;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 3; ++k) st[co * kRowD + (cig * 4 + i) * 3 + k] = gD[a][k][i];
            st[kSt1 + co * kRow1 + cig * 4 + i] = g1[a][i];
        }
        if (cig == 0) {
            st[kStB + co] = gbd[a];
            st[kStB + kF + co] = gb1[a];
        }
    }
    __syncthreads();
    float *part = partials + (long long)blockIdx.x * kGradFloats;
    for (int e = threadIdx.x; e < kWd; e += kTcnThreads) part[e] = st[(e / (3 * kF)) * kRowD + e % (3 * kF)];
    for (int e = threadIdx.x; e < kW1; e += kTcnThreads) part[kWd + e] = st[kSt1 + (e / kF) * kRow1 + e % kF];
    for (int e = threadIdx.x; e < 2 * kF; e += kTcnThreads) part[kWd + kW1 + e] = st[kStB + e];
}

// grads[layer][e] = sum over slots (ascending) of partials[layer][slot][e]
__global__ void __launch_bounds__(256)
tcn_reduce_grads_kernel(const float *__restrict__ partials, float *__restrict__ grads, int n_slots) {
    pdl_wait();
    const int layer = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kGradFloats) return;
    const float *p = partials + (long long)layer * n_slots * kGradFloats + e;
    float s = 0.f;
    for (int z = 0; z < n_slots; ++z) s += p[(long long)z * kGradFloats];
    grads[(long long)layer * kGradFloats + e] = s;
}

// ------------------------------------------------------------------------------------------------ layer backward (input)
// dx[t][ci] = dout[t][ci] + sum_k sum_co Wd[co][ci][k] dpre[t - off_k][co]
template <int RPW>
__global__ void __launch_bounds__(kTcnThreads)
tcn_layer_bwd_input_kernel(const float *__restrict__ dpre, const float *__restrict__ dout, const float *__restrict__ pack,
                           float *__restrict__ dx, TcnGeom g) {
    pdl_wait();
    constexpr int TT = 4 * RPW;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    float *ws = smem;             // WdB [k][co][ci]
    float *xs = ws + kWd;         // taps of dpre
    const long long t0 = (long long)blockIdx.x * TT;
    start_weight_copy(ws, pack + kOffWdB, kWd * 4, &bar);
    stage_taps<TT>(xs, dpre, g, t0, -1);
    __syncthreads();
    bar_wait(&bar, 0);
    const int lane = threadIdx.x & 31, cg = lane & 15, half = lane >> 4, r0 = (threadIdx.x >> 5) * RPW;
    float acc[RPW][4] = {};
#pragma unroll
    for (int k = 0; k < 3; ++k) mac_rows<RPW>(acc, xs + (k * TT + r0) * kF, ws + k * kW1 + cg * 4, half);
    half_sum<RPW>(acc);
    if (half != 0) return;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const long long t = t0 + r0 + j;
        if (t >= g.T) continue;
        const float4 d = __ldg(reinterpret_cast<const float4 *>(dout + t * kF) + cg);
        *reinterpret_cast<float4 *>(dx + t * kF + cg * 4) =
            make_float4(d.x + acc[j][0], d.y + acc[j][1], d.z + acc[j][2], d.w + acc[j][3]);
    }
}

// ------------------------------------------------------------------------------------------------ stage ends
// logits[c][t] = b[c] + sum_ci W[c][ci] x[t][ci]      (conv_out_classes, models_TCN.py:90,96; C <= 8)
__global__ void __launch_bounds__(128)
tcn_out_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                   float *__restrict__ logits, long long T, int C) {
    pdl_wait();
    __shared__ float ws[kMaxClasses * kF];
    __shared__ float bs[kMaxClasses];
    for (int i = threadIdx.x; i < C * kF; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < C) bs[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    float acc[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) acc[c] = c < C ? bs[c] : 0.f;
    const float4 *row = reinterpret_cast<const float4 *>(x + t * kF);
#pragma unroll 4
    for (int i = 0; i < kF / 4; ++i) {
        const float4 v = __ldg(row + i);
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
            if (c < C) {
                const float *wc = ws + c * kF + i * 4;
                acc[c] = fmaf(v.x, wc[0], acc[c]); acc[c] = fmaf(v.y, wc[1], acc[c]);
                acc[c] = fmaf(v.z, wc[2], acc[c]); acc[c] = fmaf(v.w, wc[3], acc[c]);
            }
    }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) logits[(long long)c * T + t] = acc[c];
}

// dx[t][ci] = sum_c W[c][ci] dl[c][t];   dl_t[t][c] = dl[c][t] (row-major copy for the weight-gradient GEMM)
__global__ void __launch_bounds__(256)
tcn_out_bwd_kernel(const float *__restrict__ dl, const float *__restrict__ w, float *__restrict__ dx,
                   float *__restrict__ dl_t, long long T, int C) {
    pdl_wait();
    __shared__ float ws[kMaxClasses * kF];
    for (int i = threadIdx.x; i < C * kF; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long t = idx / (kF / 4);
    const int c4 = (int)(idx % (kF / 4));
    if (t >= T) return;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) {
            const float d = __ldg(dl + (long long)c * T + t);
            const float *wc = ws + c * kF + c4 * 4;
            o[0] = fmaf(d, wc[0], o[0]); o[1] = fmaf(d, wc[1], o[1]);
            o[2] = fmaf(d, wc[2], o[2]); o[3] = fmaf(d, wc[3], o[3]);
            if (c4 == 0) dl_t[t * C + c] = d;
        }
    *reinterpret_cast<float4 *>(dx + t * kF + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// p[t][c] = softmax_c(logits[c][t])   (F.softmax(out, dim=1) between stages, models_TCN.py:48)
__global__ void __launch_bounds__(128)
tcn_softmax_fwd_kernel(const float *__restrict__ logits, float *__restrict__ p, long long T, int C) {
    pdl_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    float v[kMaxClasses], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) { v[c] = logits[(long long)c * T + t]; m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) { v[c] = expf(v[c] - m); s += v[c]; }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) p[t * C + c] = v[c] / s;
}

// dlogits[c][t] = p[t][c] * (dp[t][c] - sum_j p[t][j] dp[t][j])
__global__ void __launch_bounds__(128)
tcn_softmax_bwd_kernel(const float *__restrict__ p, const float *__restrict__ dp, float *__restrict__ dlogits, long long T,
                       int C) {
    pdl_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    float pv[kMaxClasses], dv[kMaxClasses], dot = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) { pv[c] = p[t * C + c]; dv[c] = dp[t * C + c]; dot = fmaf(pv[c], dv[c], dot); }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) dlogits[(long long)c * T + t] = pv[c] * (dv[c] - dot);
}

// ------------------------------------------------------------------------------------------------ bf16 tcgen05 layer (inference)
// Bulk inference (ragged batches of many videos: ensemble inference, BASELINE config 5) is bound by the fp32 SIMT issue rate
// of the kernels above (~17 TFLOP/s measured).  The layer is a dense contraction -- [128 frames x 192] x [192 x 64], then
// [128 x 64] x [64 x 64] -- so the inference path runs it on the tensor cores: one CTA per 128-frame tile,
//   * A operand of tap k = rows t0 + off_k .. + 127 of the bf16 activation copy [T, 64]: one frame is one 128-byte row = one
//     row of the 128B swizzle atom, so a tap is ONE TMA box load at a row offset (rows before / after the table are zero-filled
//     by TMA; rows that belong to a neighbouring video of a ragged batch are zeroed in shared memory before the MMA);
//   * B operand = the layer's weights as bf16 [tap*64 + co][ci] (K-major), one TMA box of 256 x 64;
//   * D1 (64 TMEM columns) = sum of 3 taps x 4 K-steps of tcgen05.mma (M = 128, N = 64, K = 16); epilogue 1: tcgen05.ld, + b_d,
//     ReLU, bf16, written as the K-major swizzled A operand of the 1x1 convolution; D2 = 4 more MMAs; epilogue 2: + b_1 + the
//     fp32 residual, staged through padded shared memory so that the tile leaves as two coalesced sweeps (fp32 residual stream
//     [T, 64] for the next layer's epilogue and the class convolution, bf16 copy [T, 64] for the next layer's TMA).
// The residual stream stays fp32: only the MMA operands are rounded to bf16 (2e-2 bar of the bf16 mode).
constexpr int kTileF = 128;                                  // frames per CTA = UMMA M
constexpr int kATapBytes = kTileF * kF * 2;                  // 16 KB
constexpr int kWTileBytes = 4 * kF * kF * 2;                 // 32 KB: 3 taps + 1x1
constexpr int kStageRow = kF + 1;                            // padded fp32 staging row (bank-conflict-free row-per-thread access)
constexpr size_t kSmemBf16 = 3 * kATapBytes + kWTileBytes + kATapBytes + 1024 /*align*/ + 512 /*bias*/ + 64 /*barriers, TMEM slot*/;   // 99 904 B: two CTAs per SM
static_assert(kTileF * kStageRow * 4 <= 3 * kATapBytes, "the fp32 staging tile reuses the tap buffers");

__global__ void __launch_bounds__(kTileF)
tcn_layer_fwd_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                          const float *__restrict__ res_in, float *__restrict__ res_out, __nv_bfloat16 *__restrict__ opn_out,
                          const float *__restrict__ bias /* b_d | b_1 */, int layer, TcnGeom g) {
    pdl_wait();
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *a_sm = smem;                               // [3][128 rows][128 B]
    unsigned char *w_sm = a_sm + 3 * kATapBytes;              // [256 rows][128 B]
    unsigned char *y_sm = w_sm + kWTileBytes;                 // [128 rows][128 B]
    float *bias_sm = reinterpret_cast<float *>(y_sm + kATapBytes);            // [128]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(bias_sm + 2 * kF);
    uint64_t *mma_bar = full_bar + 1;                         // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_bar + 2);
    float *stage = reinterpret_cast<float *>(a_sm);          // [128][65] fp32, valid once the tap MMAs have completed

    const int tid = threadIdx.x, warp = tid >> 5;
    const long long t0 = (long long)blockIdx.x * kTileF;
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_w) : "memory");
        bar_init(full_bar, 1); bar_init(&mma_bar[0], 1); bar_init(&mma_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    bias_sm[tid] = __ldg(bias + tid);                         // 128 threads, 128 floats
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        bar_expect_tx(full_bar, 3 * kATapBytes + kWTileBytes);
#pragma unroll
        for (int k = 0; k < 3; ++k) tma_load_2d(a_sm + k * kATapBytes, &tmap_a, full_bar, 0, (int)(t0 + g.off(k)));
        tma_load_2d(w_sm, &tmap_w, full_bar, 0, layer * 4 * kF);
    }
    // The fp32 residual tile (this CTA's rows of res_in: one contiguous 32 KB block) is fetched into registers NOW, so its
    // DRAM / L2 latency runs under the TMA loads, the tap MMAs and epilogue 1 (first version: fetched after the tap MMAs,
    // 68 % of the warp stalls were long_scoreboard -- profiles/r1_ncu_tcn.md); it is parked in shared memory once the tap
    // buffers are free.
    const long long rows = min((long long)kTileF, g.T - t0);
    float4 pre[kF / 4];
#pragma unroll
    for (int i = 0; i < kF / 4; ++i) {
        const int e = tid + i * kTileF, r = e / (kF / 4), c4 = e % (kF / 4);
        pre[i] = r < rows ? __ldg(reinterpret_cast<const float4 *>(res_in + (t0 + r) * kF) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    bar_wait(full_bar, 0);
    // ragged batches: a tap row that belongs to another video is zero (TMA already zero-filled rows outside the table)
    {
        const long long t = t0 + tid;
        if (t < g.T) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (!tap_ok(g, t, g.off(k))) {
                    uint4 *row = reinterpret_cast<uint4 *>(a_sm + k * kATapBytes + tid * 128);
#pragma unroll
                    for (int c = 0; c < 8; ++c) row[c] = make_uint4(0u, 0u, 0u, 0u);
                }
        }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    const uint32_t idesc = make_idesc(kTileF, kF, false, false);
    if (tid == 0) {
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s)       // K-major, SW128: 8-row groups 1024 B apart (SBO), one UMMA_K step = 32 B
                umma_bf16(tmem_base, make_smem_desc(s_addr(a_sm + k * kATapBytes) + s * 32, 16, 1024),
                          make_smem_desc(s_addr(w_sm + k * kF * 128) + s * 32, 16, 1024), idesc, (k | s) ? 1u : 0u);
        umma_commit(&mma_bar[0]);
    }
    // epilogue 1: y = relu(D1 + b_d) -> bf16 -> K-major swizzled operand tile of the 1x1 convolution (row = this thread)
    bar_wait(&mma_bar[0], 0);
    tcgen05_fence_after();
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    {
        uint32_t v[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            tmem_ld32(t_lane + (uint32_t)(h * 32), v);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = fmaxf(__uint_as_float(v[c * 8 + j]) + bias_sm[h * 32 + c * 8 + j], 0.f);
                const int slot = h * 4 + c;
                *reinterpret_cast<uint4 *>(y_sm + tid * 128 + ((slot ^ (tid & 7)) << 4)) =
                    make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
            }
        }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
        tcgen05_fence_after();
#pragma unroll
        for (int s = 0; s < 4; ++s)
            umma_bf16(tmem_base + kF, make_smem_desc(s_addr(y_sm) + s * 32, 16, 1024),
                      make_smem_desc(s_addr(w_sm + 3 * kF * 128) + s * 32, 16, 1024), idesc, s ? 1u : 0u);
        umma_commit(&mma_bar[1]);
    }
    // the tap buffers are free (their MMAs completed before mma_bar[0]): park the prefetched residual tile there
#pragma unroll
    for (int i = 0; i < kF / 4; ++i) {
        const int e = tid + i * kTileF, r = e / (kF / 4), c4 = e % (kF / 4);
        float *d = stage + r * kStageRow + c4 * 4;
        d[0] = pre[i].x; d[1] = pre[i].y; d[2] = pre[i].z; d[3] = pre[i].w;
    }
    __syncthreads();
    // epilogue 2: out = residual + D2 + b_1 (row = this thread), written back into the staging tile
    bar_wait(&mma_bar[1], 0);
    tcgen05_fence_after();
    {
        uint32_t v[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            tmem_ld32(t_lane + (uint32_t)(kF + h * 32), v);
            if (tid < rows) {
                float *d = stage + tid * kStageRow + h * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) d[j] += __uint_as_float(v[j]) + bias_sm[kF + h * 32 + j];
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    // coalesced sweeps out: fp32 residual stream and its bf16 copy
    for (int e = tid; e < rows * (kF / 4); e += kTileF) {
        const int r = e / (kF / 4), c4 = e % (kF / 4);
        const float *d = stage + r * kStageRow + c4 * 4;
        *reinterpret_cast<float4 *>(res_out + (t0 + r) * kF + c4 * 4) = make_float4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<uint2 *>(opn_out + (t0 + r) * kF + c4 * 4) = make_uint2(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]));
    }
    if (warp == 0) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128) : "memory");
    }
}

// wb16[layer][tap*64 + co][ci] (tap 3 = the 1x1 convolution) = bf16 of the fp32 pack's WdB | W1B block
__global__ void __launch_bounds__(256)
tcn_pack_bf16_kernel(const float *__restrict__ pack, __nv_bfloat16 *__restrict__ wb16) {
    pdl_wait();
    const int layer = blockIdx.y;
    const float *src = pack + (long long)layer * kPackFloats + kOffWdB;
    __nv_bfloat16 *dst = wb16 + (long long)layer * 4 * kF * kF;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * kF * kF; i += gridDim.x * blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
}

static int make_geom(TcnGeom &g, int64_t T, int32_t dilation, int32_t causal, const int32_t *tloc, const int32_t *trem) {
    g.T = T; g.tloc = tloc; g.trem = trem;
    if (causal) { g.off0 = -2 * dilation; g.off1 = -dilation; g.off2 = 0; g.centre = 2; }
    else        { g.off0 = -dilation;     g.off1 = 0;         g.off2 = dilation; g.centre = 1; }
    return 0;
}

constexpr size_t smem_fwd(int TT) { return (size_t)(kWd + kW1 + 3 * TT * kF + TT * kF) * 4; }    // 69.6 / 73.7 / 81.9 KB
// working set (W1 + tap / y / dz / dpre tiles), and at the end the padded gradient record staged for a coalesced copy out
constexpr size_t smem_bwd_h(int TT) {
    const size_t work = (size_t)(kW1 + 3 * TT * kF + 3 * TT * kF) * 4, record = (size_t)(kF * (3 * kF + 1) + kF * (kF + 1) + 2 * kF) * 4;
    return work > record ? work : record;   // 66.6 KB
}
constexpr size_t smem_bwd_i(int TT) { return (size_t)(kWd + 3 * TT * kF) * 4; }                   // 52.2 / 55.3 / 61.4 KB

// Frames per warp: cut the per-thread chain while the grid is small (one video), amortise the weight staging when it is
// large (ragged batches): RPW = 1 up to two waves of 4-frame CTAs, then 2, then 4.
static int pick_rpw(long long T) {
    const long long two_waves = 2LL * num_sms();
    if ((T + 3) / 4 <= two_waves) return 1;
    if ((T + 7) / 8 <= two_waves) return 2;
    return 4;
}
// The hidden backward keeps weight-gradient partials per CTA, so it wants few, longer CTAs: 8-frame tiles up to 80 slots.
static int pick_rpw_hidden(long long T) { return (T + 7) / 8 <= kMaxSlots ? 2 : 4; }

// > 48 KB of dynamic shared memory needs the opt-in; once per kernel and process (one device per process, DESIGN.md
// section 5).  Keyed by the kernel's address: the RPW instantiations share one function-pointer TYPE.
template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    static std::atomic<const void *> done[16];
    const void *key = reinterpret_cast<const void *>(kernel);
    for (auto &d : done)
        if (d.load(std::memory_order_acquire) == key) return B200MED_OK;
    const int e = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                             "cudaFuncSetAttribute(tcn)");
    if (e == B200MED_OK)
        for (auto &d : done) {
            const void *expected = nullptr;
            if (d.compare_exchange_strong(expected, key)) break;
        }
    return e;
}

}  // namespace b200med

using namespace b200med;
#define TCN_API extern "C" __attribute__((visibility("default")))

TCN_API int32_t b200med_tcn_slots(int64_t T) {
    const int tt = 4 * pick_rpw_hidden(T);
    const long long tiles = (T + tt - 1) / tt;
    return (int32_t)(tiles < 1 ? 1 : (tiles < kMaxSlots ? tiles : kMaxSlots));
}

TCN_API int b200med_tcn_pack(const void *const *param_ptrs, int32_t n_layers, float *packed, void *stream) {
    B200MED_REQUIRE(n_layers >= 1, "bad layer count");
    B200MED_REQUIRE(param_ptrs && packed, "null pointer");
    B200MED_REQUIRE((uintptr_t)packed % 16 == 0, "packed must be 16-byte aligned");
    launch_k(tcn_pack_kernel, dim3(8, (unsigned)n_layers), 256, 0, (cudaStream_t)stream, 
        reinterpret_cast<const float *const *>(param_ptrs), packed);
    return after_launch("tcn_pack_kernel");
}

TCN_API int b200med_tcn_layer_fwd(const float *x, const float *pack, float *out, float *y_save, int64_t T, int32_t dilation,
                                  int32_t causal, const int32_t *tloc, const int32_t *trem, float drop_p, uint64_t seed,
                                  const uint64_t *seed_dev, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(T >= 0 && dilation >= 1, "bad shape");
    B200MED_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability must be in [0, 1)");
    if (T == 0) return B200MED_OK;
    B200MED_REQUIRE(x && pack && out, "null pointer");
    B200MED_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)pack % 16 == 0 && (uintptr_t)out % 16 == 0 &&
                    (uintptr_t)y_save % 16 == 0, "pointers must be 16-byte aligned");
    TcnGeom g; make_geom(g, T, dilation, causal, tloc, trem);
#define TCN_LAUNCH_FWD(R)                                                                                          \
    {                                                                                                             \
        if (int e = opt_in_smem(tcn_layer_fwd_kernel<R>, smem_fwd(4 * R))) return e;                              \
        launch_k(tcn_layer_fwd_kernel<R>, (unsigned)((T + 4 * R - 1) / (4 * R)), kTcnThreads, smem_fwd(4 * R), (cudaStream_t)stream,  \
            x, pack, out, y_save, g, drop_p, seed, reinterpret_cast<const unsigned long long *>(seed_dev), drop_base); \
    }
    switch (pick_rpw(T)) {
        case 1: TCN_LAUNCH_FWD(1) break;
        case 2: TCN_LAUNCH_FWD(2) break;
        default: TCN_LAUNCH_FWD(4) break;
    }
    return after_launch("tcn_layer_fwd_kernel");
}

TCN_API int b200med_tcn_layer_bwd_hidden(const float *dout, const float *x, const float *y, const float *pack, float *dpre,
                                         float *partials, int32_t n_slots, int64_t T, int32_t dilation, int32_t causal,
                                         const int32_t *tloc, const int32_t *trem, float drop_p, uint64_t seed,
                                         const uint64_t *seed_dev, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(T >= 1 && dilation >= 1 && n_slots >= 1, "bad shape");
    B200MED_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability must be in [0, 1)");
    B200MED_REQUIRE(dout && x && y && pack && dpre && partials, "null pointer");
    B200MED_REQUIRE((uintptr_t)dout % 16 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)pack % 16 == 0 &&
                    (uintptr_t)dpre % 16 == 0, "pointers must be 16-byte aligned");
    TcnGeom g; make_geom(g, T, dilation, causal, tloc, trem);
    B200MED_REQUIRE(n_slots <= b200med_tcn_slots(T), "n_slots must not exceed b200med_tcn_slots(T)");
    if (pick_rpw_hidden(T) == 2) {
        if (int e = opt_in_smem(tcn_layer_bwd_hidden_kernel<2>, smem_bwd_h(8))) return e;
        launch_k(tcn_layer_bwd_hidden_kernel<2>, (unsigned)n_slots, kTcnThreads, smem_bwd_h(8), (cudaStream_t)stream, 
            dout, x, y, pack, dpre, partials, g, drop_p, seed, reinterpret_cast<const unsigned long long *>(seed_dev), drop_base);
    } else {
        if (int e = opt_in_smem(tcn_layer_bwd_hidden_kernel<4>, smem_bwd_h(16))) return e;
        launch_k(tcn_layer_bwd_hidden_kernel<4>, (unsigned)n_slots, kTcnThreads, smem_bwd_h(16), (cudaStream_t)stream, 
            dout, x, y, pack, dpre, partials, g, drop_p, seed, reinterpret_cast<const unsigned long long *>(seed_dev), drop_base);
    }
    return after_launch("tcn_layer_bwd_hidden_kernel");
}

TCN_API int b200med_tcn_layer_bwd_input(const float *dpre, const float *dout, const float *pack, float *dx, int64_t T,
                                        int32_t dilation, int32_t causal, const int32_t *tloc, const int32_t *trem,
                                        void *stream) {
    B200MED_REQUIRE(T >= 1 && dilation >= 1, "bad shape");
    B200MED_REQUIRE(dpre && dout && pack && dx, "null pointer");
    B200MED_REQUIRE((uintptr_t)dpre % 16 == 0 && (uintptr_t)dout % 16 == 0 && (uintptr_t)pack % 16 == 0 && (uintptr_t)dx % 16 == 0,
                    "pointers must be 16-byte aligned");
    TcnGeom g; make_geom(g, T, dilation, causal, tloc, trem);
#define TCN_LAUNCH_BWD_I(R)                                                                                        \
    {                                                                                                             \
        if (int e = opt_in_smem(tcn_layer_bwd_input_kernel<R>, smem_bwd_i(4 * R))) return e;                      \
        launch_k(tcn_layer_bwd_input_kernel<R>, (unsigned)((T + 4 * R - 1) / (4 * R)), kTcnThreads, smem_bwd_i(4 * R), (cudaStream_t)stream,  \
            dpre, dout, pack, dx, g);                                                                             \
    }
    switch (pick_rpw(T)) {
        case 1: TCN_LAUNCH_BWD_I(1) break;
        case 2: TCN_LAUNCH_BWD_I(2) break;
        default: TCN_LAUNCH_BWD_I(4) break;
    }
    return after_launch("tcn_layer_bwd_input_kernel");
}

TCN_API int b200med_tcn_reduce_grads(const float *partials, int32_t n_layers, int32_t n_slots, float *grads, void *stream) {
    B200MED_REQUIRE(n_layers >= 1 && n_slots >= 1, "bad shape");
    B200MED_REQUIRE(partials && grads, "null pointer");
    launch_k(tcn_reduce_grads_kernel, dim3((kGradFloats + 255) / 256, (unsigned)n_layers), 256, 0, (cudaStream_t)stream, 
        partials, grads, n_slots);
    return after_launch("tcn_reduce_grads_kernel");
}

TCN_API int b200med_tcn_out_fwd(const float *x, const float *w, const float *b, float *logits, int64_t T, int32_t C,
                                void *stream) {
    B200MED_REQUIRE(T >= 0 && C >= 1 && C <= kMaxClasses, "1 <= C <= 8 classes");
    if (T == 0) return B200MED_OK;
    B200MED_REQUIRE(x && w && b && logits, "null pointer");
    B200MED_REQUIRE((uintptr_t)x % 16 == 0, "x must be 16-byte aligned");
    launch_k(tcn_out_fwd_kernel, (unsigned)((T + 127) / 128), 128, 0, (cudaStream_t)stream, x, w, b, logits, T, C);
    return after_launch("tcn_out_fwd_kernel");
}

TCN_API int b200med_tcn_out_bwd(const float *dlogits, const float *w, float *dx, float *dlogits_t, int64_t T, int32_t C,
                                void *stream) {
    B200MED_REQUIRE(T >= 1 && C >= 1 && C <= kMaxClasses, "1 <= C <= 8 classes");
    B200MED_REQUIRE(dlogits && w && dx && dlogits_t, "null pointer");
    B200MED_REQUIRE((uintptr_t)dx % 16 == 0, "dx must be 16-byte aligned");
    const long long n = T * (kF / 4);
    launch_k(tcn_out_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, dlogits, w, dx, dlogits_t, T, C);
    return after_launch("tcn_out_bwd_kernel");
}

TCN_API int b200med_tcn_softmax_fwd(const float *logits, float *p, int64_t T, int32_t C, void *stream) {
    B200MED_REQUIRE(T >= 0 && C >= 1 && C <= kMaxClasses, "1 <= C <= 8 classes");
    if (T == 0) return B200MED_OK;
    B200MED_REQUIRE(logits && p, "null pointer");
    launch_k(tcn_softmax_fwd_kernel, (unsigned)((T + 127) / 128), 128, 0, (cudaStream_t)stream, logits, p, T, C);
    return after_launch("tcn_softmax_fwd_kernel");
}

TCN_API int b200med_tcn_softmax_bwd(const float *p, const float *dp, float *dlogits, int64_t T, int32_t C, void *stream) {
    B200MED_REQUIRE(T >= 1 && C >= 1 && C <= kMaxClasses, "1 <= C <= 8 classes");
    B200MED_REQUIRE(p && dp && dlogits, "null pointer");
    launch_k(tcn_softmax_bwd_kernel, (unsigned)((T + 127) / 128), 128, 0, (cudaStream_t)stream, p, dp, dlogits, T, C);
    return after_launch("tcn_softmax_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------ whole stage, one call
// The per-kernel entry points above cost one host round trip each (~100 per train step when driven from Python: the
// step was launch-bound at 2.2 ms for 0.3 ms of kernels).  These two run a whole SingleStageModel forward / backward
// from C: same kernels, same order, one call.

static inline size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

TCN_API int b200med_tcn_stage_fwd(const float *x, int32_t in_dim, int32_t softmax_in, const float *in_w, const float *in_b,
                                  const void *const *layer_ptrs, int32_t n_layers, const float *out_w, const float *out_b,
                                  int32_t C, int64_t T, int32_t causal, const int32_t *tloc, const int32_t *trem,
                                  const float *drop_p_host, uint64_t seed, const uint64_t *seed_dev, uint64_t layer_base,
                                  int32_t keep, float *p_in, float *acts, float *ys, float *pack, float *logits, void *stream) {
    B200MED_REQUIRE(T >= 1 && n_layers >= 1 && n_layers <= 30 && in_dim >= 1, "bad shape");
    B200MED_REQUIRE(x && in_w && in_b && layer_ptrs && out_w && out_b && acts && pack && logits, "null pointer");
    B200MED_REQUIRE(!keep || ys, "keep = 1 needs the y buffer");
    B200MED_REQUIRE(!softmax_in || (p_in && in_dim == C), "softmax_in needs p_in and in_dim == C");
    const float *xin = x;
    if (softmax_in) {
        if (int e = b200med_tcn_softmax_fwd(x, p_in, T, C, stream)) return e;
        xin = p_in;
    }
    if (int e = b200med_linear_fwd_f32(xin, in_w, in_b, acts, T, kF, in_dim, 0, stream)) return e;
    if (int e = b200med_tcn_pack(layer_ptrs, n_layers, pack, stream)) return e;
    const size_t plane = (size_t)T * kF;
    for (int l = 0; l < n_layers; ++l) {
        const float *src = acts + (keep ? (size_t)l : (size_t)(l & 1)) * plane;
        float *dst = acts + (keep ? (size_t)(l + 1) : (size_t)((l + 1) & 1)) * plane;
        if (int e = b200med_tcn_layer_fwd(src, pack + (size_t)l * kPackFloats, dst, keep ? ys + (size_t)l * plane : nullptr, T,
                                          1 << l, causal, tloc, trem, drop_p_host ? drop_p_host[l] : 0.f, seed, seed_dev,
                                          (layer_base + (uint64_t)l) << 40, stream))
            return e;
    }
    const float *last = acts + (keep ? (size_t)n_layers : (size_t)(n_layers & 1)) * plane;
    return b200med_tcn_out_fwd(last, out_w, out_b, logits, T, C, stream);
}

TCN_API int64_t b200med_tcn_stage_bwd_ws_bytes(int64_t T, int32_t in_dim, int32_t C, int32_t n_layers) {
    const int64_t slots = b200med_tcn_slots(T);
    int64_t wg = b200med_linear_bwd_weight_ws_bytes(T, C, kF);
    const int64_t wg2 = b200med_linear_bwd_weight_ws_bytes(T, kF, in_dim);
    if (wg2 > wg) wg = wg2;
    int64_t n = 0;
    n += 3 * (int64_t)align_up((size_t)T * kF * 4, 256);                         // dA, spare, dpre
    n += (int64_t)align_up((size_t)T * C * 4, 256);                              // dlogits^T
    n += (int64_t)align_up((size_t)n_layers * slots * kGradFloats * 4, 256);     // weight-gradient partials
    n += (int64_t)align_up((size_t)T * in_dim * 4, 256);                         // d(softmax output) before the softmax backward
    n += (int64_t)align_up((size_t)wg, 256);
    return n + 256;
}

// layer_grads [n_layers][B200MED_TCN_GRAD_FLOATS]; dx: [T, in_dim] (or [C, T] with softmax_in) or NULL when the input needs no gradient.
TCN_API int b200med_tcn_stage_bwd(const float *dlogits, const float *xin, int32_t in_dim, int32_t softmax_in, const float *in_w,
                                  const float *out_w, int32_t C, int32_t n_layers, int64_t T, int32_t causal,
                                  const int32_t *tloc, const int32_t *trem, const float *drop_p_host, uint64_t seed,
                                  const uint64_t *seed_dev, uint64_t layer_base, const float *acts, const float *ys,
                                  const float *pack, void *workspace,
                                  float *d_in_w, float *d_in_b, float *layer_grads, float *d_out_w, float *d_out_b, float *dx,
                                  void *stream) {
    B200MED_REQUIRE(T >= 1 && n_layers >= 1 && n_layers <= 30 && in_dim >= 1, "bad shape");
    B200MED_REQUIRE(dlogits && xin && in_w && out_w && acts && ys && pack && workspace && d_in_w && d_in_b && layer_grads &&
                    d_out_w && d_out_b, "null pointer");
    B200MED_REQUIRE((uintptr_t)workspace % 256 == 0, "workspace must be 256-byte aligned");
    const int slots = b200med_tcn_slots(T);
    const size_t plane = (size_t)T * kF;
    char *w = reinterpret_cast<char *>(workspace);
    auto take = [&](size_t bytes) { char *p = w; w += align_up(bytes, 256); return p; };
    float *dA = reinterpret_cast<float *>(take(plane * 4));
    float *spare = reinterpret_cast<float *>(take(plane * 4));
    float *dpre = reinterpret_cast<float *>(take(plane * 4));
    float *dl_t = reinterpret_cast<float *>(take((size_t)T * C * 4));
    float *partials = reinterpret_cast<float *>(take((size_t)n_layers * slots * kGradFloats * 4));
    float *dxin = reinterpret_cast<float *>(take((size_t)T * in_dim * 4));
    void *wg = w;
    if (int e = b200med_tcn_out_bwd(dlogits, out_w, dA, dl_t, T, C, stream)) return e;
    if (int e = b200med_linear_bwd_weight_f32(dl_t, acts + (size_t)n_layers * plane, d_out_w, d_out_b, T, C, kF, 0, wg, stream)) return e;
    for (int l = n_layers - 1; l >= 0; --l) {
        const float *pk = pack + (size_t)l * kPackFloats;
        if (int e = b200med_tcn_layer_bwd_hidden(dA, acts + (size_t)l * plane, ys + (size_t)l * plane, pk, dpre,
                                                 partials + (size_t)l * slots * kGradFloats, slots, T, 1 << l, causal, tloc, trem,
                                                 drop_p_host ? drop_p_host[l] : 0.f, seed, seed_dev,
                                                 (layer_base + (uint64_t)l) << 40, stream))
            return e;
        if (int e = b200med_tcn_layer_bwd_input(dpre, dA, pk, spare, T, 1 << l, causal, tloc, trem, stream)) return e;
        float *t = dA; dA = spare; spare = t;
    }
    if (int e = b200med_tcn_reduce_grads(partials, n_layers, slots, layer_grads, stream)) return e;
    if (int e = b200med_linear_bwd_weight_f32(dA, xin, d_in_w, d_in_b, T, kF, in_dim, 0, wg, stream)) return e;
    if (dx) {
        if (softmax_in) {
            if (int e = b200med_linear_bwd_data_f32(dA, in_w, nullptr, dxin, T, kF, in_dim, stream)) return e;
            return b200med_tcn_softmax_bwd(xin, dxin, dx, T, C, stream);
        }
        return b200med_linear_bwd_data_f32(dA, in_w, nullptr, dx, T, kF, in_dim, stream);
    }
    return B200MED_OK;
}


// Inference-only stage forward on the bf16 tcgen05 layer kernel (eval mode: no dropout, nothing kept for a backward).
//   res [2][T][64] f32 and opn [2][T][64] bf16: ping-pong residual stream / operand copy; wb16 [n_layers][256][64] bf16.
TCN_API int b200med_tcn_stage_fwd_bf16(const float *x, int32_t in_dim, int32_t softmax_in, const float *in_w, const float *in_b,
                                       const void *const *layer_ptrs, int32_t n_layers, const float *out_w, const float *out_b,
                                       int32_t C, int64_t T, int32_t causal, const int32_t *tloc, const int32_t *trem,
                                       float *p_in, float *res, void *opn, float *pack, void *wb16, float *logits, void *stream) {
    B200MED_REQUIRE(T >= 1 && n_layers >= 1 && n_layers <= 30 && in_dim >= 1, "bad shape");
    B200MED_REQUIRE(T + 2 * (1LL << (n_layers - 1)) < (1LL << 31), "T too large for 32-bit TMA coordinates");
    B200MED_REQUIRE(x && in_w && in_b && layer_ptrs && out_w && out_b && res && opn && pack && wb16 && logits, "null pointer");
    B200MED_REQUIRE(!softmax_in || (p_in && in_dim == C), "softmax_in needs p_in and in_dim == C");
    B200MED_REQUIRE((uintptr_t)opn % 128 == 0 && (uintptr_t)wb16 % 128 == 0 && (uintptr_t)res % 16 == 0, "buffers must be 128-byte aligned");
    if (!b200med_has_tcgen05()) { set_error("the bf16 TeCNo path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    const float *xin = x;
    if (softmax_in) {
        if (int e = b200med_tcn_softmax_fwd(x, p_in, T, C, stream)) return e;
        xin = p_in;
    }
    const size_t plane = (size_t)T * kF;
    __nv_bfloat16 *opn16 = reinterpret_cast<__nv_bfloat16 *>(opn);
    if (int e = b200med_linear_fwd_f32(xin, in_w, in_b, res, T, kF, in_dim, 0, stream)) return e;
    if (int e = b200med_cast_f32_to_bf16(res, opn16, (int64_t)plane, stream)) return e;
    if (int e = b200med_tcn_pack(layer_ptrs, n_layers, pack, stream)) return e;
    launch_k(tcn_pack_bf16_kernel, dim3(8, (unsigned)n_layers), 256, 0, st, pack, reinterpret_cast<__nv_bfloat16 *>(wb16));
    if (int e = after_launch("tcn_pack_bf16_kernel")) return e;
    CUtensorMap map_a[2], map_w;
    for (int i = 0; i < 2; ++i)
        if (int e = make_tmap(&map_a[i], opn16 + (size_t)i * plane, kF, T, kF, kF, kTileF)) return e;
    if (int e = make_tmap(&map_w, wb16, kF, (long long)n_layers * 4 * kF, kF, kF, 4 * kF)) return e;
    if (int e = opt_in_smem(tcn_layer_fwd_bf16_kernel, kSmemBf16)) return e;
    for (int l = 0; l < n_layers; ++l) {
        TcnGeom g; make_geom(g, T, 1 << l, causal, tloc, trem);
        const int src = l & 1, dst = (l + 1) & 1;
        launch_k(tcn_layer_fwd_bf16_kernel, (unsigned)((T + kTileF - 1) / kTileF), kTileF, kSmemBf16, st, 
            map_a[src], map_w, res + (size_t)src * plane, res + (size_t)dst * plane, opn16 + (size_t)dst * plane,
            pack + (size_t)l * kPackFloats + kOffBias, l, g);
        if (int e = after_launch("tcn_layer_fwd_bf16_kernel")) return e;
    }
    return b200med_tcn_out_fwd(res + (size_t)(n_layers & 1) * plane, out_w, out_b, logits, T, C, stream);
}
