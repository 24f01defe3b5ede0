// LSTM head recurrence (MED/modeling/models.py:161, 204-206: nn.LSTM(58, 128, num_layers=3, dropout=.2)) as ONE
// persistent kernel per layer and direction, instead of one GEMM + one cell kernel per time step (lstm.cu).
//
// The recurrence is independent per window, so a CTA owns 128 windows (TMEM lanes = rows) and walks all W steps:
//
//   forward   gates_t [128, 4H] = XG_t (x-part + bias, one big tcgen05 GEMM over all W*B rows beforehand)
//                               + h_{t-1} [128, H] * W_hh^T          <- tcgen05.mma, W_hh resident in shared memory
//             W_hh (bf16, 128 KB) is loaded ONCE by TMA; h_{t-1} lives in a 32 KB swizzled operand tile that the cell
//             epilogue rewrites every step; the gate accumulator fills the whole TMEM (512 columns = i|f|g|o x 128).
//             Epilogue (8 warps, one row per lane, 64 units per thread, c_t in registers): tcgen05.ld -> + XG ->
//             sigmoid/tanh (MUFU tanh.approx) -> c_t, h_t.  h_t goes to the operand tile (generic->async proxy fence)
//             and from there by TMA STORE to A_l[t+1] (wgrad operand); dropout(h_t) to a second tile -> A_{l+1}[t].
//
//   backward  dG_t = cell'(gates_t, c_t, c_{t-1}, dh_t, dc_t);   dh_{t-1} = dG_t [128, 4H] * W_hh [4H, H]
//             K = 4H = 512: the bf16 dG_t tile (128 KB) does not fit beside W_hh, so the epilogue hands it to the MMA
//             in four 32 KB chunks (8 units x 4 gates x 4 unit-quarters = 8 K-steps each) through a 2-deep ring; the same
//             ring slots are TMA-STOREd to dG in HBM for the batched dX = dG * W_ih and dW = dG^T [x|h] GEMMs.  So that a
//             ring slot is a contiguous 128-column block of a row-major matrix, dG's COLUMNS ARE PERMUTED:
//                 column' = chunk*128 + unit_quarter*32 + gate*8 + i   <->   gate column = gate*H + unit_quarter*32 + chunk*8 + i
//             (the host permutes W_ih's rows / un-permutes dW's rows and db accordingly, lstm_stack.py).  W_hh is loaded
//             once by TMA as 16-row boxes in exactly the MMA's K order (MN-major B operand).  The dh accumulator is
//             double-buffered in TMEM (2 x 128 columns).
//
// Global layouts.  A TMEM lane is a row, so a warp touches 32 ROWS at once; in a row-major matrix that is 32 cache
// lines per 16-byte access (measured: 19 us per step, all of it LSU wavefronts).  Tensors that are only ever produced
// and consumed by such row-per-lane code (XG and activated gates in fp16 -- bounded values, 4x finer than bf16 at the same
// size --, c and the dX of the layer above in f32) therefore use the
// ROW-BLOCK-INTERLEAVED layout  [rows/32][cols/V][32 rows][V]  (V = elements per 16 bytes): one warp access = 512
// contiguous bytes.  Tensors consumed by TMA-fed GEMMs (A_l, dG) stay row-major and are written by TMA store.
// The batch is padded to a multiple of 32 rows per time step (pad rows carry zeros / finite values and zero gradients).
#include "tcgen05.cuh"

namespace b200med {

constexpr int kRecH = 128;         // hidden size this kernel's shared-memory / TMEM plan is built for
constexpr int kRecRows = 128;      // windows per CTA (= TMEM lanes)
// lane 0 of warp 0 also issues the TMA copies and the tcgen05.mma instructions (no dedicated MMA warp: registers are per sub-partition)
constexpr uint32_t kWhhBytes = 4 * kRecH * kRecH * 2;  // 128 KB

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// Counter-based dropout mask, two decisions per 32-bit hash: keep(e) is a pure function of (seed, element index), so the
// backward pass regenerates it instead of storing it.  16-bit resolution of p.
__device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t pair_index) {
    uint32_t h = pair_index * 0x9E3779B1u + seed * 0x85EBCA77u + 0x165667B1u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}
// scale (1/(1-p)) or 0 for N consecutive elements starting at the EVEN element index e0
template <int N>
__device__ __forceinline__ void drop_scales(uint32_t seed, uint32_t e0, uint32_t thr16, float keep_scale, float *sc) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        const uint32_t h = drop_hash(seed, (e0 + j) >> 1);
        sc[j] = (h & 0xFFFFu) >= thr16 ? keep_scale : 0.0f;
        sc[j + 1] = (h >> 16) >= thr16 ? keep_scale : 0.0f;
    }
}

__device__ __forceinline__ uint4 ldg_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_f4(const void *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4 &u, float *f) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xFFFF0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float *f) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void unpack_f16x8(const uint4 &u, float *f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
        f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
}
// activated gates lie in [-1, 1]: no saturation needed
__device__ __forceinline__ uint32_t pack_f16x2_nosat(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint4 pack_f16x8(const float *f) {
    return make_uint4(pack_f16x2_nosat(f[0], f[1]), pack_f16x2_nosat(f[2], f[3]), pack_f16x2_nosat(f[4], f[5]),
                      pack_f16x2_nosat(f[6], f[7]));
}
// L2 prefetch of a contiguous range (a 32-row block of a row-block-interleaved matrix is contiguous over all columns)
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ unsigned char *align_1024(unsigned char *p) {
    return reinterpret_cast<unsigned char *>(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
}

// =========================================================================================== forward
struct RecFwdParams {
    const __half *xg;          // row-block-interleaved fp16 [W*Bpad, 4H]: x-part pre-activations + bias (i, f, o rows HALVED)
    __half *gact;              // row-block-interleaved fp16 [W*Bpad, 4H] activated gates        (kSave)
    float *c;                  // row-block-interleaved [W*Bpad, H] cell states                    (kSave)
    float *h_out;              // [B, H] row-major h_{W-1} (null: skip)
    long long B, Bpad;
    int W;
    int hoff;                  // kSave: TMA-store h_t into A_l[t+1][:, hoff:hoff+H]
    float drop_p;              // kUp: TMA-store dropout(h_t) into A_{l+1}[t][:, 0:H]
    const uint32_t *seed;
    uint32_t drop_base;
};

constexpr int kFwdThreads = 512;   // 16 warps: lane quarter = warp % 4 (TMEM rule), unit quarter = warp / 4 (32 units each)
constexpr size_t kRecFwdSmem = 1024 + kWhhBytes + 32768 + 32768 + 64;

// kSave: training (gates, c, h saved for the backward); kUp: a layer above consumes dropout(h_t); kDrop: drop_p > 0.
// W_hh / W_ih / bias rows of the sigmoid gates (i, f, o) arrive pre-multiplied by 0.5 (host, exact), so that
// sigmoid(z) = 0.5 * tanh(z/2) + 0.5 is one MUFU and one FMA.
template <bool kSave, bool kUp, bool kDrop>
__global__ void __launch_bounds__(kFwdThreads, 1)
lstm_rec_fwd_kernel(const __grid_constant__ CUtensorMap tmap_whh, const __grid_constant__ CUtensorMap tmap_next,
                    const __grid_constant__ CUtensorMap tmap_up, const RecFwdParams p) {
    pdl_wait();
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = align_1024(smem_dyn);
    unsigned char *w_sm = smem;                     // [k-block 0..1][256-row half 0..1][256 rows][128 B]
    unsigned char *h_sm = smem + kWhhBytes;         // h_t:          [k-block 0..1][128 rows][128 B], 128B-swizzled K-major
    unsigned char *u_sm = h_sm + 32768;             // dropout(h_t): same layout
    uint64_t *w_full = reinterpret_cast<uint64_t *>(u_sm + 32768);
    uint64_t *h_ready = w_full + 1, *acc_full = w_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        bar_init(w_full, 1);
        bar_init(h_ready, kFwdThreads / 32);
        bar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int m0 = blockIdx.x * kRecRows;
    const int W = p.W;

    if (threadIdx.x == 0 && W > 1) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_whh) : "memory");
        bar_expect_tx(w_full, kWhhBytes);
        for (int kb = 0; kb < 2; ++kb)
            for (int nh = 0; nh < 2; ++nh)
                tma_load_2d(w_sm + kb * 65536 + nh * 32768, &tmap_whh, w_full, kb * 64, nh * 256);
    }
    const uint32_t idesc = make_idesc(128, 256, false, false);
    const uint32_t ha = s_addr(h_sm), wa = s_addr(w_sm);

    const int q = warp & 3, uq = warp >> 2;
    const int row = q * 32 + lane;
    const long long b = (long long)m0 + row;
    // Bpad is a multiple of 32 and a warp owns 32 consecutive rows: validity is warp-uniform
    const bool ok = __shfl_sync(0xffffffffu, (int)(b < p.Bpad), 0) != 0;
    const uint32_t seed = (kDrop && p.seed) ? *p.seed : 0u;
    const float keep_scale = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const uint32_t thr16 = (uint32_t)(p.drop_p * 65536.0f);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(uq * 32);
    // this thread's 32 units live in k-block uq/2, 16-byte slots (uq%2)*4 .. +3 of its row
    unsigned char *h_row = h_sm + (uq >> 1) * 16384 + row * 128;
    unsigned char *u_row = u_sm + (uq >> 1) * 16384 + row * 128;
    const int slot0 = (uq & 1) * 4;
    float cst[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) cst[j] = 0.0f;

    // row-block index of this warp's 32 rows at step t: (t*Bpad + m0)/32 + q
    auto rblk = [&](int t) -> long long { return ((long long)t * p.Bpad + m0) / 32 + q; };
    uint4 xc[4];
    auto load_xg = [&](uint4 (&x)[4], int t, int c) {
        const __half *base = p.xg + ((rblk(t) * 64 + (uq * 32 + c * 8) / 8) * 32 + lane) * 8;
#pragma unroll
        for (int g = 0; g < 4; ++g) x[g] = ldg_u4(base + g * 16 * 256);   // gate g: + 128 columns = 16 groups of 256 elements
    };
    if (ok) {
        if (uq == 0 && lane == 0) prefetch_l2_bulk(p.xg + rblk(0) * (32 * 4 * kRecH), 32 * 4 * kRecH * 2);
        load_xg(xc, 0, 0);
    }

    for (int t = 0; t < W; ++t) {
        const bool have_acc = t > 0;
        const bool feed = t + 1 < W;
        if (have_acc) {
            bar_wait(acc_full, (uint32_t)((t - 1) & 1));
            tcgen05_fence_after();
        }
        if (ok) {
            const long long rb = rblk(t);
            if (uq == 0 && lane == 0 && t + 1 < W) prefetch_l2_bulk(p.xg + rblk(t + 1) * (32 * 4 * kRecH), 32 * 4 * kRecH * 2);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int u0 = uq * 32 + c * 8;
                uint4 xn[4];
                if (c < 3) load_xg(xn, t, c + 1);
                else if (t + 1 < W) load_xg(xn, t + 1, 0);
                float gi[8], gg[8], gf[8], go[8];
                unpack_f16x8(xc[0], gi); unpack_f16x8(xc[2], gg);
                if (have_acc) {
                    uint32_t a0[8], a2[8];
                    tmem_ld8_nowait(t_lane + (uint32_t)(0 * kRecH + c * 8), a0);
                    tmem_ld8_nowait(t_lane + (uint32_t)(2 * kRecH + c * 8), a2);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) { gi[j] += __uint_as_float(a0[j]); gg[j] += __uint_as_float(a2[j]); }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { gi[j] = fmaf(0.5f, tanh_fast(gi[j]), 0.5f); gg[j] = tanh_fast(gg[j]); }
                unpack_f16x8(xc[1], gf); unpack_f16x8(xc[3], go);
                if (have_acc) {
                    uint32_t a1[8], a3[8];
                    tmem_ld8_nowait(t_lane + (uint32_t)(1 * kRecH + c * 8), a1);
                    tmem_ld8_nowait(t_lane + (uint32_t)(3 * kRecH + c * 8), a3);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) { gf[j] += __uint_as_float(a1[j]); go[j] += __uint_as_float(a3[j]); }
                }
                float h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    gf[j] = fmaf(0.5f, tanh_fast(gf[j]), 0.5f);
                    go[j] = fmaf(0.5f, tanh_fast(go[j]), 0.5f);
                    cst[c * 8 + j] = fmaf(gf[j], cst[c * 8 + j], gi[j] * gg[j]);
                    h[j] = go[j] * tanh_fast(cst[c * 8 + j]);
                }
                if (kSave) {
                    __half *gdst = p.gact + ((rb * 64 + u0 / 8) * 32 + lane) * 8;
                    *reinterpret_cast<uint4 *>(gdst) = pack_f16x8(gi);
                    *reinterpret_cast<uint4 *>(gdst + 16 * 256) = pack_f16x8(gf);
                    *reinterpret_cast<uint4 *>(gdst + 32 * 256) = pack_f16x8(gg);
                    *reinterpret_cast<uint4 *>(gdst + 48 * 256) = pack_f16x8(go);
                    float *cdst = p.c + ((rb * 32 + u0 / 4) * 32 + lane) * 4;
                    *reinterpret_cast<float4 *>(cdst) = make_float4(cst[c * 8], cst[c * 8 + 1], cst[c * 8 + 2], cst[c * 8 + 3]);
                    *reinterpret_cast<float4 *>(cdst + 128) = make_float4(cst[c * 8 + 4], cst[c * 8 + 5], cst[c * 8 + 6], cst[c * 8 + 7]);
                }
                // operand tile of the next step's MMA and source of the TMA store (128B swizzle: 16-byte slot ^ row%8)
                *reinterpret_cast<uint4 *>(h_row + (((slot0 + c) ^ (row & 7)) << 4)) = pack_bf16x8(h);
                if (kUp) {
                    float hv[8];
                    if (kDrop) {
                        float sc[8];
                        drop_scales<8>(seed, p.drop_base + (uint32_t)(((long long)t * p.Bpad + b) * kRecH + u0), thr16, keep_scale, sc);
#pragma unroll
                        for (int j = 0; j < 8; ++j) hv[j] = h[j] * sc[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) hv[j] = h[j];
                    }
                    *reinterpret_cast<uint4 *>(u_row + (((slot0 + c) ^ (row & 7)) << 4)) = pack_bf16x8(hv);
                }
                if (p.h_out && t == W - 1 && b < p.B) {
                    float4 *dst = reinterpret_cast<float4 *>(p.h_out + b * kRecH + u0);
                    dst[0] = make_float4(h[0], h[1], h[2], h[3]);
                    dst[1] = make_float4(h[4], h[5], h[6], h[7]);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) xc[g] = xn[g];
            }
        }
        if (feed || kSave || kUp) {
            fence_proxy_async_smem();    // h_t / dropout(h_t) (generic-proxy stores) -> visible to tcgen05.mma and TMA
            tcgen05_fence_before();      // this step's tcgen05.ld are complete before the next MMA overwrites TMEM
            __syncwarp();
            if (lane == 0) bar_arrive(h_ready);
            if (warp == 0) {
                if (lane == 0) {
                    bar_wait(h_ready, (uint32_t)(t & 1));      // all 16 warps have published their part of h_t
                    tcgen05_fence_after();
                    if (kSave && feed) {
                        tma_store_3d(&tmap_next, h_sm, p.hoff, m0, t + 1);
                        tma_store_3d(&tmap_next, h_sm + 16384, p.hoff + 64, m0, t + 1);
                    }
                    if (kUp) {
                        tma_store_3d(&tmap_up, u_sm, 0, m0, t);
                        tma_store_3d(&tmap_up, u_sm + 16384, 64, m0, t);
                    }
                    bulk_commit();
                    if (feed) {
                        if (t == 0) bar_wait(w_full, 0);
                        // gates_{t+1} = h_t * W_hh^T
#pragma unroll
                        for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t da = make_smem_desc(ha + kb * 16384 + k * 32, 16, 1024);
                                    const uint64_t db = make_smem_desc(wa + kb * 65536 + nh * 32768 + k * 32, 16, 1024);
                                    umma_bf16(tmem_base + (uint32_t)(nh * 256), da, db, idesc, (kb | k) ? 1u : 0u);
                                }
                            }
                        }
                    }
                    bulk_wait_read<0>();    // the tiles may be rewritten once the commit below releases the other warps
                    if (feed) umma_commit(acc_full);
                }
                __syncwarp();
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// =========================================================================================== backward
struct RecBwdParams {
    const __half *gact;         // row-block-interleaved fp16 [W*Bpad, 4H] activated gates
    const float *c;             // row-block-interleaved [W*Bpad, H]
    const float *dh_top;        // [B, H] row-major gradient of h_{W-1} (top layer), or null
    const float *dh_up;         // row-block-interleaved [W*Bpad, up_cols] f32: dX of the layer above, or null
    int up_cols;
    long long B, Bpad;
    int W;
    float drop_p;
    const uint32_t *seed;
    uint32_t drop_base;
};

constexpr int kBwdThreads = 512;   // 16 warps: lane quarter = warp % 4, unit quarter = warp / 4 (32 units each)
constexpr size_t kRecBwdSmem = 1024 + kWhhBytes + 65536 + 128;

// K order of the dh GEMM (and column order of dG in HBM): k' = chunk*128 + unit_quarter*32 + gate*8 + i stands for gate
// column gate*H + unit_quarter*32 + chunk*8 + i.  Chunk c of a step = the 8 units every thread finishes in its c-th pass.
template <bool kDrop>
__global__ void __launch_bounds__(kBwdThreads, 1)
lstm_rec_bwd_kernel(const __grid_constant__ CUtensorMap tmap_whh, const __grid_constant__ CUtensorMap tmap_dg,
                    const RecBwdParams p) {
    pdl_wait();
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = align_1024(smem_dyn);
    unsigned char *b_sm = smem;                 // W_hh as MN-major B operand: [k-block 0..7][n-half 0..1][64 k][128 B]
    unsigned char *a_sm = smem + kWhhBytes;     // dG chunk ring: [slot 0..1][k-block 0..1][128 rows][128 B]
    uint64_t *w_full = reinterpret_cast<uint64_t *>(a_sm + 65536);
    uint64_t *chunk_ready = w_full + 1;         // [2], 16 arrivals (one per warp)
    uint64_t *slot_free = w_full + 3;           // [2], 2 arrivals: MMA done with the slot, TMA store done reading it
    uint64_t *acc_full = w_full + 5;            // tcgen05.commit
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        bar_init(w_full, 1);
        bar_init(&chunk_ready[0], kBwdThreads / 32); bar_init(&chunk_ready[1], kBwdThreads / 32);
        bar_init(&slot_free[0], 2); bar_init(&slot_free[1], 2);
        bar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int m0 = blockIdx.x * kRecRows;
    const int W = p.W;

    if (threadIdx.x == 0 && W > 1) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_whh) : "memory");
        bar_expect_tx(w_full, kWhhBytes);
        // 8-row boxes of W_hh in the MMA's K order (see above): k-block kb' = 2*chunk + (unit_quarter >> 1)
        for (int c = 0; c < 4; ++c)
            for (int uqq = 0; uqq < 4; ++uqq)
                for (int g = 0; g < 4; ++g)
                    for (int nh = 0; nh < 2; ++nh)
                        tma_load_2d(b_sm + (2 * c + (uqq >> 1)) * 16384 + nh * 8192 + ((uqq & 1) * 4 + g) * 1024, &tmap_whh,
                                    w_full, nh * 64, g * kRecH + uqq * 32 + c * 8);
    }
    const uint32_t idesc = make_idesc(128, 128, false, true);
    const uint32_t aa = s_addr(a_sm), ba = s_addr(b_sm);

    const int q = warp & 3, uq = warp >> 2;
    const int row = q * 32 + lane;
    const long long b = (long long)m0 + row;
    // Bpad is a multiple of 32 and a warp owns 32 consecutive rows: validity is warp-uniform
    const bool ok = __shfl_sync(0xffffffffu, (int)(b < p.Bpad), 0) != 0;
    const uint32_t seed = (kDrop && p.seed) ? *p.seed : 0u;
    const float keep_scale = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const uint32_t thr16 = (uint32_t)(p.drop_p * 65536.0f);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(uq * 32);
    unsigned char *a_row = a_sm + (uq >> 1) * 16384 + row * 128;     // + slot * 32768
    const int slot16 = (uq & 1) * 4;                                 // 16-byte slot of gate 0 within the 128-byte row
    float dc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) dc[j] = 0.0f;

    auto rblk = [&](int t) -> long long { return ((long long)t * p.Bpad + m0) / 32 + q; };
    struct Regs { uint4 ga[4]; float4 ct[2], cp[2], du[2]; };
    auto load = [&](Regs &r, int t, int c) {
        const int u = uq * 32 + c * 8;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            const long long rb = rblk(t);
            const __half *gs = p.gact + ((rb * 64 + u / 8) * 32 + lane) * 8;
#pragma unroll
            for (int g = 0; g < 4; ++g) r.ga[g] = ldg_u4(gs + g * 16 * 256);
            const float *cs = p.c + ((rb * 32 + u / 4) * 32 + lane) * 4;
            r.ct[0] = ldg_f4(cs); r.ct[1] = ldg_f4(cs + 128);
            if (t > 0) {
                const float *cq = p.c + ((rblk(t - 1) * 32 + u / 4) * 32 + lane) * 4;
                r.cp[0] = ldg_f4(cq); r.cp[1] = ldg_f4(cq + 128);
            } else { r.cp[0] = z4; r.cp[1] = z4; }
            if (p.dh_up) {
                const float *dq = p.dh_up + ((rb * (p.up_cols / 4) + u / 4) * 32 + lane) * 4;
                r.du[0] = ldg_f4(dq); r.du[1] = ldg_f4(dq + 128);
            } else if (p.dh_top && t == W - 1 && b < p.B) {
                const float4 *dq = reinterpret_cast<const float4 *>(p.dh_top + b * kRecH + u);
                r.du[0] = __ldg(dq); r.du[1] = __ldg(dq + 1);
            } else { r.du[0] = z4; r.du[1] = z4; }
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) r.ga[g] = make_uint4(0, 0, 0, 0);
            r.ct[0] = z4; r.ct[1] = z4; r.cp[0] = z4; r.cp[1] = z4; r.du[0] = z4; r.du[1] = z4;
        }
    };
    Regs cur;
    load(cur, W - 1, 0);

    for (int t = W - 1; t >= 0; --t) {
        const int step = W - 1 - t;
        const bool have_rec = step > 0;
        const bool feed = t > 0;
        if (have_rec) {
            bar_wait(acc_full, (uint32_t)((step - 1) & 1));
            tcgen05_fence_after();
        }
        const uint32_t acc_prev = t_lane + (uint32_t)(((step - 1) & 1) * 128);
        if (ok && uq == 0 && lane == 0 && t > 0) {
            const long long rbp = rblk(t - 1);
            prefetch_l2_bulk(p.gact + rbp * (32 * 4 * kRecH), 32 * 4 * kRecH * 2);
            if (t > 1) prefetch_l2_bulk(p.c + rblk(t - 2) * (32 * kRecH), 32 * kRecH * 4);
            if (p.dh_up) prefetch_l2_bulk(p.dh_up + rbp * (32 * (long long)p.up_cols), 32 * (uint32_t)p.up_cols * 4);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int u = uq * 32 + c * 8;
            Regs nxt;
            if (c < 3) load(nxt, t, c + 1);
            else if (t > 0) load(nxt, t - 1, 0);
            uint32_t rec[8];
            if (have_rec) {
                tmem_ld8_nowait(acc_prev + (uint32_t)(c * 8), rec);
                tmem_wait_ld();
            }
            float gi[8], gf[8], gg[8], go[8];
            unpack_f16x8(cur.ga[0], gi); unpack_f16x8(cur.ga[1], gf);
            unpack_f16x8(cur.ga[2], gg); unpack_f16x8(cur.ga[3], go);
            const float ct[8] = {cur.ct[0].x, cur.ct[0].y, cur.ct[0].z, cur.ct[0].w, cur.ct[1].x, cur.ct[1].y, cur.ct[1].z, cur.ct[1].w};
            const float cp[8] = {cur.cp[0].x, cur.cp[0].y, cur.cp[0].z, cur.cp[0].w, cur.cp[1].x, cur.cp[1].y, cur.cp[1].z, cur.cp[1].w};
            float dh[8] = {cur.du[0].x, cur.du[0].y, cur.du[0].z, cur.du[0].w, cur.du[1].x, cur.du[1].y, cur.du[1].z, cur.du[1].w};
            if (kDrop) {
                float sc8[8];
                drop_scales<8>(seed, p.drop_base + (uint32_t)(((long long)t * p.Bpad + b) * kRecH + u), thr16, keep_scale, sc8);
#pragma unroll
                for (int j = 0; j < 8; ++j) dh[j] *= sc8[j];
            }
            if (have_rec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) dh[j] += __uint_as_float(rec[j]);
            }
            float di[8], df[8], dg[8], d_o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float tc = tanh_fast(ct[j]);
                const float dct = fmaf(dh[j] * go[j], 1.0f - tc * tc, dc[c * 8 + j]);
                di[j] = dct * gg[j] * gi[j] * (1.0f - gi[j]);
                df[j] = dct * cp[j] * gf[j] * (1.0f - gf[j]);
                dg[j] = dct * gi[j] * (1.0f - gg[j] * gg[j]);
                d_o[j] = dh[j] * tc * go[j] * (1.0f - go[j]);
                dc[c * 8 + j] = dct * gf[j];
            }
            const uint4 pk[4] = {pack_bf16x8(di), pack_bf16x8(df), pack_bf16x8(dg), pack_bf16x8(d_o)};
            const int slot = c & 1;
            const int n = 2 * step + (c >> 1);            // this is the n-th fill of the ring slot
            if (n > 0) bar_wait(&slot_free[slot], (uint32_t)((n - 1) & 1));
#pragma unroll
            for (int g = 0; g < 4; ++g)
                *reinterpret_cast<uint4 *>(a_row + slot * 32768 + (((slot16 + g) ^ (row & 7)) << 4)) = pk[g];
            fence_proxy_async_smem();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(&chunk_ready[slot]);
            if (warp == 0) {
                if (lane == 0) {
                    bar_wait(&chunk_ready[slot], (uint32_t)(n & 1));
                    tcgen05_fence_after();
                    // dG_t[:, columns' c*128 .. +128] -> HBM
                    tma_store_3d(&tmap_dg, a_sm + slot * 32768, c * 128, m0, t);
                    tma_store_3d(&tmap_dg, a_sm + slot * 32768 + 16384, c * 128 + 64, m0, t);
                    bulk_commit();
                    if (feed) {
                        if (n == 0 && slot == 0) bar_wait(w_full, 0);
                        // dh_{t-1} += dG_t[:, this chunk's 128 K columns] * W_hh[those rows, :]
                        const uint32_t acc = tmem_base + (uint32_t)((step & 1) * 128);
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t da = make_smem_desc(aa + slot * 32768 + kb * 16384 + ks * 32, 16, 1024);
                                const uint64_t db = make_smem_desc(ba + (2 * c + kb) * 16384 + ks * 2048, 8192, 1024);
                                umma_bf16(acc, da, db, idesc, (c | kb | ks) ? 1u : 0u);
                            }
                        }
                        umma_commit(&slot_free[slot]);
                        if (c == 3) umma_commit(acc_full);
                    } else {
                        bar_arrive(&slot_free[slot]);
                    }
                    // the PREVIOUS chunk's store has finished reading its slot once at most one group is pending
                    if (step > 0 || c > 0) {
                        bulk_wait_read<1>();
                        bar_arrive(&slot_free[slot ^ 1]);
                    }
                }
                __syncwarp();
            }
            cur = nxt;
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256u) : "memory");
    }
}

}  // namespace b200med

using namespace b200med;

// 3-D map over a time-major operand buffer A [W, Bpad, ld] bf16: box {64 columns, 128 rows, 1 step}
static int make_tmap_a(CUtensorMap *tm, const void *A, int64_t W, int64_t Bpad, int64_t ld) {
    const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)Bpad, (cuuint64_t)W};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)Bpad * ld * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    return make_tmap_nd(tm, A, 3, dims, strides, box);
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_rec_fwd(
    const void *xg, const void *whh_bf16, void *gact, float *c, void *a_next, int32_t ld_next, int32_t hoff, void *a_up,
    int32_t ld_up, float *h_out, int64_t B, int64_t Bpad, int32_t W, int32_t H, float drop_p, const uint32_t *seed,
    uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(H == kRecH, "the persistent recurrence kernel is built for hidden_size = 128");
    B200MED_REQUIRE(B >= 1 && W >= 1 && Bpad >= B && Bpad % 32 == 0 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(xg && whh_bf16, "null pointer");
    B200MED_REQUIRE(((uintptr_t)xg % 16 == 0) && ((uintptr_t)whh_bf16 % 16 == 0), "operands must be 16-byte aligned");
    B200MED_REQUIRE(!a_next || (ld_next % 8 == 0 && hoff % 8 == 0 && hoff + H <= ld_next && (uintptr_t)a_next % 16 == 0), "bad a_next");
    B200MED_REQUIRE(!a_up || (ld_up % 8 == 0 && H <= ld_up && (uintptr_t)a_up % 16 == 0), "bad a_up");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    CUtensorMap tm, tn, tu;
    if (int e = make_tmap(&tm, whh_bf16, kRecH, 4 * kRecH, kRecH, 64, 256)) return e;   // box {64 k, 256 gate rows}
    tn = tm; tu = tm;
    if (a_next) if (int e = make_tmap_a(&tn, a_next, W, Bpad, ld_next)) return e;
    if (a_up) if (int e = make_tmap_a(&tu, a_up, W, Bpad, ld_up)) return e;
    RecFwdParams p{};
    p.xg = reinterpret_cast<const __half *>(xg);
    p.gact = reinterpret_cast<__half *>(gact);
    p.c = c; p.h_out = h_out; p.B = B; p.Bpad = Bpad; p.W = W;
    p.hoff = hoff;
    p.drop_p = a_up ? drop_p : 0.0f; p.seed = seed; p.drop_base = (uint32_t)drop_base;
    const unsigned grid = (unsigned)((Bpad + kRecRows - 1) / kRecRows);
    const bool save = gact != nullptr, up = a_up != nullptr, drp = up && drop_p > 0.0f;
    B200MED_REQUIRE(save == (c != nullptr) && save == (a_next != nullptr), "gact, c and a_next are saved together (training) or not at all");
    auto launch = [&](auto kern) -> int {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRecFwdSmem),
                               "cudaFuncSetAttribute(lstm_rec_fwd)")) return e;
        launch_k(kern, grid, kFwdThreads, kRecFwdSmem, (cudaStream_t)stream, tm, tn, tu, p);
        return B200MED_OK;
    };
    int e;
    if (save) e = up ? (drp ? launch(lstm_rec_fwd_kernel<true, true, true>) : launch(lstm_rec_fwd_kernel<true, true, false>))
                     : launch(lstm_rec_fwd_kernel<true, false, false>);
    else e = up ? (drp ? launch(lstm_rec_fwd_kernel<false, true, true>) : launch(lstm_rec_fwd_kernel<false, true, false>))
                : launch(lstm_rec_fwd_kernel<false, false, false>);
    if (e) return e;
    return after_launch("lstm_rec_fwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_rec_bwd(
    const void *gact, const float *c, const void *whh_bf16, const float *dh_top, const float *dh_up, int32_t up_cols, void *dG,
    int64_t B, int64_t Bpad, int32_t W, int32_t H, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(H == kRecH, "the persistent recurrence kernel is built for hidden_size = 128");
    B200MED_REQUIRE(B >= 1 && W >= 1 && Bpad >= B && Bpad % 32 == 0 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(gact && c && whh_bf16 && dG, "null pointer");
    B200MED_REQUIRE(!dh_up || (up_cols >= H && up_cols % 4 == 0 && (uintptr_t)dh_up % 16 == 0), "bad dh_up");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    CUtensorMap tm, tg;
    if (int e = make_tmap(&tm, whh_bf16, kRecH, 4 * kRecH, kRecH, 64, 8)) return e;     // box {64 n, 8 gate rows}
    if (int e = make_tmap_a(&tg, dG, W, Bpad, 4 * kRecH)) return e;   // dG [W, Bpad, 4H] row-major, box {64, 128, 1}
    RecBwdParams p{};
    p.gact = reinterpret_cast<const __half *>(gact);
    p.c = c; p.dh_top = dh_top; p.dh_up = dh_up; p.up_cols = up_cols;
    p.B = B; p.Bpad = Bpad; p.W = W; p.drop_p = dh_up ? drop_p : 0.0f; p.seed = seed; p.drop_base = (uint32_t)drop_base;
    const unsigned grid = (unsigned)((Bpad + kRecRows - 1) / kRecRows);
    auto launch = [&](auto kern) -> int {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRecBwdSmem),
                               "cudaFuncSetAttribute(lstm_rec_bwd)")) return e;
        launch_k(kern, grid, kBwdThreads, kRecBwdSmem, (cudaStream_t)stream, tm, tg, p);
        return B200MED_OK;
    };
    if (int e = (dh_up && drop_p > 0.0f) ? launch(lstm_rec_bwd_kernel<true>) : launch(lstm_rec_bwd_kernel<false>)) return e;
    return after_launch("lstm_rec_bwd_kernel");
}
