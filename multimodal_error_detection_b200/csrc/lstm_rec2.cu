// LSTM head recurrence, second generation (MED/modeling/models.py:161, 204-206): the forward recurrence of one layer as ONE
// launch of 2-CTA clusters on 128 SMs, with the x-part of the gates FUSED into the recurrence.
//
// Why.  The first-generation kernel (lstm_rec.cu) gives a CTA 128 windows (one per TMEM lane): 64 CTAs at B = 8192, i.e. 84 of
// the 148 SMs idle, and it reads the x-part of the gates (XG, 134 MB per layer) that a separate GEMM wrote.  Here a CLUSTER of
// two CTAs owns 128 windows and runs ONE `tcgen05.mma.cta_group::2` (M = 128) per k-step:
//   * each CTA stages ITS OWN 64 windows of the A operand [x_t | h_{t-1}]  (64 rows x 128 B per k-block),
//   * each CTA holds HALF of the weights: the gate rows of 64 hidden units (B operand half, 256 rows x (Kx + 128) bf16 =
//     96 / 128 KB -- the whole [W_ih | W_hh] of a layer does not fit one SM, half of it does),
//   * the accumulator of a pair-MMA with M = 128 lands in BOTH CTAs as 64 rows x N: TMEM lanes 0..63 hold the first N/2
//     columns, lanes 64..127 the second N/2 (the "2x2" layout) -- so in each CTA all 128 lanes work, lane L and lane L + 64
//     sharing window L and splitting its hidden units.  Gate rows are ordered so that every thread finds i, f, g, o of ITS
//     units under its own lane: MMA #0 = [i f](units 0..63) | [i f](units 64..127), MMA #1 = [g o] likewise.
// No activation is exchanged between the CTAs: each CTA owns all 128 units of its 64 windows; only mbarrier signals cross.
//
// Per step t (gates_t = x_t W_ih^T + h_{t-1} W_hh^T + b, two TMEM buffers of 256 columns):
//   control thread (warp 16):  x_t tiles arrive by TMA two steps ahead; the x-part MMAs of step t+2 are issued right behind the
//                              h-part MMAs of step t+1, so they run UNDER the cell math of step t+1;
//   16 epilogue warps:         tcgen05.ld -> + bias -> sigmoid / tanh (MUFU tanh.approx; the i, f, o rows of the weights
//                              arrive halved so that sigmoid(z) = 0.5 tanh(z/2) + 0.5 is one MUFU + one FMA) -> c_t, h_t;
//                              h_t -> the (double-buffered) swizzled operand tile of step t+1 and from there by TMA store to
//                              A_l[t+1] (operand of the weight-gradient GEMM), dropout(h_t) -> A_{l+1}[t];
//                              activated gates (fp16) and c_t (f32) saved in the row-block-interleaved layout of lstm_rec.cu.
// HBM per (window, step): 128-256 B x_t in, 1 KB gates + 512 B c + 2 x 256 B h out -- the 1 KB XG write + read is gone.
#include "tcgen05.cuh"

namespace b200med {

constexpr int kR2H = 128;                 // hidden size
constexpr int kR2Rows = 64;               // windows per CTA (128 per cluster)
constexpr int kR2EpiWarps = 16;
constexpr int kR2Threads = (kR2EpiWarps + 1) * 32;   // + the control warp

__device__ __forceinline__ float r2_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t r2_drop_hash(uint32_t seed, uint32_t pair_index) {      // == drop_hash of lstm_rec.cu
    uint32_t h = pair_index * 0x9E3779B1u + seed * 0x85EBCA77u + 0x165667B1u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}
__device__ __forceinline__ void r2_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ uint32_t r2_pack_f16x2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
// 3-D TMA load whose completion is signalled on an mbarrier of the pair's LEADER CTA
__device__ __forceinline__ void tma_load_3d_pair(void *dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(s_addr(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Rec2FwdParams {
    const float *bias;         // [512] in the packed gate-row order (see lstm_pack_weights2_kernel), i / f / o entries halved
    __half *gact;              // row-block-interleaved fp16 [W*Bpad, 4H] activated gates          (kSave)
    float *c;                  // row-block-interleaved [W*Bpad, H] cell states                      (kSave)
    float *h_out;              // [B, H] row-major h_{W-1} (null: skip)
    long long B, Bpad;
    int W;
    int kx;                    // padded input width (64 or 128) = column of h_{t-1} inside A_l
    float drop_p;
    const uint32_t *seed;
    uint32_t drop_base;
};

template <int KXB>             // k-blocks (64 columns) of the x part: 1 (layer 0: 58 -> 64) or 2 (upper layers: 128)
struct Rec2Smem {
    static constexpr uint32_t kWBytes = (KXB + 2) * 32768;            // [k-block][256 gate rows][128 B]
    static constexpr uint32_t kXTile = KXB * 8192;                    // [k-block][64 rows][128 B]
    static constexpr uint32_t kHTile = 2 * 8192;
    static constexpr uint32_t kOffX = kWBytes;                        // 2 buffers
    static constexpr uint32_t kOffH = kOffX + 2 * kXTile;             // 2 buffers
    static constexpr uint32_t kOffU = kOffH + 2 * kHTile;             // 2 buffers
    static constexpr uint32_t kOffBias = kOffU + 2 * kHTile;          // 512 floats
    static constexpr uint32_t kOffBar = kOffBias + 2048;
    static constexpr uint32_t kUsed = kOffBar + 256;
    static constexpr uint32_t kBytes = kUsed + 512;                   // + slack for the 1024-byte alignment of the tiles
    static_assert(kBytes <= 232448, "over the 227 KB of shared memory a CTA can have");
};

// kSave: training (gates, c, h saved for the backward); kUp: a layer above consumes dropout(h_t); kDrop: drop_p > 0.
template <int KXB, bool kSave, bool kUp, bool kDrop>
__global__ void __launch_bounds__(kR2Threads, 1)
lstm_rec2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_up, const Rec2FwdParams p) {
    pdl_wait();
    using S = Rec2Smem<KXB>;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0 && (uint32_t)(smem - smem_dyn) + S::kUsed > S::kBytes) __trap();      // the 512 B of slack were not enough
    unsigned char *w_sm = smem;
    unsigned char *x_sm = smem + S::kOffX;
    unsigned char *h_sm = smem + S::kOffH;
    unsigned char *u_sm = smem + S::kOffU;
    float *bias_sm = reinterpret_cast<float *>(smem + S::kOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::kOffBar);
    uint64_t *w_full = bars;             // leader: expect_tx + peer arrive
    uint64_t *x_full = bars + 1;         // [2] leader: expect_tx + peer arrive
    uint64_t *x_free = bars + 3;         // [2] tcgen05.commit multicast: the x-part MMAs have read the tile
    uint64_t *acc_full = bars + 5;       // [2] tcgen05.commit multicast: gates of a step are complete
    uint64_t *h_ready = bars + 7;        // [2 buffers][2 halves] leader: the 16 epilogue warps of BOTH CTAs have published that half of h_t
    uint64_t *h_local = bars + 11;       // [2] the 16 epilogue warps of this CTA have published h_t / dropout(h_t)
    uint64_t *tile_free = bars + 13;     // [2] the TMA stores of the h / u tiles of two steps ago have read them
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int W = p.W;
    const int m0 = (int)(blockIdx.x >> 1) * 2 * kR2Rows + (int)rank * kR2Rows;      // first window of this CTA

    if (threadIdx.x == 0) {
        bar_init(w_full, 2);
        for (int i = 0; i < 2; ++i) {
            bar_init(&x_full[i], 2); bar_init(&x_free[i], 1); bar_init(&acc_full[i], 1);
            bar_init(&h_ready[2 * i], 2 * kR2EpiWarps); bar_init(&h_ready[2 * i + 1], 2 * kR2EpiWarps);
            bar_init(&h_local[i], kR2EpiWarps); bar_init(&tile_free[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kR2EpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 512; i += kR2Threads) bias_sm[i] = p.bias[i];
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                  // both CTAs' barriers exist before any remote arrive / TMA signal
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kR2EpiWarps) {
        // ============================================================================ control thread
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_w) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
            const uint32_t idesc = make_idesc(128, 256, false, false);
            const uint32_t wa = s_addr(w_sm), xa = s_addr(x_sm), ha = s_addr(h_sm);
            {   // weights: this CTA's 256 gate rows, every k-block; completion on the LEADER's barrier (its MMA reads both halves)
                const uint32_t lbar = mapa_rank(w_full, 0);
                if (leader) bar_expect_tx(w_full, 2 * S::kWBytes);
                for (int kb = 0; kb < KXB + 2; ++kb) tma_load_2d_pair(w_sm + kb * 32768, &tmap_w, lbar, kb * 64, (int)rank * 256);
                if (!leader) bar_arrive_cluster(lbar);
            }
            auto load_x = [&](int t) {
                const int b = t & 1;
                const uint32_t lbar = mapa_rank(&x_full[b], 0);
                if (leader) bar_expect_tx(&x_full[b], 2 * S::kXTile);
                for (int kb = 0; kb < KXB; ++kb) tma_load_3d_pair(x_sm + b * S::kXTile + kb * 8192, &tmap_a, lbar, kb * 64, m0, t);
                if (!leader) bar_arrive_cluster(lbar);
            };
            auto x_part = [&](int t) {          // leader: gates_t (buffer t & 1) = x_t W_ih^T
                const int b = t & 1;
                bar_wait(&x_full[b], (uint32_t)((t >> 1) & 1));
                tcgen05_fence_after();
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int kb = 0; kb < KXB; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = make_smem_desc(xa + b * S::kXTile + kb * 8192 + k * 32, 16, 1024);
                            const uint64_t db = make_smem_desc(wa + kb * 32768 + j * 16384 + k * 32, 16, 1024);
                            umma_bf16_pair(tmem_base + (uint32_t)(b * 256 + j * 128), da, db, idesc, (kb | k) ? 1u : 0u);
                        }
                umma_commit_pair(&x_free[b]);
            };
            // leader: gates_t += h_{t-1} W_hh^T, K-steps of hidden units [32 half, 32 half + 32) of both unit blocks: the cell math
            // publishes h in two halves (chunk 0 = units 0..31 of each block, chunk 1 = 32..63), so the first half of the K loop
            // runs UNDER the second half of the cell math
            auto h_part = [&](int t, int half) {
                const int b = t & 1, hb = (t - 1) & 1;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k2 = 0; k2 < 2; ++k2) {
                            const int k = half * 2 + k2;
                            const uint64_t da = make_smem_desc(ha + hb * S::kHTile + kb * 8192 + k * 32, 16, 1024);
                            const uint64_t db = make_smem_desc(wa + (KXB + kb) * 32768 + j * 16384 + k * 32, 16, 1024);
                            umma_bf16_pair(tmem_base + (uint32_t)(b * 256 + j * 128), da, db, idesc, 1u);
                        }
                if (half == 1) umma_commit_pair(&acc_full[b]);
            };
            load_x(0);
            if (W > 1) load_x(1);
            if (leader) {
                bar_wait(w_full, 0);
                x_part(0);
                umma_commit_pair(&acc_full[0]);          // h_{-1} = 0: the gates of step 0 are the x-part alone
                if (W > 1) x_part(1);
            }
            if (W > 2) { bar_wait(&x_free[0], 0); load_x(2); }
            for (int t = 0; t < W; ++t) {
                const int b = t & 1;
                const uint32_t par = (uint32_t)((t >> 1) & 1);
                // critical path first: the h-part MMAs of step t+1 wait for nothing but h_t (both CTAs) ...
                if (leader && t + 1 < W) {
                    bar_wait(&h_ready[2 * b], par);      // both CTAs: first half of h_t published
                    tcgen05_fence_after();
                    h_part(t + 1, 0);
                    bar_wait(&h_ready[2 * b + 1], par);  // second half published, TMEM buffer b drained
                    tcgen05_fence_after();
                    h_part(t + 1, 1);
                }
                // ... then this CTA's h_t / dropout(h_t) tiles leave for HBM (they are complete and fenced: h_local)
                bar_wait(&h_local[b], par);
                if (kSave && t + 1 < W) {
                    tma_store_3d(&tmap_a, h_sm + b * S::kHTile, p.kx, m0, t + 1);
                    tma_store_3d(&tmap_a, h_sm + b * S::kHTile + 8192, p.kx + 64, m0, t + 1);
                }
                if (kUp) {
                    tma_store_3d(&tmap_up, u_sm + b * S::kHTile, 0, m0, t);
                    tma_store_3d(&tmap_up, u_sm + b * S::kHTile + 8192, 64, m0, t);
                }
                bulk_commit();
                if (leader && t + 2 < W) x_part(t + 2);  // runs under the cell math of step t+1
                if (t + 3 < W) {                          // x tile of step t+3 into the buffer the x-part of step t+1 has read
                    bar_wait(&x_free[(t + 1) & 1], (uint32_t)(((t + 1) >> 1) & 1));
                    load_x(t + 3);
                }
                bulk_wait_read<1>();                      // every store group but this step's has read its tiles
                if (t >= 1) bar_arrive(&tile_free[(t - 1) & 1]);
            }
            bulk_wait_all();
        }
        __syncwarp();
    } else {
        // ============================================================================ cell math (16 warps)
        const int q = warp & 3, cq = warp >> 2;
        const int row_l = (q & 1) * 32 + lane;            // window within the CTA
        const int uh = q >> 1;                            // hidden-unit half: TMEM lanes 64..127 hold units 64..127
        const long long brow = (long long)m0 + row_l;
        const bool ok = __shfl_sync(0xffffffffu, (int)(brow < p.Bpad), 0) != 0;     // warp-uniform (Bpad % 32 == 0)
        const uint32_t seed = (kDrop && p.seed) ? *p.seed : 0u;
        const float keep_scale = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
        const uint32_t thr16 = (uint32_t)(p.drop_p * 65536.0f);
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 8);
        const int ubase = uh * 64 + cq * 8;               // this thread's hidden units: [ubase, +8) (chunk 0) and [ubase + 32, +8) (chunk 1)
        const float *bsm = bias_sm + uh * 256 + cq * 8;   // + (gate >> 1) * 128 + (gate & 1) * 64 + chunk * 32 + unit offset
        float cst[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) cst[j] = 0.0f;

        for (int t = 0; t < W; ++t) {
            const int b = t & 1;
            const uint32_t par = (uint32_t)((t >> 1) & 1);
            bar_wait(&acc_full[b], par);
            tcgen05_fence_after();
            if (t >= 2) bar_wait(&tile_free[b], (uint32_t)(((t - 2) >> 1) & 1));
            if (ok) {
                const long long rb = ((long long)t * p.Bpad + m0) / 32 + (q & 1);
                unsigned char *h_row = h_sm + b * S::kHTile + uh * 8192 + row_l * 128;
                unsigned char *u_row = u_sm + b * S::kHTile + uh * 8192 + row_l * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int u0 = ubase + c * 32;
                    uint32_t a[4][8];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        r2_ld8(t_lane + (uint32_t)(b * 256 + (g >> 1) * 128 + (g & 1) * 64 + c * 32), a[g]);
                    tmem_wait_ld();
                    float gi[8], gf[8], gg[8], go[8], h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float zi = __uint_as_float(a[0][j]) + bsm[c * 32 + j];
                        const float zf = __uint_as_float(a[1][j]) + bsm[64 + c * 32 + j];
                        const float zg = __uint_as_float(a[2][j]) + bsm[128 + c * 32 + j];
                        const float zo = __uint_as_float(a[3][j]) + bsm[192 + c * 32 + j];
                        gi[j] = fmaf(0.5f, r2_tanh(zi), 0.5f);
                        gf[j] = fmaf(0.5f, r2_tanh(zf), 0.5f);
                        gg[j] = r2_tanh(zg);
                        go[j] = fmaf(0.5f, r2_tanh(zo), 0.5f);
                        cst[c * 8 + j] = fmaf(gf[j], cst[c * 8 + j], gi[j] * gg[j]);
                        h[j] = go[j] * r2_tanh(cst[c * 8 + j]);
                    }
                    if (kSave) {
                        __half *gdst = p.gact + ((rb * 64 + u0 / 8) * 32 + lane) * 8;
                        *reinterpret_cast<uint4 *>(gdst) = make_uint4(r2_pack_f16x2(gi[0], gi[1]), r2_pack_f16x2(gi[2], gi[3]),
                                                                      r2_pack_f16x2(gi[4], gi[5]), r2_pack_f16x2(gi[6], gi[7]));
                        *reinterpret_cast<uint4 *>(gdst + 16 * 256) = make_uint4(r2_pack_f16x2(gf[0], gf[1]), r2_pack_f16x2(gf[2], gf[3]),
                                                                                 r2_pack_f16x2(gf[4], gf[5]), r2_pack_f16x2(gf[6], gf[7]));
                        *reinterpret_cast<uint4 *>(gdst + 32 * 256) = make_uint4(r2_pack_f16x2(gg[0], gg[1]), r2_pack_f16x2(gg[2], gg[3]),
                                                                                 r2_pack_f16x2(gg[4], gg[5]), r2_pack_f16x2(gg[6], gg[7]));
                        *reinterpret_cast<uint4 *>(gdst + 48 * 256) = make_uint4(r2_pack_f16x2(go[0], go[1]), r2_pack_f16x2(go[2], go[3]),
                                                                                 r2_pack_f16x2(go[4], go[5]), r2_pack_f16x2(go[6], go[7]));
                        float *cdst = p.c + ((rb * 32 + u0 / 4) * 32 + lane) * 4;
                        *reinterpret_cast<float4 *>(cdst) = make_float4(cst[c * 8], cst[c * 8 + 1], cst[c * 8 + 2], cst[c * 8 + 3]);
                        *reinterpret_cast<float4 *>(cdst + 128) = make_float4(cst[c * 8 + 4], cst[c * 8 + 5], cst[c * 8 + 6], cst[c * 8 + 7]);
                    }
                    // operand tile of the next step's MMA and source of the TMA store (128B swizzle: 16-byte slot ^ row % 8)
                    const int slot = (c * 4 + cq) ^ (row_l & 7);
                    *reinterpret_cast<uint4 *>(h_row + (slot << 4)) =
                        make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
                    if (kUp) {
                        float hv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hv[j] = h[j];
                        if (kDrop) {
                            const uint32_t e0 = p.drop_base + (uint32_t)(((long long)t * p.Bpad + brow) * kR2H + u0);
#pragma unroll
                            for (int j = 0; j < 8; j += 2) {
                                const uint32_t hh = r2_drop_hash(seed, (e0 + j) >> 1);
                                hv[j] = (hh & 0xFFFFu) >= thr16 ? h[j] * keep_scale : 0.0f;
                                hv[j + 1] = (hh >> 16) >= thr16 ? h[j + 1] * keep_scale : 0.0f;
                            }
                        }
                        *reinterpret_cast<uint4 *>(u_row + (slot << 4)) =
                            make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]), pack_bf16x2(hv[6], hv[7]));
                    }
                    if (p.h_out && t == W - 1 && brow < p.B) {
                        float4 *dst = reinterpret_cast<float4 *>(p.h_out + brow * kR2H + u0);
                        dst[0] = make_float4(h[0], h[1], h[2], h[3]);
                        dst[1] = make_float4(h[4], h[5], h[6], h[7]);
                    }
                    if (c == 0) {            // first half of h_t is in the operand tile: its K-steps may start
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (leader) bar_arrive(&h_ready[2 * b]);
                            else bar_arrive_cluster(mapa_rank(&h_ready[2 * b], 0));
                        }
                    }
                }
            } else if (lane == 0) {          // a warp of padding rows still takes part in the first-half handshake
                if (leader) bar_arrive(&h_ready[2 * b]);
                else bar_arrive_cluster(mapa_rank(&h_ready[2 * b], 0));
            }
            fence_proxy_async_smem();    // h_t / dropout(h_t) (generic-proxy stores) -> visible to tcgen05.mma and TMA
            tcgen05_fence_before();      // this step's tcgen05.ld are complete before the x-part of step t+2 overwrites the buffer
            __syncwarp();
            if (lane == 0) {
                bar_arrive(&h_local[b]);
                if (leader) bar_arrive(&h_ready[2 * b + 1]);
                else bar_arrive_cluster(mapa_rank(&h_ready[2 * b + 1], 0));
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                  // no CTA of the pair leaves while its partner may still signal it
    if (warp == kR2EpiWarps) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// Weights of one layer in the operand order of the kernel above.  Row rho = R*256 + j*128 + gh*64 + u of wp [512, kx + 128] holds
// gate (2j + gh) of hidden unit R*64 + u: columns [0, kx) = W_ih (zero-padded from `in`), [kx, kx + 128) = W_hh; rows of the
// sigmoid gates (i, f, o) and their biases are HALVED (sigmoid(z) = 0.5 tanh(z/2) + 0.5; exact in bf16 / fp32).
__global__ void lstm_pack_weights2_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                                          const float *__restrict__ b_ih, const float *__restrict__ b_hh, int in, int kx,
                                          __nv_bfloat16 *__restrict__ wp, float *__restrict__ bias_p) {
    pdl_wait();
    const int kc = kx + kR2H;
    const int total = 4 * kR2H * kc;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int rho = e / kc, k = e - rho * kc;
        const int R = rho >> 8, j = (rho >> 7) & 1, gh = (rho >> 6) & 1, u = rho & 63;
        const int gate = 2 * j + gh, unit = R * 64 + u;
        const int src = gate * kR2H + unit;
        const float sc = gate == 2 ? 1.0f : 0.5f;
        float v = 0.0f;
        if (k < kx) { if (k < in) v = w_ih[(long long)src * in + k]; }
        else v = w_hh[(long long)src * kR2H + (k - kx)];
        wp[e] = __float2bfloat16_rn(v * sc);
        if (k == 0) bias_p[rho] = (b_ih[src] + b_hh[src]) * sc;
    }
}

// =========================================================================================================== backward
// dG_t = cell'(gates_t, c_t, c_{t-1}, dh_t, dc_t);   [dX_t | dh_{t-1}] = dG_t [64 windows, 4H] * [W_ih | W_hh] [4H, Kx + H]
//
// Same pair scheme as the forward: each CTA owns 64 windows (all 128 TMEM lanes busy: lane L and L + 64 share a window and split
// its hidden units), the weights are split by OUTPUT column (B operand half: (Kx + 128) / 2 rows x 512 K, 96 / 128 KB), the
// K = 4H = 512 reduction runs over the whole bf16 dG_t tile (64 rows x 1 KB = 64 KB, resident) in two chunks of 256 so that the
// MMAs of chunk 0 run under the cell math of chunk 1.  The data gradient of the layer input, dX_t = dG_t W_ih -- a separate
// GEMM over all W*B rows in the first generation (134 MB of dG re-read per layer) -- falls out of the SAME accumulator:
// output columns are ordered [dX half | dh half] per lane half, so every thread finds dh of its own hidden units under its lane
// and drains a slice of dX to HBM in the layout the layer below reads.
//
// K order of the product = column order of dG in HBM (the weight-gradient GEMM consumes it; lstm_stack.py un-permutes):
//     k' = chunk*256 + unit_half*128 + warp_column*32 + gate*8 + i    <->   gate row  gate*H + unit_half*64 + warp_column*16 + chunk*8 + i
// i.e. the 64 bytes a thread produces per chunk (4 gates x 8 units, bf16) are contiguous in its row of the operand tile.
struct Rec2BwdParams {
    const __half *gact;         // row-block-interleaved fp16 [W*Bpad, 4H] activated gates
    const float *c;             // row-block-interleaved [W*Bpad, H]
    const float *dh_top;        // [B, H] row-major gradient of h_{W-1} (top layer), or null
    const float *dh_up;         // row-block-interleaved [W*Bpad, 128] f32: dX of the layer above, or null
    float *dx;                  // OUT dX of this layer: row-block-interleaved [W*Bpad, 128] (kx = 128) or row-major [W*Bpad, 64] (kx = 64)
    long long B, Bpad;
    int W;
    float drop_p;
    const uint32_t *seed;
    uint32_t drop_base;
};

constexpr int kR2BwdThreads = 512;      // 16 warps; lane 0 of warp 0 also drives TMA / MMA (registers are the limit, not threads)

template <int KXB>
struct Rec2BwdSmem {
    static constexpr int kNHalf = KXB * 32 + 64;                       // output columns per CTA: dX half + dh half (96 | 128)
    static constexpr uint32_t kBTile = kNHalf * 128;                   // one k-block of the weight half
    static constexpr uint32_t kBBytes = 8 * kBTile;                    // K = 512 = 8 k-blocks
    static constexpr uint32_t kOffA = kBBytes;                         // dG tile: [8 k-blocks][64 rows][128 B] = 64 KB
    static constexpr uint32_t kOffBar = kOffA + 65536;
    static constexpr uint32_t kUsed = kOffBar + 256;
    static constexpr uint32_t kBytes = kUsed + 1024;
    static_assert(kBytes <= 232448, "over the 227 KB of shared memory a CTA can have");
};

template <int KXB, bool kDrop>
__global__ void __launch_bounds__(kR2BwdThreads, 1)
lstm_rec2_bwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_dg, const Rec2BwdParams p) {
    pdl_wait();
    using S = Rec2BwdSmem<KXB>;
    constexpr int kNH = S::kNHalf;
    constexpr int kXH = KXB * 32;                 // dX columns per lane half
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *b_sm = smem;
    unsigned char *a_sm = smem + S::kOffA;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::kOffBar);
    uint64_t *w_full = bars;              // leader: expect_tx + peer arrive
    uint64_t *chunk_ready = bars + 1;     // [2] leader: the 16 warps of BOTH CTAs have written their part of the chunk
    uint64_t *chunk_local = bars + 3;     // [2] the 16 warps of this CTA have written their part of the chunk
    uint64_t *acc_full = bars + 5;        // tcgen05.commit multicast
    uint64_t *tile_free = bars + 6;       // the TMA stores of this step's dG tile have read it
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int W = p.W;
    const int m0 = (int)(blockIdx.x >> 1) * 2 * kR2Rows + (int)rank * kR2Rows;

    if (threadIdx.x == 0) {
        bar_init(w_full, 2);
        for (int i = 0; i < 2; ++i) { bar_init(&chunk_ready[i], 2 * 16); bar_init(&chunk_local[i], 16); }
        bar_init(acc_full, 1);
        bar_init(tile_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_dg) : "memory");
        const uint32_t lbar = mapa_rank(w_full, 0);
        if (leader) bar_expect_tx(w_full, 2 * S::kBBytes);
        for (int kb = 0; kb < 8; ++kb) tma_load_2d_pair(b_sm + kb * S::kBTile, &tmap_w, lbar, kb * 64, (int)rank * kNH);
        if (!leader) bar_arrive_cluster(lbar);
    }
    const uint32_t idesc = make_idesc(128, 2 * kNH, false, false);
    const uint32_t aa = s_addr(a_sm), ba = s_addr(b_sm);

    const int q = warp & 3, cq = warp >> 2;
    const int row_l = (q & 1) * 32 + lane;
    const int uh = q >> 1;
    const long long brow = (long long)m0 + row_l;
    const bool ok = __shfl_sync(0xffffffffu, (int)(brow < p.Bpad), 0) != 0;
    const uint32_t seed = (kDrop && p.seed) ? *p.seed : 0u;
    const float keep_scale = kDrop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const uint32_t thr16 = (uint32_t)(p.drop_p * 65536.0f);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ubase = uh * 64 + cq * 16;
    float dc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dc[j] = 0.0f;

    auto rblk = [&](int t) -> long long { return ((long long)t * p.Bpad + m0) / 32 + (q & 1); };
    struct Regs { uint4 ga[4]; float4 ct[2], cp[2], du[2]; };
    auto load = [&](Regs &r, int t, int c) {
        const int u = ubase + c * 8;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            const long long rb = rblk(t);
            const __half *gs = p.gact + ((rb * 64 + u / 8) * 32 + lane) * 8;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r.ga[g].x), "=r"(r.ga[g].y), "=r"(r.ga[g].z), "=r"(r.ga[g].w) : "l"(gs + g * 16 * 256));
            }
            const float *cs = p.c + ((rb * 32 + u / 4) * 32 + lane) * 4;
            r.ct[0] = ldg_stream(reinterpret_cast<const float4 *>(cs)); r.ct[1] = ldg_stream(reinterpret_cast<const float4 *>(cs + 128));
            if (t > 0) {
                const float *cq_ = p.c + ((rblk(t - 1) * 32 + u / 4) * 32 + lane) * 4;
                r.cp[0] = ldg_stream(reinterpret_cast<const float4 *>(cq_)); r.cp[1] = ldg_stream(reinterpret_cast<const float4 *>(cq_ + 128));
            } else { r.cp[0] = z4; r.cp[1] = z4; }
            if (p.dh_up) {
                const float *dq = p.dh_up + ((rb * 32 + u / 4) * 32 + lane) * 4;
                r.du[0] = ldg_stream(reinterpret_cast<const float4 *>(dq)); r.du[1] = ldg_stream(reinterpret_cast<const float4 *>(dq + 128));
            } else if (p.dh_top && t == W - 1 && brow < p.B) {
                const float4 *dq = reinterpret_cast<const float4 *>(p.dh_top + brow * kR2H + u);
                r.du[0] = __ldg(dq); r.du[1] = __ldg(dq + 1);
            } else { r.du[0] = z4; r.du[1] = z4; }
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) r.ga[g] = make_uint4(0, 0, 0, 0);
            r.ct[0] = z4; r.ct[1] = z4; r.cp[0] = z4; r.cp[1] = z4; r.du[0] = z4; r.du[1] = z4;
        }
    };
    // dX_t: this thread's slice of the layer-input gradient, out of the accumulator of step t
    auto drain_dx = [&](int t, uint32_t acc_col) {
        if (!ok) return;
        const long long rb = rblk(t);
        if constexpr (KXB == 2) {            // 16 columns = the hidden units [ubase, +16) of the layer below: its dh_up, interleaved
            uint32_t v[16];
            r2_ld8(t_lane + acc_col + (uint32_t)(cq * 16), v);
            r2_ld8(t_lane + acc_col + (uint32_t)(cq * 16 + 8), v + 8);
            tmem_wait_ld();
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) {
                float *dst = p.dx + ((rb * 32 + (ubase + h4 * 4) / 4) * 32 + lane) * 4;
                *reinterpret_cast<float4 *>(dst) = make_float4(__uint_as_float(v[h4 * 4]), __uint_as_float(v[h4 * 4 + 1]),
                                                               __uint_as_float(v[h4 * 4 + 2]), __uint_as_float(v[h4 * 4 + 3]));
            }
        } else {                             // 8 columns [uh*32 + cq*8, +8) of the 64-wide (padded) input gradient, row-major
            uint32_t v[8];
            r2_ld8(t_lane + acc_col + (uint32_t)(cq * 8), v);
            tmem_wait_ld();
            float4 *dst = reinterpret_cast<float4 *>(p.dx + ((long long)t * p.Bpad + brow) * 64 + uh * 32 + cq * 8);
            dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
            dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
    };

    Regs cur;
    load(cur, W - 1, 0);
    for (int t = W - 1; t >= 0; --t) {
        const int step = W - 1 - t;
        const bool have_rec = step > 0;
        const uint32_t acc_prev = (uint32_t)(((step - 1) & 1) * 128);
        if (have_rec) {
            bar_wait(acc_full, (uint32_t)((step - 1) & 1));
            tcgen05_fence_after();
            drain_dx(t + 1, acc_prev);
            bar_wait(tile_free, (uint32_t)((step - 1) & 1));       // the stores of the previous step's dG tile have read it
        }
        if (ok && cq == 0 && uh == 0 && lane == 0 && t > 0) {
            const long long rbp = rblk(t - 1);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p.gact + rbp * (32 * 4 * kR2H)), "r"(32 * 4 * kR2H * 2) : "memory");
            if (t > 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p.c + rblk(t - 2) * (32 * kR2H)), "r"(32 * kR2H * 4) : "memory");
            if (p.dh_up) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p.dh_up + rbp * (32 * kR2H)), "r"(32 * kR2H * 4) : "memory");
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int u = ubase + c * 8;
            Regs nxt;
            if (c == 0) load(nxt, t, 1);
            else if (t > 0) load(nxt, t - 1, 0);
            uint32_t rec[8];
            if (have_rec) {
                r2_ld8(t_lane + acc_prev + (uint32_t)(kXH + cq * 16 + c * 8), rec);
                tmem_wait_ld();
            }
            float gi[8], gf[8], gg[8], go[8];
            {
                const uint32_t wi[4] = {cur.ga[0].x, cur.ga[0].y, cur.ga[0].z, cur.ga[0].w};
                const uint32_t wf[4] = {cur.ga[1].x, cur.ga[1].y, cur.ga[1].z, cur.ga[1].w};
                const uint32_t wg[4] = {cur.ga[2].x, cur.ga[2].y, cur.ga[2].z, cur.ga[2].w};
                const uint32_t wo[4] = {cur.ga[3].x, cur.ga[3].y, cur.ga[3].z, cur.ga[3].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 v;
                    v = __half22float2(*reinterpret_cast<const __half2 *>(&wi[i])); gi[2 * i] = v.x; gi[2 * i + 1] = v.y;
                    v = __half22float2(*reinterpret_cast<const __half2 *>(&wf[i])); gf[2 * i] = v.x; gf[2 * i + 1] = v.y;
                    v = __half22float2(*reinterpret_cast<const __half2 *>(&wg[i])); gg[2 * i] = v.x; gg[2 * i + 1] = v.y;
                    v = __half22float2(*reinterpret_cast<const __half2 *>(&wo[i])); go[2 * i] = v.x; go[2 * i + 1] = v.y;
                }
            }
            const float ct[8] = {cur.ct[0].x, cur.ct[0].y, cur.ct[0].z, cur.ct[0].w, cur.ct[1].x, cur.ct[1].y, cur.ct[1].z, cur.ct[1].w};
            const float cp[8] = {cur.cp[0].x, cur.cp[0].y, cur.cp[0].z, cur.cp[0].w, cur.cp[1].x, cur.cp[1].y, cur.cp[1].z, cur.cp[1].w};
            float dh[8] = {cur.du[0].x, cur.du[0].y, cur.du[0].z, cur.du[0].w, cur.du[1].x, cur.du[1].y, cur.du[1].z, cur.du[1].w};
            if (kDrop) {
                const uint32_t e0 = p.drop_base + (uint32_t)(((long long)t * p.Bpad + brow) * kR2H + u);
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const uint32_t hh = r2_drop_hash(seed, (e0 + j) >> 1);
                    dh[j] = (hh & 0xFFFFu) >= thr16 ? dh[j] * keep_scale : 0.0f;
                    dh[j + 1] = (hh >> 16) >= thr16 ? dh[j + 1] * keep_scale : 0.0f;
                }
            }
            if (have_rec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) dh[j] += __uint_as_float(rec[j]);
            }
            float di[8], df[8], dg[8], d_o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float tc = r2_tanh(ct[j]);
                const float dct = fmaf(dh[j] * go[j], 1.0f - tc * tc, dc[c * 8 + j]);
                di[j] = dct * gg[j] * gi[j] * (1.0f - gi[j]);
                df[j] = dct * cp[j] * gf[j] * (1.0f - gf[j]);
                dg[j] = dct * gi[j] * (1.0f - gg[j] * gg[j]);
                d_o[j] = dh[j] * tc * go[j] * (1.0f - go[j]);
                dc[c * 8 + j] = dct * gf[j];
            }
            // this thread's 64 bytes of the chunk: k-block c*4 + uh*2 + cq/2, 16-byte slots (cq%2)*4 + gate of its row
            unsigned char *a_row = a_sm + (c * 4 + uh * 2 + (cq >> 1)) * 8192 + row_l * 128;
            const int s0 = (cq & 1) * 4;
            *reinterpret_cast<uint4 *>(a_row + (((s0 + 0) ^ (row_l & 7)) << 4)) =
                make_uint4(pack_bf16x2(di[0], di[1]), pack_bf16x2(di[2], di[3]), pack_bf16x2(di[4], di[5]), pack_bf16x2(di[6], di[7]));
            *reinterpret_cast<uint4 *>(a_row + (((s0 + 1) ^ (row_l & 7)) << 4)) =
                make_uint4(pack_bf16x2(df[0], df[1]), pack_bf16x2(df[2], df[3]), pack_bf16x2(df[4], df[5]), pack_bf16x2(df[6], df[7]));
            *reinterpret_cast<uint4 *>(a_row + (((s0 + 2) ^ (row_l & 7)) << 4)) =
                make_uint4(pack_bf16x2(dg[0], dg[1]), pack_bf16x2(dg[2], dg[3]), pack_bf16x2(dg[4], dg[5]), pack_bf16x2(dg[6], dg[7]));
            *reinterpret_cast<uint4 *>(a_row + (((s0 + 3) ^ (row_l & 7)) << 4)) =
                make_uint4(pack_bf16x2(d_o[0], d_o[1]), pack_bf16x2(d_o[2], d_o[3]), pack_bf16x2(d_o[4], d_o[5]), pack_bf16x2(d_o[6], d_o[7]));
            fence_proxy_async_smem();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                bar_arrive(&chunk_local[c]);
                if (leader) bar_arrive(&chunk_ready[c]);
                else bar_arrive_cluster(mapa_rank(&chunk_ready[c], 0));
            }
            if (threadIdx.x == 0) {
                const uint32_t par = (uint32_t)(step & 1);
                bar_wait(&chunk_local[c], par);
                // dG_t[:, columns' c*256 .. +256] of this CTA's 64 windows -> HBM (operand of the weight-gradient GEMM)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) tma_store_3d(&tmap_dg, a_sm + (c * 4 + kb) * 8192, (c * 4 + kb) * 64, m0, t);
                bulk_commit();
                if (leader) {
                    if (step == 0 && c == 0) bar_wait(w_full, 0);
                    bar_wait(&chunk_ready[c], par);
                    tcgen05_fence_after();
                    const uint32_t acc = tmem_base + (uint32_t)((step & 1) * 128);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = make_smem_desc(aa + (c * 4 + kb) * 8192 + ks * 32, 16, 1024);
                            const uint64_t db = make_smem_desc(ba + (c * 4 + kb) * S::kBTile + ks * 32, 16, 1024);
                            umma_bf16_pair(acc, da, db, idesc, (c | kb | ks) ? 1u : 0u);
                        }
                    if (c == 1) umma_commit_pair(acc_full);
                }
                if (c == 1) {
                    bulk_wait_read<0>();          // both chunks' stores have read the tile: the next step may rewrite it
                    bar_arrive(tile_free);
                }
            }
            __syncwarp();
            cur = nxt;
        }
    }
    // the accumulator of the last step holds dX_0 (dh_{-1} is not needed)
    bar_wait(acc_full, (uint32_t)((W - 1) & 1));
    tcgen05_fence_after();
    drain_dx(0, (uint32_t)(((W - 1) & 1) * 128));
    if (threadIdx.x == 0) bulk_wait_all();

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256u) : "memory");
    }
}

// Backward operand of one layer: wt [kx + 128, 512] bf16, K-major.  Row nu = R*(N/2) + w (N/2 = kx/2 + 64): w < kx/2 -> input column
// R*kx/2 + w of W_ih (zero beyond `in`), else hidden unit R*64 + (w - kx/2) of W_hh; column k' = c*256 + uh*128 + cq*32 + gate*8 + i
// stands for gate row gate*128 + uh*64 + cq*16 + c*8 + i.  perm [512] i32 OUT: that gate row for every k' (the host un-permutes
// the weight / bias gradients with it).
__global__ void lstm_pack_weights2_bwd_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh, int in, int kx,
                                              __nv_bfloat16 *__restrict__ wt, int32_t *__restrict__ perm) {
    pdl_wait();
    const int n = kx + kR2H, nh = n / 2, xh = kx / 2;
    const int total = n * 512;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int nu = e / 512, kp = e - nu * 512;
        const int c = kp >> 8, uh = (kp >> 7) & 1, cq = (kp >> 5) & 3, gate = (kp >> 3) & 3, i = kp & 7;
        const int src = gate * kR2H + uh * 64 + cq * 16 + c * 8 + i;
        const int R = nu / nh, w = nu - R * nh;
        float v;
        if (w < xh) { const int col = R * xh + w; v = col < in ? w_ih[(long long)src * in + col] : 0.0f; }
        else v = w_hh[(long long)src * kR2H + R * 64 + (w - xh)];
        wt[e] = __float2bfloat16_rn(v);
        if (nu == 0 && perm) perm[kp] = src;
    }
}


// Weight / bias gradients of one layer out of the permuted product dWp [512 (dG column order), kx + 128] = dG^T [x | h_prev] and
// dbp [512] = column sums of dG: un-permute the gate rows and split the columns into nn.LSTM's four parameter gradients.
__global__ void lstm_unpack_grads2_kernel(const float *__restrict__ dwp, const float *__restrict__ dbp, int in, int kx,
                                          float *__restrict__ dw_ih, float *__restrict__ dw_hh, float *__restrict__ db_ih,
                                          float *__restrict__ db_hh) {
    pdl_wait();
    const int kc = kx + kR2H;
    const int total = 512 * kc;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kp = e / kc, col = e - kp * kc;
        const int c = kp >> 8, uh = (kp >> 7) & 1, cq = (kp >> 5) & 3, gate = (kp >> 3) & 3, i = kp & 7;
        const int orig = gate * kR2H + uh * 64 + cq * 16 + c * 8 + i;
        const float v = dwp[e];
        if (col < in) dw_ih[(long long)orig * in + col] = v;
        else if (col >= kx) dw_hh[(long long)orig * kR2H + (col - kx)] = v;
        if (col == 0) { const float b = dbp[kp]; db_ih[orig] = b; db_hh[orig] = b; }
    }
}

}  // namespace b200med

using namespace b200med;

// 3-D map over a time-major operand buffer A [W, Bpad, ld] bf16: box {64 columns, 64 rows, 1 step}
static int make_tmap_a64(CUtensorMap *tm, const void *A, int64_t W, int64_t Bpad, int64_t ld) {
    const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)Bpad, (cuuint64_t)W};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)Bpad * ld * 2};
    const cuuint32_t box[3] = {64, 64, 1};
    return make_tmap_nd(tm, A, 3, dims, strides, box);
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_pack_weights2(const float *w_ih, const float *w_hh, const float *b_ih,
                                                                                const float *b_hh, int32_t in, int32_t kx, void *wp,
                                                                                float *bias_p, void *stream) {
    B200MED_REQUIRE(in >= 1 && (kx == 64 || kx == 128) && in <= kx, "kx must be 64 or 128 and >= the input width");
    B200MED_REQUIRE(w_ih && w_hh && b_ih && b_hh && wp && bias_p, "null pointer");
    const int total = 4 * kR2H * (kx + kR2H);
    launch_k(lstm_pack_weights2_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, w_ih, w_hh, b_ih, b_hh, in, kx,
                                                                                    reinterpret_cast<__nv_bfloat16 *>(wp), bias_p);
    return after_launch("lstm_pack_weights2_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_rec2_fwd(
    void *a_l, int32_t ld_l, int32_t kx, const void *wp, const float *bias_p, void *gact, float *c, void *a_up, int32_t ld_up,
    float *h_out, int64_t B, int64_t Bpad, int32_t W, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(B >= 1 && W >= 1 && Bpad >= B && Bpad % 32 == 0 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE((kx == 64 || kx == 128) && ld_l >= kx + kR2H && ld_l % 8 == 0, "A_l holds [x (kx = 64 | 128 columns) | h (128)]");
    B200MED_REQUIRE(a_l && wp && bias_p && ((uintptr_t)a_l % 16 == 0) && ((uintptr_t)wp % 16 == 0), "null / unaligned pointer");
    B200MED_REQUIRE(!a_up || (ld_up % 8 == 0 && ld_up >= kR2H && (uintptr_t)a_up % 16 == 0), "bad a_up");
    const bool save = gact != nullptr, up = a_up != nullptr, drp = up && drop_p > 0.0f;
    B200MED_REQUIRE(save == (c != nullptr), "gact and c are saved together (training) or not at all");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    CUtensorMap tw, ta, tu;
    if (int e = make_tmap(&tw, wp, kx + kR2H, 4 * kR2H, kx + kR2H, 64, 256)) return e;      // box {64 k, 256 gate rows}
    if (int e = make_tmap_a64(&ta, a_l, W, Bpad, ld_l)) return e;
    tu = ta;
    if (a_up) if (int e = make_tmap_a64(&tu, a_up, W, Bpad, ld_up)) return e;
    Rec2FwdParams p{};
    p.bias = bias_p; p.gact = reinterpret_cast<__half *>(gact); p.c = c; p.h_out = h_out;
    p.B = B; p.Bpad = Bpad; p.W = W; p.kx = kx;
    p.drop_p = up ? drop_p : 0.0f; p.seed = seed; p.drop_base = (uint32_t)drop_base;
    const unsigned clusters = (unsigned)((Bpad + 2 * kR2Rows - 1) / (2 * kR2Rows));
    auto launch = [&](auto kern, size_t smem) -> int {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(lstm_rec2_fwd)")) return e;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kR2Threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1);
        return check_cuda(cudaLaunchKernelEx(&cfg, kern, tw, ta, tu, p), "cudaLaunchKernelEx(lstm_rec2_fwd)");
    };
    int e;
#define B200MED_R2(KXB)                                                                                                      \
    (save ? (up ? (drp ? launch(lstm_rec2_fwd_kernel<KXB, true, true, true>, Rec2Smem<KXB>::kBytes)                            \
                       : launch(lstm_rec2_fwd_kernel<KXB, true, true, false>, Rec2Smem<KXB>::kBytes))                          \
                : launch(lstm_rec2_fwd_kernel<KXB, true, false, false>, Rec2Smem<KXB>::kBytes))                                \
          : (up ? (drp ? launch(lstm_rec2_fwd_kernel<KXB, false, true, true>, Rec2Smem<KXB>::kBytes)                           \
                       : launch(lstm_rec2_fwd_kernel<KXB, false, true, false>, Rec2Smem<KXB>::kBytes))                         \
                : launch(lstm_rec2_fwd_kernel<KXB, false, false, false>, Rec2Smem<KXB>::kBytes)))
    e = kx == 64 ? B200MED_R2(1) : B200MED_R2(2);
#undef B200MED_R2
    if (e) return e;
    return after_launch("lstm_rec2_fwd_kernel");
}


extern "C" __attribute__((visibility("default"))) int b200med_lstm_pack_weights2_bwd(const float *w_ih, const float *w_hh, int32_t in, int32_t kx,
                                                                                    void *wt, int32_t *perm, void *stream) {
    B200MED_REQUIRE(in >= 1 && (kx == 64 || kx == 128) && in <= kx, "kx must be 64 or 128 and >= the input width");
    B200MED_REQUIRE(w_ih && w_hh && wt, "null pointer");
    const int total = (kx + kR2H) * 512;
    launch_k(lstm_pack_weights2_bwd_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, w_ih, w_hh, in, kx,
                                                                                        reinterpret_cast<__nv_bfloat16 *>(wt), perm);
    return after_launch("lstm_pack_weights2_bwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_rec2_bwd(
    const void *gact, const float *c, const void *wt, int32_t kx, const float *dh_top, const float *dh_up, void *dG, float *dx,
    int64_t B, int64_t Bpad, int32_t W, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream) {
    B200MED_REQUIRE(B >= 1 && W >= 1 && Bpad >= B && Bpad % 32 == 0 && drop_p >= 0.0f && drop_p < 1.0f, "bad shape");
    B200MED_REQUIRE(kx == 64 || kx == 128, "kx must be 64 or 128");
    B200MED_REQUIRE(gact && c && wt && dG && dx && ((uintptr_t)wt % 16 == 0) && ((uintptr_t)dG % 16 == 0) && ((uintptr_t)dx % 16 == 0),
                    "null / unaligned pointer");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    CUtensorMap tw, tg;
    if (int e = make_tmap(&tw, wt, 512, kx + kR2H, 512, 64, (kx + kR2H) / 2)) return e;       // box {64 k, N/2 output columns}
    if (int e = make_tmap_a64(&tg, dG, W, Bpad, 4 * kR2H)) return e;                            // dG [W, Bpad, 512], box {64, 64, 1}
    Rec2BwdParams p{};
    p.gact = reinterpret_cast<const __half *>(gact); p.c = c; p.dh_top = dh_top; p.dh_up = dh_up; p.dx = dx;
    p.B = B; p.Bpad = Bpad; p.W = W; p.drop_p = dh_up ? drop_p : 0.0f; p.seed = seed; p.drop_base = (uint32_t)drop_base;
    const unsigned clusters = (unsigned)((Bpad + 2 * kR2Rows - 1) / (2 * kR2Rows));
    auto launch = [&](auto kern, size_t smem) -> int {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(lstm_rec2_bwd)")) return e;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kR2BwdThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1);
        return check_cuda(cudaLaunchKernelEx(&cfg, kern, tw, tg, p), "cudaLaunchKernelEx(lstm_rec2_bwd)");
    };
    const bool drp = dh_up && drop_p > 0.0f;
    int e;
    if (kx == 64) e = drp ? launch(lstm_rec2_bwd_kernel<1, true>, Rec2BwdSmem<1>::kBytes) : launch(lstm_rec2_bwd_kernel<1, false>, Rec2BwdSmem<1>::kBytes);
    else e = drp ? launch(lstm_rec2_bwd_kernel<2, true>, Rec2BwdSmem<2>::kBytes) : launch(lstm_rec2_bwd_kernel<2, false>, Rec2BwdSmem<2>::kBytes);
    if (e) return e;
    return after_launch("lstm_rec2_bwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_lstm_unpack_grads2(const float *dwp, const float *dbp, int32_t in, int32_t kx,
                                                                                float *dw_ih, float *dw_hh, float *db_ih, float *db_hh,
                                                                                void *stream) {
    B200MED_REQUIRE(in >= 1 && (kx == 64 || kx == 128) && in <= kx, "kx must be 64 or 128 and >= the input width");
    B200MED_REQUIRE(dwp && dbp && dw_ih && dw_hh && db_ih && db_hh, "null pointer");
    const int total = 512 * (kx + kR2H);
    launch_k(lstm_unpack_grads2_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, dwp, dbp, in, kx, dw_ih, dw_hh, db_ih, db_hh);
    return after_launch("lstm_unpack_grads2_kernel");
}
