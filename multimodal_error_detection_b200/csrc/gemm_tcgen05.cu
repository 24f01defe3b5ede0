// K2 (throughput mode): bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// tensor memory, operands staged by TMA), for the FeatureExtractor MLP forward, data-gradient and
// weight-gradient products (MED/modeling/models.py:19-35 and its autograd backward).
//
//   D[M,N] = A[M,K] * B[N,K]^T      A, B bf16; fp32 accumulation in TMEM; D bf16 or fp32
//
// Either operand may be K-major (stored [rows, K]) or MN-major (stored [K, rows]); the weight
// gradient dW = dY^T X reduces over the ROW index of both activations, i.e. both are MN-major.
//
// Kernel anatomy (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) into a kStages ring,
//               completion by mbarrier complete_tx.
//   warp 1      allocates TMEM (2 accumulators x BLOCK_N columns), then one elected lane issues
//               tcgen05.mma.cta_group::1.kind::f16 (128 x BLOCK_N x 16) and tcgen05.commit to the
//               ring's "empty" barriers / the accumulator's "full" barrier.
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / ReLU / ReLU-mask -> 128-bit
//               global stores; releases the accumulator so the MMA warp can start the next tile while
//               this one is drained (double-buffered TMEM).
// Roofline: tensor pipe.  128x256x16 per instruction = 128 cycles at cta_group::1, operand traffic
// 12 KB per instruction = 96 B/cycle of shared-memory bandwidth (below the 128 B/cycle port), which
// is why BLOCK_N = 256 is the default tile for the wide layers.
#include "tcgen05.cuh"
#include <stdlib.h>

namespace b200med {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr uint32_t kATileBytes = BLOCK_M * BLOCK_K * 2;  // 16 KB

struct GemmParams {
    long long M, N, K;
    long long ldd;          // elements between output rows (of D, or of a split-K partial slab = N)
    int a_kmajor, b_kmajor;
    int out_f32;            // 1: fp32 output, 0: bf16
    int relu;
    int split_k;            // >= 1
    int kb_per_split;       // k-blocks per split
    int tiles_m, tiles_n;
    int out_f16;            // 1: fp16 output (16-bit paths write half instead of bfloat16); needs split_k == 1
    int out_rbi;            // 1: row-block-interleaved output [M/32][N/V][32][V] (split_k == 1, no mask)
    const float *bias;      // [N] or null (ignored when split_k > 1)
    const __nv_bfloat16 *mask;  // [M, ldd] or null (ignored when split_k > 1)
    void *D;                // output, or the fp32 partial workspace when split_k > 1
};

// kStage: the epilogue stages the bf16 output tile in shared memory (and fetches the ReLU-mask tile into the same buffer)
// kDouble (staged, single CTA, short K loops): TWO staging buffers and a 2-deep operand ring.  With one buffer a tile's mask
// load could only start after the previous tile's TMA store had read the buffer, so load latency + store read sat between
// the drains of consecutive tiles: the ReLU-masked data gradients (K = 32 / 256: one to four k-blocks per 128 x 256 tile) ran
// at 6.8 us per tile against 4.4 us of HBM time.  Now the mask of tile i+1 lands while tile i drains.
template <int BLOCK_N, bool kPair = false, bool kStage = false, bool kDouble = false>
struct GemmCfg {
    static_assert(!kDouble || (kStage && !kPair), "double staging: staged single-CTA kernel only");
    // CTA pair: each CTA stages its own 128 rows of A and HALF of the B tile's rows
    static constexpr int kBRows = kPair ? BLOCK_N / 2 : BLOCK_N;
    static constexpr uint32_t kBTileBytes = kBRows * BLOCK_K * 2;
    static constexpr uint32_t kStageBytes = kATileBytes + kBTileBytes;
    static constexpr uint32_t kStagingOne = kStage ? BLOCK_M * BLOCK_N * 2 : 0;       // 64-column boxes of [128 rows][128 B]
    static constexpr uint32_t kStagingBytes = (kDouble ? 2 : 1) * kStagingOne;
    static constexpr int kStages = kDouble ? 2 : kStage ? (kPair ? 5 : ((BLOCK_N >= 256) ? 3 : 5))
                                          : (kPair ? 6 : ((BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8)));
    static constexpr int kAccStages = 2;
    static constexpr uint32_t kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;  // power of two: BLOCK_N in {32,64,128,256}
    static constexpr size_t kBaseBytes = (size_t)kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    // bias in shared memory: the whole vector (N <= 512, loaded once) where 2 KB fit, else the current column tile's slice
    static constexpr int kBiasFloats = (kBaseBytes + 2048 <= 232448) ? 512 : 256;
    static constexpr size_t kSmemBytes = kBaseBytes + kBiasFloats * 4;
};

// EPI selects the epilogue at compile time (smaller code, no mode branches in the drain loop):
//   0 row-major output with bias / ReLU / ReLU-mask, 1 row-block-interleaved output with bias / ReLU, 2 split-K fp32 partial.
//   3 row-major bf16 output staged in shared memory and written by TMA store (bias / ReLU; the ReLU-mask tile is TMA-loaded
//     into the same staging buffer): a warp's 16-byte accesses hit 8 conflict-free shared-memory slots instead of 32 lines.
constexpr int kEpiRowMajor = 0, kEpiRbi = 1, kEpiPartial = 2, kEpiTmaBf16 = 3;

// kPair: the kernel is launched as clusters of 2 CTAs; a pair owns a 256 x BLOCK_N output tile (cta_group::2 MMA, M = 256):
// CTA r stages rows [128 r, 128 r + 128) of the A tile and rows [BLOCK_N/2 r, +BLOCK_N/2) of the B tile, so every byte of B
// is fetched from L2 once per 256 output rows instead of once per 128 -- the 1-CTA kernel needs 96 B/cycle/SM of L2 -> SMEM
// traffic at full MMA rate, the pair 64 B/cycle/SM.
template <int BLOCK_N, int EPI, bool kPair, bool kDouble = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_mask,
                         const GemmParams p) {
    pdl_wait();
    using Cfg = GemmCfg<BLOCK_N, kPair, EPI == kEpiTmaBf16, kDouble>;
    constexpr int kTileM = kPair ? 2 * BLOCK_M : BLOCK_M;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const long long cta_id = kPair ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
    const long long cta_stride = kPair ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment: required by the 128B swizzle pattern shared by TMA and the UMMA descriptors
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *staging = smem + (size_t)Cfg::kStages * Cfg::kStageBytes;      // 1024-byte aligned (stages are multiples of 1 KB)
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(staging + Cfg::kStagingBytes);
    uint64_t *empty_bar = full_bar + Cfg::kStages;
    uint64_t *acc_full = empty_bar + Cfg::kStages;
    uint64_t *acc_empty = acc_full + Cfg::kAccStages;
    uint64_t *mask_bar = acc_empty + Cfg::kAccStages;        // [2] TMA-store epilogue: "the mask tile is in staging buffer b"
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mask_bar + 2);
    float *bias_sm = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(full_bar) + 256);   // [256]: bias of the tile's columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb_total = (int)((p.K + BLOCK_K - 1) / BLOCK_K);
    const long long tiles_mn = (long long)p.tiles_m * p.tiles_n;
    const long long total_tiles = tiles_mn * p.split_k;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
        // pair: the leader's full barrier collects its own arrive.expect_tx and the peer's arrive; its acc_empty barrier
        // collects the 8 epilogue warps of both CTAs
        for (int s = 0; s < Cfg::kStages; ++s) { bar_init(&full_bar[s], kPair ? 2 : 1); bar_init(&empty_bar[s], 1); }
        for (int s = 0; s < Cfg::kAccStages; ++s) { bar_init(&acc_full[s], 1); bar_init(&acc_empty[s], kPair ? 16 : 8); }
        bar_init(&mask_bar[0], 1); bar_init(&mask_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(s_addr(tmem_slot)), "r"(Cfg::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(s_addr(tmem_slot)), "r"(Cfg::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();      // both CTAs' barriers are initialised before any remote arrive / TMA signal
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = cta_id; tile < total_tiles; tile += cta_stride) {
                const int z = (int)(tile / tiles_mn);
                const long long mn = tile % tiles_mn;
                const int m0 = (int)(mn / p.tiles_n) * kTileM + (int)rank * BLOCK_M;
                const int n0 = (int)(mn % p.tiles_n) * BLOCK_N + (int)rank * (kPair ? Cfg::kBRows : 0);
                const int kb0 = z * p.kb_per_split;
                const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    bar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char *sa = smem + (size_t)stage * Cfg::kStageBytes;
                    unsigned char *sb = sa + kATileBytes;
                    const int k0 = kb * BLOCK_K;
                    if constexpr (kPair) {
                        const uint32_t lbar = mapa_rank(&full_bar[stage], 0);         // the leader's barrier for this stage
                        if (leader) bar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        if (p.a_kmajor) {
                            tma_load_2d_pair(sa, &tmap_a, lbar, k0, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tmap_a, lbar, m0 + 64 * j, k0);
                        }
                        if (p.b_kmajor) {
                            tma_load_2d_pair(sb, &tmap_b, lbar, k0, n0);              // box {64 k, BLOCK_N / 2 n}
                        } else {
#pragma unroll
                            for (int j = 0; j < Cfg::kBRows / 64; ++j) tma_load_2d_pair(sb + j * 8192, &tmap_b, lbar, n0 + 64 * j, k0);
                        }
                        if (!leader) bar_arrive_cluster(lbar);
                    } else {
                        bar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        if (p.a_kmajor) {
                            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);           // box {64 k, 128 m}
                        } else {
#pragma unroll
                            for (int j = 0; j < BLOCK_M / 64; ++j)                          // boxes {64 m, 64 k}
                                tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, k0);
                        }
                        if (p.b_kmajor) {
                            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);           // box {64 k, BLOCK_N n}
                        } else {
#pragma unroll
                            for (int j = 0; j < (BLOCK_N >= 64 ? BLOCK_N / 64 : 1); ++j)
                                tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, k0);
                        }
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc(kTileM, BLOCK_N, !p.a_kmajor, !p.b_kmajor);
            // K-major, SW128: 8-row groups are 1024 B apart (SBO); LBO unused (1).  One UMMA_K step = 32 B.
            // MN-major, SW128: 64-element MN atoms are 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
            // one UMMA_K step = 16 k-rows = 2048 B.
            const uint32_t a_lbo = p.a_kmajor ? 16 : 8192, b_lbo = p.b_kmajor ? 16 : 8192;
            const uint32_t a_kstep = p.a_kmajor ? 32 : 2048, b_kstep = p.b_kmajor ? 32 : 2048;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = cta_id; tile < total_tiles; tile += cta_stride) {
                const int z = (int)(tile / tiles_mn);
                const int kb0 = z * p.kb_per_split;
                const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
                bar_wait(&acc_empty[acc], acc_phase ^ 1);  // epilogue (of both CTAs of a pair) has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = kb0; kb < kb1; ++kb) {
                    bar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = s_addr(smem + (size_t)stage * Cfg::kStageBytes);
                    const uint32_t sb = sa + kATileBytes;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
                        const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
                        if constexpr (kPair) umma_bf16_pair(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs have read it
                    if constexpr (kPair) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue
                if constexpr (kPair) umma_commit_pair(&acc_full[acc]); else umma_commit(&acc_full[acc]);
                if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        // Two warps per TMEM lane quarter, each draining half of the tile's columns in 32-column chunks; the tcgen05.ld
        // of chunk i+1 is in flight while chunk i goes through bias / ReLU / mask / conversion and its stores.
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may read
        const int half = (warp - 2) >> 2;
        constexpr int kChunks = BLOCK_N >= 64 ? BLOCK_N / 64 : 1;       // chunks per warp
        const int c_first = BLOCK_N >= 64 ? half * (BLOCK_N / 2) : 0;
        const bool has_work = BLOCK_N >= 64 || half == 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr bool staged = EPI == kEpiTmaBf16;
        constexpr bool partial_k = EPI == kEpiPartial;
        constexpr int kBoxes = BLOCK_N / 64;                      // TMA boxes of 64 columns x 128 rows per tile
        const bool elect = warp == 2 && lane == 0;                // issues the TMA stores / mask loads of this CTA
        const bool use_mask = staged && p.mask != nullptr;
        uint32_t mask_phase = 0;                                  // kDouble: bit b = phase of buffer b's barrier
        unsigned char *const staging0 = staging;
        int sbuf = 0;                                             // kDouble: buffer of the current tile
        auto load_mask_tile = [&](long long t, int b) {
            const long long mn_ = t % tiles_mn;
            const int m_ = (int)(mn_ / p.tiles_n) * kTileM + (int)rank * BLOCK_M, n_ = (int)(mn_ % p.tiles_n) * BLOCK_N;
            unsigned char *dst = staging0 + (size_t)b * Cfg::kStagingOne;
            bar_expect_tx(&mask_bar[b], Cfg::kStagingOne);
#pragma unroll
            for (int bx = 0; bx < kBoxes; ++bx) tma_load_2d(dst + bx * 16384, &tmap_mask, &mask_bar[b], n_ + 64 * bx, m_);
        };
        if (use_mask && elect && cta_id < total_tiles) load_mask_tile(cta_id, 0);
        // N <= 512 (every GEMM of the train step that has a bias): the whole bias vector is loaded once per CTA
        const bool bias_all = !partial_k && p.bias && p.N <= Cfg::kBiasFloats;
        if (bias_all) {
            for (int e = threadIdx.x - 64; e < Cfg::kBiasFloats; e += 256) bias_sm[e] = e < p.N ? __ldg(p.bias + e) : 0.0f;
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        for (long long tile = cta_id; tile < total_tiles; tile += cta_stride) {
            const int z = (int)(tile / tiles_mn);
            const long long mn = tile % tiles_mn;
            const long long m0 = (mn / p.tiles_n) * kTileM + rank * BLOCK_M, n0 = (mn % p.tiles_n) * BLOCK_N;
            if (!partial_k && p.bias && !bias_all) {
                // the tile's bias slice goes through shared memory once (LDS broadcast in the drain loop instead of eight
                // dependent global loads per 32-column chunk, which were the top stall of the small-K GEMMs)
                asm volatile("bar.sync 2, 256;" ::: "memory");        // everyone is done with the previous tile's slice
                const int e_tid = threadIdx.x - 64;
                if (e_tid < BLOCK_N) bias_sm[e_tid] = n0 + e_tid < p.N ? __ldg(p.bias + n0 + e_tid) : 0.0f;
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            const float *bias_tile = bias_all ? bias_sm + n0 : bias_sm;
            bar_wait(&acc_full[acc], acc_phase);
            tcgen05_fence_after();
            if constexpr (staged && kDouble) {
                staging = staging0 + (size_t)sbuf * Cfg::kStagingOne;
                if (use_mask) {
                    if (elect && tile + cta_stride < total_tiles) {
                        bulk_wait_read<0>();                      // the previous tile's store has read the OTHER buffer ...
                        load_mask_tile(tile + cta_stride, sbuf ^ 1);      // ... so the next tile's mask lands there under this drain
                    }
                    bar_wait(&mask_bar[sbuf], (mask_phase >> sbuf) & 1u);
                    mask_phase ^= 1u << sbuf;
                } else {
                    if (elect) bulk_wait_read<1>();               // the store of two tiles ago has read this buffer
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            } else if constexpr (staged) {
                if (use_mask) {
                    bar_wait(&mask_bar[0], mask_phase);           // the mask tile has landed (and the previous store has been read)
                    mask_phase ^= 1;
                } else {
                    if (elect) bulk_wait_read<0>();               // the previous tile's TMA store has finished reading the buffer
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
            const long long m = m0 + quarter * 32 + lane;
            const bool row_ok = m < p.M;
            constexpr bool partial = EPI == kEpiPartial;
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + c_first);
            uint32_t vbuf[2][32];
            if (has_work) {
                tmem_ld32_nowait(t_addr, vbuf[0]);
                tmem_wait_ld();
            }
#pragma unroll
            for (int ci = 0; ci < kChunks; ++ci) {
                if (!has_work) break;
                const int c0 = c_first + ci * 32;
                uint32_t (&v)[32] = vbuf[ci & 1];
                if (ci + 1 < kChunks) tmem_ld32_nowait(t_addr + (uint32_t)((ci + 1) * 32), vbuf[(ci + 1) & 1]);
                const long long n_base = n0 + c0;
                if (row_ok && n_base < p.N) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    const bool full = n_base + 32 <= p.N;
                    if (!partial) {
                        if (p.bias) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 b4 = *reinterpret_cast<const float4 *>(bias_tile + c0 + 4 * q);
                                f[4 * q] += b4.x; f[4 * q + 1] += b4.y; f[4 * q + 2] += b4.z; f[4 * q + 3] += b4.w;
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
                        }
                        if constexpr (staged) {
                            // this lane's 64 bytes of the tile row: four 16-byte slots of box c0/64, 128B swizzle
                            unsigned char *srow = staging + (c0 >> 6) * 16384 + (quarter * 32 + lane) * 128;
                            const int slot0 = (c0 & 63) >> 3, sw = (quarter * 32 + lane) & 7;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 *sp = reinterpret_cast<uint4 *>(srow + (((slot0 + q) ^ sw) << 4));
                                if (use_mask) {
                                    const uint4 u = *sp;
                                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {
                                        const uint32_t lo = w[h] & 0xFFFFu, hi = w[h] >> 16;
                                        if (!((lo & 0x8000u) == 0 && (lo & 0x7FFFu) != 0)) f[q * 8 + h * 2] = 0.0f;
                                        if (!((hi & 0x8000u) == 0 && (hi & 0x7FFFu) != 0)) f[q * 8 + h * 2 + 1] = 0.0f;
                                    }
                                }
                                *sp = make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                                 pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
                            }
                        }
                        if (EPI == kEpiRowMajor && p.mask) {
                            const __nv_bfloat16 *mk = p.mask + m * p.ldd + n_base;
                            if (full && (((uintptr_t)mk) % 16 == 0)) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(mk) + q);
                                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {
                                        // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                                        const uint32_t lo = w[h] & 0xFFFFu, hi = w[h] >> 16;
                                        if (!((lo & 0x8000u) == 0 && (lo & 0x7FFFu) != 0)) f[q * 8 + h * 2] = 0.0f;
                                        if (!((hi & 0x8000u) == 0 && (hi & 0x7FFFu) != 0)) f[q * 8 + h * 2 + 1] = 0.0f;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n_base + j < p.N && !(__bfloat162float(mk[j]) > 0.0f)) f[j] = 0.0f;
                            }
                        }
                    }
                    if constexpr (staged) {
                        // output already sits in the staging tile
                    } else if (EPI == kEpiRbi) {
                        // one 16-byte vector per lane and column group: a warp writes 512 contiguous bytes
                        const long long rb = m >> 5;
                        const int rl = (int)(m & 31);
                        if (p.out_f32) {
                            float *base = reinterpret_cast<float *>(p.D);
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                if (n_base + 4 * q < p.N)
                                    *reinterpret_cast<float4 *>(base + ((rb * (p.N >> 2) + ((n_base >> 2) + q)) * 32 + rl) * 4) =
                                        make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                        } else {
                            __nv_bfloat16 *base = reinterpret_cast<__nv_bfloat16 *>(p.D);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (n_base + 8 * q < p.N)
                                    *reinterpret_cast<uint4 *>(base + ((rb * (p.N >> 3) + ((n_base >> 3) + q)) * 32 + rl) * 8) =
                                        p.out_f16
                                            ? make_uint4(pack_f16x2(f[8 * q], f[8 * q + 1]), pack_f16x2(f[8 * q + 2], f[8 * q + 3]),
                                                         pack_f16x2(f[8 * q + 4], f[8 * q + 5]), pack_f16x2(f[8 * q + 6], f[8 * q + 7]))
                                            : make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                                         pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
                        }
                    } else if (p.out_f32 || partial) {
                        float *dst = reinterpret_cast<float *>(p.D) +
                                     (partial ? (long long)z * p.M * p.N + m * p.N : m * p.ldd) + n_base;
                        if (full && (((uintptr_t)dst) % 16 == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4 *>(dst)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n_base + j < p.N) dst[j] = f[j];
                        }
                    } else {
                        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.D) + m * p.ldd + n_base;
                        if (full && (((uintptr_t)dst) % 16 == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                reinterpret_cast<uint4 *>(dst)[q] =
                                    make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                               pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n_base + j < p.N) dst[j] = __float2bfloat16_rn(f[j]);
                        }
                    }
                }
                if (ci + 1 < kChunks) tmem_wait_ld();
            }
            // all of this warp's TMEM reads are complete (tcgen05.wait::ld above) -> release the accumulator
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kPair && !leader) bar_arrive_cluster(mapa_rank(&acc_empty[acc], 0));   // the MMA issuer lives in the leader
                else bar_arrive(&acc_empty[acc]);
            }
            if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
            if constexpr (staged) {
                fence_proxy_async_smem();                         // staged tile (generic stores) -> visible to the TMA store
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (elect) {
#pragma unroll
                    for (int bx = 0; bx < kBoxes; ++bx)
                        if (n0 + 64 * bx < p.N) tma_store_2d(&tmap_d, staging + bx * 16384, (int)n0 + 64 * bx, (int)m0);
                    bulk_commit();
                    if (!kDouble && use_mask && tile + cta_stride < total_tiles) {
                        bulk_wait_read<0>();                      // the store has read the buffer: the next mask tile may land
                        load_mask_tile(tile + cta_stride, 0);
                    }
                }
                if constexpr (kDouble) sbuf ^= 1;
            }
        }
        if constexpr (staged) {
            if (elect) bulk_wait_all();
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();      // no CTA of the pair leaves while its partner may still signal it
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
    }
}

// D = sum_z partial[z] (ascending z, fixed order) + bias, ReLU, ReLU-mask, dtype conversion.
__global__ void splitk_reduce_kernel(const float *__restrict__ part, int split, long long M, long long N, long long ldd,
                                     const float *__restrict__ bias, int relu, const __nv_bfloat16 *__restrict__ mask,
                                     void *__restrict__ D, int out_f32) {
    pdl_wait();
    const long long total = M * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long m = e / N, n = e - m * N;
        float s = 0.0f;
        for (int z = 0; z < split; ++z) s += part[(long long)z * total + e];
        if (bias) s += bias[n];
        if (relu) s = fmaxf(s, 0.0f);
        if (mask && !(__bfloat162float(mask[m * ldd + n]) > 0.0f)) s = 0.0f;
        if (out_f32) reinterpret_cast<float *>(D)[m * ldd + n] = s;
        else reinterpret_cast<__nv_bfloat16 *>(D)[m * ldd + n] = __float2bfloat16_rn(s);
    }
}

// Widths the 256-column CTA-pair tile takes: whole tiles, or a last tile at least three quarters full (N = 192: the LSTM
// layer-0 weight gradient dG^T [x | h] with 58 -> 64 + 128 columns; TMA zero-fills the missing rows of the B tile, the
// epilogues mask columns >= N) -- the 128-column single-CTA kernel ran that product at 77 us against 50 us for N = 256.
static bool pair_n_ok(long long N) { return N % 64 == 0 && (N % 256 == 0 || N % 256 >= 192); }

static int pick_block_n(long long N, int b_kmajor) {
    if (N >= 256) return 256;
    if (N > 64) return 128;
    if (N > 32 || !b_kmajor) return 64;  // MN-major B tiles are built from 64-element atoms
    return 32;
}

template <int BLOCK_N, int EPI, bool kPair = false, bool kDouble = false>
static int launch_gemm_tc(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &td, const CUtensorMap &tm,
                          const GemmParams &p, cudaStream_t st) {
    using Cfg = GemmCfg<BLOCK_N, kPair, EPI == kEpiTmaBf16, kDouble>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, EPI, kPair, kDouble>;
    static bool attr_set = false;
    if (!attr_set) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes),
                               "cudaFuncSetAttribute(gemm_bf16_tcgen05)")) return e;
        attr_set = true;
    }
    const long long tiles = (long long)p.tiles_m * p.tiles_n * p.split_k;
    if constexpr (kPair) {
        const long long pairs = num_sms() / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * (tiles < pairs ? tiles : pairs)));
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1);      // (the cluster scheduling policy makes no difference here, measured)
        if (int e = check_cuda(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, tm, p), "cudaLaunchKernelEx(gemm_bf16_tcgen05 pair)")) return e;
        return after_launch("gemm_bf16_tcgen05_kernel(pair)");
    } else {
        const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
        launch_k(kern, grid, kGemmThreads, Cfg::kSmemBytes, st, ta, tb, td, tm, p);
        return after_launch("gemm_bf16_tcgen05_kernel");
    }
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int b200med_has_tcgen05(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

static bool gemm_pair_allowed() {
    static const bool v = []() { const char *e = getenv("B200MED_GEMM_PAIR"); return !(e && e[0] == '0'); }();
    return v;
}

// Split-K choice for the weight-gradient GEMMs (small M x N, K = all rows of the batch): the persistent kernel deals
// tiles * split work items over the SMs (or SM pairs), so the cost is WAVES x items' length, not the item count -- the first
// heuristic ("2 items per SM") gave the layer-1 gradient 160 items on 74 pairs = 3 waves at 72 % (245 us; 9 splits = 2 full
// waves).  Cost model in units of one 128 x 256 x 64 k-block per SM: waves * (k-blocks per split + 6 for the partial-tile
// epilogue) + the fixed-order reduction of the partials (~3.9 per split and 2^20 output elements, measured).
extern "C" __attribute__((visibility("default"))) int32_t b200med_gemm_bf16_pick_split(int64_t M, int64_t N, int64_t K,
                                                                                       int32_t b_kmajor) {
    if (M < 1 || N < 1 || K < 1) return 1;
    int block_n = pick_block_n(N, b_kmajor);
    const int nkb = (int)((K + BLOCK_K - 1) / BLOCK_K);
    if (nkb < 16) return 1;
    if (N == 192 && M >= 256 && gemm_pair_allowed()) block_n = 256;      // three quarters of a pair tile (see pair_n_ok)
    const int sms = num_sms();
    double best = 1e300;
    int best_s = 1;
    for (int s = 1; s <= 160 && nkb / s >= 8; ++s) {
        const int kb_per = (nkb + s - 1) / s;
        if ((nkb + kb_per - 1) / kb_per != s) continue;
        const bool pair = gemm_pair_allowed() && block_n == 256 && M >= 256 && pair_n_ok(N) && nkb / s >= 8;
        const long long tiles = pair ? ((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((N + 255) / 256)
                                     : ((M + BLOCK_M - 1) / BLOCK_M) * ((N + block_n - 1) / block_n);
        const long long units = pair ? sms / 2 : sms;
        const long long waves = (tiles * s + units - 1) / units;
        const double cost = (double)waves * (kb_per + 6.0) * (block_n / 256.0) +
                            (s > 1 ? 3.9 * s * (double)M * (double)N / 1048576.0 : 0.0);
        if (cost < best - 1e-9) { best = cost; best_s = s; }
    }
    return best_s;
}

extern "C" __attribute__((visibility("default"))) int64_t b200med_gemm_bf16_ws_bytes(int64_t M, int64_t N, int64_t K, int32_t split_k) {
    (void)K;
    return split_k > 1 ? (int64_t)split_k * M * N * 4 : 0;
}

extern "C" __attribute__((visibility("default"))) int b200med_gemm_bf16(const void *A, const void *B, void *D, const float *bias, const void *mask,
                                 int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd,
                                 int32_t a_kmajor, int32_t b_kmajor, int32_t out_dtype, int32_t relu,
                                 int32_t split_k, int32_t out_layout, void *workspace, void *stream) {
    B200MED_REQUIRE(M >= 1 && N >= 1 && K >= 1, "bad shape");
    B200MED_REQUIRE(A && B && D, "null pointer");
    B200MED_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "operand leading dimensions must be multiples of 8 elements (16 bytes)");
    B200MED_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "operands must be 16-byte aligned");
    B200MED_REQUIRE(out_dtype == B200MED_BF16 || out_dtype == B200MED_F32 || out_dtype == B200MED_F16, "bad out dtype");
    B200MED_REQUIRE(out_dtype != B200MED_F16 || out_layout == B200MED_LAYOUT_RBI32, "fp16 output is only implemented for the RBI32 layout");
    B200MED_REQUIRE(out_layout == B200MED_LAYOUT_ROWMAJOR || out_layout == B200MED_LAYOUT_RBI32, "bad out_layout");
    B200MED_REQUIRE(out_layout == B200MED_LAYOUT_RBI32 || ldd >= N, "ldd < N");
    B200MED_REQUIRE(out_layout == B200MED_LAYOUT_ROWMAJOR || (split_k <= 1 && !mask && N % (out_dtype == B200MED_F32 ? 4 : 8) == 0),
                    "RBI32 output needs split_k = 1, no mask and N a multiple of the 16-byte vector width");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;

    int block_n = pick_block_n(N, b_kmajor);
    const int nkb = (int)((K + BLOCK_K - 1) / BLOCK_K);
    if (split_k < 1) split_k = 1;
    if (split_k > nkb) split_k = nkb;
    int kb_per = (nkb + split_k - 1) / split_k;
    split_k = (nkb + kb_per - 1) / kb_per;
    if (N == 192 && M >= 256 && gemm_pair_allowed() && nkb / split_k >= 8) block_n = 256;    // three quarters of a pair tile
    B200MED_REQUIRE(split_k == 1 || workspace, "split_k > 1 needs a workspace");

    // CTA-pair mode (cta_group::2, 256 x 256 tiles): the wide GEMMs whose K loop is long enough to be bound by the
    // L2 -> shared-memory operand traffic of the 1-CTA kernel (FE layer 1 forward 231 -> 210 us, weight gradient 260 -> 238 us,
    // measured A/B on one box).  B200MED_GEMM_PAIR=0 switches it off.
    const bool pair = gemm_pair_allowed() && block_n == 256 && M >= 256 && pair_n_ok(N) && nkb / split_k >= 8;

    CUtensorMap ta, tb;
    // K-major operand: rows x K, K contiguous -> box {64 k, rows}.  MN-major: K x rows -> box {64 rows, 64 k}.
    if (int e = a_kmajor ? make_tmap(&ta, A, K, M, lda, BLOCK_K, BLOCK_M) : make_tmap(&ta, A, M, K, lda, 64, BLOCK_K)) return e;
    if (int e = b_kmajor ? make_tmap(&tb, B, K, N, ldb, BLOCK_K, pair ? block_n / 2 : block_n)
                         : make_tmap(&tb, B, N, K, ldb, 64, BLOCK_K)) return e;

    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.ldd = ldd;
    p.a_kmajor = a_kmajor; p.b_kmajor = b_kmajor;
    p.out_f32 = out_dtype == B200MED_F32; p.relu = relu;
    p.split_k = split_k; p.kb_per_split = kb_per;
    p.tiles_m = pair ? (int)((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) : (int)((M + BLOCK_M - 1) / BLOCK_M);
    p.tiles_n = (int)((N + block_n - 1) / block_n);
    p.bias = split_k > 1 ? nullptr : bias;
    p.mask = split_k > 1 ? nullptr : reinterpret_cast<const __nv_bfloat16 *>(mask);
    p.D = split_k > 1 ? workspace : D;
    p.out_rbi = out_layout == B200MED_LAYOUT_RBI32;
    p.out_f16 = out_dtype == B200MED_F16;

    int e;
    // staged TMA-store epilogue: bf16 row-major output (and mask) whose rows TMA can address
    static const bool stage_allowed = []() { const char *v = getenv("B200MED_GEMM_TMA_EPILOGUE"); return !(v && v[0] == '0'); }();
    const bool staged = stage_allowed && split_k == 1 && !p.out_rbi && out_dtype == B200MED_BF16 && block_n >= 128 &&
                        ldd % 8 == 0 && (uintptr_t)D % 16 == 0 && (!mask || (uintptr_t)mask % 16 == 0);
    const int epi = split_k > 1 ? kEpiPartial : (p.out_rbi ? kEpiRbi : (staged ? kEpiTmaBf16 : kEpiRowMajor));
    CUtensorMap td = ta, tm = ta;
    if (staged) {
        if (int e2 = make_tmap(&td, D, N, M, ldd, 64, BLOCK_M)) return e2;               // box {64 columns, 128 rows}
        if (mask) if (int e2 = make_tmap(&tm, mask, N, M, ldd, 64, BLOCK_M)) return e2;
    }
#define B200MED_GEMM_EPI(BN, PAIR)                                                                                  \
    (epi == kEpiPartial ? launch_gemm_tc<BN, kEpiPartial, PAIR>(ta, tb, td, tm, p, st)                                  \
     : epi == kEpiRbi   ? launch_gemm_tc<BN, kEpiRbi, PAIR>(ta, tb, td, tm, p, st)                                      \
                        : launch_gemm_tc<BN, kEpiRowMajor, PAIR>(ta, tb, td, tm, p, st))
    if (pair) {
        e = epi == kEpiTmaBf16 ? launch_gemm_tc<256, kEpiTmaBf16, true>(ta, tb, td, tm, p, st) : B200MED_GEMM_EPI(256, true);
    } else if (block_n == 256) {
        static const bool dbl_allowed = []() { const char *v = getenv("B200MED_GEMM_DOUBLE_STAGING"); return !(v && v[0] == '0'); }();
        if (epi == kEpiTmaBf16 && kb_per <= 4 && dbl_allowed) e = launch_gemm_tc<256, kEpiTmaBf16, false, true>(ta, tb, td, tm, p, st);
        else e = epi == kEpiTmaBf16 ? launch_gemm_tc<256, kEpiTmaBf16, false>(ta, tb, td, tm, p, st) : B200MED_GEMM_EPI(256, false);
    } else if (block_n == 128) {
        e = epi == kEpiTmaBf16 ? launch_gemm_tc<128, kEpiTmaBf16, false>(ta, tb, td, tm, p, st) : B200MED_GEMM_EPI(128, false);
    } else if (block_n == 64) {
        e = B200MED_GEMM_EPI(64, false);
    } else {
        e = B200MED_GEMM_EPI(32, false);
    }
#undef B200MED_GEMM_EPI
    if (e) return e;
    if (split_k > 1) {
        const long long total = M * N;
        const long long want = (total + 255) / 256, cap = (long long)num_sms() * 8;
        launch_k(splitk_reduce_kernel, (unsigned)(want < cap ? want : cap), 256, 0, st, 
            reinterpret_cast<const float *>(workspace), split_k, M, N, ldd, bias, relu,
            reinterpret_cast<const __nv_bfloat16 *>(mask), D, out_dtype == B200MED_F32);
        return after_launch("splitk_reduce_kernel");
    }
    return B200MED_OK;
}
