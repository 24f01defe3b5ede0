// K1 + K2 fused: the window gather / standardise of the image stream as the A-OPERAND PRODUCER of the FeatureExtractor's first
// layer (MED/dataset/CustomWindowDataset.py:53-60 + MED/modeling/models.py:19-35 first Linear + ReLU; modeling_utils.py:40-41).
//
//     Xb [m, :] = bf16( (table[starts[m / W] + m % W, :] - mean) * (1 / std) )          (side output: operand of dW1 in the backward)
//     Y  [m, :] = bf16( relu( Xb[m, :] W1^T + b1 ) )                                      m = window * W + step, N = 512, K = 2048
//
// Unfused, the step writes the bf16 batch (0.54 GB at B = 8192) and the GEMM reads it back; here the fp32 table rows go
// HBM -> shared memory (TMA boxes of 64 columns x W rows, one per window: a window's frames are W consecutive table rows) ->
// converter warps (standardise, bf16, 128B-swizzled operand tile) -> tcgen05.mma, and the bf16 tile leaves ONCE by TMA store.
// HBM traffic per row: 8 KB read + 4 KB + 1 KB written, the GEMM's 4 KB re-read is gone; the kernel is HBM-bound
// (algorithmic 1.61 GB at B = 8192 against 0.275 TFLOP: 170 flop/B, under the ridge), the MMAs run in its shadow.
//
// CTA pair (cluster of 2, `tcgen05.mma.cta_group::2`, M = 256): CTA r gathers and converts rows [128 r, +128) of the pair's
// 256-row tile and stages HALF of every W1 k-block (256 of the 512 output rows; every byte of W1 comes from L2 once per 256
// rows), the leader issues two N = 256 MMAs per k-step into the 512 TMEM columns.  Warp roles per CTA: 0 TMA producer
// (gather boxes), 1 MMA issuer (leader) / TMEM allocator, 2..5 converters, 6..13 epilogue (bias, ReLU, bf16, stores), 14 W1 tile
// producer (its own warp: the gather producer must run ahead through the staging ring whatever the operand ring does).
#include "tcgen05.cuh"
#include <stdlib.h>

namespace b200med {

constexpr int kGgConvWarps = 8;          // converter warps: 16 rows of the 128-row tile each
constexpr int kGgThreads = (2 + kGgConvWarps + 8 + 1) * 32;
// Three rings, sized independently (template parameters FS / AS / BS): fp32 staging (bytes in flight are what an HBM-bound
// kernel lives on), the bf16 A operand, the W1 (B operand) tiles.  With ONE operand ring the W1 tile of k-step i+2 could only be
// requested when the MMAs of k-step i were done -- its L2 latency under a saturated memory system sat on the MMA chain.
constexpr int kGgN = 512;
constexpr uint32_t kGgFBytes = 128 * 256;            // fp32 staging stage: 128 rows x 64 columns
constexpr uint32_t kGgABytes = 128 * 128;            // bf16 A stage: 128 rows x 64 k
constexpr uint32_t kGgBBytes = 256 * 128;            // bf16 B stage: this CTA's 2 x 128 rows of the W1 k-block
constexpr int kGgMaxStages = 6;
template <int FS, int AS, int BS>
struct GgSmem {
    static constexpr uint32_t kOffA = FS * kGgFBytes;
    static constexpr uint32_t kOffB = kOffA + AS * kGgABytes;
    static constexpr uint32_t kOffBias = kOffB + BS * kGgBBytes;
    static constexpr uint32_t kOffBar = kOffBias + kGgN * 4;
    static constexpr uint32_t kUsed = kOffBar + 256;
    static constexpr uint32_t kBytes = kUsed + 512;  // + slack for the 1024-byte alignment (the declaration asks for it)
    static_assert(kBytes <= 232448, "over the 227 KB of shared memory a CTA can have");
    static_assert(FS <= kGgMaxStages && AS <= kGgMaxStages && BS <= kGgMaxStages && 2 * FS + 2 * AS + 2 * BS + 3 <= 32, "barrier block");
};

struct GatherGemmParams {
    const int32_t *starts;     // [B] first table row of every window
    const float *mean, *stdv;  // [K] or null (no standardisation)
    const float *bias;         // [512] or null
    __nv_bfloat16 *y;          // [B*W, 512]
    long long B, M;            // windows, rows = B*W
    int W, K, relu;
    long long table_rows;
    int debug;                 // experiments only (bits): 1 no Xb store, 2 no Y store, 4 no MMAs, 8 no conversion, 16 no W1 loads,
                               // 32 no evict-first hints, 64 one MMA per k-step instead of eight
};

__device__ __forceinline__ void tma_load_2d_f32(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(s_addr(dst)), "l"(map), "r"(s_addr(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tma_load_2d_f32_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        :: "r"(s_addr(dst)), "l"(map), "r"(s_addr(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *map, uint32_t src, int c0, int c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void bar_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
        if (!done) {
            __nanosleep(256);
            if (spins > (1u << 22)) __trap();
        }
    }
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// Explicit shared-space accesses: through generic pointers the compiler emits LD.E / ST.E and, unable to tell the staging buffer
// from the operand tile, orders every load behind the previous store (one exposed load latency per pass, measured).
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(a), "r"(x), "r"(y));
}
__device__ __forceinline__ void stg256(void *p, const uint32_t *v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// One warp's share of a tile's epilogue: 4 chunks of 32 accumulator columns [col0, col0 + 128) of the warp's TMEM lane quarter
// (a lane owns one output row): + bias, ReLU, bf16, two 256-bit stores per chunk (whole 32-byte sectors; the 16-byte stores of
// the first version left every sector half-written per instruction).  SIXTEEN warps per CTA share a tile -- the 8 epilogue warps
// and the 8 converter warps, which have nothing to convert while the accumulator is being drained (the MMAs of the next tile
// wait for it): the accumulator fills all 512 TMEM columns, so the MMAs idle for the length of the epilogue.
__device__ __forceinline__ void gg_epilogue_part(uint32_t tmem_base, int quarter, int col0, long long m, const GatherGemmParams &p,
                                                 uint32_t bias_s) {
    const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0;
    uint32_t vbuf[2][32];
    tmem_ld32_nowait(t_addr, vbuf[0]);
    tmem_wait_ld();
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
        uint32_t (&v)[32] = vbuf[ci & 1];
        if (ci + 1 < 4) tmem_ld32_nowait(t_addr + (uint32_t)((ci + 1) * 32), vbuf[(ci + 1) & 1]);
        const int c0 = col0 + ci * 32;
        uint32_t o[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = lds128(bias_s + (uint32_t)((c0 + 4 * j4) * 4));          // same address in every lane: broadcast
            float a0 = __uint_as_float(v[4 * j4]) + b4.x, a1 = __uint_as_float(v[4 * j4 + 1]) + b4.y;
            float a2 = __uint_as_float(v[4 * j4 + 2]) + b4.z, a3 = __uint_as_float(v[4 * j4 + 3]) + b4.w;
            if (p.relu) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); a2 = fmaxf(a2, 0.0f); a3 = fmaxf(a3, 0.0f); }
            o[2 * j4] = pack_bf16x2(a0, a1);
            o[2 * j4 + 1] = pack_bf16x2(a2, a3);
        }
        if (m < p.M && !(p.debug & 2)) {
            __nv_bfloat16 *dst = p.y + m * kGgN + c0;
            stg256(dst, o);
            stg256(dst + 16, o + 8);
        }
        if (ci + 1 < 4) tmem_wait_ld();
    }
    tcgen05_fence_before();
}

template <int FS, int AS, int BS>
__global__ void __launch_bounds__(kGgThreads, 1)
gather_gemm_kernel(const __grid_constant__ CUtensorMap tmap_table, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_xb, const GatherGemmParams p) {
    pdl_wait();
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    using S = GgSmem<FS, AS, BS>;
    if (threadIdx.x == 0 && (uint32_t)(smem - smem_dyn) + S::kUsed > S::kBytes) __trap();
    unsigned char *f_sm = smem;
    unsigned char *a_sm = smem + S::kOffA;
    unsigned char *b_sm = smem + S::kOffB;
    float *bias_sm = reinterpret_cast<float *>(smem + S::kOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::kOffBar);
    uint64_t *f_full = bars;                         // [FS] gather boxes of a k-block have landed (tx)
    uint64_t *f_empty = f_full + FS;                 // [FS] the converter warps are done with the staging buffer
    uint64_t *a_full = f_empty + FS;                 // [AS] leader: converter warps of BOTH CTAs have written their A tile
    uint64_t *a_empty = a_full + AS;                 // [AS] tcgen05.commit multicast: the MMAs have read A[s]
    uint64_t *b_full = a_empty + AS;                 // [BS] leader: both halves of the W1 k-block have landed
    uint64_t *b_empty = b_full + BS;                 // [BS] tcgen05.commit multicast: the MMAs have read B[s]
    uint64_t *acc_full = b_empty + BS;               // tcgen05.commit multicast
    uint64_t *acc_empty = acc_full + 1;              // leader: the 16 draining warps of both CTAs are done with the accumulator
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long long pair_id = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
    const long long tiles = (p.M + 255) / 256;
    const int nkb = p.K / 64;
    const int wins = 128 / p.W;                       // windows per CTA tile

    if (threadIdx.x == 0) {
        for (int i = 0; i < FS; ++i) { bar_init(&f_full[i], 1); bar_init(&f_empty[i], kGgConvWarps); }
        for (int i = 0; i < AS; ++i) { bar_init(&a_full[i], 2 * kGgConvWarps); bar_init(&a_empty[i], 1); }
        for (int i = 0; i < BS; ++i) { bar_init(&b_full[i], 2); bar_init(&b_empty[i], 1); }
        bar_init(acc_full, 1);
        bar_init(acc_empty, 2 * (8 + kGgConvWarps));      // every epilogue and converter warp of both CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kGgN; i += kGgThreads) bias_sm[i] = p.bias ? p.bias[i] : 0.0f;
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_table) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_w) : "memory");
            // the table slice of a step (1 GB) is read once: evict-first keeps W1 and the layer's output in L2 instead
            const uint64_t pol = policy_evict_first();
            const bool hint = !(p.debug & 32);
            long long it = 0;
            for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
                const long long m0 = tile * 256 + (long long)rank * 128;
                int st[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const long long b = m0 / p.W + w;
                    int s = (w < wins && b < p.B) ? p.starts[b] : 0;
                    if (s < 0 || (long long)s + p.W > p.table_rows) __trap();      // a window outside the table: the reference raises IndexError
                    st[w] = s;
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int f = (int)(it % FS);
                    const uint32_t fpar = (uint32_t)((it / FS) & 1);
                    bar_wait(&f_empty[f], fpar ^ 1);
                    bar_expect_tx(&f_full[f], kGgFBytes);
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (w < wins) {
                            if (hint) tma_load_2d_f32_hint(f_sm + f * kGgFBytes + w * p.W * 256, &tmap_table, &f_full[f], kb * 64, st[w], pol);
                            else tma_load_2d_f32(f_sm + f * kGgFBytes + w * p.W * 256, &tmap_table, &f_full[f], kb * 64, st[w]);
                        }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================== MMA issuer (leader CTA)
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc(256, 256, false, false);
            long long it = 0;
            uint32_t acc_par = 0;
            for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
                bar_wait(acc_empty, acc_par ^ 1);
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int sa_i = (int)(it % AS), sb_i = (int)(it % BS);
                    bar_wait(&a_full[sa_i], (uint32_t)((it / AS) & 1));
                    bar_wait(&b_full[sb_i], (uint32_t)((it / BS) & 1));
                    tcgen05_fence_after();
                    const uint32_t sa = s_addr(a_sm + sa_i * kGgABytes), sb = s_addr(b_sm + sb_i * kGgBBytes);
#pragma unroll
                    for (int nh = 0; nh < 2; ++nh)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if ((p.debug & 4) || ((p.debug & 64) && (nh | k))) continue;
                            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
                            const uint64_t db = make_smem_desc(sb + nh * 16384 + k * 32, 16, 1024);
                            umma_bf16_pair(tmem_base + (uint32_t)(nh * 256), da, db, idesc, (kb | k) ? 1u : 0u);
                        }
                    umma_commit_pair(&a_empty[sa_i]);
                    umma_commit_pair(&b_empty[sb_i]);
                }
                umma_commit_pair(acc_full);
                acc_par ^= 1;
            }
        }
        __syncwarp();
    } else if (warp == 2 + kGgConvWarps + 8) {
        // ================================================================== W1 tile producer
        if (lane == 0) {
            long long it = 0;
            for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % BS);
                    const uint32_t spar = (uint32_t)((it / BS) & 1);
                    bar_wait(&b_empty[s], spar ^ 1);
                    const uint32_t lbar = mapa_rank(&b_full[s], 0);
                    if (p.debug & 16) {
                        if (leader) bar_arrive(&b_full[s]); else bar_arrive_cluster(lbar);
                        continue;
                    }
                    if (leader) bar_expect_tx(&b_full[s], 2 * kGgBBytes);
                    // this CTA's half of the W1 k-block: rows [256 nh + 128 r, +128) for the two N = 256 MMAs
                    tma_load_2d_pair(b_sm + s * kGgBBytes, &tmap_w, lbar, kb * 64, (int)rank * 128);
                    tma_load_2d_pair(b_sm + s * kGgBBytes + 16384, &tmap_w, lbar, kb * 64, 256 + (int)rank * 128);
                    if (!leader) bar_arrive_cluster(lbar);
                }
            }
        }
        __syncwarp();
    } else if (warp < 2 + kGgConvWarps) {
        // ================================================================== converters (8 warps, 256 threads)
        // Warp cw owns rows [16 cw, +16) of the CTA's 128-row tile END TO END: it converts them and its lane 0 sends them to Xb with
        // its own TMA store (box of 16 rows), so no barrier between the converter warps is ever needed.
        // lane -> (column group of 4 fp32 = 16 B, row parity): a warp pass reads 2 rows x 256 B = 512 contiguous bytes (LDS.128,
        // conflict-free) and writes 2 x 128 B of the 128B-swizzled bf16 operand tile (STS.64, conflict-free).
        const int cw = warp - 2;
        const int cg = lane & 15, rp = lane >> 4;
        const uint32_t f_lane = s_addr(f_sm) + (uint32_t)((cw * 16 + rp) * 256 + cg * 16);       // + f * 32 KB + pass * 512
        uint32_t a_lane[4];                          // pass ps -> row 16 cw + 2 ps + rp; row % 8 = 2 (ps % 4) + rp picks the swizzle
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r7 = j * 2 + rp;
            a_lane[j] = s_addr(a_sm) + (uint32_t)((cw * 16 + r7) * 128 + ((((cg >> 1) ^ r7) << 4) | ((cg & 1) << 3)));
        }
        const uint64_t pol = policy_evict_first();
        const bool hint = !(p.debug & 32);
        // mean / std of k-step i+1 are fetched while k-step i is converted (their L2 latency sat on the critical path before)
        float4 m_nx = make_float4(0.f, 0.f, 0.f, 0.f), s_nx = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.mean) {
            m_nx = __ldg(reinterpret_cast<const float4 *>(p.mean + cg * 4));
            s_nx = __ldg(reinterpret_cast<const float4 *>(p.stdv + cg * 4));
        }
        // The converter warps take column blocks 2 and 3 of their TMEM lane quarter in every tile's epilogue.  They join it when
        // they have nothing else to do: the first AS k-steps of the NEXT tile are converted (the A ring is full and stays full
        // until the MMAs of that tile may start, i.e. until the accumulator is drained), or after the last tile.
        const int e_quarter = warp & 3, e_block = 2 + ((warp - 2) >> 2);
        const uint32_t bias_s = s_addr(bias_sm);
        uint32_t acc_par = 0;
        auto drain = [&](long long tile_done) {
            const long long m = tile_done * 256 + (long long)rank * 128 + e_quarter * 32 + lane;
            bar_wait(acc_full, acc_par);
            tcgen05_fence_after();
            gg_epilogue_part(tmem_base, e_quarter, e_block * 128, m, p, bias_s);
            __syncwarp();
            if (lane == 0) {
                if (leader) bar_arrive(acc_empty);
                else bar_arrive_cluster(mapa_rank(acc_empty, 0));
            }
            acc_par ^= 1;
        };
        const int drain_kb = AS < nkb ? AS : nkb - 1;      // short K: the ring holds a whole tile
        long long it = 0;
        for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
            const long long m0 = tile * 256 + (long long)rank * 128;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                if (kb == drain_kb && tile != pair_id) drain(tile - pair_stride);
                const int f = (int)(it % FS), s = (int)(it % AS);
                const uint32_t fpar = (uint32_t)((it / FS) & 1), spar = (uint32_t)((it / AS) & 1);
                const float mu0 = m_nx.x, mu1 = m_nx.y, mu2 = m_nx.z, mu3 = m_nx.w;
                const float iv0 = 1.0f / s_nx.x, iv1 = 1.0f / s_nx.y, iv2 = 1.0f / s_nx.z, iv3 = 1.0f / s_nx.w;
                if (p.mean) {
                    const int kn = (kb + 1 == nkb) ? 0 : kb + 1;
                    m_nx = __ldg(reinterpret_cast<const float4 *>(p.mean + kn * 64 + cg * 4));
                    s_nx = __ldg(reinterpret_cast<const float4 *>(p.stdv + kn * 64 + cg * 4));
                }
                bar_wait(&a_empty[s], spar ^ 1);                   // the MMAs of the slot's previous use are done ...
                if (lane == 0) bulk_wait_read<AS - 1>();            // ... and so is this warp's TMA store that read it
                __syncwarp();
                bar_wait(&f_full[f], fpar);
                const uint32_t fsrc = f_lane + (uint32_t)f * kGgFBytes;
                const uint32_t aoff = (uint32_t)s * kGgABytes;
                if (!(p.debug & 8)) {
                    float4 x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = lds128(fsrc + (uint32_t)(j * 512));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float y0 = (x[j].x - mu0) * iv0, y1 = (x[j].y - mu1) * iv1, y2 = (x[j].z - mu2) * iv2, y3 = (x[j].w - mu3) * iv3;
                        sts64(a_lane[j & 3] + aoff + (uint32_t)((j >> 2) * 1024), pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
                    }
                }
                fence_proxy_async_smem();                           // generic-proxy writes -> visible to tcgen05.mma / TMA
                __syncwarp();
                if (lane == 0) {
                    bar_arrive(&f_empty[f]);
                    if (leader) bar_arrive(&a_full[s]);
                    else bar_arrive_cluster(mapa_rank(&a_full[s], 0));
                    if (!(p.debug & 1)) {
                        const uint32_t src = s_addr(a_sm) + aoff + (uint32_t)(cw * 16 * 128);
                        if (hint) tma_store_2d_hint(&tmap_xb, src, kb * 64, (int)m0 + cw * 16, pol);
                        else tma_store_2d(&tmap_xb, a_sm + s * kGgABytes + cw * 16 * 128, kb * 64, (int)m0 + cw * 16);
                    }
                    bulk_commit();
                }
            }
        }
        drain(tiles - 1 - ((tiles - 1 - pair_id) % pair_stride));      // this pair's last tile
        if (lane == 0) bulk_wait_all();
        __syncwarp();
    } else {
        // ================================================================== epilogue warps (8): column blocks 0 and 1 of their quarter
        const int quarter = warp & 3, block = (warp - 2 - kGgConvWarps) >> 2;
        const uint32_t bias_s = s_addr(bias_sm);
        uint32_t acc_par = 0;
        for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
            const long long m = tile * 256 + (long long)rank * 128 + quarter * 32 + lane;
            bar_wait_sleep(acc_full, acc_par);       // a tile takes ~35 us: do not poll the issue slots away from the converters
            tcgen05_fence_after();
            gg_epilogue_part(tmem_base, quarter, block * 128, m, p, bias_s);
            __syncwarp();
            if (lane == 0) {
                if (leader) bar_arrive(acc_empty);
                else bar_arrive_cluster(mapa_rank(acc_empty, 0));
            }
            acc_par ^= 1;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace b200med

using namespace b200med;

// fp32 table [rows, K]: box {64 columns, W rows}, no swizzle (the converters read it row-wise)
static int make_tmap_table_f32(CUtensorMap *map, const void *ptr, long long K, long long rows, int W, int box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return B200MED_E_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)W};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(table f32) failed (CUresult %d)", (int)r); return B200MED_E_CUDA; }
    return B200MED_OK;
}

extern "C" __attribute__((visibility("default"))) int b200med_gather_linear_bf16(
    const float *table, int64_t table_rows, const float *mean, const float *stdv, const int32_t *starts, int64_t B, int32_t W,
    const void *w_bf16, const float *bias, int32_t relu, void *xb, void *y, int32_t N, int32_t K, void *stream) {
    B200MED_REQUIRE(B >= 1 && table_rows >= W && N == kGgN && K >= 64 && K % 64 == 0, "N must be 512 and K a multiple of 64");
    B200MED_REQUIRE(W >= 16 && W <= 128 && 128 % W == 0, "window length must be 16, 32, 64 or 128 (TMA boxes of W rows tile the 128-row operand)");
    B200MED_REQUIRE(table && starts && w_bf16 && y, "null pointer");      // xb may be NULL: inference keeps no bf16 batch
    B200MED_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must both be given or both NULL");
    B200MED_REQUIRE(((uintptr_t)table % 16 == 0) && ((uintptr_t)w_bf16 % 16 == 0) && ((uintptr_t)xb % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                    (!mean || (((uintptr_t)mean % 16 == 0) && ((uintptr_t)stdv % 16 == 0))), "operands must be 16-byte aligned");
    B200MED_REQUIRE((uintptr_t)y % 32 == 0, "y must be 32-byte aligned (256-bit stores)");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    const long long M = B * (long long)W;
    CUtensorMap tt, tw, tx;
    static int dbg_env = -1;
    if (dbg_env < 0) {
        const char *e2 = getenv("B200MED_GG_DEBUG");      // experiments only (scripts/bench_gather_gemm.py)
        dbg_env = e2 ? atoi(e2) : 0;
    }
    if (int e = make_tmap_table_f32(&tt, table, K, table_rows, W, 64)) return e;
    if (int e = make_tmap(&tw, w_bf16, K, N, K, 64, 128)) return e;            // W1 [512, K] bf16: box {64 k, 128 rows}
    if (xb) { if (int e = make_tmap(&tx, xb, K, M, K, 64, 128 / kGgConvWarps)) return e; }  // Xb [M, K] bf16: box {64 k, 16 rows}: one per converter warp
    else tx = tw;                                                                              // never dereferenced: no Xb store
    GatherGemmParams p{};
    p.starts = starts; p.mean = mean; p.stdv = stdv; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16 *>(y);
    p.B = B; p.M = M; p.W = W; p.K = K; p.relu = relu; p.table_rows = table_rows; p.debug = dbg_env | (xb ? 0 : 1);      // bit 1: the converter warps skip their TMA stores
    // ring depths: (fp32 staging, A, B); the alternatives are kept for scripts/bench_gather_gemm.py (B200MED_GG_RINGS)
    static int rings = -1;
    if (rings < 0) {
        const char *e3 = getenv("B200MED_GG_RINGS");
        rings = e3 ? atoi(e3) : 323;
    }
    auto kern = gather_gemm_kernel<3, 2, 3>;
    uint32_t smem_bytes = GgSmem<3, 2, 3>::kBytes;
    if (rings == 422) { kern = gather_gemm_kernel<4, 2, 2>; smem_bytes = GgSmem<4, 2, 2>::kBytes; }
    else if (rings == 224) { kern = gather_gemm_kernel<2, 2, 4>; smem_bytes = GgSmem<2, 2, 4>::kBytes; }
    else if (rings == 233) { kern = gather_gemm_kernel<2, 3, 3>; smem_bytes = GgSmem<2, 3, 3>::kBytes; }
    else if (rings == 332) { kern = gather_gemm_kernel<3, 3, 2>; smem_bytes = GgSmem<3, 3, 2>::kBytes; }
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes),
                           "cudaFuncSetAttribute(gather_gemm)")) return e;
    const long long tiles = (M + 255) / 256, pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * (tiles < pairs ? tiles : pairs)));
    cfg.blockDim = dim3(kGgThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1);
    if (int e = check_cuda(cudaLaunchKernelEx(&cfg, kern, tt, tw, tx, p), "cudaLaunchKernelEx(gather_gemm)")) return e;
    return after_launch("gather_gemm_kernel");
}
