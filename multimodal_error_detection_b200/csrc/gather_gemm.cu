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

constexpr int kGgThreads = 15 * 32;
// Bytes in flight are what an HBM-bound kernel lives on: 4 fp32 staging buffers (128 KB per SM under way; with 2 the kernel
// reached 0.61 of the copy peak, measured) and a 2-deep operand ring (the MMAs need half the time the loads need).
constexpr int kGgFStages = 4;            // fp32 staging, in k-steps (128 rows x 256 B each)
constexpr int kGgStages = 2;             // operand ring: A (16 KB) + B (32 KB) per stage
constexpr int kGgN = 512;
constexpr uint32_t kGgFBytes = 128 * 256;
constexpr uint32_t kGgABytes = 128 * 128;
constexpr uint32_t kGgBBytes = 256 * 128;
constexpr uint32_t kGgOffA = kGgFStages * kGgFBytes;
constexpr uint32_t kGgOffB = kGgOffA + kGgStages * kGgABytes;
constexpr uint32_t kGgOffBias = kGgOffB + kGgStages * kGgBBytes;
constexpr uint32_t kGgOffBar = kGgOffBias + kGgN * 4;
constexpr uint32_t kGgUsed = kGgOffBar + 256;
constexpr uint32_t kGgSmem = kGgUsed + 512;          // + slack for the 1024-byte alignment (the declaration asks for it)
static_assert(kGgSmem <= 232448, "over the 227 KB of shared memory a CTA can have");

struct GatherGemmParams {
    const int32_t *starts;     // [B] first table row of every window
    const float *mean, *stdv;  // [K] or null (no standardisation)
    const float *bias;         // [512] or null
    __nv_bfloat16 *y;          // [B*W, 512]
    long long B, M;            // windows, rows = B*W
    int W, K, relu;
    long long table_rows;
    int debug;                 // experiments only: 1 = no Xb store, 2 = no Y store
};

__device__ __forceinline__ void tma_load_2d_f32(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(s_addr(dst)), "l"(map), "r"(s_addr(bar)), "r"(c0), "r"(c1) : "memory");
}

// FK: k-steps per gather box (box = 64 FK fp32 columns x W rows; FK * 256 contiguous bytes per table row and request)
template <int FK>
__global__ void __launch_bounds__(kGgThreads, 1)
gather_gemm_kernel(const __grid_constant__ CUtensorMap tmap_table, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_xb, const GatherGemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0 && (uint32_t)(smem - smem_dyn) + kGgUsed > kGgSmem) __trap();
    unsigned char *f_sm = smem;
    unsigned char *a_sm = smem + kGgOffA;
    unsigned char *b_sm = smem + kGgOffB;
    float *bias_sm = reinterpret_cast<float *>(smem + kGgOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kGgOffBar);
    uint64_t *f_full = bars;                         // [F] gather boxes of a k-block have landed (tx)
    uint64_t *f_empty = bars + kGgFStages;           // [F] the 4 converter warps are done with the staging buffer
    uint64_t *a_full = f_empty + kGgFStages;         // [S] leader: converter warps of BOTH CTAs have written their A tile
    uint64_t *b_full = a_full + kGgStages;           // [S] leader: both halves of the W1 k-block have landed
    uint64_t *slot_empty = b_full + kGgStages;       // [S] tcgen05.commit multicast: the MMAs have read A[s] and B[s]
    uint64_t *acc_full = slot_empty + kGgStages;     // tcgen05.commit multicast
    uint64_t *acc_empty = acc_full + 1;              // leader: the 8 epilogue warps of both CTAs have drained the accumulator
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long long pair_id = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
    const long long tiles = (p.M + 255) / 256;
    const int nkb = p.K / 64;
    const int wins = 128 / p.W;                       // windows per CTA tile
    constexpr int NF = kGgFStages / FK;               // staging buffers
    constexpr uint32_t kFBuf = kGgFBytes * FK;        // bytes per staging buffer: 128 rows x (256 FK) B

    if (threadIdx.x == 0) {
        for (int i = 0; i < NF; ++i) { bar_init(&f_full[i], 1); bar_init(&f_empty[i], 4); }
        for (int i = 0; i < kGgStages; ++i) { bar_init(&a_full[i], 8); bar_init(&b_full[i], 2); bar_init(&slot_empty[i], 1); }
        bar_init(acc_full, 1);
        bar_init(acc_empty, 16);
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
                for (int kb = 0; kb < nkb; kb += FK, ++it) {
                    const int f = (int)(it % NF);
                    const uint32_t fpar = (uint32_t)((it / NF) & 1);
                    bar_wait(&f_empty[f], fpar ^ 1);
                    bar_expect_tx(&f_full[f], kFBuf);
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (w < wins)
                            tma_load_2d_f32(f_sm + f * kFBuf + w * p.W * (256 * FK), &tmap_table, &f_full[f], kb * 64, st[w]);
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
                    const int s = (int)(it % kGgStages);
                    const uint32_t spar = (uint32_t)((it / kGgStages) & 1);
                    bar_wait(&a_full[s], spar);
                    bar_wait(&b_full[s], spar);
                    tcgen05_fence_after();
                    const uint32_t sa = s_addr(a_sm + s * kGgABytes), sb = s_addr(b_sm + s * kGgBBytes);
#pragma unroll
                    for (int nh = 0; nh < 2; ++nh)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
                            const uint64_t db = make_smem_desc(sb + nh * 16384 + k * 32, 16, 1024);
                            umma_bf16_pair(tmem_base + (uint32_t)(nh * 256), da, db, idesc, (kb | k) ? 1u : 0u);
                        }
                    umma_commit_pair(&slot_empty[s]);
                }
                umma_commit_pair(acc_full);
                acc_par ^= 1;
            }
        }
        __syncwarp();
    } else if (warp == 14) {
        // ================================================================== W1 tile producer
        if (lane == 0) {
            long long it = 0;
            for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % kGgStages);
                    const uint32_t spar = (uint32_t)((it / kGgStages) & 1);
                    bar_wait(&slot_empty[s], spar ^ 1);
                    const uint32_t lbar = mapa_rank(&b_full[s], 0);
                    if (leader) bar_expect_tx(&b_full[s], 2 * kGgBBytes);
                    // this CTA's half of the W1 k-block: rows [256 nh + 128 r, +128) for the two N = 256 MMAs
                    tma_load_2d_pair(b_sm + s * kGgBBytes, &tmap_w, lbar, kb * 64, (int)rank * 128);
                    tma_load_2d_pair(b_sm + s * kGgBBytes + 16384, &tmap_w, lbar, kb * 64, 256 + (int)rank * 128);
                    if (!leader) bar_arrive_cluster(lbar);
                }
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ================================================================== converters (4 warps, 128 threads)
        // lane -> (column group of 4 fp32 = 16 B, row parity): a warp reads 2 rows x 256 B = 512 contiguous bytes per pass
        const int cw = warp - 2;
        const int cg = lane & 15, rp = lane >> 4;
        const bool issuer = cw == 0 && lane == 0;       // issues the TMA stores of the bf16 tile
        long long it = 0;
        for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
            const long long m0 = tile * 256 + (long long)rank * 128;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int f = (int)((it / FK) % NF), s = (int)(it % kGgStages);
                const int sub = (int)(it % FK);                     // k-step inside the staging buffer (K % (64 FK) == 0)
                const uint32_t fpar = (uint32_t)((it / FK / NF) & 1), spar = (uint32_t)((it / kGgStages) & 1);
                float mu[4] = {0.f, 0.f, 0.f, 0.f}, iv[4] = {1.f, 1.f, 1.f, 1.f};
                if (p.mean) {
                    const float4 m4 = __ldg(reinterpret_cast<const float4 *>(p.mean + kb * 64 + cg * 4));
                    const float4 s4 = __ldg(reinterpret_cast<const float4 *>(p.stdv + kb * 64 + cg * 4));
                    mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
                    iv[0] = 1.0f / s4.x; iv[1] = 1.0f / s4.y; iv[2] = 1.0f / s4.z; iv[3] = 1.0f / s4.w;
                }
                bar_wait(&slot_empty[s], spar ^ 1);                // the MMAs of the slot's previous use are done ...
                if (issuer) bulk_wait_read<kGgStages - 1>();        // ... and so is the TMA store that read it
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (sub == 0) bar_wait(&f_full[f], fpar);
                const unsigned char *fsrc = f_sm + f * kFBuf + sub * 256;
                unsigned char *adst = a_sm + s * kGgABytes;
#pragma unroll 4
                for (int ps = 0; ps < 16; ++ps) {
                    const int row = cw * 32 + ps * 2 + rp;
                    const float4 x = *reinterpret_cast<const float4 *>(fsrc + row * (256 * FK) + cg * 16);
                    const float y0 = (x.x - mu[0]) * iv[0], y1 = (x.y - mu[1]) * iv[1], y2 = (x.z - mu[2]) * iv[2], y3 = (x.w - mu[3]) * iv[3];
                    // 4 bf16 = 8 bytes: half of 16-byte slot cg/2 of the row, 128B swizzle (slot ^ row % 8)
                    unsigned char *dst = adst + row * 128 + ((((cg >> 1) ^ (row & 7)) << 4) | ((cg & 1) << 3));
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
                }
                fence_proxy_async_smem();                           // generic-proxy writes -> visible to tcgen05.mma / TMA
                __syncwarp();
                if (lane == 0) {
                    if (sub == FK - 1) bar_arrive(&f_empty[f]);
                    if (leader) bar_arrive(&a_full[s]);
                    else bar_arrive_cluster(mapa_rank(&a_full[s], 0));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");     // the whole tile is written: it may leave for Xb
                if (issuer) {
                    if (!(p.debug & 1)) tma_store_2d(&tmap_xb, adst, kb * 64, (int)m0);
                    bulk_commit();
                }
            }
        }
        if (issuer) bulk_wait_all();
    } else {
        // ================================================================== epilogue (8 warps): bias, ReLU, bf16, stores
        const int quarter = warp & 3, half = (warp - 6) >> 2;
        uint32_t acc_par = 0;
        for (long long tile = pair_id; tile < tiles; tile += pair_stride) {
            const long long m = tile * 256 + (long long)rank * 128 + quarter * 32 + lane;
            bar_wait(acc_full, acc_par);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 256);
            uint32_t vbuf[2][32];
            tmem_ld32_nowait(t_addr, vbuf[0]);
            tmem_wait_ld();
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                uint32_t (&v)[32] = vbuf[ci & 1];
                if (ci + 1 < 8) tmem_ld32_nowait(t_addr + (uint32_t)((ci + 1) * 32), vbuf[(ci + 1) & 1]);
                const int c0 = half * 256 + ci * 32;
                if (m < p.M && !(p.debug & 2)) {
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(v[2 * j]) + bias_sm[c0 + 2 * j];
                        float b = __uint_as_float(v[2 * j + 1]) + bias_sm[c0 + 2 * j + 1];
                        if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        o[j] = pack_bf16x2(a, b);
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(p.y + m * kGgN + c0);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
                }
                if (ci + 1 < 8) tmem_wait_ld();
            }
            tcgen05_fence_before();
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
    B200MED_REQUIRE(table && starts && w_bf16 && xb && y, "null pointer");
    B200MED_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean and std must both be given or both NULL");
    B200MED_REQUIRE(((uintptr_t)table % 16 == 0) && ((uintptr_t)w_bf16 % 16 == 0) && ((uintptr_t)xb % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                    (!mean || (((uintptr_t)mean % 16 == 0) && ((uintptr_t)stdv % 16 == 0))), "operands must be 16-byte aligned");
    if (!b200med_has_tcgen05()) { set_error("tcgen05 path needs a compute-capability 10.x device"); return B200MED_E_UNSUPPORTED; }
    const long long M = B * (long long)W;
    CUtensorMap tt, tw, tx;
    static int fk_env = -1, dbg_env = 0;
    if (fk_env < 0) {
        const char *e1 = getenv("B200MED_GG_FK"), *e2 = getenv("B200MED_GG_DEBUG");
        fk_env = e1 ? atoi(e1) : 1; dbg_env = e2 ? atoi(e2) : 0;
    }
    const int FK = (fk_env == 2 && K % 128 == 0) ? 2 : 1;
    if (int e = make_tmap_table_f32(&tt, table, K, table_rows, W, 64 * FK)) return e;
    if (int e = make_tmap(&tw, w_bf16, K, N, K, 64, 128)) return e;            // W1 [512, K] bf16: box {64 k, 128 rows}
    if (int e = make_tmap(&tx, xb, K, M, K, 64, 128)) return e;                 // Xb [M, K] bf16: box {64 k, 128 rows}
    GatherGemmParams p{};
    p.starts = starts; p.mean = mean; p.stdv = stdv; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16 *>(y);
    p.B = B; p.M = M; p.W = W; p.K = K; p.relu = relu; p.table_rows = table_rows; p.debug = dbg_env;
    static bool attr_set = false;
    if (!attr_set) {
        if (int e = check_cuda(cudaFuncSetAttribute(gather_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGgSmem),
                               "cudaFuncSetAttribute(gather_gemm)")) return e;
        if (int e = check_cuda(cudaFuncSetAttribute(gather_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGgSmem),
                               "cudaFuncSetAttribute(gather_gemm)")) return e;
        attr_set = true;
    }
    const long long tiles = (M + 255) / 256, pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * (tiles < pairs ? tiles : pairs)));
    cfg.blockDim = dim3(kGgThreads);
    cfg.dynamicSmemBytes = kGgSmem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (int e = check_cuda(FK == 2 ? cudaLaunchKernelEx(&cfg, gather_gemm_kernel<2>, tt, tw, tx, p)
                                   : cudaLaunchKernelEx(&cfg, gather_gemm_kernel<1>, tt, tw, tx, p), "cudaLaunchKernelEx(gather_gemm)")) return e;
    return after_launch("gather_gemm_kernel");
}
