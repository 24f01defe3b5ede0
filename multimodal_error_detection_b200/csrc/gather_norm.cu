// K1: fused window gather + per-stream standardise + concat, from the device-resident frame table.
//
//   out_s[b*W + t, col_s + d] = (table_s[starts[b] + t, d] - mean_s[d]) / std_s[d]
//
// Roofline: HBM bandwidth.  Algorithmic bytes per window = W * sum_s D_s * (s_in + s_out)
// (SURVEY.md section 8d): 165 920 B at W=10, D=2048+26, f32->f32; 199 104 B at W=16, f32->bf16.
// A window's W frames are CONTIGUOUS rows of the table (window_data never crosses a subject and
// subjects are contiguous), so the "gather" is B independent contiguous runs of W*D elements --
// fully coalesced 128-bit traffic; only the run starts are data dependent.
//
// Three device paths, chosen per stream on the host:
//   wide     D*4 bytes a multiple of the CTA's vector footprint (the 2048-d image stream): every
//            thread owns fixed columns, keeps mean/std in registers and streams U rows at a time
//            with 128-bit L1-bypassing loads -> U*16 B in flight per thread.
//   tma      same mapping, but the rows of a chunk are staged global->shared by one elected thread
//            with cp.async.bulk (1-D TMA) into an mbarrier-guarded ring, so the LSU only carries the
//            shared-memory reads and the global stores.
//   generic  any D / alignment (26-d kinematics, concat offsets): one element per thread, the
//            window's run is still read contiguously.
// All streams of a call are served by ONE launch: a CTA first decodes which stream its work unit
// belongs to (block-uniform branch).
#include "common.cuh"

namespace b200med {

constexpr int kThreads = 256;
constexpr int kRowsPerUnit = 4;  // U: rows in flight per thread in the wide path

struct StreamParams {
    const void *table;
    const float *mean;
    const float *stdv;
    void *out;
    int dim, table_dtype, out_dtype, out_ld, out_col, stat_rows, exact_div, path;  // path: 0 generic, 1 wide
    int table_rows;        // rows of the table (0 = unknown): a window that leaves the table traps instead of reading past it
    long long unit_begin;  // first global work-unit index of this stream
    int slabs;             // wide: column slabs per row; generic: unused
    int chunks;            // wide: row chunks per window
};

struct GatherParams {
    StreamParams s[B200MED_MAX_STREAMS];
    int n_streams;
    int W;
    long long B;
    long long total_units;
    const int32_t *starts;
};

template <bool EXACT>
__device__ __forceinline__ float standardise(float x, float mean, float sd_or_inv) {
    // EXACT: IEEE-754 subtract then divide, the op order of CustomWindowDataset.py:58,60.
    return EXACT ? __fdiv_rn(__fsub_rn(x, mean), sd_or_inv) : (x - mean) * sd_or_inv;
}

// A window start outside the table is an index bug of the caller (the reference raises IndexError): trap -- the launch
// fails with a CUDA error -- instead of reading past the table.
__device__ __forceinline__ long long checked_start(const StreamParams &sp, const int32_t *__restrict__ starts, long long b, int W) {
    const long long s = starts[b];
    if (sp.table_rows > 0 && (s < 0 || s + W > (long long)sp.table_rows)) __trap();
    return s;
}

// ---- wide path: f32 table, VPT consecutive floats per thread, U rows in flight ------------------
template <typename OutT, bool EXACT>
__device__ __forceinline__ void wide_unit(const StreamParams &sp, const int32_t *__restrict__ starts, int W,
                                          long long unit) {
    constexpr int VPT = sizeof(OutT) == 2 ? 8 : 4;  // keep the STORE 128-bit wide
    constexpr int NV = VPT / 4;
    const int slab = (int)(unit % sp.slabs);
    const long long rest = unit / sp.slabs;
    const int chunk = (int)(rest % sp.chunks);
    const long long b = rest / sp.chunks;
    const int col = (slab * kThreads + threadIdx.x) * VPT;
    const int t0 = chunk * kRowsPerUnit;
    const int rows = min(kRowsPerUnit, W - t0);
    const long long src_row = checked_start(sp, starts, b, W) + t0;
    const float *src = reinterpret_cast<const float *>(sp.table) + src_row * sp.dim + col;

    float4 x[kRowsPerUnit][NV];
#pragma unroll
    for (int r = 0; r < kRowsPerUnit; ++r)
        if (r < rows) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
                x[r][v] = ldg_stream(reinterpret_cast<const float4 *>(src + (long long)r * sp.dim) + v);
        }

    float4 mu[NV], sd[NV];
    const bool per_step = sp.stat_rows > 1;
    auto load_stats = [&](int t) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            if (sp.mean) {
                mu[v] = __ldg(reinterpret_cast<const float4 *>(sp.mean + (long long)t * sp.dim + col) + v);
                sd[v] = __ldg(reinterpret_cast<const float4 *>(sp.stdv + (long long)t * sp.dim + col) + v);
                if (!EXACT) sd[v] = make_float4(1.0f / sd[v].x, 1.0f / sd[v].y, 1.0f / sd[v].z, 1.0f / sd[v].w);
            } else {
                mu[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                sd[v] = make_float4(1.f, 1.f, 1.f, 1.f);
            }
        }
    };
    if (!per_step) load_stats(0);

    OutT *dst = reinterpret_cast<OutT *>(sp.out) + ((long long)b * W + t0) * sp.out_ld + sp.out_col + col;
#pragma unroll
    for (int r = 0; r < kRowsPerUnit; ++r)
        if (r < rows) {
            if (per_step) load_stats(t0 + r);
            float y[VPT];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                y[4 * v + 0] = standardise<EXACT>(x[r][v].x, mu[v].x, sd[v].x);
                y[4 * v + 1] = standardise<EXACT>(x[r][v].y, mu[v].y, sd[v].y);
                y[4 * v + 2] = standardise<EXACT>(x[r][v].z, mu[v].z, sd[v].z);
                y[4 * v + 3] = standardise<EXACT>(x[r][v].w, mu[v].w, sd[v].w);
            }
            OutT *d = dst + (long long)r * sp.out_ld;
            if constexpr (sizeof(OutT) == 4) {
                stg_stream(reinterpret_cast<float4 *>(d), make_float4(y[0], y[1], y[2], y[3]));
            } else {
                stg_stream(reinterpret_cast<uint4 *>(d),
                           make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                                      pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7])));
            }
        }
}

// ---- generic path: one element per thread over the window's contiguous run -----------------------
__device__ __forceinline__ float load_elem(const void *table, int dtype, long long idx) {
    return dtype == B200MED_F32 ? __ldg(reinterpret_cast<const float *>(table) + idx)
                                : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(table)[idx]);
}

__device__ __forceinline__ void generic_unit(const StreamParams &sp, const int32_t *__restrict__ starts, int W,
                                             long long unit) {
    // unit -> (window b, slice of the W*dim run); kThreads*4 elements per unit
    const int run = W * sp.dim;
    const int per_unit = kThreads * 4;
    const int units_per_window = (run + per_unit - 1) / per_unit;
    const long long b = unit / units_per_window;
    const int u = (int)(unit % units_per_window);
    const long long src0 = checked_start(sp, starts, b, W) * sp.dim;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = u * per_unit + k * kThreads + threadIdx.x;
        if (e >= run) break;
        const int t = e / sp.dim, d = e - t * sp.dim;
        float x = load_elem(sp.table, sp.table_dtype, src0 + e);
        if (sp.mean) {
            const int srow = sp.stat_rows > 1 ? t : 0;
            const float mu = __ldg(sp.mean + (long long)srow * sp.dim + d);
            const float sd = __ldg(sp.stdv + (long long)srow * sp.dim + d);
            x = sp.exact_div ? __fdiv_rn(__fsub_rn(x, mu), sd) : (x - mu) * (1.0f / sd);
        }
        const long long o = ((long long)b * W + t) * sp.out_ld + sp.out_col + d;
        if (sp.out_dtype == B200MED_F32) reinterpret_cast<float *>(sp.out)[o] = x;
        else reinterpret_cast<__nv_bfloat16 *>(sp.out)[o] = __float2bfloat16_rn(x);
    }
}

// Stream 0 may be a wide stream (template-specialised on its output type / division mode); all other
// streams go through the generic path.  One launch serves every stream of the call.
template <typename OutT, bool EXACT, bool HAS_WIDE>
__global__ void __launch_bounds__(kThreads, (sizeof(OutT) == 2 ? 3 : 4))
gather_norm_kernel(const __grid_constant__ GatherParams p) {
    pdl_wait();
    const long long wide_units = HAS_WIDE ? p.s[1].unit_begin : 0;  // s[1].unit_begin == units of stream 0
    for (long long unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        if (HAS_WIDE && unit < wide_units) {
            wide_unit<OutT, EXACT>(p.s[0], p.starts, p.W, unit);
        } else {
            int si = HAS_WIDE ? 1 : 0;
#pragma unroll
            for (int k = 1; k < B200MED_MAX_STREAMS; ++k)
                if (k < p.n_streams && unit >= p.s[k].unit_begin) si = k;
            // copy the (small) descriptor out of the parameter bank once per unit
            StreamParams sp;
#pragma unroll
            for (int k = 0; k < B200MED_MAX_STREAMS; ++k)
                if (k == si) sp = p.s[k];
            generic_unit(sp, p.starts, p.W, unit - sp.unit_begin);
        }
    }
}

// ---- TMA staging variant (wide f32 streams only) --------------------------------------------------
// Ring of kStages shared-memory buffers, each holding one chunk (U rows x D floats, contiguous in
// the table).  Thread 0 is the producer: it arms the stage's mbarrier with the byte count and issues
// ONE cp.async.bulk per chunk; all threads consume.  A stage is re-armed only after the CTA-wide
// barrier that follows its consumption, so no separate "empty" barrier is needed.

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin (like every mbarrier wait of this library): a bulk copy that never lands -- a bad source address -- traps
// and surfaces as a CUDA error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// R = rows per stage (stage = R*D*4 bytes), S = ring depth, T = threads per CTA.
template <typename OutT, bool EXACT, int R, int S, int T>
__global__ void __launch_bounds__(T)
gather_norm_tma_kernel(const __grid_constant__ GatherParams p) {
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[S];
    const StreamParams &sp = p.s[0];
    const int D = sp.dim;
    const int W = p.W;
    const int chunks = sp.chunks;
    const long long n_units = (long long)p.B * chunks;  // unit = (window, row chunk), all columns
    const size_t stage_bytes = (size_t)R * D * sizeof(float);
    float *stage_ptr[S];
#pragma unroll
    for (int s = 0; s < S; ++s) stage_ptr[s] = reinterpret_cast<float *>(smem_raw + s * stage_bytes);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](long long unit, int stage) {
        const int chunk = (int)(unit % chunks);
        const long long b = unit / chunks;
        const int t0 = chunk * R;
        const int rows = min(R, W - t0);
        const uint32_t bytes = (uint32_t)((size_t)rows * D * sizeof(float));
        const float *src = reinterpret_cast<const float *>(sp.table) + (checked_start(sp, p.starts, b, W) + t0) * D;
        mbar_expect_tx(&full_bar[stage], bytes);
        tma_bulk_g2s(stage_ptr[stage], src, bytes, &full_bar[stage]);
    };

    // prologue: fill the ring
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            const long long u = blockIdx.x + (long long)s * gridDim.x;
            if (u < n_units) issue(u, s);
        }
    }

    // Consumer mapping: a thread owns groups of 8 consecutive columns (two LDS.128 in, one 128-bit bf16 store or two
    // 128-bit f32 stores out); for D = 8*T (the 2048-d image stream with 256 threads) that is exactly one group, whose
    // mean / 1/std stay in registers for the whole kernel.
    const int ngroups = D / 8;
    const bool per_step = sp.stat_rows > 1;
    float mu[8], sd[8];
    auto load_stats = [&](int grp, int srow) {
        if (sp.mean) {
            const float4 *m4 = reinterpret_cast<const float4 *>(sp.mean + (long long)srow * D + grp * 8);
            const float4 *s4 = reinterpret_cast<const float4 *>(sp.stdv + (long long)srow * D + grp * 8);
            const float4 a0 = __ldg(m4), a1 = __ldg(m4 + 1), b0 = __ldg(s4), b1 = __ldg(s4 + 1);
            mu[0] = a0.x; mu[1] = a0.y; mu[2] = a0.z; mu[3] = a0.w; mu[4] = a1.x; mu[5] = a1.y; mu[6] = a1.z; mu[7] = a1.w;
            sd[0] = b0.x; sd[1] = b0.y; sd[2] = b0.z; sd[3] = b0.w; sd[4] = b1.x; sd[5] = b1.y; sd[6] = b1.z; sd[7] = b1.w;
            if (!EXACT) {
#pragma unroll
                for (int k = 0; k < 8; ++k) sd[k] = 1.0f / sd[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) { mu[k] = 0.0f; sd[k] = 1.0f; }
        }
    };
    const bool single_group = ngroups <= T;
    if (single_group && !per_step && (int)threadIdx.x < ngroups) load_stats(threadIdx.x, 0);

    int it = 0;
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int stage = it % S;
        const uint32_t parity = (uint32_t)((it / S) & 1);
        const int chunk = (int)(unit % chunks);
        const long long b = unit / chunks;
        const int t0 = chunk * R;
        const int rows = min(R, W - t0);
        mbar_wait(&full_bar[stage], parity);
        const float *buf = stage_ptr[stage];
        OutT *dst_base = reinterpret_cast<OutT *>(sp.out) + ((long long)b * W + t0) * sp.out_ld + sp.out_col;
        for (int grp = threadIdx.x; grp < ngroups; grp += T) {
            if (!single_group && !per_step) load_stats(grp, 0);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < rows) {
                    if (per_step) load_stats(grp, t0 + r);
                    const float4 x0 = *reinterpret_cast<const float4 *>(buf + (size_t)r * D + grp * 8);
                    const float4 x1 = *reinterpret_cast<const float4 *>(buf + (size_t)r * D + grp * 8 + 4);
                    const float y0 = standardise<EXACT>(x0.x, mu[0], sd[0]), y1 = standardise<EXACT>(x0.y, mu[1], sd[1]),
                                y2 = standardise<EXACT>(x0.z, mu[2], sd[2]), y3 = standardise<EXACT>(x0.w, mu[3], sd[3]),
                                y4 = standardise<EXACT>(x1.x, mu[4], sd[4]), y5 = standardise<EXACT>(x1.y, mu[5], sd[5]),
                                y6 = standardise<EXACT>(x1.z, mu[6], sd[6]), y7 = standardise<EXACT>(x1.w, mu[7], sd[7]);
                    OutT *d = dst_base + (long long)r * sp.out_ld + grp * 8;
                    if constexpr (sizeof(OutT) == 4) {
                        stg_stream(reinterpret_cast<float4 *>(d), make_float4(y0, y1, y2, y3));
                        stg_stream(reinterpret_cast<float4 *>(d) + 1, make_float4(y4, y5, y6, y7));
                    } else {
                        stg_stream(reinterpret_cast<uint4 *>(d), make_uint4(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3),
                                                                           pack_bf16x2(y4, y5), pack_bf16x2(y6, y7)));
                    }
                }
            }
        }
        __syncthreads();  // everyone is done reading this stage
        if (threadIdx.x == 0) {
            const long long next = unit + (long long)S * gridDim.x;
            if (next < n_units) issue(next, stage);
        }
    }
}

__global__ void standardise_rows_kernel(const float *__restrict__ x, const float *__restrict__ mean,
                                        const float *__restrict__ stdv, float *__restrict__ out, long long rows,
                                        int dim, int out_ld, int out_col) {
    pdl_wait();
    const long long n = rows * dim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim;
        const int d = (int)(i - r * dim);
        out[r * out_ld + out_col + d] = __fdiv_rn(__fsub_rn(x[i], mean[d]), stdv[d]);
    }
}

}  // namespace b200med

using namespace b200med;

// which device path the last b200med_gather_norm call of this thread took for stream 0 (tests assert that the
// benchmarked instantiation is the one they compared with the oracle): 1 = LDG kernel only, 2.. = TMA staging ring shape
static thread_local int g_last_gather_variant = 0;
extern "C" __attribute__((visibility("default"))) int b200med_gather_last_variant(void) { return g_last_gather_variant; }

extern "C" __attribute__((visibility("default"))) int b200med_gather_norm(const b200med_stream_desc *streams, int32_t n_streams, const int32_t *starts,
                                   int64_t B, int32_t W, int32_t variant, void *stream) {
    B200MED_REQUIRE(streams && n_streams >= 1 && n_streams <= B200MED_MAX_STREAMS, "1..8 streams");
    B200MED_REQUIRE(B >= 0 && W >= 1, "need B >= 0 and W >= 1");
    if (B == 0) return B200MED_OK;
    B200MED_REQUIRE(starts, "null starts");
    // bits 8..23 of `variant`: optional cap on the number of SMs the launch may occupy (0 = all).  A gather that is issued
    // on a side stream next to other kernels leaves them the rest of the GPU (one CTA per SM, 192 KB of shared memory).
    const int sm_cap = (variant >> 8) & 0xFFFF;
    variant &= 0xFF;
    const int sms = sm_cap > 0 && sm_cap < num_sms() ? sm_cap : num_sms();
    GatherParams p{};
    p.n_streams = n_streams;
    p.W = W;
    p.B = B;
    p.starts = starts;
    long long units = 0;
    for (int i = 0; i < n_streams; ++i) {
        const b200med_stream_desc &d = streams[i];
        B200MED_REQUIRE(d.table && d.out && d.dim >= 1, "stream needs table, out, dim >= 1");
        B200MED_REQUIRE((d.mean == nullptr) == (d.stdv == nullptr), "mean and std must both be given or both NULL");
        B200MED_REQUIRE(d.stat_rows == 1 || d.stat_rows == W, "stat_rows must be 1 or W");
        B200MED_REQUIRE(d.out_ld >= d.out_col + d.dim, "out_ld too small for out_col + dim");
        B200MED_REQUIRE(d.table_dtype == B200MED_F32 || d.table_dtype == B200MED_BF16, "bad table dtype");
        B200MED_REQUIRE(d.out_dtype == B200MED_F32 || d.out_dtype == B200MED_BF16, "bad out dtype");
        StreamParams &sp = p.s[i];
        sp.table = d.table; sp.mean = d.mean; sp.stdv = d.stdv; sp.out = d.out;
        sp.dim = d.dim; sp.table_dtype = d.table_dtype; sp.out_dtype = d.out_dtype;
        sp.out_ld = d.out_ld; sp.out_col = d.out_col; sp.stat_rows = d.stat_rows; sp.exact_div = d.exact_div;
        sp.table_rows = d.table_rows > 0 ? d.table_rows : 0;
        const int vpt = d.out_dtype == B200MED_BF16 ? 8 : 4;
        const size_t out_es = d.out_dtype == B200MED_BF16 ? 2 : 4;
        const bool aligned = ((uintptr_t)d.table % 16 == 0) && ((uintptr_t)d.out % 16 == 0) &&
                             ((d.out_ld * out_es) % 16 == 0) && ((d.out_col * out_es) % 16 == 0) &&
                             (!d.mean || (((uintptr_t)d.mean % 16 == 0) && ((uintptr_t)d.stdv % 16 == 0)));
        const bool wide = d.table_dtype == B200MED_F32 && aligned && (d.dim % (kThreads * vpt) == 0);
        sp.path = wide ? 1 : 0;
        sp.unit_begin = units;
        if (wide) {
            sp.slabs = d.dim / (kThreads * vpt);
            sp.chunks = (W + kRowsPerUnit - 1) / kRowsPerUnit;
            units += (long long)B * sp.chunks * sp.slabs;
        } else {
            const long long run = (long long)W * d.dim;
            units += (long long)B * ((run + kThreads * 4 - 1) / (kThreads * 4));
        }
    }
    p.total_units = units;
    cudaStream_t st = (cudaStream_t)stream;

    // TMA staging: only for a single wide f32 stream (the image stream); other streams of the call
    // go through the LDG kernel in a second launch.
    // variant 0 = measured best per output type (bench_gather, B200): f32 output -> LDG path (0.84-0.85 of the copy
    // peak at B=8192); bf16 output -> TMA staging ring, 8-row stages when W is a multiple of 8 (0.85), else 2-row stages.
    if (variant == 0 && p.s[0].path == 1 && p.s[0].out_dtype == B200MED_BF16 && B * (long long)W >= 4096)
        variant = (W % 8 == 0) ? 5 : 4;
    g_last_gather_variant = 1;
    if (variant >= 2 && variant != 3 && p.s[0].path == 1) {
        g_last_gather_variant = variant;
        GatherParams pt = p;
        pt.n_streams = 1;
        // variant -> (rows per stage, stages, threads): 2 = (4,3,256)  4 = (2,4,256)  5 = (8,3,256)  6 = (4,3,512)  7 = (2,6,256)
        int R = 4, S = 3, T = 256;
        if (variant == 4) { R = 2; S = 4; }
        else if (variant == 5) { R = 8; S = 3; }
        else if (variant == 6) { T = 512; }
        else if (variant == 7) { R = 2; S = 6; }
        pt.s[0].chunks = (W + R - 1) / R;
        const size_t smem = (size_t)S * R * pt.s[0].dim * sizeof(float);
        B200MED_REQUIRE(smem <= 200 * 1024, "row too wide for the TMA staging ring");
        const long long n_units = (long long)B * pt.s[0].chunks;
        const int per_sm = (int)((220 * 1024) / (smem + 1024));
        const long long cap_t = (long long)(per_sm < 1 ? 1 : per_sm) * sms;
        const int grid = (int)(n_units < cap_t ? n_units : cap_t);
        auto launch = [&](auto kern, int threads) -> int {
            if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                   "cudaFuncSetAttribute(gather_norm_tma)")) return e;
            launch_k(kern, grid, threads, smem, st, pt);
            return after_launch("gather_norm_tma_kernel");
        };
        const bool f32 = pt.s[0].out_dtype == B200MED_F32, ex = pt.s[0].exact_div != 0;
        int e = B200MED_E_UNSUPPORTED;
#define B200MED_TMA_CASE(RR, SS, TT)                                                                              \
        if (R == RR && S == SS && T == TT) {                                                                          \
            if (f32) e = ex ? launch(gather_norm_tma_kernel<float, true, RR, SS, TT>, TT)                             \
                            : launch(gather_norm_tma_kernel<float, false, RR, SS, TT>, TT);                           \
            else e = ex ? launch(gather_norm_tma_kernel<__nv_bfloat16, true, RR, SS, TT>, TT)                         \
                        : launch(gather_norm_tma_kernel<__nv_bfloat16, false, RR, SS, TT>, TT);                       \
        }
        B200MED_TMA_CASE(4, 3, 256)
        B200MED_TMA_CASE(2, 4, 256)
        B200MED_TMA_CASE(8, 3, 256)
        B200MED_TMA_CASE(4, 3, 512)
        B200MED_TMA_CASE(2, 6, 256)
#undef B200MED_TMA_CASE
        if (e) return e;
        if (n_streams == 1) return B200MED_OK;
        // remaining streams
        GatherParams pr{};
        pr.n_streams = n_streams - 1; pr.W = W; pr.B = B; pr.starts = starts;
        const long long shift = p.s[1].unit_begin;
        for (int i = 1; i < n_streams; ++i) { pr.s[i - 1] = p.s[i]; pr.s[i - 1].unit_begin -= shift; }
        pr.total_units = units - shift;
        p = pr;
    }
    // stream order for the single launch: at most one wide stream, placed first
    int wide_idx = -1;
    for (int i = 0; i < p.n_streams; ++i)
        if (p.s[i].path == 1) { wide_idx = i; break; }
    GatherParams q{};
    q.n_streams = p.n_streams; q.W = W; q.B = B; q.starts = starts;
    {
        int k = 0;
        long long u = 0;
        auto push = [&](const StreamParams &src, bool as_wide) {
            StreamParams sp = src;
            sp.unit_begin = u;
            if (as_wide) {
                u += (long long)B * sp.chunks * sp.slabs;
            } else {
                sp.path = 0;
                const long long run = (long long)W * sp.dim;
                u += (long long)B * ((run + kThreads * 4 - 1) / (kThreads * 4));
            }
            q.s[k++] = sp;
        };
        if (wide_idx >= 0) push(p.s[wide_idx], true);
        for (int i = 0; i < p.n_streams; ++i)
            if (i != wide_idx) push(p.s[i], false);
        if (k < B200MED_MAX_STREAMS) q.s[k].unit_begin = u;  // sentinel: s[1].unit_begin must exist
        q.total_units = u;
    }
    const long long max_grid = (long long)sms * 8;  // up to 8 CTAs of 256 threads per SM
    const int grid = (int)(q.total_units < max_grid ? q.total_units : max_grid);
    if (wide_idx < 0) {
        launch_k(gather_norm_kernel<float, true, false>, grid, kThreads, 0, st, q);
    } else if (q.s[0].out_dtype == B200MED_F32) {
        if (q.s[0].exact_div) launch_k(gather_norm_kernel<float, true, true>, grid, kThreads, 0, st, q);
        else launch_k(gather_norm_kernel<float, false, true>, grid, kThreads, 0, st, q);
    } else {
        if (q.s[0].exact_div) launch_k(gather_norm_kernel<__nv_bfloat16, true, true>, grid, kThreads, 0, st, q);
        else launch_k(gather_norm_kernel<__nv_bfloat16, false, true>, grid, kThreads, 0, st, q);
    }
    return after_launch("gather_norm_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_standardise_rows(const float *x, const float *mean, const float *stdv, float *out,
                                        int64_t rows, int32_t dim, int32_t out_ld, int32_t out_col, void *stream) {
    B200MED_REQUIRE(rows >= 0 && dim >= 1 && out_ld >= out_col + dim, "bad shape");
    if (rows == 0) return B200MED_OK;
    B200MED_REQUIRE(x && mean && stdv && out, "null pointer");
    const long long n = rows * dim;
    const long long want = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    launch_k(standardise_rows_kernel, (unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream, 
        x, mean, stdv, out, rows, dim, out_ld, out_col);
    return after_launch("standardise_rows_kernel");
}
