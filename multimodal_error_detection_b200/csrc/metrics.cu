// Ranking metric on the device: area under the ROC curve of (score, binary label) pairs.
//   roc_auc  <- sklearn.metrics.roc_auc_score as the reference calls it on the stored per-sample probabilities
//               (MED/modeling/modeling_utils.py:1124, 1243) and as north_star's "frame-level F1/AUC" bar names it.
//
// AUC = P(score_pos > score_neg) + 0.5 P(score_pos == score_neg) (Mann-Whitney; equal to the trapezoid area sklearn
// integrates, ties included).  All counting is INTEGER work, so the result does not depend on any reduction order:
//   1. keys: the scores of the NEGATIVE samples as order-preserving uint32 keys, every other slot = 0xFFFFFFFF (sorts last);
//   2. bitonic sort of the padded key array (shared-memory stages for spans <= 2048 keys, global stages above);
//   3. every POSITIVE sample binary-searches the sorted negatives: 2 * (#neg below) + (#neg equal) summed in uint64;
//   4. auc = that sum / (2 n_pos n_neg) in fp64.
// Latency-bound (a test fold has 10^2..10^5 samples); the sort is n log^2 n compare-exchanges on 4-byte keys.
#include "common.cuh"

namespace b200med {

constexpr int kSortBlock = 1024;            // threads per sorting CTA
constexpr int kSortSpan = 2 * kSortBlock;   // keys a CTA sorts / merges in shared memory

__device__ __forceinline__ uint32_t order_key(float x) {
    x += 0.0f;                               // -0.0 -> +0.0: the two compare equal as floats
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// stats: [0] n_pos, [1] n_neg, [2] 2*less + equal, [3] unused
__global__ void auc_keys_kernel(const float *__restrict__ score, const float *__restrict__ label, long long n,
                                long long npad, uint32_t *__restrict__ keys, unsigned long long *__restrict__ stats) {
    pdl_wait();
    unsigned int neg = 0, pos = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npad; i += (long long)gridDim.x * blockDim.x) {
        uint32_t k = 0xFFFFFFFFu;
        if (i < n) {
            if (label[i] > 0.5f) ++pos;
            else { k = order_key(score[i]); ++neg; }
        }
        keys[i] = k;
    }
    const long long nn = warp_sum((long long)neg), pp = warp_sum((long long)pos);
    if ((threadIdx.x & 31) == 0) {
        if (nn) atomicAdd(&stats[1], (unsigned long long)nn);
        if (pp) atomicAdd(&stats[0], (unsigned long long)pp);
    }
}

__device__ __forceinline__ void cmp_swap(uint32_t &a, uint32_t &b, bool up) {
    if ((a > b) == up) { const uint32_t t = a; a = b; b = t; }
}

// FULL = true: sort each span of kSortSpan keys (stages k = 2 .. kSortSpan); the direction of a span alternates with its
// index so that the global stages can merge them.  FULL = false: finish stage k (k > kSortSpan) for j = kSortBlock .. 1.
template <bool FULL>
__global__ void __launch_bounds__(kSortBlock) bitonic_shared_kernel(uint32_t *__restrict__ keys, long long k_stage) {
    pdl_wait();
    __shared__ uint32_t s[kSortSpan];
    const long long base = (long long)blockIdx.x * kSortSpan;
    s[threadIdx.x] = keys[base + threadIdx.x];
    s[threadIdx.x + kSortBlock] = keys[base + threadIdx.x + kSortBlock];
    __syncthreads();
    if (FULL) {
        for (int k = 2; k <= kSortSpan; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int t = threadIdx.x;
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));        // index with bit j cleared
                const bool up = (((base + lo) & k) == 0);
                cmp_swap(s[lo], s[lo | j], up);
                __syncthreads();
            }
        }
    } else {
        for (int j = kSortBlock; j > 0; j >>= 1) {
            const int t = threadIdx.x;
            const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const bool up = (((base + lo) & k_stage) == 0);
            cmp_swap(s[lo], s[lo | j], up);
            __syncthreads();
        }
    }
    keys[base + threadIdx.x] = s[threadIdx.x];
    keys[base + threadIdx.x + kSortBlock] = s[threadIdx.x + kSortBlock];
}

__global__ void bitonic_global_kernel(uint32_t *__restrict__ keys, long long npad, long long k, long long j) {
    pdl_wait();
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < npad / 2; t += (long long)gridDim.x * blockDim.x) {
        const long long lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const bool up = ((lo & k) == 0);
        uint32_t a = keys[lo], b = keys[lo | j];
        if ((a > b) == up) { keys[lo] = b; keys[lo | j] = a; }
    }
}

__global__ void auc_count_kernel(const float *__restrict__ score, const float *__restrict__ label, long long n,
                                 const uint32_t *__restrict__ sorted, unsigned long long *__restrict__ stats) {
    pdl_wait();
    const long long n_neg = (long long)stats[1];
    long long acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (!(label[i] > 0.5f)) continue;
        const uint32_t key = order_key(score[i]);
        long long lo = 0, hi = n_neg;                 // first index with sorted[idx] >= key
        while (lo < hi) { const long long m = (lo + hi) >> 1; if (sorted[m] < key) lo = m + 1; else hi = m; }
        const long long lb = lo;
        hi = n_neg;                                   // first index with sorted[idx] > key
        while (lo < hi) { const long long m = (lo + hi) >> 1; if (sorted[m] <= key) lo = m + 1; else hi = m; }
        acc += 2 * lb + (lo - lb);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&stats[2], (unsigned long long)acc);
}

__global__ void auc_final_kernel(const unsigned long long *__restrict__ stats, double *__restrict__ auc) {
    pdl_wait();
    const double np = (double)stats[0], nn = (double)stats[1];
    // one class only: sklearn raises ValueError; the host wrapper does the same from stats, the device value is NaN
    *auc = (np > 0 && nn > 0) ? (double)stats[2] / (2.0 * np * nn) : __longlong_as_double(0x7FF8000000000000LL);
}

}  // namespace b200med

using namespace b200med;

static long long auc_npad(long long n) {
    long long p = kSortSpan;
    while (p < n) p <<= 1;
    return p;
}

extern "C" __attribute__((visibility("default"))) int64_t b200med_roc_auc_ws_bytes(int64_t n) {
    return n < 0 ? 0 : (int64_t)(auc_npad(n) * sizeof(uint32_t));
}

extern "C" __attribute__((visibility("default"))) int b200med_roc_auc(const float *scores, const float *labels, int64_t n, double *auc,
                                int64_t *stats, void *workspace, void *stream) {
    B200MED_REQUIRE(n >= 1 && n <= (1LL << 30), "1 .. 2^30 samples");
    B200MED_REQUIRE(scores && labels && auc && stats && workspace, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long npad = auc_npad(n);
    uint32_t *keys = reinterpret_cast<uint32_t *>(workspace);
    unsigned long long *s64 = reinterpret_cast<unsigned long long *>(stats);
    if (int e = check_cuda(cudaMemsetAsync(stats, 0, 4 * sizeof(int64_t), st), "cudaMemsetAsync(stats)")) return e;
    const long long cap = (long long)num_sms() * 8;
    auto grid_for = [&](long long items, int threads) { const long long w = (items + threads - 1) / threads; return (unsigned)(w < cap ? (w < 1 ? 1 : w) : cap); };
    launch_k(auc_keys_kernel, grid_for(npad, 256), 256, 0, st, scores, labels, n, npad, keys, s64);
    if (int e = after_launch("auc_keys_kernel")) return e;
    launch_k(bitonic_shared_kernel<true>, (unsigned)(npad / kSortSpan), kSortBlock, 0, st, keys, 0);
    if (int e = after_launch("bitonic_shared_kernel")) return e;
    for (long long k = 2LL * kSortSpan; k <= npad; k <<= 1) {
        for (long long j = k >> 1; j > kSortBlock; j >>= 1) {
            launch_k(bitonic_global_kernel, grid_for(npad / 2, 256), 256, 0, st, keys, npad, k, j);
            if (int e = after_launch("bitonic_global_kernel")) return e;
        }
        launch_k(bitonic_shared_kernel<false>, (unsigned)(npad / kSortSpan), kSortBlock, 0, st, keys, k);
        if (int e = after_launch("bitonic_shared_kernel")) return e;
    }
    launch_k(auc_count_kernel, grid_for(n, 256), 256, 0, st, scores, labels, n, keys, s64);
    if (int e = after_launch("auc_count_kernel")) return e;
    launch_k(auc_final_kernel, 1, 1, 0, st, s64, auc);
    return after_launch("auc_final_kernel");
}
