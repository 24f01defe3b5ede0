// K0: window index construction and powerset label transform (integer work, bit-exact bar).
//
// The walk over one subject is inherently sequential (the stride grid re-phases after every
// gesture boundary) but subjects are independent, so the table is processed
// with one thread per subject: a fold has 10^1..10^5 subjects, each a few hundred frames (2-4 KB
// of gesture ids, L1/L2 resident).  Algorithmic bytes: 4 B/frame read, 4 B/window written -- three
// orders of magnitude below K1, so this kernel is sized for latency, not bandwidth.
#include "common.cuh"

namespace b200med {

// One subject's walk.  EMIT=false counts, EMIT=true writes.  Float compares on purpose: the
// reference compares Python floats taken with .item() (dataset_utils.py:220-223), so NaN != NaN
// skips a frame and -0.0 counts as the zero gesture, exactly as there.
template <bool EMIT>
__device__ __forceinline__ long long walk_subject(const float *__restrict__ g, long long base, long long n,
                                                  int W, int S, long long out_base, int32_t *starts) {
    long long pos = 0;
    while (pos < n && !(g[base + pos] != 0.0f)) ++pos;  // first non-zero gesture (:211-212)
    if (pos >= n) return -1;
    long long count = 0;
    while (pos < n - W) {                                // strict bound (:214)
        const float a = g[base + pos], b = g[base + pos + W - 1];
        if (a != b) { pos += 1; continue; }              // end points only (:220-226)
        if (EMIT) starts[out_base + count] = (int32_t)(base + pos);
        ++count;
        pos += S;                                        // (:239)
    }
    return count;
}

// status holds the SMALLEST failing subject index (or INT32_MAX when none); finalised by the scan.
__global__ void window_count_kernel(const float *__restrict__ g, const int64_t *__restrict__ off,
                                     long long n_subj, int W, int S, int64_t *__restrict__ counts,
                                     int32_t *__restrict__ status) {
    pdl_wait();
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n_subj) return;
    const long long c = walk_subject<false>(g, off[s], off[s + 1] - off[s], W, S, 0, nullptr);
    if (c < 0) atomicMin(status, (int32_t)s);
    counts[s] = c < 0 ? 0 : c;
}

// In-place exclusive scan of counts[0..n) into counts[0..n]; single block, fixed order.
__global__ void exclusive_scan_kernel(int64_t *__restrict__ data, long long n, int32_t *__restrict__ status) {
    pdl_wait();
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const long long v = i < n ? data[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_tot[warp - 1] : 0) + incl - v;
        if (i < n) data[i] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += warp_tot[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        data[n] = carry;
        if (*status == INT32_MAX) *status = -1;
    }
}

__global__ void window_fill_kernel(const float *__restrict__ g, const int64_t *__restrict__ off,
                                   const int64_t *__restrict__ win_off, long long n_subj, int W, int S,
                                   const float *__restrict__ e5, int32_t *__restrict__ starts,
                                   float *__restrict__ g_win, float *__restrict__ e5_win,
                                   int32_t *__restrict__ subj_win) {
    pdl_wait();
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n_subj) return;
    const long long o = win_off[s];
    const long long c = walk_subject<true>(g, off[s], off[s + 1] - off[s], W, S, o, starts);
    // labels of the first frame of every window (:232-233)
    for (long long k = 0; k < c; ++k) {
        const long long row = starts[o + k];
        if (g_win) g_win[o + k] = g[row];
        if (subj_win) subj_win[o + k] = (int32_t)s;
        if (e5 && e5_win) {
#pragma unroll
            for (int j = 0; j < 5; ++j) e5_win[(o + k) * 5 + j] = e5[row * 5 + j];
        }
    }
}

__global__ void set_i32_kernel(int32_t *p, int32_t v) {
    pdl_wait(); *p = v; }

// One thread per row; first matching rule wins, same order as dataset_utils.py:796-843.
__global__ void powerset_kernel(const float *__restrict__ e5, long long n, int delete_nd,
                                int32_t *__restrict__ e7, uint8_t *__restrict__ nd_mask) {
    pdl_wait();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float oov = e5[i * 5 + 0], nd = e5[i * 5 + 1], ma = e5[i * 5 + 2], np_ = e5[i * 5 + 3],
                err = e5[i * 5 + 4];
    int32_t o[7] = {0, 0, 0, 0, 0, 0, 0};
    uint8_t m = 0;
    if (err == 1.0f) {
        o[6] = 1;
        const bool single = __fadd_rn(__fadd_rn(__fadd_rn(oov, nd), ma), np_) == 1.0f;
        if ((oov == 1.0f && single) || (oov == 1.0f && nd == 1.0f)) o[1] = 1;
        else if ((ma == 1.0f && single) || (ma == 1.0f && nd == 1.0f)) o[2] = 1;
        else if ((np_ == 1.0f && single) || (np_ == 1.0f && oov == 1.0f)) o[3] = 1;
        else if (oov == 1.0f && ma == 1.0f) o[4] = 1;
        else if (ma == 1.0f && np_ == 1.0f) o[5] = 1;
        else if (nd == 1.0f) { if (delete_nd) { o[6] = 0; m = 1; } }
    } else {
        o[0] = 1;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) e7[i * 7 + j] = o[j];
    nd_mask[i] = m;
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int b200med_window_count(const float *g, const int64_t *subj_offsets, int64_t n_subjects, int32_t W,
                                    int32_t S, int64_t *win_offsets, int32_t *status, void *stream) {
    B200MED_REQUIRE(g && subj_offsets && win_offsets && status, "null pointer");
    B200MED_REQUIRE(n_subjects >= 0 && W >= 1 && S >= 1, "need n_subjects >= 0, W >= 1, S >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    launch_k(set_i32_kernel, 1, 1, 0, st, status, INT32_MAX);
    if (int e = after_launch("set_i32_kernel")) return e;
    if (n_subjects > 0) {
        const int threads = 128;
        const long long blocks = (n_subjects + threads - 1) / threads;
        launch_k(window_count_kernel, (unsigned)blocks, threads, 0, st, g, subj_offsets, n_subjects, W, S, win_offsets, status);
        if (int e = after_launch("window_count_kernel")) return e;
    }
    launch_k(exclusive_scan_kernel, 1, 1024, 0, st, win_offsets, n_subjects, status);
    return after_launch("exclusive_scan_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_window_fill(const float *g, const int64_t *subj_offsets, const int64_t *win_offsets,
                                   int64_t n_subjects, int32_t W, int32_t S, const float *e5, int32_t *starts,
                                   float *g_win, float *e5_win, int32_t *subj_win, void *stream) {
    B200MED_REQUIRE(g && subj_offsets && win_offsets && starts, "null pointer");
    B200MED_REQUIRE(n_subjects >= 0 && W >= 1 && S >= 1, "need n_subjects >= 0, W >= 1, S >= 1");
    if (n_subjects == 0) return B200MED_OK;
    const int threads = 128;
    const long long blocks = (n_subjects + threads - 1) / threads;
    launch_k(window_fill_kernel, (unsigned)blocks, threads, 0, (cudaStream_t)stream, 
        g, subj_offsets, win_offsets, n_subjects, W, S, e5, starts, g_win, e5_win, subj_win);
    return after_launch("window_fill_kernel");
}

extern "C" __attribute__((visibility("default"))) int b200med_powerset(const float *e5, int64_t n, int32_t delete_nd, int32_t *e7, uint8_t *nd_mask,
                                void *stream) {
    B200MED_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return B200MED_OK;
    B200MED_REQUIRE(e5 && e7 && nd_mask, "null pointer");
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    launch_k(powerset_kernel, (unsigned)blocks, threads, 0, (cudaStream_t)stream, e5, n, delete_nd, e7, nd_mask);
    return after_launch("powerset_kernel");
}
