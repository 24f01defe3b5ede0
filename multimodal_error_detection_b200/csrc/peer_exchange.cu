// Data-parallel gradient exchange over NVLink peer memory: ONE kernel per rank and step that sums the flat gradient buffers of
// all ranks of a box (the `loss.backward()` + DDP-style average of SURVEY section 8e; the reference trains one process,
// MED/modeling/modeling_utils.py:363-365) -- replaces the NCCL all-reduce between the backward and the Adam kernel.
//
// Every rank maps every other rank's gradient buffer and flag block (CUDA IPC, exchanged once through torch.distributed).
// The kernel is a reduce-scatter + all-gather by direct loads / stores:
//   A  "my gradients are complete" -> flag in every peer (release.sys); wait for all peers' flags (acquire.sys);
//   B  rank r owns slice r: s = sum over q = 0 .. world-1 of G_q[slice r] (peer loads, FIXED order: every rank ends up with the
//      same bits, computed once by the slice's owner) and stores s into G_q[slice r] of EVERY rank (peer stores);
//   C  fence, "my slice is everywhere" -> flags; wait for all peers; the kernel ends when this rank's buffer is complete.
// 6.4 MB of gradients at 8 ranks: 2 x 7/8 x 6.4 MB per rank over NVLink (~16 us at 700 GB/s) + two flag round trips, against
// ~0.1 ms for the NCCL call inside the captured step.  Every spin is bounded (two minutes, then trap): a rank that never launches
// its kernel is an error on the others, not a hung box.  The kernel may be captured in a CUDA graph (all arguments are device-resident).
#include "common.cuh"

namespace b200med {

constexpr int kPeerMaxRanks = 16;
// flag block (uint32): [0, 16) phase-A arrivals, [16, 32) phase-C arrivals, [32] epoch of the last finished exchange,
// [33] CTAs of this rank that have pushed their part, [34] CTAs that have left
constexpr int kPeerFlagWords = 64;

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float *p, const float4 &v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Ranks may arrive seconds apart (a rank that captures a graph, pages its table in, runs a host-side baseline): the wait is
// bounded in TIME -- two minutes -- not in polls; past that a peer never launched its kernel and this is an error, not a wait.
constexpr unsigned long long kPeerWaitNs = 120ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void wait_flag(const uint32_t *p, uint32_t epoch) {
    // epochs only grow; the comparison is wrap-safe
    unsigned long long t0 = 0;
    for (unsigned int spins = 0; (int32_t)(ld_acquire_sys(p) - epoch) < 0; ++spins) {
        __nanosleep(spins < 64 ? 32 : 256);
        if ((spins & 1023u) == 1023u) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kPeerWaitNs) __trap();
        }
    }
}

__global__ void __launch_bounds__(512)
peer_allreduce_kernel(float *const *__restrict__ bufs, uint32_t *const *__restrict__ flags, int rank, int world, long long n) {
    pdl_wait();
    uint32_t *mine = flags[rank];
    __shared__ uint32_t epoch_sm;
    if (threadIdx.x == 0) epoch_sm = ld_acquire_sys(mine + 32) + 1;      // bumped only by the LAST CTA of a launch to leave
    __syncthreads();
    const uint32_t epoch = epoch_sm;

    // ---- A: the backward of this rank is complete (stream order); tell every peer, wait for every peer
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x < world) wait_flag(mine + threadIdx.x, epoch);
    __syncthreads();

    // ---- B: reduce slice `rank` over all ranks in rank order, write the sum into every rank's buffer
    const long long per = ((n + world - 1) / world + 3) / 4 * 4;
    const long long lo = (long long)rank * per, hi = (lo + per < n) ? lo + per : n;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += stride) {
        if (i + 4 <= hi) {
            float4 s = ld_sys_f4(bufs[0] + i);
            for (int q = 1; q < world; ++q) {
                const float4 v = ld_sys_f4(bufs[q] + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            for (int q = 0; q < world; ++q) st_sys_f4(bufs[q] + i, s);
        } else {                                     // ragged end of the buffer (n not a multiple of 4)
            for (long long j = i; j < hi; ++j) {
                float s = 0.0f;
                for (int q = 0; q < world; ++q) s += *reinterpret_cast<volatile float *>(bufs[q] + j);
                for (int q = 0; q < world; ++q) *reinterpret_cast<volatile float *>(bufs[q] + j) = s;
            }
        }
    }

    // ---- C: all of this rank's stores are out -> signal; wait until every peer's slice has landed here
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t last_sm;
    if (threadIdx.x == 0) last_sm = (atomicAdd(mine + 33, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (last_sm && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(flags[threadIdx.x] + 16 + rank, epoch);
    }
    if (threadIdx.x < world) wait_flag(mine + 16 + threadIdx.x, epoch);
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(mine + 34, 1u) == gridDim.x - 1) {
        mine[33] = 0; mine[34] = 0;
        __threadfence();
        st_release_sys(mine + 32, epoch);
    }
}

}  // namespace b200med

using namespace b200med;

extern "C" __attribute__((visibility("default"))) int b200med_peer_alloc(int64_t bytes, void **ptr) {
    B200MED_REQUIRE(bytes > 0 && ptr, "bad arguments");
    // plain cudaMalloc (not a pooled / virtual-memory allocation): what cudaIpcGetMemHandle can export
    if (int e = check_cuda(cudaMalloc(ptr, (size_t)bytes), "cudaMalloc(peer buffer)")) return e;
    return check_cuda(cudaMemset(*ptr, 0, (size_t)bytes), "cudaMemset(peer buffer)");
}
extern "C" __attribute__((visibility("default"))) int b200med_peer_free(void *ptr) {
    return ptr ? check_cuda(cudaFree(ptr), "cudaFree(peer buffer)") : B200MED_OK;
}
extern "C" __attribute__((visibility("default"))) int b200med_peer_export(const void *ptr, void *handle64) {
    B200MED_REQUIRE(ptr && handle64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (int e = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)), "cudaIpcGetMemHandle")) return e;
    memcpy(handle64, &h, 64);
    return B200MED_OK;
}
extern "C" __attribute__((visibility("default"))) int b200med_peer_import(const void *handle64, void **ptr) {
    B200MED_REQUIRE(handle64 && ptr, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return check_cuda(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
extern "C" __attribute__((visibility("default"))) int b200med_peer_close(void *ptr) {
    return ptr ? check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle") : B200MED_OK;
}
extern "C" __attribute__((visibility("default"))) int64_t b200med_peer_flag_bytes(void) { return kPeerFlagWords * 4; }

extern "C" __attribute__((visibility("default"))) int b200med_peer_allreduce_f32(void *const *bufs_dev, void *const *flags_dev, int32_t rank,
                                                                                 int32_t world, int64_t n, void *stream) {
    B200MED_REQUIRE(bufs_dev && flags_dev && n >= 1, "bad arguments");
    B200MED_REQUIRE(world >= 1 && world <= kPeerMaxRanks && rank >= 0 && rank < world, "rank / world out of range (at most 16 ranks)");
    // one CTA per SM at most (every CTA must be resident: they meet at the flags), 2 KB of loads in flight per thread group
    const long long per = (n + world - 1) / world;
    long long ctas = (per + 512 * 4 - 1) / (512 * 4);
    const long long cap = num_sms();
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    peer_allreduce_kernel<<<(unsigned)ctas, 512, 0, (cudaStream_t)stream>>>(      // plain launch: its CTAs spin on each other and on the peers, they do not sit next to a draining predecessor
        reinterpret_cast<float *const *>(bufs_dev), reinterpret_cast<uint32_t *const *>(flags_dev), rank, world, n);
    return after_launch("peer_allreduce_kernel");
}
