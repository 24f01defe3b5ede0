// Shared helpers for libb200med.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/b200med.h"

namespace b200med {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;
extern thread_local int g_sm_limit;   // b200med_set_sm_limit(): SMs the persistent kernels launched by this thread may fill
extern std::atomic<int> g_pdl;        // b200med_set_pdl(): launches carry the programmatic-stream-serialization attribute

inline int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return B200MED_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B200MED_E_CUDA;
}

inline int after_launch(const char *kernel) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_cuda(cudaGetLastError(), kernel);
}

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return (g_sm_limit > 0 && g_sm_limit < sms) ? g_sm_limit : sms;
}

// Programmatic dependent launch.  EVERY kernel of this library starts with pdl_wait() (griddepcontrol.wait: returns once
// all grids this launch depends on have completed and their writes are visible; a no-op for a launch without the attribute),
// and every launch goes through launch_k / pdl_attr, so that inside a stream -- and inside a captured CUDA graph, where the
// edge between two such kernel nodes becomes a programmatic one -- the CTAs of kernel k+1 are scheduled while kernel k drains
// instead of after its completion has travelled back to the front end (~2-3 us of the ~5 us a launch costs in the
// replayed step).  Nothing issues griddepcontrol.launch_dependents early: the implicit trigger at the exit of the primary's
// CTAs keeps a dependent from occupying SMs the primary still needs.  Ordering stays transitive because no CTA of any kernel
// can leave before its own wait has returned.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline int pdl_attr(cudaLaunchAttribute *a) {      // fills a[0], returns the number of attributes written (0 = switched off)
    if (!g_pdl.load(std::memory_order_relaxed)) return 0;
    a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a[0].val.programmaticStreamSerializationAllowed = 1;
    return 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = (unsigned)pdl_attr(attr);
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define B200MED_REQUIRE(cond, msg)                       \
    do {                                                 \
        if (!(cond)) {                                   \
            b200med::set_error("%s: %s", __func__, msg); \
            return B200MED_E_ARG;                        \
        }                                                \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming load / store (data touched once: keep it out of L1).
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

// fp16 pair, saturated to the finite range by the conversion itself (F2FP.SATFINITE: one instruction; used for bounded
// quantities -- LSTM gate pre-activations / activations)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

}  // namespace b200med
