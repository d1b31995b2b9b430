// Shared helpers for libjck_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/jck_b200.h"

namespace jck {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

int set_error(int code, const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Count a launch and turn a launch failure into an error code (no sync: asynchronous faults
// surface at the caller's next synchronisation, as with any CUDA library).
#define JCK_LAUNCH_CHECK(name)                                                          \
    do {                                                                                \
        ::jck::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return ::jck::set_error(JCK_E_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
    } while (0)

#define JCK_REQUIRE(cond, ...)                                            \
    do {                                                                  \
        if (!(cond)) return ::jck::set_error(JCK_E_BADARG, __VA_ARGS__);  \
    } while (0)

// activation element access in either dtype
template <typename T> __device__ __forceinline__ float ld_act(const T* p);
template <> __device__ __forceinline__ float ld_act<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_act<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_act(T* p, float v);
template <> __device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Stride-2 pad-1 k4 tap geometry.
//   down (large -> small): input row iy = 2*oy + ky - 1.
//   up   (small -> large), output row 2*i + py takes, for tap t in {0,1}:
//       py = 0: (ky = 1, di = 0), (ky = 3, di = -1)      py = 1: (ky = 2, di = 0), (ky = 0, di = +1)
__host__ __device__ __forceinline__ int up_k(int parity, int t) { return parity == 0 ? (t == 0 ? 1 : 3) : (t == 0 ? 2 : 0); }
__host__ __device__ __forceinline__ int up_d(int parity, int t) { return t == 0 ? 0 : (parity == 0 ? -1 : 1); }

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch ---------------------------------------------------------------------------
// The train step is ~150 dependent launches in one CUDA graph; a full kernel -> kernel dependency costs a drain +
// launch gap of a few microseconds each.  Every kernel here is launched with programmatic stream serialization:
// it signals `launch_dependents` on entry (its successor may be scheduled as soon as all of this grid's CTAs are
// resident) and executes `griddepcontrol.wait` before its first global-memory access -- which blocks until the
// predecessor grid has COMPLETED and flushed, so data dependencies are exactly those of a normal launch; only the
// launch latency, CTA scheduling and the on-chip prologue (barrier init, TMEM allocation, descriptor prefetch)
// overlap the predecessor's tail.  Both instructions are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }

bool pdl_enabled();   // JCK_PDL=0 in the environment turns the attribute off (A/B timing)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only: the "already configured" memo of a
// launcher is kept per device (one process per GPU is the rule here, but the C ABI does not forbid a caller with several).
struct DeviceOnce {
    bool flags[64] = {};
    int dev() const { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : -1; }
    bool done() const { const int d = dev(); return d >= 0 && flags[d]; }
    void mark() { const int d = dev(); if (d >= 0) flags[d] = true; }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace jck
