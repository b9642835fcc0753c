// nodey_common.cuh -- shared helpers for the sm_100a kernels (error plumbing, launch geometry,
// explicitly rounded arithmetic).  Everything is compiled with -fmad=false; fused multiply-adds
// appear only where written as __fmaf_rn (the polyphase FIR, whose oracle order uses fmaf).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nodey_cuda.h"

namespace nodey {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define NODEY_CUDA_OK(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) return ::nodey::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define NODEY_LAUNCH_OK()                                                          \
    do {                                                                           \
        cudaError_t _e = cudaPeekAtLastError();                                    \
        if (_e != cudaSuccess) return ::nodey::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define NODEY_REQUIRE(cond, code, ...)                                             \
    do {                                                                           \
        if (!(cond)) { ::nodey::set_error(__VA_ARGS__); return (code); }           \
    } while (0)

// Every kernel launch site opens one of these: it counts the launch (nodey_profile_launches) and,
// while profiling is enabled, brackets the launch with CUDA events on the launching stream so that
// nodey_profile_report() can give per-kernel device time (bench.py's roofline leg).
struct LaunchScope {
    LaunchScope(const char* name, cudaStream_t st, double algo_bytes = 0.0);
    ~LaunchScope();
    void* rec;
    cudaStream_t st;
};

#define NODEY_LAUNCH(name, stream, ...)                                            \
    do {                                                                           \
        ::nodey::LaunchScope _ls(name, stream);                                    \
        __VA_ARGS__;                                                               \
    } while (0)

// SM count of the current device (cached); grids are sized in multiples of it.
int sm_count();

// stream-ordered device memory with caching of large blocks (runtime.cu)
int device_alloc(void** out, size_t bytes, cudaStream_t stream);
int device_free(void* p, cudaStream_t stream);


// grid for a grid-stride streaming kernel: ctas_per_sm resident CTAs on every SM, capped by work
inline int stream_grid(int64_t work_items, int block, int ctas_per_sm)
{
    int64_t need = (work_items + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

inline cudaStream_t as_stream(nodey_stream_t s) { return (cudaStream_t)s; }

inline bool fmt_planar(int fmt) { return fmt >= NODEY_FMT_U8P; }
inline int fmt_bytes(int fmt)
{
    switch (fmt) {
    case NODEY_FMT_S16: case NODEY_FMT_S16P: return 2;
    case NODEY_FMT_S32: case NODEY_FMT_S32P: case NODEY_FMT_FLT: case NODEY_FMT_FLTP: return 4;
    default: return 0;
    }
}

// ---- device helpers -------------------------------------------------------------------------
// streaming (read-once / write-once) accesses: keep them out of L1
__device__ __forceinline__ float4 ld_stream4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int4 ld_stream4i(const int4* p)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4i(int4* p, int4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// x86 cvttss2si as the reference's integer gain compiles to: truncation, INT_MIN when out of range
__device__ __forceinline__ int x86_trunc(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return __float2int_rz(f);
}

}  // namespace nodey
