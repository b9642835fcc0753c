// stream_kernels.cu -- the HBM-bound streaming nodes: gain (A3), sample extraction (A8), channel
// split (N1), swr format conversion / rematrix, N-input mix (A4), stereo merges (A5, A6) and the
// synthetic source.  One pass over the data, 128-bit accesses, grid = resident CTAs x SM count
// with a grid-stride loop.  Arithmetic is spelled with *_rn intrinsics so the rounding sequence
// is exactly the reference's (separate multiply and add, IEEE division).
#include "nodey_common.cuh"

#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace nodey {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorName(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return NODEY_E_NOMEM;
    return NODEY_E_CUDA;
}

int sm_count()
{
    // cached per device ordinal: a process may render on several GPUs (one Runner per device)
    static std::atomic<int> cached[64];
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64) { const int c = cached[dev].load(std::memory_order_relaxed); if (c) return c; }
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
}

// ---- launch accounting ---------------------------------------------------------------------------
struct LaunchRec { const char* name; cudaEvent_t a, b; double bytes; };
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_profiling{0};
static std::mutex g_prof_mu;
static std::vector<LaunchRec*> g_recs;

LaunchScope::LaunchScope(const char* name, cudaStream_t s, double algo_bytes) : rec(nullptr), st(s)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    LaunchRec* r = new LaunchRec{name, nullptr, nullptr, algo_bytes};
    if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->a, st);
    rec = r;
}

LaunchScope::~LaunchScope()
{
    if (!rec) return;
    LaunchRec* r = (LaunchRec*)rec;
    cudaEventRecord(r->b, st);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_recs.push_back(r);
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

constexpr int kBlock = 256;
constexpr int kCtasPerSm = 8;

// ------------------------------------------------------------------------------------------------
// gain
// ------------------------------------------------------------------------------------------------
// `vec`: bit 0 = both pointers 16-byte aligned (128-bit accesses), bit 1 = in place (dst == src): the non-coherent
// streaming load must not be used on memory this kernel writes, so in-place calls read with plain loads.  No
// __restrict__ on these kernels for the same reason; partially overlapping buffers are rejected by the entry points.
__global__ void __launch_bounds__(kBlock) gain_f32_kernel(float* dst, const float* src, int64_t n, float v, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = (vec & 1) ? n / 4 : 0;
    const bool inplace = (vec & 2) != 0;
    for (int64_t i = tid; i < nvec; i += stride) {
        float4 x = inplace ? reinterpret_cast<const float4*>(src)[i] : ld_stream4(reinterpret_cast<const float4*>(src) + i);
        x.x = __fmul_rn(x.x, v); x.y = __fmul_rn(x.y, v); x.z = __fmul_rn(x.z, v); x.w = __fmul_rn(x.w, v);
        st_stream4(reinterpret_cast<float4*>(dst) + i, x);
    }
    for (int64_t i = nvec * 4 + tid; i < n; i += stride) dst[i] = __fmul_rn(src[i], v);
}

// a batch of float streams in one launch (the per-track audio_volume_adjust nodes of a render): blockIdx.y = stream,
// pointers, lengths and gains ride in the kernel parameters
constexpr int kMaxGainBatch = 256;
struct GainBatch {
    float* dst[kMaxGainBatch];
    const float* src[kMaxGainBatch];
    long long n[kMaxGainBatch];
    float vol[kMaxGainBatch];
    int vec[kMaxGainBatch];
};

__global__ void __launch_bounds__(kBlock) gain_f32_batch_kernel(const __grid_constant__ GainBatch b)
{
    const int t = blockIdx.y;
    float* dst = b.dst[t];
    const float* src = b.src[t];
    const int64_t n = b.n[t];
    const float v = b.vol[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = (b.vec[t] & 1) ? n / 4 : 0;
    const bool inplace = (b.vec[t] & 2) != 0;
    for (int64_t i = tid; i < nvec; i += stride) {
        float4 x = inplace ? reinterpret_cast<const float4*>(src)[i] : ld_stream4(reinterpret_cast<const float4*>(src) + i);
        x.x = __fmul_rn(x.x, v); x.y = __fmul_rn(x.y, v); x.z = __fmul_rn(x.z, v); x.w = __fmul_rn(x.w, v);
        st_stream4(reinterpret_cast<float4*>(dst) + i, x);
    }
    for (int64_t i = nvec * 4 + tid; i < n; i += stride) dst[i] = __fmul_rn(src[i], v);
}

__device__ __forceinline__ short gain_s16_one(short s, float v)
{
    return (short)(unsigned short)(unsigned)x86_trunc(__fmul_rn((float)s, v));
}

__global__ void __launch_bounds__(kBlock) gain_s16_kernel(short* dst, const short* src, int64_t n, float v, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = (vec & 1) ? n / 8 : 0;
    const bool inplace = (vec & 2) != 0;
    for (int64_t i = tid; i < nvec; i += stride) {
        int4 x = inplace ? reinterpret_cast<const int4*>(src)[i] : ld_stream4i(reinterpret_cast<const int4*>(src) + i);
        int* w = reinterpret_cast<int*>(&x);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const short lo = (short)(w[k] & 0xffff), hi = (short)((unsigned)w[k] >> 16);
            const unsigned rl = (unsigned short)gain_s16_one(lo, v), rh = (unsigned short)gain_s16_one(hi, v);
            w[k] = (int)(rl | (rh << 16));
        }
        st_stream4i(reinterpret_cast<int4*>(dst) + i, x);
    }
    for (int64_t i = nvec * 8 + tid; i < n; i += stride) dst[i] = gain_s16_one(src[i], v);
}

__global__ void __launch_bounds__(kBlock) gain_s32_kernel(int* dst, const int* src, int64_t n, float v, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = (vec & 1) ? n / 4 : 0;
    const bool inplace = (vec & 2) != 0;
    for (int64_t i = tid; i < nvec; i += stride) {
        int4 x = inplace ? reinterpret_cast<const int4*>(src)[i] : ld_stream4i(reinterpret_cast<const int4*>(src) + i);
        x.x = x86_trunc(__fmul_rn((float)x.x, v)); x.y = x86_trunc(__fmul_rn((float)x.y, v));
        x.z = x86_trunc(__fmul_rn((float)x.z, v)); x.w = x86_trunc(__fmul_rn((float)x.w, v));
        st_stream4i(reinterpret_cast<int4*>(dst) + i, x);
    }
    for (int64_t i = nvec * 4 + tid; i < n; i += stride) dst[i] = x86_trunc(__fmul_rn((float)src[i], v));
}

// ------------------------------------------------------------------------------------------------
// sample extraction -> interleaved float (A8).  One thread = 4 frames.
// ------------------------------------------------------------------------------------------------
template <int FMT>
__device__ __forceinline__ float extract_one(const void* p, int64_t i)
{
    if (FMT == NODEY_FMT_S16) return __fdiv_rn((float)((const short*)p)[i], 32768.0f);
    if (FMT == NODEY_FMT_S16P) return __fdiv_rn((float)((const short*)p)[i], 32767.0f);
    if (FMT == NODEY_FMT_S32) return __fdiv_rn((float)((const int*)p)[i], 2147483648.0f);
    if (FMT == NODEY_FMT_S32P) return __double2float_rn(__ddiv_rn((double)((const int*)p)[i], 2147483647.0));
    return ((const float*)p)[i];
}

// scalar reference form of the four scales (audio-velocity.cpp:186-218)
template <int FMT>
__device__ __forceinline__ float extract_scale(int v)
{
    // dividing by a power of two is exact, so the S16 / S32 scales are multiplications; 32767 and 2147483647 need
    // the correctly rounded division
    if (FMT == NODEY_FMT_S16) return __fmul_rn((float)v, 3.0517578125e-05f);                 // / 32768.0f
    if (FMT == NODEY_FMT_S16P) return __fdiv_rn((float)v, 32767.0f);
    if (FMT == NODEY_FMT_S32) return __fmul_rn((float)v, 4.656612873077393e-10f);            // / 2147483648.0f
    if (FMT == NODEY_FMT_S32P) return __double2float_rn(__ddiv_rn((double)v, 2147483647.0));
    return __int_as_float(v);
}

// K samples of one 128-bit load -> floats
template <int FMT>
__device__ __forceinline__ void extract_vec(const int4 raw, float (&o)[8])
{
    if (FMT == NODEY_FMT_S16 || FMT == NODEY_FMT_S16P) {
        const int w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            o[2 * k] = extract_scale<FMT>((int)(short)(w[k] & 0xffff));
            o[2 * k + 1] = extract_scale<FMT>(w[k] >> 16);
        }
    } else {
        o[0] = extract_scale<FMT>(raw.x); o[1] = extract_scale<FMT>(raw.y);
        o[2] = extract_scale<FMT>(raw.z); o[3] = extract_scale<FMT>(raw.w);
    }
}

// packed (or mono planar) source: 128-bit loads and stores, grid-stride; VEC = 0: any alignment, one sample per thread
template <int FMT, int VEC>
__global__ void __launch_bounds__(kBlock) extract_packed_kernel(float* __restrict__ dst, const void* __restrict__ src, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int K = (FMT == NODEY_FMT_S16 || FMT == NODEY_FMT_S16P) ? 8 : 4;     // samples per 128-bit load
    const int64_t nvec = VEC ? n / K : 0;
    for (int64_t i = tid; i < nvec; i += stride) {
        float o[8];
        extract_vec<FMT>(ld_stream4i(reinterpret_cast<const int4*>(src) + i), o);
        float4* d = reinterpret_cast<float4*>(dst) + i * (K / 4);
        st_stream4(d, make_float4(o[0], o[1], o[2], o[3]));
        if (K == 8) st_stream4(d + 1, make_float4(o[4], o[5], o[6], o[7]));
    }
    for (int64_t i = nvec * K + tid; i < n; i += stride) dst[i] = extract_one<FMT>(src, i);
}

// planar stereo source -> interleaved: one 128-bit load per plane, 2 or 4 128-bit stores
template <int FMT, int VEC>
__global__ void __launch_bounds__(kBlock) extract_planar_kernel(float* __restrict__ dst, const void* __restrict__ p0,
                                                                const void* __restrict__ p1, int64_t nframes)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int K = FMT == NODEY_FMT_S16P ? 8 : 4;                                 // frames per 128-bit load
    const int64_t nvec = VEC ? nframes / K : 0;
    for (int64_t i = tid; i < nvec; i += stride) {
        float l[8], r[8];
        extract_vec<FMT>(ld_stream4i(reinterpret_cast<const int4*>(p0) + i), l);
        extract_vec<FMT>(ld_stream4i(reinterpret_cast<const int4*>(p1) + i), r);
        float4* d = reinterpret_cast<float4*>(dst) + i * (K / 2);
#pragma unroll
        for (int k = 0; k < K / 2; k++) st_stream4(d + k, make_float4(l[2 * k], r[2 * k], l[2 * k + 1], r[2 * k + 1]));
    }
    for (int64_t i = nvec * K + tid; i < nframes; i += stride) {
        float2 v;
        v.x = extract_one<FMT>(p0, i);
        v.y = extract_one<FMT>(p1, i);
        reinterpret_cast<float2*>(dst)[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// split (N1): packed stereo -> two planes, by 32-bit or 16-bit word
// ------------------------------------------------------------------------------------------------
template <typename W>
__global__ void __launch_bounds__(kBlock) split_kernel(W* __restrict__ l, W* __restrict__ r, const W* __restrict__ src, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        l[i] = src[2 * i];
        r[i] = src[2 * i + 1];
    }
}

// 4 frames per thread, 2x128-bit loads -> 2x128-bit stores (32-bit samples)
__global__ void __launch_bounds__(kBlock) split32_vec_kernel(int* __restrict__ l, int* __restrict__ r, const int* __restrict__ src, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = n / 4;
    for (int64_t i = tid; i < nvec; i += stride) {
        const int4 a = ld_stream4i(reinterpret_cast<const int4*>(src) + 2 * i);
        const int4 b = ld_stream4i(reinterpret_cast<const int4*>(src) + 2 * i + 1);
        st_stream4i(reinterpret_cast<int4*>(l) + i, make_int4(a.x, a.z, b.x, b.z));
        st_stream4i(reinterpret_cast<int4*>(r) + i, make_int4(a.y, a.w, b.y, b.w));
    }
    for (int64_t i = nvec * 4 + tid; i < n; i += stride) { l[i] = src[2 * i]; r[i] = src[2 * i + 1]; }
}

// ------------------------------------------------------------------------------------------------
// swr audioconvert + rematrix, no rate change
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float swr_in_sample(const void* p, int fmt, int64_t i)
{
    switch (fmt) {
    case NODEY_FMT_S16: case NODEY_FMT_S16P: return __fmul_rn((float)((const short*)p)[i], 3.0517578125e-05f);
    case NODEY_FMT_S32: case NODEY_FMT_S32P: return __fmul_rn((float)((const int*)p)[i], 4.656612873077393e-10f);
    default: return ((const float*)p)[i];
    }
}

__global__ void __launch_bounds__(kBlock) to_fltp_kernel(float* __restrict__ dl, float* __restrict__ dr,
                                                         const void* __restrict__ p0, const void* __restrict__ p1,
                                                         int fmt, int planar, int nch, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (nch == 1) {
            const float m = __fmul_rn(swr_in_sample(p0, fmt, i), 0.70710678118654752440f);
            dl[i] = m; dr[i] = m;
        } else if (planar) {
            dl[i] = swr_in_sample(p0, fmt, i);
            dr[i] = swr_in_sample(p1, fmt, i);
        } else {
            dl[i] = swr_in_sample(p0, fmt, 2 * i);
            dr[i] = swr_in_sample(p0, fmt, 2 * i + 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// N-input mix (A4): 4 frames per thread and channel, inputs walked in order
// ------------------------------------------------------------------------------------------------
struct MixArgs {
    const float* l[NODEY_MAX_MIX_INPUTS];
    const float* r[NODEY_MAX_MIX_INPUTS];
    long long len[NODEY_MAX_MIX_INPUTS];
    float vol[NODEY_MAX_MIX_INPUTS];
    float gain[NODEY_MAX_MIX_INPUTS];   // gain of an audio_volume_adjust node folded into this input (1 = none)
    int has_gain;
    int nin;
    int vec;
};

__device__ __forceinline__ float4 load4_zero_tail(const float* p, int64_t j, int64_t len, int vec)
{
    if (vec && j + 4 <= len) return ld_stream4(reinterpret_cast<const float4*>(p + j));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < len) v.x = p[j];
    if (j + 1 < len) v.y = p[j + 1];
    if (j + 2 < len) v.z = p[j + 2];
    if (j + 3 < len) v.w = p[j + 3];
    return v;
}

__device__ __forceinline__ void store4_tail(float* p, int64_t j, int64_t n, float4 v, int vec)
{
    if (vec && j + 4 <= n) { st_stream4(reinterpret_cast<float4*>(p + j), v); return; }
    if (j < n) p[j] = v.x;
    if (j + 1 < n) p[j + 1] = v.y;
    if (j + 2 < n) p[j + 2] = v.z;
    if (j + 3 < n) p[j + 3] = v.w;
}

__device__ __forceinline__ float4 mac4(float4 acc, float4 x, float v)
{
    acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, v));
    acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, v));
    acc.z = __fadd_rn(acc.z, __fmul_rn(x.z, v));
    acc.w = __fadd_rn(acc.w, __fmul_rn(x.w, v));
    return acc;
}

__global__ void __launch_bounds__(kBlock) mix_kernel(float* __restrict__ out_l, float* __restrict__ out_r,
                                                     const __grid_constant__ MixArgs a, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t npack = (n + 3) / 4;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npack; pk += stride) {
        const int64_t j = pk * 4;
        float4 tl = make_float4(0.f, 0.f, 0.f, 0.f), tr = tl;
        for (int i = 0; i < a.nin; i++) {
            float4 xl, xr;
            if (a.r[i]) {
                xl = load4_zero_tail(a.l[i], j, a.len[i], a.vec);
                xr = load4_zero_tail(a.r[i], j, a.len[i], a.vec);
            } else {
                // interleaved stereo input: frames j .. j+3 are 8 consecutive floats (de-interleaved in registers)
                const float4 f0 = load4_zero_tail(a.l[i], 2 * j, 2 * a.len[i], a.vec);
                const float4 f1 = load4_zero_tail(a.l[i], 2 * j + 4, 2 * a.len[i], a.vec);
                xl = make_float4(f0.x, f0.z, f1.x, f1.z);
                xr = make_float4(f0.y, f0.w, f1.y, f1.w);
            }
            if (a.has_gain) {
                // an audio_volume_adjust node in front of this input, folded in: its product x * gain is rounded to float
                // exactly as the node's own pass would have stored it (audio-vol.cpp:75-100), then mixed
                const float g = a.gain[i];
                xl = make_float4(__fmul_rn(xl.x, g), __fmul_rn(xl.y, g), __fmul_rn(xl.z, g), __fmul_rn(xl.w, g));
                xr = make_float4(__fmul_rn(xr.x, g), __fmul_rn(xr.y, g), __fmul_rn(xr.z, g), __fmul_rn(xr.w, g));
            }
            tl = mac4(tl, xl, a.vol[i]);
            tr = mac4(tr, xr, a.vol[i]);
        }
        store4_tail(out_l, j, n, tl, a.vec);
        store4_tail(out_r, j, n, tr, a.vec);
    }
}

// ------------------------------------------------------------------------------------------------
// bimix v1 (A5)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bimix_one(float a, float b, float k)
{
    return __fmul_rn(__fadd_rn(__fdiv_rn(a, 2.0f), __fdiv_rn(b, 2.0f)), k);
}

__global__ void __launch_bounds__(kBlock) bimix_kernel(float* __restrict__ out_l, float* __restrict__ out_r,
                                                       const float* __restrict__ ll, const float* __restrict__ lr, int64_t len_l,
                                                       const float* __restrict__ rl, const float* __restrict__ rr, int64_t len_r,
                                                       float bias_minus, float bias_plus, int64_t n, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t npack = (n + 3) / 4;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npack; pk += stride) {
        const int64_t j = pk * 4;
        const float4 a = load4_zero_tail(ll, j, len_l, vec), b = load4_zero_tail(lr, j, len_l, vec);
        const float4 c = load4_zero_tail(rl, j, len_r, vec), d = load4_zero_tail(rr, j, len_r, vec);
        float4 ol, orr;
        ol.x = bimix_one(a.x, b.x, bias_minus); ol.y = bimix_one(a.y, b.y, bias_minus);
        ol.z = bimix_one(a.z, b.z, bias_minus); ol.w = bimix_one(a.w, b.w, bias_minus);
        orr.x = bimix_one(c.x, d.x, bias_plus); orr.y = bimix_one(c.y, d.y, bias_plus);
        orr.z = bimix_one(c.z, d.z, bias_plus); orr.w = bimix_one(c.w, d.w, bias_plus);
        store4_tail(out_l, j, n, ol, vec);
        store4_tail(out_r, j, n, orr, vec);
    }
}

// ------------------------------------------------------------------------------------------------
// bimix v2 pieces (A6)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) downmix_half_kernel(float* __restrict__ dst, const float* __restrict__ l,
                                                              const float* __restrict__ r, int64_t n, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t npack = (n + 3) / 4;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npack; pk += stride) {
        const int64_t j = pk * 4;
        const float4 a = load4_zero_tail(l, j, n, vec), b = load4_zero_tail(r, j, n, vec);
        float4 o;
        o.x = __fmul_rn(__fadd_rn(a.x, b.x), 0.5f); o.y = __fmul_rn(__fadd_rn(a.y, b.y), 0.5f);
        o.z = __fmul_rn(__fadd_rn(a.z, b.z), 0.5f); o.w = __fmul_rn(__fadd_rn(a.w, b.w), 0.5f);
        store4_tail(dst, j, n, o, vec);
    }
}

// preview sink (audio-io.cpp:598-599): 48 kHz stereo planes -> packed frames, every sample clamped to [-1, 1]
// with std::clamp's comparisons (a NaN passes through)
__device__ __forceinline__ float clamp_unit(float v) { return v < -1.0f ? -1.0f : (1.0f < v ? 1.0f : v); }

__global__ void __launch_bounds__(kBlock) pack_clamp_kernel(float* __restrict__ dst, const float* __restrict__ l,
                                                            const float* __restrict__ r, int64_t n, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t npack = (n + 3) / 4;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npack; pk += stride) {
        const int64_t j = pk * 4;
        const float4 a = load4_zero_tail(l, j, n, vec), b = load4_zero_tail(r, j, n, vec);
        const float4 o0 = make_float4(clamp_unit(a.x), clamp_unit(b.x), clamp_unit(a.y), clamp_unit(b.y));
        const float4 o1 = make_float4(clamp_unit(a.z), clamp_unit(b.z), clamp_unit(a.w), clamp_unit(b.w));
        store4_tail(dst, 2 * j, 2 * n, o0, vec);
        store4_tail(dst, 2 * j + 4, 2 * n, o1, vec);
    }
}

struct MergeSeg { long long out_start, len, l, r; };

__global__ void __launch_bounds__(kBlock) merge_segments_kernel(float2* __restrict__ out, const float* __restrict__ left,
                                                                const float* __restrict__ right,
                                                                const MergeSeg* __restrict__ segs, int nseg, int64_t total)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
        int lo = 0, hi = nseg - 1;       // last segment with out_start <= j
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (segs[mid].out_start <= j) lo = mid; else hi = mid - 1;
        }
        const MergeSeg s = segs[lo];
        const int64_t d = j - s.out_start;
        float2 v;
        v.x = s.l >= 0 ? left[s.l + d] : 0.0f;
        v.y = s.r >= 0 ? right[s.r + d] : 0.0f;
        out[j] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// synthetic source (bit-identical to oracle/nodey_oracle.c orc_synth_f32)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned synth_hash(unsigned seed, unsigned long long n)
{
    unsigned long long z = n + ((unsigned long long)seed << 32) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (unsigned)(z >> 32);
}

__device__ __forceinline__ float synth_sin_turns(float t)
{
    if (t > 0.25f) t = __fsub_rn(0.5f, t);
    else if (t < -0.25f) t = __fsub_rn(-0.5f, t);
    const float y = __fmul_rn(6.28318530717958647692f, t);
    const float y2 = __fmul_rn(y, y);
    float p = 2.7557319e-6f;
    p = __fmaf_rn(p, y2, -1.9841270e-4f);
    p = __fmaf_rn(p, y2, 8.3333333e-3f);
    p = __fmaf_rn(p, y2, -1.6666667e-1f);
    p = __fmaf_rn(p, y2, 1.0f);
    return __fmul_rn(y, p);
}

__global__ void __launch_bounds__(kBlock) synth_kernel(float* __restrict__ f32, short* __restrict__ s16, int64_t nframes,
                                                       int nch, unsigned step0, unsigned step1, unsigned seed0, int64_t frame0)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = nframes * nch;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int c = (int)(e % nch);
        const unsigned long long n = (unsigned long long)(frame0 + e / nch);
        const unsigned step = c ? step1 : step0;
        const unsigned ph = (unsigned)(n * (unsigned long long)step);
        const float t = __fmul_rn(__int2float_rn((int)ph), 2.3283064365386963e-10f);
        const float s = synth_sin_turns(t);
        const float u = __fsub_rn(__fmul_rn(__uint2float_rn(synth_hash(seed0 + (unsigned)c, n) >> 8), 1.1920928955078125e-7f), 1.0f);
        const float x = __fmaf_rn(0.05f, u, __fmul_rn(0.5f, s));
        if (f32) f32[e] = x;
        if (s16) {
            int v = __float2int_rn(__fmul_rn(x, 32767.0f));
            v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
            s16[e] = (short)v;
        }
    }
}

}  // namespace nodey

using namespace nodey;

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int nodey_version(void) { return 100; }

const char* nodey_last_error(void) { return g_err; }

void nodey_profile_enable(int on) { g_profiling.store(on ? 1 : 0); }

uint64_t nodey_profile_launches(void) { return g_launches.load(); }

/* JSON object {"kernel": {"launches": n, "ms": total device ms, "bytes": algorithmic bytes}, ...} of every
 * launch recorded since the last report; synchronises the device.  Returns the length written. */
int nodey_profile_report(char* buf, int cap)
{
    cudaDeviceSynchronize();
    std::vector<LaunchRec*> recs;
    { std::lock_guard<std::mutex> lock(g_prof_mu); recs.swap(g_recs); }
    struct Agg { unsigned long long n = 0; double ms = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    for (LaunchRec* r : recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) { Agg& g = agg[r->name]; g.n++; g.ms += ms; g.bytes += r->bytes; }
        cudaEventDestroy(r->a); cudaEventDestroy(r->b);
        delete r;
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f, \"bytes\": %.0f}", first ? "" : ", ",
                 kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes);
        out += line;
        first = false;
    }
    out += "}";
    if (buf && cap > 0) { snprintf(buf, (size_t)cap, "%s", out.c_str()); }
    return (int)out.size();
}

int nodey_device_info(int* sms, int* major, int* minor, int64_t* total_mem)
{
    int dev = 0;
    NODEY_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    NODEY_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
    return NODEY_OK;
}

int nodey_synth(float* dst_f32, int16_t* dst_s16, int64_t nframes, int nch, int sample_rate, int track,
                int64_t frame0, nodey_stream_t stream)
{
    NODEY_REQUIRE(nch == 1 || nch == 2, NODEY_E_INVALID, "nodey_synth: nch must be 1 or 2 (got %d)", nch);
    NODEY_REQUIRE(sample_rate > 0 && nframes >= 0, NODEY_E_INVALID, "nodey_synth: bad size");
    if (nframes == 0) return NODEY_OK;
    unsigned step[2];
    for (int c = 0; c < 2; c++) {
        const int e = (7 * track + 4 * c) % 36;
        const double f = 220.0 * pow(2.0, (double)e / 12.0);
        step[c] = (unsigned)llrint(f / (double)sample_rate * 4294967296.0);
    }
    const unsigned seed0 = 0xA0D10u + 131u * (unsigned)track;
    NODEY_LAUNCH("synth_kernel", as_stream(stream), synth_kernel<<<stream_grid(nframes * nch, kBlock, kCtasPerSm), kBlock, 0, as_stream(stream)>>>(
        dst_f32, dst_s16, nframes, nch, step[0], step[1], seed0, frame0));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_gain(void* dst, const void* src, int fmt, int64_t n, float volume, nodey_stream_t stream)
{
    NODEY_REQUIRE(n >= 0 && (n == 0 || (dst && src)), NODEY_E_INVALID, "nodey_gain: null buffer or negative size");
    if (n == 0) return NODEY_OK;
    {
        // in place (dst == src) is allowed; any other overlap would let one thread read what another already wrote
        const size_t bytes = (size_t)n * (size_t)(fmt_bytes(fmt) ? fmt_bytes(fmt) : 1);
        const char* a = (const char*)dst; const char* b = (const char*)src;
        NODEY_REQUIRE(a == b || a + bytes <= b || b + bytes <= a, NODEY_E_INVALID, "nodey_gain: dst and src overlap partially");
    }
    const int vec = ((aligned16(dst) && aligned16(src)) ? 1 : 0) | (dst == src ? 2 : 0);
    cudaStream_t st = as_stream(stream);
    switch (fmt) {
    case NODEY_FMT_FLT: case NODEY_FMT_FLTP:
        NODEY_LAUNCH("gain_f32_kernel", st, gain_f32_kernel<<<stream_grid(n / 4 + 1, kBlock, kCtasPerSm), kBlock, 0, st>>>((float*)dst, (const float*)src, n, volume, vec));
        break;
    case NODEY_FMT_S16: case NODEY_FMT_S16P:
        NODEY_LAUNCH("gain_s16_kernel", st, gain_s16_kernel<<<stream_grid(n / 8 + 1, kBlock, kCtasPerSm), kBlock, 0, st>>>((short*)dst, (const short*)src, n, volume, vec));
        break;
    case NODEY_FMT_S32: case NODEY_FMT_S32P:
        NODEY_LAUNCH("gain_s32_kernel", st, gain_s32_kernel<<<stream_grid(n / 4 + 1, kBlock, kCtasPerSm), kBlock, 0, st>>>((int*)dst, (const int*)src, n, volume, vec));
        break;
    default:
        set_error("Audio format is not support (Include FLT, S16, S32): %d", fmt);
        return NODEY_E_FORMAT;
    }
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_gain_tracks(void* const* dst, const void* const* src, const int64_t* n, const float* volumes, int fmt, int ntracks,
                      nodey_stream_t stream)
{
    NODEY_REQUIRE(dst && src && n && volumes && ntracks >= 0, NODEY_E_INVALID, "nodey_gain_tracks: null argument");
    NODEY_REQUIRE(fmt == NODEY_FMT_FLT || fmt == NODEY_FMT_FLTP, NODEY_E_FORMAT, "nodey_gain_tracks: float streams only (format %d)", fmt);
    cudaStream_t st = as_stream(stream);
    static thread_local GainBatch b;
    for (int first = 0; first < ntracks; first += kMaxGainBatch) {
        const int cnt = ntracks - first < kMaxGainBatch ? ntracks - first : kMaxGainBatch;
        int64_t longest = 0;
        for (int t = 0; t < cnt; t++) {
            NODEY_REQUIRE(n[first + t] >= 0 && (n[first + t] == 0 || (dst[first + t] && src[first + t])), NODEY_E_INVALID, "nodey_gain_tracks: null buffer or negative size");
            b.dst[t] = (float*)dst[first + t]; b.src[t] = (const float*)src[first + t]; b.n[t] = n[first + t]; b.vol[t] = volumes[first + t];
            {
                const char* pa = (const char*)dst[first + t]; const char* pb = (const char*)src[first + t];
                const size_t bytes = (size_t)n[first + t] * sizeof(float);
                NODEY_REQUIRE(pa == pb || pa + bytes <= pb || pb + bytes <= pa, NODEY_E_INVALID, "nodey_gain_tracks: dst and src of stream %d overlap partially", first + t);
            }
            b.vec[t] = ((aligned16(dst[first + t]) && aligned16(src[first + t])) ? 1 : 0) | (dst[first + t] == src[first + t] ? 2 : 0);
            if (n[first + t] > longest) longest = n[first + t];
        }
        if (longest == 0) continue;
        // about kCtasPerSm CTAs per SM in total, at least one per stream
        int64_t gx = ((int64_t)sm_count() * kCtasPerSm + cnt - 1) / cnt;
        const int64_t need = (longest / 4 + kBlock) / kBlock;
        if (gx > need) gx = need;
        if (gx < 1) gx = 1;
        NODEY_LAUNCH("gain_f32_kernel", st, gain_f32_batch_kernel<<<dim3((unsigned)gx, (unsigned)cnt), kBlock, 0, st>>>(b));
        NODEY_LAUNCH_OK();
    }
    return NODEY_OK;
}

int nodey_extract_interleaved(float* dst, const void* p0, const void* p1, int fmt, int64_t nframes, int nch,
                              nodey_stream_t stream)
{
    NODEY_REQUIRE(nch == 1 || nch == 2, NODEY_E_INVALID, "nodey_extract_interleaved: nch must be 1 or 2");
    NODEY_REQUIRE(nframes >= 0, NODEY_E_INVALID, "nodey_extract_interleaved: negative size");
    if (nframes == 0) return NODEY_OK;
    cudaStream_t st = as_stream(stream);
    const int64_t n = nframes * nch;
    const bool vec = aligned16(dst) && aligned16(p0) && (nch == 1 || fmt < NODEY_FMT_U8P || aligned16(p1));
    const int g = stream_grid((n + 7) / 8, kBlock, kCtasPerSm);
#define NODEY_EXTRACT_PACKED(F) do { if (vec) NODEY_LAUNCH("extract_packed_kernel", st, extract_packed_kernel<F, 1><<<g, kBlock, 0, st>>>(dst, p0, n)); \
                                     else NODEY_LAUNCH("extract_packed_kernel", st, extract_packed_kernel<F, 0><<<stream_grid(n, kBlock, kCtasPerSm), kBlock, 0, st>>>(dst, p0, n)); } while (0)
#define NODEY_EXTRACT_PLANAR(F) do { if (nch == 1) NODEY_EXTRACT_PACKED(F); \
                                     else if (vec) NODEY_LAUNCH("extract_planar_kernel", st, extract_planar_kernel<F, 1><<<g, kBlock, 0, st>>>(dst, p0, p1, nframes)); \
                                     else NODEY_LAUNCH("extract_planar_kernel", st, extract_planar_kernel<F, 0><<<stream_grid(nframes, kBlock, kCtasPerSm), kBlock, 0, st>>>(dst, p0, p1, nframes)); } while (0)
    switch (fmt) {
    case NODEY_FMT_FLT:
        NODEY_CUDA_OK(cudaMemcpyAsync(dst, p0, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        return NODEY_OK;
    case NODEY_FMT_S16: NODEY_EXTRACT_PACKED(NODEY_FMT_S16); break;
    case NODEY_FMT_S32: NODEY_EXTRACT_PACKED(NODEY_FMT_S32); break;
    case NODEY_FMT_FLTP: NODEY_EXTRACT_PLANAR(NODEY_FMT_FLTP); break;
    case NODEY_FMT_S16P: NODEY_EXTRACT_PLANAR(NODEY_FMT_S16P); break;
    case NODEY_FMT_S32P: NODEY_EXTRACT_PLANAR(NODEY_FMT_S32P); break;
    default:
        set_error("Unsupported sample format: %d", fmt);
        return NODEY_E_FORMAT;
    }
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_split(void* dl, void* dr, const void* p0, const void* p1, int fmt, int64_t n, nodey_stream_t stream)
{
    NODEY_REQUIRE(n >= 0, NODEY_E_INVALID, "nodey_split: negative size");
    if (n == 0) return NODEY_OK;
    cudaStream_t st = as_stream(stream);
    switch (fmt) {
    case NODEY_FMT_FLT: case NODEY_FMT_S32:
        if (aligned16(dl) && aligned16(dr) && aligned16(p0))
            NODEY_LAUNCH("split32_vec_kernel", st, split32_vec_kernel<<<stream_grid(n / 4 + 1, kBlock, kCtasPerSm), kBlock, 0, st>>>((int*)dl, (int*)dr, (const int*)p0, n));
        else
            NODEY_LAUNCH("split_kernel", st, split_kernel<int><<<stream_grid(n, kBlock, kCtasPerSm), kBlock, 0, st>>>((int*)dl, (int*)dr, (const int*)p0, n));
        break;
    case NODEY_FMT_S16:
        NODEY_LAUNCH("split_kernel", st, split_kernel<short><<<stream_grid(n, kBlock, kCtasPerSm), kBlock, 0, st>>>((short*)dl, (short*)dr, (const short*)p0, n));
        break;
    case NODEY_FMT_FLTP: case NODEY_FMT_S32P: case NODEY_FMT_S16P: {
        const size_t bytes = (size_t)n * (size_t)fmt_bytes(fmt);
        NODEY_CUDA_OK(cudaMemcpyAsync(dl, p0, bytes, cudaMemcpyDeviceToDevice, st));
        NODEY_CUDA_OK(cudaMemcpyAsync(dr, p1, bytes, cudaMemcpyDeviceToDevice, st));
        return NODEY_OK;
    }
    default:
        set_error("nodey_split: unsupported sample format %d", fmt);
        return NODEY_E_FORMAT;
    }
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_to_fltp_stereo(float* dl, float* dr, const void* p0, const void* p1, int fmt, int nch, int64_t n,
                         nodey_stream_t stream)
{
    NODEY_REQUIRE(nch == 1 || nch == 2, NODEY_E_INVALID, "Invalid channel layout: %d", nch);
    NODEY_REQUIRE(fmt_bytes(fmt) != 0, NODEY_E_FORMAT, "nodey_to_fltp_stereo: unsupported sample format %d", fmt);
    if (n <= 0) return n == 0 ? NODEY_OK : NODEY_E_INVALID;
    NODEY_LAUNCH("to_fltp_kernel", as_stream(stream), to_fltp_kernel<<<stream_grid(n, kBlock, kCtasPerSm), kBlock, 0, as_stream(stream)>>>(dl, dr, p0, p1, fmt,
                                                                                       fmt_planar(fmt) ? 1 : 0, nch, n));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_mix(float* out_l, float* out_r, const float* const* in_l, const float* const* in_r, const int64_t* in_len,
              const float* volumes, int nin, int64_t n, nodey_stream_t stream)
{
    return nodey_mix_gains(out_l, out_r, in_l, in_r, in_len, volumes, nullptr, nin, n, stream);
}

int nodey_mix_gains(float* out_l, float* out_r, const float* const* in_l, const float* const* in_r, const int64_t* in_len,
                    const float* volumes, const float* gains, int nin, int64_t n, nodey_stream_t stream)
{
    NODEY_REQUIRE(nin >= 1 && nin <= NODEY_MAX_MIX_INPUTS, NODEY_E_RANGE, "nodey_mix: input_num %d outside 1..16", nin);
    if (n <= 0) return n == 0 ? NODEY_OK : NODEY_E_INVALID;
    MixArgs a;
    memset(&a, 0, sizeof(a));
    a.nin = nin;
    a.vec = aligned16(out_l) && aligned16(out_r);
    a.has_gain = gains != nullptr;
    for (int i = 0; i < nin; i++) {
        a.l[i] = in_l[i]; a.r[i] = in_r[i]; a.len[i] = in_len[i]; a.vol[i] = volumes[i];
        a.gain[i] = gains ? gains[i] : 1.0f;
        a.vec = a.vec && aligned16(in_l[i]) && (in_r[i] == nullptr || aligned16(in_r[i]));
    }
    NODEY_LAUNCH("mix_kernel", as_stream(stream), mix_kernel<<<stream_grid((n + 3) / 4, kBlock, kCtasPerSm), kBlock, 0, as_stream(stream)>>>(out_l, out_r, a, n));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_bimix(float* out_l, float* out_r, const float* ll, const float* lr, int64_t len_l, const float* rl,
                const float* rr, int64_t len_r, float bias, int64_t n, nodey_stream_t stream)
{
    if (n <= 0) return n == 0 ? NODEY_OK : NODEY_E_INVALID;
    const int vec = aligned16(out_l) && aligned16(out_r) && aligned16(ll) && aligned16(lr) && aligned16(rl) && aligned16(rr);
    const float bias_minus = 1 - bias, bias_plus = 1 + bias;
    NODEY_LAUNCH("bimix_kernel", as_stream(stream), bimix_kernel<<<stream_grid((n + 3) / 4, kBlock, kCtasPerSm), kBlock, 0, as_stream(stream)>>>(
        out_l, out_r, ll, lr, len_l, rl, rr, len_r, bias_minus, bias_plus, n, vec));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_downmix_half(float* dst, const float* l, const float* r, int64_t n, nodey_stream_t stream)
{
    if (n <= 0) return n == 0 ? NODEY_OK : NODEY_E_INVALID;
    const int vec = aligned16(dst) && aligned16(l) && aligned16(r);
    NODEY_LAUNCH("downmix_half_kernel", as_stream(stream), downmix_half_kernel<<<stream_grid((n + 3) / 4, kBlock, kCtasPerSm), kBlock, 0, as_stream(stream)>>>(dst, l, r, n, vec));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_preview_pack(float* dst, const float* l, const float* r, int64_t n, nodey_stream_t stream)
{
    if (n <= 0) return n == 0 ? NODEY_OK : NODEY_E_INVALID;
    NODEY_REQUIRE(dst && l && r, NODEY_E_INVALID, "nodey_preview_pack: null buffer");
    const int vec = aligned16(dst) && aligned16(l) && aligned16(r);
    NODEY_LAUNCH("pack_clamp_kernel", as_stream(stream), pack_clamp_kernel<<<stream_grid((n + 3) / 4, kBlock, kCtasPerSm), kBlock, 0, as_stream(stream)>>>(dst, l, r, n, vec));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_merge_segments(float* out, const float* left, const float* right, const int64_t* seg_out_start,
                         const int64_t* seg_len, const int64_t* seg_l, const int64_t* seg_r, int nseg,
                         nodey_stream_t stream)
{
    NODEY_REQUIRE(nseg >= 0, NODEY_E_INVALID, "nodey_merge_segments: negative segment count");
    if (nseg == 0) return NODEY_OK;
    MergeSeg* h = (MergeSeg*)malloc(sizeof(MergeSeg) * (size_t)nseg);
    NODEY_REQUIRE(h, NODEY_E_NOMEM, "nodey_merge_segments: host allocation failed");
    int64_t total = 0;
    for (int i = 0; i < nseg; i++) {
        h[i].out_start = seg_out_start[i]; h[i].len = seg_len[i]; h[i].l = seg_l[i]; h[i].r = seg_r[i];
        if (seg_out_start[i] + seg_len[i] > total) total = seg_out_start[i] + seg_len[i];
    }
    cudaStream_t st = as_stream(stream);
    MergeSeg* d = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&d, sizeof(MergeSeg) * (size_t)nseg, st);
    if (e != cudaSuccess) { free(h); return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__); }
    e = cudaMemcpyAsync(d, h, sizeof(MergeSeg) * (size_t)nseg, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // h is pageable: keep it alive until the copy landed
    free(h);
    if (e != cudaSuccess) { cudaFreeAsync(d, st); return cuda_fail(e, "segment upload", __FILE__, __LINE__); }
    if (total > 0)
        NODEY_LAUNCH("merge_segments_kernel", st, merge_segments_kernel<<<stream_grid(total, kBlock, kCtasPerSm), kBlock, 0, st>>>((float2*)out, left, right, d, nseg, total));
    e = cudaPeekAtLastError();
    cudaFreeAsync(d, st);
    if (e != cudaSuccess) return cuda_fail(e, "merge_segments_kernel", __FILE__, __LINE__);
    return NODEY_OK;
}

}  // extern "C"
