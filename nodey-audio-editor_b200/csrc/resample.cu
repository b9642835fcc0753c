// resample.cu -- K5: libswresample-equivalent polyphase FIR (A7; call sites audio-amix.cpp:263-290,
// audio-bimix.cpp:259-294, sw-resample.hpp:63-69) and its fusion with the N-input mix (A4).
//
// Output k sits at phase position pos_k = index0 + k*dst_incr_div (+ carry of k*dst_incr_mod) in
// units of 1/phase_count input samples:  y[k] = sum_i x[pos_k / P - center + i] * h[pos_k % P][i],
// x mirrored about sample 0 on the left and reflected at the end after a flush.  The accumulation
// order is the oracle's: one accumulator, ascending taps, fused multiply-add.
//
// Three kernels:
//  * resample_generic_kernel: one thread per output frame, any plan (incl. the interpolating
//    path when the ratio is not exact), taps and samples straight from global/L1.  Correctness
//    baseline and fallback.
//  * resample_tile2_kernel: the production path for exact-rational plans with at most 160 phases
//    (one phase group per warp; 44.1 -> 48 kHz: 160 phases, 20 groups).  A CTA owns a tile of 32 or 64
//    periods (period = P outputs <-> D input frames).  Tap table and packed-float input tiles are
//    staged by the TMA engine (cp.async.bulk + mbarrier), the tiles double buffered; warp w owns a
//    group of G=8 consecutive phases, lane b owns period b (and b + 32).  A thread keeps G x channels
//    (x periods) accumulators in registers and walks the union window of its group once: one
//    shared-memory sample feeds G FMAs, and the G taps of a window position are a warp-uniform
//    broadcast read (dense per-group tap matrix, zero where a phase's window does not reach).  The
//    accumulators survive across the <= 16 inputs of a mix (audio_amix fused; config 3) and leave as
//    128-bit stores; one launch covers a batch of tracks.
//  * resample_tile_kernel: the same arithmetic with shared staging rows and a copy-out pass, for
//    plans with more phase groups than warps (e.g. 22.05 -> 48 kHz: 320 phases).
#include "nodey_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct nodey_resampler {
    int in_rate, out_rate;
    int resample;                       // 0 when rates are equal
    int phase_count, filter_length, filter_alloc;
    int src_incr, dst_incr, dst_incr_div, dst_incr_mod;
    int index0;
    std::vector<float> bank;            // host copy, (phase_count + 1) * filter_alloc
    float* d_bank = nullptr;            // device copy
    // tile kernel tables (exact-rational plans only)
    int tile_ok = 0;
    int G = 8, n_groups = 0, wmax = 0, s0 = 0, span = 0;   // span = s_{P-1} - s_0
    float* d_hq = nullptr;              // [n_groups][wmax][G]
    int* d_group_start = nullptr;       // [n_groups] : s_{qG} - s_0
    int device = 0;
};

namespace nodey {

static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a < 0 ? -a : a; }

static double bessel_i0(double x)
{
    double v = 1, lastv = 0, t = 1;
    x = x * x / 4;
    for (int i = 1; v != lastv; i++) {
        lastv = v;
        t *= x / ((double)i * (double)i);
        v += t;
    }
    return v;
}

// Kaiser-windowed sinc as libswresample's build_filter() stores it for FLTP (libswresample/resample.c).
// The tap sum is taken for phase 0 only and divides every phase (so only phase 0 has exactly unit DC gain),
// and with factor == 1 (upsampling) sin(x) comes from a per-phase value with alternating sign.  Both were
// found by comparing impulse responses with the real library (tests/test_swr_real.py: bit identical).
static void build_filter(float* bank, double factor, int tap_count, int alloc, int phase_count, double beta)
{
    const int center = (tap_count - 1) / 2;
    std::vector<double> tab((size_t)tap_count);
    const int ph_nb = (phase_count % 2) ? phase_count : phase_count / 2 + 1;
    double norm = 0;
    if (factor > 1.0) factor = 1.0;
    for (int ph = 0; ph < ph_nb; ph++) {
        double s = (factor == 1.0) ? sin(M_PI * ph / phase_count) * ((center & 1) ? 1 : -1) : 0;
        for (int i = 0; i < tap_count; i++) {
            const double x = M_PI * ((double)(i - center) - (double)ph / phase_count) * factor;
            double y;
            if (x == 0) y = 1.0;
            else if (factor == 1.0) { y = s / x; s = -s; }
            else y = sin(x) / x;
            const double w = 2.0 * x / (factor * tap_count * M_PI);
            const double a = 1 - w * w;
            y *= bessel_i0(beta * sqrt(a > 0 ? a : 0));
            tab[(size_t)i] = y;
            if (!ph) norm += y;
        }
        for (int i = 0; i < tap_count; i++) bank[ph * alloc + i] = (float)(tab[(size_t)i] * 1.0 / norm);
        if (phase_count % 2) continue;
        for (int i = 0; i < tap_count; i++)
            bank[(phase_count - ph) * alloc + tap_count - 1 - i] = bank[ph * alloc + i];
    }
}

// ---- source access with on-the-fly swr input conversion ------------------------------------------
struct SrcDesc {
    const void* p0;
    const void* p1;
    long long n;          // real input frames
    long long reflect;    // reflected frames appended by the flush
    int fmt, nch, planar;
};

__device__ __forceinline__ float src_scalar(const void* p, int fmt, long long i)
{
    switch (fmt) {
    case NODEY_FMT_S16: case NODEY_FMT_S16P: return __fmul_rn((float)((const short*)p)[i], 3.0517578125e-05f);
    case NODEY_FMT_S32: case NODEY_FMT_S32P: return __fmul_rn((float)((const int*)p)[i], 4.656612873077393e-10f);
    default: return ((const float*)p)[i];
    }
}

// extended-signal frame f (mirror / reflection applied) as a stereo pair after rematrix
__device__ __forceinline__ float2 src_frame(const SrcDesc& s, long long f)
{
    if (f < 0) f = -f;
    if (f >= s.n) f = 2 * s.n - 1 - f;
    if (f < 0 || f >= s.n) return make_float2(0.f, 0.f);
    if (s.nch == 1) {
        const float m = __fmul_rn(src_scalar(s.p0, s.fmt, f), 0.70710678118654752440f);
        return make_float2(m, m);
    }
    if (s.planar) return make_float2(src_scalar(s.p0, s.fmt, f), src_scalar(s.p1, s.fmt, f));
    if (s.fmt == NODEY_FMT_FLT) return reinterpret_cast<const float2*>(s.p0)[f];
    return make_float2(src_scalar(s.p0, s.fmt, 2 * f), src_scalar(s.p0, s.fmt, 2 * f + 1));
}

struct PlanDev {
    const float* bank;
    int P, L, alloc, div, mod, src_incr, index0;
};

// ---- generic kernel ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resample_generic_kernel(float* __restrict__ out_l, float* __restrict__ out_r,
                                                               const SrcDesc s, const PlanDev pl, long long out_frames)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int center = (pl.L - 1) / 2;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < out_frames; k += stride) {
        const long long f = k * (long long)pl.mod;
        const long long index = (long long)pl.index0 + k * (long long)pl.div + f / pl.src_incr;
        const int frac = (int)(f % pl.src_incr);
        const long long start = index / pl.P - center;
        const float* taps = pl.bank + (size_t)pl.alloc * (size_t)(index % pl.P);
        float al = 0.f, ar = 0.f, bl = 0.f, br = 0.f;
        for (int i = 0; i < pl.L; i++) {
            const float2 x = src_frame(s, start + i);
            const float h = __ldg(taps + i);
            al = __fmaf_rn(x.x, h, al);
            ar = __fmaf_rn(x.y, h, ar);
            if (pl.mod) {
                const float h2 = __ldg(taps + pl.alloc + i);
                bl = __fmaf_rn(x.x, h2, bl);
                br = __fmaf_rn(x.y, h2, br);
            }
        }
        if (pl.mod) {
            // resample_linear: val += (v2 - val) * (float)frac / src_incr
            al = __fadd_rn(al, __fdiv_rn(__fmul_rn(__fsub_rn(bl, al), (float)frac), (float)pl.src_incr));
            ar = __fadd_rn(ar, __fdiv_rn(__fmul_rn(__fsub_rn(br, ar), (float)frac), (float)pl.src_incr));
        }
        out_l[k] = al;
        out_r[k] = s.nch == 1 ? al : ar;
    }
}

// ---- tile kernel --------------------------------------------------------------------------------------
constexpr int kG = 8;          // phases per warp group
constexpr int kNB = 32;        // periods per tile = lanes

struct TileArgs {
    SrcDesc src[NODEY_MAX_MIX_INPUTS];
    long long out_len[NODEY_MAX_MIX_INPUTS];   // frames this input contributes (zeros after)
    float vol[NODEY_MAX_MIX_INPUTS];
    int nin;
    int mix;                 // 0: plain resample of src[0] (no volume multiply)
    const float* hq;         // [n_groups][wmax][G]
    const int* group_start;  // [n_groups]
    int P, D, L, center, n_groups, wmax, s0, span;
    int in_tile;             // frames of input staged per tile
    int out_stride;          // padded staging row stride (odd)
    long long out_frames;
    long long n_tiles;
    int vec_out;             // 128-bit stores allowed: planes 16-byte aligned and P a multiple of 4
    int ntracks;             // pipelined kernel: tracks of the batch (per-track planes in TrackPlanes), 1 otherwise
    long long out_track_stride;   // floats between the output planes of consecutive tracks
    long long tile0, tile1;  // pipelined kernel: this launch covers tiles [tile0, tile1) of every track (tile1 = 0: all) --
                             // a render cut along time, so that the next node can start on a prefix (nodey_resample_tracks_chunk)
};

// fast staging test: the whole tile lies inside the real input (no mirror / reflection / zero fill)
__device__ __forceinline__ bool tile_interior(const SrcDesc& s, long long in0, int frames)
{
    return in0 >= 0 && in0 + frames <= s.n;
}

template <int CH>
__global__ void __launch_bounds__(640, 2) resample_tile_kernel(float* __restrict__ out_l, float* __restrict__ out_r,
                                                            const __grid_constant__ TileArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [hq: n_groups*wmax*G floats][group_start: n_groups ints][stage L: NB*out_stride][stage R][input tile]
    float* s_hq = reinterpret_cast<float*>(smem_raw);
    int* s_gs = reinterpret_cast<int*>(s_hq + a.n_groups * a.wmax * kG);
    float* s_out_l = reinterpret_cast<float*>(s_gs + ((a.n_groups + 3) & ~3));
    float* s_out_r = s_out_l + kNB * a.out_stride;
    float* s_in = s_out_r + kNB * a.out_stride;     // CH floats per frame, frame-major

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;

    for (int i = tid; i < a.n_groups * a.wmax * kG; i += nthr) s_hq[i] = a.hq[i];
    for (int i = tid; i < a.n_groups; i += nthr) s_gs[i] = a.group_start[i];

    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const long long k0 = tile * (long long)kNB * a.P;               // first output frame of the tile
        const long long in0 = tile * (long long)kNB * a.D + a.s0 - a.center;  // first staged input frame
        // direct mode (a single input, one group per warp -- e.g. the per-track 44.1 -> 48 kHz audio_amix(1): 20 groups
        // on 20 warps): a thread's 8 outputs leave its registers as 128-bit stores; no staging rows, no copy-out pass
        const bool direct = a.nin == 1 && nwarps >= a.n_groups;

        for (int inp = 0; inp < a.nin; inp++) {
            __syncthreads();   // previous use of s_in (and of the staging rows on a new tile) is over
            const SrcDesc& s = a.src[inp];
            if (k0 < a.out_len[inp]) {
                if (CH == 2) {
                    float2* dst = reinterpret_cast<float2*>(s_in);
                    if (s.fmt == NODEY_FMT_FLT && tile_interior(s, in0, a.in_tile)) {
                        // interior tile of a packed float source: plain coalesced copy, no per-frame edge logic
                        const float2* src = reinterpret_cast<const float2*>(s.p0) + in0;
                        for (int f = tid; f < a.in_tile; f += nthr) dst[f] = __ldg(src + f);
                    } else {
                        for (int f = tid; f < a.in_tile; f += nthr) dst[f] = src_frame(s, in0 + f);
                    }
                } else {
                    for (int f = tid; f < a.in_tile; f += nthr) s_in[f] = src_frame(s, in0 + f).x;
                }
            }
            __syncthreads();
            const float vol = a.vol[inp];
            for (int q = warp; q < a.n_groups; q += nwarps) {
                float acc[kG][CH];
#pragma unroll
                for (int g = 0; g < kG; g++)
#pragma unroll
                    for (int c = 0; c < CH; c++) acc[g][c] = 0.f;
                const long long kq = k0 + (long long)lane * a.P + (long long)q * kG;   // first output of this thread
                if (k0 < a.out_len[inp] && kq < a.out_frames) {
                    const float* hq = s_hq + q * a.wmax * kG;
                    const int base = lane * a.D + s_gs[q];
                    if (CH == 2) {
                        const float2* x2 = reinterpret_cast<const float2*>(s_in) + base;
#pragma unroll 4
                        for (int m = 0; m < a.wmax; m++) {
                            const float2 x = x2[m];
                            const float4 h0 = *reinterpret_cast<const float4*>(hq + m * kG);
                            const float4 h1 = *reinterpret_cast<const float4*>(hq + m * kG + 4);
                            acc[0][0] = __fmaf_rn(x.x, h0.x, acc[0][0]); acc[0][CH - 1] = __fmaf_rn(x.y, h0.x, acc[0][CH - 1]);
                            acc[1][0] = __fmaf_rn(x.x, h0.y, acc[1][0]); acc[1][CH - 1] = __fmaf_rn(x.y, h0.y, acc[1][CH - 1]);
                            acc[2][0] = __fmaf_rn(x.x, h0.z, acc[2][0]); acc[2][CH - 1] = __fmaf_rn(x.y, h0.z, acc[2][CH - 1]);
                            acc[3][0] = __fmaf_rn(x.x, h0.w, acc[3][0]); acc[3][CH - 1] = __fmaf_rn(x.y, h0.w, acc[3][CH - 1]);
                            acc[4][0] = __fmaf_rn(x.x, h1.x, acc[4][0]); acc[4][CH - 1] = __fmaf_rn(x.y, h1.x, acc[4][CH - 1]);
                            acc[5][0] = __fmaf_rn(x.x, h1.y, acc[5][0]); acc[5][CH - 1] = __fmaf_rn(x.y, h1.y, acc[5][CH - 1]);
                            acc[6][0] = __fmaf_rn(x.x, h1.z, acc[6][0]); acc[6][CH - 1] = __fmaf_rn(x.y, h1.z, acc[6][CH - 1]);
                            acc[7][0] = __fmaf_rn(x.x, h1.w, acc[7][0]); acc[7][CH - 1] = __fmaf_rn(x.y, h1.w, acc[7][CH - 1]);
                        }
                    } else {
                        const float* x1 = s_in + base;
#pragma unroll 4
                        for (int m = 0; m < a.wmax; m++) {
                            const float x = x1[m];
                            const float4 h0 = *reinterpret_cast<const float4*>(hq + m * kG);
                            const float4 h1 = *reinterpret_cast<const float4*>(hq + m * kG + 4);
                            acc[0][0] = __fmaf_rn(x, h0.x, acc[0][0]); acc[1][0] = __fmaf_rn(x, h0.y, acc[1][0]);
                            acc[2][0] = __fmaf_rn(x, h0.z, acc[2][0]); acc[3][0] = __fmaf_rn(x, h0.w, acc[3][0]);
                            acc[4][0] = __fmaf_rn(x, h1.x, acc[4][0]); acc[5][0] = __fmaf_rn(x, h1.y, acc[5][0]);
                            acc[6][0] = __fmaf_rn(x, h1.z, acc[6][0]); acc[7][0] = __fmaf_rn(x, h1.w, acc[7][0]);
                        }
                    }
                }
                // accumulate in input order: temp += data * volume (audio-amix.cpp:300-304)
                if (direct) {
                    const int ng = a.P - q * kG < kG ? a.P - q * kG : kG;             // phases of the last group may run short
                    float o[kG][2];
#pragma unroll
                    for (int g = 0; g < kG; g++) {
                        const bool live = (kq + g) < a.out_len[inp];
                        float vl = live ? acc[g][0] : 0.f;
                        float vr = live ? acc[g][CH - 1] : 0.f;
                        if (a.mix) {
                            vl = __fadd_rn(0.f, __fmul_rn(vl, vol));      // 0.0f: the reference's zeroed temp buffer
                            vr = __fadd_rn(0.f, __fmul_rn(vr, vol));
                        }
                        o[g][0] = vl; o[g][1] = vr;
                    }
                    if (a.vec_out && ng == kG && kq + kG <= a.out_frames) {
                        float4* gl = reinterpret_cast<float4*>(out_l + kq);
                        float4* gr = reinterpret_cast<float4*>(out_r + kq);
                        gl[0] = make_float4(o[0][0], o[1][0], o[2][0], o[3][0]);
                        gl[1] = make_float4(o[4][0], o[5][0], o[6][0], o[7][0]);
                        gr[0] = make_float4(o[0][1], o[1][1], o[2][1], o[3][1]);
                        gr[1] = make_float4(o[4][1], o[5][1], o[6][1], o[7][1]);
                    } else {
#pragma unroll
                        for (int g = 0; g < kG; g++)
                            if (g < ng && kq + g < a.out_frames) { out_l[kq + g] = o[g][0]; out_r[kq + g] = o[g][1]; }
                    }
                    continue;
                }
                float* rl = s_out_l + lane * a.out_stride + q * kG;
                float* rr = s_out_r + lane * a.out_stride + q * kG;
                if (q * kG + kG <= a.P && kq + kG <= a.out_len[inp]) {
                    // whole group live: no per-output tests
#pragma unroll
                    for (int g = 0; g < kG; g++) {
                        float vl = acc[g][0], vr = acc[g][CH - 1];
                        if (a.mix) {
                            const float pl = inp ? rl[g] : 0.f, pr = inp ? rr[g] : 0.f;
                            vl = __fadd_rn(pl, __fmul_rn(vl, vol));
                            vr = __fadd_rn(pr, __fmul_rn(vr, vol));
                        }
                        rl[g] = vl;
                        rr[g] = vr;
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < kG; g++) {
                        if (q * kG + g >= a.P) break;
                        const bool live = (kq + g) < a.out_len[inp];
                        float vl = live ? acc[g][0] : 0.f;
                        float vr = live ? acc[g][CH - 1] : 0.f;
                        if (a.mix) {
                            const float pl = inp ? rl[g] : 0.f, pr = inp ? rr[g] : 0.f;
                            vl = __fadd_rn(pl, __fmul_rn(vl, vol));
                            vr = __fadd_rn(pr, __fmul_rn(vr, vol));
                        }
                        rl[g] = vl;
                        rr[g] = vr;
                    }
                }
            }
        }
        if (direct) continue;
        __syncthreads();
        // coalesced copy-out of the tile: warp w takes rows (periods) w, w + nwarps, ...; a row is P consecutive
        // output frames, so the lanes walk it in 128-byte steps -- no division, no per-element index arithmetic
        const long long remain = a.out_frames - k0;
        const int n_out = (int)(remain < (long long)kNB * a.P ? remain : (long long)kNB * a.P);
        for (int b = warp; b < kNB; b += nwarps) {
            const int row0 = b * a.P;
            if (row0 >= n_out) break;
            const int len = n_out - row0 < a.P ? n_out - row0 : a.P;
            const float* sl = s_out_l + b * a.out_stride;
            const float* sr = s_out_r + b * a.out_stride;
            float* gl = out_l + k0 + row0;
            float* gr = out_r + k0 + row0;
            for (int t = lane; t < len; t += 32) { gl[t] = sl[t]; gr[t] = sr[t]; }
        }
    }
}

// ---- tile kernel, pipelined ------------------------------------------------------------------------------
// Same arithmetic as resample_tile_kernel for plans with one phase group per warp (P <= 160, e.g. 44.1 -> 48 kHz),
// restructured after the ncu source view showed 39 % of the stall samples on the input staging loop:
//  * the input tile is DOUBLE BUFFERED and filled with cp.async (LDGSTS): the tile of the next (track, tile, input)
//    stage is in flight while the current one is filtered; one barrier per stage;
//  * a thread keeps its 8 outputs in registers across the inputs of a mix (MIX) and writes them with 128-bit
//    stores: no shared staging rows, no copy-out pass;
//  * a launch covers a BATCH of tracks (per-track source planes and volume ride in the kernel parameters), so the
//    256 per-track audio_amix(1) resamplers of a render are one persistent grid instead of 256 launches.
constexpr int kMaxResampleBatch = 256;
struct TrackPlanes {
    const void* p0[kMaxResampleBatch];
    const void* p1[kMaxResampleBatch];
    float vol[kMaxResampleBatch];
};

__device__ __forceinline__ void rs_cp_async4(float* dst_smem, const float* src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void rs_cp_async8(float* dst_smem, const float* src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(src) : "memory");
}

// PT = periods per thread (tile = 32 * PT periods): the warp-uniform 128-bit tap loads cost about three shared-memory
// wavefronts each (ncu: 7.7 wavefronts per window position and warp, 2 of them the samples), so with one period per
// thread the LSU pipe, not the FMA pipe, set the pace; two periods per thread halve the tap traffic per FMA.
// ---- TMA (bulk async copy engine) staging: one elected thread issues the copy, an mbarrier counts the bytes ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* m, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(m)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* m, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(m)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* m)
{
    // 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion signalled on the mbarrier
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(m)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* m, unsigned parity)
{
    unsigned done = 0, spins = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(m)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 22)) __trap();      // a byte-count mismatch must not hang the device
    }
}

template <int CH, bool MIX, int PT>
__global__ void __launch_bounds__(640, (MIX || PT > 1) ? 1 : 2) resample_tile2_kernel(float* __restrict__ out_l, float* __restrict__ out_r,
                                                                                      const __grid_constant__ TileArgs a,
                                                                                      const __grid_constant__ TrackPlanes tp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [hq: n_groups*wmax*G floats][group_start: n_groups ints][input tile 0][input tile 1]
    float* s_hq = reinterpret_cast<float*>(smem_raw);
    int* s_gs = reinterpret_cast<int*>(s_hq + a.n_groups * a.wmax * kG);
    float* s_in0 = reinterpret_cast<float*>(s_gs + ((a.n_groups + 3) & ~3));
    constexpr int NBT = kNB * PT;                       // periods per tile
    const int in_tile = (NBT - 1) * a.D + a.span + a.wmax + 1;
    const int buf_floats = ((in_tile + 2) * CH + 3) & ~3;      // + 2 frames: a TMA-staged tile starts on an even frame
    __shared__ __align__(8) unsigned long long mbar[3];        // tile buffer 0, tile buffer 1, tap table

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int q = tid >> 5, lane = tid & 31;           // one phase group per warp

    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_init(&mbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the dense tap table (25 KB for 44.1 -> 48 kHz) is staged by the TMA engine: one bulk copy, one elected thread
    const unsigned hq_bytes = (unsigned)(a.n_groups * a.wmax * kG * sizeof(float));
    const bool hq_tma = (hq_bytes & 15u) == 0 && (((uintptr_t)a.hq) & 15) == 0;
    if (hq_tma) {
        if (tid == 0) { mbar_expect_tx(&mbar[2], hq_bytes); tma_bulk_g2s(s_hq, a.hq, hq_bytes, &mbar[2]); }
    } else {
        for (int i = tid; i < a.n_groups * a.wmax * kG; i += nthr) s_hq[i] = a.hq[i];
    }
    for (int i = tid; i < a.n_groups; i += nthr) s_gs[i] = a.group_start[i];

    const long long total = a.n_tiles * (long long)a.ntracks;
    const int nin = MIX ? a.nin : 1;                  // compile-time 1 without MIX: the accumulators die with the stage

    // stage (item, inp): fill buffer `b` with the input tile -- asynchronously when the tile is an interior run of float
    // frames.  Returns -1 when the tile was handed to the TMA engine (wait on mbar[b]), else 0; *lead = frames in front of
    // the tile's first frame inside the buffer.
    const auto stage = [&](long long item, int inp, int b, int* lead) -> int {
        float* dst = s_in0 + b * buf_floats;
        *lead = 0;
        const long long track = item / a.n_tiles, tile = item - track * a.n_tiles + a.tile0;
        const long long k0 = tile * (long long)NBT * a.P;
        if (k0 >= a.out_len[inp]) return 0;                     // contributes zeros: nothing is read
        const long long in0 = tile * (long long)NBT * a.D + a.s0 - a.center;
        SrcDesc s = a.src[inp];
        if (!MIX) { s.p0 = tp.p0[track]; s.p1 = tp.p1[track]; }
        const bool interior = tile_interior(s, in0, in_tile);
        if (CH == 2 && interior && s.fmt == NODEY_FMT_FLT && (((uintptr_t)s.p0) & 15) == 0) {
            // packed float frames: ONE bulk copy by the TMA engine.  It wants 16-byte alignment and size: start on the
            // even frame at or before in0 and take an even number of frames (the buffer has two spare frames).
            const long long in0e = in0 & ~1ll;
            const int ld = (int)(in0 - in0e);
            const int nfr = (in_tile + ld + 1) & ~1;
            if (in0e + nfr <= s.n) {
                if (tid == 0) {
                    mbar_expect_tx(&mbar[b], (unsigned)nfr * 8u);
                    tma_bulk_g2s(dst, reinterpret_cast<const float*>(s.p0) + 2 * in0e, (unsigned)nfr * 8u, &mbar[b]);
                }
                *lead = ld;
                return -1;
            }
        }
        if (CH == 2 && interior && s.fmt == NODEY_FMT_FLT && (((uintptr_t)s.p0) & 7) == 0) {
            const float* src = reinterpret_cast<const float*>(s.p0) + 2 * in0;
            for (int f = tid; f < in_tile; f += nthr) rs_cp_async8(dst + 2 * f, src + 2 * f);
        } else if (CH == 2 && interior && s.fmt == NODEY_FMT_FLTP) {
            const float* sl = reinterpret_cast<const float*>(s.p0) + in0;
            const float* sr = reinterpret_cast<const float*>(s.p1) + in0;
            for (int f = tid; f < in_tile; f += nthr) { rs_cp_async4(dst + 2 * f, sl + f); rs_cp_async4(dst + 2 * f + 1, sr + f); }
        } else if (CH == 2) {
            float2* d2 = reinterpret_cast<float2*>(dst);
            for (int f = tid; f < in_tile; f += nthr) d2[f] = src_frame(s, in0 + f);
        } else {
            for (int f = tid; f < in_tile; f += nthr) dst[f] = src_frame(s, in0 + f).x;
        }
        return 0;
    };

    long long item = blockIdx.x;
    int inp = 0, cur = 0;
    int lead_cur = 0, lead_nxt = 0, tma_cur = 0, tma_nxt = 0;     // per buffer: frames of lead-in, staged by TMA?
    unsigned phase = 0;                                            // bit b: parity the next wait on mbar[b] uses
    if (item < total) tma_cur = stage(item, 0, 0, &lead_cur);
    if (hq_tma) mbar_wait(&mbar[2], 0);
    float macc[PT][kG][2];
#pragma unroll
    for (int p = 0; p < PT; p++)
#pragma unroll
        for (int g = 0; g < kG; g++) macc[p][g][0] = macc[p][g][1] = 0.f;

    while (item < total) {
        if (tma_cur) { mbar_wait(&mbar[cur], (phase >> cur) & 1u); phase ^= 1u << cur; }
        else asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic accesses of the other buffer before the TMA rewrites it
        __syncthreads();          // this stage's tile has landed; everybody is done reading the other buffer
        long long nitem = item; int ninp = inp + 1;
        if (ninp >= nin) { ninp = 0; nitem += gridDim.x; }
        tma_nxt = 0; lead_nxt = 0;
        if (nitem < total) tma_nxt = stage(nitem, ninp, cur ^ 1, &lead_nxt);

        const long long track = item / a.n_tiles, tile = item - track * a.n_tiles + a.tile0;
        const long long k0 = tile * (long long)NBT * a.P;
        const float* s_in = s_in0 + cur * buf_floats + lead_cur * CH;
        const float vol = MIX ? a.vol[inp] : tp.vol[track];
        if (q < a.n_groups) {
            float acc[PT][kG][CH];
#pragma unroll
            for (int p = 0; p < PT; p++)
#pragma unroll
                for (int g = 0; g < kG; g++)
#pragma unroll
                    for (int c = 0; c < CH; c++) acc[p][g][c] = 0.f;
            // thread = periods lane, lane + 32, ... of the tile; first output of period p: kq + p * 32 * P
            const long long kq = k0 + (long long)lane * a.P + (long long)q * kG;
            if (k0 < a.out_len[inp] && kq < a.out_frames) {
                const float* hq = s_hq + q * a.wmax * kG;
                const int base = lane * a.D + s_gs[q];
                const int pstep = kNB * a.D;                       // input frames between a thread's periods
                if (CH == 2) {
                    // both channels of a frame meet the same tap: ONE packed fused multiply-add (fma.rn.f32x2, per component
                    // the fmaf the oracle's FIR is defined with) with the tap as a scalar operand -- half the issue slots
                    const float2* x2 = reinterpret_cast<const float2*>(s_in) + base;
                    float2 a2[PT][kG];
#pragma unroll
                    for (int p = 0; p < PT; p++)
#pragma unroll
                        for (int g = 0; g < kG; g++) a2[p][g] = make_float2(0.f, 0.f);
#pragma unroll 3
                    for (int m = 0; m < a.wmax; m++) {
                        const float4 h0 = *reinterpret_cast<const float4*>(hq + m * kG);
                        const float4 h1 = *reinterpret_cast<const float4*>(hq + m * kG + 4);
#pragma unroll
                        for (int p = 0; p < PT; p++) {
                            const float2 x = x2[m + p * pstep];
                            a2[p][0] = __ffma2_rn(x, make_float2(h0.x, h0.x), a2[p][0]);
                            a2[p][1] = __ffma2_rn(x, make_float2(h0.y, h0.y), a2[p][1]);
                            a2[p][2] = __ffma2_rn(x, make_float2(h0.z, h0.z), a2[p][2]);
                            a2[p][3] = __ffma2_rn(x, make_float2(h0.w, h0.w), a2[p][3]);
                            a2[p][4] = __ffma2_rn(x, make_float2(h1.x, h1.x), a2[p][4]);
                            a2[p][5] = __ffma2_rn(x, make_float2(h1.y, h1.y), a2[p][5]);
                            a2[p][6] = __ffma2_rn(x, make_float2(h1.z, h1.z), a2[p][6]);
                            a2[p][7] = __ffma2_rn(x, make_float2(h1.w, h1.w), a2[p][7]);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < PT; p++)
#pragma unroll
                        for (int g = 0; g < kG; g++) { acc[p][g][0] = a2[p][g].x; acc[p][g][CH - 1] = a2[p][g].y; }
                } else {
                    const float* x1 = s_in + base;
#pragma unroll 3
                    for (int m = 0; m < a.wmax; m++) {
                        const float4 h0 = *reinterpret_cast<const float4*>(hq + m * kG);
                        const float4 h1 = *reinterpret_cast<const float4*>(hq + m * kG + 4);
#pragma unroll
                        for (int p = 0; p < PT; p++) {
                            const float x = x1[m + p * pstep];
                            acc[p][0][0] = __fmaf_rn(x, h0.x, acc[p][0][0]); acc[p][1][0] = __fmaf_rn(x, h0.y, acc[p][1][0]);
                            acc[p][2][0] = __fmaf_rn(x, h0.z, acc[p][2][0]); acc[p][3][0] = __fmaf_rn(x, h0.w, acc[p][3][0]);
                            acc[p][4][0] = __fmaf_rn(x, h1.x, acc[p][4][0]); acc[p][5][0] = __fmaf_rn(x, h1.y, acc[p][5][0]);
                            acc[p][6][0] = __fmaf_rn(x, h1.z, acc[p][6][0]); acc[p][7][0] = __fmaf_rn(x, h1.w, acc[p][7][0]);
                        }
                    }
                }
            }
            float* gl0 = out_l + track * a.out_track_stride;
            float* gr0 = out_r + track * a.out_track_stride;
            const int ng = a.P - q * kG < kG ? a.P - q * kG : kG;             // phases of the last group may run short
#pragma unroll
            for (int p = 0; p < PT; p++) {
                const long long kp = kq + (long long)p * kNB * a.P;
                // accumulate in input order: temp += data * volume, temp zeroed first (audio-amix.cpp:296-304)
#pragma unroll
                for (int g = 0; g < kG; g++) {
                    const bool live = (kp + g) < a.out_len[inp];
                    float vl = live ? acc[p][g][0] : 0.f;
                    float vr = live ? acc[p][g][CH - 1] : 0.f;
                    if (a.mix) {
                        vl = __fadd_rn((MIX && inp) ? macc[p][g][0] : 0.f, __fmul_rn(vl, vol));
                        vr = __fadd_rn((MIX && inp) ? macc[p][g][1] : 0.f, __fmul_rn(vr, vol));
                    }
                    macc[p][g][0] = vl; macc[p][g][1] = vr;
                }
                if (inp == nin - 1) {
                    if (a.vec_out && ng == kG && kp + kG <= a.out_frames) {
                        float4* gl = reinterpret_cast<float4*>(gl0 + kp);
                        float4* gr = reinterpret_cast<float4*>(gr0 + kp);
                        gl[0] = make_float4(macc[p][0][0], macc[p][1][0], macc[p][2][0], macc[p][3][0]);
                        gl[1] = make_float4(macc[p][4][0], macc[p][5][0], macc[p][6][0], macc[p][7][0]);
                        gr[0] = make_float4(macc[p][0][1], macc[p][1][1], macc[p][2][1], macc[p][3][1]);
                        gr[1] = make_float4(macc[p][4][1], macc[p][5][1], macc[p][6][1], macc[p][7][1]);
                    } else {
#pragma unroll
                        for (int g = 0; g < kG; g++)
                            if (g < ng && kp + g < a.out_frames) { gl0[kp + g] = macc[p][g][0]; gr0[kp + g] = macc[p][g][1]; }
                    }
                }
            }
        }
        item = nitem; inp = ninp; cur ^= 1;
        lead_cur = lead_nxt; tma_cur = tma_nxt;
    }
}

static int upload(const void* host, size_t bytes, void** dev)
{
    NODEY_CUDA_OK(cudaMalloc(dev, bytes));
    NODEY_CUDA_OK(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    return NODEY_OK;
}

static int64_t plan_pos_index(const nodey_resampler* r, int64_t k)
{
    const int64_t f = k * (int64_t)r->dst_incr_mod;
    return (int64_t)r->index0 + k * (int64_t)r->dst_incr_div + f / r->src_incr;
}

// outputs available from n real frames plus `reflect` reflected ones
static int64_t plan_producible(const nodey_resampler* r, int64_t n, int64_t reflect)
{
    const int64_t L = r->filter_length, P = r->phase_count, center = (L - 1) / 2;
    // invert_initial_buffer() waits for filter_length + 1 samples, and the samples resample_flush() reflects count: a
    // stream a little shorter than the filter (22..32 frames at 32 taps) gets there at the flush and does produce output
    // (pinned against the real library: tests/test_swr_real.py::test_streams_shorter_than_the_filter)
    if (n + reflect < L + 1) return 0;
    const int64_t max_s = n + reflect - L + center;
    if (max_s < 0) return 0;
    int64_t lo = 0, hi = (max_s + 2) * P / (r->dst_incr_div > 0 ? r->dst_incr_div : 1) + 4;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (plan_pos_index(r, mid) / P <= max_s) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static int64_t plan_reflect(const nodey_resampler* r, int64_t n, int64_t produced)
{
    if (n < r->filter_length + 1) return (n + 1) / 2;      // nothing consumed yet: in_buffer_count is the whole input
    const int64_t wstart = plan_pos_index(r, produced) / r->phase_count - (r->filter_length - 1) / 2;
    int64_t held = n - wstart;
    if (held > r->filter_length) held = r->filter_length;
    if (held < 0) held = 0;
    return (held + 1) / 2;
}

}  // namespace nodey

using namespace nodey;

// resample_init() with the library defaults the reference leaves untouched (host-only numbers + bank)
static void plan_init_math(nodey_resampler* r, int index_mask_quirk)
{
    const int in_rate = r->in_rate, out_rate = r->out_rate;
    r->resample = in_rate != out_rate;
    if (!r->resample) {
        r->phase_count = 1; r->filter_length = 1; r->filter_alloc = 8;
        r->src_incr = r->dst_incr = 1; r->dst_incr_div = 1; r->dst_incr_mod = 0; r->index0 = 0;
        return;
    }
    double factor = (double)out_rate * 0.97 / in_rate;
    if (factor > 1.0) factor = 1.0;
    int phase_count = 1 << 10;
    int filter_length = (int)ceil(32 / factor);
    if (filter_length < 1) filter_length = 1;
    if (filter_length > 1) filter_length = (filter_length + 1) & ~1;
    {
        const int64_t g = gcd64(out_rate, in_rate);
        const int64_t exact = out_rate / g;
        if (exact <= phase_count) phase_count = (int)exact;
    }
    r->phase_count = phase_count;
    r->filter_length = filter_length;
    r->filter_alloc = (filter_length + 7) & ~7;
    r->bank.assign((size_t)r->filter_alloc * (size_t)(phase_count + 1), 0.f);
    build_filter(r->bank.data(), factor, filter_length, r->filter_alloc, phase_count, 9.0);
    memcpy(r->bank.data() + (size_t)r->filter_alloc * phase_count + 1, r->bank.data(), sizeof(float) * (size_t)(r->filter_alloc - 1));
    r->bank[(size_t)r->filter_alloc * phase_count] = r->bank[(size_t)r->filter_alloc - 1];
    {
        int64_t num = out_rate, den = (int64_t)in_rate * phase_count;
        const int64_t g = gcd64(num, den);
        num /= g; den /= g;
        while (den < (1 << 20) && num < (1 << 20)) { den *= 2; num *= 2; }
        r->src_incr = (int)num; r->dst_incr = (int)den;
    }
    r->dst_incr_div = r->dst_incr / r->src_incr;
    r->dst_incr_mod = r->dst_incr % r->src_incr;
    {
        const int idx = -phase_count * ((filter_length - 1) / 2);
        r->index0 = index_mask_quirk ? (idx & (phase_count - 1)) : 0;
    }
}

static size_t tile_geometry(const nodey_resampler* r, TileArgs& a, int ch)
{
    a.P = r->phase_count; a.D = r->dst_incr_div; a.L = r->filter_length; a.center = (r->filter_length - 1) / 2;
    a.n_groups = r->n_groups; a.wmax = r->wmax; a.s0 = r->s0; a.span = r->span;
    a.in_tile = (kNB - 1) * a.D + a.span + a.wmax + 1;
    a.out_stride = (a.n_groups * kG) | 1;
    return sizeof(float) * ((size_t)a.n_groups * a.wmax * kG + (size_t)((a.n_groups + 3) & ~3) +
                            2 * (size_t)kNB * a.out_stride + (size_t)a.in_tile * ch + 4);
}


extern "C" {

int nodey_resampler_create(nodey_resampler** out, int in_rate, int out_rate, int index_mask_quirk)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_resampler_create: null out pointer");
    NODEY_REQUIRE(in_rate > 0 && out_rate > 0, NODEY_E_INVALID, "nodey_resampler_create: bad sample rate");
    nodey_resampler* r = new nodey_resampler();
    r->in_rate = in_rate; r->out_rate = out_rate;
    r->resample = in_rate != out_rate;
    cudaGetDevice(&r->device);
    if (!r->resample) {
        r->phase_count = 1; r->filter_length = 1; r->filter_alloc = 8;
        r->src_incr = r->dst_incr = 1; r->dst_incr_div = 1; r->dst_incr_mod = 0; r->index0 = 0;
        *out = r;
        return NODEY_OK;
    }
    plan_init_math(r, index_mask_quirk);
    int rc = upload(r->bank.data(), r->bank.size() * sizeof(float), (void**)&r->d_bank);
    if (rc != NODEY_OK) { delete r; return rc; }

    // tile-kernel tables: only for exact plans (no inter-phase interpolation)
    const int phase_count = r->phase_count, filter_length = r->filter_length;
    if (r->dst_incr_mod == 0 && phase_count <= 1024 && filter_length <= 256) {
        const int P = phase_count, D = r->dst_incr_div, L = filter_length;
        std::vector<int> s_t((size_t)P), ph_t((size_t)P);
        for (int t = 0; t < P; t++) {
            const int64_t pos = (int64_t)r->index0 + (int64_t)t * D;
            s_t[(size_t)t] = (int)(pos / P);
            ph_t[(size_t)t] = (int)(pos % P);
        }
        r->n_groups = (P + kG - 1) / kG;
        r->s0 = s_t[0];
        r->span = s_t[(size_t)P - 1] - s_t[0];
        int wmax = 0;
        std::vector<int> gs((size_t)r->n_groups);
        for (int q = 0; q < r->n_groups; q++) {
            const int t0 = q * kG, t1 = (t0 + kG - 1 < P ? t0 + kG - 1 : P - 1);
            const int w = s_t[(size_t)t1] - s_t[(size_t)t0] + L;
            if (w > wmax) wmax = w;
            gs[(size_t)q] = s_t[(size_t)t0] - s_t[0];
        }
        r->wmax = wmax;
        std::vector<float> hq((size_t)r->n_groups * (size_t)wmax * kG, 0.f);
        for (int q = 0; q < r->n_groups; q++)
            for (int g = 0; g < kG; g++) {
                const int t = q * kG + g;
                if (t >= P) break;
                const int shift = s_t[(size_t)t] - s_t[(size_t)(q * kG)];
                for (int i = 0; i < L; i++)
                    hq[((size_t)q * (size_t)wmax + (size_t)(shift + i)) * kG + (size_t)g] =
                        r->bank[(size_t)r->filter_alloc * (size_t)ph_t[(size_t)t] + (size_t)i];
            }
        rc = upload(hq.data(), hq.size() * sizeof(float), (void**)&r->d_hq);
        if (rc == NODEY_OK) rc = upload(gs.data(), gs.size() * sizeof(int), (void**)&r->d_group_start);
        if (rc != NODEY_OK) { nodey_resampler_destroy(r); return rc; }
        TileArgs probe;
        r->tile_ok = tile_geometry(r, probe, 2) <= 227 * 1024;   // else: generic kernel (plan too big for one SM)
    }
    *out = r;
    return NODEY_OK;
}

void nodey_resampler_destroy(nodey_resampler* r)
{
    if (!r) return;
    if (r->d_bank) cudaFree(r->d_bank);
    if (r->d_hq) cudaFree(r->d_hq);
    if (r->d_group_start) cudaFree(r->d_group_start);
    delete r;
}

int nodey_resampler_info(const nodey_resampler* r, int info[8])
{
    NODEY_REQUIRE(r && info, NODEY_E_INVALID, "nodey_resampler_info: null argument");
    info[0] = r->phase_count; info[1] = r->filter_length; info[2] = r->filter_alloc;
    info[3] = r->dst_incr_div; info[4] = r->dst_incr_mod; info[5] = r->src_incr;
    info[6] = r->index0; info[7] = r->resample && r->dst_incr_mod != 0;
    return NODEY_OK;
}

const float* nodey_resampler_filter_bank(const nodey_resampler* r) { return r && r->resample ? r->bank.data() : nullptr; }

int64_t nodey_resampler_out_count(const nodey_resampler* r, int64_t in_frames, int flush)
{
    if (!r || in_frames < 0) return NODEY_E_INVALID;
    if (!r->resample) return in_frames;
    int64_t n = plan_producible(r, in_frames, 0);
    if (flush) n = plan_producible(r, in_frames, plan_reflect(r, in_frames, n));
    return n;
}

/* streaming bookkeeping of swr_convert (host only): outputs available from n_in real frames plus `reflect`
 * reflected ones, and the reflection length resample_flush() appends when `produced` outputs were taken */
int64_t nodey_resampler_producible(const nodey_resampler* r, int64_t n_in, int64_t reflect)
{
    if (!r || n_in < 0 || reflect < 0) return NODEY_E_INVALID;
    if (!r->resample) return n_in;
    return plan_producible(r, n_in, reflect);
}

int64_t nodey_resampler_flush_reflect(const nodey_resampler* r, int64_t n_in, int64_t produced)
{
    if (!r || n_in < 0 || produced < 0) return NODEY_E_INVALID;
    if (!r->resample) return 0;
    return plan_reflect(r, n_in, produced);
}

// Reflection length of a flushed conversion that has to deliver `want` outputs.  resample_flush() appends
// (min(in_buffer_count, filter_length) + 1) / 2 mirrored frames: a conversion drained with ample capacity holds
// filter_length - 1 frames at that point, one drained through small output capacities (audio_amix's nb next to a
// faster input) still holds more than filter_length and so reflects ONE frame more -- and can return one more output
// (tests/golden/swr_real_amix_capped.npz: 10213 instead of 10212).  The mirrored frames are the same sequence either
// way, so the shortest reflection that yields `want` reproduces the library; -1: not even the longest one does.
static int64_t flush_reflect_for(const nodey_resampler* r, int64_t in_frames, int64_t want)
{
    int64_t reflect = plan_reflect(r, in_frames, plan_producible(r, in_frames, 0));
    const int64_t longest = in_frames < r->filter_length ? (in_frames + 1) / 2 : (r->filter_length + 1) / 2;
    while (plan_producible(r, in_frames, reflect) < want) {
        if (reflect >= longest) return -1;
        reflect++;
    }
    return reflect;
}

// outputs a conversion of in_frames frames can deliver at most (flush: with the longest reflection swr_convert can append)
static int64_t max_out_count(const nodey_resampler* r, int64_t in_frames, int flush)
{
    if (!r->resample) return in_frames;
    if (!flush) return plan_producible(r, in_frames, 0);
    const int64_t usual = plan_reflect(r, in_frames, plan_producible(r, in_frames, 0));
    const int64_t longest = in_frames < r->filter_length ? (in_frames + 1) / 2 : (r->filter_length + 1) / 2;
    return plan_producible(r, in_frames, longest > usual ? longest : usual);
}

static int fill_src(SrcDesc* s, const nodey_resampler* r, const void* p0, const void* p1, int fmt, int nch,
                    int64_t in_frames, int flush, int64_t want = 0)
{
    NODEY_REQUIRE(nch == 1 || nch == 2, NODEY_E_INVALID, "Invalid channel layout: %d", nch);
    NODEY_REQUIRE(fmt_bytes(fmt) != 0, NODEY_E_FORMAT, "resampler: unsupported sample format %d", fmt);
    NODEY_REQUIRE(in_frames >= 0, NODEY_E_INVALID, "resampler: negative input size");
    s->p0 = p0; s->p1 = p1; s->n = in_frames; s->fmt = fmt; s->nch = nch; s->planar = fmt_planar(fmt) ? 1 : 0;
    s->reflect = 0;
    if (flush && r->resample) {
        const int64_t reflect = flush_reflect_for(r, in_frames, want);
        NODEY_REQUIRE(reflect >= 0, NODEY_E_RANGE, "resampler: %lld outputs exceed what swr would produce from %lld frames", (long long)want, (long long)in_frames);
        s->reflect = reflect;
    }
    return NODEY_OK;
}

static int launch_tile(const nodey_resampler* r, float* out_l, float* out_r, TileArgs& a, int ch, cudaStream_t st)
{
    a.hq = r->d_hq; a.group_start = r->d_group_start;
    const size_t smem = tile_geometry(r, a, ch);
    a.n_tiles = (a.out_frames + (int64_t)kNB * a.P - 1) / ((int64_t)kNB * a.P);
    a.vec_out = (((uintptr_t)out_l | (uintptr_t)out_r) & 15) == 0 && a.P % 4 == 0;
    NODEY_REQUIRE(smem <= 227 * 1024, NODEY_E_RANGE, "resample tile kernel: plan needs %zu bytes of shared memory", smem);
    int threads = 32 * (a.n_groups < 20 ? a.n_groups : 20);
    const int ctas_per_sm = smem > 113 * 1024 ? 1 : 2;
    int grid = (int)(a.n_tiles < (int64_t)sm_count() * ctas_per_sm ? a.n_tiles : (int64_t)sm_count() * ctas_per_sm);
    if (grid < 1) grid = 1;
    if (ch == 2) {
        NODEY_CUDA_OK(cudaFuncSetAttribute(resample_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NODEY_CUDA_OK(cudaFuncSetAttribute(resample_tile_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        NODEY_LAUNCH("resample_tile_kernel", st, resample_tile_kernel<2><<<grid, threads, smem, st>>>(out_l, out_r, a));
    } else {
        NODEY_CUDA_OK(cudaFuncSetAttribute(resample_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NODEY_CUDA_OK(cudaFuncSetAttribute(resample_tile_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        NODEY_LAUNCH("resample_tile_kernel", st, resample_tile_kernel<1><<<grid, threads, smem, st>>>(out_l, out_r, a));
    }
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

// pipelined tile kernel: plans with one phase group per warp
static bool tile2_ok(const nodey_resampler* r) { return r->tile_ok && r->n_groups <= 20; }

// periods per thread and shared memory of the pipelined kernel for this plan (a, ch): two periods per thread when the
// double-buffered 64-period tile fits one SM
static int tile2_shape(const nodey_resampler* r, TileArgs& a, int ch, int force_pt, size_t* smem_out)
{
    a.hq = r->d_hq; a.group_start = r->d_group_start;
    tile_geometry(r, a, ch);
    const size_t table = (size_t)a.n_groups * a.wmax * kG + (size_t)((a.n_groups + 3) & ~3);
    const auto smem_for = [&](int pt) {
        const size_t in_tile = (size_t)(kNB * pt - 1) * a.D + a.span + a.wmax + 1;
        return sizeof(float) * (table + 2 * (((in_tile + 2) * ch + 3) & ~(size_t)3));
    };
    int pt = smem_for(2) <= 227 * 1024 ? 2 : 1;
    if (force_pt) pt = force_pt;
    *smem_out = smem_for(pt);
    return pt;
}

static int launch_tile2(const nodey_resampler* r, float* out_l, float* out_r, TileArgs& a, const TrackPlanes& tp, int ch, cudaStream_t st, int force_pt = 0)
{
    size_t smem = 0;
    const int pt = tile2_shape(r, a, ch, force_pt, &smem);
    NODEY_REQUIRE(smem <= 227 * 1024, NODEY_E_RANGE, "resample tile kernel: plan needs %zu bytes of shared memory", smem);
    const int64_t per_tile = (int64_t)kNB * pt * a.P;
    {
        const int64_t all = (a.out_frames + per_tile - 1) / per_tile;
        const int64_t t1 = a.tile1 > 0 && a.tile1 < all ? a.tile1 : all;
        if (a.tile0 < 0) a.tile0 = 0;
        a.n_tiles = t1 - a.tile0;
        if (a.n_tiles <= 0) return NODEY_OK;
    }
    if (a.ntracks < 1) a.ntracks = 1;
    a.vec_out = (((uintptr_t)out_l | (uintptr_t)out_r) & 15) == 0 && a.P % 4 == 0 && a.out_track_stride % 4 == 0;
    const bool mix = a.nin > 1;
    const int threads = 32 * a.n_groups;
    const int ctas_per_sm = (mix || pt > 1 || smem > 113 * 1024) ? 1 : 2;
    const int64_t total = a.n_tiles * a.ntracks;
    int grid = (int)(total < (int64_t)sm_count() * ctas_per_sm ? total : (int64_t)sm_count() * ctas_per_sm);
    if (grid < 1) grid = 1;
    void (*kern)(float*, float*, TileArgs, TrackPlanes);
    if (pt == 2)
        kern = ch == 2 ? (mix ? resample_tile2_kernel<2, true, 2> : resample_tile2_kernel<2, false, 2>)
                       : (mix ? resample_tile2_kernel<1, true, 2> : resample_tile2_kernel<1, false, 2>);
    else
        kern = ch == 2 ? (mix ? resample_tile2_kernel<2, true, 1> : resample_tile2_kernel<2, false, 1>)
                       : (mix ? resample_tile2_kernel<1, true, 1> : resample_tile2_kernel<1, false, 1>);
    NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    NODEY_LAUNCH("resample_tile_kernel", st, kern<<<grid, threads, smem, st>>>(out_l, out_r, a, tp));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

// mode: 0 auto, 1 force generic, 2 force the staging-row tile kernel, 3 force the pipelined tile kernel,
// 4 the pipelined kernel with one period per thread
// (testing hook, not in the public header)
int nodey_resampler_run_mode(const nodey_resampler* r, float* out_l, float* out_r, const void* p0, const void* p1,
                             int fmt, int nch, int64_t in_frames, int flush, int64_t out_frames, int mode,
                             nodey_stream_t stream)
{
    NODEY_REQUIRE(r && out_l && out_r, NODEY_E_INVALID, "nodey_resampler_run: null argument");
    cudaStream_t st = as_stream(stream);
    if (!r->resample) {
        NODEY_REQUIRE(out_frames <= in_frames, NODEY_E_RANGE, "nodey_resampler_run: out_frames exceeds input");
        return nodey_to_fltp_stereo(out_l, out_r, p0, p1, fmt, nch, out_frames, stream);
    }
    const int64_t avail = max_out_count(r, in_frames, flush);
    NODEY_REQUIRE(out_frames >= 0 && out_frames <= avail, NODEY_E_RANGE,
                  "nodey_resampler_run: out_frames %lld exceeds what swr would produce (%lld)", (long long)out_frames, (long long)avail);
    if (out_frames == 0) return NODEY_OK;
    if ((mode == 0 && r->tile_ok) || mode == 2 || mode == 3 || mode == 4) {
        NODEY_REQUIRE(r->tile_ok, NODEY_E_RANGE, "tile kernel unavailable for this plan");
        NODEY_REQUIRE(mode < 3 || tile2_ok(r), NODEY_E_RANGE, "pipelined tile kernel unavailable for this plan");
        TileArgs a;
        memset(&a, 0, sizeof(a));
        int rc = fill_src(&a.src[0], r, p0, p1, fmt, nch, in_frames, flush, out_frames);
        if (rc != NODEY_OK) return rc;
        a.out_len[0] = out_frames; a.vol[0] = 1.f; a.nin = 1; a.mix = 0; a.out_frames = out_frames;
        if (mode >= 3 || (mode == 0 && tile2_ok(r))) {
            static thread_local TrackPlanes tp;
            tp.p0[0] = p0; tp.p1[0] = p1; tp.vol[0] = 1.f;
            a.ntracks = 1; a.out_track_stride = 0;
            return launch_tile2(r, out_l, out_r, a, tp, nch, st, mode == 4 ? 1 : 0);
        }
        return launch_tile(r, out_l, out_r, a, nch, st);
    }
    SrcDesc s;
    int rc = fill_src(&s, r, p0, p1, fmt, nch, in_frames, flush, out_frames);
    if (rc != NODEY_OK) return rc;
    PlanDev pl{r->d_bank, r->phase_count, r->filter_length, r->filter_alloc, r->dst_incr_div, r->dst_incr_mod, r->src_incr, r->index0};
    NODEY_LAUNCH("resample_generic_kernel", st, resample_generic_kernel<<<stream_grid(out_frames, 256, 8), 256, 0, st>>>(out_l, out_r, s, pl, out_frames));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

int nodey_resampler_run(const nodey_resampler* r, float* out_l, float* out_r, const void* p0, const void* p1, int fmt,
                        int nch, int64_t in_frames, int flush, int64_t out_frames, nodey_stream_t stream)
{
    return nodey_resampler_run_mode(r, out_l, out_r, p0, p1, fmt, nch, in_frames, flush, out_frames, 0, stream);
}

int nodey_resample_mix(const nodey_resampler* r, float* out_l, float* out_r, const void* const* plane0,
                       const void* const* plane1, const int* fmt, const int* nch, const int64_t* in_frames,
                       const int64_t* out_len, const float* volumes, int nin, int flush, int64_t out_frames,
                       nodey_stream_t stream)
{
    NODEY_REQUIRE(r && out_l && out_r, NODEY_E_INVALID, "nodey_resample_mix: null argument");
    NODEY_REQUIRE(nin >= 1 && nin <= NODEY_MAX_MIX_INPUTS, NODEY_E_RANGE, "nodey_resample_mix: input_num %d outside 1..16", nin);
    NODEY_REQUIRE(r->resample && r->tile_ok, NODEY_E_RANGE, "nodey_resample_mix: plan has no tile kernel (use nodey_resampler_run + nodey_mix)");
    if (out_frames <= 0) return out_frames == 0 ? NODEY_OK : NODEY_E_INVALID;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    int ch = nch[0];
    for (int i = 0; i < nin; i++) {
        NODEY_REQUIRE(nch[i] == ch, NODEY_E_INVALID, "nodey_resample_mix: inputs must share a channel count");
        const int64_t avail = max_out_count(r, in_frames[i], flush);
        NODEY_REQUIRE(out_len[i] >= 0 && out_len[i] <= avail, NODEY_E_RANGE,
                      "nodey_resample_mix: out_len[%d]=%lld exceeds what swr would produce (%lld)", i, (long long)out_len[i], (long long)avail);
        int rc = fill_src(&a.src[i], r, plane0[i], plane1 ? plane1[i] : nullptr, fmt[i], nch[i], in_frames[i], flush, out_len[i]);
        if (rc != NODEY_OK) return rc;
        a.out_len[i] = out_len[i];
        a.vol[i] = volumes[i];
    }
    a.nin = nin; a.mix = 1; a.out_frames = out_frames;
    if (tile2_ok(r)) {
        static thread_local TrackPlanes tp;
        tp.p0[0] = plane0[0]; tp.p1[0] = plane1 ? plane1[0] : nullptr; tp.vol[0] = volumes[0];
        a.ntracks = 1; a.out_track_stride = 0;
        return launch_tile2(r, out_l, out_r, a, tp, ch, as_stream(stream));
    }
    return launch_tile(r, out_l, out_r, a, ch, as_stream(stream));
}

int nodey_resample_tracks(const nodey_resampler* r, float* out_l, float* out_r, int64_t out_track_stride,
                          const void* const* plane0, const void* const* plane1, int fmt, int nch, int64_t in_frames,
                          const float* volumes, int ntracks, int flush, int64_t out_len, int64_t out_frames,
                          nodey_stream_t stream)
{
    NODEY_REQUIRE(r && out_l && out_r && plane0 && volumes, NODEY_E_INVALID, "nodey_resample_tracks: null argument");
    NODEY_REQUIRE(ntracks >= 1 && ntracks <= kMaxResampleBatch, NODEY_E_RANGE, "nodey_resample_tracks: 1..%d tracks per call", kMaxResampleBatch);
    NODEY_REQUIRE(r->resample && tile2_ok(r), NODEY_E_RANGE, "nodey_resample_tracks: plan has no pipelined tile kernel (use nodey_resample_mix per track)");
    NODEY_REQUIRE(out_track_stride >= out_frames, NODEY_E_INVALID, "nodey_resample_tracks: out_track_stride smaller than out_frames");
    if (out_frames <= 0) return out_frames == 0 ? NODEY_OK : NODEY_E_INVALID;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    const int64_t avail = max_out_count(r, in_frames, flush);
    NODEY_REQUIRE(out_len >= 0 && out_len <= avail, NODEY_E_RANGE,
                  "nodey_resample_tracks: out_len %lld exceeds what swr would produce (%lld)", (long long)out_len, (long long)avail);
    int rc = fill_src(&a.src[0], r, plane0[0], plane1 ? plane1[0] : nullptr, fmt, nch, in_frames, flush, out_len);
    if (rc != NODEY_OK) return rc;
    a.out_len[0] = out_len; a.vol[0] = volumes[0]; a.nin = 1; a.mix = 1; a.out_frames = out_frames;
    a.ntracks = ntracks; a.out_track_stride = out_track_stride;
    static thread_local TrackPlanes tp;
    for (int t = 0; t < ntracks; t++) {
        NODEY_REQUIRE(plane0[t], NODEY_E_INVALID, "nodey_resample_tracks: null plane");
        tp.p0[t] = plane0[t]; tp.p1[t] = plane1 ? plane1[t] : nullptr; tp.vol[t] = volumes[t];
    }
    return launch_tile2(r, out_l, out_r, a, tp, nch, as_stream(stream));
}

/* time-segment sharding of one long stream: see nodey_cuda.h */
// chunk c of n: tiles [c * per, (c + 1) * per) of the pipelined kernel's tiling
static void resample_chunk_range(const nodey_resampler* r, int ch, int64_t out_frames, int c, int n, int64_t* tile0, int64_t* tile1,
                                 int64_t* per_tile_out, int64_t* in_need, int64_t in_frames)
{
    TileArgs a;
    memset(&a, 0, sizeof(a));
    size_t smem = 0;
    const int pt = tile2_shape(r, a, ch, 0, &smem);
    const int64_t per_tile = (int64_t)kNB * pt * a.P;
    const int64_t all = (out_frames + per_tile - 1) / per_tile;
    const int64_t per = (all + n - 1) / n > 0 ? (all + n - 1) / n : 1;
    *tile0 = (int64_t)c * per < all ? (int64_t)c * per : all;
    *tile1 = c >= n - 1 ? all : ((int64_t)(c + 1) * per < all ? (int64_t)(c + 1) * per : all);
    *per_tile_out = per_tile;
    // input a tile reads: in_tile frames (+ 2 for the aligned bulk copy) from tile * periods * D + s0 - center
    const int64_t in_tile = (int64_t)(kNB * pt - 1) * a.D + a.span + a.wmax + 1;
    const int64_t need = (*tile1 - 1) * (int64_t)kNB * pt * a.D + a.s0 - a.center + in_tile + 4;
    *in_need = (c >= n - 1 || need > in_frames) ? in_frames : (need < 0 ? 0 : need);
}

int nodey_resample_tracks_chunks(const nodey_resampler* r, int nch, int64_t in_frames, int64_t out_frames, int want_chunks,
                                 int64_t* in_need, int64_t* out_ready, int cap)
{
    NODEY_REQUIRE(r && (nch == 1 || nch == 2) && in_frames >= 0 && out_frames >= 0, NODEY_E_INVALID, "nodey_resample_tracks_chunks: bad argument");
    NODEY_REQUIRE(r->resample && tile2_ok(r), NODEY_E_RANGE, "nodey_resample_tracks_chunks: plan has no pipelined tile kernel");
    int64_t t0, t1, per_tile, need;
    resample_chunk_range(r, nch, out_frames, 0, 1, &t0, &t1, &per_tile, &need, in_frames);
    int n = want_chunks < 1 ? 1 : want_chunks;
    if ((int64_t)n > t1 / 4) n = (int)(t1 / 4);           // at least four tiles per launch and track
    if (n < 1) n = 1;
    if (t1 > 0) {                                         // no empty chunk at the end: ceil(tiles / ceil(tiles / n)) launches
        const int64_t per = (t1 + n - 1) / n;
        n = (int)((t1 + per - 1) / per);
    }
    for (int c = 0; c < n && c < cap; c++) {
        resample_chunk_range(r, nch, out_frames, c, n, &t0, &t1, &per_tile, &need, in_frames);
        if (in_need) in_need[c] = need;
        if (out_ready) out_ready[c] = (c == n - 1 || t1 * per_tile > out_frames) ? out_frames : t1 * per_tile;
    }
    return n;
}

int nodey_resample_tracks_chunk(const nodey_resampler* r, float* out_l, float* out_r, int64_t out_track_stride,
                                const void* const* plane0, const void* const* plane1, int fmt, int nch, int64_t in_frames,
                                const float* volumes, int ntracks, int flush, int64_t out_len, int64_t out_frames,
                                int chunk, int nchunks, nodey_stream_t stream)
{
    NODEY_REQUIRE(r && out_l && out_r && plane0 && volumes, NODEY_E_INVALID, "nodey_resample_tracks_chunk: null argument");
    NODEY_REQUIRE(ntracks >= 1 && ntracks <= kMaxResampleBatch, NODEY_E_RANGE, "nodey_resample_tracks_chunk: 1..%d tracks per call", kMaxResampleBatch);
    NODEY_REQUIRE(r->resample && tile2_ok(r), NODEY_E_RANGE, "nodey_resample_tracks_chunk: plan has no pipelined tile kernel (use nodey_resample_mix per track)");
    NODEY_REQUIRE(out_track_stride >= out_frames, NODEY_E_INVALID, "nodey_resample_tracks_chunk: out_track_stride smaller than out_frames");
    NODEY_REQUIRE(nchunks >= 1 && chunk >= 0 && chunk < nchunks, NODEY_E_RANGE, "nodey_resample_tracks_chunk: chunk %d of %d", chunk, nchunks);
    if (out_frames <= 0) return out_frames == 0 ? NODEY_OK : NODEY_E_INVALID;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    const int64_t avail = max_out_count(r, in_frames, flush);
    NODEY_REQUIRE(out_len >= 0 && out_len <= avail, NODEY_E_RANGE,
                  "nodey_resample_tracks_chunk: out_len %lld exceeds what swr would produce (%lld)", (long long)out_len, (long long)avail);
    int rc = fill_src(&a.src[0], r, plane0[0], plane1 ? plane1[0] : nullptr, fmt, nch, in_frames, flush, out_len);
    if (rc != NODEY_OK) return rc;
    a.out_len[0] = out_len; a.vol[0] = volumes[0]; a.nin = 1; a.mix = 1; a.out_frames = out_frames;
    a.ntracks = ntracks; a.out_track_stride = out_track_stride;
    int64_t per_tile, need;
    {
        int64_t t0, t1;
        resample_chunk_range(r, nch, out_frames, chunk, nchunks, &t0, &t1, &per_tile, &need, in_frames);
        if (t1 <= t0) return NODEY_OK;
        a.tile0 = t0; a.tile1 = t1;
    }
    static thread_local TrackPlanes tp;
    for (int t = 0; t < ntracks; t++) {
        NODEY_REQUIRE(plane0[t], NODEY_E_INVALID, "nodey_resample_tracks_chunk: null plane");
        tp.p0[t] = plane0[t]; tp.p1[t] = plane1 ? plane1[t] : nullptr; tp.vol[t] = volumes[t];
    }
    return launch_tile2(r, out_l, out_r, a, tp, nch, as_stream(stream));
}

int nodey_resampler_segment(const nodey_resampler* r, int64_t n_in, int64_t k0, int64_t k1,
                            int64_t* in0, int64_t* in1, int64_t* skip, int* flush)
{
    NODEY_REQUIRE(r && in0 && in1 && skip && flush, NODEY_E_INVALID, "nodey_resampler_segment: null argument");
    NODEY_REQUIRE(n_in >= 0 && k0 >= 0 && k1 >= k0, NODEY_E_INVALID, "nodey_resampler_segment: bad range");
    if (!r->resample) { *in0 = k0; *in1 = k1 < n_in ? k1 : n_in; *skip = 0; *flush = 0; return NODEY_OK; }
    NODEY_REQUIRE(r->dst_incr_mod == 0, NODEY_E_RANGE, "nodey_resampler_segment: only exact-rational plans can be cut into segments");
    const int64_t P = r->phase_count, D = r->dst_incr_div, L = r->filter_length, center = (L - 1) / 2;
    const int64_t total = nodey_resampler_out_count(r, n_in, 1);
    NODEY_REQUIRE(k1 <= total, NODEY_E_RANGE, "nodey_resampler_segment: k1 %lld beyond the %lld outputs of the stream", (long long)k1, (long long)total);
    NODEY_REQUIRE(k0 % P == 0, NODEY_E_INVALID, "nodey_resampler_segment: k0 must be a multiple of the phase count %lld", (long long)P);
    // one period of P outputs consumes exactly D input frames, so a conversion started a whole number of periods
    // before the segment reproduces the phases; the lead-in covers at least `center` input frames, so only
    // outputs that are dropped see the mirrored left edge
    int64_t lead = (center + D - 1) / D;
    if (lead < 1) lead = 1;
    const int64_t periods = k0 / P;
    *in0 = periods <= lead ? 0 : (periods - lead) * D;
    *skip = periods <= lead ? k0 : lead * P;
    *in1 = n_in; *flush = 1;
    if (k1 < total && k1 > k0) {
        // last output needs input frames up to s - center + L - 1 of the slice
        const int64_t s_last = plan_pos_index(r, *skip + (k1 - k0) - 1) / P;
        int64_t need = s_last - center + L;
        if (need < L + 1) need = L + 1;
        if (*in0 + need <= n_in) { *in1 = *in0 + need; *flush = 0; }
    }
    // a conversion cannot start before filter_length + 1 frames are there (the initial mirror): a slice that runs to the
    // end of a short stream and is shorter than that starts more whole periods early (tiny streams cut into many segments)
    if (*in0 > 0 && *in1 - *in0 < L + 1) {
        lead += ((L + 1) - (*in1 - *in0) + D - 1) / D;
        *in0 = periods <= lead ? 0 : (periods - lead) * D;
        *skip = periods <= lead ? k0 : lead * P;
    }
    return NODEY_OK;
}

/* audio_amix frame bookkeeping, audio-amix.cpp:149-322 (see nodey_cuda.h) */
int64_t nodey_amix_plan(const int* in_rate, int nin, const int64_t* run_off, const int64_t* run_len,
                        const int64_t* run_count, int index_mask_quirk,
                        int32_t* seg_input, int64_t* seg_out_start, int64_t* seg_src_start, int64_t* seg_len,
                        int64_t seg_cap, int64_t* nseg_out,
                        int64_t* out_run_len, int64_t* out_run_count, int64_t out_run_cap, int64_t* n_out_runs)
{
    if (!in_rate || !run_off || !run_len || !run_count || nin < 1 || nin > NODEY_MAX_MIX_INPUTS) {
        set_error("nodey_amix_plan: input_num %d outside 1..16 or null argument", nin);
        return NODEY_E_RANGE;
    }
    struct In {
        nodey_resampler plan;
        int64_t run = 0, run_end = 0, left = 0;      // cursor into the frame-size runs
        int64_t n_in = 0, produced = 0, reflect = 0;
        int flushed = 0;
    };
    std::vector<In> st((size_t)nin);
    for (int i = 0; i < nin; i++) {
        In& s = st[(size_t)i];
        if (in_rate[i] <= 0 || run_off[i + 1] < run_off[i]) { set_error("nodey_amix_plan: bad input %d", i); return NODEY_E_INVALID; }
        s.plan.in_rate = in_rate[i]; s.plan.out_rate = 48000;
        plan_init_math(&s.plan, index_mask_quirk);
        s.run = run_off[i]; s.run_end = run_off[i + 1];
        while (s.run < s.run_end && (run_count[s.run] <= 0 || run_len[s.run] <= 0)) s.run++;
        s.left = s.run < s.run_end ? run_count[s.run] : 0;
    }
    int64_t written = 0, nseg = 0, nruns = 0;
    std::vector<int64_t> last_seg((size_t)nin, -1);
    for (;;) {
        int64_t nb = INT64_MAX;
        for (int i = 0; i < nin; i++) {
            In& s = st[(size_t)i];
            if (s.run < s.run_end && run_len[s.run] < nb) nb = run_len[s.run];
        }
        if (nb == INT64_MAX) nb = 1152;
        int count = 0;
        for (int i = 0; i < nin; i++) {
            In& s = st[(size_t)i];
            int64_t got;
            if (s.run < s.run_end) {
                s.n_in += run_len[s.run];
                if (--s.left == 0) {
                    s.run++;
                    while (s.run < s.run_end && (run_count[s.run] <= 0 || run_len[s.run] <= 0)) s.run++;
                    s.left = s.run < s.run_end ? run_count[s.run] : 0;
                }
                const int64_t avail = s.plan.resample ? plan_producible(&s.plan, s.n_in, 0) : s.n_in;
                got = avail - s.produced;
                if (got > nb) got = nb;
            } else {
                if (!s.flushed) { s.flushed = 1; if (s.plan.resample) s.reflect = plan_reflect(&s.plan, s.n_in, s.produced); }
                const int64_t avail = s.plan.resample ? plan_producible(&s.plan, s.n_in, s.reflect) : s.n_in;
                got = avail - s.produced;
                if (got > nb) got = nb;
                if (got < nb) count++;
            }
            if (got < 0) got = 0;
            if (got > 0) {
                const int64_t ls = last_seg[(size_t)i];
                if (ls >= 0 && ls < seg_cap && seg_out_start[ls] + seg_len[ls] == written && seg_src_start[ls] + seg_len[ls] == s.produced) {
                    seg_len[ls] += got;
                } else {
                    if (nseg < seg_cap) { seg_input[nseg] = i; seg_out_start[nseg] = written; seg_src_start[nseg] = s.produced; seg_len[nseg] = got; }
                    last_seg[(size_t)i] = nseg;
                    nseg++;
                }
                s.produced += got;
            }
        }
        // the emitted frame of this iteration has nb samples: run-length encode the frame sizes
        if (nruns > 0 && nruns <= out_run_cap && out_run_len && out_run_len[nruns - 1] == nb) out_run_count[nruns - 1]++;
        else {
            if (out_run_len && nruns < out_run_cap) { out_run_len[nruns] = nb; out_run_count[nruns] = 1; }
            nruns++;
        }
        written += nb;
        if (count == nin) break;
    }
    if (nseg_out) *nseg_out = nseg;
    if (n_out_runs) *n_out_runs = nruns;
    return written;
}

}  // extern "C"
