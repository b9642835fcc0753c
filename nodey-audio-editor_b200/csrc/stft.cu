// stft.cu -- K8: the audio_spectrum node (N2; new node, FFTW r2c convention: unnormalised forward
// DFT, e^{-2 pi i k n / N}).  frame m = x[m*hop .. m*hop + 4096) * periodic Hann -> 2049 complex bins.
//
// One CTA of 128 threads owns one (frame, channel) item at a time and walks items grid-stride
// (persistent: grid = resident CTAs x SM count).  The 4096 real samples are packed as 2048 complex
// points z[n] = x[2n] + i x[2n+1]; a Stockham autosort FFT runs as radix 16 x 16 x 8 with 16 points per
// thread in registers:
//   pass 1  radix 16 (no twiddles), window fused into the coalesced global load
//   pass 2  radix 16, twiddles W_256
//   pass 3  radix 8, twiddles W_2048 -- thread t takes the butterflies t and 256 - t (thread 0: 0 and 128).
//           Output Z[jb + 256 r] of one is the mirror partner Z[2048 - k] of the other's Z[(256 - jb) + 256 (7 - r)],
//           so the real-input untangling X[k] = (Z[k] + Z*[M-k])/2 - (i/2) W_N^k (Z[k] - Z*[M-k]) happens in
//           registers, a pair (k, M-k) shares its sums, and the 2049 bins go straight to global memory.
// Two shared-memory exchanges and two barriers per item (the first version ran 8 x 8 x 8 x 4 with four exchanges,
// four barriers and a separate untangling read: 147 KB of shared traffic per item, now 64 KB; ncu had it latency
// bound at 34 % issue utilisation).  Nothing that is the same for every item is loaded (round 2; ncu had the L1 / shared
// data pipe at 71 % of its wavefront peak with twiddle and window loads a third of the wavefronts): five unit phasors per
// thread, read once from a double-precision table rounded to float, stay in registers, and the twiddles of passes 2 and 3,
// the untangling twiddles and the Hann window are computed from them per item (powers with chains of at most four
// multiplications, rotations by sixteenth roots of unity; measured error 3.7e-7 of the frame peak against the 1e-5 bar).
// Exchange slots are per-thread base pointers plus compile-time offsets; complex additions and subtractions are packed
// (add.rn.f32x2 / fma.rn.f32x2 with -1: the same roundings as the scalar forms, half the issue slots).
// HBM traffic per item: 4 KB of new input (each sample is reused by 4 overlapping frames out of L2) + 16.4 KB of output
// -> write-dominated.  A pure write stream reaches 3.9 TB/s on this GPU (memset / fill, tools/write_peak.py) against the
// 6.5 TB/s of a copy that counts read + write bytes, so the ceiling of this 1 : 4 read : write mix is about 4.9 TB/s =
// 75 % of the copy peak the roofline fraction is quoted against.
//
// This file is compiled WITHOUT -fmad=false: the oracle evaluates the DFT in double, so the float
// FFT is a tolerance comparison and fused multiply-adds only make it more accurate.
#include "nodey_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <mutex>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace nodey {

constexpr int kNfft = 4096;
constexpr int kM = kNfft / 2;        // complex points
constexpr int kBins = kNfft / 2 + 1;
constexpr int kThreads = 128;
// resident CTAs per SM (80 registers, 35 KB of shared memory each).  Measured, 10 min stereo: 4, 5 and 6 CTAs within 2 % of each
// other, also with the next item's samples prefetched into registers or the twiddle powers kept in registers across items
// (4 CTAs, 128 registers): the kernel is bound by its own instruction stream -- without its stores it takes 0.297 of
// 0.337 ms, without its loads 0.28 ms
constexpr int kStftResident = 6;

// one float2 of padding per 16 keeps the strided stores of the exchanges at the 2-wavefront minimum of a
// 256-byte warp access
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }
constexpr int kBufLen = kM + (kM >> 4);

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void fft8(float2 (&a)[8])
{
    const float s = 0.70710678118654752440f;
    float2 b[8];
#pragma unroll
    for (int i = 0; i < 4; i++) { b[i] = cadd(a[i], a[i + 4]); b[i + 4] = csub(a[i], a[i + 4]); }
    b[5] = make_float2(s * (b[5].x + b[5].y), s * (b[5].y - b[5].x));        // * (1 - i)/sqrt2
    b[6] = mul_mi(b[6]);
    b[7] = make_float2(s * (b[7].y - b[7].x), -s * (b[7].x + b[7].y));       // * (-1 - i)/sqrt2
    float2 c[8];
#pragma unroll
    for (int o = 0; o < 8; o += 4) {
        c[o + 0] = cadd(b[o + 0], b[o + 2]); c[o + 2] = csub(b[o + 0], b[o + 2]);
        c[o + 1] = cadd(b[o + 1], b[o + 3]); c[o + 3] = mul_mi(csub(b[o + 1], b[o + 3]));
    }
    a[0] = cadd(c[0], c[1]); a[4] = csub(c[0], c[1]); a[2] = cadd(c[2], c[3]); a[6] = csub(c[2], c[3]);
    a[1] = cadd(c[4], c[5]); a[5] = csub(c[4], c[5]); a[3] = cadd(c[6], c[7]); a[7] = csub(c[6], c[7]);
}

__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
    const float2 s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));
    a0 = cadd(s02, s13); a2 = csub(s02, s13); a1 = cadd(d02, d13); a3 = csub(d02, d13);
}

// forward 16-point DFT, natural order in and out: n = n1 + 4 n2, k = 4 k1 + k2;
// W16^{nk} = W4^{n2 k2} * W16^{n1 k2} * W4^{n1 k1}
__device__ __forceinline__ void fft16(float2 (&a)[16])
{
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    // 4-point DFTs over n2 (stride 4): a[n1 + 4 k2] = B[n1][k2]
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) fft4(a[n1], a[n1 + 4], a[n1 + 8], a[n1 + 12]);
    // twiddles W16^{n1 k2}
    a[5] = cmul(a[5], make_float2(c1, -s1));          // W16^1
    a[9] = cmul(a[9], make_float2(h, -h));            // W16^2
    a[13] = cmul(a[13], make_float2(s1, -c1));        // W16^3
    a[6] = cmul(a[6], make_float2(h, -h));            // W16^2
    a[10] = mul_mi(a[10]);                            // W16^4
    a[14] = cmul(a[14], make_float2(-h, -h));         // W16^6
    a[7] = cmul(a[7], make_float2(s1, -c1));          // W16^3
    a[11] = cmul(a[11], make_float2(-h, -h));         // W16^6
    a[15] = cmul(a[15], make_float2(-c1, s1));        // W16^9
    // 4-point DFTs over n1: X[4 k1 + k2] = sum_n1 a[n1 + 4 k2] W4^{n1 k1}
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) fft4(a[4 * k2], a[4 * k2 + 1], a[4 * k2 + 2], a[4 * k2 + 3]);
    // a[4 k2 + k1] holds X[4 k1 + k2]: transpose into natural order
    float2 t;
    t = a[1]; a[1] = a[4]; a[4] = t;
    t = a[2]; a[2] = a[8]; a[8] = t;
    t = a[3]; a[3] = a[12]; a[12] = t;
    t = a[6]; a[6] = a[9]; a[9] = t;
    t = a[7]; a[7] = a[13]; a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}

__device__ __forceinline__ float2 csqr(float2 a) { return make_float2(a.x * a.x - a.y * a.y, (a.x + a.x) * a.y); }
__host__ __device__ constexpr int top_bit(int r) { int h = 1; while (2 * h <= r) h *= 2; return h; }
// powers b^1 .. b^(N-1) of a unit phasor: even powers by squaring, odd ones as b^(top bit) * b^(rest), so no chain is longer
// than four multiplications for N = 16 (a power is within a few ulp of the rounded exact value; the 1e-5 bar of the
// spectrum is against the frame's peak).  p[0] is not used.
template <int N>
__device__ __forceinline__ void cpowers(float2 b, float2 (&p)[N])
{
    p[1] = b;
#pragma unroll
    for (int r = 2; r < N; r++) p[r] = (r & 1) ? cmul(p[top_bit(r)], p[r - top_bit(r)]) : csqr(p[r / 2]);
}

// a * e^{-i pi q / 8} for a compile-time q (the sixteenth roots of unity): trivial factors cost no multiplication
template <int Q>
__device__ __forceinline__ float2 mul_w16(float2 a)
{
    constexpr int q = ((Q % 16) + 16) % 16;
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    if constexpr (q == 0) return a;
    else if constexpr (q == 4) return mul_mi(a);
    else if constexpr (q == 8) return make_float2(-a.x, -a.y);
    else if constexpr (q == 12) return make_float2(-a.y, a.x);
    else if constexpr (q == 2) return make_float2(h * (a.x + a.y), h * (a.y - a.x));
    else if constexpr (q == 6) return make_float2(h * (a.y - a.x), -h * (a.x + a.y));
    else if constexpr (q == 10) return make_float2(-h * (a.x + a.y), h * (a.x - a.y));
    else if constexpr (q == 14) return make_float2(h * (a.x - a.y), h * (a.x + a.y));
    else {
        // e^{-i pi q / 8} = (cos, -sin)
        constexpr float cs[16] = {1.f, c1, h, s1, 0.f, -s1, -h, -c1, -1.f, -c1, -h, -s1, 0.f, s1, h, c1};
        constexpr float sn[16] = {0.f, s1, h, c1, 1.f, c1, h, s1, 0.f, -s1, -h, -c1, -1.f, -c1, -h, -s1};
        return cmul(a, make_float2(cs[q], -sn[q]));
    }
}

struct StftArgs {
    float2* out;            // [nch][frames][kBins]
    const float* x;
    long long ch_stride;    // elements between channel planes (planar) or 1 (interleaved)
    int x_stride;           // elements between consecutive samples of one channel
    int nch, hop;
    long long frames;       // per channel
    const float2* tw;       // W_4096^i, i < 4096
    const float* win;       // periodic Hann, float
    int vec_ok;             // float2 loads allowed (x_stride == 1, 8-byte aligned, even hop)
    int inter2;             // interleaved stereo, 16-byte aligned, even hop: one float4 load = two frames of both channels
};

__device__ __forceinline__ void st_bin(float2* p, float2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// bins k and M - k from zk = Z[k] and zmc = conj(Z[M - k]): e = zk + zmc, t = W_N^k (zk - zmc),
// X[k] = (e - i t)/2, X[M - k] = conj-mirrored combination of the same e and t
__device__ __forceinline__ void untangle_pair(float2 zk, float2 zmc, float2 w, float2& xk, float2& xm)
{
    const float2 e = cadd(zk, zmc), t = cmul(w, csub(zk, zmc));
    xk = make_float2(0.5f * (e.x + t.y), 0.5f * (e.y - t.x));
    xm = make_float2(0.5f * (e.x - t.y), -0.5f * (e.y + t.x));
}

// periodic Hann value at n = n0 + 256 R (n0 = 2j or 2j + 1): 0.5 - 0.5 cos(theta0 + R pi/8) = 0.5 + ce C_R + se S_R with
// ce = -0.5 cos(theta0), se = 0.5 sin(theta0) held per thread -- two fused multiply-adds instead of a load per sample
template <int R>
__device__ __forceinline__ float hann_at(float ce, float se)
{
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    constexpr float C[16] = {1.f, c1, h, s1, 0.f, -s1, -h, -c1, -1.f, -c1, -h, -s1, 0.f, s1, h, c1};
    constexpr float S[16] = {0.f, s1, h, c1, 1.f, c1, h, s1, 0.f, -s1, -h, -c1, -1.f, -c1, -h, -s1};
    if constexpr (R == 0) return 0.5f + ce;
    else if constexpr (R == 8) return 0.5f - ce;
    else if constexpr (R == 4) return 0.5f + se;
    else if constexpr (R == 12) return 0.5f - se;
    else return fmaf(se, S[R], fmaf(ce, C[R], 0.5f));
}

template <int VEC, int R>
__device__ __forceinline__ float2 load_point(const StftArgs& a, const float* p, long long frame, int c, int j)
{
    const int n = 2 * (j + 128 * R);
    if (VEC == 1) return __ldg(reinterpret_cast<const float2*>(p + n));
    if (VEC == 2) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.x + (frame * (long long)a.hop + n) * 2));
        return c ? make_float2(q.y, q.w) : make_float2(q.x, q.z);
    }
    return make_float2(__ldg(p + n * a.x_stride), __ldg(p + (n + 1) * a.x_stride));
}

template <int VEC, int R = 0>
__device__ __forceinline__ void load_raw(float2 (&v)[16], const StftArgs& a, const float* p, long long frame, int c, int j)
{
    if constexpr (R < 16) {
        v[R] = load_point<VEC, R>(a, p, frame, c, j);
        load_raw<VEC, R + 1>(v, a, p, frame, c, j);
    }
}

template <int R = 0>
__device__ __forceinline__ void apply_window(float2 (&v)[16], const float2 (&raw)[16], float ce0, float se0, float ce1, float se1)
{
    if constexpr (R < 16) {
        v[R] = make_float2(raw[R].x * hann_at<R>(ce0, se0), raw[R].y * hann_at<R>(ce1, se1));
        apply_window<R + 1>(v, raw, ce0, se0, ce1, se1);
    }
}

template <int R = 1>
__device__ __forceinline__ void twiddle_mirror(float2 (&w)[8], const float2 (&p3)[8], bool first)
{
    // butterfly 256 - j: W_2048^{(256 - j) r} = e^{-i pi r / 4} conj(W_2048^{j r});  thread 0 takes butterfly 128: e^{-i pi r / 8}
    if constexpr (R < 8) {
        w[R] = first ? mul_w16<R>(w[R]) : cmul(w[R], mul_w16<2 * R>(cconj(p3[R])));
        twiddle_mirror<R + 1>(w, p3, first);
    }
}

template <int R = 0>
__device__ __forceinline__ void untangle_store(float2* o, const float2 (&u)[8], const float2 (&w)[8], float2 bu, int j)
{
    // Z[j + 256 r] = u[r];  its partner Z[2048 - j - 256 r] = Z[(256 - j) + 256 (7 - r)] = w[7 - r];  W_4096^{j + 256 r} = W_4096^j e^{-i pi r / 8}
    if constexpr (R < 8) {
        const int k = j + 256 * R;
        float2 xk, xm;
        untangle_pair(u[R], cconj(w[7 - R]), mul_w16<R>(bu), xk, xm);
        st_bin(o + k, xk);
        st_bin(o + (kM - k), xm);
        untangle_store<R + 1>(o, u, w, bu, j);
    }
}

// VEC 1: channel samples are contiguous and 8-byte aligned (planar input): one 64-bit load per complex point
// VEC 2: interleaved stereo: one 128-bit load holds the complex point of BOTH channels; a CTA transforms the two channels
//        of a frame back to back, so the second channel's loads hit the lines the first one brought into L1
//        (the scalar path issues two strided 32-bit loads per point: twice the instructions and wavefronts)
//
// The kernel is bound by the L1 / shared-memory data pipe (ncu: 71 % of its wavefront peak at 47 % issue utilisation, with
// twiddle and window loads a third of the wavefronts), so everything that does not depend on the item is COMPUTED from five
// per-thread phasors held in registers instead of loaded: the twiddles of pass 2 (powers of W_256^k), of pass 3 (powers of
// W_2048^j; the mirror butterfly's are their conjugates times an eighth root of unity), of the untangling (W_4096^j times
// a sixteenth root of unity) and the Hann window (a rotation by r pi/8 of the thread's cos / sin pair).
template <int VEC>
__global__ void __launch_bounds__(kThreads, kStftResident) stft4096_kernel(const __grid_constant__ StftArgs a)
{
    extern __shared__ __align__(16) float2 stft_smem[];      // 35 KB
    float2* bufA = stft_smem;
    float2* bufB = bufA + kBufLen;
    const int j = threadIdx.x;
    const long long items = a.frames * a.nch;
    const float2 b2 = __ldg(a.tw + (j & 15) * 16);           // W_256^k, k = j & 15
    const float2 b3 = __ldg(a.tw + 2 * j);                   // W_2048^j = (cos, -sin)(2 pi 2j / 4096): also the window phasor of n = 2j
    const float2 bu = __ldg(a.tw + j);                       // W_4096^j
    const float2 b1 = __ldg(a.tw + 2 * j + 1);               // window phasor of n = 2j + 1
    const float ce0 = -0.5f * b3.x, se0 = -0.5f * b3.y, ce1 = -0.5f * b1.x, se1 = -0.5f * b1.y;
    // padded exchange slots as base + compile-time offset: pad(i) = i + (i >> 4), and every index below is a per-thread
    // constant plus a multiple of 16 (pad(c + 16 m) = pad(c) + 17 m when c < 16 ... in general the shift distributes because
    // the per-thread parts are multiples of 16 or stay below 16)
    const int jb0 = j, jb1 = j ? 256 - j : 128;
    float2* const s1 = bufA + 17 * j;                                   // pass 1 stores: pad(16 j + r) = 17 j + r
    const float2* const l2 = bufA + j + (j >> 4);                       // pass 2 loads:  pad(j + 128 r) = pad(j) + 136 r
    float2* const s2 = bufB + (j & ~15) * 17 + (j & 15);                // pass 2 stores: pad((j & ~15) 16 + k + 16 r) = 17 (j & ~15) + k + 17 r
    const float2* const l3a = bufB + jb0 + (jb0 >> 4);                  // pass 3 loads:  pad(jb + 256 r) = pad(jb) + 272 r
    const float2* const l3b = bufB + jb1 + (jb1 >> 4);

    // item -> (frame, channel): channel fastest within a CTA's own sequence of items when VEC == 2 (see above), otherwise
    // across CTAs (consecutive CTAs take the channels of one frame)
    const auto locate = [&](long long it, long long& frame, int& c) -> bool {
        if (it >= items) return false;
        long long item = it;
        if (VEC == 2) {
            // CTA b owns frames b, b + grid, ...; `it` walks 2 * (its frames)
            const long long local = it - blockIdx.x;                 // 0, 1, 2, ... within this CTA
            const long long frame_v = blockIdx.x + (local >> 1) * (long long)gridDim.x;
            if (frame_v >= a.frames) return false;
            item = frame_v * 2 + (local & 1);
        }
        frame = item / a.nch;
        c = (int)(item - frame * a.nch);
        return true;
    };
    const auto fetch = [&](float2 (&raw)[16], long long frame, int c) {
        load_raw<VEC>(raw, a, a.x + (long long)c * a.ch_stride + frame * (long long)a.hop * a.x_stride, frame, c, j);
    };
    const long long step = VEC == 2 ? 1 : gridDim.x;
    long long it = blockIdx.x, frame = 0; int c = 0;
    bool valid = locate(it, frame, c);
    while (valid) {
        float2 v[16];
        // ---- pass 1: radix 16, Ns = 1 (no twiddles); window applied to the loaded samples ----
        {
            float2 raw[16];
            fetch(raw, frame, c);
            apply_window(v, raw, ce0, se0, ce1, se1);
        }
        fft16(v);
#pragma unroll
        for (int r = 0; r < 16; r++) s1[r] = v[r];
        __syncthreads();
        const long long frame_cur = frame; const int c_cur = c;
        it += step;
        valid = locate(it, frame, c);

        // ---- pass 2: radix 16, Ns = 16: twiddles W_256^{k r} ----
        {
            float2 p2[16];
            cpowers<16>(b2, p2);
#pragma unroll
            for (int r = 0; r < 16; r++) {
                v[r] = l2[136 * r];
                if (r) v[r] = cmul(v[r], p2[r]);
            }
            fft16(v);
#pragma unroll
            for (int r = 0; r < 16; r++) s2[17 * r] = v[r];
        }
        __syncthreads();

        // ---- pass 3: radix 8, Ns = 256: twiddles W_2048^{jb r}; butterflies jb0 = j and its mirror jb1 ----
        float2 u[8], w[8];
        {
            float2 p3[8];
            cpowers<8>(b3, p3);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                u[r] = l3a[272 * r];
                w[r] = l3b[272 * r];
                if (r) u[r] = cmul(u[r], p3[r]);
            }
            twiddle_mirror(w, p3, j == 0);
        }
        fft8(u);
        fft8(w);
        // ---- real-input untangling in registers, bins straight to global memory ----
        float2* o = a.out + ((long long)c_cur * a.frames + frame_cur) * kBins;
        if (j) {
            untangle_store(o, u, w, bu, j);
        } else {
            // butterfly 0: Z[256 r] = u[r], partner Z[256 (8 - r)]; bins 0 and 2048 are real; bin 1024 is its own partner
            st_bin(o, make_float2(u[0].x + u[0].y, 0.f));
            st_bin(o + kM, make_float2(u[0].x - u[0].y, 0.f));        // k = M: W_N^M = -1
#pragma unroll
            for (int r = 1; r < 4; r++) {
                float2 xk, xm;
                untangle_pair(u[r], cconj(u[8 - r]), __ldg(a.tw + 256 * r), xk, xm);
                st_bin(o + 256 * r, xk);
                st_bin(o + (kM - 256 * r), xm);
            }
            {
                float2 xk, xm;
                untangle_pair(u[4], cconj(u[4]), __ldg(a.tw + 1024), xk, xm);
                st_bin(o + 1024, xk);
            }
            // butterfly 128: Z[128 + 256 r] = w[r], partner Z[128 + 256 (7 - r)] = w[7 - r]
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int k = 128 + 256 * r;
                float2 xk, xm;
                untangle_pair(w[r], cconj(w[7 - r]), __ldg(a.tw + k), xk, xm);
                st_bin(o + k, xk);
                st_bin(o + (kM - k), xm);
            }
        }
        // no barrier needed here: the next item's pass 1 only writes bufA (last read before the second barrier
        // above), and its first barrier orders every thread's bufB reads above before pass 2 overwrites bufB
    }
}

struct StftTables { float2* tw = nullptr; float* win = nullptr; };

static int get_tables(StftTables* out)
{
    static std::mutex mu;
    static StftTables per_dev[64];
    int dev = 0;
    NODEY_CUDA_OK(cudaGetDevice(&dev));
    NODEY_REQUIRE(dev >= 0 && dev < 64, NODEY_E_INVALID, "stft: device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!per_dev[dev].tw) {
        static float2 h_tw[kNfft];
        static float h_win[kNfft];
        for (int i = 0; i < kNfft; i++) {
            const double ang = -2.0 * M_PI * (double)i / (double)kNfft;
            h_tw[i] = make_float2((float)cos(ang), (float)sin(ang));
            h_win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)kNfft));
        }
        float2* tw = nullptr; float* win = nullptr;
        NODEY_CUDA_OK(cudaMalloc((void**)&tw, sizeof(h_tw)));
        NODEY_CUDA_OK(cudaMalloc((void**)&win, sizeof(h_win)));
        NODEY_CUDA_OK(cudaMemcpy(tw, h_tw, sizeof(h_tw), cudaMemcpyHostToDevice));
        NODEY_CUDA_OK(cudaMemcpy(win, h_win, sizeof(h_win), cudaMemcpyHostToDevice));
        per_dev[dev].tw = tw; per_dev[dev].win = win;
    }
    *out = per_dev[dev];
    return NODEY_OK;
}

}  // namespace nodey

using namespace nodey;

extern "C" {

int64_t nodey_stft_frames(int64_t nframes, int nfft, int hop)
{
    if (nfft <= 0 || hop <= 0 || nframes < nfft) return 0;
    return (nframes - nfft) / hop + 1;
}

int nodey_stft(float* out_complex, const float* x, int64_t nframes, int nch, int interleaved, int64_t plane_stride,
               int nfft, int hop, nodey_stream_t stream)
{
    NODEY_REQUIRE(nfft == kNfft, NODEY_E_RANGE, "nodey_stft: fft_size %d not supported (4096 only)", nfft);
    NODEY_REQUIRE(hop > 0, NODEY_E_RANGE, "nodey_stft: hop must be positive");
    NODEY_REQUIRE(nch >= 1 && nch <= 8, NODEY_E_INVALID, "nodey_stft: bad channel count %d", nch);
    NODEY_REQUIRE(nframes >= 0, NODEY_E_INVALID, "nodey_stft: negative size");
    const int64_t frames = nodey_stft_frames(nframes, nfft, hop);
    if (frames == 0) return NODEY_OK;
    NODEY_REQUIRE(out_complex && x, NODEY_E_INVALID, "nodey_stft: null buffer");
    StftTables t;
    int rc = get_tables(&t);
    if (rc != NODEY_OK) return rc;
    StftArgs a;
    a.out = reinterpret_cast<float2*>(out_complex);
    a.x = x;
    a.nch = nch; a.hop = hop; a.frames = frames;
    a.x_stride = interleaved ? nch : 1;
    a.ch_stride = interleaved ? 1 : plane_stride;
    a.tw = t.tw; a.win = t.win;
    a.vec_ok = (a.x_stride == 1) && (((uintptr_t)x & 7) == 0) && (hop % 2 == 0) && (a.ch_stride % 2 == 0);
    a.inter2 = interleaved && nch == 2 && (((uintptr_t)x & 15) == 0) && (hop % 2 == 0);
    const int64_t items = frames * nch;
    const int64_t cap = (int64_t)sm_count() * kStftResident;
    const int grid = (int)(items < cap ? items : cap);
    constexpr size_t smem = sizeof(float2) * (2 * kBufLen);
    void (*kern)(StftArgs) = a.inter2 ? stft4096_kernel<2> : (a.vec_ok ? stft4096_kernel<1> : stft4096_kernel<0>);
    NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    NODEY_LAUNCH("stft4096_kernel", as_stream(stream), kern<<<grid, kThreads, smem, as_stream(stream)>>>(a));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

}  // extern "C"
