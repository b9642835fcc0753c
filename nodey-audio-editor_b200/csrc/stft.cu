// stft.cu -- K8: the audio_spectrum node (N2; new node, FFTW r2c convention: unnormalised forward
// DFT, e^{-2 pi i k n / N}).  frame m = x[m*hop .. m*hop + 4096) * periodic Hann -> 2049 complex bins.
//
// One CTA (256 threads) owns one (frame, channel) item at a time and walks items grid-stride
// (persistent: grid = resident CTAs x SM count).  The 4096 real samples are packed as 2048 complex
// points z[n] = x[2n] + i x[2n+1]; a Stockham autosort FFT runs as radix 8 x 8 x 8 x 4 passes with
// the butterflies in registers and the exchanges through two padded shared-memory buffers; the
// real-input untangling X[k] = (Z[k] + Z*[M-k])/2 - (i/2) W_N^k (Z[k] - Z*[M-k]) is fused into the
// coalesced store of the 2049 bins.  Twiddles and window come from double-precision tables
// rounded once to float (L1-resident, 48 KB).  HBM traffic per item: 4 KB of new input (each
// sample is reused by 4 overlapping frames out of L2) + 16.4 KB of output -> write-dominated.
//
// This file is compiled WITHOUT -fmad=false: the oracle evaluates the DFT in double, so the float
// FFT is a tolerance comparison and fused multiply-adds only make it more accurate.
#include "nodey_common.cuh"

#include <math.h>
#include <mutex>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace nodey {

constexpr int kNfft = 4096;
constexpr int kM = kNfft / 2;        // complex points
constexpr int kBins = kNfft / 2 + 1;
constexpr int kThreads = 256;

// one float2 of padding per 16 keeps the strided stores of the first two passes at the 2-wavefront
// minimum (STFT bank analysis)
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }
constexpr int kBufLen = kM + (kM >> 4);

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

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

__device__ __forceinline__ void fft4(float2 (&a)[4])
{
    const float2 s02 = cadd(a[0], a[2]), d02 = csub(a[0], a[2]);
    const float2 s13 = cadd(a[1], a[3]), d13 = mul_mi(csub(a[1], a[3]));
    a[0] = cadd(s02, s13); a[2] = csub(s02, s13); a[1] = cadd(d02, d13); a[3] = csub(d02, d13);
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
};

__global__ void __launch_bounds__(kThreads, 3) stft4096_kernel(const __grid_constant__ StftArgs a)
{
    __shared__ float2 bufA[kBufLen];
    __shared__ float2 bufB[kBufLen];
    const int j = threadIdx.x;
    const long long items = a.frames * a.nch;

    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long frame = item / a.nch;
        const int c = (int)(item - frame * a.nch);
        const float* p = a.x + (long long)c * a.ch_stride + frame * (long long)a.hop * a.x_stride;

        float2 v[8];
        // ---- pass 1: radix 8, Ns = 1 (no twiddles); window fused into the load ----
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int n = 2 * (j + 256 * r);
            float2 s;
            if (a.vec_ok) s = __ldg(reinterpret_cast<const float2*>(p + n));
            else { s.x = __ldg(p + (long long)n * a.x_stride); s.y = __ldg(p + (long long)(n + 1) * a.x_stride); }
            const float2 w = __ldg(reinterpret_cast<const float2*>(a.win + n));
            v[r] = make_float2(__fmul_rn(s.x, w.x), __fmul_rn(s.y, w.y));
        }
        fft8(v);
#pragma unroll
        for (int r = 0; r < 8; r++) bufA[pad(8 * j + r)] = v[r];
        __syncthreads();

        // ---- pass 2: radix 8, Ns = 8 ----
        {
            const int k = j & 7;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                v[r] = bufA[pad(j + 256 * r)];
                if (r) v[r] = cmul(v[r], __ldg(a.tw + k * r * 64));
            }
            fft8(v);
            const int base = (j & ~7) * 8 + k;
#pragma unroll
            for (int r = 0; r < 8; r++) bufB[pad(base + 8 * r)] = v[r];
        }
        __syncthreads();

        // ---- pass 3: radix 8, Ns = 64 ----
        {
            const int k = j & 63;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                v[r] = bufB[pad(j + 256 * r)];
                if (r) v[r] = cmul(v[r], __ldg(a.tw + k * r * 8));
            }
            fft8(v);
            const int base = (j & ~63) * 8 + k;
#pragma unroll
            for (int r = 0; r < 8; r++) bufA[pad(base + 64 * r)] = v[r];
        }
        __syncthreads();

        // ---- pass 4: radix 4, Ns = 512 (two butterflies per thread) ----
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = j + 256 * h;
            float2 u[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                u[r] = bufA[pad(k + 512 * r)];
                if (r) u[r] = cmul(u[r], __ldg(a.tw + k * r * 2));
            }
            fft4(u);
#pragma unroll
            for (int r = 0; r < 4; r++) bufB[pad(k + 512 * r)] = u[r];
        }
        __syncthreads();

        // ---- real-input untangling fused with the store ----
        float2* o = a.out + ((long long)c * a.frames + frame) * kBins;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = j + 256 * q;
            const float2 zk = bufB[pad(k)];
            float2 zm = bufB[pad((kM - k) & (kM - 1))];
            zm.y = -zm.y;
            const float2 e = cadd(zk, zm), d = csub(zk, zm);
            const float2 t = cmul(__ldg(a.tw + k), d);        // W_N^k * (Z[k] - Z*[M-k])
            // X = e/2 - (i/2) t  ->  (e.x + t.y)/2 , (e.y - t.x)/2
            const float2 X = make_float2(0.5f * (e.x + t.y), 0.5f * (e.y - t.x));
            asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(o + k), "f"(X.x), "f"(X.y) : "memory");
        }
        if (j == 0) {
            const float2 z0 = bufB[pad(0)];
            o[kM] = make_float2(z0.x - z0.y, 0.f);            // k = M: W_N^M = -1
        }
        // no barrier needed here: the next item's pass 1 only writes bufA, and its first barrier
        // orders every thread's bufB reads above before pass 2 overwrites bufB
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
    const int64_t items = frames * nch;
    const int64_t cap = (int64_t)sm_count() * 3;
    const int grid = (int)(items < cap ? items : cap);
    NODEY_LAUNCH("stft4096_kernel", as_stream(stream), stft4096_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(a));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

}  // extern "C"
