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
// bound at 34 % issue utilisation).  Twiddles and window come from double-precision tables rounded once to
// float; the twiddles of passes 2 and 3 are re-laid out per CTA in shared memory (16 KB), window and untangling
// twiddles are read coalesced through L1.  HBM traffic per item: 4 KB of new input (each sample is reused by 4 overlapping
// frames out of L2) + 16.4 KB of output -> write-dominated.
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
constexpr int kThreads = 128;

// one float2 of padding per 16 keeps the strided stores of the exchanges at the 2-wavefront minimum of a
// 256-byte warp access
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }
constexpr int kBufLen = kM + (kM >> 4);

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
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

// VEC 1: channel samples are contiguous and 8-byte aligned (planar input): one 64-bit load per complex point
// VEC 2: interleaved stereo: one 128-bit load holds the complex point of BOTH channels; a CTA transforms the two channels
//        of a frame back to back, so the second channel's loads hit the lines the first one brought into L1
//        (the scalar path issues two strided 32-bit loads per point: twice the instructions and wavefronts)
template <int VEC>
__global__ void __launch_bounds__(kThreads, 4) stft4096_kernel(const __grid_constant__ StftArgs a)
{
    extern __shared__ __align__(16) float2 stft_smem[];      // 51 KB: above the static limit
    float2* bufA = stft_smem;
    float2* bufB = bufA + kBufLen;
    float2 (*tw2)[16] = reinterpret_cast<float2 (*)[16]>(bufB + kBufLen);
    float2 (*tw3)[256] = reinterpret_cast<float2 (*)[256]>(bufB + kBufLen + 15 * 16);
    // twiddles of passes 2 and 3 by (r, butterfly): consecutive threads read consecutive words.  Read straight from
    // the W_4096 table they were strided by 16 r resp. 2 r entries across a warp -- one L1 wavefront per lane, and ncu
    // showed the L1 data pipe 91 % busy with the kernel at 35 % issue utilisation.
    // tw2[15][16]: W_256^{k r}, r = 1..15, k < 16;  tw3[7][256]: W_2048^{jb r}, r = 1..7, jb < 256
    const int j = threadIdx.x;
    const long long items = a.frames * a.nch;
    for (int i = j; i < 15 * 16; i += kThreads) { const int r = i / 16 + 1, k = i % 16; tw2[r - 1][k] = a.tw[k * r * 16]; }
    for (int i = j; i < 7 * 256; i += kThreads) { const int r = i / 256 + 1, jb = i % 256; tw3[r - 1][jb] = a.tw[jb * r * 2]; }
    // (the first barrier of the item loop orders these stores before their first use in pass 2)

    // item -> (frame, channel): channel fastest within a CTA's own sequence of items when VEC == 2 (see above), otherwise
    // across CTAs (consecutive CTAs take the channels of one frame)
    for (long long it = blockIdx.x; it < items; it += (VEC == 2 ? 1 : gridDim.x)) {
        long long item = it;
        if (VEC == 2) {
            // CTA b owns frames b, b + grid, ...; `it` walks 2 * (its frames)
            const long long local = it - blockIdx.x;                 // 0, 1, 2, ... within this CTA
            const long long frame_v = blockIdx.x + (local >> 1) * (long long)gridDim.x;
            if (frame_v >= a.frames) break;
            item = frame_v * 2 + (local & 1);
        }
        const long long frame = item / a.nch;
        const int c = (int)(item - frame * a.nch);
        const float* p = a.x + (long long)c * a.ch_stride + frame * (long long)a.hop * a.x_stride;

        float2 v[16];
        // ---- pass 1: radix 16, Ns = 1 (no twiddles); window fused into the load ----
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int n = 2 * (j + 128 * r);
            float2 s;
            if (VEC == 1) s = __ldg(reinterpret_cast<const float2*>(p + n));
            else if (VEC == 2) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(a.x + (frame * (long long)a.hop + n) * 2));
                s = c ? make_float2(q.y, q.w) : make_float2(q.x, q.z);
            }
            else { s.x = __ldg(p + n * a.x_stride); s.y = __ldg(p + (n + 1) * a.x_stride); }
            const float2 w = __ldg(reinterpret_cast<const float2*>(a.win + n));
            v[r] = make_float2(__fmul_rn(s.x, w.x), __fmul_rn(s.y, w.y));
        }
        fft16(v);
#pragma unroll
        for (int r = 0; r < 16; r++) bufA[pad(16 * j + r)] = v[r];
        __syncthreads();

        // ---- pass 2: radix 16, Ns = 16: W_256^{k r} = W_4096^{16 k r} ----
        {
            const int k = j & 15;
#pragma unroll
            for (int r = 0; r < 16; r++) {
                v[r] = bufA[pad(j + 128 * r)];
                if (r) v[r] = cmul(v[r], tw2[r - 1][k]);
            }
            fft16(v);
            const int base = (j & ~15) * 16 + k;
#pragma unroll
            for (int r = 0; r < 16; r++) bufB[pad(base + 16 * r)] = v[r];
        }
        __syncthreads();

        // ---- pass 3: radix 8, Ns = 256: W_2048^{jb r} = W_4096^{2 jb r}; butterflies jb0 and its mirror jb1 ----
        const int jb0 = j, jb1 = j ? 256 - j : 128;
        float2 u[8], w[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            u[r] = bufB[pad(jb0 + 256 * r)];
            w[r] = bufB[pad(jb1 + 256 * r)];
            if (r) { u[r] = cmul(u[r], tw3[r - 1][jb0]); w[r] = cmul(w[r], tw3[r - 1][jb1]); }
        }
        fft8(u);
        fft8(w);
        // ---- real-input untangling in registers, bins straight to global memory ----
        float2* o = a.out + ((long long)c * a.frames + frame) * kBins;
        if (j) {
            // Z[j + 256 r] = u[r];  its partner Z[2048 - j - 256 r] = Z[(256 - j) + 256 (7 - r)] = w[7 - r]
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int k = j + 256 * r;
                float2 xk, xm;
                untangle_pair(u[r], cconj(w[7 - r]), __ldg(a.tw + k), xk, xm);
                st_bin(o + k, xk);
                st_bin(o + (kM - k), xm);
            }
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
    const int64_t cap = (int64_t)sm_count() * 4;
    const int grid = (int)(items < cap ? items : cap);
    constexpr size_t smem = sizeof(float2) * (2 * kBufLen + 15 * 16 + 7 * 256);
    void (*kern)(StftArgs) = a.inter2 ? stft4096_kernel<2> : (a.vec_ok ? stft4096_kernel<1> : stft4096_kernel<0>);
    NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    NODEY_LAUNCH("stft4096_kernel", as_stream(stream), kern<<<grid, kThreads, smem, as_stream(stream)>>>(a));
    NODEY_LAUNCH_OK();
    return NODEY_OK;
}

}  // extern "C"
