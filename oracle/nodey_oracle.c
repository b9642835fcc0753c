/*
 * nodey_oracle.c -- CPU restatement of the reference processors (see nodey_oracle.h header).
 * TEST INFRASTRUCTURE ONLY.  Pinning status (the reference ships no tests or fixtures):
 *   - libswresample model: PINNED against a real libswresample (6.1.100, FFmpeg 8.0.1, found in this image;
 *     oracle/real_swr.py, tests/golden/swr_real.npz, tests/test_swr_real.py): filter bank / positions / edges
 *     bit exact through impulse responses, per-call sample counts exact, values within 1e-6, the audio_amix
 *     loop and the preview conversion end to end;
 *   - SoundTouch model (TDStretch, AAFilter, InterpolateCubic): PARITY UNPINNED -- no SoundTouch binary or
 *     source exists in the image; restated from upstream knowledge, behavioural pins only;
 *   - in-tree arithmetic (gain, mixers, extraction, alignment): restated literally from src/processor/.
 * Citations are relative to /root/reference.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  -ffp-contract=off matters:
 * the reference builds at -O3 without -march=native / -ffast-math, so no multiply-add is ever
 * fused in its in-tree loops or in SoundTouch; the oracle keeps every rounding step.
 * The one place a fused multiply-add is used on purpose is the polyphase FIR (see swr_fir()).
 */
#include "nodey_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

const char* orc_version(void) { return "nodey-oracle r1 (libswresample pinned to the real library; SoundTouch parity unpinned)"; }

/* ======================================================================================= */
/* synthetic source                                                                         */
/* ======================================================================================= */

/* splitmix64-style counter hash: one 32-bit draw per (seed, n). */
uint32_t orc_synth_hash(uint32_t seed, uint64_t n)
{
    uint64_t z = n + ((uint64_t)seed << 32) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

/* sin(2*pi*t) for t in [-0.5,0.5) as a fixed float polynomial so that CPU and GPU generators
 * agree bit for bit (every step is an IEEE float op or a fused multiply-add). */
static inline float synth_sin_turns(float t)
{
    if (t > 0.25f) t = 0.5f - t;
    else if (t < -0.25f) t = -0.5f - t;
    const float y = 6.28318530717958647692f * t;
    const float y2 = y * y;
    float p = 2.7557319e-6f;                 /*  1/9! */
    p = fmaf(p, y2, -1.9841270e-4f);         /* -1/7! */
    p = fmaf(p, y2, 8.3333333e-3f);          /*  1/5! */
    p = fmaf(p, y2, -1.6666667e-1f);         /* -1/3! */
    p = fmaf(p, y2, 1.0f);
    return y * p;
}

static uint32_t synth_phase_step(int sample_rate, int track, int ch)
{
    const int e = (7 * track + 4 * ch) % 36;
    const double f = 220.0 * pow(2.0, (double)e / 12.0);
    return (uint32_t)llrint(f / (double)sample_rate * 4294967296.0);
}

/* x = 0.5*sin(2*pi*f*n/sr) + 0.05*u  (SURVEY.md 8d).  Interleaved when nch == 2. */
void orc_synth_f32(float* dst, int64_t nframes, int nch, int sample_rate, int track, int64_t frame0)
{
    for (int c = 0; c < nch; c++) {
        const uint32_t step = synth_phase_step(sample_rate, track, c);
        const uint32_t seed = 0xA0D10u + 131u * (uint32_t)track + (uint32_t)c;
        for (int64_t i = 0; i < nframes; i++) {
            const uint64_t n = (uint64_t)(frame0 + i);
            const uint32_t ph = (uint32_t)(n * (uint64_t)step);
            const float t = (float)(int32_t)ph * 2.3283064365386963e-10f; /* 2^-32 */
            const float s = synth_sin_turns(t);
            const float u = (float)(orc_synth_hash(seed, n) >> 8) * 1.1920928955078125e-7f - 1.0f;
            dst[i * nch + c] = fmaf(0.05f, u, 0.5f * s);
        }
    }
}

void orc_f32_to_s16(int16_t* dst, const float* src, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        long v = lrintf(src[i] * 32767.0f);
        if (v > 32767) v = 32767;
        if (v < -32768) v = -32768;
        dst[i] = (int16_t)v;
    }
}

/* ======================================================================================= */
/* A3 gain -- audio-vol.cpp:75-100: std::copy then `dst[i] *= volume` on the typed sample.   */
/* For integers that is T(float(x) * volume): x86 truncating conversion (cvttss2si), no      */
/* clamp; int16 keeps the low 16 bits, int32 yields INT_MIN when out of range (App. C2).    */
/* ======================================================================================= */
static inline int32_t x86_cvttss2si(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int32_t)f;
}

void orc_gain(void* dst, const void* src, int fmt, int64_t n, float volume)
{
    switch (fmt) {
    case ORC_FMT_FLT: case ORC_FMT_FLTP: {
        const float* s = (const float*)src; float* d = (float*)dst;
        for (int64_t i = 0; i < n; i++) d[i] = s[i] * volume;
        break;
    }
    case ORC_FMT_S16: case ORC_FMT_S16P: {
        const int16_t* s = (const int16_t*)src; int16_t* d = (int16_t*)dst;
        for (int64_t i = 0; i < n; i++) d[i] = (int16_t)(uint16_t)(uint32_t)x86_cvttss2si((float)s[i] * volume);
        break;
    }
    case ORC_FMT_S32: case ORC_FMT_S32P: {
        const int32_t* s = (const int32_t*)src; int32_t* d = (int32_t*)dst;
        for (int64_t i = 0; i < n; i++) d[i] = x86_cvttss2si((float)s[i] * volume);
        break;
    }
    default: break;
    }
}

/* ======================================================================================= */
/* A8 extract_samples_interleaved -- audio-velocity.cpp:150-232.  Note the four different    */
/* integer scales (App. C6).                                                                 */
/* ======================================================================================= */
int orc_extract_interleaved(float* dst, const void* p0, const void* p1, int fmt, int64_t nframes, int nch)
{
    const void* planes[2] = {p0, p1};
    switch (fmt) {
    case ORC_FMT_FLT:
        memcpy(dst, p0, sizeof(float) * (size_t)(nframes * nch));
        return 0;
    case ORC_FMT_FLTP:
        for (int c = 0; c < nch; c++)
            for (int64_t i = 0; i < nframes; i++) dst[i * nch + c] = ((const float*)planes[c])[i];
        return 0;
    case ORC_FMT_S16:
        for (int64_t i = 0; i < nframes * nch; i++) dst[i] = (float)((const int16_t*)p0)[i] / 32768.0f;
        return 0;
    case ORC_FMT_S16P:
        for (int c = 0; c < nch; c++)
            for (int64_t i = 0; i < nframes; i++)
                dst[i * nch + c] = (float)((const int16_t*)planes[c])[i] / (float)32767;
        return 0;
    case ORC_FMT_S32:
        for (int64_t i = 0; i < nframes * nch; i++) dst[i] = (float)((const int32_t*)p0)[i] / 2147483648.0f;
        return 0;
    case ORC_FMT_S32P:
        for (int c = 0; c < nch; c++)
            for (int64_t i = 0; i < nframes; i++)
                dst[i * nch + c] = (float)((double)((const int32_t*)planes[c])[i] / (double)2147483647);
        return 0;
    default:
        return -1; /* reference throws Runtime_error("Unsupported sample format") */
    }
}

/* N1 channel split: pure routing, bit exact. Output format = planar/mono of the same sample type. */
void orc_split(void* dst_l, void* dst_r, const void* p0, const void* p1, int fmt, int64_t n)
{
    switch (fmt) {
    case ORC_FMT_FLT: case ORC_FMT_S32:
        for (int64_t i = 0; i < n; i++) {
            ((uint32_t*)dst_l)[i] = ((const uint32_t*)p0)[2 * i];
            ((uint32_t*)dst_r)[i] = ((const uint32_t*)p0)[2 * i + 1];
        }
        break;
    case ORC_FMT_S16:
        for (int64_t i = 0; i < n; i++) {
            ((uint16_t*)dst_l)[i] = ((const uint16_t*)p0)[2 * i];
            ((uint16_t*)dst_r)[i] = ((const uint16_t*)p0)[2 * i + 1];
        }
        break;
    case ORC_FMT_FLTP: case ORC_FMT_S32P:
        memcpy(dst_l, p0, 4 * (size_t)n); memcpy(dst_r, p1, 4 * (size_t)n); break;
    case ORC_FMT_S16P:
        memcpy(dst_l, p0, 2 * (size_t)n); memcpy(dst_r, p1, 2 * (size_t)n); break;
    default: break;
    }
}

/* ======================================================================================= */
/* libswresample model (FFmpeg 7.1 defaults; App. B1).  Call sites in the reference:         */
/*   audio-amix.cpp:212-290, audio-bimix.cpp:198-294, sw-resample.cpp:8-23 / .hpp:63-69.     */
/* Everything the reference leaves at its default is fixed here:  filter_size 32,            */
/* phase_shift 10, linear_interp 1, exact_rational 1, Kaiser beta 9, cutoff 0.97, no dither. */
/* Output is always stereo FLTP (every call site asks for that).                             */
/* ======================================================================================= */
struct orc_swr {
    int in_rate, out_rate, in_fmt, in_ch;
    int resample;               /* 0: rates equal => format conversion only */
    int phase_count, filter_length, filter_alloc;
    int src_incr, dst_incr, dst_incr_div, dst_incr_mod;
    int index_mask_quirk;
    float* bank;                /* (phase_count + 1) * filter_alloc */
    /* streaming state */
    int inited;                 /* initial mirror done (needs filter_length+1 samples) */
    int64_t index0;             /* phase index after the initial mirror */
    int flushed;
    int64_t n_in;               /* input frames received so far (real samples) */
    int64_t reflect;            /* reflected samples appended by flush */
    int64_t produced;           /* output frames returned so far */
    float* hist[2];             /* converted input history, two float planes */
    int64_t hist_cap;
    int64_t passthrough_buffered; /* non-resample path: frames held in the in_buffer */
};

static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a < 0 ? -a : a; }

/* modified Bessel I0, power series (libswresample's bessel() evaluates the same function) */
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

/* build_filter() (libswresample/resample.c): windowed sinc, Kaiser window, stored as float.
 * Two details of the real routine, both pinned against the real library's impulse responses
 * (tests/test_swr_real.py):
 *   - `norm` is accumulated for phase 0 ONLY (`if (!ph) norm += y;`) and every phase is divided by it,
 *     so only phase 0 has exactly unit DC gain;
 *   - when factor == 1 (upsampling) sin(x) is taken from a per-phase table with alternating sign:
 *     sin(pi*(i - center) - pi*ph/P) = +-sin(pi*ph/P). */
static void swr_build_filter(float* bank, double factor, int tap_count, int alloc, int phase_count, double beta)
{
    const int center = (tap_count - 1) / 2;
    double* tab = (double*)malloc(sizeof(double) * (size_t)tap_count);
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
            tab[i] = y;
            if (!ph) norm += y;
        }
        for (int i = 0; i < tap_count; i++) bank[ph * alloc + i] = (float)(tab[i] * 1.0 / norm);
        if (phase_count % 2) continue;
        for (int i = 0; i < tap_count; i++)
            bank[(phase_count - ph) * alloc + tap_count - 1 - i] = bank[ph * alloc + i];
    }
    free(tab);
}

orc_swr* orc_swr_create(int in_rate, int out_rate, int in_fmt, int in_ch, int index_mask_quirk)
{
    orc_swr* s = (orc_swr*)calloc(1, sizeof(*s));
    s->in_rate = in_rate; s->out_rate = out_rate; s->in_fmt = in_fmt; s->in_ch = in_ch;
    s->index_mask_quirk = index_mask_quirk;
    s->resample = in_rate != out_rate;
    if (!s->resample) return s;

    /* resample_init() */
    const double cutoff = 0.97;
    double factor = (double)out_rate * cutoff / in_rate;
    if (factor > 1.0) factor = 1.0;
    int phase_count = 1 << 10;
    int filter_length = (int)ceil(32 / factor);
    if (filter_length < 1) filter_length = 1;
    if (filter_length > 1) filter_length = (filter_length + 1) & ~1;
    {   /* exact_rational */
        const int64_t g = gcd64(out_rate, in_rate);
        const int64_t exact = out_rate / g;
        if (exact <= phase_count) phase_count = (int)exact;
    }
    s->phase_count = phase_count;
    s->filter_length = filter_length;
    s->filter_alloc = (filter_length + 7) & ~7;
    s->bank = (float*)calloc((size_t)s->filter_alloc * (size_t)(phase_count + 1), sizeof(float));
    swr_build_filter(s->bank, factor, filter_length, s->filter_alloc, phase_count, 9.0);
    /* extra phase = phase 0 advanced by one input sample */
    memcpy(s->bank + (size_t)s->filter_alloc * phase_count + 1, s->bank, sizeof(float) * (size_t)(s->filter_alloc - 1));
    s->bank[(size_t)s->filter_alloc * phase_count] = s->bank[s->filter_alloc - 1];

    /* av_reduce(&src_incr, &dst_incr, out_rate, in_rate * phase_count, INT32_MAX/2) then scale up */
    {
        int64_t num = out_rate, den = (int64_t)in_rate * phase_count;
        const int64_t g = gcd64(num, den);
        num /= g; den /= g;
        while (den < (1 << 20) && num < (1 << 20)) { den *= 2; num *= 2; }
        s->src_incr = (int)num; s->dst_incr = (int)den;
    }
    s->dst_incr_div = s->dst_incr / s->src_incr;
    s->dst_incr_mod = s->dst_incr % s->src_incr;
    {
        /* index = -phase_count * ((filter_length-1)/2); after invert_initial_buffer():
         * literal `index &= phase_count - 1`, or the intended index mod phase_count (= 0). */
        const int idx = -phase_count * ((filter_length - 1) / 2);
        s->index0 = index_mask_quirk ? (idx & (phase_count - 1)) : 0;
    }
    return s;
}

void orc_swr_free(orc_swr* s)
{
    if (!s) return;
    free(s->bank); free(s->hist[0]); free(s->hist[1]); free(s);
}

void orc_swr_plan(const orc_swr* s, int out[8])
{
    out[0] = s->phase_count; out[1] = s->filter_length; out[2] = s->filter_alloc;
    out[3] = s->dst_incr_div; out[4] = s->dst_incr_mod; out[5] = s->src_incr;
    out[6] = (int)s->index0; out[7] = s->resample && s->dst_incr_mod != 0;
}

const float* orc_swr_filter_bank(const orc_swr* s) { return s->bank; }

/* audioconvert + rematrix front end: any supported input -> two float planes.
 *   s16 -> x * (1/32768), s32 -> x * (1/2^31), float copied;
 *   mono -> both outputs = in * sqrt(1/2) (float rematrix coefficient, no normalisation). */
static void swr_append_input(orc_swr* s, const void* in0, const void* in1, int n)
{
    if (s->n_in + n > s->hist_cap) {
        int64_t cap = s->hist_cap ? s->hist_cap * 2 : 65536;
        while (cap < s->n_in + n) cap *= 2;
        s->hist[0] = (float*)realloc(s->hist[0], sizeof(float) * (size_t)cap);
        s->hist[1] = (float*)realloc(s->hist[1], sizeof(float) * (size_t)cap);
        s->hist_cap = cap;
    }
    const int planar = s->in_fmt >= ORC_FMT_U8P;
    const void* pl[2] = {in0, in1};
    for (int c = 0; c < s->in_ch; c++) {
        float* d = s->hist[c] + s->n_in;
        const int stride = planar ? 1 : s->in_ch;
        const int off = planar ? 0 : c;
        const void* src = planar ? pl[c] : in0;
        switch (s->in_fmt) {
        case ORC_FMT_FLT: case ORC_FMT_FLTP:
            for (int i = 0; i < n; i++) d[i] = ((const float*)src)[i * stride + off];
            break;
        case ORC_FMT_S16: case ORC_FMT_S16P:
            for (int i = 0; i < n; i++) d[i] = ((const int16_t*)src)[i * stride + off] * (1.0f / (1 << 15));
            break;
        case ORC_FMT_S32: case ORC_FMT_S32P:
            for (int i = 0; i < n; i++) d[i] = ((const int32_t*)src)[i * stride + off] * (1.0f / (1U << 31));
            break;
        default: break;
        }
    }
    if (s->in_ch == 1) {
        const float k = (float)0.70710678118654752440; /* M_SQRT1_2 as the float matrix entry */
        float* d0 = s->hist[0] + s->n_in; float* d1 = s->hist[1] + s->n_in;
        for (int i = 0; i < n; i++) { d0[i] = d0[i] * k; d1[i] = d0[i]; }
    }
    s->n_in += n;
}

/* extended signal: left mirror x[-n] = x[n]; right reflection (after flush) x[N+j] = x[N-1-j] */
static inline float swr_ext(const orc_swr* s, int c, int64_t i)
{
    if (i < 0) i = -i;
    if (i >= s->n_in) i = 2 * s->n_in - 1 - i;
    return s->hist[c][i];
}

/* Phase position of output k in units of 1/phase_count input samples, plus the fractional
 * numerator for the interpolating path. */
static inline void swr_pos(const orc_swr* s, int64_t k, int64_t* index, int* frac)
{
    const int64_t f = k * (int64_t)s->dst_incr_mod;
    *index = s->index0 + k * (int64_t)s->dst_incr_div + f / s->src_incr;
    *frac = (int)(f % s->src_incr);
}

/* FIR evaluation order.  libswresample has several (C template with two partial sums, SSE/AVX/
 * FMA3 assembly with 4/8 lanes), so summation order is not part of the algorithm; parity against
 * the real library is a tolerance (1e-5) question.  The oracle fixes ONE order that the CUDA kernel
 * reproduces exactly: a single accumulator, ascending taps, fused multiply-add.  That keeps the
 * whole chain bit-reproducible, which the WSOLA arg-max downstream needs. */
static inline float swr_fir(const orc_swr* s, int c, int64_t start, const float* taps)
{
    float acc = 0.0f;
    for (int i = 0; i < s->filter_length; i++) acc = fmaf(swr_ext(s, c, start + i), taps[i], acc);
    return acc;
}

/* how many outputs exist for the samples held so far */
static int64_t swr_producible(const orc_swr* s)
{
    const int64_t L = s->filter_length, P = s->phase_count;
    const int64_t center = (L - 1) / 2;
    /* invert_initial_buffer() waits until filter_length + 1 samples are there; the samples resample_flush() reflects
     * count (a stream of 22..32 frames at 32 taps gets its 33 samples that way and DOES produce output at the flush:
     * pinned against the real library, tests/test_swr_real.py::test_streams_shorter_than_the_filter) */
    if (s->n_in + s->reflect < L + 1) return 0;
    /* window of output k starts at sample s_k - center; it must end inside n_in + reflect */
    const int64_t max_s = s->n_in + s->reflect - L + center;   /* s_k <= max_s */
    if (max_s < 0) return 0;
    /* count k >= 0 with floor(index_k / P) <= max_s, index_k monotone: binary search */
    int64_t lo = 0, hi = (max_s + 2) * P / (s->dst_incr_div > 0 ? s->dst_incr_div : 1) + 4;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        int64_t idx; int fr; swr_pos(s, mid, &idx, &fr);
        if (idx / P <= max_s) lo = mid + 1; else hi = mid;
    }
    return lo;
}

int orc_swr_convert(orc_swr* s, float* out_l, float* out_r, int out_count, const void* in0, const void* in1, int in_count)
{
    float* out[2] = {out_l, out_r};
    if (!s->resample) {
        /* swr_convert() non-resampling branch: pass min(out_count, buffered + in), hold the rest */
        if (in0) swr_append_input(s, in0, in1, in_count);
        int64_t avail = s->n_in - s->produced;
        int n = (int)(avail < out_count ? avail : out_count);
        for (int c = 0; c < 2; c++) memcpy(out[c], s->hist[c] + s->produced, sizeof(float) * (size_t)n);
        s->produced += n;
        return n;
    }
    if (!in0) {
        if (!s->flushed) {
            /* resample_flush(): reflection = (min(in_buffer_count, filter_length) + 1) / 2 where
             * in_buffer_count = samples from the next window start to the end of input */
            s->flushed = 1;
            if (s->n_in < s->filter_length + 1) {
                /* nothing has been consumed yet: in_buffer_count is the whole input */
                s->reflect = (s->n_in + 1) / 2;
            } else {
                int64_t idx; int fr; swr_pos(s, s->produced, &idx, &fr);
                const int64_t wstart = idx / s->phase_count - (s->filter_length - 1) / 2;
                int64_t held = s->n_in - wstart;
                if (held > s->filter_length) held = s->filter_length;
                if (held < 0) held = 0;
                s->reflect = (held + 1) / 2;
            }
        }
    } else {
        swr_append_input(s, in0, in1, in_count);
    }
    const int64_t total = swr_producible(s);
    int64_t n = total - s->produced;
    if (n > out_count) n = out_count;
    if (n < 0) n = 0;
    const int64_t P = s->phase_count, center = (s->filter_length - 1) / 2;
    for (int64_t j = 0; j < n; j++) {
        int64_t idx; int fr; swr_pos(s, s->produced + j, &idx, &fr);
        const int64_t start = idx / P - center;
        const float* taps = s->bank + (size_t)s->filter_alloc * (size_t)(idx % P);
        for (int c = 0; c < 2; c++) {
            if (s->in_ch == 1 && c == 1) { out[1][j] = out[0][j]; continue; }
            float v = swr_fir(s, c, start, taps);
            if (s->dst_incr_mod) {
                /* resample_linear: val += (v2 - val) * (float)frac / src_incr */
                const float v2 = swr_fir(s, c, start, taps + s->filter_alloc);
                v += (v2 - v) * (float)fr / (float)s->src_incr;
            }
            out[c][j] = v;
        }
    }
    s->produced += n;
    return (int)n;
}

int64_t orc_swr_whole(int in_rate, int out_rate, int in_fmt, int in_ch, int quirk,
                      const void* in0, const void* in1, int64_t in_frames, int do_flush,
                      float* out_l, float* out_r, int64_t out_cap)
{
    orc_swr* s = orc_swr_create(in_rate, out_rate, in_fmt, in_ch, quirk);
    int64_t done = 0;
    const int chunk = 1 << 20;
    const int bps = (in_fmt == ORC_FMT_S16 || in_fmt == ORC_FMT_S16P) ? 2 : 4;
    const int planar = in_fmt >= ORC_FMT_U8P;
    for (int64_t pos = 0; pos < in_frames; pos += chunk) {
        const int n = (int)((in_frames - pos) < chunk ? (in_frames - pos) : chunk);
        const char* p0 = (const char*)in0 + (size_t)pos * (size_t)bps * (size_t)(planar ? 1 : in_ch);
        const char* p1 = in1 ? (const char*)in1 + (size_t)pos * (size_t)bps : NULL;
        const int64_t room = out_cap - done;
        done += orc_swr_convert(s, out_l + done, out_r + done, (int)(room > INT_MAX ? INT_MAX : room), p0, p1, n);
    }
    if (do_flush) {
        for (;;) {
            const int64_t room = out_cap - done;
            if (room <= 0) break;
            const int n = orc_swr_convert(s, out_l + done, out_r + done, (int)(room > INT_MAX ? INT_MAX : room), NULL, NULL, 0);
            done += n;
            if (n == 0) break;
        }
    }
    orc_swr_free(s);
    return done;
}

int64_t orc_swr_out_count(int in_rate, int out_rate, int quirk, int64_t in_frames, int do_flush)
{
    orc_swr* s = orc_swr_create(in_rate, out_rate, ORC_FMT_FLT, 1, quirk);
    int64_t r;
    if (!s->resample) r = in_frames;
    else {
        s->n_in = in_frames;
        r = swr_producible(s);
        if (do_flush && in_frames < s->filter_length + 1) {
            s->reflect = (in_frames + 1) / 2;
            r = swr_producible(s);
        } else if (do_flush) {
            int64_t idx; int fr; swr_pos(s, r, &idx, &fr);
            int64_t held = in_frames - (idx / s->phase_count - (s->filter_length - 1) / 2);
            if (held > s->filter_length) held = s->filter_length;
            if (held < 0) held = 0;
            s->reflect = (held + 1) / 2;
            r = swr_producible(s);
        }
    }
    orc_swr_free(s);
    return r;
}

/* ======================================================================================= */
/* frame-stream helpers: the reference moves AVFrames of `frame_size` samples               */
/* ======================================================================================= */
static int track_bps(const orc_track* t) { return (t->fmt == ORC_FMT_S16 || t->fmt == ORC_FMT_S16P) ? 2 : 4; }

/* frame boundaries of a track: uniform frame_size chunks, or the run-length encoded sizes when the
 * producer was a node whose frames are not uniform (an amix output: nb per iteration) */
typedef struct { int64_t n; int64_t* start; } frame_index;

static frame_index fi_build(const orc_track* t)
{
    frame_index fi;
    int64_t n = 0;
    if (t->nruns > 0) { for (int r = 0; r < t->nruns; r++) if (t->run_len[r] > 0 && t->run_count[r] > 0) n += t->run_count[r]; }
    else n = (t->nframes + t->frame_size - 1) / t->frame_size;
    fi.n = n;
    fi.start = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t pos = 0, k = 0;
    if (t->nruns > 0) {
        for (int r = 0; r < t->nruns; r++)
            for (int64_t c = 0; c < t->run_count[r] && t->run_len[r] > 0; c++) { fi.start[k++] = pos; pos += t->run_len[r]; }
        fi.start[n] = pos;
    } else {
        for (; k < n; k++) { fi.start[k] = pos; pos += t->frame_size; }
        fi.start[n] = t->nframes;
    }
    return fi;
}
static int64_t fi_frames(const frame_index* fi) { return fi->n; }
static int fi_len(const frame_index* fi, int64_t f) { return (int)(fi->start[f + 1] - fi->start[f]); }
static void fi_ptrs(const orc_track* t, const frame_index* fi, int64_t f, const void** p0, const void** p1)
{
    const int planar = t->fmt >= ORC_FMT_U8P;
    const size_t off = (size_t)fi->start[f] * (size_t)track_bps(t);
    *p0 = (const char*)t->plane0 + off * (size_t)(planar ? 1 : t->ch);
    *p1 = (planar && t->plane1) ? (const char*)t->plane1 + off : NULL;
}

/* ======================================================================================= */
/* A4 audio_amix -- audio-amix.cpp:149-322, restated iteration by iteration:                 */
/*  - iteration m takes frame m of every input that still has one (buffers[i].front());      */
/*  - nb = min nb_samples over present frames, 1152 if none (:190-195);                      */
/*  - each input: swr_convert(out nb, whole frame) or the flush call (:261-291); the temp    */
/*    buffers come from av_samples_alloc (zero filled), so a short convert leaves silence;   */
/*  - mix: temp += data * volume in input order, from 0.0f (:296-307);                       */
/*  - stop when every input was in flush mode and returned < nb in the same pass (:290,320). */
/* ======================================================================================= */
int64_t orc_amix(const orc_track* in, int nin, const float* volumes, int quirk,
                 float* out_l, float* out_r, int64_t out_cap)
{
    orc_swr** sw = (orc_swr**)calloc((size_t)nin, sizeof(*sw));
    float** dl = (float**)calloc((size_t)nin, sizeof(float*));
    float** dr = (float**)calloc((size_t)nin, sizeof(float*));
    frame_index* fi = (frame_index*)calloc((size_t)nin, sizeof(*fi));
    int maxfs = 1152;
    for (int i = 0; i < nin; i++) {
        sw[i] = orc_swr_create(in[i].rate, 48000, in[i].fmt, in[i].ch, quirk);
        fi[i] = fi_build(&in[i]);
        for (int64_t f = 0; f < fi[i].n; f++) if (fi_len(&fi[i], f) > maxfs) maxfs = fi_len(&fi[i], f);
    }
    for (int i = 0; i < nin; i++) {
        dl[i] = (float*)malloc(sizeof(float) * (size_t)maxfs);
        dr[i] = (float*)malloc(sizeof(float) * (size_t)maxfs);
    }
    int64_t written = 0;
    for (int64_t m = 0;; m++) {
        int nb = INT_MAX;
        for (int i = 0; i < nin; i++)
            if (m < fi_frames(&fi[i])) { const int n = fi_len(&fi[i], m); if (n < nb) nb = n; }
        if (nb == INT_MAX) nb = 1152;
        if (written + nb > out_cap) break;
        int count = 0;
        for (int i = 0; i < nin; i++) {
            memset(dl[i], 0, sizeof(float) * (size_t)nb);
            memset(dr[i], 0, sizeof(float) * (size_t)nb);
            if (m < fi_frames(&fi[i])) {
                const void *p0, *p1; fi_ptrs(&in[i], &fi[i], m, &p0, &p1);
                orc_swr_convert(sw[i], dl[i], dr[i], nb, p0, p1, fi_len(&fi[i], m));
            } else {
                const int c = orc_swr_convert(sw[i], dl[i], dr[i], nb, NULL, NULL, 0);
                if (c < nb) count++;
            }
        }
        for (int j = 0; j < nb; j++) {
            float tl = 0.0f, tr = 0.0f;
            for (int i = 0; i < nin; i++) {
                tl += dl[i][j] * volumes[i];
                tr += dr[i][j] * volumes[i];
            }
            out_l[written + j] = tl;
            out_r[written + j] = tr;
        }
        written += nb;
        if (count == nin) break;
    }
    for (int i = 0; i < nin; i++) { orc_swr_free(sw[i]); free(dl[i]); free(dr[i]); free(fi[i].start); }
    free(sw); free(dl); free(dr); free(fi);
    return written;
}

/* ======================================================================================= */
/* A5 audio_bimix -- audio-bimix.cpp:137-329.  outL=(ll/2+lr/2)*(1-bias), outR=(rl/2+rr/2)*  */
/* (1+bias) (:310-317).  Termination restated literally including the App. C5 slip: in the   */
/* right-side flush branch the count lands in convert_count_l (:294).                        */
/* ======================================================================================= */
int64_t orc_bimix(const orc_track* l, const orc_track* r, float bias, int quirk,
                  float* out_l, float* out_r, int64_t out_cap)
{
    orc_swr* sl = orc_swr_create(l->rate, 48000, l->fmt, l->ch, quirk);
    orc_swr* sr = orc_swr_create(r->rate, 48000, r->fmt, r->ch, quirk);
    frame_index fl_ = fi_build(l), fr_ = fi_build(r);
    int maxfs = 1152;
    for (int64_t f = 0; f < fl_.n; f++) if (fi_len(&fl_, f) > maxfs) maxfs = fi_len(&fl_, f);
    for (int64_t f = 0; f < fr_.n; f++) if (fi_len(&fr_, f) > maxfs) maxfs = fi_len(&fr_, f);
    float* d1[2]; float* d2[2];
    for (int c = 0; c < 2; c++) { d1[c] = (float*)malloc(sizeof(float) * (size_t)maxfs); d2[c] = (float*)malloc(sizeof(float) * (size_t)maxfs); }
    const float bias_minus = (1 - bias), bias_plus = (1 + bias);
    int64_t written = 0;
    for (int64_t m = 0;; m++) {
        const int has_l = m < fi_frames(&fl_), has_r = m < fi_frames(&fr_);
        int nb = 0;
        /* :178-183 -- note the if / if-else-if-else shape: both present => min, then the chain */
        if (has_r && has_l) { const int a = fi_len(&fr_, m), b = fi_len(&fl_, m); nb = a < b ? a : b; }
        if (!has_r && has_l) nb = fi_len(&fl_, m);
        else if (has_r && !has_l) nb = fi_len(&fr_, m);
        else nb = 1152;
        if (nb > maxfs) nb = maxfs;
        if (written + nb > out_cap) break;
        for (int c = 0; c < 2; c++) { memset(d1[c], 0, sizeof(float) * (size_t)nb); memset(d2[c], 0, sizeof(float) * (size_t)nb); }
        int cl = 0, cr = 0;
        const void *p0, *p1;
        if (has_l) { fi_ptrs(l, &fl_, m, &p0, &p1); cl = orc_swr_convert(sl, d1[0], d1[1], nb, p0, p1, fi_len(&fl_, m)); }
        else cl = orc_swr_convert(sl, d1[0], d1[1], nb, NULL, NULL, 0);
        if (has_r) { fi_ptrs(r, &fr_, m, &p0, &p1); cr = orc_swr_convert(sr, d2[0], d2[1], nb, p0, p1, fi_len(&fr_, m)); }
        else cl = orc_swr_convert(sr, d2[0], d2[1], nb, NULL, NULL, 0);
        for (int i = 0; i < nb; i++) {
            out_l[written + i] = (d1[0][i] / 2 + d1[1][i] / 2) * bias_minus;
            out_r[written + i] = (d2[0][i] / 2 + d2[1][i] / 2) * bias_plus;
        }
        written += nb;
        if (cr == 0 && cl == 0) break;
    }
    for (int c = 0; c < 2; c++) { free(d1[c]); free(d2[c]); }
    orc_swr_free(sl); orc_swr_free(sr); free(fl_.start); free(fr_.start);
    return written;
}

/* ======================================================================================= */
/* A6 audio_bimix_v2 -- audio-bimix.cpp:475-877.  Each side: swr -> stereo FLTP -> mono      */
/* (l+r)*0.5 (:625-627); frames carry the END time of their block (App. C12); the two lists  */
/* are merged by time with zero fill (:777-872) and tails are zero padded (:732-775).        */
/* The resampler is never flushed.  Both inputs are consumed in lock step (one frame of each */
/* per loop pass), which is what the fibre scheduling yields when both sources keep up.      */
/* ======================================================================================= */
typedef struct v2_frame { float* s; int64_t n, cap; double t; struct v2_frame* next; } v2_frame;
typedef struct { v2_frame *head, *tail; } v2_list;

static void v2_push(v2_list* l, v2_frame* f) { f->next = NULL; if (l->tail) l->tail->next = f; else l->head = f; l->tail = f; }
static void v2_pop(v2_list* l) { v2_frame* f = l->head; l->head = f->next; if (!l->head) l->tail = NULL; free(f->s); free(f); }
static void v2_drop(v2_frame* f, int64_t count)
{
    memmove(f->s, f->s + count, sizeof(float) * (size_t)(f->n - count));
    f->n -= count;
    f->t += (double)count / 48000;
}

typedef struct { float* out; int64_t n, cap; int have_pts; double pts0; } v2_sink;
static void v2_emit(v2_sink* k, const float* inter, int64_t frames, double t)
{
    if (!k->have_pts) { k->have_pts = 1; k->pts0 = t; }
    if (k->n + frames > k->cap) frames = k->cap - k->n;
    if (frames <= 0) return;
    memcpy(k->out + 2 * k->n, inter, sizeof(float) * 2 * (size_t)frames);
    k->n += frames;
}

static void v2_feed(orc_swr* sw, const orc_track* t, const frame_index* fi, int64_t m, double* time, v2_list* list)
{
    const void *p0, *p1; fi_ptrs(t, fi, m, &p0, &p1);
    const int n = fi_len(fi, m);
    float* b0 = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    float* b1 = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    const int got = orc_swr_convert(sw, b0, b1, n * 2, p0, p1, n);
    *time += (double)got / 48000;
    v2_frame* f = (v2_frame*)calloc(1, sizeof(*f));
    f->t = *time; f->n = got; f->s = (float*)malloc(sizeof(float) * (size_t)(got > 0 ? got : 1));
    for (int i = 0; i < got; i++) f->s[i] = (float)((b0[i] + b1[i]) * 0.5);
    v2_push(list, f);
    free(b0); free(b1);
}

int64_t orc_bimix_v2(const orc_track* l, const orc_track* r, int quirk,
                     float* out, int64_t out_cap, double* out_pts0)
{
    orc_swr* sl = orc_swr_create(l->rate, 48000, l->fmt, l->ch, quirk);
    orc_swr* sr = orc_swr_create(r->rate, 48000, r->fmt, r->ch, quirk);
    frame_index fil = fi_build(l), fir = fi_build(r);
    v2_list fl = {0, 0}, fr = {0, 0};
    v2_sink sink = {out, 0, out_cap, 0, 0.0};
    double time_l = l->pts0, time_r = r->pts0;
    int eof_l = 0, eof_r = 0;
    int64_t ml = 0, mr = 0;
    float* tmp = NULL; int64_t tmp_cap = 0;
#define V2_TMP(nfr) do { if ((nfr) * 2 > tmp_cap) { tmp_cap = (nfr) * 2 + 64; tmp = (float*)realloc(tmp, sizeof(float) * (size_t)tmp_cap); } } while (0)
    for (;;) {
        if (!eof_l) { if (ml < fi_frames(&fil)) v2_feed(sl, l, &fil, ml++, &time_l, &fl); else eof_l = 1; }
        if (!eof_r) { if (mr < fi_frames(&fir)) v2_feed(sr, r, &fir, mr++, &time_r, &fr); else eof_r = 1; }

        if (!fl.head && !fr.head && eof_l && eof_r) break;
        if (!fr.head && eof_r) {
            if (!fl.head) continue;
            v2_frame* f = fl.head; V2_TMP(f->n);
            for (int64_t i = 0; i < f->n; i++) { tmp[2 * i] = f->s[i]; tmp[2 * i + 1] = 0; }
            v2_emit(&sink, tmp, f->n, f->t); v2_pop(&fl); continue;
        }
        if (!fl.head && eof_l) {
            if (!fr.head) continue;
            v2_frame* f = fr.head; V2_TMP(f->n);
            for (int64_t i = 0; i < f->n; i++) { tmp[2 * i] = 0; tmp[2 * i + 1] = f->s[i]; }
            v2_emit(&sink, tmp, f->n, f->t); v2_pop(&fr); continue;
        }
        while (fl.head && fr.head) {
            const int left_earlier = fl.head->t < fr.head->t;
            const int eo = left_earlier ? 0 : 1, lo = left_earlier ? 1 : 0;
            v2_list* es = left_earlier ? &fl : &fr; v2_list* ls = left_earlier ? &fr : &fl;
            const double eb = es->head->t, lb = ls->head->t;
            const double ee = eb + (double)es->head->n / 48000, le = lb + (double)ls->head->n / 48000;
            if (ee <= lb) {
                v2_frame* f = es->head; V2_TMP(f->n);
                for (int64_t i = 0; i < f->n; i++) { tmp[2 * i + eo] = f->s[i]; tmp[2 * i + lo] = 0; }
                v2_emit(&sink, tmp, f->n, eb); v2_pop(es); continue;
            }
            const double fe = ee < le ? ee : le;
            const int64_t un = (int64_t)round((lb - eb) * 48000);
            int64_t al = (int64_t)round((fe - lb) * 48000);
            /* size_t arithmetic in the reference: es.size() - un, unsigned */
            { const uint64_t room = (uint64_t)es->head->n - (uint64_t)un; if ((uint64_t)al > room) al = (int64_t)room; }
            if (al > ls->head->n) al = ls->head->n;
            V2_TMP(un + al);
            for (int64_t i = 0; i < un; i++) { tmp[2 * i + eo] = es->head->s[i]; tmp[2 * i + lo] = 0; }
            for (int64_t i = 0; i < al; i++) { tmp[2 * (i + un) + eo] = es->head->s[i + un]; tmp[2 * (i + un) + lo] = ls->head->s[i]; }
            if (ee <= le) { v2_pop(es); v2_drop(ls->head, al); }
            else { v2_pop(ls); v2_drop(es->head, un + al); }
            if (es->head && es->head->n == 0) v2_pop(es);
            if (ls->head && ls->head->n == 0) v2_pop(ls);
            v2_emit(&sink, tmp, un + al, eb);
        }
    }
#undef V2_TMP
    while (fl.head) v2_pop(&fl);
    while (fr.head) v2_pop(&fr);
    free(tmp); orc_swr_free(sl); orc_swr_free(sr); free(fil.start); free(fir.start);
    if (out_pts0) *out_pts0 = sink.pts0;
    return sink.n;
}

/* ======================================================================================= */
/* SoundTouch 2.3.2 model (App. B2), float-sample x86 build: TDStretch with the SSE          */
/* cross-correlation (4 lane sums), AAFilter 64 taps (SSE stereo / double-sum mono),         */
/* cubic transposer.  Every multiply and add rounds separately.                              */
/* ======================================================================================= */
typedef struct { float* d; int64_t n, cap, pos; int ch; } fifo;   /* n frames valid from pos */

static void fifo_init(fifo* f, int ch) { memset(f, 0, sizeof(*f)); f->ch = ch; }
static void fifo_free(fifo* f) { free(f->d); }
static float* fifo_begin(fifo* f) { return f->d + (size_t)f->pos * (size_t)f->ch; }
static float* fifo_end(fifo* f, int64_t slack)
{
    if (f->pos + f->n + slack > f->cap) {
        if (f->pos > 0) { memmove(f->d, fifo_begin(f), sizeof(float) * (size_t)(f->n * f->ch)); f->pos = 0; }
        if (f->n + slack > f->cap) {
            int64_t cap = f->cap ? f->cap : 4096;
            while (cap < f->n + slack) cap *= 2;
            f->d = (float*)realloc(f->d, sizeof(float) * (size_t)(cap * f->ch));
            f->cap = cap;
        }
    }
    return f->d + (size_t)(f->pos + f->n) * (size_t)f->ch;
}
static void fifo_put(fifo* f, const float* s, int64_t n) { float* e = fifo_end(f, n); memcpy(e, s, sizeof(float) * (size_t)(n * f->ch)); f->n += n; }
static void fifo_commit(fifo* f, int64_t n) { f->n += n; }
static void fifo_take(fifo* f, int64_t n) { if (n >= f->n) { f->n = 0; f->pos = 0; } else { f->pos += n; f->n -= n; } }

typedef struct {
    int ch, sample_rate;
    double tempo, nominal_skip, skip_fract;
    int overlap, seek_window, seek_length, sample_req;
    int is_beginning;
    float* mid;
    fifo in, out;
    int32_t* trace; int64_t trace_n, trace_cap;
    int64_t n_seq;
} tdstretch;

static void tds_calc_seq(tdstretch* t)
{
    /* TDStretch::calcSeqParameters(), automatic sequence / seek window */
    double seq = (90.0 - ((40.0 - 90.0) / (2.0 - 0.5)) * 0.5) + ((40.0 - 90.0) / (2.0 - 0.5)) * t->tempo;
    seq = seq < 40.0 ? 40.0 : (seq > 90.0 ? 90.0 : seq);
    const int sequence_ms = (int)(seq + 0.5);
    double seek = (20.0 - ((15.0 - 20.0) / (2.0 - 0.5)) * 0.5) + ((15.0 - 20.0) / (2.0 - 0.5)) * t->tempo;
    seek = seek < 15.0 ? 15.0 : (seek > 20.0 ? 20.0 : seek);
    const int seek_ms = (int)(seek + 0.5);
    t->seek_window = (t->sample_rate * sequence_ms) / 1000;
    if (t->seek_window < 2 * t->overlap) t->seek_window = 2 * t->overlap;
    t->seek_length = (t->sample_rate * seek_ms) / 1000;
}

static void tds_init(tdstretch* t, int sample_rate, int ch, double tempo)
{
    memset(t, 0, sizeof(*t));
    t->ch = ch; t->sample_rate = sample_rate; t->tempo = tempo;
    /* calculateOverlapLength(8 ms): divisible by 8, at least 16 */
    int ovl = (sample_rate * 8) / 1000;
    if (ovl < 16) ovl = 16;
    ovl -= ovl % 8;
    t->overlap = ovl;
    tds_calc_seq(t);
    /* setTempo() */
    t->nominal_skip = tempo * (t->seek_window - t->overlap);
    const int intskip = (int)(t->nominal_skip + 0.5);
    const int a = intskip + t->overlap;
    t->sample_req = (a > t->seek_window ? a : t->seek_window) + t->seek_length;
    t->mid = (float*)calloc((size_t)(ovl * ch), sizeof(float));
    t->is_beginning = 1;
    t->skip_fract = 0;
    fifo_init(&t->in, ch); fifo_init(&t->out, ch);
}

static void tds_free(tdstretch* t) { free(t->mid); fifo_free(&t->in); fifo_free(&t->out); }

/* TDStretchSSE::calcCrossCorr(): four lane sums over channels*overlap/16 blocks of 16 floats,
 * horizontal add ((l0+l1)+l2)+l3, then double division by sqrt(norm) (norm < 1e-9 -> 1). */
static double tds_cross_corr(const tdstretch* t, const float* p1, const float* p2)
{
    float s[4] = {0, 0, 0, 0}, nrm[4] = {0, 0, 0, 0};
    const int blocks = t->ch * t->overlap / 16;
    for (int b = 0; b < blocks; b++) {
        for (int q = 0; q < 4; q++)
            for (int l = 0; l < 4; l++) {
                const float a = p1[4 * q + l];
                s[l] = s[l] + a * p2[4 * q + l];
                nrm[l] = nrm[l] + a * a;
            }
        p1 += 16; p2 += 16;
    }
    const float norm = ((nrm[0] + nrm[1]) + nrm[2]) + nrm[3];
    const float sum = ((s[0] + s[1]) + s[2]) + s[3];
    return (double)sum / sqrt(norm < 1e-9 ? 1.0 : (double)norm);
}

/* TDStretch::seekBestOverlapPositionFull() */
static int tds_seek(const tdstretch* t, const float* ref)
{
    int best = 0;
    double best_corr = tds_cross_corr(t, ref, t->mid);
    best_corr = (best_corr + 0.1) * 0.75;
    for (int i = 1; i < t->seek_length; i++) {
        double corr = tds_cross_corr(t, ref + t->ch * i, t->mid);
        const double tmp = (double)(2 * i - t->seek_length) / (double)t->seek_length;
        corr = ((corr + 0.1) * (1.0 - 0.25 * tmp * tmp));
        if (corr > best_corr) { best_corr = corr; best = i; }
    }
    return best;
}

static void tds_overlap(const tdstretch* t, float* out, const float* in)
{
    if (t->ch == 2) {
        /* overlapStereo(): f1 ramps up by repeated float adds of 1/overlap */
        const float scale = 1.0f / (float)t->overlap;
        float f1 = 0, f2 = 1.0f;
        for (int i = 0; i < 2 * t->overlap; i += 2) {
            out[i] = in[i] * f1 + t->mid[i] * f2;
            out[i + 1] = in[i + 1] * f1 + t->mid[i + 1] * f2;
            f1 += scale; f2 -= scale;
        }
    } else {
        /* overlapMono() */
        float m1 = 0, m2 = (float)t->overlap;
        for (int i = 0; i < t->overlap; i++) {
            out[i] = (in[i] * m1 + t->mid[i] * m2) / t->overlap;
            m1 += 1; m2 -= 1;
        }
    }
}

static void tds_process(tdstretch* t)
{
    while (t->in.n >= t->sample_req) {
        int offset = 0;
        const float* ib = fifo_begin(&t->in);
        if (!t->is_beginning) {
            offset = tds_seek(t, ib);
            if (t->trace) {
                if (t->trace_n < t->trace_cap) t->trace[t->trace_n] = offset;
                t->trace_n++;
            }
            float* oe = fifo_end(&t->out, t->overlap);
            tds_overlap(t, oe, ib + t->ch * offset);
            fifo_commit(&t->out, t->overlap);
            offset += t->overlap;
        } else {
            t->is_beginning = 0;
            const int skip = (int)(t->tempo * t->overlap + 0.5 * t->seek_length + 0.5);
            t->skip_fract -= skip;
            if (t->skip_fract <= -t->nominal_skip) t->skip_fract = -t->nominal_skip;
        }
        if (t->in.n < (offset + t->seek_window - t->overlap)) continue;
        const int temp = t->seek_window - 2 * t->overlap;
        fifo_put(&t->out, ib + t->ch * offset, temp);
        ib = fifo_begin(&t->in);
        memcpy(t->mid, ib + t->ch * (offset + temp), sizeof(float) * (size_t)(t->ch * t->overlap));
        t->skip_fract += t->nominal_skip;
        const int ovl_skip = (int)t->skip_fract;
        t->skip_fract -= ovl_skip;
        fifo_take(&t->in, ovl_skip);
        t->n_seq++;
    }
}

/* ---- AAFilter (64 taps) + FIRFilter ---- */
typedef struct { float h[64]; int len; } aafilter;

static void aa_design(aafilter* f, double cutoff)
{
    /* AAFilter::calculateCoeffs(): sinc * Hamming, scaled to sum 16384, +-0.5 rounding offset kept
     * (float build does not truncate), then divided by 2^14 (FIRFilterSSE::setCoefficients). */
    const int length = 64;
    double work[64], sum = 0;
    const double wc = 2.0 * M_PI * cutoff;
    const double temp_coeff = (2 * M_PI) / (double)length;
    for (int i = 0; i < length; i++) {
        const double cnt = (double)i - (double)(length / 2);
        double temp = cnt * wc;
        const double h = (temp != 0) ? sin(temp) / temp : 1.0;
        const double w = 0.54 + 0.46 * cos(temp_coeff * cnt);
        temp = w * h;
        work[i] = temp;
        sum += temp;
    }
    const double scale = 16384.0f / sum;
    for (int i = 0; i < length; i++) {
        double temp = work[i] * scale;
        temp += (temp >= 0) ? 0.5 : -0.5;
        const float c = (float)temp;
        f->h[i] = c / 16384.0f;
    }
    f->len = length;
}

/* FIRFilterSSE::evaluateFilterStereo(): even taps and odd taps in separate accumulators, added at
 * the end; (numSamples - length) & ~1 outputs.  Mono: FIRFilter::evaluateFilterMono(), double sum. */
static int64_t aa_evaluate(const aafilter* f, fifo* dst, fifo* src)
{
    const int64_t n = src->n;
    if (n < f->len) return 0;
    int64_t count;
    const float* s = fifo_begin(src);
    if (src->ch == 2) {
        count = (n - f->len) & ~(int64_t)1;
        if (count < 2) return 0;
        float* d = fifo_end(dst, count);
        s = fifo_begin(src);
        for (int64_t j = 0; j < count; j++) {
            float e0 = 0, e1 = 0, o0 = 0, o1 = 0;
            const float* p = s + 2 * j;
            for (int i = 0; i < f->len; i += 2) {
                e0 = e0 + p[2 * i] * f->h[i];
                e1 = e1 + p[2 * i + 1] * f->h[i];
                o0 = o0 + p[2 * i + 2] * f->h[i + 1];
                o1 = o1 + p[2 * i + 3] * f->h[i + 1];
            }
            d[2 * j] = o0 + e0;
            d[2 * j + 1] = o1 + e1;
        }
    } else {
        count = n - f->len;
        if (count <= 0) return 0;
        float* d = fifo_end(dst, count);
        s = fifo_begin(src);
        for (int64_t j = 0; j < count; j++) {
            double sum = 0;
            for (int i = 0; i < f->len; i++) sum += s[j + i] * f->h[i];
            d[j] = (float)sum;
        }
    }
    fifo_commit(dst, count);
    fifo_take(src, count);
    return count;
}

/* ---- InterpolateCubic ---- */
static const float cubic_coeffs[16] = {
    -0.5f, 1.0f, -0.5f, 0.0f,
    1.5f, -2.5f, 0.0f, 1.0f,
    -1.5f, 2.0f, 0.5f, 0.0f,
    0.5f, -0.5f, 0.0f, 0.0f};

typedef struct { double rate, fract; } cubic;

static int64_t cubic_transpose(cubic* c, fifo* dst, fifo* src)
{
    const int64_t nsrc = src->n;
    const int64_t demand = (int64_t)((double)nsrc / c->rate) + 8;
    float* d = fifo_end(dst, demand);
    const float* p = fifo_begin(src);
    const int ch = src->ch;
    const int64_t end = nsrc - 4;
    int64_t used = 0, i = 0;
    while (used < end) {
        const float x3 = 1.0f;
        const float x2 = (float)c->fract;
        const float x1 = x2 * x2;
        const float x0 = x1 * x2;
        const float y0 = cubic_coeffs[0] * x0 + cubic_coeffs[1] * x1 + cubic_coeffs[2] * x2 + cubic_coeffs[3] * x3;
        const float y1 = cubic_coeffs[4] * x0 + cubic_coeffs[5] * x1 + cubic_coeffs[6] * x2 + cubic_coeffs[7] * x3;
        const float y2 = cubic_coeffs[8] * x0 + cubic_coeffs[9] * x1 + cubic_coeffs[10] * x2 + cubic_coeffs[11] * x3;
        const float y3 = cubic_coeffs[12] * x0 + cubic_coeffs[13] * x1 + cubic_coeffs[14] * x2 + cubic_coeffs[15] * x3;
        if (ch == 2) {
            d[2 * i] = y0 * p[0] + y1 * p[2] + y2 * p[4] + y3 * p[6];
            d[2 * i + 1] = y0 * p[1] + y1 * p[3] + y2 * p[5] + y3 * p[7];
        } else {
            d[i] = y0 * p[0] + y1 * p[1] + y2 * p[2] + y3 * p[3];
        }
        i++;
        c->fract += c->rate;
        const int whole = (int)c->fract;
        c->fract -= whole;
        p += ch * whole;
        used += whole;
    }
    fifo_commit(dst, i);
    fifo_take(src, used);
    return i;
}

/* ---- RateTransposer ---- */
typedef struct { cubic tr; aafilter aa; fifo in, mid, out; } ratetransposer;

static void rt_init(ratetransposer* r, int ch, double rate)
{
    /* buffers are created stereo and pre-filled with latency (1 + 32) silent frames before
     * setChannels() runs, so a mono stream starts with 66 silent samples (FIFOSampleBuffer::
     * setChannels re-interprets the stored floats). */
    fifo_init(&r->in, ch); fifo_init(&r->mid, ch); fifo_init(&r->out, ch);
    const int prefill = (1 + 32) * 2 / ch;
    float* e = fifo_end(&r->in, prefill);
    memset(e, 0, sizeof(float) * (size_t)(prefill * ch));
    fifo_commit(&r->in, prefill);
    r->tr.rate = rate; r->tr.fract = 0;
    aa_design(&r->aa, rate > 1.0 ? 0.5 / rate : 0.5 * rate);
}
static void rt_free(ratetransposer* r) { fifo_free(&r->in); fifo_free(&r->mid); fifo_free(&r->out); }

static void rt_put(ratetransposer* r, const float* s, int64_t n)
{
    if (n == 0) return;
    fifo_put(&r->in, s, n);
    if (r->tr.rate < 1.0f) {
        cubic_transpose(&r->tr, &r->mid, &r->in);
        aa_evaluate(&r->aa, &r->out, &r->mid);
    } else {
        aa_evaluate(&r->aa, &r->mid, &r->in);
        cubic_transpose(&r->tr, &r->out, &r->mid);
    }
}

/* ---- SoundTouch facade ---- */
typedef struct {
    int ch;
    double rate, tempo;
    double expected_out; long samples_output;
    tdstretch td; ratetransposer rt;
    int td_first;     /* rate > 1: TDStretch then RateTransposer */
} soundtouch;

static fifo* st_out(soundtouch* s) { return s->td_first ? &s->rt.out : &s->td.out; }

static void st_put(soundtouch* s, const float* x, int64_t n)
{
    s->expected_out += (double)n / (s->rate * s->tempo);
    if (!s->td_first) {
        rt_put(&s->rt, x, n);
        /* pTDStretch->moveSamples(*pRateTransposer) */
        fifo_put(&s->td.in, fifo_begin(&s->rt.out), s->rt.out.n);
        fifo_take(&s->rt.out, s->rt.out.n);
        tds_process(&s->td);
    } else {
        fifo_put(&s->td.in, x, n);
        tds_process(&s->td);
        const int64_t m = s->td.out.n;
        rt_put(&s->rt, fifo_begin(&s->td.out), m);
        fifo_take(&s->td.out, m);
    }
}

float orc_pitch_node_factor(float semitones) { return powf(2.0f, semitones / 12.0f); }
float orc_velocity_node_pitch(float velocity, int keep_pitch) { return keep_pitch ? 1 / velocity : 1; }

int64_t orc_soundtouch(const float* in, int64_t nframes, int nch, int sample_rate,
                       float rate_arg, float pitch_arg, int frame_size,
                       float* out, int64_t out_cap,
                       int32_t* offsets, int64_t offsets_cap, orc_st_info* info)
{
    soundtouch s; memset(&s, 0, sizeof(s));
    /* setRate(velocity); setPitch(pitch): virtualRate, virtualPitch doubles; virtualTempo = 1 */
    const double vrate = (double)rate_arg, vpitch = (double)pitch_arg;
    s.ch = nch;
    s.tempo = 1.0 / vpitch;
    s.rate = vpitch * vrate;
    s.td_first = !(s.rate <= 1.0f);
    tds_init(&s.td, sample_rate, nch, s.tempo);
    rt_init(&s.rt, nch, s.rate);
    s.td.trace = offsets; s.td.trace_cap = offsets_cap;

    int64_t got = 0;
#define ST_DRAIN() do { fifo* o = st_out(&s); int64_t n = o->n; if (got + n > out_cap) n = out_cap - got; \
        if (n > 0) { memcpy(out + (size_t)got * nch, fifo_begin(o), sizeof(float) * (size_t)(n * nch)); got += n; } \
        s.samples_output += (long)o->n; fifo_take(o, o->n); } while (0)
    for (int64_t pos = 0; pos < nframes; pos += frame_size) {
        const int64_t n = (nframes - pos) < frame_size ? (nframes - pos) : frame_size;
        st_put(&s, in + (size_t)pos * nch, n);
        ST_DRAIN();   /* receiveSamples() timing does not change content; drain eagerly */
    }
    /* flush(): feed 128-frame silent blocks (<= 200) until enough output, then trim to the
     * expected total: numStillExpected = (long)(samplesExpectedOut + 0.5) - samplesOutput */
    {
        int still = (int)((long)(s.expected_out + 0.5) - s.samples_output);
        if (still < 0) still = 0;
        float* zeros = (float*)calloc((size_t)(128 * nch), sizeof(float));
        for (int i = 0; (still > (int)st_out(&s)->n) && (i < 200); i++) st_put(&s, zeros, 128);
        free(zeros);
        fifo* o = st_out(&s);
        if (o->n > still) o->n = still;   /* adjustAmountOfSamples() */
        ST_DRAIN();
    }
#undef ST_DRAIN
    if (info) {
        info->sample_rate = sample_rate; info->channels = nch;
        info->rate = s.rate; info->tempo = s.tempo;
        info->overlap = s.td.overlap; info->seek_window = s.td.seek_window;
        info->seek_length = s.td.seek_length; info->sample_req = s.td.sample_req;
        info->nominal_skip = s.td.nominal_skip; info->tdstretch_first = s.td_first;
        info->n_sequences = s.td.n_seq;
    }
    if (offsets && offsets_cap > 0 && info) { /* trace_n may exceed cap; caller sees n_sequences */ }
    tds_free(&s.td); rt_free(&s.rt);
    return got;
}

/* ---- the streaming object itself: SoundTouch's public calls, for the reference-loop restatement below and for the
 * stand-in library of the pin harness (tests/fake_soundtouch) ---- */
struct orc_st { soundtouch s; int sample_rate; };

orc_st* orc_st_create(int sample_rate, int nch, float rate_arg, float pitch_arg)
{
    orc_st* h = (orc_st*)calloc(1, sizeof(orc_st));
    if (!h) return NULL;
    const double vrate = (double)rate_arg, vpitch = (double)pitch_arg;
    h->sample_rate = sample_rate;
    h->s.ch = nch;
    h->s.tempo = 1.0 / vpitch;
    h->s.rate = vpitch * vrate;
    h->s.td_first = !(h->s.rate <= 1.0f);
    tds_init(&h->s.td, sample_rate, nch, h->s.tempo);
    rt_init(&h->s.rt, nch, h->s.rate);
    return h;
}

void orc_st_destroy(orc_st* h)
{
    if (!h) return;
    tds_free(&h->s.td); rt_free(&h->s.rt);
    free(h);
}

void orc_st_put(orc_st* h, const float* x, int64_t n) { if (n > 0) st_put(&h->s, x, n); }

int64_t orc_st_num_samples(orc_st* h) { return st_out(&h->s)->n; }

/* SoundTouch::receiveSamples(): samplesOutput counts what was handed out */
int64_t orc_st_receive(orc_st* h, float* out, int64_t max_frames)
{
    fifo* o = st_out(&h->s);
    int64_t n = o->n < max_frames ? o->n : max_frames;
    if (n < 0) n = 0;
    if (n > 0) memcpy(out, fifo_begin(o), sizeof(float) * (size_t)(n * h->s.ch));
    fifo_take(o, n);
    h->s.samples_output += (long)n;
    return n;
}

/* SoundTouch::flush(): numStillExpected from what was RECEIVED so far; silent 128-frame blocks (at most 200) until
 * the output holds that much; trim the output to it; clear TDStretch's input */
void orc_st_flush(orc_st* h)
{
    soundtouch* s = &h->s;
    int still = (int)((long)(s->expected_out + 0.5) - s->samples_output);
    if (still < 0) still = 0;
    float* zeros = (float*)calloc((size_t)(128 * s->ch), sizeof(float));
    for (int i = 0; (still > (int)st_out(s)->n) && (i < 200); i++) st_put(s, zeros, 128);
    free(zeros);
    fifo* o = st_out(s);
    if (o->n > still) o->n = still;
    s->td.in.n = 0;                /* pTDStretch->clearInput() */
}

/* soundtouch_process_payload (audio-velocity.cpp:286-441) with a frame available at every iteration: one putSamples
 * per loop turn, receiveSamples(min(numSamples, 3 * 1152 / velocity)) whenever more than 1152 / velocity are queued,
 * and at end of input either the early `break` (output FIFO empty: SURVEY.md App. C7, the tail is lost) or flush() +
 * one last receive.  chunk_sizes: the frame sizes the node pushes downstream.  *flushed: 1 when flush() was reached. */
int64_t orc_soundtouch_reference_loop(const float* in, int64_t nframes, int nch, int sample_rate, float velocity, float pitch_arg,
                                      int frame_size, float* out, int64_t out_cap, int64_t* chunk_sizes, int64_t chunk_cap,
                                      int64_t* nchunks, int* flushed)
{
    orc_st* h = NULL;
    const double time_ratio = 1.0f / velocity;                 /* const double time_ratio = 1.0f / velocity; (:289) */
    int64_t pos = 0, got = 0, chunks = 0;
    int eof = 0, did_flush = 0;
    for (;;) {
        if (!eof) {
            if (pos >= nframes) eof = 1;                       /* try_pop fails and the stream is at EOF */
            else {
                const int64_t n = (nframes - pos) < frame_size ? (nframes - pos) : frame_size;
                if (!h) h = orc_st_create(sample_rate, nch, velocity, pitch_arg);
                orc_st_put(h, in + (size_t)pos * nch, n);
                pos += n;
            }
        }
        if (h) {
            if (orc_st_num_samples(h) == 0 && eof) break;      /* :414 */
            const uint32_t min_samples = (uint32_t)(time_ratio * 1152);
            const uint32_t max_samples = (uint32_t)(time_ratio * 1152 * 3);
            int64_t want = -1;
            if ((uint64_t)orc_st_num_samples(h) > min_samples) {
                want = orc_st_num_samples(h);
                if (want > (int64_t)max_samples) want = max_samples;
            } else if (eof) {
                orc_st_flush(h);
                did_flush = 1;
                want = orc_st_num_samples(h);
                if (want <= 0) break;
            }
            if (want > 0) {
                int64_t room = out_cap - got;
                float* scratch = NULL;
                float* dst = out + (size_t)got * nch;
                if (want > room) { scratch = (float*)malloc(sizeof(float) * (size_t)(want * nch)); dst = scratch; }
                const int64_t n = orc_st_receive(h, dst, want);
                if (scratch) { if (room > 0) memcpy(out + (size_t)got * nch, scratch, sizeof(float) * (size_t)(room * nch)); free(scratch); }
                got += n < room ? n : (room > 0 ? room : 0);
                if (chunk_sizes && chunks < chunk_cap) chunk_sizes[chunks] = n;
                chunks++;
                if (did_flush) break;
            }
        } else if (eof) break;                                 /* empty input: no SoundTouch object was ever made */
    }
    if (nchunks) *nchunks = chunks;
    if (flushed) *flushed = did_flush;
    orc_st_destroy(h);
    return got;
}

/* ======================================================================================= */
/* N2 spectrum (new node; FFTW's r2c convention, unnormalised, exp(-2*pi*i*k*n/N)).          */
/* Window and windowing product are float32; the DFT itself is evaluated in double and       */
/* rounded once, so it is the exact-arithmetic reference for the float32 GPU FFT.            */
/* ======================================================================================= */
void orc_hann_window(float* w, int nfft)
{
    for (int n = 0; n < nfft; n++) w[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)nfft));
}

int64_t orc_stft_frames(int64_t nframes, int nfft, int hop)
{
    if (nframes < nfft) return 0;
    return (nframes - nfft) / hop + 1;
}

static void fft_double(double* re, double* im, int n)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1;
        for (int k = 0; k < half; k++) {
            const double ang = -2.0 * M_PI * (double)k / (double)len;
            const double wr = cos(ang), wi = sin(ang);
            for (int i = k; i < n; i += len) {
                const int j = i + half;
                const double xr = re[j] * wr - im[j] * wi, xi = re[j] * wi + im[j] * wr;
                re[j] = re[i] - xr; im[j] = im[i] - xi;
                re[i] += xr; im[i] += xi;
            }
        }
    }
}

int64_t orc_stft(const float* x, int64_t nframes, int nfft, int hop, float* out)
{
    const int64_t m = orc_stft_frames(nframes, nfft, hop);
    float* w = (float*)malloc(sizeof(float) * (size_t)nfft);
    double* re = (double*)malloc(sizeof(double) * (size_t)nfft);
    double* im = (double*)malloc(sizeof(double) * (size_t)nfft);
    orc_hann_window(w, nfft);
    const int bins = nfft / 2 + 1;
    for (int64_t f = 0; f < m; f++) {
        const float* p = x + f * hop;
        for (int n = 0; n < nfft; n++) { re[n] = (double)(p[n] * w[n]); im[n] = 0; }
        fft_double(re, im, nfft);
        float* o = out + (size_t)f * (size_t)bins * 2;
        for (int k = 0; k < bins; k++) { o[2 * k] = (float)re[k]; o[2 * k + 1] = (float)im[k]; }
    }
    free(w); free(re); free(im);
    return m;
}
