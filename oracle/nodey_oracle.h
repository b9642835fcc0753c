/*
 * nodey_oracle.h -- CPU restatement of the Nodey Audio Editor processor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline.  The product path is the CUDA library
 * (include/nodey_cuda.h, include/nodey_engine.h) and fails loudly without it.
 *
 * PINNING STATUS.  The reference (Stehsaer/nodey-audio-editor) ships no tests, golden vectors
 * or fixtures, cannot be built in this image (needs FFmpeg 7.1, SoundTouch 2.3.2, Boost.Fiber,
 * jsoncpp, SDL2, ImGui, LAME, <print>), and the heavy arithmetic lives in un-vendored
 * libraries:  libswresample (FFmpeg 7.1, xmake.lua:12) and SoundTouch 2.3.2 (xmake.lua:16).
 *   - libswresample: restated here and PINNED against the real library -- a stock libswresample
 *     6.1.100 (FFmpeg 8.0.1) ships inside this image's opencv wheel; oracle/real_swr.py drives it the
 *     way the reference's call sites do, tests/golden/swr_real.npz holds its outputs and
 *     tests/test_swr_real.py compares (filter bank bit exact, per-call counts exact, values 1e-6).
 *   - SoundTouch: PARITY UNPINNED.  No binary or source of it exists here; its published algorithm is
 *     restated from upstream knowledge and anchored on the reference's call sites and on behavioural
 *     pins (tests/test_oracle.py).
 *   - In-tree arithmetic (gain, mixers, sample extraction) is restated literally from
 *     /root/reference/src/processor/ (cited per function).
 *
 * Sample formats use FFmpeg's AVSampleFormat numbering so the values in reference frames map 1:1.
 */
#ifndef NODEY_ORACLE_H
#define NODEY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_FMT_U8 = 0, ORC_FMT_S16 = 1, ORC_FMT_S32 = 2, ORC_FMT_FLT = 3, ORC_FMT_DBL = 4,
    ORC_FMT_U8P = 5, ORC_FMT_S16P = 6, ORC_FMT_S32P = 7, ORC_FMT_FLTP = 8, ORC_FMT_DBLP = 9
};

/* ---- synthetic source ------------------------------------------------- */
void orc_synth_f32(float* dst, int64_t nframes, int nch, int sample_rate, int track, int64_t frame0);
void orc_f32_to_s16(int16_t* dst, const float* src, int64_t n);
uint32_t orc_synth_hash(uint32_t seed, uint64_t n);

/* ---- A3 gain: src/processor/audio-vol.cpp:75-100 -------------------------------------- */
void orc_gain(void* dst, const void* src, int fmt, int64_t n_elems, float volume);

/* ---- A8 sample extraction: src/processor/audio-velocity.cpp:150-232 ------------------- */
int orc_extract_interleaved(float* dst, const void* plane0, const void* plane1, int fmt,
                            int64_t nframes, int nch);

/* ---- libswresample model (App. B1): format conversion + rematrix + polyphase FIR ------- */
typedef struct orc_swr orc_swr;
/* index_mask_quirk: 1 = restate `c->index &= c->phase_count - 1` literally (initial phase 128 for
 * 160 phases); 0 = mathematically intended phase 0. */
orc_swr* orc_swr_create(int in_rate, int out_rate, int in_fmt, int in_ch, int index_mask_quirk);
void orc_swr_free(orc_swr* s);
/* swr_convert(): in planes NULL => flush call.  Output is stereo float planar. Returns count. */
int orc_swr_convert(orc_swr* s, float* out_l, float* out_r, int out_count,
                    const void* in0, const void* in1, int in_count);
/* plan introspection: phase_count, filter_length, filter_alloc, dst_incr_div, dst_incr_mod,
 * src_incr, index0 (after initial mirror), linear (1 when the interpolating path is active) */
void orc_swr_plan(const orc_swr* s, int out[8]);
const float* orc_swr_filter_bank(const orc_swr* s);  /* (phase_count+1) x filter_alloc */
/* whole-buffer convenience: all input at once, optional flush; returns produced frames */
int64_t orc_swr_whole(int in_rate, int out_rate, int in_fmt, int in_ch, int index_mask_quirk,
                      const void* in0, const void* in1, int64_t in_frames, int do_flush,
                      float* out_l, float* out_r, int64_t out_cap);
int64_t orc_swr_out_count(int in_rate, int out_rate, int index_mask_quirk, int64_t in_frames,
                          int do_flush);

/* ---- tracks as the reference's frame streams see them ---------------------------------- */
typedef struct {
    const void* plane0;   /* packed: the interleaved buffer; planar: channel 0 */
    const void* plane1;   /* planar channel 1, else NULL */
    int fmt, rate, ch;
    int64_t nframes;
    int frame_size;       /* decoder frame size, 1152 unless stated */
    double pts0;          /* seconds, start time of frame 0 */
    /* optional run-length encoded frame sizes (nruns > 0 overrides frame_size): frames of a stream
     * produced by audio_amix are nb samples long per iteration, not uniform */
    const int64_t* run_len;
    const int64_t* run_count;
    int nruns;
} orc_track;

/* A4 audio_amix: src/processor/audio-amix.cpp:86-324.  Returns frames written (FLTP, 48 kHz). */
int64_t orc_amix(const orc_track* in, int nin, const float* volumes, int index_mask_quirk,
                 float* out_l, float* out_r, int64_t out_cap);
/* A5 audio_bimix: src/processor/audio-bimix.cpp:83-331 */
int64_t orc_bimix(const orc_track* l, const orc_track* r, float bias, int index_mask_quirk,
                  float* out_l, float* out_r, int64_t out_cap);
/* A6 audio_bimix_v2: src/processor/audio-bimix.cpp:475-877 (FLT interleaved out).
 * out_pts0 receives the pts (seconds) of the first emitted frame. */
int64_t orc_bimix_v2(const orc_track* l, const orc_track* r, int index_mask_quirk,
                     float* out_interleaved, int64_t out_cap, double* out_pts0);

/* ---- SoundTouch 2.3.2 model (App. B2), driven like audio-velocity.cpp:265-443 ----------- */
typedef struct {
    int sample_rate, channels;
    double rate, tempo;            /* effective */
    int overlap, seek_window, seek_length, sample_req;
    double nominal_skip;
    int tdstretch_first;           /* 1 when rate > 1 */
    int64_t n_sequences;
} orc_st_info;
/* in: interleaved float.  frame_size = putSamples chunk (reference frame size).  Returns output
 * frames (interleaved float written to out).  offsets (optional) receives the WSOLA offset trace. */
int64_t orc_soundtouch(const float* in, int64_t nframes, int nch, int sample_rate,
                       float rate_arg, float pitch_arg, int frame_size,
                       float* out, int64_t out_cap,
                       int32_t* offsets, int64_t offsets_cap, orc_st_info* info);
/* parameter helpers the nodes use (audio-velocity.cpp:445-477) */
float orc_pitch_node_factor(float semitones);       /* std::pow(2.0f, pitch / 12.0f) */
float orc_velocity_node_pitch(float velocity, int keep_pitch);

/* ---- N1 channel split (new node), N2 spectrum (new node) -------------------------------- */
void orc_split(void* dst_l, void* dst_r, const void* plane0, const void* plane1, int fmt, int64_t nframes);
/* The streaming object (SoundTouch's public calls on the same model): what the whole-buffer function above drives
 * eagerly.  Used by the literal restatement of the node's loop below and by the stand-in library of the pin harness. */
typedef struct orc_st orc_st;
orc_st* orc_st_create(int sample_rate, int nch, float rate_arg, float pitch_arg);
void orc_st_destroy(orc_st* h);
void orc_st_put(orc_st* h, const float* x, int64_t nframes);
int64_t orc_st_num_samples(orc_st* h);
int64_t orc_st_receive(orc_st* h, float* out, int64_t max_frames);
void orc_st_flush(orc_st* h);
/* soundtouch_process_payload, src/processor/audio-velocity.cpp:286-441, literally (one frame per loop turn): the
 * reference's receive sizes and its early break before flush() */
int64_t orc_soundtouch_reference_loop(const float* in, int64_t nframes, int nch, int sample_rate, float velocity, float pitch_arg,
                                      int frame_size, float* out, int64_t out_cap, int64_t* chunk_sizes, int64_t chunk_cap,
                                      int64_t* nchunks, int* flushed);

int64_t orc_stft_frames(int64_t nframes, int nfft, int hop);
/* out: [frames][nfft/2+1] complex64 (re,im interleaved); window: periodic Hann in float32 */
int64_t orc_stft(const float* x, int64_t nframes, int nfft, int hop, float* out);
void orc_hann_window(float* w, int nfft);

const char* orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
