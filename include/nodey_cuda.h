/*
 * nodey_cuda.h -- C ABI of libnodey_cuda.so: the sm_100a kernels behind the Nodey Audio Editor
 * processor nodes.  This is the drop-in boundary for the offline-render hot path: plain pointers
 * and sizes, no C++ / torch types.  The reference has no FFI of its own (single C++ process);
 * each entry point names the reference code it replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - every function returns 0 on success, a negative NODEY_E_* code otherwise; the message is
 *    available per thread from nodey_last_error();
 *  - all sample pointers are DEVICE pointers unless the parameter is documented as host;
 *  - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default);
 *  - sample formats use FFmpeg's AVSampleFormat numbering, as stored in the reference's frames
 *    (AVFrame::format): S16=1 S32=2 FLT=3 S16P=6 S32P=7 FLTP=8;
 *  - "frames" = samples per channel.
 */
#ifndef NODEY_CUDA_H
#define NODEY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nodey_stream_t;   /* cudaStream_t */
typedef void* nodey_event_t;    /* cudaEvent_t */

enum {
    NODEY_FMT_U8 = 0, NODEY_FMT_S16 = 1, NODEY_FMT_S32 = 2, NODEY_FMT_FLT = 3, NODEY_FMT_DBL = 4,
    NODEY_FMT_U8P = 5, NODEY_FMT_S16P = 6, NODEY_FMT_S32P = 7, NODEY_FMT_FLTP = 8, NODEY_FMT_DBLP = 9
};

enum {
    NODEY_OK = 0,
    NODEY_E_INVALID = -1,      /* bad argument */
    NODEY_E_FORMAT = -2,       /* unsupported sample format (reference: Runtime_error "format is not support") */
    NODEY_E_CUDA = -3,         /* CUDA runtime error, text in nodey_last_error() */
    NODEY_E_NOMEM = -4,
    NODEY_E_RANGE = -5,        /* parameter outside the range the reference accepts */
    NODEY_E_COMM = -6          /* NCCL missing or an NCCL call failed (nodey_bus_*), text in nodey_last_error() */
};

#define NODEY_MAX_MIX_INPUTS 16   /* audio-amix.cpp:342 clamps input_num to 1..16 */

int nodey_version(void);
const char* nodey_last_error(void);
/* sm count, cc major, cc minor, total bytes: fails (NODEY_E_CUDA) when no CUDA device is present */
int nodey_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);

/* Device runtime behind the ABI (the reference has none: it is a CPU program).  Host layers bind these
 * instead of the CUDA runtime so that the only native dependency is this library.
 * nodey_malloc / nodey_free are stream ordered: a block freed on a stream may be handed out again to
 * that stream at once, to other streams after the free point has been reached (large blocks are
 * cached by the library, small ones use cudaMallocAsync).  Streams are blocking streams, i.e. ordered
 * against the legacy default stream (NULL). */
int nodey_set_device(int ordinal);
int nodey_get_device(int* ordinal);
int nodey_device_count(int* count);
int nodey_device_synchronize(void);
int nodey_stream_create(nodey_stream_t* out);
int nodey_stream_create_priority(nodey_stream_t* out, int high_priority);   /* high: scheduled first when SM slots free up */
int nodey_stream_destroy(nodey_stream_t s);
int nodey_stream_synchronize(nodey_stream_t s);
int nodey_event_create(nodey_event_t* out, int timing);
int nodey_event_destroy(nodey_event_t e);
int nodey_event_record(nodey_event_t e, nodey_stream_t s);
int nodey_event_synchronize(nodey_event_t e);
int nodey_event_elapsed_ms(float* ms, nodey_event_t start, nodey_event_t stop);
int nodey_stream_wait_event(nodey_stream_t s, nodey_event_t e);
int nodey_malloc(void** out, size_t bytes, nodey_stream_t s);
int nodey_free(void* p, nodey_stream_t s);
int nodey_trim_memory(void);   /* return cached blocks to the driver */
/* device memory held through nodey_malloc right now and its high-water mark (blocks >= 1 MiB); reset_peak != 0
 * restarts the mark at the current level */
int nodey_memory_stats(int64_t* live_bytes, int64_t* peak_bytes, int reset_peak);
/* device memory the library has obtained from the driver and not returned (held by callers + cached for reuse): the
 * footprint a render really has, and its high-water mark since the last nodey_memory_stats(..., reset_peak = 1) */
int nodey_memory_reserved(int64_t* reserved_bytes, int64_t* peak_reserved_bytes);
/* reuse_pending != 0: nodey_malloc may hand out a block that another stream has freed although the device has not
 * reached that free yet; the requesting stream then waits for it (less memory, more cross-stream ordering).  Default 0:
 * such a block is only reused by the stream that freed it, or once the free has completed. */
int nodey_set_memory_policy(int reuse_pending);
int nodey_memset(void* dst, int value, size_t bytes, nodey_stream_t s);
int nodey_memcpy_h2d(void* dst, const void* src_host, size_t bytes, nodey_stream_t s);
/* rows x width_bytes from host rows src_pitch apart to device rows dst_pitch apart, one enqueue (cudaMemcpy2DAsync) */
int nodey_memcpy2d_h2d(void* dst, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t width_bytes, size_t rows, nodey_stream_t s);
int nodey_memcpy_d2h(void* dst_host, const void* src, size_t bytes, nodey_stream_t s);
int nodey_memcpy_d2d(void* dst, const void* src, size_t bytes, nodey_stream_t s);
int nodey_host_alloc(void** out, size_t bytes);      /* pinned host memory */
int nodey_host_free(void* p);

/* Synthetic source of SURVEY.md 8(d): x = 0.5 sin(2 pi f n / sr) + 0.05 u, bit-identical to the
 * oracle generator.  dst: interleaved float [nframes][nch]; optional s16 copy clip(lrintf(x*32767)). */
int nodey_synth(float* dst_f32, int16_t* dst_s16, int64_t nframes, int nch, int sample_rate,
                int track, int64_t frame0, nodey_stream_t stream);

/* A3  audio_volume_adjust -- change_volume<T>, src/processor/audio-vol.cpp:75-100.
 * dst[i] = T(src[i] * volume) over n_elems samples of one plane (packed: frames*channels).
 * Integer formats follow the x86 truncating conversion the reference compiles to (no clamp). */
int nodey_gain(void* dst, const void* src, int fmt, int64_t n_elems, float volume, nodey_stream_t stream);
/* The same for a batch of float streams in one launch (HOST arrays of ntracks entries: device pointers, sample
 * counts, gains): the per-track audio_volume_adjust nodes of one graph level. */
int nodey_gain_tracks(void* const* dst, const void* const* src, const int64_t* n, const float* volumes, int fmt,
                      int ntracks, nodey_stream_t stream);

/* A8  extract_samples_interleaved, src/processor/audio-velocity.cpp:150-232.
 * Any of the six formats -> interleaved float, with the reference's four integer scales. */
int nodey_extract_interleaved(float* dst, const void* plane0, const void* plane1, int fmt,
                              int64_t nframes, int nch, nodey_stream_t stream);

/* N1  audio_channel_split (new node): stereo -> two mono planes of the same sample
 * type; pure routing, bit exact. */
int nodey_split(void* dst_l, void* dst_r, const void* plane0, const void* plane1, int fmt,
                int64_t nframes, nodey_stream_t stream);

/* swr format conversion + rematrix with no rate change (libswresample audioconvert/rematrix as
 * configured at audio-amix.cpp:212-243): any format, mono|stereo -> stereo float planar.
 * mono -> both channels = in * sqrt(1/2). */
int nodey_to_fltp_stereo(float* dst_l, float* dst_r, const void* plane0, const void* plane1, int fmt,
                         int nch, int64_t nframes, nodey_stream_t stream);

/* A4  audio_amix mix loop, src/processor/audio-amix.cpp:293-307.
 * out[j] = ((0 + in0[j]*v0) + in1[j]*v1) + ... in input order; input i contributes zeros for
 * j >= in_len[i] (the reference's zero-filled temp buffers).  in_l / in_r / in_len / volumes are
 * HOST arrays of nin entries (device pointers inside).  in_r[i] == NULL: in_l[i] holds interleaved
 * stereo float frames (swr's FLT -> FLTP conversion is a pure de-interleave, done in registers). */
int nodey_mix(float* out_l, float* out_r, const float* const* in_l, const float* const* in_r,
              const int64_t* in_len, const float* volumes, int nin, int64_t nframes,
              nodey_stream_t stream);
/* The same with an audio_volume_adjust node folded into every input (gains: HOST array of nin floats, NULL = none):
 * out = sum_i (in_i * gain_i) * vol_i, the inner product rounded to float first -- exactly what the gain node's own pass
 * would have stored (dst = T(src * volume), audio-vol.cpp:75-100) -- so the bus is bit identical and the gain pass with
 * its intermediate stream disappears.  Float inputs only (the integer gain truncates, that stays a kernel of its own). */
int nodey_mix_gains(float* out_l, float* out_r, const float* const* in_l, const float* const* in_r,
                    const int64_t* in_len, const float* volumes, const float* gains, int nin, int64_t nframes,
                    nodey_stream_t stream);

/* A5  audio_bimix mix loop, src/processor/audio-bimix.cpp:310-317.
 * outL = (ll/2 + lr/2) * (1 - bias), outR = (rl/2 + rr/2) * (1 + bias); zeros past len_l / len_r. */
int nodey_bimix(float* out_l, float* out_r, const float* ll, const float* lr, int64_t len_l,
                const float* rl, const float* rr, int64_t len_r, float bias, int64_t nframes,
                nodey_stream_t stream);

/* Preview sink, src/processor/audio-io.cpp:478-638 (do_preview): after swr has brought the stream to 48 kHz
 * stereo float, every sample is clamped to [-1, 1] (std::clamp, :598-599) and queued as packed frames.
 * dst[2j] = clamp(l[j]), dst[2j+1] = clamp(r[j]). */
int nodey_preview_pack(float* dst, const float* l, const float* r, int64_t nframes, nodey_stream_t stream);

/* A6  audio_bimix_v2 pieces, src/processor/audio-bimix.cpp:625-627 and :777-872.
 * downmix: dst = (l + r) * 0.5.  merge: interleaved stereo out; for frame j of segment s
 * (seg_out_start[s] <= j < seg_out_start[s] + seg_len[s]) L = left[seg_l[s] + d] or 0 when
 * seg_l[s] < 0, same for R.  Segment arrays are HOST arrays (nseg entries). */
int nodey_downmix_half(float* dst, const float* l, const float* r, int64_t n, nodey_stream_t stream);
int nodey_merge_segments(float* out_interleaved, const float* left, const float* right,
                         const int64_t* seg_out_start, const int64_t* seg_len,
                         const int64_t* seg_l, const int64_t* seg_r, int nseg, nodey_stream_t stream);

/* A7  Audio_resampler / swr_convert rate conversion (src/utility/sw-resample.cpp:8-23,
 * include/utility/sw-resample.hpp:55-70; call sites audio-amix.cpp:263-290,
 * audio-bimix.cpp:259-294): libswresample with all-default options -> Kaiser(9) windowed sinc,
 * 32 taps (more when down-sampling), exact-rational polyphase, mirrored start, reflected flush.
 * The plan owns the device copy of the filter bank. */
typedef struct nodey_resampler nodey_resampler;
int nodey_resampler_create(nodey_resampler** out, int in_rate, int out_rate, int index_mask_quirk);
void nodey_resampler_destroy(nodey_resampler* r);
/* info[8]: phase_count, filter_length, filter_alloc, dst_incr_div, dst_incr_mod, src_incr, index0, linear */
int nodey_resampler_info(const nodey_resampler* r, int info[8]);
/* HOST pointer to the (phase_count+1) x filter_alloc float filter bank */
const float* nodey_resampler_filter_bank(const nodey_resampler* r);
/* frames swr would return in total for in_frames of input (flush: after swr_convert(NULL) drain) */
int64_t nodey_resampler_out_count(const nodey_resampler* r, int64_t in_frames, int flush);
/* streaming bookkeeping of swr_convert (host only): outputs available from n_in real frames plus
 * `reflect` reflected ones; reflection length resample_flush() appends once `produced` were taken */
int64_t nodey_resampler_producible(const nodey_resampler* r, int64_t n_in, int64_t reflect);
int64_t nodey_resampler_flush_reflect(const nodey_resampler* r, int64_t n_in, int64_t produced);
/* Whole-track conversion: source in its native format (converted on load, mono rematrixed),
 * stereo float planar out.  out_frames <= nodey_resampler_out_count(), with one exception the library has too: a
 * flushed conversion that was drained through small output capacities (audio_amix's nb next to a faster input) still
 * buffers more than filter_length frames when resample_flush() runs, reflects one frame more and can return one more
 * output than the same input converted with ample capacity; flush = 1 accepts that count as well (the bookkeeping
 * that says when it happens is nodey_resampler_flush_reflect / nodey_amix_plan). */
int nodey_resampler_run(const nodey_resampler* r, float* out_l, float* out_r,
                        const void* plane0, const void* plane1, int fmt, int nch, int64_t in_frames,
                        int flush, int64_t out_frames, nodey_stream_t stream);
/* Test hook: as nodey_resampler_run with the kernel forced (mode 0 auto, 1 generic one-thread-per-
 * output kernel, 2 tiled kernel; 2 fails with NODEY_E_RANGE when the plan has no tile tables). */
int nodey_resampler_run_mode(const nodey_resampler* r, float* out_l, float* out_r,
                             const void* plane0, const void* plane1, int fmt, int nch, int64_t in_frames,
                             int flush, int64_t out_frames, int mode, nodey_stream_t stream);
/* Fused A7 + A4 for inputs that share one plan (same source rate): out = sum_i vol_i * resample(in_i)
 * in input order, zeros past each input's own length.  Host arrays of nin entries. */
int nodey_resample_mix(const nodey_resampler* r, float* out_l, float* out_r,
                       const void* const* plane0, const void* const* plane1, const int* fmt,
                       const int* nch, const int64_t* in_frames, const int64_t* out_len,
                       const float* volumes, int nin, int flush, int64_t out_frames,
                       nodey_stream_t stream);
/* The same for a BATCH of independent single-input mixers (the per-track audio_amix(1) resamplers of a render) in
 * one launch: track t reads plane0[t] / plane1[t] (HOST arrays of device pointers; all tracks share format,
 * channel count and length), is scaled by volumes[t] (temp = 0 + data*volume, audio-amix.cpp:296-304) and
 * written to out_l/out_r + t*out_track_stride.  At most 256 tracks per call; exact-rational plans with at most
 * 160 phases (NODEY_E_RANGE otherwise: fall back to nodey_resample_mix per track). */
int nodey_resample_tracks(const nodey_resampler* r, float* out_l, float* out_r, int64_t out_track_stride,
                          const void* const* plane0, const void* const* plane1, int fmt, int nch, int64_t in_frames,
                          const float* volumes, int ntracks, int flush, int64_t out_len, int64_t out_frames,
                          nodey_stream_t stream);

/* nodey_resample_tracks cut into launches along time (like nodey_soundtouch_chunks): chunk c may run once input frames
 * [0, in_need[c]) of every track are there and makes output frames [0, out_ready[c]) final; the last chunk needs the whole
 * input (the flush reflects its end).  Chunks run in order, with the same arguments; the result is bit identical to the
 * one-launch call.  nodey_resample_tracks_chunks returns the number of chunks (<= want_chunks, at least four tiles each). */
int nodey_resample_tracks_chunks(const nodey_resampler* r, int nch, int64_t in_frames, int64_t out_frames, int want_chunks,
                                 int64_t* in_need, int64_t* out_ready, int cap);
int nodey_resample_tracks_chunk(const nodey_resampler* r, float* out_l, float* out_r, int64_t out_track_stride,
                                const void* const* plane0, const void* const* plane1, int fmt, int nch, int64_t in_frames,
                                const float* volumes, int ntracks, int flush, int64_t out_len, int64_t out_frames,
                                int chunk, int nchunks, nodey_stream_t stream);

/* Time-segment sharding of one long stream across GPUs: outputs [k0, k1) of the flushed
 * conversion of n_in frames depend only on the input slice [*in0, *in1).  Run the resampler on that slice with
 * flush = *flush and out_frames = *skip + (k1 - k0), drop the first *skip outputs: the rest is bit identical
 * to the whole-stream result (one period of phase_count outputs consumes exactly dst_incr_div input frames, so
 * a conversion started whole periods early reproduces every phase; the lead-in spans at least (taps-1)/2 input
 * frames, so only dropped outputs see the mirrored left edge).  k0 must be a multiple of phase_count; exact-rational plans only (NODEY_E_RANGE otherwise).
 * Streaming nodes (gain, split, merge, mix) cut anywhere without halo; the spectrum node needs
 * [m0*hop, (m1-1)*hop + nfft) for frames [m0, m1); SoundTouch state is sequential and cannot be cut. */
int nodey_resampler_segment(const nodey_resampler* r, int64_t n_in, int64_t k0, int64_t k1,
                            int64_t* in0, int64_t* in1, int64_t* skip, int* flush);

/* A4  audio_amix frame bookkeeping (host only, no device work), audio-amix.cpp:149-322.  The node pulls
 * one frame per input and iteration, asks swr for nb = min(frame sizes) frames (1152 once the inputs
 * ran dry) and zero fills what swr did not deliver, until every input is flushed.  Inputs are
 * described by their rate and their frame sizes, run-length encoded: input i owns runs
 * run_off[i] .. run_off[i+1]-1, run r = run_count[r] frames of run_len[r] samples.  Returns the
 * node's total output frames (whole iterations, i.e. zero padded) and
 *   - where every input's resampled frames land: segment k copies seg_len[k] frames of input
 *     seg_input[k], from its resampled frame seg_src_start[k], to output frame seg_out_start[k]
 *     (up-sampled inputs: one segment from 0; inputs above 48 kHz deliver less than nb per iteration
 *     and leave gaps -- reference behaviour);
 *   - the frame sizes of the node's own output stream, run-length encoded (out_run_*).
 * At most seg_cap / out_run_cap entries are written; *nseg_out / *n_out_runs receive the numbers
 * needed.  Negative return = error code. */
int64_t nodey_amix_plan(const int* in_rate, int nin, const int64_t* run_off, const int64_t* run_len,
                        const int64_t* run_count, int index_mask_quirk,
                        int32_t* seg_input, int64_t* seg_out_start, int64_t* seg_src_start, int64_t* seg_len,
                        int64_t seg_cap, int64_t* nseg_out,
                        int64_t* out_run_len, int64_t* out_run_count, int64_t out_run_cap, int64_t* n_out_runs);

/* Launch accounting: every kernel launch of the library is counted; with profiling enabled each
 * launch is bracketed by CUDA events on its stream and nodey_profile_report() writes a JSON object
 * {"kernel": {"launches": n, "ms": device ms, "bytes": 0}} of the launches since the last report. */
void nodey_profile_enable(int on);
uint64_t nodey_profile_launches(void);
int nodey_profile_report(char* buf, int cap);

/* A9  pitch_modifier / velocity_modifier -- soundtouch_process_payload, src/processor/audio-velocity.cpp:265-443:
 * new SoundTouch; setSampleRate; setChannels; setRate(rate); setPitch(pitch) (:367-390), then
 * putSamples per frame, receiveSamples, flush (:399-435).  SoundTouch 2.3.2 float build with default
 * settings (TDStretch WSOLA, 64-tap AA filter, cubic transposer); see SURVEY.md App. B2.
 *   Pitch_modifier   : rate 1,        pitch powf(2, semitones/12)          (:462-477)
 *   Velocity_modifier: rate velocity, pitch keep_pitch ? 1/velocity : 1    (:445-460)
 * The object is the whole-track equivalent of the streaming one: `ntracks` tracks of equal length
 * (interleaved float, track t at in + t*in_stride floats) are rendered in one call.  frame_size is
 * the putSamples chunk (it only enters the expected-output bookkeeping of flush()).
 * offsets (optional, device): the WSOLA offset trace, n_sequences-1 ints per track. */
typedef struct nodey_soundtouch nodey_soundtouch;
int nodey_soundtouch_create(nodey_soundtouch** out, int sample_rate, int channels, float rate, float pitch);
void nodey_soundtouch_destroy(nodey_soundtouch* s);
/* info_i[8]: overlap, seek_window, seek_length, sample_req, tdstretch_first, prefill, channels, sample_rate
 * info_d[3]: effective rate, effective tempo, nominal_skip */
int nodey_soundtouch_info(const nodey_soundtouch* s, int info_i[8], double info_d[3]);
/* frames the node emits in total for in_frames of input (after flush); n_sequences optional */
int64_t nodey_soundtouch_out_frames(nodey_soundtouch* s, int64_t in_frames, int frame_size, int64_t* n_sequences);
/* Test hook: cluster size of the WSOLA offsets kernel (thread-block cluster of 1, 2 or 4 CTAs per track;
 * 0 = automatic: the largest that keeps tracks * cluster * 2 <= SM count). */
int nodey_soundtouch_set_cluster(nodey_soundtouch* s, int cluster);
/* Test hook: candidates per thread of the WSOLA offsets kernel (4, 8 or 11..16; 0 = automatic: the smallest count that
 * fits one (SSE lane, candidate class) stream into one warp when a track runs on one CTA, 8 resp. 4 with clusters). */
int nodey_soundtouch_set_candidates_per_thread(nodey_soundtouch* s, int kt);
/* Test hook: 1 = run cross-fade, AA FIR and cubic transposer as separate kernels even where the fused
 * tail kernel applies (stereo, TDStretch-first order). */
int nodey_soundtouch_set_unfused(nodey_soundtouch* s, int unfused);
int nodey_soundtouch_run(nodey_soundtouch* s, float* out, int64_t out_stride, const float* in, int64_t in_stride,
                         int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                         int32_t* offsets, int64_t offsets_stride, nodey_stream_t stream);
/* Same, with the tracks read in place through per-track pointers (HOST arrays of device pointers, at most
 * 256 tracks per call): in_a[t] = interleaved float frames, or -- for stereo with in_b != NULL -- the left
 * plane with in_b[t] the right plane.  extract_samples_interleaved (audio-velocity.cpp:150-232) is the
 * identity on float samples, so FLT / FLTP streams need no staging copy. */
int nodey_soundtouch_run_tracks(nodey_soundtouch* s, float* out, int64_t out_stride, const float* const* in_a, const float* const* in_b,
                                int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                                int32_t* offsets, int64_t offsets_stride, nodey_stream_t stream);

/* SURVEY.md App. C7 compatibility: the node loop of the reference (audio-velocity.cpp:286-441), run with an input frame
 * available at every turn, receives min(numSamples, 3 * 1152 / velocity) samples whenever more than 1152 / velocity are
 * queued and leaves through `if (numSamples() == 0 && input_stream_eof) break;` (:414) before flush() whenever its last
 * receive emptied SoundTouch's output -- the normal case -- so the tail SoundTouch still holds is never emitted.  This
 * returns what that loop emits in total (always a prefix of the canonical render: pass it to run() as out_frames) and
 * the frame sizes it pushes downstream (run-length encoded; *n_runs receives the number of runs needed), *flushed = 1
 * when flush() was reached.  velocity: the node's velocity (1 for pitch_modifier).  Host arithmetic only. */
int64_t nodey_soundtouch_reference_schedule(nodey_soundtouch* s, int64_t in_frames, int frame_size, float velocity,
                                            int64_t* run_len, int64_t* run_count, int64_t run_cap, int64_t* n_runs, int* flushed);

/* The same render cut into launches along time, so that a consumer can start before the producer has finished (the
 * reference pipelines its nodes frame by frame, audio-velocity.cpp:286-441; here the unit is a chunk of WSOLA sequences):
 * nodey_soundtouch_chunks plans at most want_chunks launches and returns how many there are (1 for the paths that
 * cannot be cut: mono, rate <= 1); chunk c may run once input frames [0, in_need[c]) are final and makes output frames
 * [0, out_ready[c]) final (the last chunk needs / makes everything).  run_chunk / run_tracks_chunk are run / run_tracks
 * restricted to chunk `chunk` of `nchunks`; chunks run in order on one stream, with the SAME arguments, and `offsets`
 * (required here: n_sequences - 1 ints per track, device) carries the chain from one chunk to the next.  The result is
 * bit identical to the one-launch render.
 * phase: 0 = the chunk's WSOLA search and its tail (cross-fade + AA FIR + cubic transposer over the tiles those sequences
 * complete) on `stream`; 1 = the search only, 2 = the tail only -- so that a caller can keep the sequential search chain
 * on a stream of its own, back to back, and run the tails next to it (tail c after search c; search c + 1 does not wait
 * for tail c).  A chunk's input and output ranges (in_need / out_ready) are those of both phases together. */
int nodey_soundtouch_chunks(nodey_soundtouch* s, int64_t in_frames, int frame_size, int64_t out_frames, int want_chunks,
                            int64_t* in_need, int64_t* out_ready, int cap);
int nodey_soundtouch_run_chunk(nodey_soundtouch* s, float* out, int64_t out_stride, const float* in, int64_t in_stride,
                               int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                               int32_t* offsets, int64_t offsets_stride, int chunk, int nchunks, int phase, nodey_stream_t stream);
int nodey_soundtouch_run_tracks_chunk(nodey_soundtouch* s, float* out, int64_t out_stride, const float* const* in_a, const float* const* in_b,
                                      int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                                      int32_t* offsets, int64_t offsets_stride, int chunk, int nchunks, int phase, nodey_stream_t stream);

/* N2  audio_spectrum (new node; FFTW r2c convention, unnormalised):
 * per channel, frame m = x[m*hop .. m*hop+nfft) * periodic Hann; out[ch][m][0..nfft/2] complex64
 * (re, im interleaved).  interleaved != 0: sample f of channel c at x[f*nch + c]; otherwise planar
 * with channel c at x + c*plane_stride.  nfft must be 4096 in this release. */
int64_t nodey_stft_frames(int64_t nframes, int nfft, int hop);
int nodey_stft(float* out_complex, const float* x, int64_t nframes, int nch, int interleaved,
               int64_t plane_stride, int nfft, int hop, nodey_stream_t stream);

/* Master-bus reduce across the GPUs of one box.  The reference is a single-process CPU program
 * and has no counterpart; what is replaced is the LAST sum of the graph: the master audio_amix adds its inputs in
 * order (audio-amix.cpp:293-307), and when the tracks are sharded over ranks (one process per GPU) every rank holds
 * a partial FLTP bus that still has to be added.  This is the only exchange step of the render, hence the only
 * collective: ncclReduce / ncclAllReduce(sum, float32) over NVLink, both planes in one NCCL group, asynchronous on
 * `stream`.  The sum across ranks is not the sequential input order of the CPU mixer, so the reduced bus is compared
 * within 1e-5, not bit for bit (deterministic for a fixed rank count).
 * NCCL is bound at run time (dlopen of libnccl.so.2, NODEY_NCCL_LIB overrides): single-GPU users never load it, and
 * a process that already carries an NCCL shares it.  NODEY_E_COMM when it is missing -- there is no fallback.
 * Set-up: rank 0 calls nodey_bus_unique_id and hands the NODEY_BUS_ID_BYTES bytes to every rank (launcher, file,
 * socket: out of band); every rank then calls nodey_bus_create on ITS device (nodey_set_device first); the call
 * returns when all ranks have joined. */
#define NODEY_BUS_ID_BYTES 128
typedef struct nodey_bus nodey_bus;
int nodey_bus_nccl_version(int* version);
int nodey_bus_unique_id(void* id);
int nodey_bus_create(nodey_bus** out, const void* id, int rank, int nranks);
void nodey_bus_destroy(nodey_bus* bus);
int nodey_bus_info(const nodey_bus* bus, int* rank, int* nranks, int* device);
/* recv = sum over ranks of send, on `root` only (recv may be NULL elsewhere; send == recv is allowed on the root).
 * send_r / recv_r = second plane of the FLTP bus, NULL for a mono bus.  nframes must agree on every rank. */
int nodey_bus_reduce(nodey_bus* bus, const float* send_l, const float* send_r, float* recv_l, float* recv_r,
                     int64_t nframes, int root, nodey_stream_t stream);
/* the same with the sum delivered to every rank (when each rank goes on with a time slice of the bus, e.g. its share
 * of the spectrum frames) */
int nodey_bus_allreduce(nodey_bus* bus, const float* send_l, const float* send_r, float* recv_l, float* recv_r,
                        int64_t nframes, nodey_stream_t stream);

/* Peer memory: the master mix as one kernel over NVLink instead of "mix, then reduce".  nodey_peer_alloc makes a
 * block that can be exported to the other processes of the box (CUDA IPC; plain cudaMalloc, not the stream-ordered
 * pool); nodey_peer_export writes its NODEY_PEER_HANDLE_BYTES handle; a rank that nodey_peer_open()s the handle holds an
 * ordinary device pointer whose loads travel over NVLink.  nodey_mix reads its inputs wherever they live, so the root
 * runs the graph's master audio_amix (audio-amix.cpp:293-307) over the group mixes of EVERY rank in the graph's input
 * order: compute and exchange are one kernel, and the bus is BIT IDENTICAL to the one-GPU render (nodey_bus_reduce adds
 * partial buses in another order: 1e-5).  Cross-process ordering is the caller's: a peer's block may be read once that
 * peer has synchronised its writes and said so (a barrier), and rewritten once the reader has finished. */
#define NODEY_PEER_HANDLE_BYTES 64
int nodey_peer_alloc(void** out, size_t bytes);
int nodey_peer_free(void* p);
int nodey_peer_export(const void* p, void* handle);
int nodey_peer_open(void** out, const void* handle);
int nodey_peer_close(void* p);

#ifdef __cplusplus
}
#endif
#endif
