/*
 * nodey_engine.h -- C facade over the C++ host layer (nodey-audio-editor_b200/host): load a Nodey project
 * (the JSON written by Graph::serialize, src/infra/graph.cpp:284-372), bind PCM sources to the
 * audio_input node, run the graph with infra::Runner::create_and_run (src/infra/runner.cpp:142-154) and
 * read what arrived at audio_output or on any link.  This is what a non-C++ caller (tests, bench.py, a
 * scripting front end) binds; C++ callers use the infra:: / processor:: classes directly.
 *
 * All functions return 0 on success, negative on failure (nodey_engine_last_error() has the text).
 * Device pointers returned here stay valid until the engine is run again or destroyed.
 */
#ifndef NODEY_ENGINE_H
#define NODEY_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nodey_engine nodey_engine;

const char* nodey_engine_last_error(void);

/* infra::register_all_processors() + Graph::deserialize(Json).  NODEY_ENGINE_E_FILE for Invalid_file_error. */
int nodey_engine_create(nodey_engine** out, const char* project_json);
void nodey_engine_destroy(nodey_engine* e);
/* Registers the example processors that are not part of the reference's set (src/register.cpp:14-23):
 * "frame_gain_example", a node written against the reference's FRAME interface (Audio_stream::try_pop /
 * try_push / set_eof, src/processor/audio-stream.cpp:60-80) -- the frame-streaming compatibility mode that lets
 * processors which were never ported to device buffers run inside this engine.  Call before nodey_engine_create. */
int nodey_engine_register_examples(void);

/* Graph::serialize() with a 2-space indent (src/frontend/app.cpp:837-839).  Returns the length needed. */
int nodey_engine_serialize(nodey_engine* e, char* buf, int cap);
/* Graph::check_graph(): 0, or NODEY_ENGINE_E_GRAPH (mismatched pin, multiple input, loop) */
int nodey_engine_check(nodey_engine* e);
int nodey_engine_node_count(nodey_engine* e);
/* k-th node in id order: id, identifier (copied into buf), graph level (-1 when the graph is invalid) */
int nodey_engine_node_info(nodey_engine* e, int k, int* node_id, char* identifier, int cap, int* level);

/* audio_volume_adjust's gain is not part of the project file: programmatic setter */
int nodey_engine_set_volume(nodey_engine* e, int node_id, float volume);

/* PCM source for pin output_{index} of the audio_input node (replaces the file decode of
 * src/processor/audio-io.cpp:69-300).  The memory must stay valid until nodey_engine_run returns;
 * host memory should be pinned for the upload to overlap.  data1 = second plane of planar stereo. */
int nodey_engine_bind_source(nodey_engine* e, int index, const void* data, const void* data1, int on_device,
                             int fmt, int sample_rate, int channels, int64_t frames, int frame_size,
                             double pts_seconds);

/* Runner::create_and_run + wait.  Errors of nodes (Processor::Runtime_error ...) -> NODEY_ENGINE_E_NODE. */
int nodey_engine_run(nodey_engine* e);

/* Product on output pin `pin` of node `node_id` after a run (first link on that pin).
 * kind: 1 = audio stream, 2 = spectrum.  audio: fmt/rate/channels/frames/pts, planes.
 * spectrum: fmt = 3, channels, frames = STFT frames, plane0 = complex64 [channels][frames][bins], bins in *extra. */
int nodey_engine_product(nodey_engine* e, int node_id, const char* pin, int* kind, int* fmt, int* sample_rate,
                         int* channels, int64_t* frames, double* pts_seconds, void** plane0, void** plane1, int* extra);
/* what arrived at the audio_output sink (same fields) */
int nodey_engine_output(nodey_engine* e, int* fmt, int* sample_rate, int* channels, int64_t* frames,
                        double* pts_seconds, void** plane0, void** plane1);
/* Process_context::export_path of the sink (src/frontend/app.cpp:2067-2073): "" keeps the result in memory only, a
 * path ending in ".wav" makes the export also write an interleaved 32-bit float WAV, with do_export's pts rule:
 * (int)((frame stamp - time) * sample_rate) frames of silence in front of every frame where that is positive
 * (src/processor/audio-io.cpp:833-839; nodey_engine_export_plan below states the arithmetic). */
int nodey_engine_set_export_path(nodey_engine* e, const char* path);
/* Any other non-empty path is an MP3 export like the reference's (Audio_output::do_export, audio-io.cpp:640-841): the
 * stream goes to LAME in its own sample format, frame by frame in the sizes its producer recorded, with the same
 * parameters (quality 2, CBR at `kbps`, 48 kHz out), the same silence rule and -- like the reference -- without a final
 * lame_encode_flush.  libmp3lame is bound at run time (libmp3lame.so.0, NODEY_LAME_LIB overrides); when it cannot be
 * loaded the run fails with NODEY_ENGINE_E_NODE ("MP3 encoder not available").  kbps: 8..320, default 320
 * (src/frontend/app.cpp:595).  nodey_engine_mp3_available: 1 when the encoder library could be bound. */
int nodey_engine_set_export_kbps(nodey_engine* e, int kbps);
int nodey_engine_mp3_available(void);
/* The same MP3 leg on a HOST stream (packed: plane0; planar stereo: plane0 / plane1), cut into frame_size frames:
 * for callers that hold the rendered PCM already (e.g. the reduced master bus of a multi-GPU render).
 * *time_inout = Process_context::time before / after (NULL = start at 0). */
int nodey_engine_encode_mp3(const char* path, const void* plane0, const void* plane1, int fmt, int sample_rate, int channels,
                            int64_t frames, int frame_size, double pts_seconds, int kbps, double* time_inout);
/* What the audio_input node publishes for a RIFF/WAVE file, from the header alone (no device needed): the sample format as
 * the reference's decoder would hand it on (PCM 16 -> S16, PCM 24 / 32 -> S32 with 24-bit samples in the upper three bytes,
 * IEEE float 32 -> FLT: libavcodec/pcm.c), rate, channels, sample frames, and the frame size of the stream -- libavformat's
 * wav demuxer reads packets of 4096 bytes rounded down to whole blocks and the PCM decoder returns one frame per packet
 * (src/processor/audio-io.cpp:176-222 pushes every decoded frame as it is), i.e. 4096 / block_align sample frames: 1024 for
 * 16-bit stereo, 512 for float stereo, 682 for 24-bit stereo.  Any output pointer may be NULL.
 * NODEY_ENGINE_E_FILE for files the source node refuses ("Cannot open audio file"). */
int nodey_engine_probe_wav(const char* path, int* fmt, int* sample_rate, int* channels, int64_t* frames, int* frame_size);
/* do_export's per-frame bookkeeping as plain arithmetic (audio-io.cpp:826-839): for a stream of `frames` samples per channel
 * cut into the given frame runs, whose frames are stamped by rule `stamp` -- 0: exact start times from `origin` (decoder
 * stamps), 1: running END time from `origin` truncated to whole microseconds (audio_amix / audio_bimix,
 * audio-amix.cpp:199-201), 2: running start time from `origin` through a float of microseconds (SoundTouch nodes,
 * audio-velocity.cpp:238-249, 313-318) -- fills silence[k] = samples of silence the export encodes in front of frame k
 * (up to cap entries; any of the output pointers may be NULL) and *time_inout = Process_context::time before / after.
 * Returns the number of frames.  nodey_engine_product_stamp: rule and origin of an audio product of the last run. */
int nodey_engine_export_plan(int stamp, double origin, int sample_rate, int64_t frames, const int64_t* run_len, const int64_t* run_count,
                             int n_runs, double* time_inout, int64_t* silence, double* frame_pts, int cap);
int nodey_engine_product_stamp(nodey_engine* e, int node_id, const char* pin, int* stamp, double* origin);
/* Preview instead of export (the reference's Preview state, src/frontend/app.cpp:2001-2040 -> Audio_output::do_preview,
 * src/processor/audio-io.cpp:478-638): the sink brings the stream to 48 kHz stereo float frame by frame without a
 * final flush, clamps to [-1, 1] and queues packed frames.  nodey_engine_preview returns that queue content (device
 * pointer, packed stereo float) and the chunk sizes the sink callback received (one per input frame); the return
 * value is the number of chunks. */
int nodey_engine_set_preview(nodey_engine* e, int preview);
int nodey_engine_preview(nodey_engine* e, int64_t* frames, void** packed, int64_t* chunk_len, int chunk_cap);
/* Scheduling knobs of this engine's runs (infra::Runner::Schedule): "wave_pins" (waves of n source pins), "wave_pattern"
 * ("32,64,48": explicit wave sizes, the last one repeats), "compute_lanes" (1..4), "side_streams" (0: no chunk-wise overlap
 * of a chain's nodes), "stream_priority" (0: no priority for the WSOLA search streams), "stream_chunks" (1..64 launches a
 * stream is cut into along time), "trace" / "timing" (development output on stderr).  0 resp. -1 = automatic: the values
 * DESIGN.md 2.1 measured for BASELINE's render shapes on a B200.  The NODEY_* environment variables of the same names are
 * development overrides read once per run; an explicit setting wins over them. */
int nodey_engine_set_schedule(nodey_engine* e, const char* key, const char* value);
/* Memory policy of the runs started after the call (process wide; infra::Runner::release_products).  0 (default): every
 * link keeps its product until the next run, like the reference's channels (include/infra/runner.hpp:60-84), so any
 * product can be read back with nodey_engine_product.  1: a link lets go of its product once its consumer has enqueued
 * its work; only the sink's stream (nodey_engine_output) and spectrum products survive the run, and a render holds one
 * or two levels of intermediates per wave instead of all of them. */
int nodey_engine_set_release_products(int release);
/* frame sizes (run-length encoded) of an audio product: fills up to cap pairs, returns the count */
int nodey_engine_product_runs(nodey_engine* e, int node_id, const char* pin, int64_t* run_len, int64_t* run_count, int cap);

/* Diagnostics of the last run (the reference's overlay, src/frontend/app.cpp:1556-1592, lists node
 * states and the fill of every link's channel -- links here are published once, so the per-link figure is replaced by
 * the device time of every (wave, level) step of the Runner).  nodey_engine_diagnostics: the overlay's "Audio" block as
 * text ("N Running | N Finished | N Errors", then one line per step); returns the length needed.
 * nodey_engine_level_timings: the same figures as arrays (any pointer may be NULL); device_ms = span between the
 * step's first and last command on its lane, start_ms = device time since the run's first command; returns the
 * number of steps. */
int nodey_engine_diagnostics(nodey_engine* e, char* buf, int cap);
int nodey_engine_level_timings(nodey_engine* e, int* wave, int* level, int* lane, int* nodes, double* enqueue_ms, double* device_ms,
                               double* start_ms, int cap);

enum { NODEY_ENGINE_E_INVALID = -1, NODEY_ENGINE_E_FILE = -2, NODEY_ENGINE_E_GRAPH = -3, NODEY_ENGINE_E_NODE = -4 };

#ifdef __cplusplus
}
#endif
#endif
