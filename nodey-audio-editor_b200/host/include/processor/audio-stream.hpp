// processor/audio-stream.hpp -- the product that travels over audio links.
// Reference: Audio_stream = bounded channel of AVFrame-backed Audio_frames + EOF flag
// (include/processor/audio-stream.hpp:22-83).  Here a stream is published ONCE as a device-resident
// whole-track buffer (Audio_buffer): samples in the layout the reference's frames would carry
// (AVSampleFormat numbering, packed or planar), plus the metadata the reference reads from frames:
// sample rate, channel count, start pts and the frame sizes (run-length encoded) that the
// frame-by-frame bookkeeping of amix / bimix depends on.
#pragma once

#include "infra/exec-context.hpp"
#include "infra/processor.hpp"

#include <cstdint>
#include <memory>
#include <mutex>
#include <optional>
#include <utility>
#include <vector>

namespace processor
{
	// AVSampleFormat values (libavutil/samplefmt.h), as stored in AVFrame::format by the reference
	enum Sample_format : int
	{
		FMT_U8 = 0, FMT_S16 = 1, FMT_S32 = 2, FMT_FLT = 3, FMT_DBL = 4,
		FMT_U8P = 5, FMT_S16P = 6, FMT_S32P = 7, FMT_FLTP = 8, FMT_DBLP = 9
	};

	inline bool format_is_planar(int fmt) { return fmt >= FMT_U8P; }
	inline int format_bytes(int fmt)
	{
		switch (fmt)
		{
		case FMT_S16: case FMT_S16P: return 2;
		case FMT_S32: case FMT_S32P: case FMT_FLT: case FMT_FLTP: return 4;
		default: return 0;
		}
	}

	// frame sizes of a stream, run-length encoded: {samples per frame, number of such frames}
	using Frame_runs = std::vector<std::pair<int64_t, int64_t>>;
	Frame_runs uniform_frame_runs(int64_t frames, int64_t frame_size);
	int64_t frame_runs_total(const Frame_runs& runs);

	// How the producer of a stream stamped its frames.  The reference's sinks read PER-FRAME stamps (do_export encodes
	// (int)((frame.pts * time_base - time) * rate) samples of silence before every frame, audio-io.cpp:833-839), and its
	// nodes stamp in three different ways, so the rule travels with the stream:
	enum Pts_stamp : int
	{
		STAMP_START = 0,           // decoder stamps and everything that forwards them: frame k starts at pts + at_k / rate
		STAMP_END_US = 1,          // audio_amix / audio_bimix: the running END time, `time_seconds += nb / double(rate); pts =
		                           // time_seconds * 1000000` truncated to whole microseconds (audio-amix.cpp:199-201,
		                           // audio-bimix.cpp:188-191, App. C4)
		STAMP_START_FLOAT_US = 2,  // SoundTouch nodes under the reference schedule: running START time handed on as a FLOAT of
		                           // microseconds (audio-velocity.cpp:238-249, 313-318, App. C8): 64 us steps after 10 minutes
		STAMP_LIST = 3             // frames pushed one by one by a node in the reference's style: every frame's own pts is kept
	};

	// Walks a stream's frames and returns what `frame.pts * av_q2d(frame.time_base)` is in the reference for each of them
	// (plain double arithmetic, in the reference's order of operations).
	class Frame_clock
	{
		int stamp; double origin; double rate; double t; int64_t at = 0;
		std::shared_ptr<const std::vector<double>> list; size_t index = 0;
	  public:
		// origin: STAMP_START: pts of the first frame; STAMP_END_US: the node's time_seconds before its first frame (0);
		// STAMP_START_FLOAT_US: the node's time_seconds at its first frame (= the first input frame's stamp)
		Frame_clock(int stamp, double origin, int sample_rate, std::shared_ptr<const std::vector<double>> list = nullptr)
			: stamp(stamp), origin(origin), rate((double)sample_rate), t(origin), list(std::move(list)) {}
		double next(int64_t nb)
		{
			double pts;
			switch (stamp)
			{
			case STAMP_END_US:
				t += (double)nb / rate;
				pts = (double)(int64_t)(t * 1000000) * (1 / (double)1000000);
				break;
			case STAMP_START_FLOAT_US:
				pts = (double)(int64_t)(float)(t * 1000000) * (1 / (double)1000000);
				t += (double)nb / rate;
				break;
			case STAMP_LIST:
				if (list && index < list->size()) { pts = (*list)[index++]; break; }
				[[fallthrough]];
			default:
				pts = origin + (double)at / rate;
				break;
			}
			at += nb;
			return pts;
		}
	};

	// A producer that renders a stream in chunks along time says so here: once points[k].event has completed, frames
	// [0, points[k].frames) hold their final values (ascending; the last point covers the whole stream).  A consumer that
	// can start on a prefix waits for the point it needs instead of Audio_buffer::ready -- the reference's nodes overlap the
	// same way, frame by frame through their 16-frame channels (src/processor/audio-stream.cpp:60-80).
	struct Stream_progress
	{
		struct Point { int64_t frames; std::shared_ptr<infra::Device_event> event; };
		std::vector<Point> points;
		infra::Stream_handle stream = nullptr;     // the stream the producer enqueued on
	};

	struct Audio_buffer;

	// audio_volume_adjust on a float stream does not run its pass at once: it publishes the product LAZY -- "source times
	// gain" -- so that a mixer behind it can fold the multiplication into its own read ((x * gain) rounded, then * volume:
	// the bits the node's own pass would have stored, audio-vol.cpp:75-100) and the intermediate stream never exists.
	// Every other reader gets the ordinary buffer: materialize() runs the gain kernel on first use.
	struct Lazy_gain
	{
		std::mutex mutex;
		std::shared_ptr<const Audio_buffer> source;
		float gain = 1.0f;
		bool done = false;                             // materialize() has filled block / plane / ready
	};

	struct Audio_buffer
	{
		// (mutable: filled by materialize() for a lazy product, under Lazy_gain::mutex)
		mutable std::shared_ptr<infra::Device_block> block;    // owner of the memory (may be shared by a whole batch)
		mutable void* plane[2] = {nullptr, nullptr};   // packed: plane[0]; planar: one plane per channel
		int format = FMT_FLT;
		int sample_rate = 48000;
		int channels = 2;
		int64_t frames = 0;                            // samples per channel
		Frame_runs runs;                               // how the reference would have cut it into frames
		double pts_seconds = 0.0;                      // stamp of the first frame as a consumer reads it
		int stamp = STAMP_START;                       // how the producer stamped its frames (Pts_stamp)
		double stamp_origin = 0.0;                     // Frame_clock origin for the END_US / START_FLOAT_US rules
		std::shared_ptr<const std::vector<double>> frame_pts;      // STAMP_LIST: one stamp per frame
		Frame_clock clock() const
		{
			const bool from_first = stamp == STAMP_START || stamp == STAMP_LIST;
			return Frame_clock(stamp, from_first ? pts_seconds : stamp_origin, sample_rate, frame_pts);
		}
		mutable std::shared_ptr<infra::Device_event> ready;    // recorded after the producing kernels were enqueued
		std::shared_ptr<const Stream_progress> progress;   // optional: chunk-wise completion (shared by the products of a batch)
		std::shared_ptr<Lazy_gain> lazy;               // optional: see Lazy_gain

		bool is_lazy() const { return lazy && !lazy->done; }

		size_t plane_bytes() const { return (size_t)frames * (size_t)format_bytes(format) * (format_is_planar(format) ? 1u : (size_t)channels); }
	};

	// runs the deferred gain pass of a lazy product on `stream` (no-op otherwise); afterwards plane[] / ready are valid
	void materialize(const Audio_buffer& buffer, infra::Stream_handle stream);

	// Host-side frame of the frame-streaming compatibility mode: the AVFrame subset the reference's nodes touch
	// (include/processor/audio-stream.hpp:22-42: format, sample_rate, channel count, nb_samples, pts, data planes).
	struct Audio_frame
	{
		int format = FMT_FLT;
		int sample_rate = 48000;
		int channels = 2;
		int64_t nb_samples = 0;                        // samples per channel
		double pts_seconds = 0.0;
		std::vector<uint8_t> data[2];                  // packed: data[0]; planar: one plane per channel

		uint8_t* plane(int c) { return data[format_is_planar(format) ? c : 0].data(); }
		const uint8_t* plane(int c) const { return data[format_is_planar(format) ? c : 0].data(); }
	};

	class Audio_stream : public infra::Processor::Product
	{
		mutable std::mutex mutex;
		std::shared_ptr<const Audio_buffer> buffer;
		std::atomic<bool> end_of_stream = false;

		// frame-streaming compatibility mode: a consumer written against the reference's
		// try_pop() sees the published device buffer cut into its frames (downloaded once, on first use); a
		// producer written against try_push() has its frames collected and uploaded as one buffer at set_eof()
		struct Frame_cursor { std::vector<uint8_t> host[2]; size_t run = 0; int64_t left = 0, done = 0; bool loaded = false; std::unique_ptr<Frame_clock> clock; };
		std::unique_ptr<Frame_cursor> cursor;
		std::vector<std::shared_ptr<const Audio_frame>> pushed;
		void upload_pushed();

	  public:

		Audio_stream() = default;
		Audio_stream(const Audio_stream&) = delete;
		Audio_stream& operator=(const Audio_stream&) = delete;

		// Runner: every consumer of this link has enqueued its work; the link lets go of the buffer (the memory is freed
		// in stream order once no other link holds it)
		void release() override
		{
			std::lock_guard lock(mutex);
			buffer.reset();
		}

		// producer side: hand over the rendered track and close the stream
		void publish(std::shared_ptr<const Audio_buffer> rendered)
		{
			{
				std::lock_guard lock(mutex);
				buffer = std::move(rendered);
			}
			end_of_stream.store(true);
		}

		// consumer side: nullptr when the producer closed the stream without data
		std::shared_ptr<const Audio_buffer> get() const
		{
			std::lock_guard lock(mutex);
			return buffer;
		}

		bool eof() const { return end_of_stream.load(); }
		// closes the stream; frames collected by try_push() are uploaded and published first
		void set_eof()
		{
			if (!pushed.empty()) upload_pushed();
			end_of_stream.store(true);
		}
		size_t buffered_count() const { return get() ? 1 : 0; }

		// ---- the reference's frame interface (src/processor/audio-stream.cpp:60-80) ----
		// try_push: always accepts (no 16-frame bound: nothing drains the stream concurrently); all frames of a
		// stream must share format, rate and channel count
		bool try_push(std::shared_ptr<const Audio_frame> frame);
		// try_pop: next frame of the rendered stream, in the frame sizes the producing node recorded; nullopt when
		// the stream is drained (check eof()) or nothing was published
		std::optional<std::shared_ptr<const Audio_frame>> try_pop();
	};

	// product of audio_spectrum: complex64 bins, [channels][frames][fft_size/2+1], device resident
	struct Spectrum_buffer
	{
		std::shared_ptr<infra::Device_block> block;
		float* data = nullptr;
		int channels = 0, fft_size = 4096, hop = 1024, sample_rate = 48000;
		int64_t frames = 0;
		std::shared_ptr<infra::Device_event> ready;
	};

	class Spectrum_stream : public infra::Processor::Product
	{
		mutable std::mutex mutex;
		std::shared_ptr<const Spectrum_buffer> buffer;

	  public:
		void publish(std::shared_ptr<const Spectrum_buffer> b) { std::lock_guard lock(mutex); buffer = std::move(b); }
		std::shared_ptr<const Spectrum_buffer> get() const { std::lock_guard lock(mutex); return buffer; }
	};
}
