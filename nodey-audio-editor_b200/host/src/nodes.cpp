// nodes.cpp -- the processor nodes over the C ABI (include/nodey_cuda.h).  A node only plans (pure host
// arithmetic on lengths and frame sizes) and enqueues kernels on the stream the Runner gave it; the
// published product carries the event consumers order themselves after.  Reference per node: see
// include/processor/nodes.hpp.
#include "processor/nodes.hpp"

#include "nodey_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <mutex>
#include <tuple>

namespace processor
{
	using infra::Exec_context;
	using infra::Processor;
	using Runtime_error = infra::Processor::Runtime_error;

	// ---------------------------------------------------------------------------------------------
	// helpers
	// ---------------------------------------------------------------------------------------------
	Frame_runs uniform_frame_runs(int64_t frames, int64_t frame_size)
	{
		Frame_runs runs;
		if (frames <= 0 || frame_size <= 0) return runs;
		if (frames / frame_size) runs.emplace_back(frame_size, frames / frame_size);
		if (frames % frame_size) runs.emplace_back(frames % frame_size, 1);
		return runs;
	}

	int64_t frame_runs_total(const Frame_runs& runs)
	{
		int64_t n = 0;
		for (const auto& [len, count] : runs) n += len * count;
		return n;
	}

	namespace
	{
		nodey_stream_t cur_stream() { return Exec_context::current().stream; }

		void abi(int rc, const char* node)
		{
			if (rc == NODEY_OK) return;
			const std::string text = nodey_last_error();
			if (rc == NODEY_E_NOMEM) throw std::bad_alloc();
			if (rc == NODEY_E_FORMAT) throw Runtime_error("Unsupported sample format", std::format("{} cannot process this sample format.", node), text);
			if (rc == NODEY_E_RANGE) throw Runtime_error("Parameter out of range", std::format("{} was given a value the reference rejects too.", node), text);
			throw Runtime_error("GPU kernel call failed", std::format("{} could not enqueue its work.", node), text);
		}

		struct Pin_type
		{
			static Processor::Pin_attribute audio(std::string id, std::string name, bool is_input)
			{
				return {.identifier = std::move(id), .display_name = std::move(name), .type = typeid(Audio_stream), .is_input = is_input,
						.generate_func = [] { return std::make_shared<Audio_stream>(); }};
			}
		};

		// sub-allocator over one Device_block: every product of a batch is a view into the same block
		struct Arena
		{
			std::shared_ptr<infra::Device_block> block;
			size_t used = 0;
			explicit Arena(size_t bytes) : block(std::make_shared<infra::Device_block>(bytes + 256)) {}
			static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
			void* take(size_t bytes)
			{
				void* p = (char*)block->ptr + used;
				used += padded(bytes);
				if (used > block->bytes) THROW_LOGIC_ERROR("arena overflow: {} of {}", used, block->bytes);
				return p;
			}
		};

		// wait = false: the caller orders itself after the producer (chunk by chunk through Audio_buffer::progress)
		// accept_lazy = false: a lazy gain product (Lazy_gain) is materialised here, so the caller sees an ordinary buffer
		// The "no input" error reads as the reference's does, word for word (audio-vol.cpp:113-117, audio-amix.cpp:122-126,
		// audio-bimix.cpp:110-114 and 486-490, audio-io.cpp:858-862, audio-velocity.cpp:278-282): `explain_title` and `detail_key`
		// carry the two places where the reference's own strings differ from the pattern (bimix writes "Audio channel mix
		// processor" in the explanation and names the pin 'input' whichever side is missing).
		std::shared_ptr<const Audio_buffer> require_input(const Processor::Input_map& input, const std::string& key,
														   const char* node_title, bool wait = true, bool accept_lazy = false,
														   const char* explain_title = nullptr, const char* detail_key = nullptr)
		{
			const auto item = infra::get_input_item<Audio_stream>(input, key);
			if (!item.has_value())
				throw Runtime_error(std::format("{} has no input", node_title),
									std::format("{} requires an audio stream input to function properly.", explain_title ? explain_title : node_title),
									std::format("Input item '{}' not found", detail_key ? std::string(detail_key) : key));
			auto buffer = item->get().get();
			if (!buffer)
				throw Runtime_error(std::format("{} received an empty stream", node_title),
									"The upstream node closed its stream without publishing audio.", std::format("pin '{}'", key));
			if (!accept_lazy && buffer->is_lazy()) materialize(*buffer, cur_stream());
			if (wait && buffer->ready) buffer->ready->wait_on(cur_stream());
			return buffer;
		}

		void publish(const Processor::Output_map& output, const std::string& key, const std::shared_ptr<Audio_buffer>& buffer)
		{
			if (!buffer->ready)
			{
				auto ev = std::make_shared<infra::Device_event>();
				ev->record(cur_stream());
				buffer->ready = std::move(ev);
			}
			for (auto& stream : infra::get_output_item<Audio_stream>(output, key)) stream->publish(buffer);
		}

		// the stream after `s` in the lane's cycle  stream -> side[0] -> ... -> side[3] -> stream  (a stream outside the
		// lane, e.g. the transfer lane's, is followed by the lane's own stream)
		nodey_stream_t next_in_cycle(nodey_stream_t s)
		{
			const Exec_context& ctx = Exec_context::current();
			constexpr int n = Exec_context::kSideStreams;
			if (s == ctx.stream) return ctx.side_stream[0] ? ctx.side_stream[0] : ctx.stream;
			for (int k = 0; k < n; k++)
				if (s == ctx.side_stream[k]) return k + 1 < n && ctx.side_stream[k + 1] ? ctx.side_stream[k + 1] : ctx.stream;
			return ctx.stream;
		}

		// the stream a chunk-wise consumer runs on: the one after its producer's (no chunk-wise producer: the lane's own)
		nodey_stream_t stream_after(const Stream_progress* producer)
		{
			return producer ? next_in_cycle(producer->stream) : Exec_context::current().stream;
		}

		// orders a stream after the progress points of a chunk-wise input, one point at a time
		struct Progress_waiter
		{
			const Stream_progress* progress;
			size_t waited = 0;
			void need(int64_t frames, nodey_stream_t stream)
			{
				if (!progress) return;
				while (waited < progress->points.size() && (waited == 0 || progress->points[waited - 1].frames < frames))
					progress->points[waited++].event->wait_on(stream);
			}
			void all(nodey_stream_t stream)
			{
				if (!progress) return;
				while (waited < progress->points.size()) progress->points[waited++].event->wait_on(stream);
			}
		};

		// launches of the current node go to `s` (and its allocations follow that stream's order) while this lives
		struct Stream_scope
		{
			nodey_stream_t saved;
			explicit Stream_scope(nodey_stream_t s) : saved(Exec_context::current().stream) { Exec_context::current().stream = s; }
			~Stream_scope() { Exec_context::current().stream = saved; }
			Stream_scope(const Stream_scope&) = delete;
			Stream_scope& operator=(const Stream_scope&) = delete;
		};

		// how many launches / copies a stream is cut into along time (1 = whole-track); Runner::Schedule::stream_chunks
		// overrides (bench.py times kernels one at a time with 1).  24: the 256 x 180 s render end to end 317 ms with 16 chunks,
		// 311 ms with 24, 32, 48 or 64 (what is exposed after the last uploaded byte is one chunk of every stage); with the
		// sources resident 16..64 chunks are within 1 ms of each other (191 ms, 33 ms at 32 tracks)
		int stream_chunk_count()
		{
			const int set = Exec_context::current().stream_chunks;
			return set > 0 ? std::clamp(set, 1, 64) : 24;
		}

		std::shared_ptr<Audio_buffer> new_buffer(const std::shared_ptr<infra::Device_block>& block, void* p0, void* p1, int fmt, int rate,
												 int ch, int64_t frames, Frame_runs runs, double pts)
		{
			auto b = std::make_shared<Audio_buffer>();
			b->block = block; b->plane[0] = p0; b->plane[1] = p1; b->format = fmt; b->sample_rate = rate; b->channels = ch;
			b->frames = frames; b->runs = std::move(runs); b->pts_seconds = pts;
			return b;
		}

		// frames forwarded with their stamps (gain, split: `out_frame->pts = src_frame.pts`, audio-vol.cpp:170-171)
		std::shared_ptr<Audio_buffer> inherit_stamps(std::shared_ptr<Audio_buffer> b, const Audio_buffer& in)
		{
			b->stamp = in.stamp; b->stamp_origin = in.stamp_origin; b->frame_pts = in.frame_pts;
			return b;
		}

		// audio_amix / audio_bimix: running END time from 0 in whole microseconds (App. C3 / C4); pts_seconds = the first stamp
		std::shared_ptr<Audio_buffer> end_time_stamps(std::shared_ptr<Audio_buffer> b)
		{
			b->stamp = STAMP_END_US; b->stamp_origin = 0.0;
			b->pts_seconds = 0.0;
			if (!b->runs.empty() && b->frames > 0) b->pts_seconds = b->clock().next(std::min<int64_t>(b->runs.front().first, b->frames));
			return b;
		}

		// libavutil's names of the sample formats (av_get_sample_fmt_name), for the reference's "Sample format: {}" detail
		const char* sample_format_name(int format)
		{
			static const char* const names[] = {"u8", "s16", "s32", "flt", "dbl", "u8p", "s16p", "s32p", "fltp", "dblp", "s64", "s64p"};
			return format >= 0 && format < 12 ? names[format] : "(null)";
		}

		// Streams a node cannot take.  The three nodes whose reference code checks by itself fail with its words:
		// audio_volume_adjust (audio-vol.cpp:177-182, 238-243), the SoundTouch nodes (extract_samples_interleaved,
		// audio-velocity.cpp:223-228) and audio_bimix_v2 (audio-bimix.cpp:555-560); the others leave it to swr_init.
		enum class Node_kind { other, volume, soundtouch, bimix_v2 };

		void check_channels(const Audio_buffer& b, const char* node, Node_kind kind = Node_kind::other)
		{
			if (b.channels != 1 && b.channels != 2)
			{
				if (kind == Node_kind::volume)
					throw Runtime_error("Invalid channel count", "Only mono and stereo audio are supported.", std::format("Got {} channels", b.channels));
				if (kind == Node_kind::bimix_v2)
					throw Runtime_error("Invalid audio channel layout", "Audio channel layout must be stereo or mono.",
										std::format("Invalid channel layout: {}", b.channels));
				throw Runtime_error("Invalid channel layout", std::format("{} handles mono and stereo streams.", node), std::format("channels: {}", b.channels));
			}
			if (format_bytes(b.format) == 0)
			{
				if (kind == Node_kind::volume)
					throw Runtime_error("Audio format is not support", "Audio volume processor requires an audio format properly.", "Include FLT, S16, S32");
				if (kind == Node_kind::soundtouch)
					throw Runtime_error("Unsupported sample format", "The processors do not support the given sample format.",
										std::format("Sample format: {}", sample_format_name(b.format)));
				throw Runtime_error("Audio format is not support (Include FLT, S16, S32)", std::format("{} cannot process this sample format.", node),
									std::format("AVSampleFormat {}", b.format));
			}
		}

		// process-wide plan caches (plans are immutable once built; kernels in flight keep using them)
		std::mutex plan_mutex;

		// plans own device memory (filter banks, fade ramps, position tables) on the device that was current when they
		// were built, so both caches are keyed by the device ordinal: a process may run one Runner per GPU
		int current_device()
		{
			int device = 0;
			abi(nodey_get_device(&device), "plan cache");
			return device;
		}

		nodey_resampler* resampler_for(int in_rate)
		{
			static std::map<std::pair<int, int>, nodey_resampler*> cache;
			const int device = current_device();
			std::lock_guard lock(plan_mutex);
			const auto it = cache.find({device, in_rate});
			if (it != cache.end()) return it->second;
			nodey_resampler* r = nullptr;
			abi(nodey_resampler_create(&r, in_rate, 48000, 0), "swresample plan");
			cache[{device, in_rate}] = r;
			return r;
		}

		// SoundTouch plans by (device, rate, channels, parameters): a plan holds one immutable position table per stream
		// length it has rendered, so the length is not part of the key.  Bounded: the least recently used plan goes when
		// more than kMaxSoundtouchPlans exist -- shared_ptr, so a render that is still enqueueing with it keeps it alive
		// (kernels already enqueued only need the device tables, and nodey_soundtouch_destroy frees them with cudaFree,
		// which waits for the device).
		constexpr size_t kMaxSoundtouchPlans = 32;

		std::shared_ptr<nodey_soundtouch> soundtouch_for(int rate_hz, int ch, float rate, float pitch)
		{
			using Key = std::tuple<int, int, int, uint32_t, uint32_t>;
			struct Slot { std::shared_ptr<nodey_soundtouch> plan; uint64_t used; };
			static std::map<Key, Slot> cache;
			static uint64_t clock = 0;
			uint32_t rb, pb;
			memcpy(&rb, &rate, 4); memcpy(&pb, &pitch, 4);
			const int device = current_device();
			std::lock_guard lock(plan_mutex);
			const Key key{device, rate_hz, ch, rb, pb};
			const auto it = cache.find(key);
			if (it != cache.end()) { it->second.used = ++clock; return it->second.plan; }
			nodey_soundtouch* s = nullptr;
			const int rc = nodey_soundtouch_create(&s, rate_hz, ch, rate, pitch);
			if (rc == NODEY_E_RANGE)
				// audio-velocity.cpp:371-379, word for word (the explanation names the rate where a node name was meant)
				throw Runtime_error("Unsupported sample rate", std::format("{} requires a sample rate between 8000 and 48000 Hz.", rate_hz),
									std::format("Sample rate: {}", rate_hz));
			abi(rc, "SoundTouch");
			if (cache.size() >= kMaxSoundtouchPlans)
			{
				auto oldest = cache.begin();
				for (auto k = cache.begin(); k != cache.end(); ++k)
					if (k->second.used < oldest->second.used) oldest = k;
				cache.erase(oldest);
			}
			std::shared_ptr<nodey_soundtouch> plan(s, [](nodey_soundtouch* p) { nodey_soundtouch_destroy(p); });
			cache[key] = Slot{plan, ++clock};
			return plan;
		}

		struct Runs_flat
		{
			std::vector<int64_t> off, len, count;
			explicit Runs_flat(const std::vector<const Frame_runs*>& all)
			{
				off.push_back(0);
				for (const Frame_runs* runs : all)
				{
					for (const auto& [l, c] : *runs) { len.push_back(l); count.push_back(c); }
					off.push_back((int64_t)len.size());
				}
				if (len.empty()) { len.push_back(0); count.push_back(0); }
			}
		};
	}

	// ---------------------------------------------------------------------------------------------
	// lazy gain products (audio-stream.hpp: Lazy_gain)
	// ---------------------------------------------------------------------------------------------
	void materialize(const Audio_buffer& b, infra::Stream_handle stream)
	{
		if (!b.lazy) return;
		std::lock_guard lock(b.lazy->mutex);
		if (b.lazy->done) return;
		const Audio_buffer& src = *b.lazy->source;
		if (src.is_lazy()) materialize(src, stream);
		Stream_scope scope(stream);
		if (src.ready) src.ready->wait_on(stream);
		const bool planar2 = format_is_planar(src.format) && src.channels == 2;
		const size_t plane = Arena::padded(src.plane_bytes());
		auto block = std::make_shared<infra::Device_block>(std::max<size_t>(plane * (planar2 ? 2 : 1), 256));
		void* p0 = block->ptr;
		void* p1 = planar2 ? (char*)block->ptr + plane : nullptr;
		const int64_t n = format_is_planar(src.format) ? src.frames : src.frames * src.channels;
		abi(nodey_gain(p0, src.plane[0], src.format, n, b.lazy->gain, stream), "Volume adjust");
		if (planar2) abi(nodey_gain(p1, src.plane[1], src.format, n, b.lazy->gain, stream), "Volume adjust");
		auto ev = std::make_shared<infra::Device_event>();
		ev->record(stream);
		b.block = std::move(block); b.plane[0] = p0; b.plane[1] = p1; b.ready = std::move(ev);
		b.lazy->done = true;
	}

	// ---------------------------------------------------------------------------------------------
	// frame-streaming compatibility mode of Audio_stream (reference: src/processor/audio-stream.cpp:60-80)
	// ---------------------------------------------------------------------------------------------
	bool Audio_stream::try_push(std::shared_ptr<const Audio_frame> frame)
	{
		if (!frame || end_of_stream.load()) return false;
		if (format_bytes(frame->format) == 0 || (frame->channels != 1 && frame->channels != 2) || frame->nb_samples < 0)
			throw Runtime_error("Invalid audio frame", "A processor pushed a frame this engine cannot carry.",
								std::format("format {}, channels {}, samples {}", frame->format, frame->channels, frame->nb_samples));
		if (!pushed.empty())
		{
			const Audio_frame& first = *pushed.front();
			if (first.format != frame->format || first.sample_rate != frame->sample_rate || first.channels != frame->channels)
				throw Runtime_error("Inconsistent audio frames", "All frames of one stream must share format, sample rate and channel count.",
									std::format("format {} vs {}, rate {} vs {}", first.format, frame->format, first.sample_rate, frame->sample_rate));
		}
		pushed.push_back(std::move(frame));
		return true;
	}

	// the collected frames become one device-resident buffer, with their sizes as the stream's frame runs
	void Audio_stream::upload_pushed()
	{
		const Audio_frame& first = *pushed.front();
		const bool planar2 = format_is_planar(first.format) && first.channels == 2;
		const size_t sample = (size_t)format_bytes(first.format) * (format_is_planar(first.format) ? 1u : (size_t)first.channels);
		int64_t total = 0;
		Frame_runs runs;
		for (const auto& f : pushed)
		{
			if (f->nb_samples == 0) continue;
			total += f->nb_samples;
			if (!runs.empty() && runs.back().first == f->nb_samples) runs.back().second++;
			else runs.emplace_back(f->nb_samples, 1);
		}
		const size_t plane = Arena::padded((size_t)std::max<int64_t>(total, 1) * sample);
		auto block = std::make_shared<infra::Device_block>(plane * (planar2 ? 2 : 1));
		// one contiguous host image per plane (kept alive by the buffer's deleter until the copy has landed: the
		// stream is synchronised here, uploads of this mode are not performance paths)
		for (int c = 0; c < (planar2 ? 2 : 1); c++)
		{
			std::vector<uint8_t> host((size_t)total * sample);
			size_t at = 0;
			for (const auto& f : pushed)
			{
				const size_t bytes = (size_t)f->nb_samples * sample;
				if (f->data[c].size() < bytes)
					throw Runtime_error("Invalid audio frame", "A pushed frame holds fewer samples than it declares.", std::format("{} < {}", f->data[c].size(), bytes));
				memcpy(host.data() + at, f->data[c].data(), bytes);
				at += bytes;
			}
			if (total > 0)
			{
				abi(nodey_memcpy_h2d((char*)block->ptr + c * plane, host.data(), host.size(), cur_stream()), "frame upload");
				abi(nodey_stream_synchronize(cur_stream()), "frame upload");
			}
		}
		auto b = new_buffer(block, block->ptr, planar2 ? (char*)block->ptr + plane : nullptr, first.format, first.sample_rate, first.channels,
							total, std::move(runs), first.pts_seconds);
		{
			// every pushed frame keeps its own stamp (a node in the reference's style forwards or makes them as it likes)
			auto stamps = std::make_shared<std::vector<double>>();
			for (const auto& f : pushed)
				if (f->nb_samples != 0) stamps->push_back(f->pts_seconds);
			if (!stamps->empty()) b->pts_seconds = stamps->front();
			b->stamp = STAMP_LIST; b->frame_pts = std::move(stamps);
		}
		auto ev = std::make_shared<infra::Device_event>();
		ev->record(cur_stream());
		b->ready = std::move(ev);
		pushed.clear();
		{
			std::lock_guard lock(mutex);
			buffer = std::move(b);
		}
	}

	std::optional<std::shared_ptr<const Audio_frame>> Audio_stream::try_pop()
	{
		const auto b = get();
		if (!b) return std::nullopt;
		if (!cursor) cursor = std::make_unique<Frame_cursor>();
		Frame_cursor& cur = *cursor;
		const bool planar2 = format_is_planar(b->format) && b->channels == 2;
		const size_t sample = (size_t)format_bytes(b->format) * (format_is_planar(b->format) ? 1u : (size_t)b->channels);
		if (!cur.loaded)
		{
			// download once; the frames below are cut from this image
			if (b->is_lazy()) materialize(*b, cur_stream());
			if (b->ready) b->ready->wait_on(cur_stream());
			for (int c = 0; c < (planar2 ? 2 : 1); c++)
			{
				cur.host[c].resize((size_t)b->frames * sample);
				if (b->frames > 0) abi(nodey_memcpy_d2h(cur.host[c].data(), b->plane[c], cur.host[c].size(), cur_stream()), "frame download");
			}
			abi(nodey_stream_synchronize(cur_stream()), "frame download");
			cur.loaded = true;
			cur.clock = std::make_unique<Frame_clock>(b->clock());
			cur.run = 0;
			cur.left = b->runs.empty() ? 0 : b->runs[0].second;
		}
		while (cur.run < b->runs.size() && (cur.left <= 0 || b->runs[cur.run].first <= 0))
		{
			cur.run++;
			cur.left = cur.run < b->runs.size() ? b->runs[cur.run].second : 0;
		}
		if (cur.run >= b->runs.size() || cur.done >= b->frames) return std::nullopt;
		const int64_t n = std::min<int64_t>(b->runs[cur.run].first, b->frames - cur.done);
		auto f = std::make_shared<Audio_frame>();
		f->format = b->format; f->sample_rate = b->sample_rate; f->channels = b->channels; f->nb_samples = n;
		f->pts_seconds = cur.clock->next(n);           // the frame's own stamp, by the producer's rule
		for (int c = 0; c < (planar2 ? 2 : 1); c++)
			f->data[c].assign(cur.host[c].begin() + (size_t)cur.done * sample, cur.host[c].begin() + (size_t)(cur.done + n) * sample);
		cur.done += n;
		cur.left--;
		return std::shared_ptr<const Audio_frame>(std::move(f));
	}

	// ---------------------------------------------------------------------------------------------
	// frame_gain_example: a node in the reference's own style (pop a frame, work on the host, push it on)
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Frame_gain_example::get_processor_info()
	{
		return {.identifier = "frame_gain_example", .display_name = "Frame Gain (example)", .singleton = false,
				.generate = std::make_unique<Frame_gain_example>,
				.description = "Example of a processor written against the frame interface (try_pop / try_push): scales every frame on the host."};
	}
	std::vector<Processor::Pin_attribute> Frame_gain_example::get_pin_attributes() const
	{
		return {Pin_type::audio("output", "Output", false), Pin_type::audio("input", "Input", true)};
	}
	Json::Value Frame_gain_example::serialize() const
	{
		Json::Value value;
		value["volume"] = volume;
		return value;
	}
	void Frame_gain_example::deserialize(const Json::Value& value)
	{
		if (value.isObject() && value.isMember("volume") && value["volume"].isDouble()) volume = value["volume"].asFloat();
	}
	void Frame_gain_example::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>& stop_token, std::any&)
	{
		const auto in = infra::get_input_item<Audio_stream>(input, "input");
		if (!in.has_value())
			throw Runtime_error("Frame gain has no input", "The example node requires an audio stream on pin 'input'.", "Input item 'input' not found");
		Audio_stream& stream = in->get();
		const auto outs = infra::get_output_item<Audio_stream>(output, "output");
		// the loop shape of the reference's nodes (audio-vol.cpp:102-250): pop, process, push to every link, until EOF
		while (!stop_token)
		{
			const auto popped = stream.try_pop();
			if (!popped.has_value())
			{
				if (stream.eof()) break;
				continue;       // the reference yields its fiber here; in this engine an unpublished input means an upstream error
			}
			const Audio_frame& src = **popped;
			auto dst = std::make_shared<Audio_frame>(src);
			const size_t count = (size_t)src.nb_samples * (format_is_planar(src.format) ? 1u : (size_t)src.channels);
			for (int c = 0; c < (format_is_planar(src.format) ? src.channels : 1); c++)
			{
				// change_volume<T>: dst[i] = T(src[i] * volume), float multiply, C++ truncating conversion (audio-vol.cpp:75-100)
				switch (src.format)
				{
				case FMT_FLT: case FMT_FLTP: { auto* p = (float*)dst->data[c].data(); for (size_t i = 0; i < count; i++) p[i] = p[i] * volume; break; }
				case FMT_S16: case FMT_S16P: { auto* p = (int16_t*)dst->data[c].data(); for (size_t i = 0; i < count; i++) p[i] = (int16_t)(int32_t)((float)p[i] * volume); break; }
				case FMT_S32: case FMT_S32P: { auto* p = (int32_t*)dst->data[c].data(); for (size_t i = 0; i < count; i++) { const float v = (float)p[i] * volume; p[i] = (v >= -2147483648.0f && v < 2147483648.0f) ? (int32_t)v : INT32_MIN; } break; }
				default: throw Runtime_error("Audio format is not support (Include FLT, S16, S32)", "Frame gain cannot process this sample format.", std::format("AVSampleFormat {}", src.format));
				}
			}
			for (auto& o : outs) o->try_push(dst);
		}
		for (auto& o : outs) o->set_eof();
	}

	void register_example_processors()
	{
		static std::once_flag once;
		std::call_once(once, [] { Processor::register_processor<Frame_gain_example>(); });
	}

	// ---------------------------------------------------------------------------------------------
	// audio_input
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_input::get_processor_info()
	{
		return {.identifier = "audio_input", .display_name = "Audio Input", .singleton = true, .generate = std::make_unique<Audio_input>,
				.description = "PCM sources of the render: one output pin per file slot. Sources are raw PCM buffers handed in by the host "
							   "(Pcm_source_list) or RIFF/WAVE files named by file_path."};
	}

	std::vector<Processor::Pin_attribute> Audio_input::get_pin_attributes() const
	{
		std::vector<Processor::Pin_attribute> pins;
		for (size_t i = 0; i < file_count; i++) pins.push_back(Pin_type::audio(std::format("output_{}", i), std::format("Output {}", i + 1), false));
		return pins;
	}

	void Audio_input::set_file_count(size_t n)
	{
		file_count = std::max<size_t>(n, 1);
		file_paths.resize(file_count);
	}

	Json::Value Audio_input::serialize() const
	{
		Json::Value list(Json::arrayValue);
		for (const auto& path : file_paths) list.append(path);
		Json::Value value(Json::objectValue);
		value["file_path"] = list;
		return value;
	}

	void Audio_input::deserialize(const Json::Value& value)
	{
		const auto bad = [](const char* field) {
			return Runtime_error("Failed to deserialize JSON file",
								 "Audio_input failed to serialize the JSON input because of missing or invalid fields.", std::format("Wrong field: {}", field));
		};
		if (!value.isObject() || !value.isMember("file_path") || !value["file_path"].isArray()) throw bad("file_path");
		file_paths.clear();
		for (const auto& path : value["file_path"])
		{
			if (!path.isString()) throw bad("file_path.path");
			file_paths.push_back(path.asString());
		}
		file_count = std::max<size_t>(file_paths.size(), 1);
		file_paths.resize(file_count);
	}

	namespace
	{
		// canonical RIFF/WAVE reader: PCM 16 / 24 / 32 bit and IEEE float 32 -- what the reference's libavformat / libavcodec
		// path (audio-io.cpp:69-227) hands to the nodes for such files:
		//  * sample formats: pcm_s16le -> S16, pcm_s32le -> S32, pcm_f32le -> FLT, pcm_s24le -> S32 with the 24 bits in the
		//    upper three bytes (libavcodec/pcm.c decodes le24 with a shift of 8); u8 and f64 files decode to formats every
		//    node of the reference refuses ("Audio format is not support"), here they are refused at the source;
		//  * frame sizes: the wav demuxer cuts the data chunk into packets of max_size = 4096 bytes rounded down to whole
		//    blocks (libavformat/wavdec.c, wav_read_packet) and the PCM decoder returns one frame per packet, so a frame
		//    holds 4096 / block_align sample frames: 1024 for 16-bit stereo, 2048 for 16-bit mono, 512 for 32-bit or float
		//    stereo, 682 for 24-bit stereo; the last frame takes the rest.  audio_amix's `nb` follows its inputs' frame sizes
		//    (audio-amix.cpp:190-196), so the size is part of the result.
		struct Wav_data { std::vector<char> samples; int format = FMT_S16, rate = 0, channels = 0, frame_size = 1024; int64_t frames = 0; };

		int wav_packet_frames(int file_bytes_per_sample, int channels)
		{
			const int block = file_bytes_per_sample * channels;
			int size = 4096;
			if (block > 1)
			{
				if (size < block) size = block;
				size = (size / block) * block;
			}
			return std::max(size / block, 1);
		}

		Wav_data read_wav(const std::string& path, bool header_only = false)
		{
			const auto fail = [&](const std::string& why) {
				return Runtime_error("Cannot open audio file", std::format("'{}' is not a PCM/float RIFF WAVE file this engine can read.", path), why);
			};
			std::ifstream f(path, std::ios::binary);
			if (!f) throw fail("file not found or unreadable");
			char hdr[12];
			if (!f.read(hdr, 12) || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) throw fail("missing RIFF/WAVE header");
			Wav_data w;
			int tag = 0, bits = 0;
			bool have_fmt = false, have_data = false;
			uint64_t data_bytes = 0;
			for (;;)
			{
				char ck[8];
				if (!f.read(ck, 8)) break;
				uint32_t size; memcpy(&size, ck + 4, 4);
				if (!memcmp(ck, "fmt ", 4))
				{
					if (size < 16 || size > 4096) throw fail("fmt chunk of an impossible size");
					std::vector<char> b(size);
					if (!f.read(b.data(), size)) throw fail("truncated fmt chunk");
					uint16_t t, c, bp; uint32_t r;
					memcpy(&t, b.data(), 2); memcpy(&c, b.data() + 2, 2); memcpy(&r, b.data() + 4, 4); memcpy(&bp, b.data() + 14, 2);
					if (t == 0xFFFE && size >= 26) memcpy(&t, b.data() + 24, 2);   // WAVE_FORMAT_EXTENSIBLE sub-format
					tag = t; w.channels = c; w.rate = (int)r; bits = bp; have_fmt = true;
					if (size & 1) f.seekg(1, std::ios::cur);
				}
				else if (!memcmp(ck, "data", 4))
				{
					if (!have_fmt) throw fail("data chunk before fmt chunk");
					have_data = true;
					// what is really there counts (a recording that was cut off); a size of 0 or 0xFFFFFFFF is a streamed file
					// whose writer never patched the header: the demuxer reads such a chunk to the end of the file
					const auto at = f.tellg();
					f.seekg(0, std::ios::end);
					const uint64_t rest = (uint64_t)(f.tellg() - at);
					f.seekg(at);
					data_bytes = (size == 0 || size == 0xFFFFFFFFu) ? rest : std::min<uint64_t>(size, rest);
					if (!header_only)
					{
						w.samples.resize((size_t)data_bytes);
						f.read(w.samples.data(), (std::streamsize)data_bytes);
						w.samples.resize((size_t)f.gcount());
						data_bytes = w.samples.size();
					}
					break;
				}
				else f.seekg(size + (size & 1), std::ios::cur);
			}
			if (!have_fmt || w.channels < 1 || w.channels > 2) throw fail("unsupported channel count");
			if (!have_data) throw fail("no data chunk");
			int file_bytes = 0;
			if (tag == 1 && bits == 16) { w.format = FMT_S16; file_bytes = 2; }
			else if (tag == 1 && bits == 24) { w.format = FMT_S32; file_bytes = 3; }
			else if (tag == 1 && bits == 32) { w.format = FMT_S32; file_bytes = 4; }
			else if (tag == 3 && bits == 32) { w.format = FMT_FLT; file_bytes = 4; }
			else throw fail(std::format("unsupported encoding (tag {}, {} bits)", tag, bits));
			w.frame_size = wav_packet_frames(file_bytes, w.channels);
			w.frames = (int64_t)(data_bytes / (size_t)(file_bytes * w.channels));
			if (file_bytes == 3 && !header_only)
			{
				// pcm_s24le: the three little-endian bytes become the upper three bytes of a 32-bit sample
				const size_t n = (size_t)w.frames * (size_t)w.channels;
				std::vector<char> wide(n * 4);
				const unsigned char* src = (const unsigned char*)w.samples.data();
				for (size_t i = 0; i < n; i++)
				{
					const uint32_t v = ((uint32_t)src[3 * i] << 8) | ((uint32_t)src[3 * i + 1] << 16) | ((uint32_t)src[3 * i + 2] << 24);
					memcpy(wide.data() + 4 * i, &v, 4);
				}
				w.samples.swap(wide);
			}
			return w;
		}
	}

	void probe_wav(const std::string& path, int& format, int& sample_rate, int& channels, int64_t& frames, int& frame_size)
	{
		const Wav_data w = read_wav(path, true);
		format = w.format; sample_rate = w.rate; channels = w.channels; frames = w.frames; frame_size = w.frame_size;
	}

	size_t Audio_input::upload_bytes(const std::any& user_data) const
	{
		const Pcm_source_list* bound = std::any_cast<Pcm_source_list>(&user_data);
		if (!bound)
			if (const auto* sp = std::any_cast<std::shared_ptr<Pcm_source_list>>(&user_data)) bound = sp->get();
		size_t bytes = 0;
		for (size_t i = 0; i < file_count; i++)
		{
			if (bound && i < bound->sources.size() && bound->sources[i].data)
			{
				const Pcm_source& s = bound->sources[i];
				if (!s.on_device) bytes += (size_t)s.frames * (size_t)format_bytes(s.format) * (size_t)s.channels;
			}
			else if (!file_paths[i].empty()) bytes += 64u << 20;     // a file: size unknown until it is read
		}
		return bytes;
	}

	void Audio_input::process_payload(const Input_map&, const Output_map& output, const std::atomic<bool>&, std::any& user_data)
	{
		const Pcm_source_list* bound = std::any_cast<Pcm_source_list>(&user_data);
		if (!bound)
			if (const auto* sp = std::any_cast<std::shared_ptr<Pcm_source_list>>(&user_data)) bound = sp->get();

		// Every slot that names a file must name a regular file, linked or not (audio-io.cpp:232-239, checked before any
		// file is opened); slots served by a bound PCM source (this engine's extension) or left empty are not files.
		for (size_t i = 0; i < file_count; i++)
		{
			if (bound && i < bound->sources.size() && bound->sources[i].data) continue;
			if (file_paths[i].empty()) continue;
			std::error_code ec;
			if (!std::filesystem::exists(file_paths[i], ec) || !std::filesystem::is_regular_file(file_paths[i], ec))
				throw Runtime_error(std::format("Invalid file path in slot {}", i + 1), "The specified audio file does not exist or is not a regular file.",
									std::format("File path: {}", file_paths[i]));
		}

		std::vector<Wav_data> files;     // keeps file payloads alive until the copies are enqueued... and landed
		std::vector<Pcm_source> sources(file_count);
		for (size_t i = 0; i < file_count; i++)
		{
			const auto find = output.find(std::format("output_{}", i));
			const bool used = find != output.end() && !find->second.empty();
			if (bound && i < bound->sources.size() && bound->sources[i].data) { sources[i] = bound->sources[i]; continue; }
			if (!used) continue;
			if (file_paths[i].empty())
				throw Runtime_error("No input source", std::format("Output {} of the audio input node is linked but has neither a PCM source nor a file.", i + 1),
									std::format("pin output_{}", i));
			files.push_back(read_wav(file_paths[i]));
			const Wav_data& w = files.back();
			sources[i].data = w.samples.data(); sources[i].format = w.format; sources[i].sample_rate = w.rate;
			sources[i].channels = w.channels; sources[i].frames = w.frames; sources[i].frame_size = w.frame_size;
		}

		// one arena for every host source that has to be uploaded
		size_t upload_bytes = 0;
		for (const auto& s : sources)
			if (s.data && !s.on_device)
			{
				// the same arithmetic as the take() below: a planar stereo source is two padded planes
				const bool planar2 = format_is_planar(s.format) && s.channels == 2;
				const size_t plane_bytes = (size_t)s.frames * (size_t)format_bytes(s.format) * (format_is_planar(s.format) ? 1u : (size_t)s.channels);
				upload_bytes += planar2 ? 2 * Arena::padded(plane_bytes) : Arena::padded(plane_bytes);
			}
		std::unique_ptr<Arena> arena;
		if (upload_bytes) arena = std::make_unique<Arena>(upload_bytes);

		// Uploads go wave by wave (the blocks of pins the Runner pipelines the graph over): inside a wave whose sources
		// have one length the copies are interleaved chunk by chunk along time -- chunk c of every track, then chunk c + 1
		// -- and an event after each round tells the consumers which prefix has landed (Audio_buffer::progress), so the
		// wave's resamplers and WSOLA chains start while the rest of its audio is still on the bus.
		std::vector<int> wave_first{0};
		if (const auto* waves = Exec_context::current().wave_begin; waves && !waves->empty()) wave_first = *waves;
		wave_first.push_back((int)file_count);
		constexpr size_t kMinChunkBytes = 1u << 20;
		for (size_t w = 0; w + 1 < wave_first.size(); w++)
		{
			struct Pending { size_t pin; void* p0; void* p1; size_t frame_bytes; bool planar2; std::shared_ptr<Audio_buffer> buffer; };
			std::vector<Pending> uploads;
			for (size_t i = (size_t)std::max(wave_first[w], 0); i < (size_t)wave_first[w + 1] && i < file_count; i++)
			{
				const Pcm_source& s = sources[i];
				const std::string key = std::format("output_{}", i);
				const auto find = output.find(key);
				if (!s.data || find == output.end() || find->second.empty()) continue;
				if (format_bytes(s.format) == 0 || (s.channels != 1 && s.channels != 2))
					throw Runtime_error("Unsupported sample format", "PCM sources must be S16/S32/FLT (packed or planar), mono or stereo.", key);
				const size_t frame_bytes = (size_t)format_bytes(s.format) * (format_is_planar(s.format) ? 1u : (size_t)s.channels);
				const size_t plane_bytes = (size_t)s.frames * frame_bytes;
				void* p0 = const_cast<void*>(s.data);
				void* p1 = const_cast<void*>(s.data1);
				std::shared_ptr<infra::Device_block> owner;
				const bool planar2 = format_is_planar(s.format) && s.channels == 2;
				if (!s.on_device)
				{
					owner = arena->block;
					p0 = arena->take(planar2 ? 2 * Arena::padded(plane_bytes) : plane_bytes);
					if (planar2) p1 = (char*)p0 + Arena::padded(plane_bytes);
				}
				auto buffer = new_buffer(owner, p0, p1, s.format, s.sample_rate, s.channels, s.frames, uniform_frame_runs(s.frames, s.frame_size), s.pts_seconds);
				if (s.on_device) publish(output, key, buffer);
				else uploads.push_back({i, p0, p1, frame_bytes, planar2, std::move(buffer)});
			}
			if (uploads.empty()) continue;
			// chunk-wise only when every upload of the wave has the same length (one progress object serves them all) and
			// a chunk is worth a copy of its own
			int nchunks = stream_chunk_count();
			const int64_t frames = uploads.front().buffer->frames;
			for (const Pending& u : uploads)
				if (u.buffer->frames != frames) nchunks = 1;
			while (nchunks > 1 && (size_t)(frames / nchunks) * uploads.front().frame_bytes < kMinChunkBytes) nchunks /= 2;
			if (nchunks < 1) nchunks = 1;
			// One copy per chunk ROUND when the wave's sources sit at a constant pitch in host memory (rows of one pinned
			// block: what a host that batches its tracks has) and -- always true here -- in the arena: a 2-D copy whose rows
			// are the tracks.  Thousands of separate copies would fill the driver's copy queue and block this thread (the
			// enqueue of everything downstream with it) until half of the audio has gone over the bus.
			bool rows = uploads.size() > 1 && !uploads.front().planar2;
			ptrdiff_t src_pitch = 0, dst_pitch = 0;
			if (rows)
			{
				src_pitch = (const char*)sources[uploads[1].pin].data - (const char*)sources[uploads[0].pin].data;
				dst_pitch = (const char*)uploads[1].p0 - (const char*)uploads[0].p0;
				for (size_t k = 1; k < uploads.size() && rows; k++)
					rows = uploads[k].frame_bytes == uploads[0].frame_bytes && !uploads[k].planar2 && uploads[k].buffer->frames == frames
						&& (const char*)sources[uploads[k].pin].data - (const char*)sources[uploads[k - 1].pin].data == src_pitch
						&& (const char*)uploads[k].p0 - (const char*)uploads[k - 1].p0 == dst_pitch;
				rows = rows && src_pitch > 0 && dst_pitch > 0 && (size_t)src_pitch >= (size_t)frames * uploads[0].frame_bytes;
			}
			if (!rows)      // separate copies: keep their number per render in the hundreds
				while (nchunks > 1 && (size_t)nchunks * file_count > 512) nchunks /= 2;
			auto progress = std::make_shared<Stream_progress>();
			progress->stream = cur_stream();
			for (int c = 0; c < nchunks; c++)
			{
				if (rows)
				{
					const int64_t f0 = nchunks == 1 ? 0 : ((frames * c / nchunks) & ~(int64_t)63);
					const int64_t f1 = c == nchunks - 1 ? frames : ((frames * (c + 1) / nchunks) & ~(int64_t)63);
					const size_t fb = uploads[0].frame_bytes;
					if (f1 > f0)
						abi(nodey_memcpy2d_h2d((char*)uploads[0].p0 + (size_t)f0 * fb, (size_t)dst_pitch, (const char*)sources[uploads[0].pin].data + (size_t)f0 * fb,
											   (size_t)src_pitch, (size_t)(f1 - f0) * fb, uploads.size(), cur_stream()), "audio_input");
				}
				else
				for (const Pending& u : uploads)
				{
					const int64_t total = u.buffer->frames;
					const int64_t f0 = nchunks == 1 ? 0 : ((total * c / nchunks) & ~(int64_t)63);
					const int64_t f1 = c == nchunks - 1 ? total : ((total * (c + 1) / nchunks) & ~(int64_t)63);
					if (f1 <= f0) continue;
					const Pcm_source& s = sources[u.pin];
					abi(nodey_memcpy_h2d((char*)u.p0 + (size_t)f0 * u.frame_bytes, (const char*)s.data + (size_t)f0 * u.frame_bytes,
										 (size_t)(f1 - f0) * u.frame_bytes, cur_stream()), "audio_input");
					if (u.planar2)
						abi(nodey_memcpy_h2d((char*)u.p1 + (size_t)f0 * u.frame_bytes, (const char*)s.data1 + (size_t)f0 * u.frame_bytes,
											 (size_t)(f1 - f0) * u.frame_bytes, cur_stream()), "audio_input");
				}
				auto ev = std::make_shared<infra::Device_event>();
				ev->record(cur_stream());
				progress->points.push_back({c == nchunks - 1 ? frames : ((frames * (c + 1) / nchunks) & ~(int64_t)63), ev});
			}
			for (Pending& u : uploads)
			{
				u.buffer->ready = progress->points.back().event;
				if (nchunks > 1) u.buffer->progress = progress;
				publish(output, std::format("output_{}", u.pin), u.buffer);
			}
		}
		// file payloads are pageable host memory: make sure the copies have landed before they are freed
		if (!files.empty()) abi(nodey_stream_synchronize(cur_stream()), "audio_input");
	}

	// ---------------------------------------------------------------------------------------------
	// audio_output
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_output::get_processor_info()
	{
		return {.identifier = "audio_output", .display_name = "Audio Output", .singleton = true, .generate = std::make_unique<Audio_output>,
				.description = "Sink of the render: keeps the rendered stream for the host and optionally writes it as a float WAV file."};
	}

	std::vector<Processor::Pin_attribute> Audio_output::get_pin_attributes() const { return {Pin_type::audio("input", "Input", true)}; }

	void Audio_output::process_payload(const Input_map& input, const Output_map&, const std::atomic<bool>&, std::any& user_data)
	{
		auto buffer = require_input(input, "input", "Audio output processor");
		Process_context* ctx = std::any_cast<Process_context>(&user_data);
		if (!ctx) throw std::bad_any_cast();
		ctx->rendered = buffer;
		ctx->preview.reset();
		if (!ctx->do_export)
		{
			// ---- preview (audio-io.cpp:478-638): per input frame swr_convert -> clamp -> queue; never flushed ----
			check_channels(*buffer, "Audio output");
			const nodey_resampler* plan = resampler_for(buffer->sample_rate);
			const int64_t total = nodey_resampler_producible(plan, buffer->frames, 0);
			if (total < 0) abi((int)total, "Audio output");
			// chunk sizes: what each swr_convert call returns for the stream's frames
			Frame_runs chunks;
			{
				int64_t fed = 0, made = 0;
				for (const auto& [len, count] : buffer->runs)
					for (int64_t k = 0; k < count; k++)
					{
						fed += len;
						const int64_t now = nodey_resampler_producible(plan, std::min(fed, buffer->frames), 0);
						const int64_t got = now - made;
						made = now;
						if (!chunks.empty() && chunks.back().first == got) chunks.back().second++;
						else chunks.emplace_back(got, 1);
					}
			}
			const size_t bytes = Arena::padded((size_t)std::max<int64_t>(total, 1) * 2 * sizeof(float));
			auto block = std::make_shared<infra::Device_block>(bytes);
			if (total > 0)
			{
				const size_t plane = Arena::padded((size_t)total * sizeof(float));
				infra::Device_block planes(2 * plane);
				float* pl = (float*)planes.ptr;
				float* pr = (float*)((char*)planes.ptr + plane);
				abi(nodey_resampler_run(plan, pl, pr, buffer->plane[0], buffer->plane[1], buffer->format, buffer->channels, buffer->frames,
										0, total, cur_stream()), "Audio output");
				abi(nodey_preview_pack((float*)block->ptr, pl, pr, total, cur_stream()), "Audio output");
			}
			auto pv = new_buffer(block, block->ptr, nullptr, FMT_FLT, 48000, 2, total, std::move(chunks), buffer->pts_seconds);
			auto ev = std::make_shared<infra::Device_event>();
			ev->record(cur_stream());
			pv->ready = std::move(ev);
			ctx->preview = pv;
			abi(nodey_stream_synchronize(cur_stream()), "audio_output");
			if (ctx->preview_sink && total > 0)
			{
				// the "audio device": chunks in order from one host image; a false return stops the preview
				std::vector<float> host((size_t)total * 2);
				abi(nodey_memcpy_d2h(host.data(), block->ptr, host.size() * sizeof(float), cur_stream()), "audio_output");
				abi(nodey_stream_synchronize(cur_stream()), "audio_output");
				int64_t at = 0;
				bool go = true;
				for (const auto& [len, count] : pv->runs)
					for (int64_t k = 0; k < count && go; k++)
					{
						if (len > 0) go = ctx->preview_sink(host.data() + 2 * at, len);
						at += len;
						if (ctx->time) ctx->time->store((double)at / 48000.0);
					}
			}
			if (ctx->time && !ctx->preview_sink) ctx->time->store((double)total / 48000.0);
			return;
		}
		abi(nodey_stream_synchronize(cur_stream()), "audio_output");     // the sink is where the host waits for the device
		const auto ends_with_wav = [](const std::string& path)
		{
			if (path.size() < 4) return false;
			std::string tail = path.substr(path.size() - 4);
			for (char& c : tail) c = (char)std::tolower((unsigned char)c);
			return tail == ".wav";
		};
		if (!ctx->export_path.empty() && !ends_with_wav(ctx->export_path))
		{
			// ---- MP3 like the reference (audio-io.cpp:640-841): the stream goes to LAME in its own sample format, in the
			// frame sizes its producer recorded; host/src/mp3-export.cpp holds the call sequence ----
			check_channels(*buffer, "Audio output");
			Host_stream hs;
			hs.format = buffer->format; hs.sample_rate = buffer->sample_rate; hs.channels = buffer->channels;
			hs.frames = buffer->frames; hs.pts_seconds = buffer->pts_seconds; hs.runs = buffer->runs;
			hs.stamp = buffer->stamp; hs.stamp_origin = buffer->stamp_origin; hs.frame_pts = buffer->frame_pts;
			const bool planar2 = format_is_planar(buffer->format) && buffer->channels == 2;
			const size_t plane_bytes = buffer->plane_bytes();
			std::vector<unsigned char> host[2];
			for (int c = 0; c < (planar2 ? 2 : 1); c++)
			{
				host[c].resize(std::max<size_t>(plane_bytes, 1));
				if (plane_bytes) abi(nodey_memcpy_d2h(host[c].data(), buffer->plane[c], plane_bytes, cur_stream()), "audio_output");
				hs.plane[c] = host[c].data();
			}
			abi(nodey_stream_synchronize(cur_stream()), "audio_output");
			const double end = export_mp3(hs, ctx->export_path, ctx->kbps, ctx->time ? ctx->time->load() : 0.0);
			if (ctx->time) ctx->time->store(end);
			return;
		}
		if (!ctx->export_path.empty())
		{
			// interleaved float WAV: the lossless way out for hosts without LAME, and what the parity tests read back
			check_channels(*buffer, "Audio output");
			infra::Device_block tmp((size_t)buffer->frames * (size_t)buffer->channels * sizeof(float));
			abi(nodey_extract_interleaved((float*)tmp.ptr, buffer->plane[0], buffer->plane[1], buffer->format, buffer->frames, buffer->channels, cur_stream()),
				"audio_output");
			std::vector<float> host((size_t)buffer->frames * (size_t)buffer->channels);
			abi(nodey_memcpy_d2h(host.data(), tmp.ptr, host.size() * sizeof(float), cur_stream()), "audio_output");
			abi(nodey_stream_synchronize(cur_stream()), "audio_output");
			std::ofstream f(ctx->export_path, std::ios::binary);
			if (!f)                                                      // audio-io.cpp:648-654
				throw Runtime_error("Failed to open output file", "Cannot open the output file for writing. Check if the path is valid and writable.",
									std::format("Output path: {}", ctx->export_path));
			// do_export's pts rule (audio-io.cpp:833-839): before EVERY frame, (int)((frame_begin - time) * sample_rate) samples of
			// silence are encoded when that is positive; `time` then follows the frame ends.  frame_begin is the frame's own
			// stamp, by its producer's rule (Frame_clock): the end-time stamps of amix / bimix (App. C4) make the reference
			// prepend almost one frame of silence to their exports (1023 samples in front of 1024-sample frames: the stamp is
			// truncated to microseconds) and another len_k - len_{k-1} samples wherever their frame size grows
			std::vector<Export_step> steps;
			const double end_time = export_steps(buffer->runs, buffer->frames, buffer->clock(), buffer->sample_rate, ctx->time ? ctx->time->load() : 0.0, steps);
			int64_t silence_total = 0;
			for (const Export_step& step : steps) silence_total += step.silence;
			const uint32_t data_bytes = (uint32_t)((host.size() + (size_t)silence_total * (size_t)buffer->channels) * sizeof(float));
			const uint16_t tag = 3, ch = (uint16_t)buffer->channels, bits = 32, align = (uint16_t)(4 * buffer->channels);
			const uint32_t rate = (uint32_t)buffer->sample_rate, byte_rate = rate * align, riff = 36 + data_bytes, fmt_size = 16;
			f.write("RIFF", 4); f.write((const char*)&riff, 4); f.write("WAVEfmt ", 8); f.write((const char*)&fmt_size, 4);
			f.write((const char*)&tag, 2); f.write((const char*)&ch, 2); f.write((const char*)&rate, 4); f.write((const char*)&byte_rate, 4);
			f.write((const char*)&align, 2); f.write((const char*)&bits, 2); f.write("data", 4); f.write((const char*)&data_bytes, 4);
			std::vector<float> silence;
			for (const Export_step& step : steps)
			{
				if (step.silence > 0)
				{
					silence.assign((size_t)step.silence * (size_t)buffer->channels, 0.0f);
					f.write((const char*)silence.data(), (std::streamsize)(silence.size() * sizeof(float)));
				}
				f.write((const char*)(host.data() + (size_t)step.at * (size_t)buffer->channels), (std::streamsize)((size_t)step.nb * (size_t)buffer->channels * sizeof(float)));
			}
			if (ctx->time) ctx->time->store(end_time);      // `time` ends at the end of the last frame (audio-io.cpp:838)
			return;
		}
		// no export: `time` still ends where the last frame does
		if (ctx->time)
		{
			std::vector<Export_step> steps;
			ctx->time->store(export_steps(buffer->runs, buffer->frames, buffer->clock(), buffer->sample_rate, ctx->time->load(), steps));
		}
	}

	// ---------------------------------------------------------------------------------------------
	// audio_volume_adjust  (audio-vol.cpp:75-100, 102-250)
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_vol::get_processor_info()
	{
		return {.identifier = "audio_volume_adjust", .display_name = "Adjust Volume", .singleton = false, .generate = std::make_unique<Audio_vol>,
				.description = "Multiplies every sample by the gain (0..10). Integer formats are scaled in float and truncated like the reference; "
							   "format, layout and rate pass through."};
	}

	std::vector<Processor::Pin_attribute> Audio_vol::get_pin_attributes() const
	{
		return {Pin_type::audio("output", "Output", false), Pin_type::audio("input", "Input", true)};
	}

	void Audio_vol::set_volume(float v) { volume = std::clamp(v, 0.0f, 10.0f); }

	void Audio_vol::deserialize(const Json::Value& value)
	{
		if (value.isObject() && value.isMember("volume") && value["volume"].isDouble()) set_volume(value["volume"].asFloat());
	}

	namespace
	{
		void gain_into(const Audio_buffer& in, float volume, void* p0, void* p1)
		{
			const int64_t n = format_is_planar(in.format) ? in.frames : in.frames * in.channels;
			abi(nodey_gain(p0, in.plane[0], in.format, n, volume, cur_stream()), "Volume adjust");
			if (format_is_planar(in.format) && in.channels == 2) abi(nodey_gain(p1, in.plane[1], in.format, n, volume, cur_stream()), "Volume adjust");
		}
	}

	bool Audio_vol::process_batch(const std::vector<Batch_item>& items)
	{
		std::vector<std::shared_ptr<const Audio_buffer>> ins;
		size_t bytes = 0;
		for (const auto& it : items)
		{
			ins.push_back(require_input(*it.input, "input", "Volume adjust processor"));
			check_channels(*ins.back(), "Volume adjust", Node_kind::volume);
			const bool planar2 = format_is_planar(ins.back()->format) && ins.back()->channels == 2;
			bytes += Arena::padded(ins.back()->plane_bytes()) * (planar2 ? 2 : 1);
		}
		// Float streams are published LAZY (Lazy_gain): a mixer behind this node folds the gain into its own read, any
		// other consumer triggers the pass on first use.  NODEY_EAGER_GAIN=1 runs the pass here, as round 1 did.
		static const bool eager = getenv("NODEY_EAGER_GAIN") != nullptr;
		if (!eager)
		{
			bool all_float = true;
			for (const auto& in : ins) all_float = all_float && (in->format == FMT_FLT || in->format == FMT_FLTP);
			if (all_float)
			{
				for (size_t k = 0; k < items.size(); k++)
				{
					const Audio_buffer& in = *ins[k];
					auto b = inherit_stamps(new_buffer(nullptr, nullptr, nullptr, in.format, in.sample_rate, in.channels, in.frames, in.runs, in.pts_seconds), in);
					b->lazy = std::make_shared<Lazy_gain>();
					b->lazy->source = ins[k];
					b->lazy->gain = static_cast<Audio_vol*>(items[k].processor)->volume;
					b->ready = in.ready;          // a reader of source * gain orders itself after the source
					if (!b->ready) { auto ev = std::make_shared<infra::Device_event>(); ev->record(cur_stream()); b->ready = ev; }
					publish(*items[k].output, "output", b);
				}
				return true;
			}
		}
		Arena arena(bytes);
		// float streams (planes) of the whole batch go to ONE launch; integer formats run per stream.  Products are
		// published after everything is enqueued: publish() records the event the consumers wait on.
		std::vector<void*> bd; std::vector<const void*> bs; std::vector<int64_t> bn; std::vector<float> bv;
		std::vector<std::pair<void*, void*>> planes;
		for (size_t k = 0; k < items.size(); k++)
		{
			const Audio_buffer& in = *ins[k];
			const bool planar2 = format_is_planar(in.format) && in.channels == 2;
			void* p0 = arena.take(in.plane_bytes());
			void* p1 = planar2 ? arena.take(in.plane_bytes()) : nullptr;
			planes.emplace_back(p0, p1);
			const float volume = static_cast<Audio_vol*>(items[k].processor)->volume;
			if (in.format == FMT_FLT || in.format == FMT_FLTP)
			{
				const int64_t n = format_is_planar(in.format) ? in.frames : in.frames * in.channels;
				bd.push_back(p0); bs.push_back(in.plane[0]); bn.push_back(n); bv.push_back(volume);
				if (planar2) { bd.push_back(p1); bs.push_back(in.plane[1]); bn.push_back(n); bv.push_back(volume); }
			}
			else gain_into(in, volume, p0, p1);
		}
		if (!bd.empty())
			abi(nodey_gain_tracks(bd.data(), bs.data(), bn.data(), bv.data(), FMT_FLT, (int)bd.size(), cur_stream()), "Volume adjust");
		for (size_t k = 0; k < items.size(); k++)
		{
			const Audio_buffer& in = *ins[k];
			publish(*items[k].output, "output", inherit_stamps(new_buffer(arena.block, planes[k].first, planes[k].second, in.format, in.sample_rate,
																			  in.channels, in.frames, in.runs, in.pts_seconds), in));
		}
		return true;
	}

	void Audio_vol::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>& stop, std::any& user_data)
	{
		process_batch({Batch_item{this, &input, &output, &stop, &user_data}});
	}

	// ---------------------------------------------------------------------------------------------
	// velocity_modifier / pitch_modifier  (audio-velocity.cpp:265-477)
	// ---------------------------------------------------------------------------------------------
	struct Soundtouch_params
	{
		float rate, pitch;
		bool reference_schedule;
		static Soundtouch_params of(const Processor* p)
		{
			static const bool everywhere = getenv("NODEY_REFERENCE_SCHEDULE") != nullptr;
			if (const auto* v = dynamic_cast<const Velocity_modifier*>(p))
				return {v->velocity, v->keep_pitch ? 1 / v->velocity : 1, v->reference_schedule || everywhere};   // audio-velocity.cpp:457
			const auto* q = dynamic_cast<const Pitch_modifier*>(p);
			return {1.0f, std::pow(2.0f, q->pitch / 12.0f), q->reference_schedule || everywhere};                 // audio-velocity.cpp:473-474
		}
	};

	namespace
	{
		constexpr int kSoundtouchFrame = 1152;     // canonical putSamples / output chunk
		constexpr size_t kMaxTracksPerLaunch = 256;

		bool soundtouch_batch(const std::vector<Processor::Batch_item>& items, const char* title)
		{
			struct Entry { size_t item; std::shared_ptr<const Audio_buffer> in; Soundtouch_params prm; };
			std::map<std::tuple<int, int, int64_t, uint32_t, uint32_t, const Stream_progress*, bool>, std::vector<Entry>> groups;
			for (size_t k = 0; k < items.size(); k++)
			{
				// inputs that arrive chunk by chunk are not waited for here: the chunks below wait for the prefix they need
				auto in = require_input(*items[k].input, "input", title, false);
				if (!in->progress && in->ready) in->ready->wait_on(cur_stream());
				check_channels(*in, title, Node_kind::soundtouch);
				const Soundtouch_params prm = Soundtouch_params::of(items[k].processor);
				uint32_t rb, pb;
				memcpy(&rb, &prm.rate, 4); memcpy(&pb, &prm.pitch, 4);
				const Stream_progress* pg = in->progress.get();
				groups[{in->sample_rate, in->channels, in->frames, rb, pb, pg, prm.reference_schedule}].push_back({k, std::move(in), prm});
			}
			const nodey_stream_t main_stream = cur_stream();
			for (auto& [key, all] : groups)
			{
				const auto [rate_hz, ch, n, rb_, pb_, pg_, ref_schedule] = key;
				(void)rb_; (void)pb_; (void)pg_;
				const std::shared_ptr<nodey_soundtouch> plan = soundtouch_for(rate_hz, ch, all.front().prm.rate, all.front().prm.pitch);
				nodey_soundtouch* st = plan.get();
				int64_t nseq = 0;
				int64_t m = nodey_soundtouch_out_frames(st, n, kSoundtouchFrame, &nseq);
				if (m < 0) abi((int)m, title);
				// frame sizes of the product: canonical 1152-sample frames of the complete render, or -- App. C7 switch --
				// what the reference's loop receives, without the tail it never flushes (a prefix of the same samples)
				Frame_runs out_runs = uniform_frame_runs(m, kSoundtouchFrame);
				if (ref_schedule)
				{
					std::vector<int64_t> rl(64), rc(64);
					int64_t n_runs = 0;
					int flushed = 0;
					int64_t total = 0;
					for (int attempt = 0; attempt < 2; attempt++)
					{
						total = nodey_soundtouch_reference_schedule(st, n, kSoundtouchFrame, all.front().prm.rate, rl.data(), rc.data(), (int64_t)rl.size(),
																	&n_runs, &flushed);
						if (total < 0) abi((int)total, title);
						if (n_runs <= (int64_t)rl.size()) break;
						rl.resize((size_t)n_runs); rc.resize((size_t)n_runs);
					}
					if (total > m) THROW_LOGIC_ERROR("reference schedule yields {} frames, the complete render {}", total, m);
					m = total;
					out_runs.clear();
					for (int64_t k = 0; k < n_runs; k++) out_runs.emplace_back(rl[(size_t)k], rc[(size_t)k]);
				}
				for (size_t first = 0; first < all.size(); first += kMaxTracksPerLaunch)
				{
					const size_t cnt = std::min(kMaxTracksPerLaunch, all.size() - first);
					const size_t in_stride = Arena::padded((size_t)n * ch * sizeof(float)) / sizeof(float);
					const size_t out_stride = Arena::padded((size_t)std::max<int64_t>(m, 1) * ch * sizeof(float)) / sizeof(float);
					const size_t offs_stride = (size_t)std::max<int64_t>(nseq - 1, 1);
					// A8: extract_samples_interleaved is the identity on float samples: FLT / FLTP streams are read in
					// place through per-track pointers; integer formats are converted into one contiguous batch first
					bool in_place = true;
					for (size_t k = 0; k < cnt; k++)
					{
						const int f = all[first + k].in->format;
						in_place = in_place && (f == FMT_FLT || f == FMT_FLTP) && f == all[first].in->format;
					}
					// chunk plan: the chain of a track runs as several launches so that the next node can start on a prefix
					// of this node's output (and this node on a prefix of its input, when its producer works the same way)
					const std::shared_ptr<const Stream_progress> in_progress = all[first].in->progress;
					constexpr int kMaxChunks = 64;
					int64_t in_need[kMaxChunks], out_ready[kMaxChunks];
					int nchunks = 1;
					if (in_place && m > 0)
					{
						nchunks = nodey_soundtouch_chunks(st, n, kSoundtouchFrame, m, stream_chunk_count(), in_need, out_ready, kMaxChunks);
						if (nchunks < 0) abi(nchunks, title);
					}
					const bool chunked = in_place && m > 0 && (nchunks > 1 || in_progress);
					// next to a chunk-wise producer this node runs on the lane's OTHER stream, so that the two overlap.  Its
					// memory is then allocated in that stream's order too (the allocator hands a block freed on a stream
					// straight back to that stream): nothing here may order the side stream after what the producer has
					// already enqueued on the main one
					// Two streams when chunked: `search` carries the sequential WSOLA chain, chunk after chunk with nothing in
					// between (it only ever waits for input), `run` the tails (cross-fade + FIR + cubic of chunk c after search
					// c) and the node's progress events.  Without side streams both are the lane's stream.
					const nodey_stream_t search = chunked ? stream_after(in_progress.get()) : main_stream;
					const nodey_stream_t run = chunked && search != main_stream ? next_in_cycle(search) : search;
					std::shared_ptr<infra::Device_block> offs_block;
					if (chunked)
					{
						Stream_scope on_search(search);     // the offset trace is first written (cleared) on the search stream
						offs_block = std::make_shared<infra::Device_block>(Arena::padded(offs_stride * sizeof(int32_t) * cnt));
					}
					Stream_scope scope(run);
					Arena arena(out_stride * sizeof(float) * cnt);
					float* out_base = (float*)arena.take(out_stride * sizeof(float) * cnt);
					std::shared_ptr<Stream_progress> progress;
					std::shared_ptr<infra::Device_event> done;
					if (chunked)
					{
						int32_t* offs = (int32_t*)offs_block->ptr;
						const bool planes = all[first].in->format == FMT_FLTP && ch == 2;
						std::vector<const float*> pa(cnt), pb(cnt, nullptr);
						for (size_t k = 0; k < cnt; k++)
						{
							pa[k] = (const float*)all[first + k].in->plane[0];
							if (planes) pb[k] = (const float*)all[first + k].in->plane[1];
						}
						progress = std::make_shared<Stream_progress>();
						progress->stream = run;
						Progress_waiter input{in_progress.get()};
						const bool split = search != run;
						if (split)
						{
							// the whole chain first (host order = device order on `search`), an event after every chunk
							std::vector<std::shared_ptr<infra::Device_event>> searched;
							for (int c = 0; c < nchunks; c++)
							{
								input.need(in_need[c], search);
								abi(nodey_soundtouch_run_tracks_chunk(st, out_base, (int64_t)out_stride, pa.data(), planes ? pb.data() : nullptr, (int)cnt, n,
																	  kSoundtouchFrame, m, offs, (int64_t)offs_stride, c, nchunks, 1, search), title);
								searched.push_back(std::make_shared<infra::Device_event>());
								searched.back()->record(search);
							}
							input.all(search);
							for (int c = 0; c < nchunks; c++)
							{
								searched[(size_t)c]->wait_on(run);       // also orders the tail after the input prefix the search waited for
								abi(nodey_soundtouch_run_tracks_chunk(st, out_base, (int64_t)out_stride, pa.data(), planes ? pb.data() : nullptr, (int)cnt, n,
																	  kSoundtouchFrame, m, offs, (int64_t)offs_stride, c, nchunks, 2, run), title);
								auto ev = std::make_shared<infra::Device_event>();
								ev->record(run);
								progress->points.push_back({out_ready[c], ev});
							}
						}
						else
						{
							for (int c = 0; c < nchunks; c++)
							{
								input.need(in_need[c], run);
								abi(nodey_soundtouch_run_tracks_chunk(st, out_base, (int64_t)out_stride, pa.data(), planes ? pb.data() : nullptr, (int)cnt, n,
																	  kSoundtouchFrame, m, offs, (int64_t)offs_stride, c, nchunks, 0, run), title);
								auto ev = std::make_shared<infra::Device_event>();
								ev->record(run);
								progress->points.push_back({out_ready[c], ev});
							}
							input.all(run);       // whatever the chunk plan needed, the product is complete only after the whole input is
						}
						done = progress->points.back().event;
						// (the offset trace goes out of scope below: Device_block frees are ordered after everything enqueued on
						// every stream of the lane, Lane_registry::free_ordered)
					}
					else if (in_place)
					{
						const bool planes = all[first].in->format == FMT_FLTP && ch == 2;
						std::vector<const float*> pa(cnt), pb(cnt, nullptr);
						for (size_t k = 0; k < cnt; k++)
						{
							pa[k] = (const float*)all[first + k].in->plane[0];
							if (planes) pb[k] = (const float*)all[first + k].in->plane[1];
						}
						if (m > 0)
							abi(nodey_soundtouch_run_tracks(st, out_base, (int64_t)out_stride, pa.data(), planes ? pb.data() : nullptr, (int)cnt, n,
															kSoundtouchFrame, m, nullptr, 0, cur_stream()), title);
					}
					else
					{
						infra::Device_block staging(in_stride * sizeof(float) * cnt);
						for (size_t k = 0; k < cnt; k++)
						{
							const Audio_buffer& in = *all[first + k].in;
							if (in.progress && in.ready) in.ready->wait_on(cur_stream());
							abi(nodey_extract_interleaved((float*)staging.ptr + k * in_stride, in.plane[0], in.plane[1], in.format, n, ch, cur_stream()), title);
						}
						if (m > 0)
							abi(nodey_soundtouch_run(st, out_base, (int64_t)out_stride, (const float*)staging.ptr, (int64_t)in_stride, (int)cnt, n,
													 kSoundtouchFrame, m, nullptr, 0, cur_stream()), title);
					}
					for (size_t k = 0; k < cnt; k++)
					{
						const Entry& e = all[first + k];
						// App. C8 (same switch): the reference stamps its frames with (int64_t)(float)(seconds * 1e6) microseconds
						// (construct_audio_frame_float takes `float time_us`, audio-velocity.cpp:234-250); canonical: the exact start
						// -- for EVERY frame, from a running double of seconds that starts at the first input frame's stamp (:388, :313-318)
						auto b = new_buffer(arena.block, out_base + k * out_stride, nullptr, FMT_FLT, rate_hz, ch, m, out_runs, e.in->pts_seconds);
						if (ref_schedule)
						{
							b->stamp = STAMP_START_FLOAT_US; b->stamp_origin = e.in->pts_seconds;
							b->pts_seconds = Frame_clock(STAMP_START_FLOAT_US, e.in->pts_seconds, rate_hz).next(0);
						}
						if (done) { b->ready = done; b->progress = progress; }
						publish(*items[e.item].output, "output", b);
					}
				}
			}
			return true;
		}
	}

	infra::Processor::Info Velocity_modifier::get_processor_info()
	{
		return {.identifier = "velocity_modifier", .display_name = "Velocity Modifier", .singleton = false,
				.generate = std::make_unique<Velocity_modifier>,
				.description = "Changes playback speed (0.5x..3x) with SoundTouch's rate transposer and WSOLA time stretcher; "
							   "'keep_pitch' compensates the pitch change. Output: 32-bit float interleaved at the input rate."};
	}
	std::vector<Processor::Pin_attribute> Velocity_modifier::get_pin_attributes() const
	{
		return {Pin_type::audio("output", "Output", false), Pin_type::audio("input", "Input", true)};
	}
	Json::Value Velocity_modifier::serialize() const
	{
		Json::Value value;
		value["velocity"] = velocity;
		value["keep_pitch"] = keep_pitch;
		if (reference_schedule) value["reference_schedule"] = true;      // written only when set: reference files stay as they are
		return value;
	}
	void Velocity_modifier::deserialize(const Json::Value& value)
	{
		if (value.isMember("velocity") && value["velocity"].isDouble()) velocity = value["velocity"].asFloat();
		if (value.isMember("keep_pitch") && value["keep_pitch"].isBool()) keep_pitch = value["keep_pitch"].asBool();
		if (value.isMember("reference_schedule") && value["reference_schedule"].isBool()) reference_schedule = value["reference_schedule"].asBool();
	}
	bool Velocity_modifier::process_batch(const std::vector<Batch_item>& items) { return soundtouch_batch(items, "Velocity Modifier"); }
	void Velocity_modifier::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>& stop, std::any& user_data)
	{
		soundtouch_batch({Batch_item{this, &input, &output, &stop, &user_data}}, "Velocity Modifier");
	}

	infra::Processor::Info Pitch_modifier::get_processor_info()
	{
		return {.identifier = "pitch_modifier", .display_name = "Pitch Modifier", .singleton = false, .generate = std::make_unique<Pitch_modifier>,
				.description = "Shifts the pitch by a number of semitones at constant duration (SoundTouch). "
							   "Output: 32-bit float interleaved at the input rate."};
	}
	std::vector<Processor::Pin_attribute> Pitch_modifier::get_pin_attributes() const
	{
		return {Pin_type::audio("output", "Output", false), Pin_type::audio("input", "Input", true)};
	}
	Json::Value Pitch_modifier::serialize() const
	{
		Json::Value value;
		value["pitch"] = pitch;
		if (reference_schedule) value["reference_schedule"] = true;
		return value;
	}
	void Pitch_modifier::deserialize(const Json::Value& value)
	{
		if (value.isMember("pitch") && value["pitch"].isDouble()) pitch = value["pitch"].asFloat();
		if (value.isMember("reference_schedule") && value["reference_schedule"].isBool()) reference_schedule = value["reference_schedule"].asBool();
	}
	bool Pitch_modifier::process_batch(const std::vector<Batch_item>& items) { return soundtouch_batch(items, "Pitch Modifier"); }
	void Pitch_modifier::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>& stop, std::any& user_data)
	{
		soundtouch_batch({Batch_item{this, &input, &output, &stop, &user_data}}, "Pitch Modifier");
	}

	// ---------------------------------------------------------------------------------------------
	// mixers: shared machinery
	// ---------------------------------------------------------------------------------------------
	namespace
	{
		struct Segment { int input; int64_t out_start, src_start, len; };

		// swr output of one input as two float planes [2][produced]; `contiguous` = it lands at output 0 unbroken
		struct Resampled
		{
			std::shared_ptr<infra::Device_block> block;
			float* l = nullptr; float* r = nullptr;
			int64_t len = 0;
		};

		// resample one whole input (flush: with the reflected tail) to `produced` frames
		Resampled resample_input(const Audio_buffer& in, int64_t produced, bool flush, const char* node)
		{
			Resampled out;
			out.len = produced;
			const size_t plane = Arena::padded((size_t)std::max<int64_t>(produced, 1) * sizeof(float));
			out.block = std::make_shared<infra::Device_block>(2 * plane);
			out.l = (float*)out.block->ptr;
			out.r = (float*)((char*)out.block->ptr + plane);
			if (produced > 0)
				abi(nodey_resampler_run(resampler_for(in.sample_rate), out.l, out.r, in.plane[0], in.plane[1], in.format, in.channels, in.frames,
										flush ? 1 : 0, produced, cur_stream()), node);
			return out;
		}

		// place an input's resampled frames into zero-initialised planes of `total` frames (gappy inputs)
		Resampled scatter(const Resampled& src, const std::vector<Segment>& segs, int input, int64_t total, const char* node)
		{
			Resampled out;
			out.len = total;
			const size_t plane = Arena::padded((size_t)std::max<int64_t>(total, 1) * sizeof(float));
			out.block = std::make_shared<infra::Device_block>(2 * plane);
			out.l = (float*)out.block->ptr;
			out.r = (float*)((char*)out.block->ptr + plane);
			abi(nodey_memset(out.block->ptr, 0, 2 * plane, cur_stream()), node);
			for (const Segment& s : segs)
			{
				if (s.input != input) continue;
				abi(nodey_memcpy_d2d(out.l + s.out_start, src.l + s.src_start, (size_t)s.len * sizeof(float), cur_stream()), node);
				abi(nodey_memcpy_d2d(out.r + s.out_start, src.r + s.src_start, (size_t)s.len * sizeof(float), cur_stream()), node);
			}
			return out;
		}

		bool single_front_segment(const std::vector<Segment>& segs, int input, int64_t* len)
		{
			int count = 0;
			*len = 0;
			for (const Segment& s : segs)
				if (s.input == input) { count++; if (s.out_start != 0 || s.src_start != 0) return false; *len = s.len; }
			return count <= 1;
		}

		int64_t produced_of(const std::vector<Segment>& segs, int input)
		{
			int64_t n = 0;
			for (const Segment& s : segs)
				if (s.input == input) n = std::max(n, s.src_start + s.len);
			return n;
		}
	}

	// ---------------------------------------------------------------------------------------------
	// audio_amix  (audio-amix.cpp:86-324)
	// ---------------------------------------------------------------------------------------------
	Audio_amix::Audio_amix() : volumes(2, 1.0f), locks(2, false) {}     // App. C9: a fresh node mixes at unity

	infra::Processor::Info Audio_amix::get_processor_info()
	{
		return {.identifier = "audio_amix", .display_name = "Audio Amix", .singleton = false, .generate = std::make_unique<Audio_amix>,
				.description = "Mixes 1..16 inputs into one stereo stream: every input is converted to 48 kHz stereo float planar "
							   "(libswresample defaults), scaled by its volume and summed in input order."};
	}

	std::vector<Processor::Pin_attribute> Audio_amix::get_pin_attributes() const
	{
		std::vector<Processor::Pin_attribute> pins;
		pins.push_back(Pin_type::audio("output", "Output", false));
		for (int i = 0; i < input_num; i++) pins.push_back(Pin_type::audio(std::format("input_{}", i + 1), std::format("Input {}", i + 1), true));
		return pins;
	}

	Json::Value Audio_amix::serialize() const
	{
		Json::Value value;
		value["input_num"] = input_num;
		for (int i = 0; i < input_num; i++)
		{
			value[std::format("volumes{}", i)] = i < (int)volumes.size() ? volumes[(size_t)i] : 1.0f;
			value[std::format("locks{}", i)] = i < (int)locks.size() ? (bool)locks[(size_t)i] : false;
		}
		if (start_time_stamps) value["start_time_stamps"] = true;      // written only when set: reference files stay as they are
		return value;
	}

	void Audio_amix::deserialize(const Json::Value& value)
	{
		if (!value.isMember("input_num"))
			throw Runtime_error("Failed to deserialize JSON file",
								// (the reference's text names Audio_bimix here: audio-amix.cpp:412)
								"Audio_bimix failed to serialize the JSON input because of missing or invalid fields.", "Wrong field: input_num");
		input_num = std::clamp(value["input_num"].asInt(), 1, NODEY_MAX_MIX_INPUTS);       // audio-amix.cpp:342
		locks.clear();
		volumes.clear();
		for (int i = 0; i < input_num; i++)
		{
			volumes.push_back(value[std::format("volumes{}", i)].asFloat());
			locks.push_back(value[std::format("locks{}", i)].asBool());
		}
		start_time_stamps = value.isMember("start_time_stamps") && value["start_time_stamps"].isBool() && value["start_time_stamps"].asBool();
	}

	namespace
	{
		// frame bookkeeping of the reference loop: total length, placement of every input, own frame sizes
		struct Amix_plan { int64_t total = 0; std::vector<Segment> segs; Frame_runs out_runs; };

		// everything process_payload decides before it touches the device
		struct Amix_job
		{
			std::vector<std::shared_ptr<const Audio_buffer>> ins;
			std::vector<float> vol;
			std::shared_ptr<const Amix_plan> plan;
			bool start_stamps = false;   // Audio_amix::start_time_stamps
			bool fused = false;     // same-rate inputs that need resampling, each landing unbroken at 0, exact-rational plan
			std::vector<int64_t> front_len;
		};

		Amix_job amix_prepare(int nin, const std::vector<float>& volumes, const Processor::Input_map& input)
		{
			Amix_job job;
			std::vector<const Frame_runs*> runs;
			std::vector<int> rates;
			for (int i = 0; i < nin; i++)
			{
				// a chunk-wise input is not waited for here: the batch path below consumes it chunk by chunk, every other
				// path waits in amix_execute
				job.ins.push_back(require_input(input, std::format("input_{}", i + 1), "Audio Mixer processor", false, true));
				if (!job.ins.back()->progress && job.ins.back()->ready) job.ins.back()->ready->wait_on(cur_stream());
				check_channels(*job.ins.back(), "Audio mixer");
				runs.push_back(&job.ins.back()->runs);
				rates.push_back(job.ins.back()->sample_rate);
			}
			job.vol = volumes;
			job.vol.resize((size_t)nin, 1.0f);

			// (cached: the 256 per-track amix nodes of a batch render share one plan)
			const Runs_flat flat(runs);
			static std::map<std::vector<int64_t>, std::shared_ptr<const Amix_plan>> plan_cache;
			std::vector<int64_t> key;
			for (const int r : rates) key.push_back(r);
			key.push_back(-1);
			key.insert(key.end(), flat.off.begin(), flat.off.end());
			key.push_back(-1);
			key.insert(key.end(), flat.len.begin(), flat.len.end());
			key.push_back(-1);
			key.insert(key.end(), flat.count.begin(), flat.count.end());
			{
				std::lock_guard lock(plan_mutex);
				const auto it = plan_cache.find(key);
				if (it != plan_cache.end()) job.plan = it->second;
			}
			if (!job.plan)
			{
				auto fresh = std::make_shared<Amix_plan>();
				std::vector<int32_t> seg_in(64);
				std::vector<int64_t> seg_out(64), seg_src(64), seg_len(64), run_len(64), run_cnt(64);
				int64_t nseg = 0, nrun = 0;
				for (int attempt = 0; attempt < 2; attempt++)
				{
					fresh->total = nodey_amix_plan(rates.data(), nin, flat.off.data(), flat.len.data(), flat.count.data(), 0, seg_in.data(),
												   seg_out.data(), seg_src.data(), seg_len.data(), (int64_t)seg_in.size(), &nseg, run_len.data(),
												   run_cnt.data(), (int64_t)run_len.size(), &nrun);
					if (fresh->total < 0) abi((int)fresh->total, "Audio mixer");
					if (nseg <= (int64_t)seg_in.size() && nrun <= (int64_t)run_len.size()) break;
					const size_t a = (size_t)std::max<int64_t>(nseg, 64), b = (size_t)std::max<int64_t>(nrun, 64);
					seg_in.resize(a); seg_out.resize(a); seg_src.resize(a); seg_len.resize(a); run_len.resize(b); run_cnt.resize(b);
				}
				for (int64_t k = 0; k < nseg; k++) fresh->segs.push_back({seg_in[(size_t)k], seg_out[(size_t)k], seg_src[(size_t)k], seg_len[(size_t)k]});
				for (int64_t k = 0; k < nrun; k++) fresh->out_runs.emplace_back(run_len[(size_t)k], run_cnt[(size_t)k]);
				std::lock_guard lock(plan_mutex);
				if (plan_cache.size() > 256) plan_cache.clear();
				plan_cache[key] = fresh;
				job.plan = fresh;
			}

			// fused path: same source rate (needs resampling), same channel count, every input lands unbroken at 0
			job.fused = job.plan->total > 0;
			job.front_len.assign((size_t)nin, 0);
			for (int i = 0; i < nin && job.fused; i++)
				job.fused = job.ins[(size_t)i]->sample_rate == job.ins[0]->sample_rate && job.ins[(size_t)i]->sample_rate != 48000
						 && job.ins[(size_t)i]->channels == job.ins[0]->channels && single_front_segment(job.plan->segs, i, &job.front_len[(size_t)i]);
			if (job.fused)
			{
				int info[8];
				abi(nodey_resampler_info(resampler_for(job.ins[0]->sample_rate), info), "Audio mixer");
				job.fused = info[4] == 0;     // exact-rational plan: the tiled kernel exists
			}
			return job;
		}

		void amix_publish(const Amix_job& job, const Processor::Output_map& output, const std::shared_ptr<infra::Device_block>& block,
						  float* out_l, float* out_r, const std::shared_ptr<const Stream_progress>& progress = nullptr)
		{
			// pts: the reference stamps each frame with the running END time (audio-amix.cpp:199-201, App. C4)
			Frame_runs out_runs = job.plan->out_runs;
			auto buffer = new_buffer(block, out_l, out_r, FMT_FLTP, 48000, 2, job.plan->total, std::move(out_runs), 0.0);
			if (!job.start_stamps) buffer = end_time_stamps(std::move(buffer));       // the reference's stamps unless the node opts out
			if (progress && !progress->points.empty()) { buffer->ready = progress->points.back().event; buffer->progress = progress; }
			publish(output, "output", buffer);
		}

		void amix_execute(Amix_job& job, const Processor::Output_map& output)
		{
			const int nin = (int)job.ins.size();
			const auto& ins = job.ins;
			const int64_t total = job.plan->total;
			const std::vector<Segment>& segs = job.plan->segs;
			bool fused = job.fused;
			for (const auto& in : ins)       // chunk-wise inputs were not waited for in amix_prepare
				if (in->progress && in->ready) in->ready->wait_on(cur_stream());
			// lazy gain products: only the in-place read of the mix kernel below folds the gain in; every path that
			// resamples or converts first wants the ordinary buffer
			for (int i = 0; i < nin; i++)
			{
				const Audio_buffer& in = *ins[(size_t)i];
				int64_t len = 0;
				const bool in_place = !fused && in.sample_rate == 48000 && in.channels == 2 && (in.format == FMT_FLT || in.format == FMT_FLTP)
								   && single_front_segment(segs, i, &len) && len <= in.frames;
				if (in.is_lazy() && !in_place) { materialize(in, cur_stream()); if (in.ready) in.ready->wait_on(cur_stream()); }
			}

			const size_t plane = Arena::padded((size_t)std::max<int64_t>(total, 1) * sizeof(float));
			auto block = std::make_shared<infra::Device_block>(2 * plane);
			float* out_l = (float*)block->ptr;
			float* out_r = (float*)((char*)block->ptr + plane);

			if (total > 0 && fused)
			{
				std::vector<const void*> p0, p1;
				std::vector<int> fmt, ch;
				std::vector<int64_t> in_frames;
				for (const auto& in : ins) { p0.push_back(in->plane[0]); p1.push_back(in->plane[1]); fmt.push_back(in->format); ch.push_back(in->channels); in_frames.push_back(in->frames); }
				const int rc = nodey_resample_mix(resampler_for(ins[0]->sample_rate), out_l, out_r, p0.data(), p1.data(), fmt.data(), ch.data(),
												  in_frames.data(), job.front_len.data(), job.vol.data(), nin, 1, total, cur_stream());
				if (rc == NODEY_E_RANGE) fused = false; else abi(rc, "Audio mixer");
			}
			if (total > 0 && !fused)
			{
				std::vector<Resampled> keep;
				std::vector<const float*> pl, pr;
				std::vector<int64_t> lens;
				std::vector<float> gains((size_t)nin, 1.0f);
				bool any_gain = false;
				for (int i = 0; i < nin; i++)
				{
					int64_t len = 0;
					const Audio_buffer& in = *ins[(size_t)i];
					// 48 kHz stereo float that lands unbroken at 0: swr would only copy (FLTP) or de-interleave (FLT) it;
					// the mix kernel reads it in place -- and through a lazy gain product it reads the gain node's SOURCE
					// and multiplies by the gain itself
					if (in.sample_rate == 48000 && in.channels == 2 && (in.format == FMT_FLT || in.format == FMT_FLTP)
						&& single_front_segment(segs, i, &len) && len <= in.frames)
					{
						const Audio_buffer* from = &in;
						if (in.is_lazy()) { from = in.lazy->source.get(); gains[(size_t)i] = in.lazy->gain; any_gain = true; }
						pl.push_back((const float*)from->plane[0]);
						pr.push_back(in.format == FMT_FLTP ? (const float*)from->plane[1] : nullptr);
						lens.push_back(len);
						continue;
					}
					Resampled r = resample_input(in, produced_of(segs, i), true, "Audio mixer");
					if (!single_front_segment(segs, i, &len)) { r = scatter(r, segs, i, total, "Audio mixer"); len = total; }
					pl.push_back(r.l); pr.push_back(r.r); lens.push_back(len);
					keep.push_back(std::move(r));
				}
				abi(nodey_mix_gains(out_l, out_r, pl.data(), pr.data(), lens.data(), job.vol.data(), any_gain ? gains.data() : nullptr, nin, total,
									cur_stream()), "Audio mixer");
			}
			amix_publish(job, output, block, out_l, out_r);
		}
	}

	void Audio_amix::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>&, std::any&)
	{
		Amix_job job = amix_prepare(input_num, volumes, input);
		job.start_stamps = start_time_stamps;
		amix_execute(job, output);
	}

	// Batch: the single-input mixers of one level that resample the same kind of stream (the per-track audio_amix(1)
	// nodes of a multitrack render) become ONE launch over all their tracks; every other node runs as it would alone.
	bool Audio_amix::process_batch(const std::vector<Batch_item>& items)
	{
		constexpr size_t kMaxTracks = 256;
		std::vector<Amix_job> jobs;
		std::map<std::tuple<int, int, int, int64_t, int64_t, int64_t, const Stream_progress*>, std::vector<size_t>> groups;
		for (size_t k = 0; k < items.size(); k++)
		{
			const auto* node = static_cast<const Audio_amix*>(items[k].processor);
			jobs.push_back(amix_prepare(node->input_num, node->volumes, *items[k].input));
			jobs.back().start_stamps = node->start_time_stamps;
			const Amix_job& job = jobs.back();
			if (job.ins.size() == 1 && job.fused)
			{
				const Audio_buffer& in = *job.ins[0];
				groups[{in.sample_rate, in.format, in.channels, in.frames, job.plan->total, job.front_len[0], in.progress.get()}].push_back(k);
			}
		}
		std::vector<bool> done(items.size(), false);
		for (const auto& [key, members] : groups)
		{
			if (members.size() < 2) continue;
			const auto [rate, fmt, ch, frames, total, front, pg_] = key;
			(void)pg_;
			for (size_t first = 0; first < members.size(); first += kMaxTracks)
			{
				const size_t cnt = std::min(kMaxTracks, members.size() - first);
				const size_t plane = Arena::padded((size_t)std::max<int64_t>(total, 1) * sizeof(float));
				// The batch is cut into launches along time: a chunk starts as soon as the prefix of the input it reads has
				// landed (uploads arrive chunk by chunk) and tells the next node which prefix of the output is final.
				const nodey_resampler* plan = resampler_for(rate);
				const std::shared_ptr<const Stream_progress> in_progress = jobs[members[first]].ins[0]->progress;
				constexpr int kMaxChunks = 64;
				int64_t in_need[kMaxChunks], out_ready[kMaxChunks];
				const int nchunks = nodey_resample_tracks_chunks(plan, ch, frames, total, stream_chunk_count(), in_need, out_ready, kMaxChunks);
				if (nchunks == NODEY_E_RANGE) break;      // no pipelined kernel for this plan: the members run one by one below
				if (nchunks < 0) abi(nchunks, "Audio mixer");
				const nodey_stream_t run = stream_after(in_progress.get());
				Stream_scope scope(run);
				auto block = std::make_shared<infra::Device_block>(2 * plane * cnt);
				float* base = (float*)block->ptr;
				std::vector<const void*> p0(cnt), p1(cnt);
				std::vector<float> vol(cnt);
				for (size_t t = 0; t < cnt; t++)
				{
					const Amix_job& job = jobs[members[first + t]];
					p0[t] = job.ins[0]->plane[0]; p1[t] = job.ins[0]->plane[1]; vol[t] = job.vol[0];
				}
				auto progress = std::make_shared<Stream_progress>();
				progress->stream = run;
				Progress_waiter input{in_progress.get()};
				for (int c = 0; c < nchunks; c++)
				{
					input.need(in_need[c], run);
					// track t: left plane at base + 2t * plane, right plane one plane further
					abi(nodey_resample_tracks_chunk(plan, base, base + plane / sizeof(float), (int64_t)(2 * plane / sizeof(float)), p0.data(), p1.data(),
													fmt, ch, frames, vol.data(), (int)cnt, 1, front, total, c, nchunks, run), "Audio mixer");
					auto ev = std::make_shared<infra::Device_event>();
					ev->record(run);
					progress->points.push_back({out_ready[c], ev});
				}
				input.all(run);
				if (in_progress) { auto ev = std::make_shared<infra::Device_event>(); ev->record(run); progress->points.back().event = ev; }
				for (size_t t = 0; t < cnt; t++)
				{
					const size_t k = members[first + t];
					float* out_l = base + 2 * t * (plane / sizeof(float));
					amix_publish(jobs[k], *items[k].output, block, out_l, out_l + plane / sizeof(float), nchunks > 1 || in_progress ? progress : nullptr);
					if (!(nchunks > 1 || in_progress)) { /* published with a ready event recorded by publish() on `run` */ }
					done[k] = true;
				}
			}
		}
		for (size_t k = 0; k < items.size(); k++)
			if (!done[k]) amix_execute(jobs[k], *items[k].output);
		return true;
	}

	// ---------------------------------------------------------------------------------------------
	// audio_bimix  (audio-bimix.cpp:83-331)
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_bimix::get_processor_info()
	{
		return {.identifier = "audio_bimix", .display_name = "Audio Bimix", .singleton = false, .generate = std::make_unique<Audio_bimix>,
				.description = "Builds a stereo stream from a left and a right input: each side is converted to 48 kHz stereo and folded to "
							   "one channel; 'bias' (-1..1) shifts the balance."};
	}
	std::vector<Processor::Pin_attribute> Audio_bimix::get_pin_attributes() const
	{
		return {Pin_type::audio("output", "Output", false), Pin_type::audio("input_l", "Left", true), Pin_type::audio("input_r", "Right", true)};
	}
	Json::Value Audio_bimix::serialize() const
	{
		Json::Value value;
		value["bias"] = bias;
		if (start_time_stamps) value["start_time_stamps"] = true;
		return value;
	}
	void Audio_bimix::deserialize(const Json::Value& value)
	{
		if (!value.isMember("bias") || !value["bias"].isDouble())
			throw Runtime_error("Failed to deserialize JSON file",
								"Audio_bimix failed to serialize the JSON input because of missing or invalid fields.", "Wrong field: bias");
		bias = std::clamp<float>((float)value["bias"].asDouble(), -1, 1);
		start_time_stamps = value.isMember("start_time_stamps") && value["start_time_stamps"].isBool() && value["start_time_stamps"].asBool();
	}

	namespace
	{
		// frame cursor over run-length encoded frame sizes
		struct Frame_cursor
		{
			const Frame_runs* runs; size_t run = 0; int64_t left = 0;
			explicit Frame_cursor(const Frame_runs& r) : runs(&r) { settle(); }
			void settle()
			{
				while (run < runs->size() && ((*runs)[run].first <= 0 || (*runs)[run].second <= 0)) run++;
				left = run < runs->size() ? (*runs)[run].second : 0;
			}
			bool has() const { return run < runs->size(); }
			int64_t size() const { return (*runs)[run].first; }
			void next() { if (--left == 0) { run++; settle(); } }
		};

		// swr_convert bookkeeping for one input
		struct Swr_state
		{
			const nodey_resampler* plan; int64_t n_in = 0, produced = 0, reflect = 0; bool flushed = false;
			int64_t feed(int64_t frames, int64_t out_cap)
			{
				n_in += frames;
				return take(out_cap);
			}
			int64_t flush(int64_t out_cap)
			{
				if (!flushed) { flushed = true; reflect = nodey_resampler_flush_reflect(plan, n_in, produced); }
				return take(out_cap);
			}
			int64_t take(int64_t out_cap)
			{
				int64_t got = nodey_resampler_producible(plan, n_in, reflect) - produced;
				got = std::clamp<int64_t>(got, 0, out_cap);
				produced += got;
				return got;
			}
		};

		void add_segment(std::vector<Segment>& segs, int input, int64_t out_start, int64_t src_start, int64_t len)
		{
			if (len <= 0) return;
			for (auto it = segs.rbegin(); it != segs.rend(); ++it)
				if (it->input == input)
				{
					if (it->out_start + it->len == out_start && it->src_start + it->len == src_start) { it->len += len; return; }
					break;
				}
			segs.push_back({input, out_start, src_start, len});
		}

		void add_run(Frame_runs& runs, int64_t len)
		{
			if (!runs.empty() && runs.back().first == len) runs.back().second++;
			else runs.emplace_back(len, 1);
		}
	}

	void Audio_bimix::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>&, std::any&)
	{
		auto in_l = require_input(input, "input_l", "Audio Channel mix processor", true, false, "Audio channel mix processor", "input");
		auto in_r = require_input(input, "input_r", "Audio Channel mix processor", true, false, "Audio channel mix processor", "input");
		check_channels(*in_l, "Audio bimix");
		check_channels(*in_r, "Audio bimix");

		// the reference loop, counts only (audio-bimix.cpp:137-329), including its two slips: with both
		// sides present nb ends up 1152 (:178-183), and the right-hand flush count lands in the left
		// counter (:294, App. C5)
		Frame_cursor cl(in_l->runs), cr(in_r->runs);
		Swr_state sl{resampler_for(in_l->sample_rate)}, sr{resampler_for(in_r->sample_rate)};
		std::vector<Segment> segs;
		Frame_runs out_runs;
		int64_t written = 0, max_frame = 1152;
		for (const auto& [len, c] : in_l->runs) max_frame = std::max(max_frame, len);
		for (const auto& [len, c] : in_r->runs) max_frame = std::max(max_frame, len);
		for (;;)
		{
			const bool has_l = cl.has(), has_r = cr.has();
			int64_t nb = 0;
			if (has_r && has_l) nb = std::min(cr.size(), cl.size());
			if (!has_r && has_l) nb = cl.size();
			else if (has_r && !has_l) nb = cr.size();
			else nb = 1152;
			nb = std::min(nb, max_frame);
			int64_t count_l = 0, count_r = 0;
			if (has_l) { const int64_t n = cl.size(); cl.next(); const int64_t at = sl.produced; count_l = sl.feed(n, nb); add_segment(segs, 0, written, at, count_l); }
			else { const int64_t at = sl.produced; count_l = sl.flush(nb); add_segment(segs, 0, written, at, count_l); }
			if (has_r) { const int64_t n = cr.size(); cr.next(); const int64_t at = sr.produced; count_r = sr.feed(n, nb); add_segment(segs, 1, written, at, count_r); }
			else { const int64_t at = sr.produced; const int64_t got = sr.flush(nb); add_segment(segs, 1, written, at, got); count_l = got; }
			add_run(out_runs, nb);
			written += nb;
			if (count_r == 0 && count_l == 0) break;
		}
		const int64_t total = written;
		Resampled rl = resample_input(*in_l, sl.produced, sl.flushed, "Audio bimix");
		Resampled rr = resample_input(*in_r, sr.produced, sr.flushed, "Audio bimix");
		int64_t len_l = 0, len_r = 0;
		if (!single_front_segment(segs, 0, &len_l)) { rl = scatter(rl, segs, 0, total, "Audio bimix"); len_l = total; }
		if (!single_front_segment(segs, 1, &len_r)) { rr = scatter(rr, segs, 1, total, "Audio bimix"); len_r = total; }

		const size_t plane = Arena::padded((size_t)std::max<int64_t>(total, 1) * sizeof(float));
		auto block = std::make_shared<infra::Device_block>(2 * plane);
		float* out_l = (float*)block->ptr;
		float* out_r = (float*)((char*)block->ptr + plane);
		if (total > 0) abi(nodey_bimix(out_l, out_r, rl.l, rl.r, len_l, rr.l, rr.r, len_r, bias, total, cur_stream()), "Audio bimix");
		// App. C3: the running time starts at 0; C4: frames carry their END time
		auto mixed = new_buffer(block, out_l, out_r, FMT_FLTP, 48000, 2, total, std::move(out_runs), 0.0);
		publish(output, "output", start_time_stamps ? mixed : end_time_stamps(std::move(mixed)));
	}

	// ---------------------------------------------------------------------------------------------
	// audio_bimix_v2  (audio-bimix.cpp:475-877)
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_bimix_v2::get_processor_info()
	{
		return {.identifier = "audio_bimix_v2", .display_name = "Audio Bimix V2", .singleton = false, .generate = std::make_unique<Audio_bimix_v2>,
				.description = "Stereo merge with time alignment: both inputs are converted to 48 kHz, folded to mono and placed on the left / "
							   "right channel according to their timestamps; gaps are filled with silence. Output: float interleaved."};
	}
	std::vector<Processor::Pin_attribute> Audio_bimix_v2::get_pin_attributes() const
	{
		return {Pin_type::audio("output", "Output", false), Pin_type::audio("input_l", "Left", true), Pin_type::audio("input_r", "Right", true)};
	}

	void Audio_bimix_v2::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>&, std::any&)
	{
		std::shared_ptr<const Audio_buffer> in[2] = {require_input(input, "input_l", "Audio Channel mix processor", true, false, "Audio channel mix processor", "input"),
													   require_input(input, "input_r", "Audio Channel mix processor", true, false, "Audio channel mix processor", "input")};
		check_channels(*in[0], "Audio bimix v2", Node_kind::bimix_v2);
		check_channels(*in[1], "Audio bimix v2", Node_kind::bimix_v2);

		// per side: list of resampled frames {source offset, length, END time of the block (App. C12)}
		struct Piece { int64_t src, n; double t; };
		struct Side { std::vector<Piece> list; size_t head = 0; Frame_cursor cur; Swr_state swr; double time; bool eof = false;
					  bool empty() const { return head >= list.size(); } Piece& front() { return list[head]; } void pop() { head++; } };
		Side side[2] = {{{}, 0, Frame_cursor(in[0]->runs), Swr_state{resampler_for(in[0]->sample_rate)}, in[0]->pts_seconds},
						{{}, 0, Frame_cursor(in[1]->runs), Swr_state{resampler_for(in[1]->sample_rate)}, in[1]->pts_seconds}};
		struct Out_seg { int64_t out_start, len, l, r; };
		std::vector<Out_seg> segs;
		Frame_runs out_runs;
		int64_t written = 0;
		bool have_pts = false;
		double pts0 = 0;
		const auto emit = [&](int64_t frames, int64_t l_src, int64_t r_src, double t) {
			if (!have_pts) { have_pts = true; pts0 = t; }
			if (frames <= 0) { return; }
			if (!segs.empty())
			{
				Out_seg& b = segs.back();
				const bool l_ok = (b.l < 0 && l_src < 0) || (b.l >= 0 && l_src == b.l + b.len);
				const bool r_ok = (b.r < 0 && r_src < 0) || (b.r >= 0 && r_src == b.r + b.len);
				if (l_ok && r_ok) { b.len += frames; written += frames; add_run(out_runs, frames); return; }
			}
			segs.push_back({written, frames, l_src, r_src});
			written += frames;
			add_run(out_runs, frames);
		};
		const auto feed = [&](Side& s) {
			const int64_t n = s.cur.size();
			s.cur.next();
			const int64_t at = s.swr.produced;
			const int64_t got = s.swr.feed(n, 2 * n);       // swr_convert(out 2n, in n), never flushed
			s.time += (double)got / 48000;
			s.list.push_back({at, got, s.time});
		};
		const auto drop = [](Piece& p, int64_t count) { p.src += count; p.n -= count; p.t += (double)count / 48000; };

		for (;;)
		{
			for (Side& s : side)
				if (!s.eof) { if (s.cur.has()) feed(s); else s.eof = true; }
			Side &L = side[0], &R = side[1];
			if (L.empty() && R.empty() && L.eof && R.eof) break;
			if (R.empty() && R.eof) { if (L.empty()) continue; const Piece f = L.front(); emit(f.n, f.src, -1, f.t); L.pop(); continue; }
			if (L.empty() && L.eof) { if (R.empty()) continue; const Piece f = R.front(); emit(f.n, -1, f.src, f.t); R.pop(); continue; }
			while (!L.empty() && !R.empty())
			{
				const bool left_earlier = L.front().t < R.front().t;
				Side& es = left_earlier ? L : R;
				Side& ls = left_earlier ? R : L;
				const double eb = es.front().t, lb = ls.front().t;
				const double ee = eb + (double)es.front().n / 48000, le = lb + (double)ls.front().n / 48000;
				const auto lr = [&](int64_t e_src, int64_t l_src, int64_t* l_out, int64_t* r_out) {
					*l_out = left_earlier ? e_src : l_src; *r_out = left_earlier ? l_src : e_src; };
				if (ee <= lb)
				{
					int64_t a, b; lr(es.front().src, -1, &a, &b);
					emit(es.front().n, a, b, eb);
					es.pop();
					continue;
				}
				const double fe = ee < le ? ee : le;
				const int64_t un = (int64_t)std::round((lb - eb) * 48000);
				int64_t al = (int64_t)std::round((fe - lb) * 48000);
				{ const uint64_t room = (uint64_t)es.front().n - (uint64_t)un; if ((uint64_t)al > room) al = (int64_t)room; }
				if (al > ls.front().n) al = ls.front().n;
				const int64_t e_src = es.front().src, l_src = ls.front().src;
				if (ee <= le) { es.pop(); drop(ls.front(), al); }
				else { ls.pop(); drop(es.front(), un + al); }
				if (!es.empty() && es.front().n == 0) es.pop();
				if (!ls.empty() && ls.front().n == 0) ls.pop();
				// one emitted block of un + al frames: earlier side alone, then both
				if (!have_pts) { have_pts = true; pts0 = eb; }
				{
					int64_t a, b;
					const size_t before = out_runs.size(); (void)before;
					lr(e_src, -1, &a, &b);
					Frame_runs scratch;
					std::swap(scratch, out_runs);              // both parts belong to ONE reference frame
					emit(un, a, b, eb);
					lr(e_src + un, l_src, &a, &b);
					emit(al, a, b, eb);
					std::swap(scratch, out_runs);
					add_run(out_runs, un + al);
				}
			}
		}

		// device work: swr (no flush) + mono fold per side, then one gather into the interleaved output
		float* mono[2] = {nullptr, nullptr};
		std::shared_ptr<infra::Device_block> mono_block[2];
		for (int s = 0; s < 2; s++)
		{
			const int64_t n = side[s].swr.produced;
			Resampled rs = resample_input(*in[s], n, false, "Audio bimix v2");
			mono_block[s] = std::make_shared<infra::Device_block>((size_t)std::max<int64_t>(n, 1) * sizeof(float));
			mono[s] = (float*)mono_block[s]->ptr;
			if (n > 0) abi(nodey_downmix_half(mono[s], rs.l, rs.r, n, cur_stream()), "Audio bimix v2");
		}
		auto block = std::make_shared<infra::Device_block>((size_t)std::max<int64_t>(written, 1) * 2 * sizeof(float));
		if (!segs.empty())
		{
			std::vector<int64_t> o, n, l, r;
			for (const Out_seg& s : segs) { o.push_back(s.out_start); n.push_back(s.len); l.push_back(s.l); r.push_back(s.r); }
			abi(nodey_merge_segments((float*)block->ptr, mono[0], mono[1], o.data(), n.data(), l.data(), r.data(), (int)segs.size(), cur_stream()),
				"Audio bimix v2");
		}
		publish(output, "output", new_buffer(block, block->ptr, nullptr, FMT_FLT, 48000, 2, written, std::move(out_runs), pts0));
	}

	// ---------------------------------------------------------------------------------------------
	// audio_channel_split (new)
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_channel_split::get_processor_info()
	{
		return {.identifier = "audio_channel_split", .display_name = "Channel Split", .singleton = false,
				.generate = std::make_unique<Audio_channel_split>,
				.description = "Routes the left and right channel of a stereo stream to two mono streams of the same sample type (bit exact)."};
	}
	std::vector<Processor::Pin_attribute> Audio_channel_split::get_pin_attributes() const
	{
		return {Pin_type::audio("output_l", "Left", false), Pin_type::audio("output_r", "Right", false), Pin_type::audio("input", "Input", true)};
	}
	void Audio_channel_split::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>&, std::any&)
	{
		auto in = require_input(input, "input", "Channel split");
		check_channels(*in, "Channel split");
		if (in->channels != 2)
			throw Runtime_error("Channel split needs a stereo stream", "The input of the channel split node has one channel only.",
								std::format("channels: {}", in->channels));
		const size_t bytes = (size_t)in->frames * (size_t)format_bytes(in->format);
		Arena arena(2 * Arena::padded(bytes));
		void* l = arena.take(bytes);
		void* r = arena.take(bytes);
		if (in->frames > 0) abi(nodey_split(l, r, in->plane[0], in->plane[1], in->format, in->frames, cur_stream()), "Channel split");
		auto ev = std::make_shared<infra::Device_event>();
		ev->record(cur_stream());
		auto bl = inherit_stamps(new_buffer(arena.block, l, nullptr, in->format, in->sample_rate, 1, in->frames, in->runs, in->pts_seconds), *in);
		auto br = inherit_stamps(new_buffer(arena.block, r, nullptr, in->format, in->sample_rate, 1, in->frames, in->runs, in->pts_seconds), *in);
		bl->ready = ev; br->ready = ev;
		publish(output, "output_l", bl);
		publish(output, "output_r", br);
	}

	// ---------------------------------------------------------------------------------------------
	// audio_spectrum (new)
	// ---------------------------------------------------------------------------------------------
	infra::Processor::Info Audio_spectrum::get_processor_info()
	{
		return {.identifier = "audio_spectrum", .display_name = "Spectrum", .singleton = false, .generate = std::make_unique<Audio_spectrum>,
				.description = "Short-time Fourier transform per channel: 4096-point periodic Hann window, hop 1024, unnormalised forward DFT "
							   "(FFTW r2c convention), 2049 complex bins per frame."};
	}
	std::vector<Processor::Pin_attribute> Audio_spectrum::get_pin_attributes() const
	{
		return {Processor::Pin_attribute{.identifier = "output", .display_name = "Spectrum", .type = typeid(Spectrum_stream), .is_input = false,
										 .generate_func = [] { return std::make_shared<Spectrum_stream>(); }},
				Pin_type::audio("input", "Input", true)};
	}
	Json::Value Audio_spectrum::serialize() const
	{
		Json::Value value;
		value["fft_size"] = fft_size;
		value["hop"] = hop;
		value["window"] = window;
		return value;
	}
	void Audio_spectrum::deserialize(const Json::Value& value)
	{
		if (value.isMember("fft_size") && value["fft_size"].isInt()) fft_size = value["fft_size"].asInt();
		if (value.isMember("hop") && value["hop"].isInt()) hop = value["hop"].asInt();
		if (value.isMember("window") && value["window"].isString()) window = value["window"].asString();
	}
	void Audio_spectrum::process_payload(const Input_map& input, const Output_map& output, const std::atomic<bool>&, std::any&)
	{
		auto in = require_input(input, "input", "Spectrum");
		check_channels(*in, "Spectrum");
		if (window != "hann")
			throw Runtime_error("Unsupported window", "The spectrum node implements the periodic Hann window only.", window);
		const float* x = (const float*)in->plane[0];
		int interleaved = in->format == FMT_FLT ? 1 : 0;
		int64_t plane_stride = in->format == FMT_FLTP && in->channels == 2 ? (const float*)in->plane[1] - (const float*)in->plane[0] : 0;
		std::unique_ptr<infra::Device_block> converted;
		if (in->format != FMT_FLT && !(in->format == FMT_FLTP && (in->channels == 1 || plane_stride > 0)))
		{
			converted = std::make_unique<infra::Device_block>((size_t)std::max<int64_t>(in->frames, 1) * (size_t)in->channels * sizeof(float));
			abi(nodey_extract_interleaved((float*)converted->ptr, in->plane[0], in->plane[1], in->format, in->frames, in->channels, cur_stream()), "Spectrum");
			x = (const float*)converted->ptr;
			interleaved = 1;
			plane_stride = 0;
		}
		const int64_t m = nodey_stft_frames(in->frames, fft_size, hop);
		auto spec = std::make_shared<Spectrum_buffer>();
		spec->channels = in->channels; spec->fft_size = fft_size; spec->hop = hop; spec->sample_rate = in->sample_rate; spec->frames = m;
		spec->block = std::make_shared<infra::Device_block>((size_t)std::max<int64_t>(m, 1) * (size_t)in->channels * (size_t)(fft_size / 2 + 1) * 2 * sizeof(float));
		spec->data = (float*)spec->block->ptr;
		abi(nodey_stft(spec->data, x, in->frames, in->channels, interleaved, plane_stride, fft_size, hop, cur_stream()), "Spectrum");
		spec->ready = std::make_shared<infra::Device_event>();
		spec->ready->record(cur_stream());
		{
			std::lock_guard lock(result_mutex);
			result = spec;
		}
		if (const auto find = output.find("output"); find != output.end())
			for (auto& item : find->second) std::dynamic_pointer_cast<Spectrum_stream>(item)->publish(spec);
	}
}

namespace infra
{
	// reference: src/register.cpp:14-23 (eight nodes) + the two new ones
	void register_all_processors()
	{
		static std::once_flag once;
		std::call_once(once, [] {
			Processor::register_processor<processor::Audio_input>();
			Processor::register_processor<processor::Audio_output>();
			Processor::register_processor<processor::Audio_vol>();
			Processor::register_processor<processor::Velocity_modifier>();
			Processor::register_processor<processor::Pitch_modifier>();
			Processor::register_processor<processor::Audio_amix>();
			Processor::register_processor<processor::Audio_bimix>();
			Processor::register_processor<processor::Audio_bimix_v2>();
			Processor::register_processor<processor::Audio_channel_split>();
			Processor::register_processor<processor::Audio_spectrum>();
		});
	}
}
