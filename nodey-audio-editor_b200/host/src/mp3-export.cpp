// mp3-export.cpp -- the MP3 leg of the audio_output sink.
//
// Reference: Audio_output::do_export (src/processor/audio-io.cpp:640-841) feeds every frame that arrives at
// the sink to LAME in the frame's own sample format and writes what the encoder returns.  Here the rendered
// stream is one device buffer; it is downloaded once and walked in the frame sizes its producer recorded, so
// LAME sees the call sequence the reference would have issued: parameters on the first frame
// (:805-822), `(int)((frame_begin - time) * rate)` samples of 16-bit silence in front of a frame that starts
// late (:664-693, :829-834), then the frame through the entry point of its format (:695-777), buffer sizes
// `4 * nb + 7200` resp. `1.25 * n + 7200` bytes, no final lame_encode_flush (the reference never calls it:
// the encoder's last partial MP3 frame is dropped there too).
//
// libmp3lame is bound at run time (dlopen of libmp3lame.so.0, NODEY_LAME_LIB overrides) because the library
// is an optional codec, absent from this image: without it an MP3 export is a Runtime_error that says so,
// the float WAV export (*.wav) and the in-memory result do not need it.
//
// Deliberate divergences from the reference's encode switch: its S16P and S32P cases lack a
// `break` and fall through into the next case, which encodes plane 0 a second time as interleaved data of
// another type; and its interleaved entry points are also used for mono frames, which reads past the frame.
// Both are undefined reads in the reference; here planar frames go to the planar entry point only and mono
// frames hand their single plane as both channels (LAME ignores the right one when num_channels == 1).
#include "processor/nodes.hpp"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <format>
#include <fstream>
#include <mutex>

namespace processor
{
	namespace
	{
		using lame_t = void*;
		// lame.h: MPEG_mode { STEREO = 0, JOINT_STEREO, DUAL_CHANNEL, MONO }, vbr_mode { vbr_off = 0, ... }
		constexpr int kModeStereo = 0, kModeMono = 3, kVbrOff = 0;

		struct Lame_api
		{
			void* handle = nullptr;
			std::string path, why;
			lame_t (*init)() = nullptr;
			int (*close)(lame_t) = nullptr;
			int (*set_in_samplerate)(lame_t, int) = nullptr;
			int (*set_num_channels)(lame_t, int) = nullptr;
			int (*set_quality)(lame_t, int) = nullptr;
			int (*set_mode)(lame_t, int) = nullptr;
			int (*set_out_samplerate)(lame_t, int) = nullptr;
			int (*set_VBR)(lame_t, int) = nullptr;
			int (*set_brate)(lame_t, int) = nullptr;
			int (*init_params)(lame_t) = nullptr;
			int (*encode_interleaved)(lame_t, short*, int, unsigned char*, int) = nullptr;
			int (*encode)(lame_t, const short*, const short*, int, unsigned char*, int) = nullptr;
			int (*encode_interleaved_int)(lame_t, const int*, int, unsigned char*, int) = nullptr;
			int (*encode_int)(lame_t, const int*, const int*, int, unsigned char*, int) = nullptr;
			int (*encode_interleaved_float)(lame_t, const float*, int, unsigned char*, int) = nullptr;
			int (*encode_float)(lame_t, const float*, const float*, int, unsigned char*, int) = nullptr;
		};

		std::mutex g_lame_mutex;
		Lame_api g_lame;

		// (re)binds when NODEY_LAME_LIB names another library than the one bound (tests swap in a recording stub)
		const Lame_api& lame_api()
		{
			std::lock_guard lock(g_lame_mutex);
			const char* env = getenv("NODEY_LAME_LIB");
			const std::string want = (env && *env) ? env : "libmp3lame.so.0";
			if (g_lame.path == want && (g_lame.handle || !g_lame.why.empty())) return g_lame;
			if (g_lame.handle) dlclose(g_lame.handle);
			g_lame = Lame_api{};
			g_lame.path = want;
			g_lame.handle = dlopen(want.c_str(), RTLD_NOW | RTLD_LOCAL);
			if (!g_lame.handle && !(env && *env)) g_lame.handle = dlopen("libmp3lame.so", RTLD_NOW | RTLD_LOCAL);
			if (!g_lame.handle)
			{
				const char* e = dlerror();
				g_lame.why = e ? e : "dlopen failed";
				return g_lame;
			}
			bool ok = true;
			const auto sym = [&](const char* name) -> void*
			{
				void* p = dlsym(g_lame.handle, name);
				if (!p && ok) { ok = false; g_lame.why = std::format("{} lacks {}", want, name); }
				return p;
			};
#define NODEY_LAME_BIND(field, name) g_lame.field = reinterpret_cast<decltype(g_lame.field)>(sym(name))
			NODEY_LAME_BIND(init, "lame_init");
			NODEY_LAME_BIND(close, "lame_close");
			NODEY_LAME_BIND(set_in_samplerate, "lame_set_in_samplerate");
			NODEY_LAME_BIND(set_num_channels, "lame_set_num_channels");
			NODEY_LAME_BIND(set_quality, "lame_set_quality");
			NODEY_LAME_BIND(set_mode, "lame_set_mode");
			NODEY_LAME_BIND(set_out_samplerate, "lame_set_out_samplerate");
			NODEY_LAME_BIND(set_VBR, "lame_set_VBR");
			NODEY_LAME_BIND(set_brate, "lame_set_brate");
			NODEY_LAME_BIND(init_params, "lame_init_params");
			NODEY_LAME_BIND(encode_interleaved, "lame_encode_buffer_interleaved");
			NODEY_LAME_BIND(encode, "lame_encode_buffer");
			NODEY_LAME_BIND(encode_interleaved_int, "lame_encode_buffer_interleaved_int");
			NODEY_LAME_BIND(encode_int, "lame_encode_buffer_int");
			NODEY_LAME_BIND(encode_interleaved_float, "lame_encode_buffer_interleaved_ieee_float");
			NODEY_LAME_BIND(encode_float, "lame_encode_buffer_ieee_float");
#undef NODEY_LAME_BIND
			if (!ok) { dlclose(g_lame.handle); g_lame.handle = nullptr; }
			return g_lame;
		}
	}

	bool mp3_encoder_available(std::string* why)
	{
		const Lame_api& api = lame_api();
		if (why) *why = api.handle ? std::string() : api.why;
		return api.handle != nullptr;
	}

	double export_mp3(const Host_stream& s, const std::string& path, size_t kbps, double time)
	{
		using Runtime_error = infra::Processor::Runtime_error;
		const Lame_api& api = lame_api();
		if (!api.handle)
			throw Runtime_error("MP3 encoder not available",
								"Exporting MP3 needs libmp3lame, which could not be loaded. Install LAME (or point NODEY_LAME_LIB at "
								"libmp3lame.so.0), or export to a .wav path.",
								api.why);
		if (format_bytes(s.format) == 0)
			throw Runtime_error("Unsupported sample format", "The audio sample format is not supported for encoding.",
								std::format("Sample format: {}", s.format));
		// audio-io.cpp:648-654
		std::ofstream output_file(path, std::ios::binary);
		if (!output_file.is_open())
			throw Runtime_error("Failed to open output file", "Cannot open the output file for writing. Check if the path is valid and writable.",
								std::format("Output path: {}", path));
		lame_t lame = api.init();
		if (lame == nullptr) throw std::bad_alloc();
		struct Closer { const Lame_api& api; lame_t lame; ~Closer() { api.close(lame); } } closer{api, lame};

		std::vector<unsigned char> file_buffer;
		const auto write_out = [&](int written, const char* what, const char* explanation)
		{
			if (written < 0) throw Runtime_error(what, explanation, std::format("LAME Error: {}", written));
			if (written > 0) output_file.write(reinterpret_cast<const char*>(file_buffer.data()), written);
		};

		bool lame_param_set = false;
		const int bps = format_bytes(s.format);
		const bool planar = format_is_planar(s.format);
		const size_t stride = planar ? (size_t)bps : (size_t)bps * (size_t)s.channels;     // bytes per frame inside a plane
		// the frames, their own stamps and the silence in front of each (audio-io.cpp:826-839)
		std::vector<Export_step> steps;
		const double end_time = export_steps(s.runs, s.frames, s.clock(), s.sample_rate, time, steps);
		for (const Export_step& step : steps)
			{
				const int nb = step.nb;
				const int64_t at = step.at;
				if (!lame_param_set)      // audio-io.cpp:805-822: parameters come from the first frame
				{
					lame_param_set = true;
					api.set_in_samplerate(lame, s.sample_rate);
					api.set_num_channels(lame, s.channels);
					api.set_quality(lame, 2);
					api.set_mode(lame, s.channels == 2 ? kModeStereo : kModeMono);
					api.set_out_samplerate(lame, 48000);               // config::audio::sample_rate (include/config.hpp:20)
					api.set_VBR(lame, kVbrOff);
					api.set_brate(lame, (int)kbps);
					if (api.init_params(lame) == -1)
						throw Runtime_error("Failed to initialize LAME parameters",
											"Cannot set LAME parameters for encoding. Internal error may have occurred.", "");
				}
				const int silence_samples = step.silence;
				if (silence_samples > 0)      // push_silence, :664-693: always two channels of 16-bit zeros
				{
					const int buffer_size = (int)(1.25 * silence_samples + 7200);
					file_buffer.resize((size_t)buffer_size);
					std::vector<short> silence_buffer((size_t)silence_samples * 2, 0);
					write_out(api.encode_interleaved(lame, silence_buffer.data(), silence_samples, file_buffer.data(), buffer_size),
							  "Failed to encode silence", "Cannot encode silence audio data. Internal error may have occurred.");
				}
				if (nb > 0)
				{
					const int buffer_size = 4 * nb + 7200;
					file_buffer.resize((size_t)buffer_size);
					const unsigned char* p0 = static_cast<const unsigned char*>(s.plane[0]) + (size_t)at * stride;
					const unsigned char* p1 = (planar && s.channels == 2) ? static_cast<const unsigned char*>(s.plane[1]) + (size_t)at * stride : p0;
					const bool interleaved = !planar && s.channels == 2;
					int written = 0;
					switch (s.format)
					{
					case FMT_S16: case FMT_S16P:
						written = interleaved ? api.encode_interleaved(lame, const_cast<short*>(reinterpret_cast<const short*>(p0)), nb, file_buffer.data(), buffer_size)
											  : api.encode(lame, reinterpret_cast<const short*>(p0), reinterpret_cast<const short*>(p1), nb, file_buffer.data(), buffer_size);
						break;
					case FMT_S32: case FMT_S32P:
						written = interleaved ? api.encode_interleaved_int(lame, reinterpret_cast<const int*>(p0), nb, file_buffer.data(), buffer_size)
											  : api.encode_int(lame, reinterpret_cast<const int*>(p0), reinterpret_cast<const int*>(p1), nb, file_buffer.data(), buffer_size);
						break;
					default:   // FMT_FLT, FMT_FLTP
						written = interleaved ? api.encode_interleaved_float(lame, reinterpret_cast<const float*>(p0), nb, file_buffer.data(), buffer_size)
											  : api.encode_float(lame, reinterpret_cast<const float*>(p0), reinterpret_cast<const float*>(p1), nb, file_buffer.data(), buffer_size);
						break;
					}
					write_out(written, "Failed to encode audio frame", "Cannot encode the audio frame. Internal error may have occurred.");
				}
			}
		return end_time;
	}

	double export_steps(const Frame_runs& runs, int64_t frames, Frame_clock clock, int sample_rate, double time, std::vector<Export_step>& steps)
	{
		int64_t at = 0;
		for (const auto& [len, count] : runs)
			for (int64_t k = 0; k < count && at < frames; k++)
			{
				const int nb = (int)std::min<int64_t>(len, frames - at);
				// audio-io.cpp:833-838
				const double frame_begin = clock.next(nb);
				const double frame_end = frame_begin + nb / (double)sample_rate;
				const double silence_time = frame_begin - time;
				const int silence_samples = static_cast<int>(silence_time * sample_rate);      // push_silence, :664-667
				steps.push_back({at, nb, silence_samples > 0 ? silence_samples : 0});
				time = frame_end;
				at += nb;
			}
		return time;
	}
}
