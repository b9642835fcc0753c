// processor/nodes.hpp -- the node classes.  Class names, identifiers, pin identifiers and JSON keys are
// the reference's; the two nodes the reference's README promises but never
// shipped (channel split, spectrum) are added through the same API.
//   audio_input          Audio_input        include/processor/audio-io.hpp     src/processor/audio-io.cpp:29-338
//   audio_output         Audio_output       include/processor/audio-io.hpp     src/processor/audio-io.cpp:430-844
//   audio_volume_adjust  Audio_vol          include/processor/audio-vol.hpp    src/processor/audio-vol.cpp
//   velocity_modifier    Velocity_modifier  include/processor/audio-velocity.hpp src/processor/audio-velocity.cpp
//   pitch_modifier       Pitch_modifier     (same files)
//   audio_amix           Audio_amix         include/processor/audio-amix.hpp   src/processor/audio-amix.cpp
//   audio_bimix          Audio_bimix        include/processor/audio-bimix.hpp  src/processor/audio-bimix.cpp:26-383
//   audio_bimix_v2       Audio_bimix_v2     (same files, :387-877)
//   audio_channel_split  Audio_channel_split  NEW
//   audio_spectrum       Audio_spectrum       NEW
#pragma once

#include "infra/processor.hpp"
#include "processor/audio-stream.hpp"

#include <functional>
#include <mutex>
#include <string>
#include <vector>

namespace processor
{
#define NODEY_NODE_COMMON(Class)                                                                          \
	Class(const Class&) = delete;                                                                         \
	Class& operator=(const Class&) = delete;                                                              \
	static infra::Processor::Info get_processor_info();                                                   \
	virtual Processor::Info get_processor_info_non_static() const { return get_processor_info(); }        \
	virtual std::vector<infra::Processor::Pin_attribute> get_pin_attributes() const;                      \
	virtual void process_payload(                                                                         \
		const std::map<std::string, std::shared_ptr<infra::Processor::Product>>& input,                   \
		const std::map<std::string, std::set<std::shared_ptr<infra::Processor::Product>>>& output,        \
		const std::atomic<bool>& stop_token,                                                              \
		std::any& user_data                                                                               \
	);                                                                                                    \
	virtual void draw_title() {}                                                                          \
	virtual bool draw_content(bool) { return false; }

	// ---- source ------------------------------------------------------------------------------------
	// One PCM source per output pin.  The reference demuxes/decodes files with libavformat/libavcodec
	// (not available here); this engine takes raw PCM handed in through user_data (Pcm_source_list) or
	// reads canonical RIFF/WAVE files named by file_path (PCM 16/24/32 bit, IEEE float 32) with the sample formats and
	// frame sizes libavformat's wav demuxer + PCM decoders would hand out (4096-byte packets: 4096 / block_align frames).
	struct Pcm_source
	{
		const void* data = nullptr;     // packed samples (planar: plane 0), host or device memory
		const void* data1 = nullptr;    // planar channel 1
		bool on_device = false;
		int format = FMT_FLT, sample_rate = 48000, channels = 2;
		int64_t frames = 0;
		int frame_size = 1152;          // decoder frame size the reference would see (1152 MP3, 4096 / block_align WAV, 4096 FLAC)
		double pts_seconds = 0.0;
	};
	struct Pcm_source_list
	{
		std::vector<Pcm_source> sources;    // index i feeds pin output_{i}
	};
	// What audio_input would publish for a RIFF/WAVE file (header only, no device needed): sample format as decoded
	// (24-bit PCM arrives as S32), rate, channels, sample frames and the frame size of the decoder's packets.
	// Throws Processor::Runtime_error("Cannot open audio file", ...) like the source node.
	void probe_wav(const std::string& path, int& format, int& sample_rate, int& channels, int64_t& frames, int& frame_size);

	class Audio_input : public infra::Processor
	{
		size_t file_count = 1;
		std::vector<std::string> file_paths = {""};

	  public:
		Audio_input() = default;
		virtual ~Audio_input() = default;
		NODEY_NODE_COMMON(Audio_input)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
		void set_file_count(size_t n);          // programmatic equivalent of the UI's add/remove file buttons
		virtual size_t upload_bytes(const std::any& user_data) const;
	};

	// ---- sink --------------------------------------------------------------------------------------
	class Audio_output : public infra::Processor
	{
	  public:
		// user_data of the sink.  do_export / export_path / time keep the reference's meaning
		// (audio-io.hpp:62-69); the rendered stream is additionally kept for the caller.
		struct Process_context
		{
			bool do_export = true;
			// "" = keep in memory only; *.wav = interleaved float WAV; any other path = MP3 through LAME like the
			// reference (audio-io.cpp:640-841; libmp3lame bound at run time, Runtime_error when it is absent)
			std::string export_path = "";
			size_t kbps = 0;                                 // MP3 bit rate (the editor passes 64..320, app.cpp:595-608)
			std::shared_ptr<std::atomic<double>> time = std::make_shared<std::atomic<double>>(0.0);
			std::shared_ptr<const Audio_buffer> rendered;    // OUT: what arrived at the sink (device resident)
			// do_export == false selects the reference's preview path (audio-io.cpp:478-638): the stream is brought to
			// 48 kHz stereo float by swr frame by frame WITHOUT a final flush, clamped to [-1, 1] and queued as packed
			// frames.  OUT: that queue content as one packed buffer whose frame runs are the per-frame chunk sizes.
			// preview_sink (optional) plays the role of SDL_QueueAudio: it receives the chunks in order (host memory,
			// packed stereo float) and returns false to stop the preview, like the stop token.
			std::shared_ptr<const Audio_buffer> preview;
			std::function<bool(const float* packed_frames, int64_t frames)> preview_sink;
		};

		Audio_output() = default;
		virtual ~Audio_output() = default;
		NODEY_NODE_COMMON(Audio_output)
		virtual Json::Value serialize() const { return {}; }
		virtual void deserialize(const Json::Value&) {}
	};

	// ---- MP3 leg of the sink (host/src/mp3-export.cpp) ---------------------------------------------------
	// host image of a rendered stream in its own sample format (what the reference's frames would carry)
	struct Host_stream
	{
		int format = FMT_FLT, sample_rate = 48000, channels = 2;
		int64_t frames = 0;
		double pts_seconds = 0.0;
		Frame_runs runs;
		const void* plane[2] = {nullptr, nullptr};
		int stamp = STAMP_START;            // how the producer stamped its frames (Audio_buffer::stamp / stamp_origin)
		double stamp_origin = 0.0;
		std::shared_ptr<const std::vector<double>> frame_pts;
		Frame_clock clock() const
		{
			const bool from_first = stamp == STAMP_START || stamp == STAMP_LIST;
			return Frame_clock(stamp, from_first ? pts_seconds : stamp_origin, sample_rate, frame_pts);
		}
	};
	// Audio_output::do_export's bookkeeping over the frames of a stream (audio-io.cpp:826-839): before a frame,
	// (int)((frame_begin - time) * sample_rate) samples of silence when that is positive; afterwards time = frame_begin +
	// nb / rate.  frame_begin is the frame's own stamp (Frame_clock).  Returns `time` after the last frame.
	struct Export_step { int64_t at; int nb; int silence; };
	double export_steps(const Frame_runs& runs, int64_t frames, Frame_clock clock, int sample_rate, double time, std::vector<Export_step>& steps);
	// true when libmp3lame could be bound (NODEY_LAME_LIB overrides the library name); why = the loader's message
	bool mp3_encoder_available(std::string* why = nullptr);
	// Audio_output::do_export's LAME call sequence over the stream's frames; `time` is Process_context::time going in,
	// the return value what it is afterwards (end of the last frame).  Throws Processor::Runtime_error like the reference.
	double export_mp3(const Host_stream& stream, const std::string& path, size_t kbps, double time);

	// ---- gain --------------------------------------------------------------------------------------
	class Audio_vol : public infra::Processor
	{
		float volume = 1.0;

	  public:
		Audio_vol() = default;
		virtual ~Audio_vol() = default;
		NODEY_NODE_COMMON(Audio_vol)
		virtual Json::Value serialize() const { return {}; }          // the reference never persists the gain (App. C1)
		virtual void deserialize(const Json::Value& value);           // reads an optional "volume" key if present
		virtual bool process_batch(const std::vector<Batch_item>& items);
		void set_volume(float v);                                      // clamped to [0, 10] like the UI slider
		float get_volume() const { return volume; }
	};

	// ---- SoundTouch nodes --------------------------------------------------------------------------
	class Velocity_modifier : public infra::Processor
	{
		float velocity = 1;
		bool keep_pitch = false;
		// SURVEY.md App. C7 switch (optional JSON key "reference_schedule", default off; NODEY_REFERENCE_SCHEDULE=1 turns it on
		// for every node): emit what the reference's node loop emits -- its receive sizes as frame sizes and NO flushed tail
		// when the loop leaves through its early break (audio-velocity.cpp:414) -- instead of the canonical 1152-sample
		// frames of the complete, always flushed render
		bool reference_schedule = false;

	  public:
		Velocity_modifier() = default;
		virtual ~Velocity_modifier() = default;
		NODEY_NODE_COMMON(Velocity_modifier)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
		virtual bool process_batch(const std::vector<Batch_item>& items);
		friend struct Soundtouch_params;
	};

	class Pitch_modifier : public infra::Processor
	{
		float pitch = 0;
		bool reference_schedule = false;     // see Velocity_modifier

	  public:
		Pitch_modifier() = default;
		virtual ~Pitch_modifier() = default;
		NODEY_NODE_COMMON(Pitch_modifier)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
		virtual bool process_batch(const std::vector<Batch_item>& items);
		friend struct Soundtouch_params;
	};

	// ---- mixers ------------------------------------------------------------------------------------
	class Audio_amix : public infra::Processor
	{
		int input_num = 2;
		std::vector<float> volumes;
		std::vector<bool> locks;

	  public:
		// SURVEY.md App. C4 switch (optional JSON key "start_time_stamps", default off): the reference stamps the mixer's
		// frames with their END time, truncated to microseconds (audio-amix.cpp:199-201), which makes its export start with
		// almost one frame of silence and insert more wherever the frame size grows.  Set, the frames carry exact START
		// times from 0 and an export is the mix and nothing else.  Samples are the same either way.
		bool start_time_stamps = false;
		Audio_amix();
		virtual ~Audio_amix() = default;
		NODEY_NODE_COMMON(Audio_amix)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
		virtual bool process_batch(const std::vector<Batch_item>& items);
	};

	class Audio_bimix : public infra::Processor
	{
		float bias = 0.0f;

	  public:
		bool start_time_stamps = false;      // see Audio_amix (audio-bimix.cpp:188-191 stamps the same way)
		Audio_bimix() = default;
		virtual ~Audio_bimix() = default;
		NODEY_NODE_COMMON(Audio_bimix)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
	};

	class Audio_bimix_v2 : public infra::Processor
	{
	  public:
		Audio_bimix_v2() = default;
		virtual ~Audio_bimix_v2() = default;
		NODEY_NODE_COMMON(Audio_bimix_v2)
		virtual Json::Value serialize() const { return {}; }
		virtual void deserialize(const Json::Value&) {}
	};

	// ---- new nodes ---------------------------------------------------------------------------------
	class Audio_channel_split : public infra::Processor
	{
	  public:
		Audio_channel_split() = default;
		virtual ~Audio_channel_split() = default;
		NODEY_NODE_COMMON(Audio_channel_split)
		virtual Json::Value serialize() const { return Json::Value(Json::objectValue); }
		virtual void deserialize(const Json::Value&) {}
	};

	// ---- example of a processor written against the reference's FRAME interface --------------------------
	// (frame-streaming compatibility mode): pops host frames, scales them on the CPU the way the
	// reference's change_volume<T> does (audio-vol.cpp:75-100) and pushes them on, like a third-party node that
	// was never ported to device buffers.  Not part of register_all_processors(): the host registers it
	// explicitly (processor::register_example_processors / nodey_engine_register_examples).
	class Frame_gain_example : public infra::Processor
	{
		float volume = 1.0f;

	  public:
		Frame_gain_example() = default;
		virtual ~Frame_gain_example() = default;
		NODEY_NODE_COMMON(Frame_gain_example)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
	};
	void register_example_processors();

	class Audio_spectrum : public infra::Processor
	{
		int fft_size = 4096, hop = 1024;
		std::string window = "hann";
		// The spectrum is usually the end of a branch: with no link on "output" there is no product to
		// publish to (products exist per link), so the last result is also kept on the node for the host.
		mutable std::mutex result_mutex;
		std::shared_ptr<const Spectrum_buffer> result;

	  public:
		std::shared_ptr<const Spectrum_buffer> get_result() const { std::lock_guard lock(result_mutex); return result; }
		Audio_spectrum() = default;
		virtual ~Audio_spectrum() = default;
		NODEY_NODE_COMMON(Audio_spectrum)
		virtual Json::Value serialize() const;
		virtual void deserialize(const Json::Value& value);
	};
}
