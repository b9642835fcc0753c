// engine.cpp -- C facade (include/nodey_engine.h) over Graph / Runner / the node classes.
#include "infra/graph.hpp"
#include "infra/runner.hpp"
#include "processor/nodes.hpp"

#include "nodey_engine.h"

#include <chrono>
#include <cstring>

using namespace infra;
using namespace processor;

struct nodey_engine
{
	Graph graph;
	Pcm_source_list sources;
	std::unique_ptr<Runner> runner;
	std::shared_ptr<std::any> sink_data;
	bool preview = false;
	std::string export_path;
	size_t kbps = 320;        // the editor's default bit rate (src/frontend/app.cpp:595)
	std::vector<int64_t> preview_chunks;     // chunk sizes the preview sink callback received, in order
	Runner::Schedule schedule;               // explicit settings (nodey_engine_set_schedule); automatic where unset
};

namespace
{
	thread_local std::string g_error;

	int fail(int code, const std::string& text)
	{
		g_error = text;
		return code;
	}

	void copy_text(const std::string& s, char* buf, int cap)
	{
		if (!buf || cap <= 0) return;
		const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
		memcpy(buf, s.data(), n);
		buf[n] = 0;
	}

	std::shared_ptr<Processor::Product> find_product(nodey_engine* e, int node_id, const std::string& pin)
	{
		if (!e->runner) return nullptr;
		const auto& res = e->runner->get_processor_resources();
		const auto it = res.find(node_id);
		if (it == res.end()) return nullptr;
		const auto out = it->second->output_payloads.find(pin);
		if (out == it->second->output_payloads.end() || out->second.empty()) return nullptr;
		return *out->second.begin();
	}
}

extern "C" {

const char* nodey_engine_last_error(void) { return g_error.c_str(); }

int nodey_engine_create(nodey_engine** out, const char* project_json)
{
	if (!out || !project_json) return fail(NODEY_ENGINE_E_INVALID, "nodey_engine_create: null argument");
	try
	{
		register_all_processors();
		Json::Value root;
		Json::Reader reader;
		if (!reader.parse(project_json, root)) return fail(NODEY_ENGINE_E_FILE, "Invalid File: " + reader.getFormattedErrorMessages());
		auto e = std::make_unique<nodey_engine>();
		e->graph = Graph::deserialize(root);
		*out = e.release();
		return 0;
	}
	catch (const Graph::Invalid_file_error& err) { return fail(NODEY_ENGINE_E_FILE, err.what()); }
	catch (const Processor::Runtime_error& err) { return fail(NODEY_ENGINE_E_FILE, err.what()); }
	catch (const Graph::Mismatched_pin_error& err) { return fail(NODEY_ENGINE_E_GRAPH, err.what()); }
	catch (const Graph::Multiple_input_error& err) { return fail(NODEY_ENGINE_E_GRAPH, err.what()); }
	catch (const std::exception& err) { return fail(NODEY_ENGINE_E_INVALID, err.what()); }
}

void nodey_engine_destroy(nodey_engine* e) { delete e; }

int nodey_engine_register_examples(void)
{
	try
	{
		register_all_processors();
		processor::register_example_processors();
		return 0;
	}
	catch (const std::exception& err) { return fail(NODEY_ENGINE_E_INVALID, err.what()); }
}

int nodey_engine_serialize(nodey_engine* e, char* buf, int cap)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	const std::string text = Json::writeString(e->graph.serialize(), "  ");
	copy_text(text, buf, cap);
	return (int)text.size();
}

int nodey_engine_check(nodey_engine* e)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	try { e->graph.check_graph(); return 0; }
	catch (const std::exception& err) { return fail(NODEY_ENGINE_E_GRAPH, err.what()); }
}

int nodey_engine_node_count(nodey_engine* e) { return e ? (int)e->graph.nodes.size() : NODEY_ENGINE_E_INVALID; }

int nodey_engine_node_info(nodey_engine* e, int k, int* node_id, char* identifier, int cap, int* level)
{
	if (!e || k < 0 || k >= (int)e->graph.nodes.size()) return fail(NODEY_ENGINE_E_INVALID, "node index out of range");
	auto it = e->graph.nodes.begin();
	std::advance(it, k);
	if (node_id) *node_id = it->first;
	copy_text(it->second.processor->get_processor_info_non_static().identifier, identifier, cap);
	if (level)
	{
		*level = -1;
		try
		{
			const auto levels = e->graph.topological_levels();
			for (size_t l = 0; l < levels.size(); l++)
				for (const Id_t id : levels[l])
					if (id == it->first) *level = (int)l;
		}
		catch (const std::exception&) {}
	}
	return 0;
}

int nodey_engine_set_volume(nodey_engine* e, int node_id, float volume)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	const auto it = e->graph.nodes.find(node_id);
	if (it == e->graph.nodes.end()) return fail(NODEY_ENGINE_E_INVALID, "no such node");
	auto* vol = dynamic_cast<Audio_vol*>(it->second.processor.get());
	if (!vol) return fail(NODEY_ENGINE_E_INVALID, "node is not an audio_volume_adjust");
	vol->set_volume(volume);
	return 0;
}

int nodey_engine_bind_source(nodey_engine* e, int index, const void* data, const void* data1, int on_device, int fmt, int sample_rate,
							 int channels, int64_t frames, int frame_size, double pts_seconds)
{
	if (!e || index < 0 || index > 4096) return fail(NODEY_ENGINE_E_INVALID, "bad source index");
	if ((size_t)index >= e->sources.sources.size()) e->sources.sources.resize((size_t)index + 1);
	Pcm_source& s = e->sources.sources[(size_t)index];
	s.data = data; s.data1 = data1; s.on_device = on_device != 0; s.format = fmt; s.sample_rate = sample_rate; s.channels = channels;
	s.frames = frames; s.frame_size = frame_size > 0 ? frame_size : 1152; s.pts_seconds = pts_seconds;
	return 0;
}

int nodey_engine_run(nodey_engine* e)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	try
	{
		// the environment's development overrides are read here, once per run; what the caller set explicitly wins
		const Runner::Schedule schedule = Runner::Schedule::from_environment().overlay(e->schedule);
		const bool timing = schedule.timing;      // development: host-side phases of a run on stderr
		const auto t0 = std::chrono::steady_clock::now();
		e->runner.reset();
		const auto t1 = std::chrono::steady_clock::now();
		std::map<Id_t, std::shared_ptr<std::any>> node_data;
		if (const auto in = e->graph.singleton_node_map.find("audio_input"); in != e->graph.singleton_node_map.end())
			node_data[in->second] = std::make_shared<std::any>(e->sources);
		if (const auto out = e->graph.singleton_node_map.find("audio_output"); out != e->graph.singleton_node_map.end())
		{
			Audio_output::Process_context ctx;
			ctx.do_export = !e->preview;
			ctx.export_path = e->export_path;
			ctx.kbps = e->kbps;
			e->preview_chunks.clear();
			if (e->preview)
				ctx.preview_sink = [e](const float*, int64_t frames) { e->preview_chunks.push_back(frames); return true; };
			e->sink_data = std::make_shared<std::any>(ctx);
			node_data[out->second] = e->sink_data;
		}
		e->runner = Runner::create_and_run(e->graph, std::move(node_data), schedule);
		const auto t2 = std::chrono::steady_clock::now();
		e->runner->wait();
		if (timing)
		{
			const auto t3 = std::chrono::steady_clock::now();
			const auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
			fprintf(stderr, "[nodey engine] drop previous run %.2f ms, resources + start %.2f ms, run + drain %.2f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
		}
		const std::string err = e->runner->first_error();
		if (!err.empty()) return fail(NODEY_ENGINE_E_NODE, err);
		return 0;
	}
	catch (const Graph::Mismatched_pin_error& err) { return fail(NODEY_ENGINE_E_GRAPH, err.what()); }
	catch (const Graph::Multiple_input_error& err) { return fail(NODEY_ENGINE_E_GRAPH, err.what()); }
	catch (const Graph::Loop_detected_error& err) { return fail(NODEY_ENGINE_E_GRAPH, err.what()); }
	catch (const std::exception& err) { return fail(NODEY_ENGINE_E_INVALID, err.what()); }
}

int nodey_engine_product(nodey_engine* e, int node_id, const char* pin, int* kind, int* fmt, int* sample_rate, int* channels, int64_t* frames,
						 double* pts_seconds, void** plane0, void** plane1, int* extra)
{
	if (!e || !pin) return fail(NODEY_ENGINE_E_INVALID, "null argument");
	auto product = find_product(e, node_id, pin);
	if (!product && e->runner)
	{
		// an unlinked audio_spectrum output: the node keeps its last result for the host
		const auto node = e->graph.nodes.find(node_id);
		if (node != e->graph.nodes.end())
			if (const auto* sp = dynamic_cast<const Audio_spectrum*>(node->second.processor.get()); sp && sp->get_result())
			{
				auto holder = std::make_shared<Spectrum_stream>();
				holder->publish(sp->get_result());
				product = holder;
			}
	}
	if (!product) return fail(NODEY_ENGINE_E_INVALID, "no product on that pin (not linked, or the engine has not run)");
	if (const auto audio = std::dynamic_pointer_cast<Audio_stream>(product))
	{
		const auto b = audio->get();
		if (!b) return fail(NODEY_ENGINE_E_NODE, "stream was closed without audio (or the run released its intermediates)");
		if (b->is_lazy())
		{
			// a gain product nobody materialised (its mixer folded the gain in): run the node's pass now, for the host
			try { processor::materialize(*b, nullptr); }
			catch (const std::exception& err) { return fail(NODEY_ENGINE_E_NODE, err.what()); }
		}
		if (b->ready) b->ready->synchronize();
		if (kind) *kind = 1;
		if (fmt) *fmt = b->format;
		if (sample_rate) *sample_rate = b->sample_rate;
		if (channels) *channels = b->channels;
		if (frames) *frames = b->frames;
		if (pts_seconds) *pts_seconds = b->pts_seconds;
		if (plane0) *plane0 = b->plane[0];
		if (plane1) *plane1 = b->plane[1];
		if (extra) *extra = 0;
		return 0;
	}
	if (const auto spec = std::dynamic_pointer_cast<Spectrum_stream>(product))
	{
		const auto b = spec->get();
		if (!b) return fail(NODEY_ENGINE_E_NODE, "spectrum was not published");
		if (b->ready) b->ready->synchronize();
		if (kind) *kind = 2;
		if (fmt) *fmt = FMT_FLT;
		if (sample_rate) *sample_rate = b->sample_rate;
		if (channels) *channels = b->channels;
		if (frames) *frames = b->frames;
		if (pts_seconds) *pts_seconds = 0;
		if (plane0) *plane0 = b->data;
		if (plane1) *plane1 = nullptr;
		if (extra) *extra = b->fft_size / 2 + 1;
		return 0;
	}
	return fail(NODEY_ENGINE_E_INVALID, "unknown product type");
}

int nodey_engine_set_release_products(int release)
{
	Runner::release_products(release != 0);
	return 0;
}

int nodey_engine_product_runs(nodey_engine* e, int node_id, const char* pin, int64_t* run_len, int64_t* run_count, int cap)
{
	if (!e || !pin) return fail(NODEY_ENGINE_E_INVALID, "null argument");
	const auto audio = std::dynamic_pointer_cast<Audio_stream>(find_product(e, node_id, pin));
	if (!audio || !audio->get()) return fail(NODEY_ENGINE_E_INVALID, "no audio product on that pin");
	const auto& runs = audio->get()->runs;
	for (int k = 0; k < cap && k < (int)runs.size(); k++) { run_len[k] = runs[(size_t)k].first; run_count[k] = runs[(size_t)k].second; }
	return (int)runs.size();
}

int nodey_engine_product_stamp(nodey_engine* e, int node_id, const char* pin, int* stamp, double* origin)
{
	if (!e || !pin) return fail(NODEY_ENGINE_E_INVALID, "null argument");
	const auto audio = std::dynamic_pointer_cast<Audio_stream>(find_product(e, node_id, pin));
	if (!audio || !audio->get()) return fail(NODEY_ENGINE_E_INVALID, "no audio product on that pin");
	const auto b = audio->get();
	if (stamp) *stamp = b->stamp;
	if (origin) *origin = (b->stamp == STAMP_START || b->stamp == STAMP_LIST) ? b->pts_seconds : b->stamp_origin;
	return 0;
}

int nodey_engine_probe_wav(const char* path, int* fmt, int* sample_rate, int* channels, int64_t* frames, int* frame_size)
{
	if (!path) return fail(NODEY_ENGINE_E_INVALID, "nodey_engine_probe_wav: null path");
	try
	{
		int f = 0, r = 0, c = 0, fs = 0;
		int64_t n = 0;
		processor::probe_wav(path, f, r, c, n, fs);
		if (fmt) *fmt = f;
		if (sample_rate) *sample_rate = r;
		if (channels) *channels = c;
		if (frames) *frames = n;
		if (frame_size) *frame_size = fs;
		return 0;
	}
	catch (const Processor::Runtime_error& err) { return fail(NODEY_ENGINE_E_FILE, err.what()); }
	catch (const std::exception& err) { return fail(NODEY_ENGINE_E_INVALID, err.what()); }
}

int nodey_engine_export_plan(int stamp, double origin, int sample_rate, int64_t frames, const int64_t* run_len, const int64_t* run_count,
							 int n_runs, double* time_inout, int64_t* silence, double* frame_pts, int cap)
{
	if (stamp < STAMP_START || stamp > STAMP_START_FLOAT_US || sample_rate < 1 || frames < 0 || n_runs < 0 || (n_runs > 0 && (!run_len || !run_count)))
		return fail(NODEY_ENGINE_E_INVALID, "nodey_engine_export_plan: bad argument");
	Frame_runs runs;
	for (int k = 0; k < n_runs; k++) runs.emplace_back(run_len[k], run_count[k]);
	std::vector<Export_step> steps;
	const double end = export_steps(runs, frames, Frame_clock(stamp, origin, sample_rate), sample_rate, time_inout ? *time_inout : 0.0, steps);
	if (time_inout) *time_inout = end;
	Frame_clock clock(stamp, origin, sample_rate);
	for (size_t k = 0; k < steps.size(); k++)
	{
		const double pts = clock.next(steps[k].nb);
		if ((int)k >= cap) continue;
		if (silence) silence[k] = steps[k].silence;
		if (frame_pts) frame_pts[k] = pts;
	}
	return (int)steps.size();
}

int nodey_engine_output(nodey_engine* e, int* fmt, int* sample_rate, int* channels, int64_t* frames, double* pts_seconds, void** plane0,
						void** plane1)
{
	if (!e || !e->sink_data) return fail(NODEY_ENGINE_E_INVALID, "the graph has no audio_output node or has not run");
	const auto* ctx = std::any_cast<Audio_output::Process_context>(e->sink_data.get());
	if (!ctx || !ctx->rendered) return fail(NODEY_ENGINE_E_NODE, "nothing arrived at audio_output");
	const auto& b = ctx->rendered;
	if (fmt) *fmt = b->format;
	if (sample_rate) *sample_rate = b->sample_rate;
	if (channels) *channels = b->channels;
	if (frames) *frames = b->frames;
	if (pts_seconds) *pts_seconds = b->pts_seconds;
	if (plane0) *plane0 = b->plane[0];
	if (plane1) *plane1 = b->plane[1];
	return 0;
}

int nodey_engine_set_export_path(nodey_engine* e, const char* path)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	e->export_path = path ? path : "";
	return 0;
}

int nodey_engine_diagnostics(nodey_engine* e, char* buf, int cap)
{
	if (!e || !e->runner) return fail(NODEY_ENGINE_E_INVALID, "the engine has not run");
	const std::string text = e->runner->diagnostics_text();
	if (buf && cap > 0)
	{
		const size_t n = std::min(text.size(), (size_t)cap - 1);
		memcpy(buf, text.data(), n);
		buf[n] = 0;
	}
	return (int)text.size();
}

int nodey_engine_level_timings(nodey_engine* e, int* wave, int* level, int* lane, int* nodes, double* enqueue_ms, double* device_ms,
                               double* start_ms, int cap)
{
	if (!e || !e->runner) return fail(NODEY_ENGINE_E_INVALID, "the engine has not run");
	const auto timings = e->runner->get_level_timings();
	for (int k = 0; k < (int)timings.size() && k < cap; k++)
	{
		const auto& t = timings[(size_t)k];
		if (wave) wave[k] = t.wave;
		if (level) level[k] = t.level;
		if (lane) lane[k] = t.lane;
		if (nodes) nodes[k] = (int)t.nodes;
		if (enqueue_ms) enqueue_ms[k] = t.enqueue_ms;
		if (device_ms) device_ms[k] = t.device_ms;
		if (start_ms) start_ms[k] = t.start_ms;
	}
	return (int)timings.size();
}

int nodey_engine_set_schedule(nodey_engine* e, const char* key, const char* value)
{
	if (!e || !key || !value) return fail(NODEY_ENGINE_E_INVALID, "null argument");
	const std::string k = key;
	const int v = atoi(value);
	Runner::Schedule& s = e->schedule;
	if (k == "wave_pins") s.wave_pins = std::max(0, v);
	else if (k == "wave_pattern")
	{
		s.wave_pattern.clear();
		for (const char* c = value; *c;)
		{
			const int size = atoi(c);
			if (size < 1) return fail(NODEY_ENGINE_E_INVALID, "wave_pattern: sizes must be positive, e.g. \"32,64,48\"");
			s.wave_pattern.push_back(size);
			while (*c && *c != ',') c++;
			if (*c == ',') c++;
		}
	}
	else if (k == "compute_lanes") { if (v < 0 || v > 4) return fail(NODEY_ENGINE_E_INVALID, "compute_lanes: 0 (automatic) .. 4"); s.compute_lanes = v; }
	else if (k == "side_streams") s.side_streams = v < 0 ? -1 : (v ? 1 : 0);
	else if (k == "stream_priority") s.stream_priority = v < 0 ? -1 : (v ? 1 : 0);
	else if (k == "stream_chunks") { if (v < 0 || v > 64) return fail(NODEY_ENGINE_E_INVALID, "stream_chunks: 0 (automatic) .. 64"); s.stream_chunks = v; }
	else if (k == "trace") s.trace = v != 0;
	else if (k == "timing") s.timing = v != 0;
	else return fail(NODEY_ENGINE_E_INVALID, "unknown schedule key: " + k);
	return 0;
}

int nodey_engine_set_export_kbps(nodey_engine* e, int kbps)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	if (kbps < 8 || kbps > 320) return fail(NODEY_ENGINE_E_INVALID, "MP3 bit rate must be 8..320 kbps");
	e->kbps = (size_t)kbps;
	return 0;
}

int nodey_engine_encode_mp3(const char* path, const void* plane0, const void* plane1, int fmt, int sample_rate, int channels,
                            int64_t frames, int frame_size, double pts_seconds, int kbps, double* time_inout)
{
	if (!path || !*path || (!plane0 && frames > 0) || frames < 0 || frame_size < 1 || (channels != 1 && channels != 2) || sample_rate < 1)
		return fail(NODEY_ENGINE_E_INVALID, "nodey_engine_encode_mp3: bad argument");
	if (format_is_planar(fmt) && channels == 2 && !plane1 && frames > 0)
		return fail(NODEY_ENGINE_E_INVALID, "nodey_engine_encode_mp3: planar stereo needs both planes");
	try
	{
		Host_stream hs;
		hs.format = fmt; hs.sample_rate = sample_rate; hs.channels = channels; hs.frames = frames; hs.pts_seconds = pts_seconds;
		hs.runs = uniform_frame_runs(frames, frame_size);
		hs.plane[0] = plane0; hs.plane[1] = plane1;
		const double end = export_mp3(hs, path, (size_t)std::max(0, kbps), time_inout ? *time_inout : 0.0);
		if (time_inout) *time_inout = end;
		return 0;
	}
	catch (const std::exception& err) { return fail(NODEY_ENGINE_E_NODE, err.what()); }
}

int nodey_engine_mp3_available(void)
{
	return processor::mp3_encoder_available() ? 1 : 0;
}

int nodey_engine_set_preview(nodey_engine* e, int preview)
{
	if (!e) return fail(NODEY_ENGINE_E_INVALID, "null engine");
	e->preview = preview != 0;
	return 0;
}

int nodey_engine_preview(nodey_engine* e, int64_t* frames, void** packed, int64_t* chunk_len, int chunk_cap)
{
	if (!e || !e->sink_data) return fail(NODEY_ENGINE_E_INVALID, "the graph has no audio_output node or has not run");
	const auto* ctx = std::any_cast<Audio_output::Process_context>(e->sink_data.get());
	if (!ctx || !ctx->preview) return fail(NODEY_ENGINE_E_NODE, "no preview was rendered (nodey_engine_set_preview before the run)");
	if (frames) *frames = ctx->preview->frames;
	if (packed) *packed = ctx->preview->plane[0];
	const int n = (int)e->preview_chunks.size();
	for (int k = 0; k < n && k < chunk_cap && chunk_len; k++) chunk_len[k] = e->preview_chunks[(size_t)k];
	return n;
}

}  // extern "C"
