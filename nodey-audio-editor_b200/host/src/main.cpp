// nodey_render -- headless offline render of a Nodey project file:
//   nodey_render project.json [out.wav | out.mp3] [--gpu N] [--kbps K] [--diagnostics]
// Sources are the WAV files named in the project's audio_input node; the sink writes a float WAV (*.wav) or, like
// the reference's export, an MP3 through LAME (any other path; libmp3lame is bound at run time).
#include "infra/graph.hpp"
#include "infra/runner.hpp"
#include "processor/nodes.hpp"

#include "nodey_cuda.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

int main(int argc, char** argv)
{
	if (argc < 2)
	{
		fprintf(stderr, "usage: %s project.json [out.wav | out.mp3] [--gpu N] [--kbps K] [--diagnostics]\n", argv[0]);
		return 2;
	}
	std::string out_path;
	int gpu = 0;
	size_t kbps = 320;      // the editor's default (src/frontend/app.cpp:595)
	bool diagnostics = false;   // print the overlay's Audio block (node states + device time per Runner step) to stderr
	for (int i = 2; i < argc; i++)
	{
		if (!strcmp(argv[i], "--gpu") && i + 1 < argc) gpu = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--kbps") && i + 1 < argc) kbps = (size_t)std::max(8, atoi(argv[++i]));
		else if (!strcmp(argv[i], "--diagnostics")) diagnostics = true;
		else out_path = argv[i];
	}
	try
	{
		if (nodey_set_device(gpu) != NODEY_OK) { fprintf(stderr, "no CUDA device: %s\n", nodey_last_error()); return 1; }
		infra::register_all_processors();
		std::ifstream f(argv[1]);
		if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
		std::stringstream ss;
		ss << f.rdbuf();
		Json::Value root;
		Json::Reader reader;
		if (!reader.parse(ss.str(), root)) { fprintf(stderr, "Invalid File: %s\n", reader.getFormattedErrorMessages().c_str()); return 1; }
		infra::Graph graph = infra::Graph::deserialize(root);

		std::map<infra::Id_t, std::shared_ptr<std::any>> node_data;
		std::shared_ptr<std::any> sink;
		if (const auto it = graph.singleton_node_map.find("audio_output"); it != graph.singleton_node_map.end())
		{
			processor::Audio_output::Process_context ctx;
			ctx.do_export = true;
			ctx.export_path = out_path;
			ctx.kbps = kbps;
			sink = std::make_shared<std::any>(ctx);
			node_data[it->second] = sink;
		}
		const auto t0 = std::chrono::steady_clock::now();
		auto runner = infra::Runner::create_and_run(graph, node_data);
		runner->wait();
		const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (diagnostics) fputs(runner->diagnostics_text().c_str(), stderr);
		const std::string err = runner->first_error();
		if (!err.empty()) { fprintf(stderr, "render failed: %s\n", err.c_str()); return 1; }
		double audio = 0;
		if (sink)
			if (const auto* ctx = std::any_cast<processor::Audio_output::Process_context>(sink.get())) audio = ctx->time->load();
		printf("{\"nodes\": %zu, \"links\": %zu, \"audio_seconds\": %.6f, \"wall_seconds\": %.6f, \"realtime_factor\": %.3f, \"gpu_launches\": %llu}\n",
			   graph.nodes.size(), graph.links.size(), audio, secs, secs > 0 ? audio / secs : 0.0, (unsigned long long)nodey_profile_launches());
		return 0;
	}
	catch (const std::exception& e)
	{
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
}
