// CPU unit test of the nodes' user-facing faults and names: a node whose input pin is not linked, or whose input stream has a
// channel count / sample format it cannot take, throws Processor::Runtime_error with the reference's own three strings
// (message, explanation, detail) -- the editor shows them verbatim (src/frontend/app.cpp error popup).  Every check
// below fails before the node touches the device, so no GPU is needed.  Reference strings:
//   audio-vol.cpp:113-117, 177-182, 238-243; audio-amix.cpp:122-126; audio-bimix.cpp:110-114, 486-490, 555-560;
//   audio-io.cpp:232-239, 858-862; audio-velocity.cpp:223-228, 278-282 (processor_name = Info::display_name, :458 / :475).
// Built and run by tests/test_host_graph.py.
#include "infra/processor.hpp"
#include "processor/audio-stream.hpp"
#include "processor/nodes.hpp"

#include <atomic>
#include <cstdio>
#include <map>
#include <string>

using namespace infra;
using namespace processor;

static int failures = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

struct Fault { bool thrown = false; std::string message, explanation, detail; };

static Fault run(const char* identifier, const Processor::Input_map& input, const Json::Value* info = nullptr)
{
    Fault f;
    auto node = Processor::processor_map.at(identifier).generate();
    if (info) node->deserialize(*info);
    Processor::Output_map output;
    std::atomic<bool> stop = false;
    std::any user;
    try { node->process_payload(input, output, stop, user); }
    catch (const Processor::Runtime_error& e) { f.thrown = true; f.message = e.message; f.explanation = e.explanation; f.detail = e.detail; }
    catch (const std::exception& e) { std::fprintf(stderr, "%s threw something else: %s\n", identifier, e.what()); failures++; }
    return f;
}

static bool is(const Fault& f, const char* message, const char* explanation, const char* detail)
{
    const bool ok = f.thrown && f.message == message && f.explanation == explanation && f.detail == detail;
    if (!ok) std::fprintf(stderr, "  got: \"%s\" | \"%s\" | \"%s\"\n", f.message.c_str(), f.explanation.c_str(), f.detail.c_str());
    return ok;
}

// a stream as a producer would have published it; no device memory behind it (the checks come first)
static std::shared_ptr<Processor::Product> stream_of(int format, int channels, int sample_rate = 48000)
{
    auto b = std::make_shared<Audio_buffer>();
    b->format = format; b->channels = channels; b->sample_rate = sample_rate; b->frames = 1152;
    b->runs = {{1152, 1}};
    auto s = std::make_shared<Audio_stream>();
    s->publish(b);
    return s;
}

int main()
{
    register_all_processors();
    const Processor::Input_map none;

    // ---- pins that are not linked ----
    CHECK(is(run("audio_volume_adjust", none), "Volume adjust processor has no input",
             "Volume adjust processor requires an audio stream input to function properly.", "Input item 'input' not found"));
    CHECK(is(run("audio_output", none), "Audio output processor has no input",
             "Audio output processor requires an audio stream input to function properly.", "Input item 'input' not found"));
    CHECK(is(run("audio_bimix", none), "Audio Channel mix processor has no input",
             "Audio channel mix processor requires an audio stream input to function properly.", "Input item 'input' not found"));
    CHECK(is(run("audio_bimix_v2", none), "Audio Channel mix processor has no input",
             "Audio channel mix processor requires an audio stream input to function properly.", "Input item 'input' not found"));
    CHECK(is(run("pitch_modifier", none), "Pitch Modifier has no input", "Pitch Modifier requires an audio stream input to function properly.",
             "Input item 'input' not found"));
    CHECK(is(run("velocity_modifier", none), "Velocity Modifier has no input",
             "Velocity Modifier requires an audio stream input to function properly.", "Input item 'input' not found"));
    {
        // audio_amix names the first pin that is missing (1-based pins, audio-amix.cpp:125)
        Json::Value info(Json::objectValue);
        info["input_num"] = 3;
        for (int i = 0; i < 3; i++) { info["volumes" + std::to_string(i)] = 1.0; info["locks" + std::to_string(i)] = false; }
        Processor::Input_map first_only{{"input_1", stream_of(FMT_FLT, 2)}};
        CHECK(is(run("audio_amix", first_only, &info), "Audio Mixer processor has no input",
                 "Audio Mixer processor requires an audio stream input to function properly.", "Input item 'input_2' not found"));
    }
    // bimix with one side linked: the reference names the pin 'input' whichever side is missing
    CHECK(is(run("audio_bimix", {{"input_l", stream_of(FMT_FLT, 2)}}), "Audio Channel mix processor has no input",
             "Audio channel mix processor requires an audio stream input to function properly.", "Input item 'input' not found"));

    // ---- streams a node cannot take ----
    CHECK(is(run("audio_volume_adjust", {{"input", stream_of(FMT_FLT, 6)}}), "Invalid channel count", "Only mono and stereo audio are supported.",
             "Got 6 channels"));
    CHECK(is(run("audio_volume_adjust", {{"input", stream_of(4 /* AV_SAMPLE_FMT_DBL */, 2)}}), "Audio format is not support",
             "Audio volume processor requires an audio format properly.", "Include FLT, S16, S32"));
    CHECK(is(run("pitch_modifier", {{"input", stream_of(4, 2)}}), "Unsupported sample format", "The processors do not support the given sample format.",
             "Sample format: dbl"));
    CHECK(is(run("velocity_modifier", {{"input", stream_of(0 /* AV_SAMPLE_FMT_U8 */, 1)}}), "Unsupported sample format",
             "The processors do not support the given sample format.", "Sample format: u8"));
    CHECK(is(run("audio_bimix_v2", {{"input_l", stream_of(FMT_FLT, 2)}, {"input_r", stream_of(FMT_FLT, 3)}}), "Invalid audio channel layout",
             "Audio channel layout must be stereo or mono.", "Invalid channel layout: 3"));

    // ---- audio_input: every slot that names a file must name a regular file, linked or not (audio-io.cpp:232-239) ----
    {
        Json::Value info(Json::objectValue), paths(Json::arrayValue);
        paths.append(""); paths.append("/nonexistent/dir/take.wav");
        info["file_path"] = paths;
        CHECK(is(run("audio_input", none, &info), "Invalid file path in slot 2", "The specified audio file does not exist or is not a regular file.",
                 "File path: /nonexistent/dir/take.wav"));
        Json::Value dir(Json::objectValue), one(Json::arrayValue);
        one.append("/tmp");
        dir["file_path"] = one;
        CHECK(is(run("audio_input", none, &dir), "Invalid file path in slot 1", "The specified audio file does not exist or is not a regular file.",
                 "File path: /tmp"));
    }

    // ---- what the editor's menus and pins show (Info::display_name, Pin_attribute::display_name; same files, :26-100) ----
    {
        const std::map<std::string, std::string> names{{"audio_input", "Audio Input"}, {"audio_output", "Audio Output"},
            {"audio_volume_adjust", "Adjust Volume"}, {"velocity_modifier", "Velocity Modifier"}, {"pitch_modifier", "Pitch Modifier"},
            {"audio_amix", "Audio Amix"}, {"audio_bimix", "Audio Bimix"}, {"audio_bimix_v2", "Audio Bimix V2"}};
        for (const auto& [identifier, shown] : names)
        {
            const auto& info = Processor::processor_map.at(identifier);
            CHECK(info.display_name == shown && info.identifier == identifier);
            CHECK(info.singleton == (identifier == "audio_input" || identifier == "audio_output"));
        }
        const auto pins_of = [](const char* identifier, const Json::Value* info = nullptr) {
            auto node = Processor::processor_map.at(identifier).generate();
            if (info) node->deserialize(*info);
            std::string text;
            for (const auto& p : node->get_pin_attributes()) text += p.identifier + "=" + p.display_name + (p.is_input ? "<" : ">") + " ";
            return text;
        };
        CHECK(pins_of("audio_volume_adjust") == "output=Output> input=Input< ");
        CHECK(pins_of("pitch_modifier") == "output=Output> input=Input< " && pins_of("velocity_modifier") == "output=Output> input=Input< ");
        CHECK(pins_of("audio_bimix") == "output=Output> input_l=Left< input_r=Right< ");
        CHECK(pins_of("audio_bimix_v2") == "output=Output> input_l=Left< input_r=Right< ");
        CHECK(pins_of("audio_output") == "input=Input< ");
        Json::Value two(Json::objectValue), paths(Json::arrayValue);
        paths.append(""); paths.append("");
        two["file_path"] = paths;
        CHECK(pins_of("audio_input", &two) == "output_0=Output 1> output_1=Output 2> ");      // 0-based pins, 1-based labels (audio-io.cpp:54-55)
        Json::Value mix(Json::objectValue);
        mix["input_num"] = 2;
        for (int i = 0; i < 2; i++) { mix["volumes" + std::to_string(i)] = 0.5; mix["locks" + std::to_string(i)] = false; }
        CHECK(pins_of("audio_amix", &mix) == "output=Output> input_1=Input 1< input_2=Input 2< ");
    }

    if (failures) { std::fprintf(stderr, "%d failure(s)\n", failures); return 1; }
    std::printf("node_errors_test ok\n");
    return 0;
}
