// CPU unit test of the Graph editing API the editor drives: add_node / remove_node /
// update_node_pin / add_link / remove_link / get_pin_to_node_map / get_node_input_map, with the reference's
// semantics (src/infra/graph.cpp:9-178): smallest-free-id allocation, singleton bookkeeping, links re-attached
// by pin NAME when a node's pin set changes, Mismatched / Multiple_input errors from add_link, serialize ->
// deserialize round trip of an edited graph.  No device is touched.  Built and run by tests/test_host_graph.py.
#include "infra/graph.hpp"
#include "processor/nodes.hpp"

#include <cstdio>
#include <cstdlib>

using namespace infra;

static int failures = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

template <typename E, typename F>
static bool throws(F&& f)
{
    try { f(); } catch (const E&) { return true; } catch (...) { return false; }
    return false;
}

static std::unique_ptr<Processor> make(const char* identifier) { return Processor::processor_map.at(identifier).generate(); }

static Json::Value amix_info(int n)
{
    Json::Value v(Json::objectValue);
    v["input_num"] = n;
    for (int i = 0; i < n; i++) { v["volumes" + std::to_string(i)] = 1.0; v["locks" + std::to_string(i)] = false; }
    return v;
}

int main()
{
    register_all_processors();
    register_all_processors();                                                    // the set is registered once per process
    CHECK(Processor::processor_map.size() == 10 && Processor::processor_map.count("audio_bimix_v2"));
    CHECK(throws<std::logic_error>([] { Processor::register_processor<processor::Audio_vol>(); }));   // duplicate identifier (processor.hpp:116-129)

    Graph g;
    const Id_t in = g.add_node(make("audio_input"));
    const Id_t vol = g.add_node(make("audio_volume_adjust"));
    const Id_t mix = g.add_node(make("audio_amix"));
    const Id_t out = g.add_node(make("audio_output"));
    CHECK(in == 0 && vol == 1 && mix == 2 && out == 3 && g.modified);
    CHECK(g.singleton_node_map.at("audio_input") == in && g.singleton_node_map.at("audio_output") == out);
    CHECK(g.singleton_node_map.size() == 2);

    // pins were created by update_node_pin, named like the processor's attributes
    auto& mixer = g.nodes.at(mix);
    mixer.processor->deserialize(amix_info(3));
    g.update_node_pin(mix);
    CHECK(mixer.pin_name_map.count("output") && mixer.pin_name_map.count("input_1") && mixer.pin_name_map.count("input_3"));
    CHECK(!mixer.pin_name_map.count("input_4") && mixer.pins.size() == 4);
    for (const auto& [name, pin] : mixer.pin_name_map) CHECK(g.pins.at(pin).parent == mix && g.pins.at(pin).attribute.identifier == name);

    const Id_t l0 = g.add_link(g.nodes.at(in).pin_name_map.at("output_0"), g.nodes.at(vol).pin_name_map.at("input"));
    const Id_t l1 = g.add_link(g.nodes.at(vol).pin_name_map.at("output"), mixer.pin_name_map.at("input_1"));
    const Id_t l2 = g.add_link(g.nodes.at(vol).pin_name_map.at("output"), mixer.pin_name_map.at("input_3"));    // fan-out
    const Id_t l3 = g.add_link(mixer.pin_name_map.at("output"), g.nodes.at(out).pin_name_map.at("input"));
    CHECK(l0 == 0 && l1 == 1 && l2 == 2 && l3 == 3);
    g.check_graph();

    // the reference's add_link tests the pin BEFORE it adds (check_multiple_input is true while the pin has at most one
    // link, graph.hpp:173-183): a second link into an input pin is accepted and only check_graph objects; a third is refused
    const Id_t extra = g.add_link(g.nodes.at(in).pin_name_map.at("output_0"), mixer.pin_name_map.at("input_1"));
    CHECK(extra == 4 && !g.check_multiple_input(mixer.pin_name_map.at("input_1")));
    CHECK(throws<Graph::Multiple_input_error>([&] { g.check_graph(); }));
    CHECK(throws<Graph::Multiple_input_error>([&] { g.add_link(g.nodes.at(in).pin_name_map.at("output_0"), mixer.pin_name_map.at("input_1")); }));
    g.remove_link(extra);
    CHECK(g.links.size() == 4 && g.check_multiple_input(mixer.pin_name_map.at("input_1")));
    g.check_graph();

    // maps used by the editor
    const auto pin_to_node = g.get_pin_to_node_map();
    CHECK(pin_to_node.size() == g.pins.size() && pin_to_node.at(mixer.pin_name_map.at("output")) == mix);
    const auto inputs = g.get_node_input_map();
    CHECK(inputs.at(in).empty() && inputs.at(mix).size() == 1 && inputs.at(out).size() == 1);       // both mixer inputs come from one pin

    // shrinking the mixer to two inputs: input_1 and output keep their links BY NAME, the link into input_3 goes
    mixer.processor->deserialize(amix_info(2));
    g.update_node_pin(mix);
    CHECK(mixer.pins.size() == 3 && !mixer.pin_name_map.count("input_3"));
    CHECK(g.links.size() == 3);
    {
        bool in1 = false, outl = false;
        for (const auto& [_, l] : g.links)
        {
            if (l.to == mixer.pin_name_map.at("input_1") && l.from == g.nodes.at(vol).pin_name_map.at("output")) in1 = true;
            if (l.from == mixer.pin_name_map.at("output") && l.to == g.nodes.at(out).pin_name_map.at("input")) outl = true;
        }
        CHECK(in1 && outl);
    }
    g.check_graph();

    // ids are reused smallest-first (find_empty): removing link 0 frees id 0 for the next link
    g.remove_link(l0);
    CHECK(!g.links.count(0));
    const Id_t again = g.add_link(g.nodes.at(in).pin_name_map.at("output_0"), g.nodes.at(vol).pin_name_map.at("input"));
    CHECK(again == 0);
    g.remove_link(g.nodes.at(in).pin_name_map.at("output_0"), g.nodes.at(vol).pin_name_map.at("input"));          // by endpoints
    CHECK(!g.links.count(0) && g.links.size() == 2);

    // remove_node drops the node's pins and every link touching them, and frees the singleton slot and the id
    const size_t pins_before = g.pins.size();
    g.remove_node(out);
    CHECK(!g.nodes.count(out) && !g.singleton_node_map.count("audio_output"));
    CHECK(g.pins.size() == pins_before - 1 && g.links.size() == 1);
    const Id_t out2 = g.add_node(make("audio_output"));
    CHECK(out2 == out && g.singleton_node_map.at("audio_output") == out2);

    // a loop is caught by check_graph, not by add_link
    const Id_t vol2 = g.add_node(make("audio_volume_adjust"));
    g.add_link(mixer.pin_name_map.at("output"), g.nodes.at(vol2).pin_name_map.at("input"));
    g.add_link(g.nodes.at(vol2).pin_name_map.at("output"), mixer.pin_name_map.at("input_2"));
    CHECK(throws<Graph::Loop_detected_error>([&] { g.check_graph(); }));
    g.remove_link(g.nodes.at(vol2).pin_name_map.at("output"), mixer.pin_name_map.at("input_2"));
    g.check_graph();

    // the edited graph survives serialize -> deserialize with the same ids, identifiers and links
    const Json::Value saved = g.serialize();
    const Graph h = Graph::deserialize(saved);
    CHECK(h.nodes.size() == g.nodes.size() && h.links.size() == g.links.size() && h.pins.size() == g.pins.size());
    for (const auto& [id, node] : g.nodes)
        CHECK(h.nodes.count(id) && h.nodes.at(id).processor->get_processor_info_non_static().identifier
                                       == node.processor->get_processor_info_non_static().identifier);
    CHECK(h.nodes.at(mix).pin_name_map.count("input_2") && !h.nodes.at(mix).pin_name_map.count("input_3"));
    const auto levels = h.topological_levels();
    CHECK(!levels.empty() && levels.front().size() >= 1);

    if (failures) { std::fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    std::puts("graph_edit_test ok");
    return 0;
}
