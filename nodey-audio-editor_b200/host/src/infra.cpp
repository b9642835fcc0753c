// infra.cpp -- registry, execution context, device handles, graph model and the level-batched Runner.
#include "infra/exec-context.hpp"
#include "infra/graph.hpp"
#include "infra/runner.hpp"

#include "nodey_cuda.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <typeindex>

namespace infra
{
	std::map<std::string, Processor::Info> Processor::processor_map;

	// ---------------------------------------------------------------------------------------------
	// plugin API: out-of-line parts of infra/processor.hpp and infra/graph.hpp
	// ---------------------------------------------------------------------------------------------
	static std::string runtime_error_text(const std::string& message, const std::string& explanation, const std::string& detail)
	{
		return message + " (Detail: " + detail + ") (Explanation: " + explanation + ")";
	}

	Processor::Runtime_error::Runtime_error(std::string message_, std::string explanation_, std::string detail_) :
		std::runtime_error(runtime_error_text(message_, explanation_, detail_)),
		message(std::move(message_)), explanation(std::move(explanation_)), detail(std::move(detail_))
	{
	}

	void Processor::add_to_registry(Info info)
	{
		const std::string identifier = info.identifier;
		if (!processor_map.try_emplace(identifier, std::move(info)).second)
			THROW_LOGIC_ERROR("Processor with identifier '{}' already registered", identifier);
	}

	std::shared_ptr<Processor::Product> detail::input_product(const Processor::Input_map& input, const std::string& key,
															  const std::type_info& expected)
	{
		const auto slot = input.find(key);
		if (slot == input.end()) return nullptr;
		const auto& product = slot->second;
		if (!product) THROW_LOGIC_ERROR("Found nullptr in input map for key '{}'", key);
		if (product->get_typeinfo() != expected)
			THROW_LOGIC_ERROR("Type mismatch in input map for key '{}', expected {}, got {}", key, expected.name(), product->get_typeinfo().name());
		return product;
	}

	const std::set<std::shared_ptr<Processor::Product>>& detail::output_products(const Processor::Output_map& output, const std::string& key)
	{
		const auto slot = output.find(key);
		if (slot == output.end()) THROW_LOGIC_ERROR("Key '{}' not found in output map", key);
		for (const auto& product : slot->second)
			if (!product) THROW_LOGIC_ERROR("Found nullptr in output map for key '{}'", key);
		return slot->second;
	}

	Graph::Mismatched_pin_error::Mismatched_pin_error(Id_t from_, Id_t to_) :
		std::runtime_error("Mismatch Pin: " + std::to_string(from_) + ", " + std::to_string(to_)), from(from_), to(to_)
	{
	}
	Graph::Loop_detected_error::Loop_detected_error() : std::runtime_error("Loop Detected") {}
	Graph::Multiple_input_error::Multiple_input_error(Id_t pin_) :
		std::runtime_error("Multiple Inputs in Input Pin: " + std::to_string(pin_)), pin(pin_)
	{
	}
	Graph::Invalid_file_error::Invalid_file_error(std::string message_) :
		std::runtime_error("Invalid File: " + message_), message(std::move(message_))
	{
	}

	bool Graph::check_node_type_match(Id_t from, Id_t to) const
	{
		// pin types are compared by the ADDRESS of their type_info, as the reference does (include/infra/graph.hpp:167-170)
		const std::type_info* a = &pins.at(from).attribute.type.get();
		const std::type_info* b = &pins.at(to).attribute.type.get();
		return a == b;
	}

	bool Graph::check_multiple_input(Id_t pin_id) const
	{
		const auto arrivals = std::count_if(links.begin(), links.end(), [pin_id](const auto& entry) { return entry.second.to == pin_id; });
		return arrivals <= 1;
	}

	namespace
	{
		// smallest id that is not a key of the map: ids are handed out again after a removal, like the reference's
		template <typename Map>
		Id_t find_empty(const Map& items)
		{
			Id_t candidate = 0;
			for (auto it = items.begin(); it != items.end() && it->first == candidate; ++it) ++candidate;
			return candidate;
		}
	}

	// ---------------------------------------------------------------------------------------------
	// device handles over the C ABI
	// ---------------------------------------------------------------------------------------------
	static void abi_check(int rc, const char* what)
	{
		if (rc == NODEY_OK) return;
		const std::string text = nodey_last_error();
		if (rc == NODEY_E_NOMEM) throw std::bad_alloc();
		throw Processor::Runtime_error("GPU runtime call failed", std::format("{} returned {}", what, rc), text);
	}

	Exec_context& Exec_context::current()
	{
		thread_local Exec_context ctx;
		return ctx;
	}

	Device_event::Device_event() { abi_check(nodey_event_create(&ev, 0), "nodey_event_create"); }
	Device_event::~Device_event() { nodey_event_destroy(ev); }
	void Device_event::record(Stream_handle stream) { abi_check(nodey_event_record(ev, stream), "nodey_event_record"); }
	void Device_event::wait_on(Stream_handle stream) const { abi_check(nodey_stream_wait_event(stream, ev), "nodey_stream_wait_event"); }
	void Device_event::synchronize() const { abi_check(nodey_event_synchronize(ev), "nodey_event_synchronize"); }

	Device_block::Device_block(size_t n) : bytes(n)
	{
		abi_check(nodey_malloc(&ptr, n, Exec_context::current().stream), "nodey_malloc");
	}

	Device_block::~Device_block() { Lane_registry::free_ordered(ptr); }

	namespace
	{
		struct Lane { Stream_handle stream; Event_handle event; bool compute; };
		struct Lanes
		{
			std::mutex mutex;
			std::vector<Lane> lanes;
			Stream_handle reaper = nullptr;
		};
		Lanes& lanes_state()
		{
			static Lanes s;
			return s;
		}
	}

	void Lane_registry::add(Stream_handle lane, bool compute)
	{
		Lanes& s = lanes_state();
		std::lock_guard lock(s.mutex);
		Event_handle ev = nullptr;
		abi_check(nodey_event_create(&ev, 0), "nodey_event_create");
		s.lanes.push_back({lane, ev, compute});
	}

	void Lane_registry::remove(Stream_handle lane)
	{
		Lanes& s = lanes_state();
		std::lock_guard lock(s.mutex);
		for (auto it = s.lanes.begin(); it != s.lanes.end(); ++it)
			if (it->stream == lane)
			{
				nodey_event_destroy(it->event);
				s.lanes.erase(it);
				return;
			}
	}

	void Lane_registry::free_ordered(void* ptr)
	{
		if (!ptr) return;
		Lanes& s = lanes_state();
		std::lock_guard lock(s.mutex);
		Stream_handle compute = nullptr;
		int n_compute = 0;
		for (const Lane& l : s.lanes)
			if (l.compute) { compute = l.stream; n_compute++; }
		if (s.lanes.empty())
		{
			// no render in flight: the legacy default stream orders the free after all earlier work
			nodey_free(ptr, nullptr);
			return;
		}
		if (n_compute == 1)
		{
			// one render in flight: every consumer enqueued on its compute lane, so the free goes there too
			nodey_free(ptr, compute);
			return;
		}
		// several renders at once: order the free after everything enqueued on every lane
		if (!s.reaper && nodey_stream_create(&s.reaper) != NODEY_OK) { nodey_free(ptr, nullptr); return; }
		for (Lane& l : s.lanes)
		{
			nodey_event_record(l.event, l.stream);
			nodey_stream_wait_event(s.reaper, l.event);
		}
		nodey_free(ptr, s.reaper);
	}

	// ---------------------------------------------------------------------------------------------
	// Graph (reference: src/infra/graph.cpp)
	// ---------------------------------------------------------------------------------------------
	Id_t Graph::add_node(std::unique_ptr<Processor> processor)
	{
		const Id_t id = find_empty(nodes);
		const auto info = processor->get_processor_info_non_static();
		Node node;
		node.processor = std::move(processor);
		nodes[id] = std::move(node);
		update_node_pin(id);
		if (info.singleton) singleton_node_map.emplace(info.identifier, id);
		modified = true;
		return id;
	}

	void Graph::remove_node(Id_t id)
	{
		const auto find = nodes.find(id);
		if (find == nodes.end()) return;
		const auto info = find->second.processor->get_processor_info_non_static();
		if (info.singleton)
		{
			const auto it = singleton_node_map.find(info.identifier);
			if (it == singleton_node_map.end()) THROW_LOGIC_ERROR("Singleton node ID not found");
			if (it->second != id) THROW_LOGIC_ERROR("Singleton node ID mismatch, expected {}, got {}", it->second, id);
			singleton_node_map.erase(it);
		}
		const std::set<Id_t> owned = find->second.pins;
		for (const Id_t pin : owned) pins.erase(pin);
		std::erase_if(links, [&owned](const auto& kv) { return owned.contains(kv.second.from) || owned.contains(kv.second.to); });
		nodes.erase(find);
		modified = true;
	}

	void Graph::update_node_pin(Id_t id)
	{
		Node& node = nodes[id];

		// remember what the node's current pins were linked to, by pin name
		std::map<std::string, Id_t> old_inputs;
		std::map<std::string, std::set<Id_t>> old_outputs;
		for (auto it = links.begin(); it != links.end();)
		{
			const Link link = it->second;
			if (node.pins.contains(link.from)) { old_outputs[pins.at(link.from).attribute.identifier].insert(link.to); it = links.erase(it); }
			else if (node.pins.contains(link.to)) { old_inputs[pins.at(link.to).attribute.identifier] = link.from; it = links.erase(it); }
			else ++it;
		}
		for (const Id_t pin : node.pins) pins.erase(pin);
		node.pins.clear();
		node.pin_name_map.clear();

		// re-create the pins from the processor's current attributes and restore compatible links
		for (const auto& attribute : node.processor->get_pin_attributes())
		{
			if (node.pin_name_map.contains(attribute.identifier))
				THROW_LOGIC_ERROR("Pin name {} already exists for node ID {}", attribute.identifier, id);
			const Id_t pin_id = find_empty(pins);
			node.pins.insert(pin_id);
			pins.emplace(pin_id, Pin{.parent = id, .attribute = attribute});
			node.pin_name_map.emplace(attribute.identifier, pin_id);

			if (const auto in = old_inputs.find(attribute.identifier);
				in != old_inputs.end() && attribute.type.get() == pins.at(in->second).attribute.type.get())
				links.emplace(find_empty(links), Link{.from = in->second, .to = pin_id});
			if (const auto out = old_outputs.find(attribute.identifier); out != old_outputs.end())
				for (const Id_t to : out->second)
					if (attribute.type.get() == pins.at(to).attribute.type.get())
						links.emplace(find_empty(links), Link{.from = pin_id, .to = to});
		}
		modified = true;
	}

	Id_t Graph::add_link(Id_t from, Id_t to)
	{
		if (!check_node_type_match(from, to)) throw Mismatched_pin_error{from, to};
		if (!check_multiple_input(to)) throw Multiple_input_error{to};
		const Id_t id = find_empty(links);
		links.insert_or_assign(id, Link{.from = from, .to = to});
		modified = true;
		return id;
	}

	void Graph::remove_link(Id_t id)
	{
		links.erase(id);
		modified = true;
	}

	void Graph::remove_link(Id_t from, Id_t to)
	{
		std::erase_if(links, [&](const auto& kv) { return kv.second.from == from && kv.second.to == to; });
		modified = true;
	}

	std::map<Id_t, Id_t> Graph::get_pin_to_node_map() const
	{
		std::map<Id_t, Id_t> out;
		for (const auto& [id, node] : nodes)
			for (const Id_t pin : node.pins) out[pin] = id;
		return out;
	}

	std::map<Id_t, std::set<Id_t>> Graph::get_node_input_map() const
	{
		std::map<Id_t, std::set<Id_t>> out;
		for (const auto& [id, _] : nodes) out.emplace(id, std::set<Id_t>{});
		for (const auto& [_, link] : links) out[pins.at(link.to).parent].insert(link.from);
		return out;
	}

	std::vector<std::vector<Id_t>> Graph::topological_levels() const
	{
		// A render checks and levels its graph on every run, so this is linear work on flat arrays (the 256-track project has
		// 1044 nodes and 1300 links; the first version -- check_multiple_input() per link, maps of sets -- took 10 ms per run).
		std::map<Id_t, int> arrivals;            // links per input pin
		for (const auto& [_, link] : links) arrivals[link.to]++;
		const size_t slots = nodes.empty() ? 0 : (size_t)nodes.rbegin()->first + 1;      // node ids are small non-negative integers
		std::vector<std::vector<Id_t>> successors(slots);
		std::vector<int> pending(slots, 0);      // links that still have to be satisfied (a producer linked twice counts twice, on both sides)
		for (const auto& [_, link] : links)
		{
			if (!check_node_type_match(link.from, link.to)) throw Mismatched_pin_error{link.from, link.to};
			if (arrivals.at(link.to) > 1) throw Multiple_input_error(link.to);
			const Id_t from = pins.at(link.from).parent, to = pins.at(link.to).parent;
			if (from < 0 || to < 0 || (size_t)from >= slots || (size_t)to >= slots) THROW_LOGIC_ERROR("Link between unknown nodes {} -> {}", from, to);
			successors[from].push_back(to);
			pending[to]++;
		}
		// Kahn's algorithm on node level: a node is ready when all its producers are placed
		std::vector<std::vector<Id_t>> levels;
		std::vector<Id_t> frontier;
		for (const auto& [id, _] : nodes)
			if (pending[id] == 0) frontier.push_back(id);
		if (!nodes.empty() && frontier.empty()) throw Loop_detected_error{};
		size_t placed = 0;
		while (!frontier.empty())
		{
			levels.push_back(frontier);
			placed += frontier.size();
			std::vector<Id_t> next;
			for (const Id_t id : frontier)
				for (const Id_t s : successors[id])
					if (--pending[s] == 0) next.push_back(s);
			std::sort(next.begin(), next.end());
			frontier = std::move(next);
		}
		if (placed != nodes.size()) throw Loop_detected_error{};
		return levels;
	}

	void Graph::check_graph() const { (void)topological_levels(); }

	Json::Value Graph::serialize() const
	{
		Json::Value node_json(Json::objectValue);
		for (const auto& [id, node] : nodes)
		{
			Json::Value item;
			item["identifier"] = node.processor->get_processor_info_non_static().identifier;
			item["info"] = node.processor->serialize();
			item["position"]["x"] = node.position.x;
			item["position"]["y"] = node.position.y;
			node_json[std::to_string(id)] = std::move(item);
		}
		Json::Value link_json(Json::arrayValue);
		for (const auto& [_, link] : links)
		{
			const Pin& from_pin = pins.at(link.from);
			const Pin& to_pin = pins.at(link.to);
			Json::Value item;
			item["from"]["node"] = from_pin.parent;
			item["from"]["pin"] = from_pin.attribute.identifier;
			item["to"]["node"] = to_pin.parent;
			item["to"]["pin"] = to_pin.attribute.identifier;
			link_json.append(std::move(item));
		}
		Json::Value result;
		result["nodes"] = std::move(node_json);
		result["links"] = std::move(link_json);
		return result;
	}

	Graph Graph::deserialize(const Json::Value& value)
	try
	{
		if (!value.isObject()) throw Invalid_file_error("Invalid graph format, expected object");
		const Json::Value& nodes_json = value["nodes"];
		const Json::Value& links_json = value["links"];
		if (!nodes_json.isObject()) throw Invalid_file_error("Invalid nodes format, expected object");
		if (!links_json.isArray()) throw Invalid_file_error("Invalid links format, expected array");

		Graph graph;
		for (const auto& key : nodes_json.getMemberNames())
		{
			size_t used = 0;
			Id_t id = 0;
			try { id = std::stoi(key, &used); } catch (const std::exception&) { used = 0; }
			if (used != key.length() || key.empty()) throw Invalid_file_error(std::format("Invalid node ID: {}", key));

			const Json::Value& node_json = nodes_json[key];
			if (!node_json.isObject()) throw Invalid_file_error(std::format("Invalid node JSON format for ID: {}", id));
			const std::string identifier = node_json["identifier"].asString();
			const auto meta = Processor::processor_map.find(identifier);
			if (meta == Processor::processor_map.end())
				throw Invalid_file_error(std::format("Unknown processor identifier: {}", identifier));

			std::shared_ptr<Processor> processor = meta->second.generate();
			processor->deserialize(node_json["info"]);
			if (meta->second.singleton)
			{
				if (graph.singleton_node_map.contains(identifier))
					throw Invalid_file_error(std::format("Duplicating singleton node \"{}\"", identifier));
				graph.singleton_node_map.emplace(identifier, id);
			}
			Node node;
			node.processor = std::move(processor);
			node.position = ImVec2(node_json["position"]["x"].asFloat(), node_json["position"]["y"].asFloat());
			graph.nodes.emplace(id, std::move(node));
			graph.update_node_pin(id);
		}

		for (const Json::Value& link : links_json)
		{
			if (!link.isObject()) throw Invalid_file_error("Invalid link JSON format, expected object");
			const Json::Value& from_json = link["from"];
			const Json::Value& to_json = link["to"];
			if (!from_json.isObject() || !to_json.isObject())
				throw Invalid_file_error("Invalid link 'from' or 'to' JSON format, expected object");
			const Id_t from_node = from_json["node"].asInt(), to_node = to_json["node"].asInt();
			const std::string from_pin = from_json["pin"].asString(), to_pin = to_json["pin"].asString();
			if (!graph.nodes.contains(from_node) || !graph.nodes.contains(to_node))
				throw Invalid_file_error(std::format("Link references non-existent node: {} -> {}", from_node, to_node));
			const auto& from_map = graph.nodes.at(from_node).pin_name_map;
			const auto& to_map = graph.nodes.at(to_node).pin_name_map;
			if (!from_map.contains(from_pin) || !to_map.contains(to_pin))
				throw Invalid_file_error(
					std::format("Link references non-existent pin: {}.{} -> {}.{}", from_node, from_pin, to_node, to_pin));
			graph.add_link(from_map.at(from_pin), to_map.at(to_pin));
		}
		return graph;
	}
	catch (const Json::Exception& e)
	{
		throw Invalid_file_error(std::format("Failed to deserialize graph due to JSON error: {}", e.what()));
	}

	// ---------------------------------------------------------------------------------------------
	// Runner
	// ---------------------------------------------------------------------------------------------
	void Runner::generate_processor_resources(const Graph& graph)
	{
		const auto T0 = std::chrono::steady_clock::now();
		auto TP = T0;
		const bool timing = schedule.timing;
		const auto lap = [&](const char* what) { const auto n = std::chrono::steady_clock::now(); if (timing) fprintf(stderr, "  [resources] %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(n - TP).count()); TP = n; };
		levels = graph.topological_levels();   // = check_graph(), and the schedule
		lap("levels");

		for (const auto& [id, node] : graph.nodes)
		{
			auto resource = std::make_shared<Processor_resource>();
			resource->processor = node.processor;
			for (const Id_t pin_id : node.pins)
			{
				const auto& attribute = graph.pins.at(pin_id).attribute;
				if (!attribute.is_input) resource->output_payloads.emplace(attribute.identifier, std::set<std::shared_ptr<Processor::Product>>{});
			}
			processor_resources.emplace(id, std::move(resource));
		}

		lap("resources");
		// one product per link, generated by the source pin; fan-out = several products on one output pin
		for (const auto& [idx, link] : graph.links)
		{
			const Graph::Pin& from_pin = graph.pins.at(link.from);
			const Graph::Pin& to_pin = graph.pins.at(link.to);
			std::shared_ptr<Processor::Product> product = from_pin.attribute.generate_func();
			processor_resources.at(from_pin.parent)->output_payloads[from_pin.attribute.identifier].insert(product);
			processor_resources.at(to_pin.parent)->input_payloads.emplace(to_pin.attribute.identifier, product);
			link_products.emplace(idx, product);
		}

		lap("products");
		// waves: links leaving a source node (level 0) belong to wave (position of the pin among the node's
		// output pins) / wave_size; every other node runs in the latest wave among its inputs
		// Waves hide the host -> device upload behind compute; smaller batches cost some kernel efficiency, so a render
		// whose sources are already in HBM runs as one wave, or as two halves from 128 source pins on (see below).
		size_t upload = 0;
		std::any no_data;
		if (!levels.empty())
			for (const Id_t id : levels.front())
			{
				const auto data = node_data.find(id);
				upload += graph.nodes.at(id).processor->upload_bytes(data == node_data.end() ? no_data : *data->second);
			}
		// Uploads are only a little slower than compute per track (1.15 vs 0.94 ms), so the first wave's upload and the
		// last wave's compute are exposed: both want SMALL waves.  A SoundTouch chain costs the same 40 ms for 8 or 40
		// tracks (sequential per track), so waves of 32 source pins are the smallest that still pay; consecutive waves
		// run on three rotating compute lanes so that a wave's sequential chains do not leave the SMs idle.  Measured on
		// the 256-track render (tools/e2e_trace.py): 364 ms for 32-pin waves on three lanes, 379 ms for 32/64/64/64/32 on
		// two, 437 ms for 32/96/128 on one; the 16.26 GB upload alone takes 294 ms.  Schedule::wave_pins forces waves of n
		// pins, Schedule::wave_pattern an explicit pattern.
		const bool pipelined = upload >= (1u << 30);
		const int uniform = std::max(0, schedule.wave_pins);
		int source_pins = 0;
		if (!levels.empty())
			for (const Id_t id : levels.front())
				for (const auto& attribute : graph.nodes.at(id).processor->get_pin_attributes())
					if (!attribute.is_input) source_pins++;
		wave_begin.assign(1, 0);             // first pin position of every wave
		if (!schedule.wave_pattern.empty())
		{
			// explicit wave sizes, e.g. {32, 64, 48} (the last one repeats)
			int p = 0;
			size_t k = 0;
			while (p < source_pins)
			{
				const int size = std::max(1, schedule.wave_pattern[std::min(k++, schedule.wave_pattern.size() - 1)]);
				p += size;
				if (p < source_pins) wave_begin.push_back(p);
			}
		}
		else if (uniform > 0)
			for (int p = uniform; p < source_pins; p += uniform) wave_begin.push_back(p);
		// (sources already in HBM: ONE wave.  Round 1 split such a render into two half-size waves on two lanes to fill
		// what a single stream left idle; now the nodes of a track chain overlap chunk by chunk on the lane's three streams,
		// which fills the same gaps with full-size batches: 256 tracks 216 ms as one wave, 230 ms as two.)
		else if (pipelined && source_pins > 32)
		{
			constexpr int kEdge = 32, kBody = 32;
			int p = kEdge;
			while (source_pins - p > kBody + kEdge) { wave_begin.push_back(p); p += kBody; }
			wave_begin.push_back(p);
			if (source_pins - p > kBody) wave_begin.push_back(source_pins - kEdge);
		}
		lap("wave sizes");
		const auto wave_of_pin = [&](int position) {
			return (int)(std::upper_bound(wave_begin.begin(), wave_begin.end(), position) - wave_begin.begin()) - 1;
		};
		std::map<Id_t, int> pin_position;     // output pin of a source node -> index in its node's attribute order
		if (!levels.empty())
			for (const Id_t id : levels.front())
			{
				const auto& node = graph.nodes.at(id);
				int k = 0;
				for (const auto& attribute : node.processor->get_pin_attributes())
					if (!attribute.is_input) pin_position[node.pin_name_map.at(attribute.identifier)] = k++;
			}
		std::set<Id_t> sources(levels.empty() ? std::vector<Id_t>{}.begin() : levels.front().begin(),
							   levels.empty() ? std::vector<Id_t>{}.end() : levels.front().end());
		for (const auto& [id, _] : graph.nodes) node_wave[id] = 0;
		std::map<Id_t, std::vector<Id_t>> incoming;     // node -> producer pins
		for (const auto& [_, link] : graph.links) incoming[graph.pins.at(link.to).parent].push_back(link.from);
		for (size_t l = 1; l < levels.size(); l++)
			for (const Id_t id : levels[l])
			{
				int wave = 0;
				const auto in = incoming.find(id);
				if (in != incoming.end())
					for (const Id_t from_pin : in->second)
					{
						const Id_t producer = graph.pins.at(from_pin).parent;
						wave = std::max(wave, sources.contains(producer) ? wave_of_pin(pin_position.at(from_pin)) : node_wave.at(producer));
					}
				node_wave[id] = wave;
			}
		lap("node waves");
	}

	namespace { std::atomic<bool> g_release_products{false}; }
	void Runner::release_products(bool release)
	{
		g_release_products = release;
		// released blocks are only worth something if the next allocation may take them before the device has passed the
		// free point: the allocator then orders the taker's stream after that point
		nodey_set_memory_policy(release ? 1 : 0);
	}

	Runner::~Runner()
	{
		for (auto& [_, resource] : processor_resources) resource->stop_source = true;
		if (worker.joinable()) worker.join();
	}

	void Runner::wait()
	{
		if (worker.joinable()) worker.join();
	}

	std::string Runner::first_error() const
	{
		for (const auto& [id, resource] : processor_resources)
		{
			if (resource->state != State::Error) continue;
			const std::string who = std::format("node {} ({}): ", id, resource->processor->get_processor_info_non_static().identifier);
			if (const auto* e = std::any_cast<Processor::Runtime_error>(&resource->exception)) return who + e->what();
			if (const auto* e = std::any_cast<std::runtime_error>(&resource->exception)) return who + e->what();
			if (const auto* e = std::any_cast<std::logic_error>(&resource->exception)) return who + e->what();
			return who + "unknown exception";
		}
		return "";
	}

	std::string Runner::diagnostics_text() const
	{
		// count_processor_states + the "%d Running | %d Finished | %d Errors" line of app.cpp:1558-1567
		int running = 0, finished_count = 0, errors = 0;
		for (const auto& [_, r] : processor_resources)
		{
			const State st = r->state.load();
			running += st == State::Running; finished_count += st == State::Finished; errors += st == State::Error;
		}
		std::string text = std::format("{} Running | {} Finished | {} Errors\n", running, finished_count, errors);
		for (const auto& t : get_level_timings())
			text += std::format("W{} L{} {} x{}: {:.2f} ms (lane {}, +{:.2f} ms, enqueue {:.2f} ms)\n", t.wave, t.level, t.identifier, t.nodes,
								t.device_ms, t.lane, t.start_ms, t.enqueue_ms);
		return text;
	}

	namespace
	{
		// the catch ladder of the reference's fiber body (src/infra/runner.cpp:87-136)
		template <typename F>
		bool guarded(Runner::Processor_resource& r, F&& body)
		{
			const auto name = [&] { return r.processor->get_processor_info_non_static().identifier; };
			try { body(); return true; }
			catch (const Processor::Runtime_error& e) { r.exception = e; }
			catch (const std::bad_any_cast&) { r.exception = std::logic_error(std::format("Bad any cast found in the processor \"{}\"", name())); }
			catch (const std::bad_alloc&) { r.exception = std::runtime_error(std::format("Memory allocation failed in the processor \"{}\"", name())); }
			catch (const std::bad_optional_access&) { r.exception = std::logic_error(std::format("Bad optional access found in the processor \"{}\"", name())); }
			catch (const std::runtime_error& e) { r.exception = e; }
			catch (const std::logic_error& e) { r.exception = e; }
			catch (...) { r.exception = std::exception(); }
			r.state = Runner::State::Error;
			return false;
		}
	}

	void Runner::launch_threads()
	{
		// lane 0: transfers (nodes without inputs: the sources' uploads); lanes 1..: compute (everything else).
		// One compute lane when the render is a single wave; waves rotate over up to three otherwise.
		int max_wave = 0;
		for (const auto& [_, w] : node_wave) max_wave = std::max(max_wave, w);
		int compute_lanes = max_wave > 1 ? 3 : (max_wave > 0 ? 2 : 1);
		if (schedule.compute_lanes > 0) compute_lanes = std::clamp(schedule.compute_lanes, 1, 4);
		constexpr int kMaxLanes = 5;
		const int kLanes = 1 + compute_lanes;
		nodey_stream_t lanes[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr, nullptr};
		// every compute lane has a side stream: a node that consumes a stream chunk by chunk runs there, next to its producer
		constexpr int kSides = Exec_context::kSideStreams;
		nodey_stream_t sides[kSides][kMaxLanes] = {};
		const auto release_lanes = [&] {
			const auto drop = [](nodey_stream_t& s) {
				if (!s) return;
				Lane_registry::remove(s);
				nodey_stream_destroy(s);
				s = nullptr;
			};
			for (auto& s : lanes) drop(s);
			for (auto& side : sides)
				for (auto& s : side) drop(s);
		};
		bool failed = false;
		if (device >= 0 && nodey_set_device(device) != NODEY_OK) failed = true;
		for (int k = 0; k < kLanes && !failed; k++)
		{
			if (nodey_stream_create(&lanes[k]) != NODEY_OK) { lanes[k] = nullptr; failed = true; break; }
			Lane_registry::add(lanes[k], k >= 1);
			if (k >= 1 && schedule.side_streams != 0)
			{
				// sides 0 and 2 carry the sequential WSOLA searches of a resampler -> pitch -> tempo chain (nodes.cpp: the
				// search runs on the stream after the producer's, the tails on the next one): high priority, so that a
				// chain's next launch gets its SM slots ahead of the thousands of tail CTAs it competes with
				const bool prio = schedule.stream_priority != 0;
				for (int j = 0; j < kSides; j++)
				{
					const int rc = prio ? nodey_stream_create_priority(&sides[j][k], j % 2 == 0) : nodey_stream_create(&sides[j][k]);
					if (rc != NODEY_OK) { sides[j][k] = nullptr; failed = true; break; }
					Lane_registry::add(sides[j][k], true);
				}
			}
		}
		if (failed)
		{
			release_lanes();      // lanes created before the failure must not stay registered (they would skew free_ordered)
			for (auto& [_, r] : processor_resources)
			{
				r->exception = Processor::Runtime_error("No CUDA device", "The render engine needs a CUDA device; there is no CPU fallback.", nodey_last_error());
				r->state = State::Error;
			}
			done = true;
			return;
		}

		std::any fallback;
		// NODEY_TRACE=1: per-step wall time (host enqueue, then device drain) on stderr; serialises the steps
		const bool trace = schedule.trace;

		// run one batch per node class of `ids` on lane `lane`
		const auto run_group = [&](const std::vector<Id_t>& all_ids, int lane, int level_index) {
			std::map<std::type_index, std::vector<Id_t>> groups;
			for (const Id_t id : all_ids) groups[std::type_index(typeid(*processor_resources.at(id)->processor))].push_back(id);
			for (auto& [_, ids] : groups)
			{
				if (failed) return;
				Exec_context& ctx = Exec_context::current();
				ctx.stream = lanes[lane];
				for (int k = 0; k < kSides; k++) ctx.side_stream[k] = sides[k][lane];
				ctx.wave_begin = level_index == 0 ? &wave_begin : nullptr;
				ctx.level = level_index;
				ctx.lane = lane;
				ctx.stream_chunks = schedule.stream_chunks;

				std::vector<Processor::Batch_item> items;
				for (const Id_t id : ids)
				{
					Processor_resource& r = *processor_resources.at(id);
					if (r.stop_source) continue;
					const auto data = node_data.find(id);
					items.push_back(Processor::Batch_item{
						r.processor.get(), &r.input_payloads, &r.output_payloads, &r.stop_source,
						data == node_data.end() ? &fallback : data->second.get()});
					r.state = State::Running;
				}
				if (items.empty()) continue;

				bool batched = false;
				if (items.size() > 1)
				{
					// a batch failure is reported on every node of the batch
					Processor_resource& first = *processor_resources.at(ids.front());
					const bool ok = guarded(first, [&] { batched = items.front().processor->process_batch(items); });
					if (!ok)
					{
						for (const Id_t id : ids)
						{
							Processor_resource& r = *processor_resources.at(id);
							if (&r != &first) { r.exception = first.exception; r.state = State::Error; }
						}
						failed = true;
						return;
					}
				}
				if (batched)
				{
					for (const Id_t id : ids) processor_resources.at(id)->state = State::Finished;
					continue;
				}
				for (size_t k = 0; k < items.size() && !failed; k++)
				{
					Processor_resource& r = *processor_resources.at(ids[k]);
					const auto& it = items[k];
					if (guarded(r, [&] { it.processor->process_payload(*it.input, *it.output, *it.stop_token, *it.user_data); }))
						r.state = State::Finished;
					else
						failed = true;
				}
			}
		};

		// Schedule: sources first (their uploads are enqueued on the transfer lane in pin order), then wave
		// by wave, each wave level by level on the compute lane.  A wave is the part of the graph fed by a
		// contiguous block of source pins, so wave k computes while the uploads of wave k+1 are in flight.
		struct Step_events { Level_timing timing; nodey_event_t begin = nullptr, end = nullptr; bool ok = false; };
		std::vector<Step_events> steps;
		for (int wave = 0; wave <= max_wave && !failed; wave++)
			for (size_t level_index = wave == 0 ? 0 : 1; level_index < levels.size() && !failed; level_index++)
			{
				std::vector<Id_t> ids;
				for (const Id_t id : levels[level_index])
					if (node_wave.at(id) == wave) ids.push_back(id);
				if (ids.empty()) continue;
				const auto t_begin = std::chrono::steady_clock::now();
				const int lane = level_index == 0 ? 0 : 1 + wave % compute_lanes;
				Step_events se;
				se.timing.wave = wave; se.timing.level = (int)level_index; se.timing.lane = lane; se.timing.nodes = ids.size();
				se.timing.identifier = processor_resources.at(ids.front())->processor->get_processor_info_non_static().identifier;
				const bool timed = nodey_event_create(&se.begin, 1) == NODEY_OK && nodey_event_create(&se.end, 1) == NODEY_OK
								&& nodey_event_record(se.begin, lanes[lane]) == NODEY_OK;
				run_group(ids, lane, (int)level_index);
				if (g_release_products && !failed)
					for (const Id_t id : ids)
						for (auto& [pin, product] : processor_resources.at(id)->input_payloads)
							if (product) product->release();
				se.timing.enqueue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
				se.ok = timed && nodey_event_record(se.end, lanes[lane]) == NODEY_OK;
				steps.push_back(std::move(se));
				if (trace)
				{
					const auto t_enq = std::chrono::steady_clock::now();
					for (auto& s : lanes) if (s) nodey_stream_synchronize(s);
					for (auto& side : sides) for (auto& s : side) if (s) nodey_stream_synchronize(s);
					const auto t_end = std::chrono::steady_clock::now();
					fprintf(stderr, "[nodey trace] wave %d level %zu: %zu nodes (%s...), enqueue %.2f ms, drained after %.2f ms\n", wave, level_index,
							ids.size(), processor_resources.at(ids.front())->processor->get_processor_info_non_static().identifier.c_str(),
							std::chrono::duration<double, std::milli>(t_enq - t_begin).count(),
							std::chrono::duration<double, std::milli>(t_end - t_begin).count());
				}
			}
		for (auto& s : lanes)
			if (s) nodey_stream_synchronize(s);
		for (auto& side : sides)
			for (auto& s : side)
				if (s) nodey_stream_synchronize(s);
		// everything has drained: turn the step events into timings (the run's origin is the first step's begin)
		for (auto& se : steps)
		{
			float ms = 0.0f;
			if (se.ok && !failed && nodey_event_elapsed_ms(&ms, se.begin, se.end) == NODEY_OK) se.timing.device_ms = ms;
			if (se.ok && !failed && steps.front().ok && nodey_event_elapsed_ms(&ms, steps.front().begin, se.begin) == NODEY_OK) se.timing.start_ms = ms;
		}
		for (auto& se : steps)      // only now: the first step's begin event is the origin of every start_ms above
		{
			if (se.begin) nodey_event_destroy(se.begin);
			if (se.end) nodey_event_destroy(se.end);
			level_timings.push_back(std::move(se.timing));
		}
		release_lanes();
		done = true;
	}

	Runner::Schedule Runner::Schedule::from_environment()
	{
		Schedule s;
		const auto text = [](const char* name) -> const char* { const char* v = getenv(name); return v && *v ? v : nullptr; };
		if (const char* v = text("NODEY_WAVE")) s.wave_pins = std::max(1, atoi(v));
		if (const char* v = text("NODEY_WAVES"))
			for (const char* c = v; *c;)
			{
				s.wave_pattern.push_back(std::max(1, atoi(c)));
				while (*c && *c != ',') c++;
				if (*c == ',') c++;
			}
		if (const char* v = text("NODEY_COMPUTE_LANES")) s.compute_lanes = std::clamp(atoi(v), 1, 4);
		if (getenv("NODEY_NO_SIDE_STREAMS")) s.side_streams = 0;
		if (text("NODEY_NO_STREAM_PRIORITY")) s.stream_priority = 0;
		if (const char* v = text("NODEY_ST_CHUNKS")) s.stream_chunks = std::clamp(atoi(v), 1, 64);
		s.trace = getenv("NODEY_TRACE") != nullptr;
		s.timing = getenv("NODEY_ENGINE_TIMING") != nullptr;
		return s;
	}

	Runner::Schedule& Runner::Schedule::overlay(const Schedule& over)
	{
		if (over.wave_pins > 0) wave_pins = over.wave_pins;
		if (!over.wave_pattern.empty()) wave_pattern = over.wave_pattern;
		if (over.compute_lanes > 0) compute_lanes = over.compute_lanes;
		if (over.side_streams >= 0) side_streams = over.side_streams;
		if (over.stream_priority >= 0) stream_priority = over.stream_priority;
		if (over.stream_chunks > 0) stream_chunks = over.stream_chunks;
		trace = trace || over.trace;
		timing = timing || over.timing;
		return *this;
	}

	std::unique_ptr<Runner> Runner::create_and_run(const Graph& graph, std::map<Id_t, std::shared_ptr<std::any>> node_data)
	{
		return create_and_run(graph, std::move(node_data), Schedule::from_environment());
	}

	std::unique_ptr<Runner> Runner::create_and_run(const Graph& graph, std::map<Id_t, std::shared_ptr<std::any>> node_data, const Schedule& schedule)
	{
		auto runner = std::make_unique<Runner>();
		runner->schedule = schedule;
		runner->node_data = std::move(node_data);
		runner->generate_processor_resources(graph);
		if (nodey_get_device(&runner->device) != NODEY_OK) runner->device = -1;
		runner->worker = std::thread(&Runner::launch_threads, runner.get());
		return runner;
	}
}
