// infra/graph.hpp -- node graph model and project-file (JSON) round trip.  Same public data layout,
// id allocation, validation rules and error types as the reference (include/infra/graph.hpp:21-192,
// src/infra/graph.cpp) so that existing project files load unchanged; the traversal in check_graph
// is a Kahn topological sort, which also yields the graph levels the Runner schedules by.
#pragma once

#include "processor.hpp"

#include <imgui.h>
#include <map>
#include <set>
#include <tuple>
#include <vector>

namespace infra
{
	class Graph
	{
	  public:

		// ---- project file: the JSON written by the editor (src/infra/graph.cpp:284-372) ----
		Json::Value serialize() const;
		// throws Invalid_file_error; node classes may throw Processor::Runtime_error from their own deserialize()
		static Graph deserialize(const Json::Value& value);

		// ---- validation ----
		// the four graph faults, with the reference's what() texts (include/infra/graph.hpp:89-134); constructors in infra.cpp
		struct Invalid_file_error : public std::runtime_error
		{
			std::string message;
			explicit Invalid_file_error(std::string message);
		};

		struct Multiple_input_error : public std::runtime_error
		{
			Id_t pin;
			explicit Multiple_input_error(Id_t pin);
		};

		struct Mismatched_pin_error : public std::runtime_error
		{
			Id_t from, to;
			Mismatched_pin_error(Id_t from, Id_t to);
		};

		struct Loop_detected_error : public std::runtime_error
		{
			Loop_detected_error();
		};

		// throws Mismatched_pin_error / Multiple_input_error / Loop_detected_error
		void check_graph() const;
		// check_graph() plus the result of the sort: level[k] = nodes whose inputs all lie in levels < k
		std::vector<std::vector<Id_t>> topological_levels() const;
		// true while the input pin has at most one link (so add_link accepts a second one and check_graph objects)
		bool check_multiple_input(Id_t pin_id) const;
		// both pins carry the same product type
		bool check_node_type_match(Id_t from, Id_t to) const;

		// ---- editing (ids are the smallest free ones, reused after removal) ----
		Id_t add_link(Id_t from, Id_t to);
		void remove_link(Id_t id);
		void remove_link(Id_t from, Id_t to);
		Id_t add_node(std::unique_ptr<Processor> processor);
		void remove_node(Id_t id);
		// rebuilds the node's pins from get_pin_attributes(); links are re-attached to pins of the same NAME and type
		void update_node_pin(Id_t id);

		std::map<Id_t, std::set<Id_t>> get_node_input_map() const;
		std::map<Id_t, Id_t> get_pin_to_node_map() const;

		// ---- the model itself: public data, as in the reference (the editor reads and writes it directly) ----
		struct Link
		{
			Id_t from, to;
			bool operator==(const Link& other) const { return std::tie(from, to) == std::tie(other.from, other.to); }
		};

		struct Pin
		{
			Id_t parent;
			Processor::Pin_attribute attribute;
		};

		struct Node
		{
			std::shared_ptr<Processor> processor;
			std::set<Id_t> pins;
			std::map<std::string, Id_t> pin_name_map;
			ImVec2 position = ImVec2(0, 0);
		};

		std::map<Id_t, Link> links;
		std::map<Id_t, Pin> pins;
		std::map<Id_t, Node> nodes;
		std::map<std::string, Id_t> singleton_node_map;
		bool modified = false;
	};
}
