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

		struct Node
		{
			std::shared_ptr<Processor> processor;
			std::set<Id_t> pins;
			std::map<std::string, Id_t> pin_name_map;
			ImVec2 position = ImVec2(0, 0);
		};

		struct Pin
		{
			Id_t parent;
			Processor::Pin_attribute attribute;
		};

		struct Link
		{
			Id_t from, to;
			bool operator==(const Link& other) const { return std::tie(from, to) == std::tie(other.from, other.to); }
		};

		std::map<Id_t, Node> nodes;
		std::map<Id_t, Pin> pins;
		std::map<Id_t, Link> links;
		std::map<std::string, Id_t> singleton_node_map;
		bool modified = false;

	  private:

		// smallest id not present in the map (ids are reused after removal, like the reference)
		template <typename T>
		static Id_t find_empty(const std::map<Id_t, T>& list)
		{
			Id_t expect = 0;
			for (const auto& [id, _] : list)
			{
				if (id != expect) break;
				++expect;
			}
			return expect;
		}

	  public:

		struct Mismatched_pin_error : public std::runtime_error
		{
			Id_t from, to;
			Mismatched_pin_error(Id_t from, Id_t to) :
				std::runtime_error(std::format("Mismatch Pin: {}, {}", from, to)), from(from), to(to) {}
		};

		struct Loop_detected_error : public std::runtime_error
		{
			Loop_detected_error() : std::runtime_error("Loop Detected") {}
		};

		struct Multiple_input_error : public std::runtime_error
		{
			Id_t pin;
			Multiple_input_error(Id_t pin) :
				std::runtime_error(std::format("Multiple Inputs in Input Pin: {}", pin)), pin(pin) {}
		};

		struct Invalid_file_error : public std::runtime_error
		{
			std::string message;
			Invalid_file_error(std::string message) :
				std::runtime_error(std::format("Invalid File: {}", message)), message(std::move(message)) {}
		};

		Id_t add_node(std::unique_ptr<Processor> processor);
		void remove_node(Id_t id);
		void update_node_pin(Id_t id);
		Id_t add_link(Id_t from, Id_t to);
		void remove_link(Id_t id);
		void remove_link(Id_t from, Id_t to);

		std::map<Id_t, Id_t> get_pin_to_node_map() const;
		std::map<Id_t, std::set<Id_t>> get_node_input_map() const;

		// throws Mismatched_pin_error / Multiple_input_error / Loop_detected_error
		void check_graph() const;

		// check_graph() plus the result of the sort: level[k] = nodes whose inputs all lie in levels < k
		std::vector<std::vector<Id_t>> topological_levels() const;

		bool check_node_type_match(Id_t from, Id_t to) const
		{
			return &pins.at(from).attribute.type.get() == &pins.at(to).attribute.type.get();
		}

		// true while the input pin has at most one link
		bool check_multiple_input(Id_t pin_id) const
		{
			size_t count = 0;
			for (const auto& [_, link] : links)
				if (link.to == pin_id && ++count > 1) return false;
			return true;
		}

		Json::Value serialize() const;
		static Graph deserialize(const Json::Value& value);
	};
}
