// infra/processor.hpp -- the Processor plugin API of Nodey Audio Editor, kept as the drop-in boundary
// of the B200 engine (reference: include/infra/processor.hpp:26-176).  A node class written against
// the reference header compiles against this one: same nested types (Product, Pin_attribute, Info,
// Runtime_error), same virtuals, same registry and lookup helpers.  Additions are opt-in:
//   * process_batch(): the level-batched Runner hands all nodes of one class that are ready at the
//     same graph level to ONE call, so a node can render many tracks with one kernel launch;
//   * infra::Exec_context (exec-context.hpp): the CUDA stream a node must launch on.
#pragma once

#include <any>
#include <atomic>
#include <format>
#include <functional>
#include <json/json.h>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "utility/logic-error-utility.hpp"

namespace infra
{
	using Id_t = int;

	template <typename T, typename Ty>
	concept Has_static_processor_info_func = requires {
		{ T::get_processor_info() } -> std::same_as<Ty>;
	};

	class Processor
	{
	  public:

		// what travels over a link; every product kind derives from this
		class Product
		{
		  public:
			Product() = default;
			virtual ~Product() = default;
			const std::type_info& get_typeinfo() const { return typeid(*this); }
		};

		struct Pin_attribute
		{
			std::string identifier;
			std::string display_name;
			std::reference_wrapper<const std::type_info> type;
			bool is_input;
			std::function<std::shared_ptr<Product>()> generate_func;
		};

		struct Info
		{
			std::string identifier;
			std::string display_name;
			bool singleton = false;
			std::function<std::unique_ptr<Processor>()> generate;
			std::string description;
		};

		// user-facing fault: short message, explanation, technical detail
		struct Runtime_error : public std::runtime_error
		{
			std::string message, explanation, detail;

			Runtime_error(std::string message, std::string explanation, std::string detail = "") :
				std::runtime_error(std::format("{} (Detail: {}) (Explanation: {})", message, detail, explanation)),
				message(std::move(message)),
				explanation(std::move(explanation)),
				detail(std::move(detail))
			{
			}
		};

		static std::map<std::string, Processor::Info> processor_map;

		Processor() = default;
		virtual ~Processor() = default;

		virtual std::vector<Processor::Pin_attribute> get_pin_attributes() const = 0;
		virtual Processor::Info get_processor_info_non_static() const = 0;
		virtual Json::Value serialize() const = 0;
		virtual void deserialize(const Json::Value& value) = 0;
		virtual void draw_title() = 0;
		virtual bool draw_content(bool readonly) = 0;

		using Input_map = std::map<std::string, std::shared_ptr<Processor::Product>>;
		using Output_map = std::map<std::string, std::set<std::shared_ptr<Processor::Product>>>;

		virtual void process_payload(
			const Input_map& input,
			const Output_map& output,
			const std::atomic<bool>& stop_token,
			std::any& user_data
		) = 0;

		// One node of a batch: the processor instance and its payload maps.
		struct Batch_item
		{
			Processor* processor;
			const Input_map* input;
			const Output_map* output;
			const std::atomic<bool>* stop_token;
			std::any* user_data;
		};

		// Optional: render every item (all of this node class, same graph level) in one go.  Return false
		// to let the Runner fall back to one process_payload() call per node.
		virtual bool process_batch(const std::vector<Batch_item>& /*items*/) { return false; }

		// Optional: bytes this node will copy host -> device when it runs with `user_data` (source nodes).
		// The Runner pipelines the graph in waves over the source pins only when there is an upload to hide.
		virtual size_t upload_bytes(const std::any& /*user_data*/) const { return 0; }

		template <typename T>
			requires(std::is_base_of_v<Processor, T> && Has_static_processor_info_func<T, Processor::Info>)
		static void register_processor()
		{
			const Info processor_info = T::get_processor_info();
			if (processor_map.contains(processor_info.identifier))
				THROW_LOGIC_ERROR("Processor with identifier '{}' already registered", processor_info.identifier)
			processor_map[processor_info.identifier] = processor_info;
		}
	};

	template <typename T>
	std::optional<std::reference_wrapper<T>> get_input_item(const Processor::Input_map& input, const std::string& key)
	{
		const auto find = input.find(key);
		if (find == input.end()) return std::nullopt;
		if (find->second == nullptr) THROW_LOGIC_ERROR("Found nullptr in input map for key '{}'", key);
		if (find->second->get_typeinfo() != typeid(T))
			THROW_LOGIC_ERROR(
				"Type mismatch in input map for key '{}', expected {}, got {}",
				key, typeid(T).name(), find->second->get_typeinfo().name()
			);
		return *std::dynamic_pointer_cast<T>(find->second);
	}

	template <typename T>
	std::set<std::shared_ptr<T>> get_output_item(const Processor::Output_map& output, const std::string& key)
	{
		const auto find = output.find(key);
		if (find == output.end()) THROW_LOGIC_ERROR("Key '{}' not found in output map", key);
		std::set<std::shared_ptr<T>> output_set;
		for (auto& item : find->second)
		{
			if (item == nullptr) THROW_LOGIC_ERROR("Found nullptr in output map for key '{}'", key);
			output_set.emplace(std::dynamic_pointer_cast<T>(item));
		}
		return output_set;
	}

	// registers every node class (src/register.cpp); idempotent in this engine so that a host process
	// may call it from several entry points
	void register_all_processors();
}
