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

		class Product;
		struct Info;
		// payload maps of one node: the product on each input pin, the products (one per link) on each output pin
		using Input_map = std::map<std::string, std::shared_ptr<Product>>;
		using Output_map = std::map<std::string, std::set<std::shared_ptr<Product>>>;

		// ---- registry (src/infra/processor.cpp:5, include/infra/processor.hpp:116-129) ----
		static std::map<std::string, Info> processor_map;

		template <typename T>
			requires(std::is_base_of_v<Processor, T> && Has_static_processor_info_func<T, Info>)
		static void register_processor() { add_to_registry(T::get_processor_info()); }

		// ---- what a node class implements ----
		Processor() = default;
		virtual ~Processor() = default;

		virtual Info get_processor_info_non_static() const = 0;
		struct Pin_attribute;
		virtual std::vector<Pin_attribute> get_pin_attributes() const = 0;
		virtual void deserialize(const Json::Value& value) = 0;
		virtual Json::Value serialize() const = 0;
		virtual bool draw_content(bool readonly) = 0;
		virtual void draw_title() = 0;

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

		// ---- nested types of the API ----
		// user-facing fault: short message, explanation, technical detail
		struct Runtime_error : public std::runtime_error
		{
			std::string message, explanation, detail;

			// what() reads "<message> (Detail: <detail>) (Explanation: <explanation>)" like the reference's (infra.cpp)
			Runtime_error(std::string message, std::string explanation, std::string detail = "");
		};

		// what travels over a link; every product kind derives from this
		class Product
		{
		  public:
			Product() = default;
			virtual ~Product() = default;
			const std::type_info& get_typeinfo() const { return typeid(*this); }
			// Optional (not in the reference): the Runner calls this once the consumer of the link has enqueued its work
			// and the run was asked to release intermediates (Runner::release_products); default: keep the payload
			virtual void release() {}
		};

		struct Info
		{
			std::string identifier;
			std::string display_name;
			bool singleton = false;
			std::function<std::unique_ptr<Processor>()> generate;
			std::string description;
		};

		struct Pin_attribute
		{
			std::string identifier;
			std::string display_name;
			std::reference_wrapper<const std::type_info> type;
			bool is_input;
			std::function<std::shared_ptr<Product>()> generate_func;
		};

	  private:

		// std::logic_error when the identifier is taken (infra.cpp)
		static void add_to_registry(Info info);
	};

	namespace detail
	{
		// lookups behind get_input_item / get_output_item (infra.cpp): nullptr when the key is absent from the input map;
		// std::logic_error for a null product, a product of another type, or a key absent from the output map
		std::shared_ptr<Processor::Product> input_product(const Processor::Input_map& input, const std::string& key,
														  const std::type_info& expected);
		const std::set<std::shared_ptr<Processor::Product>>& output_products(const Processor::Output_map& output, const std::string& key);
	}

	template <typename T>
	std::optional<std::reference_wrapper<T>> get_input_item(const Processor::Input_map& input, const std::string& key)
	{
		const auto product = detail::input_product(input, key, typeid(T));
		if (!product) return std::nullopt;
		return static_cast<T&>(*product);        // the type was checked by input_product
	}

	template <typename T>
	std::set<std::shared_ptr<T>> get_output_item(const Processor::Output_map& output, const std::string& key)
	{
		std::set<std::shared_ptr<T>> typed;
		for (const auto& product : detail::output_products(output, key)) typed.emplace(std::dynamic_pointer_cast<T>(product));
		return typed;
	}

	// registers every node class (src/register.cpp); idempotent in this engine so that a host process
	// may call it from several entry points
	void register_all_processors();
}
