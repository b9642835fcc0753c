// infra/runner.hpp -- executes a Graph.  Public surface of the reference Runner (create_and_run,
// State, get_processor_resources, get_link_products; include/infra/runner.hpp:20-84) with the
// Boost.Fiber scheduler replaced by a level-batched one: the DAG is sorted into levels, every level
// is handed to the nodes in batches per node class, each batch enqueues its kernels on one CUDA
// stream of a small pool, and products carry the event downstream nodes wait on.  Host code never
// waits for the device between nodes; only sinks synchronise.
#pragma once

#include "graph.hpp"

#include <any>
#include <atomic>
#include <thread>

namespace infra
{
	class Runner
	{
	  public:

		enum class State
		{
			Ready,
			Running,
			Finished,
			Error
		};

		struct Processor_resource
		{
			std::shared_ptr<Processor> processor;
			Processor::Input_map input_payloads;
			Processor::Output_map output_payloads;

			std::atomic<bool> stop_source = false;
			std::atomic<State> state = State::Ready;
			std::any exception;   // Processor::Runtime_error, std::runtime_error, std::logic_error or std::exception
		};

	  private:

		std::map<Id_t, std::shared_ptr<Processor_resource>> processor_resources;
		std::map<Id_t, std::shared_ptr<Processor::Product>> link_products;
		std::map<Id_t, std::shared_ptr<std::any>> node_data;
		std::vector<std::vector<Id_t>> levels;
		std::map<Id_t, int> node_wave;          // which block of source pins feeds the node (see launch_threads)
		std::thread worker;
		std::atomic<bool> done = false;
		int device = -1;     // CUDA device of the creating thread; the worker thread binds to it

		void generate_processor_resources(const Graph& graph);
		void launch_threads();   // body of the worker thread: walks the levels

	  public:

		Runner() = default;
		Runner(const Runner&) = delete;
		Runner(Runner&&) = delete;
		Runner& operator=(const Runner&) = delete;
		Runner& operator=(Runner&&) = delete;
		~Runner();

		// builds one product per link, starts the worker thread and returns at once (reference semantics)
		static std::unique_ptr<Runner> create_and_run(const Graph& graph, std::map<Id_t, std::shared_ptr<std::any>> node_data);

		const auto& get_processor_resources() const { return processor_resources; }
		const auto& get_link_products() const { return link_products; }

		// headless helpers (not in the reference, which polls the states from its UI loop)
		void wait();                                   // joins the worker thread
		bool finished() const { return done.load(); }
		// first error in node-id order as text, empty when every node finished
		std::string first_error() const;
	};
}
