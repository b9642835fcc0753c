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

		// Scheduling knobs of one run.  0 / -1 / empty = automatic: what the measurements behind DESIGN.md 2.1 chose on a
		// 148-SM B200 for renders shaped like BASELINE's (waves of 32 source pins on three compute lanes once the sources
		// have to be uploaded, one wave otherwise; 24 chunks along time).  A host with other shapes sets them here instead
		// of through the process environment.
		struct Schedule
		{
			int wave_pins = 0;                 // > 0: waves of that many source pins
			std::vector<int> wave_pattern;     // explicit wave sizes (the last one repeats); wins over wave_pins
			int compute_lanes = 0;             // 1..4 compute lanes the waves rotate over
			int side_streams = -1;             // 0: nodes of a chain run on their lane's one stream (no chunk-wise overlap)
			int stream_priority = -1;          // 0: the WSOLA search streams get no priority over the tails
			int stream_chunks = 0;             // 1..64 launches / copies a stream is cut into along time
			bool trace = false;                // per-step wall times on stderr (serialises the steps)
			bool timing = false;               // host phases of the set-up on stderr
			// development aid: NODEY_WAVE, NODEY_WAVES, NODEY_COMPUTE_LANES, NODEY_NO_SIDE_STREAMS, NODEY_NO_STREAM_PRIORITY,
			// NODEY_ST_CHUNKS, NODEY_TRACE, NODEY_ENGINE_TIMING -- read ONCE per call, never inside a run
			static Schedule from_environment();
			// fields set in `over` (non-automatic) replace this one's
			Schedule& overlay(const Schedule& over);
		};

		// builds one product per link, starts the worker thread and returns at once (reference semantics); the two-argument
		// form of the reference schedules by Schedule::from_environment()
		static std::unique_ptr<Runner> create_and_run(const Graph& graph, std::map<Id_t, std::shared_ptr<std::any>> node_data);
		static std::unique_ptr<Runner> create_and_run(const Graph& graph, std::map<Id_t, std::shared_ptr<std::any>> node_data, const Schedule& schedule);

		// sets every node's stop_source and joins the worker (src/infra/runner.cpp:53-63)
		~Runner();
		Runner() = default;
		Runner(Runner&&) = delete;
		Runner(const Runner&) = delete;
		Runner& operator=(Runner&&) = delete;
		Runner& operator=(const Runner&) = delete;

		enum class State
		{
			Ready,
			Running,
			Finished,
			Error
		};

		// what the UI polls per node: state, stop switch and -- in State::Error -- the exception the node threw
		struct Processor_resource
		{
			std::atomic<State> state = State::Ready;
			std::atomic<bool> stop_source = false;
			std::any exception;   // Processor::Runtime_error, std::runtime_error, std::logic_error or std::exception

			std::shared_ptr<Processor> processor;
			Processor::Output_map output_payloads;
			Processor::Input_map input_payloads;
		};

		// Memory policy of the runs created after the call (process wide, default: keep).  The reference keeps every link's
		// channel until the Runner dies; with `release` a link drops its payload as soon as its consumer has enqueued its
		// work, so a render holds one or two levels of intermediates per wave instead of all of them -- products other than
		// the sink's are then gone after the run (get_link_products() still lists the links).
		static void release_products(bool release);

		const auto& get_link_products() const { return link_products; }
		const auto& get_processor_resources() const { return processor_resources; }

		// ---- headless helpers (not in the reference, which polls the states from its UI loop) ----
		void wait();                                   // joins the worker thread
		bool finished() const { return done.load(); }
		// first error in node-id order as text, empty when every node finished
		std::string first_error() const;

		// Diagnostics.  The reference's overlay lists the fill of every link's 16-frame
		// channel (src/frontend/app.cpp:1556-1592); links here are published once, so what tells a user where the
		// time goes is the device time of every (wave, level) step: events recorded on the step's lane around its
		// enqueue.  device_ms is the span on that lane (it includes waiting for products of other lanes).
		struct Level_timing
		{
			int wave = 0, level = 0, lane = 0;
			size_t nodes = 0;
			std::string identifier;       // identifier of the step's first node class
			double enqueue_ms = 0.0;      // host time spent enqueueing the step
			double device_ms = 0.0;       // device span between the step's first and last command on its lane
			double start_ms = 0.0;        // device time from the run's first command to the step's first
		};
		// per-step device timings of the finished run (empty while running)
		std::vector<Level_timing> get_level_timings() const { return done.load() ? level_timings : std::vector<Level_timing>{}; }
		// the overlay's "Audio" block as text: node states, then one line per step
		std::string diagnostics_text() const;

	  private:

		void launch_threads();   // body of the worker thread: walks the levels
		void generate_processor_resources(const Graph& graph);

		std::thread worker;
		std::atomic<bool> done = false;
		int device = -1;     // CUDA device of the creating thread; the worker thread binds to it
		Schedule schedule;
		std::vector<std::vector<Id_t>> levels;
		std::map<Id_t, int> node_wave;          // which block of source pins feeds the node (see launch_threads)
		std::vector<int> wave_begin;            // first source-pin position of every wave
		std::vector<Level_timing> level_timings;
		std::map<Id_t, std::shared_ptr<std::any>> node_data;
		std::map<Id_t, std::shared_ptr<Processor::Product>> link_products;
		std::map<Id_t, std::shared_ptr<Processor_resource>> processor_resources;
	};
}
