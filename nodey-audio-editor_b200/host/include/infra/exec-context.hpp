// infra/exec-context.hpp -- what the level-batched Runner gives a node while it runs: the CUDA stream
// to launch on and helpers to order work across streams.  (No counterpart in the reference, whose
// nodes run as Boost fibers on one CPU thread; here a node only ENQUEUES device work.)
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

namespace infra
{
	// opaque handles so that node code does not need the CUDA headers
	using Stream_handle = void*;   // cudaStream_t
	using Event_handle = void*;    // cudaEvent_t

	struct Exec_context
	{
		Stream_handle stream = nullptr;   // launch everything of the current node (batch) on this stream
		// further streams of the lane (may be null): a node whose input arrives chunk by chunk (Audio_buffer::progress) runs
		// on the stream(s) AFTER its producer's in the cycle stream -> side[0] -> ... -> side[3] -> stream, so that a chain of
		// such nodes (resampler -> pitch -> tempo) overlaps stage by stage; a SoundTouch node takes two: one for its
		// sequential WSOLA search, which must never wait for anything but its input, one for the tails.  Products are
		// published with events recorded on those streams.
		static constexpr int kSideStreams = 4;
		Stream_handle side_stream[kSideStreams] = {nullptr, nullptr, nullptr, nullptr};
		// source level only: the pin positions at which the Runner's waves begin (ascending, first = 0; empty = one
		// wave).  A source that uploads chunk by chunk interleaves the chunks of the pins of one wave.
		const std::vector<int>* wave_begin = nullptr;
		int level = 0;                    // graph level being executed
		int lane = 0;                     // index of the stream inside the level
		int stream_chunks = 0;            // Runner::Schedule::stream_chunks of the run (0: the nodes' default)

		// the context of the calling thread (set by the Runner around process_payload / process_batch)
		static Exec_context& current();
	};

	// reference-counted CUDA event, recorded by a producer, waited on by consumers' streams
	class Device_event
	{
		Event_handle ev = nullptr;
	  public:
		Device_event();
		~Device_event();
		Device_event(const Device_event&) = delete;
		Device_event& operator=(const Device_event&) = delete;
		void record(Stream_handle stream);
		void wait_on(Stream_handle stream) const;   // cudaStreamWaitEvent
		void synchronize() const;
	};

	// The Runner's streams.  A render uses two: a transfer lane (source nodes: host->device uploads) and a
	// compute lane (every other node), ordered against each other by the events in the products.
	// Device memory is stream ordered: blocks are allocated on the lane of the node that creates them and
	// released through free_ordered(), which enqueues the free on the compute lane -- after every kernel
	// that could still read the block, without stalling anything (a free on the legacy default stream
	// would act as a barrier across all lanes) and immediately reusable by the next node's allocations.
	struct Lane_registry
	{
		static void add(Stream_handle lane, bool compute);
		static void remove(Stream_handle lane);
		static void free_ordered(void* ptr);
	};

	// owned device allocation (cudaMallocAsync on the current stream; freed stream-ordered)
	class Device_block
	{
	  public:
		void* ptr = nullptr;
		size_t bytes = 0;
		explicit Device_block(size_t bytes);
		~Device_block();
		Device_block(const Device_block&) = delete;
		Device_block& operator=(const Device_block&) = delete;
	};
}
