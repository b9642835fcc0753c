// CPU unit test of two host-side pieces that need no device: infra::Runner::Schedule (the scheduling knobs, their
// development overrides from the environment and the rule that an explicit setting wins) and processor::Frame_clock
// (the per-frame stamps of the reference's nodes: audio-amix.cpp:199-201, audio-velocity.cpp:238-249 / 313-318).
// Built and run by tests/test_host_graph.py.
#include "infra/runner.hpp"
#include "processor/audio-stream.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace infra;
using namespace processor;

static int failures = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

int main()
{
    // ---- Schedule ----
    for (const char* name : {"NODEY_WAVE", "NODEY_WAVES", "NODEY_COMPUTE_LANES", "NODEY_NO_SIDE_STREAMS", "NODEY_NO_STREAM_PRIORITY",
                             "NODEY_ST_CHUNKS", "NODEY_TRACE", "NODEY_ENGINE_TIMING"})
        unsetenv(name);
    Runner::Schedule automatic = Runner::Schedule::from_environment();
    CHECK(automatic.wave_pins == 0 && automatic.wave_pattern.empty() && automatic.compute_lanes == 0);
    CHECK(automatic.side_streams == -1 && automatic.stream_priority == -1 && automatic.stream_chunks == 0 && !automatic.trace && !automatic.timing);

    setenv("NODEY_WAVE", "16", 1);
    setenv("NODEY_WAVES", "32,64,48", 1);
    setenv("NODEY_COMPUTE_LANES", "9", 1);          // clamped
    setenv("NODEY_NO_SIDE_STREAMS", "1", 1);
    setenv("NODEY_ST_CHUNKS", "200", 1);            // clamped
    setenv("NODEY_TRACE", "1", 1);
    Runner::Schedule env = Runner::Schedule::from_environment();
    CHECK(env.wave_pins == 16 && env.wave_pattern == (std::vector<int>{32, 64, 48}) && env.compute_lanes == 4);
    CHECK(env.side_streams == 0 && env.stream_priority == -1 && env.stream_chunks == 64 && env.trace && !env.timing);

    Runner::Schedule over;                           // what a host set explicitly
    over.wave_pins = 8; over.compute_lanes = 2; over.side_streams = 1; over.stream_chunks = 4;
    Runner::Schedule merged = Runner::Schedule::from_environment().overlay(over);
    CHECK(merged.wave_pins == 8 && merged.compute_lanes == 2 && merged.side_streams == 1 && merged.stream_chunks == 4);
    CHECK(merged.wave_pattern == (std::vector<int>{32, 64, 48}) && merged.trace);        // untouched fields keep the environment's value
    setenv("NODEY_WAVES", "", 1);                    // empty = unset
    CHECK(Runner::Schedule::from_environment().wave_pattern.empty());

    // ---- Frame_clock ----
    {
        // end-time stamps, whole microseconds: 1024 / 48000 s = 21333.33 us -> 21333 us
        Frame_clock c(STAMP_END_US, 0.0, 48000);
        const double first = c.next(1024), second = c.next(1024);
        CHECK(first == 21333 * (1 / (double)1000000));
        CHECK(second == (double)(int64_t)((1024 / 48000.0 + 1024 / 48000.0) * 1000000) * (1 / (double)1000000));
    }
    {
        // float microseconds: after 600 s a float holds multiples of 64 us
        Frame_clock c(STAMP_START_FLOAT_US, 600.0, 44100);
        c.next(1152);
        const double second = c.next(1152);
        const double us = second * 1000000;
        CHECK(std::fabs(us / 64 - std::round(us / 64)) < 1e-6);
        CHECK(second != 600.0 + 1152 / 44100.0);
    }
    {
        Frame_clock c(STAMP_START, 0.25, 44100);
        CHECK(c.next(1000) == 0.25);
        CHECK(c.next(1000) == 0.25 + 1000 / 44100.0);
        auto list = std::make_shared<std::vector<double>>(std::vector<double>{0.5, 0.75});
        Frame_clock l(STAMP_LIST, 0.5, 48000, list);
        CHECK(l.next(10) == 0.5 && l.next(10) == 0.75);
        CHECK(l.next(10) == 0.5 + 20 / 48000.0);     // past the list: exact start times
    }
    if (failures) return 1;
    std::printf("schedule_clock_test ok\n");
    return 0;
}
