// runtime.cu -- device memory, streams and events behind the C ABI, so that the C++ host layer (and any
// reference-side binding) needs nothing but include/nodey_cuda.h: no CUDA headers, no torch.
#include "nodey_common.cuh"

#include <stdlib.h>
#include <map>
#include <mutex>
#include <vector>

// ---- caching device allocator -----------------------------------------------------------------------
// A render allocates the same (large) sizes in the same order step after step.  cudaMallocAsync's pool
// re-maps physical memory when sizes do not line up and stalls the enqueueing thread for hundreds of
// milliseconds on multi-GB requests, so large blocks are cached here instead: a freed block stays owned
// by the library, tagged with the stream it was freed on and an event; it is handed out again to the
// same stream at once (stream order makes that safe) or to another stream once the event has completed.
// Sizes are rounded up to 2 MiB; blocks are never split.  On out-of-memory every cached block is
// returned to the driver and the allocation retried.
namespace nodey {

namespace {
// A render runs on up to ten streams (transfer lane + three compute lanes with two side streams each).  The driver maps
// streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue see false dependencies --
// a lane waiting for an upload event then holds up another lane's chain.  Ask for 32 queues unless the host chose a value;
// this only takes effect when the library is loaded before the CUDA context exists (a host that initialises CUDA first
// sets the variable itself: bench.py, tests/conftest.py).
struct ConnectionsInit { ConnectionsInit() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); } } g_connections_init;

struct CachedBlock { void* ptr; size_t bytes; cudaStream_t stream; cudaEvent_t event; int device; };
std::mutex g_alloc_mu;
std::multimap<size_t, CachedBlock> g_free_blocks;           // by (rounded) size
std::map<void*, std::pair<size_t, int>> g_live_blocks;     // ptr -> (rounded size, device)
long long g_reserved_bytes = 0, g_peak_reserved = 0;       // bytes obtained from cudaMalloc and not returned (live + cached)
bool g_reuse_pending = false;                              // nodey_set_memory_policy: hand out blocks whose free point has not passed yet
long long g_live_bytes = 0, g_peak_bytes = 0;               // bytes of cached-allocator blocks handed out (guarded by g_alloc_mu)
void note_live_locked(long long delta) { g_live_bytes += delta; if (g_live_bytes > g_peak_bytes) g_peak_bytes = g_live_bytes; }
constexpr size_t kGranule = 2u << 20;
constexpr size_t kSmall = 1u << 20;                           // below this: cudaMallocAsync (pool handles it well)

void retain_small_pool()
{
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

void drop_cached_locked(int device)
{
    for (auto it = g_free_blocks.begin(); it != g_free_blocks.end();) {
        if (it->second.device == device) {
            cudaEventSynchronize(it->second.event);
            cudaEventDestroy(it->second.event);
            cudaFree(it->second.ptr);
            g_reserved_bytes -= (long long)it->second.bytes;
            it = g_free_blocks.erase(it);
        } else ++it;
    }
}
}  // namespace

int device_alloc(void** out, size_t bytes, cudaStream_t stream)
{
    *out = nullptr;
    if (bytes == 0) return NODEY_OK;
    if (bytes < kSmall) {
        retain_small_pool();
        NODEY_CUDA_OK(cudaMallocAsync(out, bytes, stream));
        return NODEY_OK;
    }
    int dev = 0;
    NODEY_CUDA_OK(cudaGetDevice(&dev));
    const size_t want = (bytes + kGranule - 1) / kGranule * kGranule;
    std::lock_guard<std::mutex> lock(g_alloc_mu);
    // smallest reusable block that fits with no more than 12.5 % waste; blocks freed on this very stream
    // are preferred (no event query needed: stream order already protects them)
    auto pick = g_free_blocks.end();
    for (auto it = g_free_blocks.lower_bound(want); it != g_free_blocks.end() && it->first <= want + want / 8; ++it) {
        const CachedBlock& b = it->second;
        if (b.device != dev) continue;
        if (b.stream == stream) { pick = it; break; }
        if (pick == g_free_blocks.end() && cudaEventQuery(b.event) == cudaSuccess) pick = it;
    }
    if (pick == g_free_blocks.end() && g_reuse_pending) {
        // memory-saving policy: a block another stream has freed but whose free point the device has not reached yet is
        // handed out too -- the requesting stream waits for that point first (stream-ordered reuse across streams)
        // (any block up to twice the size: in this mode a smaller footprint is worth more than a tight fit)
        for (auto it = g_free_blocks.lower_bound(want); it != g_free_blocks.end() && it->first <= 2 * want; ++it) {
            if (it->second.device != dev) continue;
            if (cudaStreamWaitEvent(stream, it->second.event, 0) == cudaSuccess) { pick = it; break; }
            cudaGetLastError();
        }
    }
    if (pick != g_free_blocks.end()) {
        const CachedBlock b = pick->second;
        g_free_blocks.erase(pick);
        cudaEventDestroy(b.event);
        g_live_blocks[b.ptr] = {b.bytes, dev};
        note_live_locked((long long)b.bytes);
        *out = b.ptr;
        return NODEY_OK;
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        drop_cached_locked(dev);
        e = cudaMalloc(out, want);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
    g_reserved_bytes += (long long)want;
    if (g_reserved_bytes > g_peak_reserved) g_peak_reserved = g_reserved_bytes;
    g_live_blocks[*out] = {want, dev};
    note_live_locked((long long)want);
    return NODEY_OK;
}

int device_free(void* p, cudaStream_t stream)
{
    if (!p) return NODEY_OK;
    {
        std::lock_guard<std::mutex> lock(g_alloc_mu);
        const auto it = g_live_blocks.find(p);
        if (it != g_live_blocks.end()) {
            CachedBlock b{p, it->second.first, stream, nullptr, it->second.second};
            g_live_blocks.erase(it);
            note_live_locked(-(long long)b.bytes);
            if (cudaEventCreateWithFlags(&b.event, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(b.event, stream) != cudaSuccess) {
                cudaGetLastError();
                cudaStreamSynchronize(stream);
                cudaFree(p);
                g_reserved_bytes -= (long long)b.bytes;
                return NODEY_OK;
            }
            g_free_blocks.emplace(b.bytes, b);
            return NODEY_OK;
        }
    }
    NODEY_CUDA_OK(cudaFreeAsync(p, stream));
    return NODEY_OK;
}

}  // namespace nodey

using namespace nodey;

extern "C" {

int nodey_set_device(int ordinal)
{
    NODEY_CUDA_OK(cudaSetDevice(ordinal));
    return NODEY_OK;
}

int nodey_get_device(int* ordinal)
{
    NODEY_REQUIRE(ordinal, NODEY_E_INVALID, "nodey_get_device: null argument");
    NODEY_CUDA_OK(cudaGetDevice(ordinal));
    return NODEY_OK;
}

int nodey_device_count(int* count)
{
    NODEY_REQUIRE(count, NODEY_E_INVALID, "nodey_device_count: null argument");
    NODEY_CUDA_OK(cudaGetDeviceCount(count));
    return NODEY_OK;
}

int nodey_device_synchronize(void)
{
    NODEY_CUDA_OK(cudaDeviceSynchronize());
    return NODEY_OK;
}

int nodey_stream_create(nodey_stream_t* out)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_stream_create: null argument");
    cudaStream_t s;
    NODEY_CUDA_OK(cudaStreamCreate(&s));    // blocking stream: ordered against the legacy default stream
    *out = (nodey_stream_t)s;
    return NODEY_OK;
}

/* high_priority != 0: the stream's kernels are picked first when SM slots free up (cudaStreamCreateWithPriority) -- for
 * latency-critical chains that share the device with bulk kernels */
int nodey_stream_create_priority(nodey_stream_t* out, int high_priority)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_stream_create_priority: null argument");
    int lo = 0, hi = 0;
    NODEY_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // numerically lower = higher priority
    cudaStream_t s;
    NODEY_CUDA_OK(cudaStreamCreateWithPriority(&s, cudaStreamDefault, high_priority ? hi : lo));
    *out = (nodey_stream_t)s;
    return NODEY_OK;
}

int nodey_stream_destroy(nodey_stream_t s)
{
    if (s) NODEY_CUDA_OK(cudaStreamDestroy(as_stream(s)));
    return NODEY_OK;
}

int nodey_stream_synchronize(nodey_stream_t s)
{
    NODEY_CUDA_OK(cudaStreamSynchronize(as_stream(s)));
    return NODEY_OK;
}

int nodey_event_create(nodey_event_t* out, int timing)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_event_create: null argument");
    cudaEvent_t e;
    NODEY_CUDA_OK(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
    *out = (nodey_event_t)e;
    return NODEY_OK;
}

int nodey_event_destroy(nodey_event_t e)
{
    if (e) NODEY_CUDA_OK(cudaEventDestroy((cudaEvent_t)e));
    return NODEY_OK;
}

int nodey_event_record(nodey_event_t e, nodey_stream_t s)
{
    NODEY_CUDA_OK(cudaEventRecord((cudaEvent_t)e, as_stream(s)));
    return NODEY_OK;
}

int nodey_event_synchronize(nodey_event_t e)
{
    NODEY_CUDA_OK(cudaEventSynchronize((cudaEvent_t)e));
    return NODEY_OK;
}

int nodey_event_elapsed_ms(float* ms, nodey_event_t a, nodey_event_t b)
{
    NODEY_REQUIRE(ms, NODEY_E_INVALID, "nodey_event_elapsed_ms: null argument");
    NODEY_CUDA_OK(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return NODEY_OK;
}

int nodey_stream_wait_event(nodey_stream_t s, nodey_event_t e)
{
    NODEY_CUDA_OK(cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)e, 0));
    return NODEY_OK;
}

int nodey_malloc(void** out, size_t bytes, nodey_stream_t s)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_malloc: null argument");
    return device_alloc(out, bytes, as_stream(s));
}

int nodey_free(void* p, nodey_stream_t s) { return device_free(p, as_stream(s)); }

/* return every cached block of the current device to the driver (synchronises) */
int nodey_trim_memory(void)
{
    int dev = 0;
    NODEY_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_alloc_mu);
    drop_cached_locked(dev);
    return NODEY_OK;
}

/* bytes of device memory the library's callers hold right now (blocks of 1 MiB and more, which is where a render's
 * footprint is) and the high-water mark since the last reset */
int nodey_memory_stats(int64_t* live_bytes, int64_t* peak_bytes, int reset_peak)
{
    std::lock_guard<std::mutex> lock(g_alloc_mu);
    if (live_bytes) *live_bytes = g_live_bytes;
    if (peak_bytes) *peak_bytes = g_peak_bytes;
    if (reset_peak) { g_peak_bytes = g_live_bytes; g_peak_reserved = g_reserved_bytes; }
    return NODEY_OK;
}

int nodey_memory_reserved(int64_t* reserved_bytes, int64_t* peak_reserved_bytes)
{
    std::lock_guard<std::mutex> lock(g_alloc_mu);
    if (reserved_bytes) *reserved_bytes = g_reserved_bytes;
    if (peak_reserved_bytes) *peak_reserved_bytes = g_peak_reserved;
    return NODEY_OK;
}

int nodey_set_memory_policy(int reuse_pending)
{
    std::lock_guard<std::mutex> lock(g_alloc_mu);
    g_reuse_pending = reuse_pending != 0;
    return NODEY_OK;
}

int nodey_memset(void* dst, int value, size_t bytes, nodey_stream_t s)
{
    if (bytes) NODEY_CUDA_OK(cudaMemsetAsync(dst, value, bytes, as_stream(s)));
    return NODEY_OK;
}

int nodey_memcpy_h2d(void* dst, const void* src, size_t bytes, nodey_stream_t s)
{
    if (bytes) NODEY_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return NODEY_OK;
}

int nodey_memcpy2d_h2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, nodey_stream_t s)
{
    if (width_bytes && rows)
        NODEY_CUDA_OK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice, as_stream(s)));
    return NODEY_OK;
}

int nodey_memcpy_d2h(void* dst, const void* src, size_t bytes, nodey_stream_t s)
{
    if (bytes) NODEY_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return NODEY_OK;
}

int nodey_memcpy_d2d(void* dst, const void* src, size_t bytes, nodey_stream_t s)
{
    if (bytes) NODEY_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return NODEY_OK;
}

int nodey_host_alloc(void** out, size_t bytes)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_host_alloc: null argument");
    NODEY_CUDA_OK(cudaMallocHost(out, bytes ? bytes : 1));
    return NODEY_OK;
}

int nodey_host_free(void* p)
{
    if (p) NODEY_CUDA_OK(cudaFreeHost(p));
    return NODEY_OK;
}

}  // extern "C"
