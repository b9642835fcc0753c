// runtime.cu -- device memory, streams and events behind the C ABI, so that the C++ host layer (and any
// reference-side binding) needs nothing but include/nodey_cuda.h: no CUDA headers, no torch.
#include "nodey_common.cuh"

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
    *out = nullptr;
    if (bytes == 0) return NODEY_OK;
    NODEY_CUDA_OK(cudaMallocAsync(out, bytes, as_stream(s)));
    return NODEY_OK;
}

int nodey_free(void* p, nodey_stream_t s)
{
    if (p) NODEY_CUDA_OK(cudaFreeAsync(p, as_stream(s)));
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
