// bus.cu -- master-bus reduce across the GPUs of one box (SURVEY.md 8b `nodey_bus_reduce`, 8e).
//
// Tracks are independent until the master bus (the last audio_amix of the graph, audio-amix.cpp:293-307,
// is a sum): every rank renders its tracks into a partial bus and the partial buses are summed once.
// That sum is the only exchange step of the path, so it is the only collective: ncclReduce /
// ncclAllReduce(sum, float32) over NVLink, both planes of the FLTP bus in one NCCL group.
//
// NCCL is bound at run time (dlopen), not at link time: libnodey_cuda.so keeps the CUDA runtime as
// its only link dependency, single-GPU users never load NCCL, and inside a process that already
// carries an NCCL (torch's bundled libnccl.so.2) the loader hands back that same library instead of
// a second copy.  No NCCL header is needed: the handful of types the six calls use are restated
// below (nccl.h 2.x: ncclUniqueId = 128 opaque bytes, ncclFloat32 = 7, ncclSum = 0).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "nodey_common.cuh"

namespace {

struct NcclId { char bytes[NODEY_BUS_ID_BYTES]; };
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    char why[512] = {0};
};

NcclApi g_api;
std::once_flag g_once;

void load_nccl()
{
    // NODEY_NCCL_LIB: explicit library; otherwise the SONAME (an NCCL already in the process wins), then the dev name
    const char* names[3] = {getenv("NODEY_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.handle) break;
        snprintf(g_api.why, sizeof(g_api.why), "%s", dlerror());
    }
    if (!g_api.handle) return;
    bool ok = true;
    const auto sym = [&](const char* name) -> void* {
        void* p = dlsym(g_api.handle, name);
        if (!p) { ok = false; snprintf(g_api.why, sizeof(g_api.why), "NCCL library lacks %s", name); }
        return p;
    };
    g_api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    g_api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    g_api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
    g_api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    g_api.Reduce = (int (*)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t))sym("ncclReduce");
    g_api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
    g_api.GroupStart = (int (*)())sym("ncclGroupStart");
    g_api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!ok) { dlclose(g_api.handle); g_api.handle = nullptr; }
}

// 0 when NCCL is bound; NODEY_E_COMM with the loader's message otherwise (no fallback: the caller must know)
int need_nccl()
{
    std::call_once(g_once, load_nccl);
    if (g_api.handle) return NODEY_OK;
    nodey::set_error("NCCL is not available: %s", g_api.why[0] ? g_api.why : "libnccl.so.2 not found");
    return NODEY_E_COMM;
}

int nccl_fail(int rc, const char* what)
{
    nodey::set_error("%s failed: %s (NCCL result %d)", what, g_api.GetErrorString ? g_api.GetErrorString(rc) : "?", rc);
    return NODEY_E_COMM;
}

#define NODEY_NCCL_OK(expr)                                          \
    do {                                                             \
        const int _r = (expr);                                       \
        if (_r != kNcclSuccess) return nccl_fail(_r, #expr);         \
    } while (0)

}  // namespace

struct nodey_bus {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

extern "C" {

int nodey_bus_nccl_version(int* version)
{
    NODEY_REQUIRE(version, NODEY_E_INVALID, "nodey_bus_nccl_version: version must not be NULL");
    if (const int rc = need_nccl()) return rc;
    NODEY_NCCL_OK(g_api.GetVersion(version));
    return NODEY_OK;
}

int nodey_bus_unique_id(void* id)
{
    NODEY_REQUIRE(id, NODEY_E_INVALID, "nodey_bus_unique_id: id must point to NODEY_BUS_ID_BYTES bytes");
    if (const int rc = need_nccl()) return rc;
    NcclId u;
    NODEY_NCCL_OK(g_api.GetUniqueId(&u));
    memcpy(id, u.bytes, NODEY_BUS_ID_BYTES);
    return NODEY_OK;
}

int nodey_bus_create(nodey_bus** out, const void* id, int rank, int nranks)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_bus_create: out must not be NULL");
    *out = nullptr;
    NODEY_REQUIRE(id, NODEY_E_INVALID, "nodey_bus_create: id must point to the NODEY_BUS_ID_BYTES bytes rank 0 made");
    NODEY_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, NODEY_E_INVALID,
                  "nodey_bus_create: rank %d outside 0..%d", rank, nranks - 1);
    if (const int rc = need_nccl()) return rc;
    int dev = 0;
    NODEY_CUDA_OK(cudaGetDevice(&dev));
    NcclId u;
    memcpy(u.bytes, id, NODEY_BUS_ID_BYTES);
    NcclComm comm = nullptr;
    NODEY_NCCL_OK(g_api.CommInitRank(&comm, nranks, u, rank));
    nodey_bus* b = new nodey_bus;
    b->comm = comm; b->rank = rank; b->nranks = nranks; b->device = dev;
    *out = b;
    return NODEY_OK;
}

void nodey_bus_destroy(nodey_bus* bus)
{
    if (!bus) return;
    if (bus->comm && g_api.CommDestroy) g_api.CommDestroy(bus->comm);
    delete bus;
}

int nodey_bus_info(const nodey_bus* bus, int* rank, int* nranks, int* device)
{
    NODEY_REQUIRE(bus, NODEY_E_INVALID, "nodey_bus_info: bus must not be NULL");
    if (rank) *rank = bus->rank;
    if (nranks) *nranks = bus->nranks;
    if (device) *device = bus->device;
    return NODEY_OK;
}

static int bus_collective(const char* who, nodey_bus* bus, const float* const* send, float* const* recv, int nplanes,
                          int64_t nframes, int root, nodey_stream_t stream)
{
    NODEY_REQUIRE(bus && bus->comm, NODEY_E_INVALID, "%s: bus must not be NULL", who);
    NODEY_REQUIRE(nplanes >= 1 && nplanes <= 2 && send && recv, NODEY_E_INVALID, "%s: 1 or 2 planes", who);
    NODEY_REQUIRE(nframes >= 0, NODEY_E_INVALID, "%s: negative frame count", who);
    NODEY_REQUIRE(root < bus->nranks, NODEY_E_INVALID, "%s: root %d outside 0..%d", who, root, bus->nranks - 1);
    for (int p = 0; p < nplanes; p++) {
        NODEY_REQUIRE(send[p], NODEY_E_INVALID, "%s: send plane %d is NULL", who, p);
        // ncclReduce reads recv only on the root; every rank of an all-reduce needs one
        NODEY_REQUIRE(recv[p] || (root >= 0 && bus->rank != root), NODEY_E_INVALID, "%s: recv plane %d is NULL", who, p);
    }
    if (nframes == 0) return NODEY_OK;
    if (const int rc = need_nccl()) return rc;
    const cudaStream_t st = nodey::as_stream(stream);
    // both planes of the bus ride in one group: one launch, one NVLink schedule
    if (nplanes > 1) NODEY_NCCL_OK(g_api.GroupStart());
    int first_bad = kNcclSuccess;
    for (int p = 0; p < nplanes; p++) {
        const int r = root >= 0
            ? g_api.Reduce(send[p], recv[p], (size_t)nframes, kNcclFloat32, kNcclSum, root, bus->comm, st)
            : g_api.AllReduce(send[p], recv[p], (size_t)nframes, kNcclFloat32, kNcclSum, bus->comm, st);
        if (r != kNcclSuccess && first_bad == kNcclSuccess) first_bad = r;
    }
    if (nplanes > 1) {
        const int r = g_api.GroupEnd();           // always closed, also after a failed call inside the group
        if (r != kNcclSuccess && first_bad == kNcclSuccess) first_bad = r;
    }
    if (first_bad != kNcclSuccess) return nccl_fail(first_bad, who);
    return NODEY_OK;
}

int nodey_bus_reduce(nodey_bus* bus, const float* send_l, const float* send_r, float* recv_l, float* recv_r,
                     int64_t nframes, int root, nodey_stream_t stream)
{
    NODEY_REQUIRE(root >= 0, NODEY_E_INVALID, "nodey_bus_reduce: root must be a rank");
    const float* s[2] = {send_l, send_r};
    float* r[2] = {recv_l, recv_r};
    return bus_collective("nodey_bus_reduce", bus, s, r, send_r ? 2 : 1, nframes, root, stream);
}

int nodey_bus_allreduce(nodey_bus* bus, const float* send_l, const float* send_r, float* recv_l, float* recv_r,
                        int64_t nframes, nodey_stream_t stream)
{
    const float* s[2] = {send_l, send_r};
    float* r[2] = {recv_l, recv_r};
    return bus_collective("nodey_bus_allreduce", bus, s, r, send_r ? 2 : 1, nframes, -1, stream);
}

// ---- peer memory: the master mix as ONE kernel over NVLink (no collective) ------------------------------------
// A block made by nodey_peer_alloc can be exported to the other processes of the box (CUDA IPC); a rank that opened
// it holds an ordinary device pointer whose loads travel over NVLink.  nodey_mix does not care where its inputs
// live, so rank 0 can run the graph's master audio_amix over the group mixes of EVERY rank in the graph's input
// order: compute and exchange in one kernel, and the bus is bit identical to the one-GPU render (a reduce of
// partial buses adds in another order).  Ordering across processes is the caller's: a peer's block may be read
// once that peer has synchronised its writes and said so (barrier), and rewritten once the reader is done.
int nodey_peer_alloc(void** out, size_t bytes)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_peer_alloc: out must not be NULL");
    *out = nullptr;
    NODEY_CUDA_OK(cudaMalloc(out, bytes ? bytes : 1));      // plain cudaMalloc: stream-ordered pool memory cannot be exported this way
    return NODEY_OK;
}

int nodey_peer_free(void* p)
{
    if (p) NODEY_CUDA_OK(cudaFree(p));
    return NODEY_OK;
}

int nodey_peer_export(const void* p, void* handle)
{
    NODEY_REQUIRE(p && handle, NODEY_E_INVALID, "nodey_peer_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == NODEY_PEER_HANDLE_BYTES, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    NODEY_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
    memcpy(handle, &h, sizeof(h));
    return NODEY_OK;
}

int nodey_peer_open(void** out, const void* handle)
{
    NODEY_REQUIRE(out && handle, NODEY_E_INVALID, "nodey_peer_open: null argument");
    *out = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    NODEY_CUDA_OK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return NODEY_OK;
}

int nodey_peer_close(void* p)
{
    if (p) NODEY_CUDA_OK(cudaIpcCloseMemHandle(p));
    return NODEY_OK;
}

}  // extern "C"
