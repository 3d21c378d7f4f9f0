/*
 * comm.cu - the communicator of a multi-GPU run, inside the C++ library (the reference has no multi-device path at all;
 * this is the B200 side of north_star's "source tree and equivalent particles replicated by an NCCL all-gather over NVLink").
 *
 * One context per GPU, either one process per GPU (onb_comm_init_rank: the ncclUniqueId travels through whatever the host
 * language has - torch.distributed in bench.py, a file, MPI) or one process driving all GPUs with one host thread per
 * context (onb_comm_init_all: the C++ drivers' -g=<n>). NCCL is loaded with dlopen at the first use, so single-GPU callers
 * neither link nor need it.
 *
 * ONE collective is all the hot path needs, IN PLACE on buffers that have the same layout on every rank:
 *   all-gather : rank r owns bytes [r*chunk, (r+1)*chunk) of every listed buffer  -> one ncclAllGather per buffer, grouped
 * (leaf records, source planes, sliced inputs as they lie; the equivalent strengths after packing them per rank, dist.cu).
 * It is enqueued on the communicator's own stream; the callers order them against the build / upward streams with events,
 * so that they overlap with the tree build of the other particle set and with the local part of the upward pass.
 *
 * A third transport, "loopback", joins contexts of ONE process (any devices, also all on the same one) with plain device
 * copies and a host barrier. It exists so that the whole distributed code path can be verified bit for bit on a single
 * GPU (tests/test_gpu_dist.py drives R contexts from R threads); it is not a product path.
 */
#include "onb_internal.h"
#include <dlfcn.h>
#include <condition_variable>
#include <mutex>
#include <cstdlib>
#include <cstring>

namespace {

// the few NCCL entry points, typed by hand (nccl.h is not needed to build the library)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
struct Nccl {
    int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
    std::string err;
};
Nccl g_nccl;
std::mutex g_nccl_mutex;
constexpr int NCCL_CHAR = 0;      // ncclInt8 / ncclChar

bool nccl_load(std::string& err) {
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (g_nccl.handle) return true;
    const char* names[] = { std::getenv("ONB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* nm : names) { if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!h) { err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return false; }
    struct { const char* name; void** fn; } syms[] = {
        {"ncclGetUniqueId", (void**)&g_nccl.GetUniqueId}, {"ncclCommInitRank", (void**)&g_nccl.CommInitRank}, {"ncclCommInitAll", (void**)&g_nccl.CommInitAll},
        {"ncclCommDestroy", (void**)&g_nccl.CommDestroy}, {"ncclGetErrorString", (void**)&g_nccl.GetErrorString}, {"ncclGroupStart", (void**)&g_nccl.GroupStart},
        {"ncclGroupEnd", (void**)&g_nccl.GroupEnd}, {"ncclAllGather", (void**)&g_nccl.AllGather},
        {"ncclGetVersion", (void**)&g_nccl.GetVersion} };
    for (auto& s : syms) { *s.fn = dlsym(h, s.name); if (!*s.fn) { err = std::string("NCCL symbol missing: ") + s.name; return false; } }
    g_nccl.handle = h;
    return true;
}

// ---- loopback transport: contexts of one process, host barrier + device copies -----------------
struct LoopGroup {
    int n = 0;
    std::mutex m; std::condition_variable cv;
    int arrived = 0; unsigned long long gen = 0;
    std::vector<std::vector<void*>> ptrs;         // per rank: the buffers of the collective in flight
    std::vector<cudaEvent_t> ready, done;         // per rank
    int refs = 0;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

}  // namespace

struct OnbComm {
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
    LoopGroup* loop = nullptr;
    cudaStream_t stream = nullptr;
};

#define ONB_NCCL(call) do { int r__ = (call); if (r__ != 0) { \
    c->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "NCCL error") + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
    return ONB_ERR_CUDA; } } while (0)

int onb_comm_rank(const onb_context* c) { return c->comm ? c->comm->rank : 0; }
int onb_comm_size(const onb_context* c) { return c->comm ? c->comm->nranks : 1; }
cudaStream_t onb_comm_stream(const onb_context* c) { return c->comm ? c->comm->stream : nullptr; }

// in-place all-gather of equal chunks: rank r owns [r*chunk_bytes, (r+1)*chunk_bytes) of every buffer
int onb_comm_allgather(onb_context* c, const std::vector<void*>& bufs, const std::vector<size_t>& chunk_bytes) {
    OnbComm* cm = c->comm;
    if (!cm || cm->nranks == 1 || bufs.empty()) return ONB_OK;
    if (cm->nccl) {
        ONB_NCCL(g_nccl.GroupStart());
        for (size_t k = 0; k < bufs.size(); ++k) {
            char* b = (char*)bufs[k];
            ONB_NCCL(g_nccl.AllGather(b + (size_t)cm->rank * chunk_bytes[k], b, chunk_bytes[k], NCCL_CHAR, cm->nccl, cm->stream));
        }
        ONB_NCCL(g_nccl.GroupEnd());
        return ONB_OK;
    }
    LoopGroup* g = cm->loop;
    ONB_CUDA(cudaEventRecord(g->ready[cm->rank], cm->stream));
    g->ptrs[cm->rank] = bufs;
    g->barrier();
    for (int q = 0; q < cm->nranks; ++q) {
        if (q == cm->rank) continue;
        ONB_CUDA(cudaStreamWaitEvent(cm->stream, g->ready[q], 0));
        for (size_t k = 0; k < bufs.size(); ++k)
            ONB_CUDA(cudaMemcpyAsync((char*)bufs[k] + (size_t)q * chunk_bytes[k], (const char*)g->ptrs[q][k] + (size_t)q * chunk_bytes[k], chunk_bytes[k], cudaMemcpyDefault, cm->stream));
    }
    ONB_CUDA(cudaEventRecord(g->done[cm->rank], cm->stream));
    g->barrier();
    for (int q = 0; q < cm->nranks; ++q) if (q != cm->rank) ONB_CUDA(cudaStreamWaitEvent(cm->stream, g->done[q], 0));   // my buffers stay untouched until every peer has read them
    g->barrier();                                                                                                       // (and the events are not re-recorded before every peer has enqueued its waits)
    return ONB_OK;
}

static int attach(onb_context* c, OnbComm* cm) {
    ONB_CUDA(cudaSetDevice(c->device));
    ONB_CUDA(cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking));
    if (!c->ev_src_planes) ONB_CUDA(cudaEventCreateWithFlags(&c->ev_src_planes, cudaEventDisableTiming));
    c->comm = cm;
    c->shard_rank = cm->rank; c->shard_n = cm->nranks;
    c->plan[0].valid = c->plan[1].valid = false;
    return ONB_OK;
}

extern "C" {

int onb_comm_unique_id(void* id, uint64_t bytes) {
    std::string err;
    if (bytes < sizeof(ncclUniqueId_t) || !nccl_load(err)) return ONB_ERR_UNSUPPORTED;
    return g_nccl.GetUniqueId((ncclUniqueId_t*)id) == 0 ? ONB_OK : ONB_ERR_CUDA;
}

int onb_comm_init_rank(onb_context* c, int rank, int nranks, const void* id, uint64_t bytes) {
    if (!c || nranks < 1 || nranks > ONB_MAX_RANKS || rank < 0 || rank >= nranks || bytes < sizeof(ncclUniqueId_t)) { if (c) c->err = "comm_init_rank: bad arguments"; return ONB_ERR_ARG; }
    if (c->comm) { c->err = "a communicator is already attached"; return ONB_ERR_ARG; }
    if (!nccl_load(c->err)) return ONB_ERR_UNSUPPORTED;
    ONB_CUDA(cudaSetDevice(c->device));
    OnbComm* cm = new OnbComm(); cm->rank = rank; cm->nranks = nranks;
    ncclUniqueId_t uid; memcpy(&uid, id, sizeof(uid));
    { int r = g_nccl.CommInitRank(&cm->nccl, nranks, uid, rank); if (r != 0) { c->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); delete cm; return ONB_ERR_CUDA; } }
    return attach(c, cm);
}

// one process, one context per device: ncclCommInitAll
int onb_comm_init_all(onb_context** ctxs, int n) {
    if (!ctxs || n < 1 || n > ONB_MAX_RANKS) return ONB_ERR_ARG;
    onb_context* c = ctxs[0];
    for (int i = 0; i < n; ++i) if (!ctxs[i] || ctxs[i]->comm) { c->err = "comm_init_all: null context or communicator already attached"; return ONB_ERR_ARG; }
    if (n == 1) return ONB_OK;
    if (!nccl_load(c->err)) return ONB_ERR_UNSUPPORTED;
    std::vector<int> devs(n); std::vector<ncclComm_t> comms(n, nullptr);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    ONB_NCCL(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        OnbComm* cm = new OnbComm(); cm->rank = i; cm->nranks = n; cm->nccl = comms[i];
        int rc = attach(ctxs[i], cm); if (rc) return rc;
    }
    return ONB_OK;
}

// test transport (see the header comment): contexts of this process, driven by one host thread each
int onb_comm_init_loopback(onb_context** ctxs, int n) {
    if (!ctxs || n < 1 || n > ONB_MAX_RANKS) return ONB_ERR_ARG;
    for (int i = 0; i < n; ++i) if (!ctxs[i] || ctxs[i]->comm) return ONB_ERR_ARG;
    LoopGroup* g = new LoopGroup(); g->n = n; g->ptrs.resize(n); g->ready.resize(n); g->done.resize(n); g->refs = n;
    for (int i = 0; i < n; ++i) {
        onb_context* c = ctxs[i];
        ONB_CUDA(cudaSetDevice(c->device));
        ONB_CUDA(cudaEventCreateWithFlags(&g->ready[i], cudaEventDisableTiming)); ONB_CUDA(cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming));
        OnbComm* cm = new OnbComm(); cm->rank = i; cm->nranks = n; cm->loop = g;
        int rc = attach(c, cm); if (rc) return rc;
    }
    return ONB_OK;
}

int onb_comm_destroy(onb_context* c) {
    if (!c || !c->comm) return ONB_OK;
    cudaSetDevice(c->device);
    OnbComm* cm = c->comm;
    if (cm->stream) { cudaStreamSynchronize(cm->stream); }
    if (cm->nccl) g_nccl.CommDestroy(cm->nccl);
    if (cm->loop) {
        LoopGroup* g = cm->loop;
        bool last;
        { std::lock_guard<std::mutex> lk(g->m); last = --g->refs == 0; }
        if (last) { for (auto e : g->ready) if (e) cudaEventDestroy(e); for (auto e : g->done) if (e) cudaEventDestroy(e); delete g; }
    }
    if (cm->stream) cudaStreamDestroy(cm->stream);
    delete cm; c->comm = nullptr;
    c->shard_rank = 0; c->shard_n = 1;
    c->plan[0].valid = c->plan[1].valid = false;
    return ONB_OK;
}

int onb_comm_info(const onb_context* c, int* rank, int* nranks, int* transport, int* nccl_version) {
    if (rank) *rank = onb_comm_rank(c);
    if (nranks) *nranks = onb_comm_size(c);
    if (transport) *transport = !c->comm ? 0 : (c->comm->nccl ? 1 : 2);
    if (nccl_version) { *nccl_version = 0; if (g_nccl.handle && g_nccl.GetVersion) g_nccl.GetVersion(nccl_version); }
    return ONB_OK;
}

}  // extern "C"
