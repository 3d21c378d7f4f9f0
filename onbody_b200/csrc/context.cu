/*
 * context.cu - the C ABI (include/onbody_b200.h): device context, host<->device copies, and the phase
 * sequence of the reference drivers (ongrav3d.cpp:600-908) expressed as calls into the CUDA kernels.
 * Host-side C++ only orchestrates; there is no CPU compute path and every entry point fails loudly when the
 * device is missing.
 */
#include "onb_internal.h"
#include <cstdlib>

int onb_comm_allgather(onb_context* c, const std::vector<void*>& bufs, const std::vector<size_t>& chunk_bytes);
cudaStream_t onb_comm_stream(const onb_context* c);

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

static std::string g_create_error;

// ---------------------------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------------------------
int onb_alloc_parts(onb_context* c, DParts& p, uint32_t n, bool are_sources) {
    p = DParts();
    // slack: bulk copies round up to 16 B; in-place all-gathers of equal leaf-aligned chunks overrun n by < one leaf per rank
    p.n = n; p.cap = ((n + 63u) & ~31u) + 288u + (uint32_t)(ONB_MAX_RANKS + 1) * (uint32_t)std::max(c->block, 128); p.PD = c->PD; p.SD = c->SD; p.OD = c->OD; p.are_sources = are_sources;
    const size_t bytes = (size_t)p.cap * sizeof(float);
    // the zero fills go on the stream of the work being enqueued (ONB_ST): onb_prepare_eval allocates the equivalent target
    // points while its target chain runs on the second stream, and a fill ordered on the context stream could land after
    // k_upward had written them
    cudaStream_t st = ONB_ST(c);
    for (int d = 0; d < c->PD; ++d) { ONB_CUDA(onb_pmalloc(c, (void**)&p.x[d], bytes)); ONB_CUDA(cudaMemsetAsync(p.x[d], 0, bytes, st)); }
    ONB_CUDA(onb_pmalloc(c, (void**)&p.r, bytes)); ONB_CUDA(cudaMemsetAsync(p.r, 0, bytes, st));
    if (are_sources) {
        for (int d = 0; d < c->SD; ++d) { ONB_CUDA(onb_pmalloc(c, (void**)&p.s[d], bytes)); ONB_CUDA(cudaMemsetAsync(p.s[d], 0, bytes, st)); }
        ONB_CUDA(onb_pmalloc(c, (void**)&p.pk0, (size_t)p.cap * sizeof(float4)));
        const bool nf2 = c->physics == ONB_VORT3D || c->physics == ONB_VORTGRAD3D;
        if (nf2) ONB_CUDA(onb_pmalloc(c, (void**)&p.pk1, (size_t)p.cap * sizeof(float4)));
        if (c->physics == ONB_GRAV3D) ONB_CUDA(onb_pmalloc(c, (void**)&p.pk2, bytes));
    } else if (c->mem_mode == ONB_MEM_LEAN && c->shard_n > 1 && &p == &c->parts[1]) {
        // lean memory mode: the output planes keep their full index space but only this rank's shard is backed by memory
        uint64_t lo = 0, hi = 0; onb_shard_range_for(n, c->block, c->shard_rank, c->shard_n, &lo, &hi);
        std::vector<std::pair<size_t, size_t>> ranges(1, std::make_pair((size_t)lo * sizeof(float), (size_t)(hi - lo) * sizeof(float)));
        for (int d = 0; d < c->OD; ++d) ONB_CUDA(onb_sparse_alloc(c, (void**)&p.u[d], bytes, ranges, st));
        p.sparse_key = c->plan_key(1); p.u_lo = (uint32_t)lo; p.u_hi = (uint32_t)hi;
    } else {
        for (int d = 0; d < c->OD; ++d) { ONB_CUDA(onb_pmalloc(c, (void**)&p.u[d], bytes)); ONB_CUDA(cudaMemsetAsync(p.u[d], 0, bytes, st)); }
    }
    if (!are_sources && c->accum64) {
        if (p.sparse_key) { c->err = "ACCUM = double is not available together with the lean memory mode of a sharded run"; return ONB_ERR_UNSUPPORTED; }
        for (int d = 0; d < c->OD; ++d) { ONB_CUDA(onb_pmalloc(c, (void**)&p.ud[d], 2 * bytes)); ONB_CUDA(cudaMemsetAsync(p.ud[d], 0, 2 * bytes, st)); }
    }
    return ONB_OK;
}
// lean memory mode: once the float4 tiles exist nothing on the evaluation path reads the SoA planes of a source set again
static void release_unpacked(onb_context* c, DParts& p) {
    if (!p.are_sources || !p.packed_valid) return;
    for (int d = 0; d < ONB_MAX_PD; ++d) { if (p.x[d]) onb_pfree(c, p.x[d]); p.x[d] = nullptr; }
    if (p.r) onb_pfree(c, p.r); p.r = nullptr;
    for (int d = 0; d < ONB_MAX_SD; ++d) { if (p.s[d]) onb_pfree(c, p.s[d]); p.s[d] = nullptr; }
    p.unpacked_released = true;
}
static int lean_after_prepare(onb_context* c) {
    if (c->mem_mode != ONB_MEM_LEAN) return ONB_OK;
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    release_unpacked(c, c->parts[0]); release_unpacked(c, c->parts[2]);
    onb_scratch_trim(c);
    return ONB_OK;
}
void onb_free_parts(onb_context* c, DParts& p) {
    for (int d = 0; d < ONB_MAX_PD; ++d) if (p.x[d]) onb_pfree(c, p.x[d]);
    if (p.r) onb_pfree(c, p.r);
    for (int d = 0; d < ONB_MAX_SD; ++d) if (p.s[d]) onb_pfree(c, p.s[d]);
    for (int d = 0; d < ONB_MAX_OD; ++d) { if (p.u[d]) onb_pfree(c, p.u[d]); if (p.ud[d]) onb_pfree(c, p.ud[d]); }
    if (p.gidx) onb_pfree(c, p.gidx);
    if (p.gidx_spare) onb_pfree(c, p.gidx_spare);
    if (p.pk0) onb_pfree(c, p.pk0); if (p.pk1) onb_pfree(c, p.pk1); if (p.pk2) onb_pfree(c, p.pk2);
    p = DParts();
}
static inline uint32_t host_log2(uint32_t x) { return x == 0 ? 0 : 31 - __builtin_clz(x); }

int onb_alloc_tree(onb_context* c, DTree& t, uint32_t n, int block) {
    const uint32_t numLeaf = 1 + (n - 1) / block;                                         // Tree.hpp:83-87
    const int levels = 1 + host_log2(2 * numLeaf - 1);
    if (t.numnodes == (1 << levels) && t.nr) {          // same shape as last time: reuse the arrays, just clear them
        const size_t fb0 = (size_t)t.numnodes * sizeof(float);
        for (int d = 0; d < c->PD; ++d) { ONB_CUDA(cudaMemsetAsync(t.x[d], 0, fb0, c->stream)); ONB_CUDA(cudaMemsetAsync(t.nc[d], 0, fb0, c->stream)); ONB_CUDA(cudaMemsetAsync(t.ns[d], 0, fb0, c->stream)); }
        ONB_CUDA(cudaMemsetAsync(t.nr, 0, fb0, c->stream)); ONB_CUDA(cudaMemsetAsync(t.pr, 0, fb0, c->stream));
        for (int d = 0; d < c->SD; ++d) ONB_CUDA(cudaMemsetAsync(t.s[d], 0, fb0, c->stream));
        ONB_CUDA(cudaMemsetAsync(t.ioffset, 0, fb0, c->stream)); ONB_CUDA(cudaMemsetAsync(t.num, 0, fb0, c->stream));
        t.built = false;
        return ONB_OK;
    }
    onb_free_tree(c, t);
    t.levels = levels;
    t.numnodes = 1 << t.levels;
    const size_t fb = (size_t)t.numnodes * sizeof(float), ub = (size_t)t.numnodes * sizeof(uint32_t);
    for (int d = 0; d < c->PD; ++d) {
        ONB_CUDA(onb_pmalloc(c, (void**)&t.x[d], fb)); ONB_CUDA(onb_pmalloc(c, (void**)&t.nc[d], fb)); ONB_CUDA(onb_pmalloc(c, (void**)&t.ns[d], fb));
        ONB_CUDA(cudaMemsetAsync(t.x[d], 0, fb, c->stream)); ONB_CUDA(cudaMemsetAsync(t.nc[d], 0, fb, c->stream)); ONB_CUDA(cudaMemsetAsync(t.ns[d], 0, fb, c->stream));
    }
    ONB_CUDA(onb_pmalloc(c, (void**)&t.nr, fb)); ONB_CUDA(onb_pmalloc(c, (void**)&t.pr, fb));
    ONB_CUDA(cudaMemsetAsync(t.nr, 0, fb, c->stream)); ONB_CUDA(cudaMemsetAsync(t.pr, 0, fb, c->stream));
    for (int d = 0; d < c->SD; ++d) { ONB_CUDA(onb_pmalloc(c, (void**)&t.s[d], fb)); ONB_CUDA(cudaMemsetAsync(t.s[d], 0, fb, c->stream)); }
    ONB_CUDA(onb_pmalloc(c, (void**)&t.ioffset, ub)); ONB_CUDA(onb_pmalloc(c, (void**)&t.num, ub));
    ONB_CUDA(cudaMemsetAsync(t.ioffset, 0, ub, c->stream)); ONB_CUDA(cudaMemsetAsync(t.num, 0, ub, c->stream));
    return ONB_OK;
}
void onb_free_tree(onb_context* c, DTree& t) {
    for (int d = 0; d < ONB_MAX_PD; ++d) { if (t.x[d]) onb_pfree(c, t.x[d]); if (t.nc[d]) onb_pfree(c, t.nc[d]); if (t.ns[d]) onb_pfree(c, t.ns[d]); }
    if (t.nr) onb_pfree(c, t.nr); if (t.pr) onb_pfree(c, t.pr);
    for (int d = 0; d < ONB_MAX_SD; ++d) if (t.s[d]) onb_pfree(c, t.s[d]);
    if (t.ioffset) onb_pfree(c, t.ioffset); if (t.num) onb_pfree(c, t.num);
    t = DTree();
}
int onb_join_copies(onb_context* c) {
    if (c->tgt_copy_pending) {
        c->tgt_copy_pending = false;
        ONB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_tgt_ready, 0));
    }
    return ONB_OK;
}
int onb_check_flag(onb_context* c, const char* what) {
    ONB_CUDA(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    const int f = *c->h_flag;
    if (f != 0) {
        c->err = std::string(what);
        ONB_CUDA(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
        return f;
    }
    return ONB_OK;
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* onb_last_create_error(void) { return g_create_error.c_str(); }

onb_context* onb_create(int physics, int device) {
    if (physics < 0 || physics > 4) { g_create_error = "unknown physics id"; return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        return nullptr;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return nullptr;
    }
    if (prop.major < 10) { g_create_error = "device is not sm_100 class (kernels are built for sm_100a only)"; return nullptr; }
    onb_context* c = new onb_context();
    static const int PDs[5] = {3, 3, 3, 2, 2}, SDs[5] = {1, 3, 3, 1, 1}, ODs[5] = {3, 3, 12, 2, 2}, FL[5] = {19, 28, 64, 13, 15};
    c->physics = physics; c->device = device; c->PD = PDs[physics]; c->SD = SDs[physics]; c->OD = ODs[physics];
    c->flops_per_pair = FL[physics]; c->has_tr = physics == ONB_VORT2DTR; c->has_fastsumm = physics != ONB_VORTGRAD3D;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c->d_flag, sizeof(int)) != cudaSuccess || cudaMallocHost(&c->h_flag, sizeof(int)) != cudaSuccess) {
        g_create_error = "context allocation failed"; delete c; return nullptr;
    }
    cudaMemset(c->d_flag, 0, sizeof(int));
    cudaMalloc(&c->d_build_stats, 16 * sizeof(unsigned long long)); cudaMemset(c->d_build_stats, 0, 16 * sizeof(unsigned long long));
    cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
    onb_set_params(c, 128, 4, ONB_ARITH_FAST);
    return c;
}

void onb_destroy(onb_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 4; ++i) onb_free_parts(c, c->parts[i]);
    for (int i = 0; i < 2; ++i) onb_free_tree(c, c->trees[i]);
    cudaStreamSynchronize(c->stream);
    if (c->d_flag) cudaFree(c->d_flag);
    if (c->d_build_stats) cudaFree(c->d_build_stats);
    onb_comm_destroy(c);
    for (int k = 0; k < 2; ++k) { if (c->d_shared[k]) cudaFree(c->d_shared[k]); if (c->rec_buf[k]) cudaFree(c->rec_buf[k]); }
    if (c->ev_src_planes) cudaEventDestroy(c->ev_src_planes);
    if (c->eq_stage) cudaFree(c->eq_stage);
    if (c->d_epnum) cudaFree(c->d_epnum);
    if (c->dtt_pool) cudaFree(c->dtt_pool);
    for (auto& e : c->ev_cache) if (e) cudaEventDestroy(e);
    for (auto& sl : c->slabs) cudaFree(sl.p);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (int k = 0; k < 2; ++k) if (c->ev_stage[k]) cudaEventDestroy(c->ev_stage[k]);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->ev_tgt_ready) cudaEventDestroy(c->ev_tgt_ready);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* onb_error(const onb_context* c) { return c ? c->err.c_str() : "null context"; }

int onb_set_params(onb_context* c, int block_size, int order, int arith) {
    if (!c) return ONB_ERR_ARG;
    if (block_size < 2) { c->err = "block size must be >= 2"; return ONB_ERR_ARG; }
    block_size = 2 * ((block_size + 1) / 2);            // the reference rounds -b up to even (ongrav3d.cpp:522, minBlkSz 2)
    if (block_size > 128) { c->err = "block sizes above 128 are not supported by the GPU build"; return ONB_ERR_UNSUPPORTED; }
    if (order == 0) { c->err = "order 0 is not a valid barycentric order (the drivers reject -o=0, ongrav3d.cpp:516)"; return ONB_ERR_ARG; }
    if (order < 0) {
        // the drivers' default when -o is omitted (ongrav3d.cpp:481): hierarchical pair-merge equivalents, up to blockSize per node
        c->block = block_size; c->order = -1; c->arith = arith; c->ncp = 0; c->num_eqps = block_size; c->ebs = 128; c->legacy = true;
        return ONB_OK;
    }
    if (order > ONB_MAX_ORDER) { c->err = "order above 20"; return ONB_ERR_ARG; }
    int ne = 1; for (int d = 0; d < c->PD; ++d) ne *= (order + 1);
    if (ne > 128) { c->err = "(order+1)^PD exceeds the 128-slot equivalent block of the GPU build"; return ONB_ERR_UNSUPPORTED; }
    c->block = block_size; c->order = order; c->arith = arith; c->ncp = order + 1; c->num_eqps = ne; c->ebs = 128; c->legacy = false;
    return ONB_OK;
}

void onb_dims(const onb_context* c, int* pd, int* sd, int* od, int* hf) { *pd = c->PD; *sd = c->SD; *od = c->OD; *hf = c->has_fastsumm; }

void onb_set_flops_per_pair(onb_context* c, int flops) { if (c && flops > 0) c->flops_per_pair = flops; }

int onb_set_shard(onb_context* c, int rank, int nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks) { c->err = "bad shard"; return ONB_ERR_ARG; }
    if (c->comm && (rank != c->shard_rank || nranks != c->shard_n)) { c->err = "the shard of a context with a communicator is its rank"; return ONB_ERR_ARG; }
    c->shard_rank = rank; c->shard_n = nranks;
    c->plan[0].valid = c->plan[1].valid = false;
    return ONB_OK;
}
// ACCUM of the reference drivers (ongrav3d.cpp:7-8, README.md:107-112): 0 = float (default), 1 = double - the pair arithmetic
// stays fp32 (STORE = float), every "+=" into a target value and the downward interpolation run in fp64
int onb_set_accum(onb_context* c, int accum_double) {
    if (!c) return ONB_ERR_ARG;
    if (accum_double && c->legacy) { c->err = "ACCUM = double needs barycentric equivalents (-o=<order>)"; return ONB_ERR_UNSUPPORTED; }
    c->accum64 = accum_double != 0;
    return ONB_OK;
}
int onb_set_memory_mode(onb_context* c, int mode) {
    if (!c || (mode != ONB_MEM_NORMAL && mode != ONB_MEM_LEAN)) return ONB_ERR_ARG;
    c->mem_mode = mode; return ONB_OK;
}

// ---------------------------------------------------------------------------------------------
// inputs
// ---------------------------------------------------------------------------------------------
// xs[PD] / ss[SD]: one pointer per plane (the planar C-ABI layout is the special case xs[d] = x + d*n)
static int set_parts(onb_context* c, int which, uint64_t n, const float* const* xs, const float* r, const float* const* ss) {
    if (n == 0 || n >= 0xfffff000ull) { c->err = "particle count out of range for one GPU"; return ONB_ERR_ARG; }
    ONB_CUDA(cudaSetDevice(c->device));
    if (which == 0) { int jrc = onb_dist_join_source_planes(c, c->stream); if (jrc) return jrc; }     // a plane gather of the previous step may still be in flight
    DParts& p = c->parts[which];
    const uint64_t want_key = (which == 1 && c->mem_mode == ONB_MEM_LEAN && c->shard_n > 1) ? c->plan_key_for(n) : 0ull;
    if (p.n != n || p.unpacked_released || p.sparse_key != want_key || (which == 1 && (p.ud[0] != nullptr) != c->accum64)) { onb_free_parts(c, p); int rc = onb_alloc_parts(c, p, (uint32_t)n, which == 0); if (rc) return rc; }
    if (p.gidx) { p.gidx_spare = p.gidx; p.gidx = nullptr; }      // no cudaFree/cudaMalloc per step: the next build takes it back
    const size_t bytes = (size_t)n * sizeof(float);
    bool async = false;
    if (c->async_inputs) {      // only buffers the copy engine reads directly can be left in flight
        cudaPointerAttributes at;
        async = true;
        std::vector<const float*> ptrs; ptrs.push_back(r);
        for (int d = 0; d < c->PD; ++d) ptrs.push_back(xs[d]);
        if (which == 0) for (int d = 0; d < c->SD; ++d) ptrs.push_back(ss[d]);
        for (const float* q : ptrs) {
            if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); async = false; break; }
            if (at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) { async = false; break; }
        }
    }
    cudaStream_t st = c->stream;
    if (async && which == 1) {
        // targets travel on the second stream, behind whatever the context stream has enqueued so far (the source copy)
        { int jrc = onb_join_copies(c); if (jrc) return jrc; }
        if (!c->ev_copy) { ONB_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming)); ONB_CUDA(cudaEventCreateWithFlags(&c->ev_tgt_ready, cudaEventDisableTiming)); }
        ONB_CUDA(cudaEventRecord(c->ev_copy, c->stream));
        ONB_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_copy, 0));
        st = c->stream2;
    }
    if (c->comm && c->shard_n > 1 && c->sliced_inputs) {
        // every context is handed the same arrays: each pulls 1/nranks of every plane over its own PCIe link and the slices
        // are replicated over NVLink (one grouped in-place all-gather), so every input byte crosses PCIe once in total
        const size_t chunk = ((n + (size_t)c->shard_n - 1) / (size_t)c->shard_n + 31) & ~(size_t)31;
        const size_t lo = std::min<size_t>(n, chunk * (size_t)c->shard_rank), hi = std::min<size_t>(n, lo + chunk);
        if ((size_t)p.cap < chunk * (size_t)c->shard_n) { c->err = "set: planes too short for the sliced input gather"; return ONB_ERR_ARG; }
        std::vector<void*> bufs; std::vector<size_t> chunks;
        auto slice = [&](float* dst, const float* src) -> int {
            if (hi > lo) ONB_CUDA(cudaMemcpyAsync(dst + lo, src + lo, (hi - lo) * sizeof(float), cudaMemcpyDefault, st));
            bufs.push_back(dst); chunks.push_back(chunk * sizeof(float));
            return ONB_OK;
        };
        for (int d = 0; d < c->PD; ++d) { int rc = slice(p.x[d], xs[d]); if (rc) return rc; }
        { int rc = slice(p.r, r); if (rc) return rc; }
        if (which == 0) for (int d = 0; d < c->SD; ++d) { int rc = slice(p.s[d], ss[d]); if (rc) return rc; }
        cudaStream_t sc = onb_comm_stream(c);
        cudaEvent_t e_in = onb_cached_event(c, 390 + which), e_out = onb_cached_event(c, 392 + which);
        ONB_CUDA(cudaEventRecord(e_in, st)); ONB_CUDA(cudaStreamWaitEvent(sc, e_in, 0));
        { int rc = onb_comm_allgather(c, bufs, chunks); if (rc) return rc; }
        ONB_CUDA(cudaEventRecord(e_out, sc)); ONB_CUDA(cudaStreamWaitEvent(st, e_out, 0));
    } else {
        for (int d = 0; d < c->PD; ++d) ONB_CUDA(cudaMemcpyAsync(p.x[d], xs[d], bytes, cudaMemcpyDefault, st));
        ONB_CUDA(cudaMemcpyAsync(p.r, r, bytes, cudaMemcpyDefault, st));
        if (which == 0) for (int d = 0; d < c->SD; ++d) ONB_CUDA(cudaMemcpyAsync(p.s[d], ss[d], bytes, cudaMemcpyDefault, st));
    }
    if (!async) ONB_CUDA(cudaStreamSynchronize(st));
    else if (which == 1) { ONB_CUDA(cudaEventRecord(c->ev_tgt_ready, c->stream2)); c->tgt_copy_pending = true; }
    p.packed_valid = false;
    c->trees[which].built = false;
    return ONB_OK;
}
int onb_set_async_inputs(onb_context* c, int on) { if (!c) return ONB_ERR_ARG; c->async_inputs = on != 0; return ONB_OK; }
int onb_set_sliced_inputs(onb_context* c, int on) { if (!c) return ONB_ERR_ARG; c->sliced_inputs = on != 0; return ONB_OK; }
int onb_set_sources(onb_context* c, uint64_t n, const float* x, const float* r, const float* s) {
    const float* xs[ONB_MAX_PD]; const float* ss[ONB_MAX_SD];
    for (int d = 0; d < ONB_MAX_PD; ++d) xs[d] = x + (size_t)d * n;
    for (int d = 0; d < ONB_MAX_SD; ++d) ss[d] = s + (size_t)d * n;
    return set_parts(c, 0, n, xs, r, ss);
}
int onb_set_targets(onb_context* c, uint64_t n, const float* x, const float* r) {
    const float* xs[ONB_MAX_PD];
    for (int d = 0; d < ONB_MAX_PD; ++d) xs[d] = x + (size_t)d * n;
    return set_parts(c, 1, n, xs, r, nullptr);
}
int onb_set_sources_planes(onb_context* c, uint64_t n, const float* const* x, const float* r, const float* const* s) { return set_parts(c, 0, n, x, r, s); }
int onb_set_targets_planes(onb_context* c, uint64_t n, const float* const* x, const float* r) { return set_parts(c, 1, n, x, r, nullptr); }

// Parts::random_in_cube(std::mt19937) Parts.hpp:99-109 and wave_strengths :169-176, on the host like the reference
int onb_driver_inputs(int physics, uint64_t n, int strength_mode, float* x, float* r, float* s) {
    if (physics < 0 || physics > 4 || n == 0) return ONB_ERR_ARG;
    static const int PDs[5] = {3, 3, 3, 2, 2}, SDs[5] = {1, 3, 3, 1, 1};
    const int PD = PDs[physics], SD = SDs[physics];
    std::mt19937 eng(12345);
    std::uniform_real_distribution<float> dist(-1.0, 1.0);
    for (int d = 0; d < PD; ++d) for (uint64_t i = 0; i < n; ++i) x[(size_t)d * n + i] = dist(eng);
    const float factor = (float)(1.0 / (float)n);
    if (s) for (int d = 0; d < SD; ++d) for (uint64_t i = 0; i < n; ++i) s[(size_t)d * n + i] = dist(eng) * factor;
    const float rad = (float)std::pow((float)n, -1.0 / (float)PD);
    for (uint64_t i = 0; i < n; ++i) r[i] = rad;
    if (s && strength_mode == 1)
        for (uint64_t i = 0; i < n; ++i) for (int d = 0; d < SD; ++d)
            s[(size_t)d * n + i] = (float)((double)factor * std::cos((d + 0.7) * 10.0 * (double)x[(size_t)d * n + i]));
    return ONB_OK;
}

// ---------------------------------------------------------------------------------------------
// phases
// ---------------------------------------------------------------------------------------------
int onb_make_tree_range(onb_context* c, int which, uint64_t lo, uint64_t hi) {
    onb_scratch_reset(c);
    if (which < 0 || which > 1 || c->parts[which].n == 0) { c->err = "make_tree: set the particles first"; return ONB_ERR_ARG; }
    ONB_CUDA(cudaSetDevice(c->device));
    int rc = onb_alloc_tree(c, c->trees[which], c->parts[which].n, c->block);
    if (rc) return rc;
    PhaseTimer tm(c, "tree");
    rc = onb_tree_build(c, c->parts[which], c->trees[which], (uint32_t)lo, (uint32_t)std::min<uint64_t>(hi, c->parts[which].n));
    tm.stop();
    return rc;
}
int onb_make_tree(onb_context* c, int which) {
    if (c->comm && c->shard_n > 1) {
        onb_scratch_reset(c);
        if (which < 0 || which > 1 || c->parts[which].n == 0) { c->err = "make_tree: set the particles first"; return ONB_ERR_ARG; }
        ONB_CUDA(cudaSetDevice(c->device));
        return onb_dist_make_trees(c, which);
    }
    return onb_make_tree_range(c, which, 0, ~0ull);
}

// Both trees at once: the two builds are independent and their top levels are latency bound (grid-wide barriers around
// short passes), so they are enqueued on two streams and overlap on the device.
int onb_make_trees_range(onb_context* c, uint64_t slo, uint64_t shi, uint64_t tlo, uint64_t thi) {
    // one build after the other: on request (diagnostics), and above 5e8 particles, where two concurrent builds would hold
    // 2 x 36 B per particle of scratch at once
    static const bool seq_env = std::getenv("ONB_SEQ_BUILDS") != nullptr;
    const bool dist = c->comm && c->shard_n > 1;
    const bool seq_builds = dist ? onb_dist_sequential_builds(c) : (seq_env || c->mem_mode == ONB_MEM_LEAN || std::max(c->parts[0].n, c->parts[1].n) > 500000000u);
    // a pending asynchronous target copy sits on stream2, where the target build is enqueued behind it: the source
    // build on the context stream need not wait for it (the streams are joined at the end of this call)
    const bool tgt_in_flight = c->tgt_copy_pending && !seq_builds;
    if (tgt_in_flight) c->tgt_copy_pending = false;
    onb_scratch_reset(c);
    if (c->parts[0].n == 0 || c->parts[1].n == 0) { c->err = "make_trees: set sources and targets first"; return ONB_ERR_ARG; }
    ONB_CUDA(cudaSetDevice(c->device));
    if (dist) return onb_dist_make_trees(c, -1);        // every rank its own range of both trees (the range arguments are the plan's)
    int rc = onb_alloc_tree(c, c->trees[0], c->parts[0].n, c->block);
    if (rc == ONB_OK) rc = onb_alloc_tree(c, c->trees[1], c->parts[1].n, c->block);
    if (rc) return rc;
    cudaEvent_t e0, e1, e2;
    ONB_CUDA(cudaEventCreate(&e0)); ONB_CUDA(cudaEventCreate(&e1)); ONB_CUDA(cudaEventCreate(&e2));
    ONB_CUDA(cudaEventRecord(e0, c->stream));
    ONB_CUDA(cudaStreamWaitEvent(c->stream2, e0, 0));
    c->concurrent_builds = !seq_builds;
    rc = onb_tree_build(c, c->parts[0], c->trees[0], (uint32_t)slo, (uint32_t)std::min<uint64_t>(shi, c->parts[0].n));
    if (rc == ONB_OK) {
        if (seq_builds) { c->slab_cur = 0; c->slab_off = 0; }     // same stream: the second build reuses the first one's scratch
        c->cur_stream = seq_builds ? nullptr : c->stream2; c->cur_stats_off = 8;
        rc = onb_tree_build(c, c->parts[1], c->trees[1], (uint32_t)tlo, (uint32_t)std::min<uint64_t>(thi, c->parts[1].n));
        c->cur_stream = nullptr; c->cur_stats_off = 0;
    }
    c->concurrent_builds = false;
    ONB_CUDA(cudaEventRecord(e1, c->stream2));
    ONB_CUDA(cudaStreamWaitEvent(c->stream, e1, 0));
    ONB_CUDA(cudaEventRecord(e2, c->stream));
    ONB_CUDA(cudaEventSynchronize(e2));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e2);
    c->phase_ms["tree"] = ms; c->phase_ms["trees"] = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    if (c->mem_mode == ONB_MEM_LEAN) onb_scratch_trim(c);
    return rc;
}
int onb_make_trees(onb_context* c) { return onb_make_trees_range(c, 0, ~0ull, 0, ~0ull); }
int onb_finish_tree(onb_context* c, int which) {
    onb_scratch_reset(c);
    if (which < 0 || which > 1) return ONB_ERR_ARG;
    ONB_CUDA(cudaSetDevice(c->device));
    PhaseTimer tm(c, "finish");
    int rc = onb_tree_finish_from_particles(c, c->parts[which], c->trees[which]);
    tm.stop();
    return rc;
}
int onb_set_build_range(onb_context* c, int which, uint64_t lo, uint64_t hi) {
    if (which < 0 || which > 1) return ONB_ERR_ARG;
    DParts& p = c->parts[which];
    p.build_lo = (uint32_t)std::min<uint64_t>(lo, p.n); p.build_hi = (uint32_t)std::min<uint64_t>(hi, p.n);
    return ONB_OK;
}
int onb_shard_particle_range(const onb_context* c, uint64_t n, int rank, int nranks, uint64_t* lo, uint64_t* hi) {
    return onb_shard_range_for(n, c->block, rank, nranks, lo, hi);
}
int onb_refine(onb_context* c, int which) {
    onb_scratch_reset(c);
    if (which < 0 || which > 1) return ONB_ERR_ARG;
    ONB_CUDA(cudaSetDevice(c->device));
    PhaseTimer tm(c, "refine");
    int rc = onb_tree_refine(c, c->parts[which], c->trees[which]);
    tm.stop();
    return rc;
}
int onb_upward(onb_context* c, int which) {
    onb_scratch_reset(c);
    if (which < 0 || which > 1) return ONB_ERR_ARG;
    ONB_CUDA(cudaSetDevice(c->device));
    if (c->legacy && which == 1) {
        c->err = "legacy equivalents (-o omitted) exist for sources only: the reference's calcEquivalents returns at once for targets (barneshut.hpp:953)";
        return ONB_ERR_UNSUPPORTED;
    }
    PhaseTimer tm(c, "upward");
    int rc;
    if (c->legacy) rc = onb_legacy_equivalents(c, c->parts[which], c->parts[which + 2], c->trees[which]);
    else if (which == 0 && c->comm && c->shard_n > 1) rc = onb_dist_upward_sources(c);      // own nodes, exchange, straddling nodes, packing
    else if (which == 1 && c->shard_n > 1) rc = onb_bary_upward_mode(c, c->parts[1], c->parts[3], c->trees[1], ONB_UP_NEED);   // only the nodes this shard evaluates
    else rc = onb_bary_upward(c, c->parts[which], c->parts[which + 2], c->trees[which]);
    if (rc == ONB_OK && which == 0 && !c->parts[2].packed_valid) rc = onb_pack_sources(c, c->parts[2]);
    if (rc == ONB_OK && which == 0 && !c->parts[0].packed_valid) rc = onb_pack_sources(c, c->parts[0]);
    tm.stop();
    return rc;
}
// Everything between the tree builds and the dual-tree evaluation in one call, the source side and the target side on
// two streams: [finish_tree(0)] -> upward(0) -> pack | [finish_tree(1)] -> refine(1) -> upward(1). The two chains are
// independent (ongrav3d.cpp:636-724 runs them one after the other) and each is a sequence of small, latency-bound
// launches, so they overlap almost perfectly. finish != 0 is the multi-GPU variant (node arrays completed bottom-up
// after the plane exchange; the refinement is restricted to the target build range [tgt_lo, tgt_hi)).
int onb_prepare_eval(onb_context* c, int finish, uint64_t tgt_lo, uint64_t tgt_hi) {
    onb_scratch_reset(c);
    ONB_CUDA(cudaSetDevice(c->device));
    if (c->legacy) { c->err = "prepare_eval: the dual tree needs barycentric equivalents (-o=<order>)"; return ONB_ERR_UNSUPPORTED; }
    if (!c->trees[0].built || !c->trees[1].built) { c->err = "prepare_eval: build both trees first"; return ONB_ERR_ARG; }
    cudaEvent_t e0, e1, e2;
    ONB_CUDA(cudaEventCreate(&e0)); ONB_CUDA(cudaEventCreate(&e1)); ONB_CUDA(cudaEventCreate(&e2));
    ONB_CUDA(cudaEventRecord(e0, c->stream));
    ONB_CUDA(cudaStreamWaitEvent(c->stream2, e0, 0));
    // source side on the context stream
    int rc = ONB_OK;
    const bool dist = c->comm && c->shard_n > 1;
    if (dist) finish = 0;       // a communicator's onb_make_trees has completed the node arrays from exchanged leaf records already
    if (finish) rc = onb_tree_finish_from_particles(c, c->parts[0], c->trees[0]);
    if (rc == ONB_OK) rc = dist ? onb_dist_upward_sources(c) : onb_bary_upward(c, c->parts[0], c->parts[2], c->trees[0]);
    if (rc == ONB_OK && !c->parts[2].packed_valid) rc = onb_pack_sources(c, c->parts[2]);
    if (rc == ONB_OK && !c->parts[0].packed_valid) rc = onb_pack_sources(c, c->parts[0]);
    // target side on the second stream
    if (rc == ONB_OK) {
        c->cur_stream = c->stream2;
        if (finish) {
            rc = onb_tree_finish_from_particles(c, c->parts[1], c->trees[1]);
            DParts& p = c->parts[1];
            p.build_lo = (uint32_t)std::min<uint64_t>(tgt_lo, p.n); p.build_hi = (uint32_t)std::min<uint64_t>(tgt_hi, p.n);
        }
        if (rc == ONB_OK) rc = onb_tree_refine(c, c->parts[1], c->trees[1], false);
        if (rc == ONB_OK) rc = onb_bary_upward_mode(c, c->parts[1], c->parts[3], c->trees[1], c->shard_n > 1 ? ONB_UP_NEED : ONB_UP_ALL);
        c->cur_stream = nullptr;
    }
    ONB_CUDA(cudaEventRecord(e1, c->stream2));
    ONB_CUDA(cudaStreamWaitEvent(c->stream, e1, 0));
    ONB_CUDA(cudaEventRecord(e2, c->stream));
    ONB_CUDA(cudaEventSynchronize(e2));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e2);
    c->phase_ms["prepare"] = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    if (rc == ONB_OK) rc = onb_check_flag(c, "refine (introsort depth limit: libstdc++ heapsort fallback is not restated)");
    if (dist) onb_dist_record_exchange_times(c);
    if (rc == ONB_OK) rc = lean_after_prepare(c);
    return rc;
}
int onb_zero_vels(onb_context* c) {
    ONB_CUDA(cudaSetDevice(c->device));
    DParts& t = c->parts[1];
    for (int d = 0; d < c->OD; ++d) if (t.u[d]) { int rc = onb_memset_plane(c, t.u[d], (size_t)t.cap, c->stream); if (rc) return rc; }
    for (int d = 0; d < c->OD; ++d) if (t.ud[d]) ONB_CUDA(cudaMemsetAsync(t.ud[d], 0, (size_t)t.cap * sizeof(double), c->stream));
    return ONB_OK;
}
int onb_naive(onb_context* c, uint64_t tskip, float* flops) {
    onb_scratch_reset(c);
    ONB_CUDA(cudaSetDevice(c->device));
    if (c->parts[0].n == 0 || c->parts[1].n == 0) { c->err = "naive: set sources and targets first"; return ONB_ERR_ARG; }
    if (tskip < 1) tskip = 1;
    PhaseTimer tm(c, "eval");
    int rc = onb_p2p_direct(c, tskip);
    tm.stop();
    c->phase_ms["p2p"] = c->phase_ms["eval"];
    if (flops) *flops = (float)(c->parts[1].n / tskip) * (float)c->parts[0].n * (float)c->flops_per_pair;   // barneshut.hpp:52
    return rc;
}
static int need_trees(onb_context* c, bool target_tree, bool eq_targets) {
    if (!c->trees[0].built || c->parts[2].n == 0) { c->err = "build the source tree and run the upward pass first"; return ONB_ERR_ARG; }
    if (target_tree && !c->trees[1].built) { c->err = "build the target tree first"; return ONB_ERR_ARG; }
    if (eq_targets && c->parts[3].n == 0) { c->err = "run the target upward pass first"; return ONB_ERR_ARG; }
    return ONB_OK;
}
int onb_treecode3(onb_context* c, float theta, float* flops) {
    onb_scratch_reset(c);
    ONB_CUDA(cudaSetDevice(c->device));
    int rc = need_trees(c, true, false); if (rc) return rc;
    PhaseTimer te(c, "eval");
    WorkList wl;
    { PhaseTimer tl(c, "lists"); rc = onb_lists_boxwise(c, theta, wl); tl.stop(); }
    if (rc == ONB_OK) { PhaseTimer tp(c, "p2p"); rc = onb_p2p_lists(c, wl, 1, 1, true); tp.stop(); }
    onb_free_worklist(c, wl);
    te.stop();
    if (flops) *flops = (float)c->flops_per_pair * (float)c->block *
                        ((float)c->stats[0] * (float)c->block + (float)c->stats[1] * (float)(c->legacy ? c->root_epnum : (uint32_t)c->num_eqps));        // :335-336
    return rc;
}
int onb_fastsumm(onb_context* c, float theta) {
    onb_scratch_reset(c);
    ONB_CUDA(cudaSetDevice(c->device));
    if (!c->has_fastsumm) { c->err = "this physics has no dual-tree method in the reference (onvortgrad3d.cpp:264)"; return ONB_ERR_UNSUPPORTED; }
    if (c->legacy) {
        c->err = "the dual tree needs barycentric equivalent targets (-o=<order>): with -o omitted the reference builds no target equivalents at all (barneshut.hpp:953)";
        return ONB_ERR_UNSUPPORTED;
    }
    int rc = need_trees(c, true, true); if (rc) return rc;
    PhaseTimer te(c, "eval");
    rc = onb_run_fastsumm(c, theta);
    te.stop();
    return rc;
}
int onb_treecode2(onb_context* c, float theta, float* flops) {
    onb_scratch_reset(c);
    ONB_CUDA(cudaSetDevice(c->device));
    int rc = need_trees(c, false, false); if (rc) return rc;
    PhaseTimer te(c, "eval");
    rc = onb_run_treecode2(c, theta, 2);
    te.stop();
    if (flops) *flops = (float)c->flops_per_pair * ((float)c->stats[0] * (float)c->block + (float)c->stats[1] * (float)(c->legacy ? c->root_epnum : (uint32_t)c->num_eqps));   // :220-221
    return rc;
}
int onb_treecode1(onb_context* c, float theta, float* flops) {
    onb_scratch_reset(c);
    ONB_CUDA(cudaSetDevice(c->device));
    if (!c->trees[0].built) { c->err = "build the source tree first"; return ONB_ERR_ARG; }
    PhaseTimer te(c, "eval");
    int rc = onb_run_treecode2(c, theta, 1);
    te.stop();
    if (flops) *flops = (float)c->flops_per_pair * ((float)c->stats[1] + (float)c->stats[0] * (float)c->block);   // :131
    return rc;
}

// ---------------------------------------------------------------------------------------------
// outputs
// ---------------------------------------------------------------------------------------------
uint64_t onb_count(const onb_context* c, int which) { return (which >= 0 && which < 4) ? c->parts[which].n : 0; }

int onb_get_parts(onb_context* c, int which, float* x, float* r, float* s, float* u, uint64_t* gidx) {
    if (which < 0 || which > 3) return ONB_ERR_ARG;
    ONB_CUDA(cudaSetDevice(c->device));
    DParts& p = c->parts[which];
    const size_t n = p.n;
    if (n == 0) return ONB_OK;
    { int jrc = onb_join_copies(c); if (jrc) return jrc; }
    if (which == 0) { int jrc = onb_dist_join_source_planes(c, c->stream); if (jrc) return jrc; }
    if ((x || r || (s && p.are_sources)) && p.unpacked_released) { c->err = "get_parts: the source planes were released after packing (lean memory mode)"; return ONB_ERR_ARG; }
    // sparse planes (lean memory mode) are copied where they are backed by memory; the rest of the caller's array is left alone
    if (x) for (int d = 0; d < c->PD; ++d) ONB_CUDA(onb_copy_plane_to_host(c, x + d * n, p.x[d], 0, n, c->stream));
    if (r) ONB_CUDA(onb_copy_plane_to_host(c, r, p.r, 0, n, c->stream));
    if (s && p.are_sources) for (int d = 0; d < c->SD; ++d) ONB_CUDA(onb_copy_plane_to_host(c, s + d * n, p.s[d], 0, n, c->stream));
    if (u && !p.are_sources && p.ud[0]) { int rrc = onb_round_outputs(c, p); if (rrc) return rrc; }      // ACCUM = double: u = (float) ud
    if (u && !p.are_sources) for (int d = 0; d < c->OD; ++d) ONB_CUDA(onb_copy_plane_to_host(c, u + d * n, p.u[d], 0, n, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    if (gidx && p.gidx) {
        std::vector<uint32_t> tmp(n);
        ONB_CUDA(cudaMemcpy(tmp.data(), p.gidx, n * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) gidx[i] = tmp[i];
    }
    return ONB_OK;
}
// ACCUM = double: the outputs of targets (which = 1) or equivalent target points (which = 3) in full precision, u is [OD][n]
int onb_get_results_f64(onb_context* c, int which, double* u) {
    if (which != 1 && which != 3) return ONB_ERR_ARG;
    ONB_CUDA(cudaSetDevice(c->device));
    DParts& p = c->parts[which];
    if (!p.ud[0]) { c->err = "get_results_f64: ACCUM = double is not selected (onb_set_accum)"; return ONB_ERR_ARG; }
    for (int d = 0; d < c->OD; ++d) ONB_CUDA(cudaMemcpyAsync(u + (size_t)d * p.n, p.ud[d], (size_t)p.n * sizeof(double), cudaMemcpyDefault, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    return ONB_OK;
}
// the output planes of this context's shard only: u is [OD][n], elements [lo,hi) of every plane are written (the multi-GPU
// drivers let every rank fill its own part of one shared host array)
int onb_get_shard_results(onb_context* c, float* u, uint64_t plane_stride, uint64_t* lo_out, uint64_t* hi_out) {
    ONB_CUDA(cudaSetDevice(c->device));
    DParts& p = c->parts[1];
    uint32_t lo, hi; onb_shard_range(c, &lo, &hi);
    if (lo_out) *lo_out = lo; if (hi_out) *hi_out = hi;
    if (!u || hi <= lo) return ONB_OK;
    if (p.ud[0]) { int rrc = onb_round_outputs(c, p); if (rrc) return rrc; }
    const size_t stride = plane_stride ? (size_t)plane_stride : (size_t)p.n;
    for (int d = 0; d < c->OD; ++d) ONB_CUDA(cudaMemcpyAsync(u + (size_t)d * stride + lo, p.u[d] + lo, (size_t)(hi - lo) * sizeof(float), cudaMemcpyDefault, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    return ONB_OK;
}

// tree order -> the caller's order for ALL output planes in one kernel: out[d][gidx[i]] (+)= u[d][i] for the targets of
// this context's range. gidx is a permutation, so the += needs no atomics.
struct UnsortArgs { const float* u[ONB_MAX_OD]; float* out[ONB_MAX_OD]; const uint32_t* g; uint32_t lo, hi; int OD, add; };
__global__ void k_unsort_planes(const UnsortArgs a) {
    const uint32_t i = a.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.hi) return;
    const uint32_t j = a.g[i];
    if (a.add) { for (int d = 0; d < a.OD; ++d) a.out[d][j] += a.u[d][i]; }
    else       { for (int d = 0; d < a.OD; ++d) a.out[d][j] = a.u[d][i]; }
}

// results of [lo,hi) (what this context evaluated: its shard, inside the range its tree build ordered) added to the caller's
// planes. Device pointers: one kernel, in place. Host pointers: one scatter kernel into a zeroed device staging area, then
// plane by plane a bulk copy into a pinned double buffer that overlaps the host's += of the previous plane.
static int add_results(onb_context* c, float* const* out) {
    ONB_CUDA(cudaSetDevice(c->device));
    DParts& p = c->parts[1];
    const size_t n = p.n;
    if (!p.gidx) { c->err = "targets have no tree order yet"; return ONB_ERR_ARG; }
    { int jrc = onb_join_copies(c); if (jrc) return jrc; }
    onb_scratch_reset(c);
    uint32_t lo, hi; onb_shard_range(c, &lo, &hi);
    lo = std::max(lo, p.build_lo); hi = std::min(hi, p.build_hi);      // a range-restricted build ordered (and indexed) only its own range
    if (hi <= lo) return ONB_OK;
    if (p.ud[0]) { int rrc = onb_round_outputs(c, p); if (rrc) return rrc; }
    bool all_device = true;
    for (int d = 0; d < c->OD; ++d) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out[d]) != cudaSuccess) { cudaGetLastError(); all_device = false; break; }
        if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) { all_device = false; break; }
    }
    UnsortArgs a; a.g = p.gidx; a.lo = lo; a.hi = hi; a.OD = c->OD;
    for (int d = 0; d < ONB_MAX_OD; ++d) { a.u[d] = p.u[d]; a.out[d] = nullptr; }
    const unsigned grid = (unsigned)((hi - lo + 255) / 256);
    if (all_device) {
        for (int d = 0; d < c->OD; ++d) a.out[d] = out[d];
        a.add = 1;
        k_unsort_planes<<<grid, 256, 0, c->stream>>>(a); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        ONB_CUDA(cudaStreamSynchronize(c->stream));
        return ONB_OK;
    }
    float* tmp = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&tmp, (size_t)c->OD * n * sizeof(float)));
    const bool partial = (size_t)(hi - lo) < n;
    if (partial) ONB_CUDA(cudaMemsetAsync(tmp, 0, (size_t)c->OD * n * sizeof(float), c->stream));
    for (int d = 0; d < c->OD; ++d) a.out[d] = tmp + (size_t)d * n;
    a.add = 0;
    k_unsort_planes<<<grid, 256, 0, c->stream>>>(a); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    if (c->h_stage_cap < 2 * n) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->h_stage_cap = 0;
        ONB_CUDA(cudaMallocHost((void**)&c->h_stage, 2 * n * sizeof(float)));
        c->h_stage_cap = 2 * n;
    }
    if (!c->ev_stage[0]) { ONB_CUDA(cudaEventCreateWithFlags(&c->ev_stage[0], cudaEventDisableTiming)); ONB_CUDA(cudaEventCreateWithFlags(&c->ev_stage[1], cudaEventDisableTiming)); }
    auto issue = [&](int d) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(c->h_stage + (size_t)(d & 1) * n, tmp + (size_t)d * n, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
        return e != cudaSuccess ? e : cudaEventRecord(c->ev_stage[d & 1], c->stream);
    };
    ONB_CUDA(issue(0));
    for (int d = 0; d < c->OD; ++d) {
        ONB_CUDA(cudaEventSynchronize(c->ev_stage[d & 1]));
        if (d + 1 < c->OD) ONB_CUDA(issue(d + 1));                     // the other half of the double buffer is free: plane d-1 was consumed
        const float* __restrict__ h = c->h_stage + (size_t)(d & 1) * n;
        float* __restrict__ ud = out[d];
        for (size_t i = 0; i < n; ++i) ud[i] += h[i];                                     // interface3dvortgrads.cpp:384-395
    }
    return ONB_OK;
}

int onb_add_results_original_order(onb_context* c, float* u) {
    float* planes[ONB_MAX_OD];
    for (int d = 0; d < c->OD; ++d) planes[d] = u + (size_t)d * c->parts[1].n;
    return add_results(c, planes);
}
int onb_add_results_planes(onb_context* c, float* const* out) { return add_results(c, out); }

int onb_tree_shape(const onb_context* c, int which, int* levels, int* numnodes) {
    if (which < 0 || which > 1) return ONB_ERR_ARG;
    *levels = c->trees[which].levels; *numnodes = c->trees[which].numnodes; return ONB_OK;
}

int onb_get_tree(onb_context* c, int which, float* x, float* nc, float* ns, float* nr, float* pr, float* s,
                 uint64_t* ioffset, uint64_t* num, uint64_t* epoffset, uint64_t* epnum) {
    if (which < 0 || which > 1) return ONB_ERR_ARG;
    ONB_CUDA(cudaSetDevice(c->device));
    DTree& t = c->trees[which];
    const size_t n = t.numnodes, fb = n * sizeof(float);
    if (n == 0) return ONB_OK;
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    for (int d = 0; d < c->PD; ++d) {
        if (x)  ONB_CUDA(cudaMemcpy(x + d * n,  t.x[d],  fb, cudaMemcpyDeviceToHost));
        if (nc) ONB_CUDA(cudaMemcpy(nc + d * n, t.nc[d], fb, cudaMemcpyDeviceToHost));
        if (ns) ONB_CUDA(cudaMemcpy(ns + d * n, t.ns[d], fb, cudaMemcpyDeviceToHost));
    }
    if (nr) ONB_CUDA(cudaMemcpy(nr, t.nr, fb, cudaMemcpyDeviceToHost));
    if (pr) ONB_CUDA(cudaMemcpy(pr, t.pr, fb, cudaMemcpyDeviceToHost));
    if (s) for (int d = 0; d < c->SD; ++d) ONB_CUDA(cudaMemcpy(s + d * n, t.s[d], fb, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> io(n), nm(n);
    ONB_CUDA(cudaMemcpy(io.data(), t.ioffset, n * 4, cudaMemcpyDeviceToHost));
    ONB_CUDA(cudaMemcpy(nm.data(), t.num, n * 4, cudaMemcpyDeviceToHost));
    const bool have_eq = c->parts[which + 2].n > 0;
    std::vector<uint32_t> en;
    if (c->legacy && which == 0 && have_eq && c->d_epnum && epnum) { en.resize(n); ONB_CUDA(cudaMemcpy(en.data(), c->d_epnum, n * 4, cudaMemcpyDeviceToHost)); }
    for (size_t i = 0; i < n; ++i) {
        if (ioffset) ioffset[i] = io[i];
        if (num) num[i] = nm[i];
        const bool nonleaf = have_eq && nm[i] > (uint32_t)c->block;
        if (epoffset) epoffset[i] = nonleaf ? (uint64_t)i * c->ebs : 0;                   // BarycentricLagrange.hpp:289
        if (epnum) epnum[i] = nonleaf ? (en.empty() ? (uint64_t)c->num_eqps : (uint64_t)en[i]) : 0;
    }
    return ONB_OK;
}

int onb_device_memory(onb_context* c, uint64_t* used, uint64_t* total) {
    ONB_CUDA(cudaSetDevice(c->device));
    size_t f = 0, t = 0;
    ONB_CUDA(cudaMemGetInfo(&f, &t));
    if (used) *used = t - f; if (total) *total = t;
    return ONB_OK;
}

int onb_get_stats(const onb_context* c, uint64_t out[9]) { for (int i = 0; i < 9; ++i) out[i] = c->stats[i]; return ONB_OK; }

double onb_phase_ms(const onb_context* c, const char* name) {
    auto it = c->phase_ms.find(name);
    return it == c->phase_ms.end() ? -1.0 : it->second;
}
uint64_t onb_last_pairs(const onb_context* c) { return c->last_pairs; }

// step timer: two CUDA events on the context's stream bracket everything issued in between (host gaps included)
int onb_timer_start(onb_context* c) {
    ONB_CUDA(cudaSetDevice(c->device));
    if (!c->ev_t0) { ONB_CUDA(cudaEventCreate(&c->ev_t0)); ONB_CUDA(cudaEventCreate(&c->ev_t1)); }
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    ONB_CUDA(cudaEventRecord(c->ev_t0, c->stream));
    return ONB_OK;
}
double onb_timer_stop_ms(onb_context* c) {
    if (!c->ev_t0) return -1.0;
    cudaSetDevice(c->device);
    if (cudaEventRecord(c->ev_t1, c->stream) != cudaSuccess || cudaEventSynchronize(c->ev_t1) != cudaSuccess) return -1.0;
    float ms = 0.f; cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
    return ms;
}
uint64_t onb_launch_count(const onb_context* c) { return c->launches; }

void* onb_device_ptr(onb_context* c, int which, int field) {
    if (which < 0 || which > 3) return nullptr;
    cudaSetDevice(c->device);
    onb_join_copies(c);
    DParts& p = c->parts[which];
    if (field >= 0 && field < 3) return p.x[field];
    if (field == 3) return p.r;
    if (field >= 4 && field < 7) return p.s[field - 4];
    if (field >= 7 && field < 7 + ONB_MAX_OD) return p.u[field - 7];
    return nullptr;
}

int onb_load_tree(onb_context* c, int which, int levels, const float* x, const float* nc, const float* ns,
                  const float* nr, const float* pr, const float* s, const uint64_t* ioffset, const uint64_t* num) {
    if (which < 0 || which > 1 || c->parts[which].n == 0) { c->err = "load_tree: set the particles first"; return ONB_ERR_ARG; }
    ONB_CUDA(cudaSetDevice(c->device));
    DTree& t = c->trees[which];
    int rc = onb_alloc_tree(c, t, c->parts[which].n, c->block);
    if (rc) return rc;
    if (t.levels != levels) { c->err = "load_tree: level count does not match Tree.hpp sizing"; return ONB_ERR_ARG; }
    ONB_CUDA(cudaStreamSynchronize(c->stream));     // the allocation's zero-fills are stream ordered; the copies below are not
    const size_t n = t.numnodes, fb = n * sizeof(float);
    for (int d = 0; d < c->PD; ++d) {
        ONB_CUDA(cudaMemcpy(t.x[d], x + d * n, fb, cudaMemcpyHostToDevice));
        ONB_CUDA(cudaMemcpy(t.nc[d], nc + d * n, fb, cudaMemcpyHostToDevice));
        ONB_CUDA(cudaMemcpy(t.ns[d], ns + d * n, fb, cudaMemcpyHostToDevice));
    }
    ONB_CUDA(cudaMemcpy(t.nr, nr, fb, cudaMemcpyHostToDevice));
    ONB_CUDA(cudaMemcpy(t.pr, pr, fb, cudaMemcpyHostToDevice));
    if (s) for (int d = 0; d < c->SD; ++d) ONB_CUDA(cudaMemcpy(t.s[d], s + d * n, fb, cudaMemcpyHostToDevice));
    std::vector<uint32_t> io(n), nm(n);
    for (size_t i = 0; i < n; ++i) { io[i] = (uint32_t)ioffset[i]; nm[i] = (uint32_t)num[i]; }
    ONB_CUDA(cudaMemcpy(t.ioffset, io.data(), n * 4, cudaMemcpyHostToDevice));
    ONB_CUDA(cudaMemcpy(t.num, nm.data(), n * 4, cudaMemcpyHostToDevice));
    if (which == 1) {   // targets loaded in tree order: original index = position
        DParts& p = c->parts[1];
        if (!p.gidx && p.gidx_spare) { p.gidx = p.gidx_spare; p.gidx_spare = nullptr; }
        if (!p.gidx) ONB_CUDA(onb_pmalloc(c, (void**)&p.gidx, (size_t)p.n * 4));
        std::vector<uint32_t> id(p.n); for (uint32_t i = 0; i < p.n; ++i) id[i] = i;
        ONB_CUDA(cudaMemcpy(p.gidx, id.data(), (size_t)p.n * 4, cudaMemcpyHostToDevice));
    }
    c->parts[which].build_lo = 0; c->parts[which].build_hi = c->parts[which].n;
    t.built = true;
    return ONB_OK;
}

}  // extern "C"
